import torch, time
x = torch.empty(421_023_600 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D pinned 421 MB: {ms:.2f} ms -> {x.numel()*4/ms/1e6:.1f} GB/s")
m = torch.empty(26_234_880, dtype=torch.uint8, device="cuda"); h = torch.empty(26_234_880, dtype=torch.uint8).pin_memory()
e0.record(); 
for _ in range(5): h.copy_(m, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"D2H 26 MB: {e0.elapsed_time(e1)/5:.2f} ms")
