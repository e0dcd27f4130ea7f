"""Host -> device bandwidth with 1, 2 and 4 concurrent copy streams (pinned memory), idle GPU."""
import torch

dev = torch.device("cuda", 0)
n = 256 << 20
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
devb = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
for k in (1, 2, 4):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams[:k]:
            s.wait_stream(torch.cuda.current_stream())
        for r in range(4):
            for i in range(k):
                with torch.cuda.stream(streams[i]):
                    devb[i].copy_(host[i], non_blocking=True)
        for s in streams[:k]:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
    print(f"{k} stream(s): {4 * k * n / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s")
