"""e2e (host buffers in / out) of bench.py's clip for different chunk plans (run under gpurun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fgvc_b200 import engine

dev = torch.device("cuda", 0)
feats, onehot = bench.build_inputs(dev, 1000)
W = bench.WORK
T, (H, Wd) = W["clip_frames"], W["feat_hw"]
clip = engine.MaskClipPropagator(T, W["channels"], H, Wd, W["objects"], W["image_hw"], bench.CFG, dev)
fh, oh = feats.cpu().pin_memory(), onehot.cpu().pin_memory()
mh = torch.empty(tuple(clip.masks.shape), dtype=torch.uint8).pin_memory()
tiles = (-(-H // 8)) * (-(-Wd // 16))
plans = {"default": None, "whole": [(0, T - 1)]}
for cr in (0.75, 1.2):
    for lc in (0.4, 2.0):
        plans[f"cr{cr}_lc{lc}"] = engine.plan_chunks(T - 1, tiles, copy_ratio=cr, launch_cost=lc)
for n in (2, 3, 4, 6, 8):
    step = -(-(T - 1) // n)
    plans[f"even{n}"] = [(a, min(T - 1, a + step)) for a in range(0, T - 1, step)]
for name, ch in plans.items():
    for _ in range(2):
        clip.run_host(fh, oh, mh, chunks=ch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        clip.run_host(fh, oh, mh, chunks=ch)
    clip.join_host()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps(dict(plan=name, ms=round(ms, 3), fps=round((T - 1) / ms * 1e3), n_chunks=len(ch) if ch else None)))
