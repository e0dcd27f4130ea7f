"""Feasibility probe: K1 of clip i+1 on one stream, the tail (gather chain + decode) of clip i on another.
Timing only (both use the same buffers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from fgvc_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
feats, onehot = bench.build_inputs(dev, 1000)
W = bench.WORK
T, (H, Wd) = W["clip_frames"], W["feat_hw"]
clip = engine.MaskClipPropagator(T, W["channels"], H, Wd, W["objects"], W["image_hw"], bench.CFG, dev)
clip.run(feats, onehot, want_maps=False)
torch.cuda.synchronize()
n = len(clip.table)
A, B = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=-1)


def timed(fn, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (A, B):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def serial():
    clip.bank.load_frames(feats, 0, normalize=True)
    clip._k1(0, n)
    clip._tail(0, n, False)


def overlapped():
    cur = torch.cuda.current_stream()
    A.wait_stream(cur); B.wait_stream(cur)
    with torch.cuda.stream(A):
        clip.bank.load_frames(feats, 0, normalize=True)
        clip._k1(0, n)
    with torch.cuda.stream(B):
        clip._tail(0, n, False)
    cur.wait_stream(A); cur.wait_stream(B)


def k1_only():
    clip._k1(0, n)


def tail_only():
    clip._tail(0, n, False)


for name, fn in (("serial", serial), ("overlapped", overlapped), ("k1_only", k1_only), ("tail_only", tail_only)):
    timed(fn, 3)
    print(f"{name}: {timed(fn):.3f} ms")
