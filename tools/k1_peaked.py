"""K1 on the bench geometry with PEAKED features (every pixel an own random direction, the same in all frames, plus
frame noise) instead of the random-init encoder's smooth ones: how much of K1's time is the candidate scan when the
thresholds rise at once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from fgvc_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
W = bench.WORK
T, (H, Wd), C = W["clip_frames"], W["feat_hw"], W["channels"]
g = torch.Generator(device="cpu").manual_seed(0)
base = torch.randn(1, C, H, Wd, generator=g)
for name, noise in (("peaked (noise 0.3)", 0.3), ("peaked (noise 1.0)", 1.0), ("i.i.d. random", None)):
    if noise is None:
        feats = torch.randn(T, C, H, Wd, generator=g)
    else:
        feats = base + noise * torch.randn(T, C, H, Wd, generator=g)
    feats = feats.to(dev)
    clip = engine.MaskClipPropagator(T, C, H, Wd, W["objects"], W["image_hw"], bench.CFG, dev)
    clip.bank.load_frames(feats, 0, normalize=True)
    n = len(clip.table)
    for _ in range(2):
        clip._k1(0, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        clip._k1(0, n)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: K1 {e0.elapsed_time(e1) / 5:.3f} ms  ({clip.plan})")
    del clip
