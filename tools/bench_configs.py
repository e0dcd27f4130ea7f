"""Secondary measurements (not the driver's bench): the point-tracking configurations of
BASELINE.json through the public tracker API, features resident in HBM.

  python tools/bench_configs.py [cfg3 cfg4 cfg5s c2f]
cfg3 : TAP-Vid-DAVIS shape, 256x256 stride 2 -> 128x128, 50 frames, 256 points, r=15, precede 5
cfg4 : JHMDB shape, 320x320 stride 2 -> 160x160, 32 frames, 15 key-points, r=15, precede 5
cfg5s: TAP-Vid-Kinetics shard, 256x256 -> 128x128, 250 frames, 128 of 1024 points (1/8 shard)
c2f  : one coarse-to-fine call, coarse 32x32 + fine 128x128, T=6, L=256, radius_fine 12
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import fgvc_b200  # noqa: E402
from fgvc_b200 import _lib, synthetic as S  # noqa: E402

CFGS = {
    "cfg3": dict(hw=(256, 256), T=50, P=256, nr=30, precede=5),
    # same clip, query times uniform in [0, T/2) (TAP-Vid query_mode="first": a with_first group per query frame)
    "cfg3g": dict(hw=(256, 256), T=50, P=256, nr=30, precede=5, spread=True),
    "cfg4": dict(hw=(320, 320), T=32, P=15, nr=30, precede=5),
    "cfg5s": dict(hw=(256, 256), T=250, P=128, nr=30, precede=5),
    # BASELINE config 5 memory-length sweep (same 1/8 point shard)
    "cfg5s_p10": dict(hw=(256, 256), T=250, P=128, nr=30, precede=10),
    "cfg5s_p20": dict(hw=(256, 256), T=250, P=128, nr=30, precede=20),
    "cfg5s_p40": dict(hw=(256, 256), T=250, P=128, nr=30, precede=40),
}


def run_points(name, reps=3):
    c = CFGS[name]
    h, w = c["hw"]
    dev = torch.device("cuda")
    enc = S.davis_encoder(2, seed=0).to(dev)
    frames = S.synthetic_video(min(c["T"], 50), h, w, seed=1000).to(dev)
    feats = S.encode(enc, frames, batch=4)
    if feats.shape[0] < c["T"]:                      # long clips: tile the encoded frames
        feats = feats.repeat((c["T"] + feats.shape[0] - 1) // feats.shape[0], 1, 1, 1)[: c["T"]].contiguous()
    qp = S.query_points(c["P"], c["T"], h, w, seed=1, first_frame_only=not c.get("spread", False))
    cfg = dict(precede_frames=c["precede"], topk=10, temperature=0.07, neighbor_range=c["nr"], with_first=True,
               with_first_neighbor=True)
    trk = fgvc_b200.VanillaTracker(backbone=torch.nn.Identity(), test_cfg=cfg)
    groups = [(int(t0), qp[qp[:, 0] == t0][:, 1:].to(dev)) for t0 in sorted(set(qp[:, 0].tolist()))]
    trk.propagate_points(feats, groups, (h, w))
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        trk.propagate_points(feats, groups, (h, w))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps(dict(config=name, frames_per_s=(c["T"] - 1) / dt, ms_per_clip=dt * 1e3,
                          launches_per_clip=(_lib.launch_count() - n0) // reps, feat=list(feats.shape), P=c["P"],
                          groups=len(groups))))


def run_c2f(reps=5):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    T, C, L = 6, 256, 256
    q, k = torch.randn(1, C, 32, 32, generator=g).relu().to(dev), torch.randn(1, C, T, 32, 32, generator=g).relu().to(dev)
    qf = torch.randn(1, C, 128, 128, generator=g).relu().to(dev)
    kf = torch.randn(1, C, T, 128, 128, generator=g).relu().to(dev)
    v = torch.rand(1, L, T, 128, 128, generator=g).to(dev)
    mask = fgvc_b200.spatial_neighbor(1, 32, 32, 24, dev, torch.float32)
    f = lambda: fgvc_b200.masked_attention_efficient_c2f(q, k, qf, kf, v, mask, temperature=0.07, topk=10, radius_fine=12)
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    print(json.dumps(dict(config="c2f", ms_per_call=(time.perf_counter() - t0) / reps * 1e3)))


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["cfg3", "cfg4", "cfg5s", "c2f"]):
        if name == "c2f":
            run_c2f()
        else:
            run_points(name)
