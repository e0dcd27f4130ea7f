"""One pass of a secondary bench configuration through the tracker API (for ncu launch lists).
  python tools/secondary_pass.py cfg5_tapvid_kinetics [precede] [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import fgvc_b200  # noqa: E402
from fgvc_b200 import synthetic as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_tapvid_kinetics"
if name == "cfg3ii":                     # the coarse-to-fine clip driver
    import json
    print(json.dumps(bench.run_c2f_clip(torch.device("cuda", 0), 0, 1)))
    sys.exit(0)
c = bench.SECONDARY[name]
precede = int(sys.argv[2]) if len(sys.argv) > 2 else c["precede"][0]
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
h, w = c["hw"]
feats = bench._secondary_feats(c, dev, 1, seed0=2000)[0]
qp = S.query_points(c["P"], c["T"], h, w, seed=1)
groups = [(0, qp[:, 1:].to(dev))]
if os.environ.get("SPREAD_GROUPS"):      # TAP-Vid query_mode='first': query times spread over [0, T/2)
    n = int(os.environ["SPREAD_GROUPS"])
    pts = qp[:, 1:].to(dev)
    groups = [((c["T"] // 2) * g // n, pts[g::n]) for g in range(n)]
cfg = dict(precede_frames=precede, topk=10, temperature=0.07, neighbor_range=c["nr"], with_first=True,
           with_first_neighbor=True)
trk = fgvc_b200.VanillaTracker(backbone=torch.nn.Identity(), test_cfg=cfg)
for _ in range(passes):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trk.propagate_points(feats, groups, (h, w))
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} precede {precede}: {e0.elapsed_time(e1):.2f} ms")
