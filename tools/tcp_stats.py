"""Prefilter engine statistics on bench.py's workload (run under gpurun): insertion rounds, overflow queue."""
import os, sys, json
os.environ["FGVC_TCP_EXP"] = str(8 | int(os.environ.get("FGVC_TCP_EXP", "0")))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fgvc_b200 import engine

dev = torch.device("cuda", 0)
feats, onehot = bench.build_inputs(dev, 1000)
W = bench.WORK
clip = engine.MaskClipPropagator(W["clip_frames"], W["channels"], *W["feat_hw"], W["objects"], W["image_hw"], bench.CFG, dev,
                                 engine_id=engine._lib.ENGINE_PREFILTER)
clip.run(feats, onehot, want_maps=False)
torch.cuda.synchronize()
ws = engine._WORKSPACES[(0, torch.cuda.current_stream().cuda_stream)]
n_jobs, n_pix = len(clip.table), W["feat_hw"][0] * W["feat_hw"][1]
off = n_jobs * clip.groups * n_pix * 4 * 10 * 8
ovf = int(ws[off:off + 4].view(torch.int32).item())
st = ws[off + 64:off + 96].view(torch.int64).tolist()
print(json.dumps(dict(queries=n_jobs * n_pix, overflow=ovf, row_scans=st[0], hot=st[1], rounds=st[2], insertions=st[3],
                      hot_frac=st[1] / max(1, st[0]), rounds_per_rowscan=st[2] / max(1, st[0]),
                      ins_per_query=st[3] / (n_jobs * n_pix))))
