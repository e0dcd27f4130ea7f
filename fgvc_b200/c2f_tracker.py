"""Coarse-to-fine point tracking of a clip: the recurrent driver around ``masked_attention_efficient_c2f``
(mmpt/models/common/local_attention.py:721-880) that BASELINE config 3-(ii) names.

The reference defines the operator but no driver: its output lives on the COARSE query
grid while its memory labels are consumed at FINE resolution.  This module closes the loop in the simplest way the
reference's own tracker suggests (vanilla_tracker.py:305-412); the test harness restates the same loop step by step
around the genuine operator (tests/golden/c2f_driver.npz):

    S_fine[0]  = gaussian heat-maps of the query points at the fine stride          (vanilla_tracker.py:204-221)
    for t = 1 .. T-1:
        mem       = [0] + [max(0, t - precede) .. t-1]                                (:346-362)
        out_c     = c2f(q = Fc[t], k = Fc[mem], q_fine = Ff[t], k_fine = Ff[mem], v = S_fine[mem])   -> [L, Hc, Wc]
        S_fine[t] = bilinear(out_c -> (Hf, Wf), align_corners=False)                 the next frames' memory labels
        traj[t]   = img2coord(bilinear(out_c -> (h, w)))                              (:396-406)

Everything stays resident in HBM across frames: both feature banks are built once per clip (two K0 launches), the fine
label bank holds every frame, a frame costs one coarse K1 (k = 1 per memory frame), one window-mode K1 on the tensor
cores, the c2f tail, one up-sampling kernel and K3.
"""
import torch

from . import _lib, engine
from .engine import FeatureBank, JobTable, LabelBank


class C2FPointTracker:
    """cfg keys: precede_frames, topk, temperature, neighbor_range (coarse), radius_fine, with_first (memory),
    with_first_neighbor, with_norm, split, mask_mode."""

    def __init__(self, cfg, engine_id=_lib.ENGINE_AUTO):
        self.cfg = cfg
        self.engine_id = engine_id

    @torch.no_grad()
    def track(self, feats_coarse, feats_fine, points_xy, image_hw):
        """feats_coarse [T,C,Hc,Wc], feats_fine [T,Cf,Hf,Wf] (CUDA fp32, Hf = s * Hc); points_xy [P,2] (x, y) image
        pixels of the points at frame 0.  Returns (traj [T,P,2] float32 CUDA, out_coarse list of [Nq_c, Lp])."""
        _lib.require_cuda()
        cfg = self.cfg
        T, C, Hc, Wc = feats_coarse.shape
        Tf, Cf, Hf, Wf = feats_fine.shape
        assert T == Tf and Hf % Hc == 0 and Wf % Wc == 0 and Hf // Hc == Wf // Wc, "fine grid must be s x the coarse grid"
        h, w = image_hw
        dev = feats_coarse.device
        stride_f = h // Hf
        split = cfg.get("split") or ("f16" if (engine.default_split(C) == "f16" and Cf % 4 == 0) else "tf32")
        normalize = cfg.get("with_norm", True)
        coarse = FeatureBank(T, C, Hc, Wc, dev, split=split)
        coarse.load_frames(feats_coarse.float(), 0, normalize=normalize)
        fine = FeatureBank(T, Cf, Hf, Wf, dev, split=split)
        fine.load_frames(feats_fine.float(), 0, normalize=normalize)
        P = points_xy.shape[0]
        pts = points_xy.to(device=dev, dtype=torch.float32).contiguous()
        labels = LabelBank(T, P, Hf, Wf, dev)
        labels.put_gaussians(pts, 0, stride_f)
        table = JobTable()
        unmasked_first = 0 if cfg.get("with_first_neighbor", True) else 1
        for t in range(1, T):
            mem = engine.memory_frames(t, cfg["precede_frames"], cfg.get("with_first", True))
            table.add(t, mem, mem, t, unmasked=unmasked_first)
        radius = cfg["neighbor_range"] // 2
        traj = torch.zeros(T, P, 2, dtype=torch.float32, device=dev)
        traj[0] = engine.gaussian_coords(pts, (h, w))
        maps = torch.empty(P, Hc, Wc, dtype=torch.float32, device=dev)
        outs = []
        for t in range(1, T):
            out = engine.c2f_propagate(coarse, fine, table, t - 1, labels, radius, cfg.get("radius_fine", 12),
                                       cfg["topk"], cfg["temperature"], cfg.get("mask_mode", "circle"), self.engine_id)
            # the coarse output becomes (a) the fine memory labels of the following frames, (b) this frame's coordinates
            _lib.call("fgvc_upsample_labels", _lib.ptr(out), Hc, Wc, labels.Lp, _lib.ptr(labels.buf), t, Hf, Wf,
                      _lib.stream_ptr())
            _lib.call("fgvc_labels_to_nchw", _lib.ptr(out), 0, labels.Lp, P, Hc * Wc, _lib.ptr(maps), _lib.stream_ptr())
            traj[t] = engine.heatmap_coords(maps, (h, w))
            outs.append(out)
        return traj, outs
