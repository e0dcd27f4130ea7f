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

    def _stage(self, feats_coarse, feats_fine, coarse, fine, normalize, n_chunks=6):
        """K0 of both feature stacks.  Device tensors: two launches, returns a no-op.  Host (pinned) tensors: chunked
        copies on a side stream + K0 per chunk on a second one; returns ``landed(t)``, which makes the current stream
        wait for the chunk frame t is in."""
        if feats_coarse.is_cuda and feats_fine.is_cuda:
            coarse.load_frames(feats_coarse.float(), 0, normalize=normalize)
            fine.load_frames(feats_fine.float(), 0, normalize=normalize)
            return lambda t: None
        dev = coarse.buf.device
        T = feats_coarse.shape[0]
        cur = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream, self._k0_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        cp, k0 = self._copy_stream, self._k0_stream
        step = -(-T // n_chunks)
        # two persistent staging buffers per stack (kept across calls): a copy waits only for the K0 that last read its buffer
        key = (step,) + tuple(feats_coarse.shape[1:]) + tuple(feats_fine.shape[1:]) + (dev.index,)
        if getattr(self, "_stage_key", None) != key:
            self._stage_c = [torch.empty((step,) + tuple(feats_coarse.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
            self._stage_f = [torch.empty((step,) + tuple(feats_fine.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
            self._stage_free, self._stage_key = [None, None], key
            cp.wait_stream(cur)                           # fresh memory: earlier kernels of this stream may still use it
        k0.wait_stream(cur)                               # the banks may still be read by the previous call's kernels
        done = []
        for i, a in enumerate(range(0, T, step)):
            b = min(T, a + step)
            fc, ff = self._stage_c[i % 2][:b - a], self._stage_f[i % 2][:b - a]
            with torch.cuda.stream(cp):
                if self._stage_free[i % 2] is not None:
                    cp.wait_event(self._stage_free[i % 2])
                fc.copy_(feats_coarse[a:b], non_blocking=True)
                ff.copy_(feats_fine[a:b], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(cp)
            with torch.cuda.stream(k0):
                k0.wait_event(ready)
                coarse.load_frames(fc, a, normalize=normalize)
                fine.load_frames(ff, a, normalize=normalize)
                ev = torch.cuda.Event()
                ev.record(k0)
            self._stage_free[i % 2] = ev
            done.append((b, ev))
        state = {"i": 0}

        def landed(t):
            while state["i"] < len(done) and done[state["i"]][0] <= t:
                state["i"] += 1                      # chunks wholly before frame t: already waited for, or implied
            # frame t lies in chunk i (its end is the first one > t): wait for it (events of earlier chunks precede it
            # on the same stream)
            if state["i"] < len(done) and not state.get(state["i"]):
                cur.wait_event(done[state["i"]][1])
                state[state["i"]] = True
        return landed

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
        dev = feats_coarse.device if feats_coarse.is_cuda else torch.device("cuda", torch.cuda.current_device())
        stride_f = h // Hf
        split = cfg.get("split") or ("f16" if (engine.default_split(C) == "f16" and Cf % 4 == 0) else "tf32")
        normalize = cfg.get("with_norm", True)
        coarse = FeatureBank(T, C, Hc, Wc, dev, split=split)
        fine = FeatureBank(T, Cf, Hf, Wf, dev, split=split)
        # host features: copied in frame chunks on a side stream; the frame loop below waits for the chunk a frame is in
        # (frame t only reads frames <= t), so the host link overlaps the propagation of the frames already there
        landed = self._stage(feats_coarse, feats_fine, coarse, fine, normalize)
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
            landed(t)
            out = engine.c2f_propagate(coarse, fine, table, t - 1, labels, radius, cfg.get("radius_fine", 12),
                                       cfg["topk"], cfg["temperature"], cfg.get("mask_mode", "circle"), self.engine_id)
            # the coarse output becomes (a) the fine memory labels of the following frames, (b) this frame's coordinates
            _lib.call("fgvc_upsample_labels", _lib.ptr(out), Hc, Wc, labels.Lp, _lib.ptr(labels.buf), t, Hf, Wf,
                      _lib.stream_ptr())
            _lib.call("fgvc_labels_to_nchw", _lib.ptr(out), 0, labels.Lp, P, Hc * Wc, _lib.ptr(maps), _lib.stream_ptr())
            traj[t] = engine.heatmap_coords(maps, (h, w))
            outs.append(out)
        landed(T - 1)            # (a one-frame clip: the banks must not be released under the staging stream)
        return traj, outs
