#!/usr/bin/env python
"""bench.py -- propagated frames/s of the label-propagation hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "davis2017_vos"): BASELINE config 2 -- VOS-style mask
propagation on a synthetic DAVIS-2017-shaped clip: 480x854 frames, random-init ResNet-18
stride-8 features (60x107x256, encoder untimed and out of scope), 64 frames per clip,
memory = first frame + 20 preceding frames, radius 12 (neighbor_range 24), top-k 10,
temperature 0.07, 11 objects.  One step = one clip = 63 propagated frames through
K0 (normalise/split) -> K1 (tcgen05 affinity + mask + top-k, one launch) -> 63 x
[K1b gather, NCHW, decode to a 480x854 uint8 mask].

value    : frames/s with the clip's features already resident in HBM.
e2e      : same through the public clip API with HOST (pinned) features in and HOST masks out.
roofline : K1 against the tensor pipe: algorithmic FLOPs 2*C*sum(in-mask pairs) per launch
           over the CUDA-event duration of the K1 launch; peak = measured bf16 dense / 2 (tf32)
           / 3 (3xTF32 issues three tensor MACs per fp32-faithful MAC).
N > 1    : one process per GPU (torchrun), one clip per rank per step (weak scaling, videos
           are independent: no data-path collective); the per-step masks are gathered with one
           NCCL all_gather inside the timed region; time = max over ranks.
--impl reference : the oracle port of the reference's op sequence (torch CPU, all host
           threads) on a bounded sample of the same workload: one full-memory frame per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORK = dict(workload="davis2017_vos", image_hw=(480, 854), stride=8, feat_hw=(60, 107), channels=256,
            clip_frames=64, precede_frames=20, neighbor_range=24, topk=10, temperature=0.07, objects=11)
CFG = dict(precede_frames=WORK["precede_frames"], topk=WORK["topk"], temperature=WORK["temperature"],
           neighbor_range=WORK["neighbor_range"], with_first=True, with_first_neighbor=True)


def in_mask_pairs(H, W, r):
    """sum over queries of in-mask keys = |{(q,k): dy^2+dx^2 < r^2}| (exact)."""
    n = 0
    for dy in range(-(r - 1), r):
        for dx in range(-(r - 1), r):
            if dy * dy + dx * dx < r * r:
                n += max(0, H - abs(dy)) * max(0, W - abs(dx))
    return n


def algorithmic_work():
    H, W = WORK["feat_hw"]
    C, T, L = WORK["channels"], WORK["clip_frames"], WORK["objects"]
    pairs = in_mask_pairs(H, W, WORK["neighbor_range"] // 2)
    entries = sum(min(t, WORK["precede_frames"]) + 1 for t in range(1, T))     # memory entries incl. duplicate
    flops = 2.0 * C * pairs * entries
    return dict(flops_per_step=flops, in_mask_pairs=pairs, mem_entries=entries)


def dense_pairs(H, W, r, split="f16"):
    """(query, key) pairs the tensor engine actually multiplies: 128-query tiles x the key boxes (16 x BH pixels)
    of their radius halo that some query of the tile can see.  Same geometry choices as the launcher
    (csrc/topk_tc16.cu: tile orientation and BH by halo cost; corner boxes skipped)."""
    reach = r - 1
    max_bh = 4 if split == "f16" else 8

    def box_cost(rows, bh):
        return -(-rows // bh) * (16 * bh + 24)

    def pick_bh(rows):
        best = max_bh
        for bh in range(max_bh - 1, 0, -1):
            if box_cost(rows, bh) < box_cost(rows, best):
                best = bh
        return best

    def halo_cost(qh, qw):
        rows, cols = min(H, qh + 2 * reach), min(W, qw + 2 * reach)
        return (-(-H // qh)) * (-(-W // qw)) * box_cost(rows, pick_bh(rows)) * (-(-cols // 16))

    QH, QW = (16, 8) if halo_cost(16, 8) < halo_cost(8, 16) else (8, 16)
    bh = pick_bh(min(H, QH + 2 * reach))
    total = 0
    for qy0 in range(0, H, QH):
        for qx0 in range(0, W, QW):
            y_lo, y_hi = max(0, qy0 - reach), min(H - 1, qy0 + QH - 1 + reach)
            x_lo, x_hi = max(0, qx0 - reach), min(W - 1, qx0 + QW - 1 + reach)
            qy1, qx1 = min(H - 1, qy0 + QH - 1), min(W - 1, qx0 + QW - 1)
            for by in range(y_lo, y_hi + 1, bh):
                for bx in range(x_lo, x_hi + 1, 16):
                    by1, bx1 = min(H - 1, by + bh - 1), min(W - 1, bx + 15)
                    dy = max(0, by - qy1, qy0 - by1)
                    dx = max(0, bx - qx1, qx0 - bx1)
                    if dy * dy + dx * dx < r * r:
                        total += 128 * 16 * bh
    return total


def k1_traffic(split):
    """DRAM bytes of one K1 launch of this workload from the committed ncu capture (or None)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))[split]
        return d["bytes"], d["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def algorithmic_bytes():
    """per propagated frame 4*[C*Nq + C*Nk*T_u] feature bytes (fp32-equivalent: every unique
    memory frame and the query frame read once) -- summed over the clip's 63 jobs; the top-k lists
    (8*k*Nq per job) are what K1 writes."""
    H, W = WORK["feat_hw"]
    C, T, k = WORK["channels"], WORK["clip_frames"], WORK["topk"]
    n = H * W
    total = 0
    for t in range(1, T):
        uniq = min(t, WORK["precede_frames"]) + (0 if t <= WORK["precede_frames"] else 1)
        total += 4 * (C * n + C * n * uniq) + 8 * k * n
    return total


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm=d.get("hbm_gbs"), src="measured")
    return dict(bf16=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def build_inputs(device, seed):
    from fgvc_b200 import synthetic as S
    h, w = WORK["image_hw"]
    T = WORK["clip_frames"]
    enc = S.davis_encoder(WORK["stride"], seed=0).to(device)
    frames = S.synthetic_video(T, h, w, seed=seed).to(device)
    feats = S.encode(enc, frames, batch=8).contiguous()
    assert tuple(feats.shape[1:]) == (WORK["channels"],) + WORK["feat_hw"], feats.shape
    seg = S.voronoi_mask(*WORK["feat_hw"], WORK["objects"], seed=seed)
    onehot = torch.nn.functional.one_hot(seg, WORK["objects"]).permute(2, 0, 1).float().contiguous().to(device)
    del enc, frames
    torch.cuda.empty_cache()
    return feats, onehot


def run_ours(args):
    import torch.distributed as dist
    from fgvc_b200 import _lib, engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    feats, onehot = build_inputs(dev, seed=1000 + rank)
    T = WORK["clip_frames"]
    H, W = WORK["feat_hw"]
    eng = _lib.ENGINE_AUTO
    clip = engine.MaskClipPropagator(T, WORK["channels"], H, W, WORK["objects"], WORK["image_hw"], CFG, dev,
                                     engine_id=eng, split=args.split)
    split = clip.bank.split
    gathered = torch.empty((world,) + tuple(clip.masks.shape), dtype=torch.uint8, device=dev) if world > 1 else None

    def step(ev=False):
        clip.run(feats, onehot, events=ev, want_maps=False)
        if world > 1:
            dist.all_gather_into_tensor(gathered, clip.masks)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(n_warm):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    k1_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    k1_ev = []
    for _ in range(args.steps):
        step(ev=True)
        k1_ev.append(clip.k1_events)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    k1_ms = [a.elapsed_time(b) for a, b in k1_ev]
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    frames_per_step = (T - 1) * world
    value = frames_per_step * args.steps / (ms / 1e3)

    if args.profile:      # under ncu: kernels only, no e2e / CPU legs, no JSON line worth keeping
        if rank == 0:
            print(json.dumps(dict(profile_run=True, ms_per_step=ms / args.steps, k1_ms=k1_ms)))
        return
    # ---- end to end: pinned host features in, host masks out, through the public clip API
    feats_host = feats.cpu().pin_memory()
    onehot_host = onehot.cpu().pin_memory()
    masks_host = torch.empty(tuple(clip.masks.shape), dtype=torch.uint8).pin_memory()

    # a stream of clips: the staging buffers are double-buffered, so the copies of clip i + 1 run under the compute
    # of clip i and K1 stays one launch per clip (engine.MaskClipPropagator.run_host)
    whole = [(0, len(clip.table))]

    def e2e_step():
        clip.run_host(feats_host, onehot_host, masks_host, chunks=whole)

    for _ in range(2):
        e2e_step()
    clip.join_host()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    clip.join_host()                      # the device->host copies of every step are inside the timed region
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_step * args.steps / (float(t.item()) / 1e3)
    h2d = feats_host.numel() * 4 + onehot_host.numel() * 4
    d2h = masks_host.numel()

    if rank == 0:
        work = algorithmic_work()
        pk = peaks()
        k1 = statistics.mean(k1_ms)
        achieved = work["flops_per_step"] / (k1 / 1e3) / 1e12
        # three tensor MACs per fp32-faithful MAC; the tf32 pipe runs at half the 16-bit rate
        peak = pk["bf16"] / 3.0 if split == "f16" else pk["bf16"] / 2.0 / 3.0
        traffic, traffic_src = k1_traffic(split)
        # dense pairs the engine multiplies: job-packed tiles for the fp16 engine (csrc/topk_tc16g.cu)
        mode_id = _lib.MASK_CIRCLE
        r = WORK["neighbor_range"] // 2
        J, aligned = (clip.plan.J, clip.plan.aligned) if split == "f16" else (1, False)
        if split == "f16":
            dense = engine.dense_pairs(clip.table, 0, len(clip.table), H, W, r, mode_id, J, aligned)
        else:
            dense = dense_pairs(H, W, r, split) * work["mem_entries"]
        kname = "affinity_topk_tc16_kernel (K1, fp16 three-term split)" if split == "f16" else \
            "affinity_topk_tc_kernel (K1, 3xTF32)"
        if split == "f16" and J > 1:
            kname = f"affinity_topk_tc16g_kernel (K1, fp16 three-term split, {J} jobs packed per query tile)"
        out = dict(metric="propagated frames/sec", value=value, unit="frames/s", n_gpus=world, steps=args.steps,
                   warmup=n_warm, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                   vs_baseline=None, dtype=("f16x3" if split == "f16" else "tf32x3") + " split, fp32 accumulate (fp32-faithful)",
                   data="synthetic",
                   config=dict(WORK, l2="inputs (420 MB features + 840 MB feature bank per clip) exceed the 126 MB L2",
                               clips_per_step_per_gpu=1, parallelism=f"videos sharded, dp{world}"),
                   e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                   gpu_launches=int(launches),
                   roofline=dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak,
                                 unit="TFLOP/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src,
                                 algorithmic_bytes=algorithmic_bytes(), k1_ms=k1,
                                 k1_share_of_step=k1 / (ms / args.steps),
                                 peak_source=f"{pk['src']} bf16 sustained {pk['bf16']} TF/s" + (" / 3 (three fp16 MMAs per MAC)" if split == "f16"
                                             else " / 2 (tf32) / 3 (3xTF32)"),
                                 flops_per_launch=work["flops_per_step"],
                                 # geometry: a 128-query tile multiplies the union of its queries' circles, in whole
                                 # key boxes.  frac_dense = what the tensor pipe itself sustains (dense MACs / peak)
                                 tile_overhead=dense / (work["in_mask_pairs"] * work["mem_entries"]),
                                 frac_dense=achieved / peak * dense / (work["in_mask_pairs"] * work["mem_entries"]),
                                 jobs_per_tile=J,
                                 # SURVEY 8d / north_star state the roofline as dense TF32 peak / 3 (3xTF32):
                                 frac_vs_3xtf32_roofline=achieved / (pk["bf16"] / 2.0 / 3.0)),
                   clocks=clocks)
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_sample(frames=2)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------- CPU baseline / reference
def _cpu_inputs(n_frames):
    """features + soft labels for n_frames consecutive full-memory query frames (CPU)."""
    from fgvc_b200 import synthetic as S
    h, w = WORK["image_hw"]
    T = WORK["precede_frames"] + 1 + n_frames
    torch.manual_seed(0)
    enc = S.davis_encoder(WORK["stride"], seed=0)
    frames = S.synthetic_video(T, h, w, seed=1000)
    feats = S.encode(enc, frames, batch=4)
    g = torch.Generator().manual_seed(1)
    labels = torch.rand(T, WORK["objects"], *WORK["feat_hw"], generator=g)
    labels = labels / labels.sum(1, keepdim=True)
    return feats, labels


def _cpu_frame(O, feats, labels, mask, t):
    mem = O.memory_frames(t, WORK["precede_frames"])
    k = feats[mem].permute(1, 0, 2, 3)[None]
    v = labels[mem].permute(1, 0, 2, 3)[None]
    return O.propagate_port(feats[t][None], k, v, mask=mask, temperature=WORK["temperature"], topk=WORK["topk"],
                            step=512)


def cpu_baseline_sample(frames=2):
    from oracle import oracle as O        # the CPU baseline leg: the one place bench.py runs the oracle
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    feats, labels = _cpu_inputs(frames)
    mask = O.neighbor_mask(*WORK["feat_hw"], WORK["neighbor_range"])
    t0 = WORK["precede_frames"] + 1
    _cpu_frame(O, feats, labels, mask, t0)                      # warm-up
    s = time.perf_counter()
    for i in range(frames):
        _cpu_frame(O, feats, labels, mask, t0 + i)
    dt = time.perf_counter() - s
    return dict(value=frames / dt, unit="frames/s", cores=cores, kind="port",
                sample=f"{frames} propagated frames of the same workload at full memory (21 entries), "
                       f"oracle port of the reference op sequence, torch CPU fp32, mask build excluded")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    n = 2
    feats, labels = _cpu_inputs(n)
    mask = O.neighbor_mask(*WORK["feat_hw"], WORK["neighbor_range"])
    t0 = WORK["precede_frames"] + 1
    for i in range(max(1, args.warmup)):
        _cpu_frame(O, feats, labels, mask, t0 + i % n)
    s = time.perf_counter()
    for i in range(args.steps):
        _cpu_frame(O, feats, labels, mask, t0 + i % n)
    dt = time.perf_counter() - s
    value = args.steps / dt
    sample = "one full-memory propagated frame (21 entries) per step, oracle port, torch CPU fp32, all host threads"
    print(json.dumps(dict(impl="reference", metric="propagated frames/sec", value=value, unit="frames/s",
                          n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=args.steps, warmup=max(1, args.warmup),
                          ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="f32", data="synthetic", config=dict(WORK),
                          cpu_baseline=dict(value=value, unit="frames/s", cores=cores, kind="port", sample=sample),
                          e2e=dict(value=value, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--split", default=None, choices=["f16", "tf32"],
                    help="feature-bank split / tensor engine (default: f16 three-term; tf32 = 3xTF32)")
    ap.add_argument("--profile", action="store_true", help="kernels only (for ncu): no e2e, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 3
        return run_reference(args)
    if args.steps is None:
        args.steps = 20
    run_ours(args)


if __name__ == "__main__":
    main()
