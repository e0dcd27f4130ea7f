"""Host-side logic that needs no GPU: job tables, memory multisets, sharding, the
distributed collect (gloo, world_size 2) and the API surface / error behaviour."""
import os
import sys

import numpy as np
import pytest

import fgvc_b200
import torch
import torch.multiprocessing as mp

from fgvc_b200 import apis, engine, ops
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_memory_frames_match_oracle():
    for t in range(1, 12):
        for p in (2, 5):
            for wf in (True, False):
                assert engine.memory_frames(t, p, wf) == O.memory_frames(t, p, wf)
    # grouped clips start at t0: the window is clipped at t0 and t0 is "first"
    assert engine.memory_frames(5, 2, True, first=3) == [3, 3, 4]
    assert engine.memory_frames(4, 5, True, first=3) == [3, 3]


def test_job_table_layout():
    tb = engine.JobTable()
    tb.add(1, [0, 0], [0, 0], 1)
    tb.add(2, [0, 0, 1], [0, 0, 1], 2, unmasked=1)
    assert tb.jobs == [(1, 0, 2, 1), (2, 2, 5, 2)]
    assert tb.mem_feat[2] == (0 | 0x40000000) and tb.mem_feat[3] == 0
    assert tb.max_mem == 3 and len(tb) == 2
    j, mf, ml = tb.device("cpu")
    assert j.dtype == torch.int32 and j.shape == (2, 4) and mf.shape == (5,)


def test_job_packing_tables_reproduce_every_memory_list():
    """JobTable.packed (host side of csrc/topk_tc16.cu): the union list of a tile group, read through union_pos,
    must give back every job's own memory multiset in its own order -- incl. frame 0 twice while t <= precede."""
    for precede, T, unmasked in ((3, 9, 0), (20, 30, 1), (5, 7, 0)):
        tb = engine.JobTable()
        for t in range(1, T):
            mem = engine.memory_frames(t, precede, True)
            tb.add(t, mem, mem, t, unmasked=unmasked)
        for J in (1, 2, 4):
            tg, uent, upos = tb.packed(0, len(tb), J, "cpu")
            assert tg.shape == (-(-len(tb) // J), 8) and upos.shape == (uent.numel(), 4)
            seen_jobs = []
            for g in tg.tolist():
                members, n, u0, u1 = g[:4], g[4], g[5], g[6]
                assert members[n:] == [-1] * (4 - n) and 1 <= n <= J
                # oldest frame first (the kernel walks the list backwards: newest first)
                slots = [int(x) & ~0x40000000 for x in uent[u0:u1].tolist()]
                assert slots == sorted(slots)
                for li in range(n):
                    job = members[li]
                    seen_jobs.append(job)
                    b, e = tb.jobs[job][1], tb.jobs[job][2]
                    got = sorted((int(upos[u, li]), int(uent[u])) for u in range(u0, u1) if int(upos[u, li]) >= 0)
                    assert [p for p, _ in got] == list(range(e - b))          # every position exactly once
                    assert [r for _, r in got] == tb.mem_feat[b:e]
                for li in range(n, 4):
                    assert (upos[u0:u1, li] == -1).all()
            assert seen_jobs == list(range(len(tb)))
            assert tb.union_sizes(0, len(tb), J) == [g[6] - g[5] for g in tg.tolist()]


def test_aligned_packing_covers_every_entry_once_and_wastes_nothing():
    """JobTable.packed(aligned=True): memory frames in J classes (slot % J), query frames of class a grouped with phase
    a.  Every memory entry of every job appears in exactly one class, every job has a tile group in EVERY class
    (possibly with no entry: its list must still be written), and in the steady state of a sliding window every
    job of a group uses every union entry (no wasted tile rows)."""
    precede, T, J = 20, 64, 4
    tb = engine.JobTable()
    for t in range(1, T):
        mem = engine.memory_frames(t, precede, True)
        tb.add(t, mem, mem, t)
    tg, uent, upos = tb.packed(0, len(tb), J, "cpu", aligned=True)
    covered = {j: [] for j in range(len(tb))}
    classes = {j: set() for j in range(len(tb))}
    for g in tg.tolist():
        members, n, u0, u1, cls = g[:4], g[4], g[5], g[6], g[7]
        assert 0 <= cls < J and 1 <= n <= J
        for u in range(u0, u1):
            assert (int(uent[u]) & ~0x40000000) % J == cls
        for li in range(n):
            classes[members[li]].add(cls)
            covered[members[li]] += [int(upos[u, li]) for u in range(u0, u1) if int(upos[u, li]) >= 0]
        # steady state (window full, first frame outside the window): no wasted rows
        if n == J and all(tb.jobs[m][0] > precede + J for m in members[:n]):
            assert (upos[u0:u1, :n] >= 0).all()
    for j in range(len(tb)):
        assert classes[j] == set(range(J))
        assert sorted(covered[j]) == list(range(tb.jobs[j][2] - tb.jobs[j][1]))
    plain = sum(tb.union_sizes(0, len(tb), J))
    aligned = sum(tb.union_sizes(0, len(tb), J, aligned=True))
    assert aligned < 0.93 * plain


def test_pick_packing_follows_the_memory_length():
    """engine.pick_packing (cost model over the exact key-box counts; needs the library, no GPU): long memories pack
    several consecutive query frames per tile, short ones do not; FGVC_PACK forces it."""
    def table(T, precede):
        tb = engine.JobTable()
        for t in range(1, T):
            mem = engine.memory_frames(t, precede, True)
            tb.add(t, mem, mem, t)
        return tb
    long_mem, short_mem = table(64, 20), table(50, 5)
    assert engine.pick_packing(long_mem, 0, len(long_mem), 60, 107, 12, 0)[0] in (2, 4)
    assert engine.pick_packing(short_mem, 0, len(short_mem), 128, 128, 15, 0)[0] in (1, 2)
    assert engine.pick_packing(long_mem, 3, 4, 60, 107, 12, 0) == (1, False)   # a single job has nothing to share
    # exact accounting used by bench.py: packing must reduce the dense pairs of the long-memory clip
    d1 = engine.dense_pairs(long_mem, 0, len(long_mem), 60, 107, 12, 0, 1)
    d4 = engine.dense_pairs(long_mem, 0, len(long_mem), 60, 107, 12, 0, 4)
    assert 0.6 * d1 < d4 < 0.9 * d1
    os.environ["FGVC_PACK"] = "2"
    try:
        assert engine.pick_packing(short_mem, 0, len(short_mem), 128, 128, 15, 0) == (2, False)
    finally:
        del os.environ["FGVC_PACK"]


def test_shared_pair_table_maps_every_entry_to_its_union_list():
    """engine.shared_pair_table: several with_first groups share K1 per (query frame, memory frame) pair."""
    T, precede = 12, 3
    tb, spans = engine.JobTable(), []
    for t0 in (0, 2, 5, 6):
        spans.append((len(tb), t0))
        for t in range(t0 + 1, T):
            mem = engine.memory_frames(t, precede, True, first=t0)
            tb.add(t, mem, mem, t, unmasked=1)
    ut, g, pr = engine.shared_pair_table(tb, spans, T)
    assert len(ut) == T - 1 and len(pr) == len(tb.mem_feat) and g == ut.max_mem
    for (q, b, e, _) in tb.jobs:
        for k in range(b, e):
            u, i = divmod(pr[k], g)
            uq, ub, ue, _ = ut.jobs[u]
            assert uq == q and ub + i < ue and ut.mem_feat[ub + i] == tb.mem_feat[k]
    for (_, b, e, _) in ut.jobs:                       # union entries are unique
        assert len(set(ut.mem_feat[b:e])) == e - b
    # a single group shares nothing: not worth it
    one = engine.JobTable()
    for t in range(1, T):
        mem = engine.memory_frames(t, precede, True)
        one.add(t, mem, mem, t)
    assert engine.shared_pair_table(one, [(0, 0)], T) is None


def test_pick_groups_fills_the_chip():
    assert engine.pick_groups(1, 60, 107, 21) >= 2          # one DAVIS frame: split the memory
    assert engine.pick_groups(63, 60, 107, 21) == 1         # a whole clip: 3528 CTAs = 23.8 waves already
    assert engine.pick_groups(1, 8, 8, 3) <= 3              # never more groups than memory frames
    assert engine.pick_groups(8, 60, 107, 21) >= 2          # 448 CTAs = 3.03 waves: split to fill the last wave


def test_sampler_matches_reference_sharding():
    ds = list(range(10))
    for world in (1, 2, 4):
        for rank in range(world):
            s = apis.DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=False)
            assert list(iter(s)) == O.shard_indices(10, rank, world)
    with pytest.raises(ValueError):
        apis.DistributedSampler(list(range(3)), num_replicas=4, rank=0, shuffle=False)


def test_spatial_neighbor_matches_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, "masks.npz"))
    for key in d.files:
        mode, H, W, nr = key.split("_")
        m = ops.spatial_neighbor(2, int(H), int(W), int(nr), "cpu", torch.float32, mode=mode)
        m_ = m.as_subclass(torch.Tensor)
        got = m_.numpy() if mode == "circle" else m_[0].numpy()
        assert (got == d[key]).all(), key
        assert ops._mask_spec(m, int(H), int(W), int(H), int(W)) == (mode, int(nr) // 2)


def test_mask_radius_recovered_from_plain_tensor():
    for mode, nr in (("circle", 8), ("square", 6), ("circle", 3)):
        m = ops.spatial_neighbor(1, 11, 13, nr, "cpu", torch.float32, mode=mode).as_subclass(torch.Tensor).clone()
        assert ops._mask_spec(m, 11, 13, 11, 13) == (mode, nr // 2)
    bad = torch.rand(11 * 13, 11 * 13) > 0.5
    with pytest.raises(NotImplementedError):
        ops._mask_spec(bad, 11, 13, 11, 13)


def test_unsupported_modes_raise_loudly():
    q, k, v = torch.randn(1, 32, 4, 4), torch.randn(1, 32, 1, 4, 4), torch.rand(1, 2, 1, 4, 4)
    with pytest.raises(fgvc_b200.FgvcError):                 # topk=None is built (dense kernel) but there is no CPU path
        ops.masked_attention_efficient(q, k, v, None, topk=None)
    with pytest.raises(NotImplementedError):
        ops.masked_attention_efficient(q, k, v, None, topk=17)
    with pytest.raises(NotImplementedError):
        ops.masked_attention_efficient_c2f(q, k, q, k, v, None, topk=None)
    with pytest.raises(NotImplementedError):
        engine.sim_params(None, 32, 0.07, sim_mode="l2-distance", normalize=False)
    assert engine.sim_params(None, 64, 0.07, mode="cosine", sim_mode="l2-distance") == (8.0, 3)
    with pytest.raises(AssertionError):
        ops.masked_attention_efficient(q, k, v, None, topk=5, mode="bogus")


def test_encoder_contract():
    from fgvc_b200.encoder import ResNetEncoder
    torch.manual_seed(0)
    enc = ResNetEncoder(depth=18, strides=(1, 2, 2, 1), out_indices=(2,), pool_type="none").eval()
    with torch.no_grad():
        y = enc(torch.randn(1, 3, 64, 96))
    assert y.shape == (1, 256, 8, 12) and (y >= 0).all()
    keys = enc.state_dict().keys()
    for k in ("conv1.conv.weight", "conv1.bn.running_mean", "layer3.0.downsample.conv.weight",
              "layer2.1.conv2.bn.weight"):
        assert k in keys


class _FakeModel(torch.nn.Module):
    def forward(self, test_mode=True, save_image=False, save_path=None, iteration=None, vid=None):
        i = int(vid[0])
        return (torch.full((1, 3, 2, 2), float(i)), torch.arange(i + 1, dtype=torch.float64), f"video{i}")


def _worker(rank, world, port, gpu_collect, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ds = [dict(vid=torch.tensor([i])) for i in range(5)]
    sampler = apis.DistributedSampler(ds, shuffle=False)
    loader = torch.utils.data.DataLoader(ds, batch_size=None, sampler=sampler)
    res = apis.multi_gpu_test(_FakeModel(), loader, gpu_collect=gpu_collect)
    if rank == 0:
        q.put([(float(r[0].mean()), r[1].tolist(), r[2]) for r in res])
    else:
        assert res is None
    dist.destroy_process_group()


@pytest.mark.parametrize("gpu_collect", [True, False])
def test_multi_gpu_test_world2_gloo(gpu_collect):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29511 + int(gpu_collect)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gpu_collect, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[2] for g in got] == [f"video{i}" for i in range(5)]          # reference order, padding cut
    assert [g[0] for g in got] == [float(i) for i in range(5)]
    assert got[3][1] == [0.0, 1.0, 2.0, 3.0]


def test_legacy_utilities_match_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, "legacy.npz"))
    a, b, img = (torch.from_numpy(d[k]) for k in ("a", "b", "img"))
    aff = ops.compute_affinity(a, b, temperature=0.07, softmax_dim=1)
    assert torch.allclose(aff, torch.from_numpy(d["aff"]), atol=1e-6)
    assert torch.allclose(ops.propagate(img, aff, topk=4), torch.from_numpy(d["prop"]), atol=1e-5)
    vid = torch.stack([img, img], dim=2)
    affs = torch.stack([aff, aff], dim=1)
    want = O.propagate_legacy_port(torch.cat([img, img], dim=0)[:2], aff, topk=None)   # shape check only
    out = ops.propagate_temporal(vid, affs, topk=None)
    assert out.shape == img.shape and want.shape == img.shape
    assert torch.allclose(out, 2 * ops.propagate(img, aff), atol=1e-5)


class _FakeTracker(torch.nn.Module):
    """forward_test contract of the tracker with a deterministic per-point 'track'."""

    def __init__(self, with_first):
        super().__init__()
        self.test_cfg = dict(with_first=with_first)

    def forward(self, test_mode=True, rgbs=None, query_points=None, trajectories=None, visibilities=None):
        T, P = rgbs.shape[1], query_points.shape[1]
        t = torch.arange(T, dtype=torch.float64).view(T, 1, 1)
        pred = (query_points[0, :, 1:].double().view(1, P, 2) + t)[None]          # x+t, y+t
        if not self.test_cfg["with_first"]:
            return trajectories, visibilities, pred, torch.zeros_like(visibilities), query_points
        perm = torch.argsort(query_points[0, :, 0], stable=True)
        return (trajectories[:, :, perm], visibilities[:, :, perm], pred[:, :, perm].float(),
                torch.zeros_like(visibilities), query_points[:, perm])


def _shard_worker(rank, world, port, with_first, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    T, P = 4, 7
    rgbs = torch.zeros(1, T, 3, 8, 8)
    qp = torch.cat([torch.randint(0, 3, (1, P, 1), generator=g).float(), torch.rand(1, P, 2, generator=g) * 8], dim=2)
    traj, vis = torch.rand(1, T, P, 2, generator=g), torch.zeros(1, T, P)
    model = _FakeTracker(with_first)
    got = apis.sharded_forward_test(model, rgbs, qp, traj, vis)
    want = model(test_mode=True, rgbs=rgbs, query_points=qp, trajectories=traj, visibilities=vis)
    ok = all(torch.allclose(a.double(), b.double()) for a, b in zip(got, want))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("with_first", [False, True])
def test_point_sharded_forward_test_world2_gloo(with_first):
    assert apis.point_shard(7, 0, 2) == (0, 3) and apis.point_shard(7, 1, 2) == (3, 7)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, 29531 + int(with_first), with_first, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _two_phase_worker(rank, world, port, q):
    """host side of the two-phase split of one long video (apis.frame_shard / gather_job_lists / gather_point_tracks):
    phase 1 = every rank fills the lists of its frame range, phase 2 = every rank tracks its slice of every group."""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ok = True
    for n_jobs in (7, 2, 1):
        lo, hi, per = apis.frame_shard(n_jobs, rank, world)
        buf = torch.full((per * world, 3, 5), -1.0)
        for j in range(lo, hi):
            buf[j] = float(j)                                  # "K1" of job j
        apis.gather_job_lists(buf, per, rank, world)
        ok &= all(bool((buf[j] == float(j)).all()) for j in range(n_jobs))
    sizes, T = [5, 1, 4, 0], 3
    pads = [-(-n // world) for n in sizes]
    local = torch.zeros(T, sum(pads), 2, dtype=torch.float64)
    off = 0
    for g, (n, pad) in enumerate(zip(sizes, pads)):
        plo, phi = apis.point_shard(n, rank, world)
        for i in range(plo, phi):                              # "track" of point i of group g
            local[:, off + i - plo, 0] = 100 * g + i
            local[:, off + i - plo, 1] = torch.arange(T, dtype=torch.float64)
        off += pad
    outs = apis.gather_point_tracks(local, sizes, rank, world)
    for g, (n, o) in enumerate(zip(sizes, outs)):
        ok &= tuple(o.shape) == (T, n, 2)
        ok &= bool((o[:, :, 0] == (100 * g + torch.arange(n, dtype=torch.float64))[None]).all())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_phase_long_video_split_world2_gloo():
    assert apis.frame_shard(7, 0, 2) == (0, 4, 4) and apis.frame_shard(7, 1, 2) == (4, 7, 4)
    assert apis.frame_shard(1, 1, 2) == (1, 1, 1) and apis.frame_shard(0, 0, 2) == (0, 0, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_phase_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_plan_chunks_cover_all_jobs_with_full_waves():
    for n, tiles in ((63, 56), (49, 128), (5, 56), (1, 56), (249, 128)):
        ch = engine.plan_chunks(n, tiles)
        assert ch[0][0] == 0 and ch[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and all(b > a for a, b in ch)
    ch = engine.plan_chunks(63, 56)
    waste = sum(-(-(b - a) * 56 // 148) for a, b in ch) / (63 * 56 / 148)
    assert waste < 1.15 and (ch[0][1] - ch[0][0]) <= 16 and len(ch) >= 3   # full waves, early start, overlap


def test_frame_runs_of_a_job_range():
    """Two-phase split: a rank reads its jobs' query frames, their memory windows and frame 0 -- nothing else."""
    from fgvc_b200 import engine, _lib
    t = engine.JobTable()
    for f in range(1, 40):
        mem = engine.memory_frames(f, 5, True)
        t.add(f, mem, mem, f, unmasked=1)
    assert engine.frame_runs(t, 0, len(t)) == [(0, 40)]
    assert engine.frame_runs(t, 19, 29) == [(0, 1), (15, 30)]          # jobs 19..28 = frames 20..29, windows from 15
    assert engine.frame_runs(t, 0, 3) == [(0, 4)]
    assert engine.frame_runs(t, 5, 5) == []
    covered = set()
    for r in range(4):
        lo, hi = r * 10, min(39, r * 10 + 10)
        for a, b in engine.frame_runs(t, lo, hi):
            covered.update(range(a, b))
    assert covered == set(range(40))


def test_in_mask_pairs_square_and_circle():
    import bench
    H, W, r = 9, 7, 2
    brute_c = sum(1 for y in range(H) for x in range(W) for v in range(H) for u in range(W)
                  if (y - v) ** 2 + (x - u) ** 2 < r * r)
    brute_s = sum(1 for y in range(H) for x in range(W) for v in range(H) for u in range(W)
                  if abs(y - v) <= r and abs(x - u) <= r)
    assert bench.in_mask_pairs(H, W, r) == brute_c and bench.in_mask_pairs(H, W, r, square=True) == brute_s
