"""Test-time data-parallel driver: same function names, arguments and return values as
mmpt/apis/test.py (``single_gpu_test`` :13, ``multi_gpu_test`` :62,
``collect_results_cpu`` :131, ``collect_results_gpu`` :192) and the rank-strided video
sharding of mmpt/datasets/samplers/distributed_sampler.py:12-56.

One process per GPU (``torch.distributed``; NCCL over NVLink on the GPU box, gloo in the
CPU tests).  Videos are independent, so there is no data-path collective: the only
exchange is the final gather of the per-video results.  Instead of pickling CUDA tensors
through two padded uint8 all_gathers or a shared file system, the tensors of each result
are packed into one typed device buffer per rank and gathered with a single
``all_gather`` (plus a tiny object gather for the shapes).
"""
import math
import pickle

import torch
import torch.distributed as dist
from torch.utils.data import Sampler


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class DistributedSampler(Sampler):
    """Rank-strided sharding; pads by wrapping so every rank gets the same count
    (distributed_sampler.py:39-56)."""

    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True, samples_per_gpu=1):
        r, w = get_dist_info()
        self.dataset = dataset
        self.num_replicas = w if num_replicas is None else num_replicas
        self.rank = r if rank is None else rank
        self.shuffle = shuffle
        self.samples_per_gpu = samples_per_gpu
        self.epoch = 0
        per = int(math.ceil(len(dataset) * 1.0 / self.num_replicas / samples_per_gpu))
        self.num_samples = per * samples_per_gpu
        self.total_size = self.num_samples * self.num_replicas
        if len(dataset) < self.num_replicas * samples_per_gpu:
            raise ValueError("You may use too small dataset and our distributed sampler cannot pad your "
                             "dataset correctly. We highly recommend you to use fewer GPUs to finish your work")

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return self.num_samples

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.epoch)
            indices = torch.randperm(n, generator=g).tolist()
        else:
            indices = list(range(n))
        indices += indices[: self.total_size - len(indices)]
        return iter(indices[self.rank: self.total_size: self.num_replicas])


def _batch_size(data):
    for v in data.values():
        if isinstance(v, torch.Tensor):
            return v.size(0)
    return 1


def single_gpu_test(model, data_loader, save_image=False, save_path=None, iteration=None):
    if save_image and save_path is None:
        raise ValueError("When 'save_image' is True, you should also set 'save_path'.")
    model.eval()
    results = []
    for idx, data in enumerate(data_loader):
        with torch.no_grad():
            results.append(model(test_mode=True, save_image=save_image, save_path=save_path, iteration=idx, **data))
    return results


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=False, save_image=False, save_path=None,
                   iteration=None, empty_cache=False):
    """Returns the ordered list of per-video results on rank 0, ``None`` elsewhere."""
    if save_image and save_path is None:
        raise ValueError("When 'save_image' is True, you should also set 'save_path'.")
    model.eval()
    results = []
    for data in data_loader:
        with torch.no_grad():
            results.append(model(test_mode=True, save_image=save_image, save_path=save_path, iteration=iteration,
                                 **data))
        if empty_cache:
            torch.cuda.empty_cache()
    size = len(data_loader.dataset)
    if gpu_collect:
        return collect_results_gpu(results, size)
    return collect_results_cpu(results, size, tmpdir)


# ------------------------------------------------------------------------------ collect
def _flatten(obj, tensors):
    """Replace tensors by placeholders, appending them to ``tensors``."""
    if isinstance(obj, torch.Tensor):
        tensors.append(obj)
        return ("__t__", len(tensors) - 1, tuple(obj.shape), str(obj.dtype).replace("torch.", ""))
    if isinstance(obj, (list, tuple)):
        out = [_flatten(o, tensors) for o in obj]
        return ("__tuple__", out) if isinstance(obj, tuple) else out
    if isinstance(obj, dict):
        return {k: _flatten(v, tensors) for k, v in obj.items()}
    return obj


def _unflatten(obj, payload, offsets):
    if isinstance(obj, tuple) and len(obj) == 4 and obj[0] == "__t__":
        _, i, shape, dt = obj
        dtype = getattr(torch, dt)
        n = int(torch.tensor(shape).prod().item()) if len(shape) else 1
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        raw = payload[offsets[i]: offsets[i] + nbytes].clone()
        return raw.view(dtype).reshape(shape)
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == "__tuple__":
        return tuple(_unflatten(o, payload, offsets) for o in obj[1])
    if isinstance(obj, list):
        return [_unflatten(o, payload, offsets) for o in obj]
    if isinstance(obj, dict):
        return {k: _unflatten(v, payload, offsets) for k, v in obj.items()}
    return obj


def _pack(result_part, device):
    tensors = []
    skeleton = _flatten(result_part, tensors)
    offsets, chunks, cur = [], [], 0
    for t in tensors:
        b = t.detach().contiguous().view(-1).view(torch.uint8).to(device)
        offsets.append(cur)
        pad = (-b.numel()) % 16          # keep every tensor 16-byte aligned for the dtype views
        chunks.append(b)
        if pad:
            chunks.append(torch.zeros(pad, dtype=torch.uint8, device=device))
        cur += b.numel() + pad
    payload = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.uint8, device=device)
    return skeleton, offsets, payload


def _interleave(part_list, size):
    ordered = []
    for res in zip(*part_list):
        ordered.extend(list(res))
    return ordered[:size]          # the sampler may have padded by wrapping


def collect_results_gpu(result_part, size):
    """One typed all_gather of the packed tensors (device = wherever the process group
    communicates: CUDA for NCCL, CPU for gloo)."""
    rank, world = get_dist_info()
    if world == 1:
        return result_part[:size]
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    skeleton, offsets, payload = _pack(result_part, device)
    meta = [None] * world
    dist.all_gather_object(meta, (skeleton, offsets, int(payload.numel())))
    longest = max(m[2] for m in meta)
    send = torch.zeros(max(longest, 16), dtype=torch.uint8, device=device)
    send[: payload.numel()] = payload
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    if rank != 0:
        return None
    parts = []
    for (skel, offs, n), buf in zip(meta, recv):
        parts.append(_unflatten(skel, buf[:n], offs))
    return _interleave(parts, size)


def collect_results_cpu(result_part, size, tmpdir=None):
    """Reference semantics (ordered list on rank 0) without the shared file system:
    results are moved to host memory and gathered as objects."""
    rank, world = get_dist_info()
    if world == 1:
        return result_part[:size]

    def to_cpu(o):
        if isinstance(o, torch.Tensor):
            return o.detach().cpu()
        if isinstance(o, tuple):
            return tuple(to_cpu(x) for x in o)
        if isinstance(o, list):
            return [to_cpu(x) for x in o]
        if isinstance(o, dict):
            return {k: to_cpu(v) for k, v in o.items()}
        return o

    blob = pickle.dumps(to_cpu(result_part))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(blob, gathered, dst=0)
    if rank != 0:
        return None
    return _interleave([pickle.loads(b) for b in gathered], size)


# ------------------------------------------------------------------ single long video: two-phase sharding
# the affinity / top-k lists depend only on features (phase 1: tensor-bound, parallel over frames),
# the recurrence lives only in the label gather (phase 2: sequential in time, parallel over label channels).  So one
# long video scales its COMPUTE across ranks by sharding phase 1 over frame ranges, exchanging the sparse lists once
# (8 * k * Nq bytes per frame), and sharding phase 2 over the tracked points; results are identical to one GPU.
def point_shard(n_points, rank, world):
    """Contiguous slice of the tracked points owned by ``rank`` (label channels never mix)."""
    lo = (n_points * rank) // world
    hi = (n_points * (rank + 1)) // world
    return lo, hi


def frame_shard(n_jobs, rank, world):
    """(lo, hi, per): jobs [lo, hi) of phase 1 owned by ``rank``; every rank owns a slot of ``per`` jobs in the
    gathered list buffer (the last ranks' slots may be partly or wholly padding)."""
    per = -(-n_jobs // world) if n_jobs else 0
    lo, hi = min(n_jobs, rank * per), min(n_jobs, (rank + 1) * per)
    return lo, hi, per


def gather_job_lists(buf, per, rank, world):
    """``buf`` [per * world, ...]: every rank filled its own slot [rank*per, (rank+1)*per); afterwards every rank
    holds all slots.  One all_gather (in place over NCCL: the send buffer is the rank's slot of the receive buffer)."""
    if world == 1 or per == 0:
        return buf
    assert buf.shape[0] == per * world and buf.is_contiguous()
    mine = buf[rank * per:(rank + 1) * per]
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(buf, mine)
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.clone())
        for r, p in enumerate(parts):
            buf[r * per:(r + 1) * per].copy_(p)
    return buf


def gather_point_tracks(local, group_sizes, rank, world):
    """Phase-2 result exchange.  ``local`` [T, sum_g pad_g, 2]: for every group g the tracks of this rank's point
    slice (point_shard of the group's size), left-aligned in a slot of pad_g = ceil(P_g / world) columns.  One
    all_gather; returns the per-group [T, P_g, 2] tracks in the original point order on every rank."""
    pads = [-(-n // world) for n in group_sizes]
    assert local.shape[1] == sum(pads)
    if world == 1:
        parts = [local]
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous())
    out, off = [], 0
    for n, pad in zip(group_sizes, pads):
        cols = []
        for r in range(world):
            lo, hi = point_shard(n, r, world)
            cols.append(parts[r][:, off:off + (hi - lo)])
        out.append(torch.cat(cols, dim=1) if cols else local[:, :0])
        off += pad
    return out


def sharded_forward_test(model, rgbs, query_points, trajectories, visibilities, **kw):
    """``forward_test`` of one (long) video over the ranks of the default process group; returns the reference's
    5-tuple (vanilla_tracker.py:227-303) on every rank, identical to the un-sharded call.
    A tracker that implements the two-phase split (``VanillaTracker``: ``forward_test(..., shard=(rank, world))``)
    shards the K1 launch over frame ranges, all-gathers the top-k lists, and propagates only its slice of the
    points; collectives: one all_gather of the lists, one of the ``[T, P_rank, 2]`` tracks.
    Any other model falls back to sharding the points only (every rank then repeats the label-independent work)."""
    rank, world = get_dist_info()
    if world == 1:
        return model(test_mode=True, rgbs=rgbs, query_points=query_points, trajectories=trajectories,
                     visibilities=visibilities, **kw)
    if getattr(model, "two_phase_sharding", False):
        return model(test_mode=True, rgbs=rgbs, query_points=query_points, trajectories=trajectories,
                     visibilities=visibilities, shard=(rank, world), **kw)
    assert rgbs.shape[0] == 1
    B, T = rgbs.shape[:2]
    P = query_points.shape[1]
    lo, hi = point_shard(P, rank, world)
    grouped = bool(model.test_cfg.get("with_first", False))
    qp = query_points[:, lo:hi]
    out = model(test_mode=True, rgbs=rgbs, query_points=qp, trajectories=trajectories[:, :, lo:hi],
                visibilities=visibilities[:, :, lo:hi], **kw) if hi > lo else None
    # undo the per-shard re-ordering by query frame so that shards concatenate in the original order
    per = max(point_shard(P, r, world)[1] - point_shard(P, r, world)[0] for r in range(world))
    dev = out[2].device if out is not None else (torch.device("cuda", torch.cuda.current_device())
                                                 if dist.get_backend() == "nccl" else torch.device("cpu"))
    dtype = torch.float32 if grouped else torch.float64
    local = torch.zeros(T, per, 2, dtype=dtype, device=dev)
    if out is not None:
        pred = out[2][0]
        if grouped:
            perm = torch.argsort(qp[0, :, 0].to(dev), stable=True)
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel(), device=dev)
            pred = pred[:, inv]
        local[:, : hi - lo] = pred.to(dtype)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    full = torch.cat([parts[r][:, : point_shard(P, r, world)[1] - point_shard(P, r, world)[0]] for r in range(world)],
                     dim=1)[None]
    query_points, trajectories, visibilities = (x.to(dev) for x in (query_points, trajectories, visibilities))
    if not grouped:
        return trajectories, visibilities, full, torch.zeros_like(visibilities), query_points
    gperm = torch.argsort(query_points[0, :, 0], stable=True)
    return (trajectories[:, :, gperm], visibilities[:, :, gperm], full[:, :, gperm].to(trajectories.dtype),
            torch.zeros_like(visibilities), query_points[:, gperm])
