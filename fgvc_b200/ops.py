"""Host-side mirror of the reference's label-propagation operators, same names and
argument meaning, computing on the B200 kernels through the C ABI.

Reference interfaces (paths relative to the FGVC repository):
  masked_attention_efficient      mmpt/models/common/local_attention.py:267-389
  masked_attention_efficient_v2   mmpt/models/common/local_attention.py:392-508
  masked_attention_efficient_c2f  mmpt/models/common/local_attention.py:721-880
  spatial_neighbor                mmpt/models/common/affinity_utils.py:75-112

Differences, all loud: N must be 1 (the reference driver asserts it and its gather
indexes batch 0 only, local_attention.py:360-362); ``topk > 16`` and ``sim_mode='l2-distance'``
without normalisation raise NotImplementedError (``topk=None``, the dense soft-max, runs as a flash-style kernel); ``step`` is accepted and
ignored (nothing is chunked: the affinity never exists in HBM); a ``mask`` tensor must be
one that ``spatial_neighbor`` produces (its radius is recovered and verified), because the
kernels evaluate the mask analytically.  There is no CPU fallback.
"""
import torch

from . import _lib, engine
from .engine import FeatureBank, JobTable, LabelBank


class NeighborMask(torch.Tensor):
    """bool [H*W, H*W] tensor returned by :func:`spatial_neighbor` that remembers how it
    was built, so the operators need not re-derive the radius from 268 M booleans."""

    @staticmethod
    def __new__(cls, data, spec):
        obj = torch.Tensor._make_subclass(cls, data)
        obj.spec = spec
        return obj

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **(kwargs or {}))


def spatial_neighbor(batches, height, width, neighbor_range, device, dtype, dim=1, mode="circle"):
    """Same contract as the reference: bool mask, [H*W, H*W] for ``circle`` and
    [batches, H*W, H*W] for ``square`` (affinity_utils.py:85-112)."""
    assert dim in [1, 2]
    assert mode in ["circle", "square"]
    ys = torch.arange(height, device=device).view(-1, 1).expand(height, width).reshape(-1)
    xs = torch.arange(width, device=device).view(1, -1).expand(height, width).reshape(-1)
    dy = ys.view(-1, 1) - ys.view(1, -1)
    dx = xs.view(-1, 1) - xs.view(1, -1)
    if mode == "circle":
        r = neighbor_range // 2
        m = (dy * dy + dx * dx) < r * r
    else:
        nr = (neighbor_range, neighbor_range) if isinstance(neighbor_range, int) else tuple(neighbor_range)
        assert nr[0] == nr[1], "square masks must be isotropic for the B200 kernels"
        m = (dy.abs() <= nr[0] // 2) & (dx.abs() <= nr[1] // 2)
        m = m.view(1, height * width, height * width).expand(batches, -1, -1)
    return NeighborMask(m, dict(height=height, width=width, mode=mode,
                                radius=(neighbor_range if isinstance(neighbor_range, int) else neighbor_range[0]) // 2))


def _mask_spec(mask, Hk, Wk, Hq, Wq):
    """(mode, radius) of a reference-style mask tensor; raises if it is not one."""
    spec = getattr(mask, "spec", None)
    if spec is not None:
        assert (spec["height"], spec["width"]) == (Hk, Wk)
        return spec["mode"], spec["radius"]
    m = mask[0] if mask.dim() == 3 else mask
    assert m.shape == (Hk * Wk, Hq * Wq)
    cached = getattr(mask, "_fgvc_spec", None)       # remembered on the tensor object itself
    if cached is not None and cached[0] == mask._version:
        return cached[1]
    if (Hk, Wk) != (Hq, Wq):
        raise NotImplementedError("masks between different query/key grids are not supported")
    m = m.bool()
    cy, cx = Hk // 2, Wk // 2
    col = m[:, cy * Wk + cx].view(Hk, Wk)
    reach = int(col[cy, cx:].sum().item()) - 1          # largest dx in-mask along the row
    found = None
    for mode, radius in (("circle", reach + 1), ("square", reach)):
        want = spatial_neighbor(1, Hk, Wk, 2 * radius, m.device, torch.float32, mode=mode)
        want = want.as_subclass(torch.Tensor)
        want = want[0] if want.dim() == 3 else want
        if torch.equal(want, m):
            found = (mode, radius)
            break
    if found is None:
        raise NotImplementedError("mask is not a spatial_neighbor circle/square mask; the B200 kernels "
                                  "evaluate the mask analytically")
    try:
        mask._fgvc_spec = (mask._version, found)
    except AttributeError:
        pass
    return found


def _check_common(query, key, value, mode, sim_mode, topk):
    assert mode in ["softmax", "cosine"]
    assert query.size(0) == key.size(0) == value.size(0)
    if query.size(0) != 1:
        raise NotImplementedError("batch size must be 1 (as in the reference driver, vanilla_tracker.py:134)")
    assert sim_mode in ["dot_product", "l2-distance"]
    if topk is not None and not (1 <= topk <= 16):
        raise NotImplementedError("topk must be in [1, 16]")
    _lib.require_cuda()
    for t in (query, key, value):
        if not t.is_cuda:
            raise _lib.FgvcError("fgvc_b200 operators take CUDA tensors; there is no CPU fallback")


def _propagate(query, key, value, radius, mask_mode, temperature, topk, normalize, non_mask_len, engine_id,
               groups=None, split=None, mode="softmax", sim_mode="dot_product"):
    C, Hq, Wq = query.shape[1:]
    T, Hk, Wk = key.shape[2:]
    L = value.size(1)
    if (Hq, Wq) != (Hk, Wk):
        raise NotImplementedError("query and key grids must match")
    temperature, flags = engine.sim_params(None, C, temperature, mode, sim_mode, normalize)
    dev = query.device
    query = query.float().contiguous()
    key = key.float().contiguous()
    value = value.float().contiguous()
    if split is None and C % 4 == 0 and engine_id == _lib.ENGINE_SIMT:
        split = "tf32"          # the CUDA-core engine then sees the exact fp32 values
    feats = FeatureBank(T + 1, C, Hk, Wk, dev, split=split)
    feats.load(key, 0, T, Hk * Wk, T * Hk * Wk, normalize)          # key[0,:,t] in place
    feats.load(query, T, 1, 0, Hq * Wq, normalize)
    labels = LabelBank(T + 1, L, Hk, Wk, dev)
    for t in range(T):
        labels.put_nchw(value[0, :, t], t, chan_stride=T * Hk * Wk)
    table = JobTable()
    table.add(T, list(range(T)), list(range(T)), T, unmasked=non_mask_len if radius is not None else T)
    r = radius if radius is not None else 1
    if topk is None:      # dense soft-max over every allowed candidate (local_attention.py:376-383)
        engine.dense_propagate(feats, table, 0, 1, labels, r, temperature, flags, mask_mode)
    else:
        lists = engine.affinity_topk(feats, table, r, topk, mask_mode, groups=groups, engine=engine_id)
        engine.gather_labels(lists, table, 0, 1, labels, temperature, flags)
    return labels.get_nchw(T).view(1, L, Hq, Wq)


def masked_attention_efficient(query, key, value, mask, temperature=1, topk=None, normalize=True, step=32,
                               non_mask_len=0, mode="softmax", sim_mode="dot_product", engine_id=_lib.ENGINE_AUTO,
                               split=None):
    """query [1,C,H,W]; key [1,C,T,H,W]; value [1,L,T,H,W]; mask bool [H*W,H*W] or None.
    Returns [1,L,H,W] on the input device (local_attention.py:267-389)."""
    _check_common(query, key, value, mode, sim_mode, topk)
    assert value.shape[2:] == key.shape[2:], f"{value.shape} {key.shape}"
    if key.ndim == 4:
        key = key.unsqueeze(2)
        value = value.unsqueeze(2)
    assert value.ndim == key.ndim == 5
    clip_len = key.size(2)
    assert 0 <= non_mask_len < clip_len
    if mask is None:
        radius, mask_mode = None, "circle"
    else:
        mask_mode, radius = _mask_spec(mask, key.shape[3], key.shape[4], query.shape[2], query.shape[3])
    return _propagate(query, key, value, radius, mask_mode, temperature, topk, normalize, non_mask_len, engine_id,
                      split=split, mode=mode, sim_mode=sim_mode)


def masked_attention_efficient_v2(query, key, value, radius, temperature=1, topk=None, normalize=True, step=32,
                                  non_mask_len=0, mode="softmax", sim_mode="dot_product",
                                  engine_id=_lib.ENGINE_AUTO, split=None):
    """Radius given directly; every memory frame is masked (local_attention.py:392-508,
    which accepts but never reads non_mask_len / sim_mode)."""
    _check_common(query, key, value, mode, "dot_product", topk)
    assert value.shape[2:] == key.shape[2:], f"{value.shape} {key.shape}"
    if key.ndim == 4:
        key = key.unsqueeze(2)
        value = value.unsqueeze(2)
    assert value.ndim == key.ndim == 5
    assert 0 <= non_mask_len < key.size(2)
    return _propagate(query, key, value, int(radius), "circle", temperature, topk, normalize, 0, engine_id,
                      split=split, mode=mode)


def masked_attention_efficient_c2f(query, key, query_fine, key_fine, value, mask, temperature=1, topk=None,
                                   normalize=True, step=32, non_mask_len=0, mode="softmax",
                                   sim_mode="dot_product", radius_fine=12, engine_id=_lib.ENGINE_AUTO, split=None):
    """Coarse-to-fine propagation (local_attention.py:721-880).  ``value`` lives on the FINE
    grid [1,L,T,s*Hk,s*Wk]; the output on the COARSE query grid [1,L,Hq,Wq]."""
    if topk is None:
        raise NotImplementedError("masked_attention_efficient_c2f: topk=None is not built (the fine stage selects)")
    _check_common(query, key, value, mode, sim_mode, topk)
    if mode != "softmax" or sim_mode != "dot_product":
        raise NotImplementedError("c2f is built for mode='softmax', sim_mode='dot_product'")
    if key.ndim == 4:
        key, value, key_fine = key.unsqueeze(2), value.unsqueeze(2), key_fine.unsqueeze(2)
    assert value.ndim == key.ndim == 5
    T = key.size(2)
    assert 0 <= non_mask_len < T
    C, Hq, Wq = query.shape[1:]
    Hk, Wk = key.shape[3:]
    Cf, Hf, Wf = key_fine.shape[1], key_fine.shape[3], key_fine.shape[4]
    L = value.size(1)
    if (Hq, Wq) != (Hk, Wk):
        raise NotImplementedError("query and key grids must match")
    assert query_fine.shape[2:] == key_fine.shape[3:] == value.shape[3:]
    if mask is None:
        radius, mask_mode, unmasked = 1, "circle", T
    else:
        mask_mode, radius = _mask_spec(mask, Hk, Wk, Hq, Wq)
        unmasked = non_mask_len
    dev = query.device
    if split is None:
        split = "f16" if (engine.default_split(C) == "f16" and Cf % 4 == 0 and engine_id != _lib.ENGINE_SIMT) else "tf32"
    coarse = FeatureBank(T + 1, C, Hk, Wk, dev, split=split)
    key = key.float().contiguous()
    coarse.load(key, 0, T, Hk * Wk, T * Hk * Wk, normalize)
    coarse.load(query.float().contiguous(), T, 1, 0, Hq * Wq, normalize)
    fine = FeatureBank(T + 1, Cf, Hf, Wf, dev, split=split)
    key_fine = key_fine.float().contiguous()
    fine.load(key_fine, 0, T, Hf * Wf, T * Hf * Wf, normalize)
    fine.load(query_fine.float().contiguous(), T, 1, 0, Hf * Wf, normalize)
    labels = LabelBank(T, L, Hf, Wf, dev)
    value = value.float().contiguous()
    for t in range(T):
        labels.put_nchw(value[0, :, t], t, chan_stride=T * Hf * Wf)
    table = JobTable()
    table.add(T, list(range(T)), list(range(T)), 0, unmasked=unmasked)
    out = engine.c2f_propagate(coarse, fine, table, 0, labels, radius, radius_fine, topk, temperature, mask_mode,
                               engine_id)
    return out[:, :L].t().reshape(1, L, Hq, Wq).contiguous()


# ------------------------------------------------------------------------------------------------
# Legacy dense-affinity utilities of the reference's API surface (mmpt/models/common/affinity_utils.py:6-73;
# SURVEY.md section 8 row a11).  Nothing in the reference calls them and they are not part of the accelerated path:
# plain torch tensor algebra on whatever device the inputs live on, written from the semantics:
#   compute_affinity : A[b, i, j] = <src[b, :, i], dst[b, :, j]> / temperature over flattened pixels, -inf where the
#                      mask is off, optional soft-max along `softmax_dim`; fully masked lines give 0, not NaN
#   propagate*       : out[b, c, j] = sum_i img[b, c, i] * A'[b, i, j], where with `topk` every column j of A keeps
#                      only what exceeds its k-th largest source entry, re-normalised to unit L1 mass
def _pixels(x):
    """[B, C, *spatial] -> [B, C, n]"""
    return x.flatten(2)


def compute_affinity(src_img, dst_img, temperature=1., normalize=True, softmax_dim=None, mask=None):
    src, dst = _pixels(src_img), _pixels(dst_img)
    if normalize:
        src, dst = (torch.nn.functional.normalize(t, p=2, dim=1) for t in (src, dst))
    scores = torch.einsum("bci,bcj->bij", src, dst) / temperature
    if mask is not None:
        scores = torch.where(mask.bool(), scores, scores.new_full((), float("-inf")))
    if softmax_dim is not None:
        scores = torch.softmax(scores, dim=softmax_dim)
    if mask is not None:
        scores = torch.nan_to_num(scores, nan=0.0)       # a line without any allowed entry soft-maxes to NaN
    return scores


def _keep_above_kth(weights, k):
    """per column (dim 1 = source pixels): subtract the k-th largest entry, drop what is not above it, unit L1 mass"""
    kth = torch.kthvalue(weights, weights.shape[1] - k + 1, dim=1, keepdim=True).values
    excess = torch.relu(weights - kth)
    return excess / excess.sum(dim=1, keepdim=True).clamp_min(1e-12)


def propagate(img, affinity, topk=None):
    weights = affinity if topk is None else _keep_above_kth(affinity, topk)
    return torch.einsum("bci,bij->bcj", _pixels(img), weights).reshape(img.shape)


def propagate_temporal(imgs, affinities, topk=None):
    B, C, T, H, W = imgs.shape
    if tuple(affinities.shape) != (B, T, H * W, H * W):
        raise AssertionError(f"affinities {tuple(affinities.shape)} do not match imgs {tuple(imgs.shape)}: "
                             f"expected {(B, T, H * W, H * W)}")
    weights = affinities.reshape(B, T * H * W, H * W)        # every (frame, pixel) of the clip is a source
    if topk is not None:
        weights = _keep_above_kth(weights, topk)
    return torch.einsum("bci,bij->bcj", _pixels(imgs), weights).reshape(B, C, H, W)
