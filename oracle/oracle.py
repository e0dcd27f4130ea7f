"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of FGVC's label-propagation path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker / the CPU
baseline.  The product path (``fgvc_b200``) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the *genuine* reference functions
executed in the build container (``oracle/ref_loader.py``): ``oracle/gen_golden.py``
writes their outputs to ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py``
checks every function below against those fixtures on every run.

Each function cites the reference lines it restates (paths relative to /root/reference).
Everything is N=1 (the reference driver asserts B == 1, vanilla_tracker.py:134).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- mask
def neighbor_mask(height, width, neighbor_range, mode="circle"):
    """bool [H*W (key), H*W (query)].  mmpt/models/common/affinity_utils.py:75-112.

    circle: euclidean distance < neighbor_range // 2 (strict);
    square: |dy| <= nr_y // 2 and |dx| <= nr_x // 2 (window clipped to the image).
    """
    ys = torch.arange(height).view(-1, 1).expand(height, width).reshape(-1)
    xs = torch.arange(width).view(1, -1).expand(height, width).reshape(-1)
    dy = ys.view(-1, 1) - ys.view(1, -1)
    dx = xs.view(-1, 1) - xs.view(1, -1)
    if mode == "circle":
        r = neighbor_range // 2
        return (dy * dy + dx * dx) < r * r
    if mode == "square":
        nr = (neighbor_range, neighbor_range) if isinstance(neighbor_range, int) else tuple(neighbor_range)
        return (dy.abs() <= nr[0] // 2) & (dx.abs() <= nr[1] // 2)
    raise AssertionError(mode)


def _unit(x, dim):
    """F.normalize(p=2, eps=1e-12) semantics, local_attention.py:308-310."""
    return x / x.norm(dim=dim, keepdim=True).clamp_min(1e-12)


# ------------------------------------------------------------------ propagation (port)
def propagate_port(query, key, value, mask=None, radius=None, temperature=1.0, topk=10,
                   normalize=True, step=512, non_mask_len=0, mode="softmax",
                   sim_mode="dot_product"):
    """fp32 torch port of ``masked_attention_efficient`` (local_attention.py:267-389)
    and, when ``radius`` is given instead of ``mask``, of ``_v2`` (:392-508, which masks
    every memory frame and ignores non_mask_len).

    query [1,C,Hq,Wq]; key [1,C,T,Hk,Wk] (or 4-D); value [1,L,T,Hk,Wk]; mask bool
    [Nk,Nq] or None.  Returns [1,L,Hq,Wq].  Same op sequence as the reference per query
    chunk: affinity GEMM / temperature -> masked_fill(-inf) -> topk -> gather ->
    softmax -> weighted sum.
    """
    assert query.shape[0] == 1 and key.shape[0] == 1 and value.shape[0] == 1
    if key.dim() == 4:
        key, value = key[:, :, None], value[:, :, None]
    C, Hq, Wq = query.shape[1:]
    T, Hk, Wk = key.shape[2:]
    L = value.shape[1]
    assert 0 <= non_mask_len < T
    Nq, Nk = Hq * Wq, Hk * Wk
    if radius is not None:
        assert mask is None and (Hq, Wq) == (Hk, Wk)
        mask = neighbor_mask(Hk, Wk, 2 * radius)
        non_mask_len = 0
    if normalize:
        query, key = _unit(query, 1), _unit(key, 1)
    qv = query.reshape(C, Nq)
    kv = key.reshape(C, T * Nk)
    vv = value.reshape(L, T * Nk)
    out = torch.zeros(L, Nq, dtype=query.dtype)
    step = step or Nq
    for s in range(0, Nq, step):
        e = min(Nq, s + step)
        if sim_mode == "dot_product":
            aff = (kv.t() @ qv[:, s:e]) / temperature            # [T*Nk, n]
        elif sim_mode == "l2-distance":
            aff = (2 * (kv.t() @ qv[:, s:e]) - kv.pow(2).sum(0)[:, None]) / math.sqrt(C)
        else:
            raise AssertionError(sim_mode)
        if mask is not None:
            keep = mask[:, s:e].repeat(T - non_mask_len, 1)
            if non_mask_len:
                keep = torch.cat([torch.ones(non_mask_len * Nk, e - s, dtype=torch.bool), keep])
            aff = aff.masked_fill(~keep, float("-inf"))
        if topk is not None:
            top_a, top_i = aff.topk(topk, dim=0)                # [k, n]
            picked = vv[:, top_i.reshape(-1)].reshape(L, topk, e - s)
            w = _weights(top_a, mode)
            out[:, s:e] = (picked * w[None]).sum(1)
        else:
            out[:, s:e] = vv @ _weights(aff, mode)
    return out.reshape(1, L, Hq, Wq)


def _weights(a, mode):
    if mode == "softmax":
        return a.softmax(dim=0)
    if mode == "cosine":
        return a.clamp(min=0) ** 2
    raise ValueError(mode)


# ----------------------------------------------------------- propagation (fp64 exact)
def propagate_exact(query, key, value, radius=None, temperature=1.0, topk=10, normalize=True,
                    masked=None, mask_mode="circle", chunk=1024):
    """fp64 brute-force statement of SURVEY.md Appendix A.2 (same semantics as
    ``propagate_port``; duplicates in the memory list are distinct candidates).

    ``masked``: per-memory-slot bool list (True -> radius mask applies); default all True
    when ``radius`` is given.  Returns dict(out [L,Hq,Wq] f64, idx [Nq,k] int64 into T*Nk
    (-1 where fewer than k candidates), aff [Nq,k] f64 (already / temperature),
    gap [Nq] = a_k - a_(k+1), +inf when there is no (k+1)-th candidate).
    ``gap`` classifies tie-ambiguous queries: an fp32 implementation may legitimately
    pick the other candidate when gap is below its rounding noise.
    """
    if key.dim() == 4:
        key, value = key[:, :, None], value[:, :, None]
    q = query[0].double()
    k = key[0].double()
    v = value[0].double()
    C, Hq, Wq = q.shape
    T, Hk, Wk = k.shape[1:]
    L = v.shape[0]
    Nq, Nk = Hq * Wq, Hk * Wk
    if normalize:
        q, k = _unit(q, 0), _unit(k, 0)
    qv = q.reshape(C, Nq)
    kv = k.reshape(C, T * Nk)
    vv = v.reshape(L, T * Nk)
    if masked is None:
        masked = [radius is not None] * T
    ky = torch.arange(Hk).view(-1, 1).expand(Hk, Wk).reshape(-1)
    kx = torch.arange(Wk).view(1, -1).expand(Hk, Wk).reshape(-1)
    out = torch.zeros(L, Nq, dtype=torch.float64)
    kk = min(topk, T * Nk)
    idx_all = torch.full((Nq, topk), -1, dtype=torch.int64)
    aff_all = torch.full((Nq, topk), float("-inf"), dtype=torch.float64)
    gap = torch.full((Nq,), float("inf"), dtype=torch.float64)
    for s in range(0, Nq, chunk):
        e = min(Nq, s + chunk)
        aff = (kv.t() @ qv[:, s:e]) / temperature                # [T*Nk, n]
        if radius is not None and any(masked):
            qy = torch.arange(s, e) // Wq
            qx = torch.arange(s, e) % Wq
            dy = ky[:, None] - qy[None]
            dx = kx[:, None] - qx[None]
            if mask_mode == "circle":
                ok = (dy * dy + dx * dx) < radius * radius
            else:
                ok = (dy.abs() <= radius) & (dx.abs() <= radius)
            keep = torch.cat([ok if m else torch.ones_like(ok) for m in masked])
            aff = aff.masked_fill(~keep, float("-inf"))
        srt, order = aff.sort(dim=0, descending=True, stable=True)
        top_a, top_i = srt[:kk], order[:kk]
        w = top_a.softmax(dim=0)
        w = torch.where(torch.isinf(top_a), torch.zeros_like(w), w)
        out[:, s:e] = (vv[:, top_i.reshape(-1)].reshape(L, kk, e - s) * w[None]).sum(1)
        idx_all[s:e, :kk] = torch.where(torch.isinf(top_a), torch.full_like(top_i, -1), top_i).t()
        aff_all[s:e, :kk] = top_a.t()
        if srt.shape[0] > kk:
            g = srt[kk - 1] - srt[kk]
            gap[s:e] = torch.where(torch.isnan(g) | torch.isinf(srt[kk]), torch.full_like(g, float("inf")), g)
    return dict(out=out.reshape(L, Hq, Wq), idx=idx_all, aff=aff_all, gap=gap)


# ------------------------------------------------------------------------------- c2f
def c2f_port(query, key, query_fine, key_fine, value, mask, temperature=1.0, topk=10,
             normalize=True, non_mask_len=0, radius_fine=12, dtype=None):
    """``masked_attention_efficient_c2f`` restated without the R^2 unfold blow-up
    (local_attention.py:721-880).

    Coarse stage: per memory frame, argmax of the masked coarse affinity (:804-837).
    Fine stage: the (2*radius_fine+1)^2 window of key_fine centred at scale*(ky,kx),
    zero padded -- padded candidates have affinity 0 and value 0 but still compete
    (:790-793, :847); top-k over T*R^2, softmax, weighted sum of fine labels (:859-871).
    Output lives on the coarse query grid.  ``dtype=torch.float64`` gives the exact form.
    """
    if key.dim() == 4:
        key, value, key_fine = key[:, :, None], value[:, :, None], key_fine[:, :, None]
    dt = dtype or query.dtype
    q, k = query[0].to(dt), key[0].to(dt)
    qf, kf, v = query_fine[0].to(dt), key_fine[0].to(dt), value[0].to(dt)
    C, Hq, Wq = q.shape
    T, Hk, Wk = k.shape[1:]
    Cf, Hkf, Wkf = kf.shape[0], kf.shape[2], kf.shape[3]
    L = v.shape[0]
    scale = Hkf // Hk
    if normalize:
        q, k, qf, kf = _unit(q, 0), _unit(k, 0), _unit(qf, 0), _unit(kf, 0)
    Nq, Nk = Hq * Wq, Hk * Wk
    aff = torch.einsum("ctk,cq->tkq", k.reshape(C, T, Nk), q.reshape(C, Nq)) / temperature
    if mask is not None:
        keep = mask.view(1, Nk, Nq).expand(T - non_mask_len, Nk, Nq)
        if non_mask_len:
            keep = torch.cat([torch.ones(non_mask_len, Nk, Nq, dtype=torch.bool), keep])
        aff = aff.masked_fill(~keep, float("-inf"))
    best = aff.softmax(dim=1).argmax(dim=1)                         # [T, Nq]
    qfs = qf[:, ::scale, ::scale].reshape(Cf, -1)                   # [Cf, Nq]
    R = 2 * radius_fine + 1
    off = torch.arange(-radius_fine, radius_fine + 1)
    kfp = F.pad(kf, (radius_fine,) * 4)                             # zero padding
    vp = F.pad(v, (radius_fine,) * 4)
    out = torch.zeros(L, Nq, dtype=dt)
    cand_a = torch.empty(T * R * R, Nq, dtype=dt)
    for i in range(Nq):
        a_i, v_i = [], []
        for t in range(T):
            cy = int(best[t, i]) // Wk * scale
            cx = int(best[t, i]) % Wk * scale
            win = kfp[:, t, cy:cy + R, cx:cx + R].reshape(Cf, R * R)
            a_i.append((win * qfs[:, i:i + 1]).sum(0) / temperature)
            v_i.append(vp[:, t, cy:cy + R, cx:cx + R].reshape(L, R * R))
        a_i = torch.cat(a_i)
        v_i = torch.cat(v_i, dim=1)
        cand_a[:, i] = a_i
        ta, ti = a_i.topk(topk)
        out[:, i] = (v_i[:, ti] * ta.softmax(0)[None]).sum(1)
    return dict(out=out.reshape(1, L, Hq, Wq), best=best, cand=cand_a)


# ------------------------------------------------------------------- heat-map -> coords
def img2coord_port(maps, topk=5):
    """``VanillaTracker.img2coord`` (vanilla_tracker.py:172-191).

    maps: numpy [T,P,h,w] fp32.  Returns [2,P,T] float64 (x row 0, y row 1):
    the ``topk`` largest pixels, weights v/(sum v + 1e-9), weighted mean of
    (idx % w, idx // w); an all-zero map gives -1.
    """
    T, P, h, w = maps.shape
    flat = maps.reshape(T, P, h * w)
    order = np.argsort(flat, axis=-1)[..., -topk:]
    vals = np.take_along_axis(flat, order, axis=-1)
    vals = vals / (vals.sum(axis=-1, keepdims=True) + 1e-9)
    xy = np.zeros((2, P, T), dtype=float)
    xy[0] = (order % w * vals).sum(-1).T
    xy[1] = (order // w * vals).sum(-1).T
    xy[:, flat.transpose(1, 0, 2).sum(-1) == 0] = -1
    return xy


def gaussian_labels(points_xy, h, w, stride, sigma=6.0):
    """``draw_gaussion_map_online`` (vanilla_tracker.py:204-221): full-res maps [P,h,w]
    and their ``[::stride, ::stride]`` sub-sampling."""
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, w)
    ys = torch.arange(h, dtype=torch.float32).view(1, h, 1)
    px = points_xy[:, 0].view(-1, 1, 1)
    py = points_xy[:, 1].view(-1, 1, 1)
    g = torch.exp(-((xs - px) ** 2 + (ys - py) ** 2) / (2 * sigma ** 2)).float()
    return g, g[:, ::stride, ::stride].contiguous()


# ------------------------------------------------------------------------------ driver
def memory_frames(t, precede_frames, with_first=True):
    """Memory multiset for query frame t (vanilla_tracker.py:346-362): frame 0 is
    prepended even when the window already contains it."""
    win = list(range(max(0, t - precede_frames), t))
    return ([0] + win) if with_first else win


def track_clip_port(feats, points_xy, image_hw, cfg, propagate=None):
    """``forward_test_main`` from the feature bank on (vanilla_tracker.py:320-412;
    SURVEY.md A.1).  feats [T,C,Hf,Wf] fp32; points_xy [P,2] (x,y) in image pixels.
    Returns (labels list of [P,Hf,Wf], traj [T,P,2] float64)."""
    T, C, Hf, Wf = feats.shape
    h, w = image_hw
    stride = h // Hf
    full, lab0 = gaussian_labels(points_xy, h, w, stride)
    nr = cfg.get("neighbor_range")
    mask = neighbor_mask(Hf, Wf, nr, cfg.get("mask_mode", "circle")) if nr is not None else None
    labels = [lab0]
    preds = [full]
    prop = propagate or propagate_port
    for t in range(1, T):
        mem = memory_frames(t, cfg["precede_frames"], cfg.get("with_first", True))
        k = feats[mem].permute(1, 0, 2, 3)[None]
        v = torch.stack([labels[m] for m in mem], dim=1)[None]
        lab = prop(feats[t][None], k, v, mask=mask, temperature=cfg["temperature"],
                   topk=cfg["topk"], step=cfg.get("step", 32),
                   normalize=cfg.get("with_norm", True),
                   non_mask_len=0 if cfg.get("with_first_neighbor", True) else 1)[0]
        labels.append(lab)
        preds.append(F.interpolate(lab[None], size=(h, w), mode="bilinear", align_corners=False)[0])
    xy = img2coord_port(torch.stack(preds).numpy())
    return labels, np.transpose(xy, (2, 1, 0))


def track_clip_c2f_port(feats_c, feats_f, points_xy, image_hw, cfg, c2f=None, mask=None):
    """The coarse-to-fine clip loop of fgvc_b200/c2f_tracker.py, restated on CPU around ``c2f`` (default: c2f_port;
    oracle/gen_golden.py passes the GENUINE masked_attention_efficient_c2f, local_attention.py:721-880).  The
    reference has no driver for this operator: the loop is the builder's design --
    coarse output -> bilinear x s -> fine memory labels of the next frames; coordinates from the coarse output
    up-sampled to the image (vanilla_tracker.py:396-406).  Returns (coarse outputs [T-1][L,Hc,Wc], traj [T,P,2])."""
    T, C, Hc, Wc = feats_c.shape
    Hf, Wf = feats_f.shape[2:]
    h, w = image_hw
    full, lab0 = gaussian_labels(points_xy, h, w, h // Hf)
    if mask is None:
        mask = neighbor_mask(Hc, Wc, cfg["neighbor_range"], cfg.get("mask_mode", "circle"))
    labels, preds, outs = [lab0], [full], []
    for t in range(1, T):
        mem = memory_frames(t, cfg["precede_frames"], cfg.get("with_first", True))
        k = feats_c[mem].permute(1, 0, 2, 3)[None].contiguous()
        kf = feats_f[mem].permute(1, 0, 2, 3)[None].contiguous()
        v = torch.stack([labels[m] for m in mem], dim=1)[None].contiguous()
        if c2f is None:
            out = c2f_port(feats_c[t][None], k, feats_f[t][None], kf, v, mask, temperature=cfg["temperature"],
                           topk=cfg["topk"], radius_fine=cfg.get("radius_fine", 12))["out"][0]
        else:
            out = c2f(feats_c[t][None], k, feats_f[t][None], kf, v, mask, temperature=cfg["temperature"],
                      topk=cfg["topk"], step=cfg.get("step", 64), radius_fine=cfg.get("radius_fine", 12))[0]
        outs.append(out)
        labels.append(F.interpolate(out[None], size=(Hf, Wf), mode="bilinear", align_corners=False)[0])
        preds.append(F.interpolate(out[None], size=(h, w), mode="bilinear", align_corners=False)[0])
    xy = img2coord_port(torch.stack(preds).numpy())
    return outs, np.transpose(xy, (2, 1, 0))


# ---------------------------------------------------------------- local-window ("HR") propagation
def correlation_unfold(query, key, radius):
    """``mmcv.ops.Correlation(max_displacement=radius)(query, key).flatten(1, 2)`` restated with F.unfold
    (call sites vanilla_tracker.py:421-441): out[b, (dy+r)*(2r+1) + (dx+r), y, x] = <query[b,:,y,x],
    key[b,:,y+dy,x+dx]>, zero outside the image, no normalisation.  mmcv-full 1.5.2 is not vendored and not
    installable here, so this restatement is PARITY UNPINNED for that one op; what the tracker
    needs from it is only that displacement i of the correlation and of F.unfold(value, padding=r) name the same key,
    which holds for any ordering shared by both."""
    B, C, H, W = query.shape
    R = 2 * radius + 1
    unf = F.unfold(key, kernel_size=(R, R), padding=radius).reshape(B, C, R * R, H * W)
    return (unf * query.reshape(B, C, 1, H * W)).sum(1).reshape(B, R * R, H, W)


def hr_propagate_port(query, key, value, radius, temperature=1.0, topk=10, normalize=True):
    """One frame of ``HRVanillaTracker.forward_test_main`` (vanilla_tracker.py:541-563): correlation of the query
    with every memory frame in a (2r+1)^2 window, top-k over K * R^2 candidates (zero-padded window positions
    included: affinity 0, value 0), temperature after the top-k, soft-max, weighted sum of the unfolded labels.
    query [1,C,H,W]; key [1,C,K,H,W]; value [1,L,K,H,W] -> [1,L,H,W]."""
    C, H, W = query.shape[1:]
    K, L = key.shape[2], value.shape[1]
    R = 2 * radius + 1
    if normalize:
        query, key = _unit(query, 1), _unit(key, 1)
    kk = key[0].transpose(0, 1)                                           # [K,C,H,W]
    corr = correlation_unfold(query.repeat(K, 1, 1, 1), kk, radius)        # [K,R^2,H,W]
    unfold_v = F.unfold(value[0].transpose(0, 1), kernel_size=(R, R), padding=radius).reshape(K, L, R * R, H, W)
    corr = corr.reshape(1, K * R * R, H, W)
    unfold_v = unfold_v.transpose(0, 1).reshape(1, L, K * R * R, H, W)
    top_a, top_i = corr.topk(k=topk, dim=1)
    top_v = torch.gather(unfold_v, 2, top_i.unsqueeze(1).expand(1, L, topk, H, W))
    w = (top_a / temperature).softmax(dim=1)
    return torch.einsum("bckhw,bkhw->bchw", top_v, w)


def nearest_resize_port(seg, size):
    """``pil_nearest_interpolate`` (common/utils.py:39-56: mmcv.imresize(..., 'nearest', backend='pillow')):
    PIL's NEAREST picks source pixel floor((dst + 0.5) * in / out).  seg [H,W] integer tensor -> [size]."""
    H, W = seg.shape
    ys = ((torch.arange(size[0], dtype=torch.float64) + 0.5) * H / size[0]).floor().clamp(0, H - 1).long()
    xs = ((torch.arange(size[1], dtype=torch.float64) + 0.5) * W / size[1]).floor().clamp(0, W - 1).long()
    return seg[ys][:, xs]


def track_masks_hr_port(feats, ref_seg, out_hw, cfg):
    """The mask loop of ``forward_test_backward_save_mem`` (vanilla_tracker.py:663-798) from the feature bank on, for
    an un-padded clip whose original size is ``out_hw``: nearest-resized one-hot labels (:694-705), local-window
    propagation per frame, decode = bilinear up-sample, per-channel min-max where max > 0, arg-max (:769-798).
    feats [T,C,Hf,Wf]; ref_seg [h,w] integer.  Returns (label maps [T,L,Hf,Wf], masks [T,h,w] long)."""
    T, C, Hf, Wf = feats.shape
    lab0 = onehot_labels(nearest_resize_port(ref_seg, (Hf, Wf)))
    labels, soft = [lab0], [lab0]          # memory labels (hardened with hard_prop) / soft predictions
    r = cfg["neighbor_range"] // 2
    for t in range(1, T):
        mem = memory_frames(t, cfg["precede_frames"], cfg.get("with_first", True))
        k = feats[mem].permute(1, 0, 2, 3)[None]
        v = torch.stack([labels[m] for m in mem], dim=1)[None]
        lab = hr_propagate_port(feats[t][None], k, v, r, temperature=cfg["temperature"], topk=cfg["topk"],
                                normalize=cfg.get("with_norm", True))[0]
        labels.append(F.one_hot(lab.argmax(0), lab.shape[0]).permute(2, 0, 1).float()
                      if cfg.get("hard_prop", False) else lab)
        soft.append(lab)
    masks = torch.stack([ref_seg.long()] + [decode_masks_port(s_, out_hw) for s_ in soft[1:]])
    return torch.stack(soft), masks


def group_by_query_frame(query_points):
    """Grouping of ``forward_test`` when test_cfg.with_first is set
    (vanilla_tracker.py:246-295): ascending unique query frame; within a group the
    original point order.  query_points [P,3]=(t,x,y).  Returns [(t, point_indices)]."""
    ts = np.unique(np.asarray(query_points[:, 0]))
    return [(int(t), np.nonzero(np.asarray(query_points[:, 0]) == t)[0]) for t in ts]


# --------------------------------------------------------------------- VOS-style decode
def onehot_labels(seg, num_classes=None):
    """int mask [Hf,Wf] -> one-hot [L,Hf,Wf] float (vanilla_tracker.py:694-705)."""
    seg = torch.as_tensor(seg).long()
    return F.one_hot(seg, -1 if num_classes is None else num_classes).permute(2, 0, 1).float()


def decode_masks_port(label, out_hw):
    """Up-sample, per-channel min-max normalise where max > 0, argmax
    (vanilla_tracker.py:769-798).  label [L,Hf,Wf] -> int64 [h,w]."""
    p = F.interpolate(label[None], size=out_hw, mode="bilinear", align_corners=False)[0]
    lo = p.amin(dim=(1, 2), keepdim=True)
    hi = p.amax(dim=(1, 2), keepdim=True)
    p = torch.where(hi > 0, (p - lo) / (hi - lo + 1e-12), p)
    return p.argmax(dim=0)


# ----------------------------------------------------------- test-time data-parallel
def shard_indices(n_items, rank, world):
    """``DistributedSampler.__iter__`` with shuffle=False, samples_per_gpu=1
    (datasets/samplers/distributed_sampler.py:39-56): pad by wrapping, rank-strided."""
    per = int(math.ceil(n_items / world))
    idx = list(range(n_items))
    idx += idx[: per * world - n_items]
    return idx[rank: per * world: world]


def interleave_results(parts, size):
    """``collect_results_*`` ordering (apis/test.py:183-186, :231-235)."""
    out = []
    for row in zip(*parts):
        out.extend(row)
    return out[:size]


# --------------------------------------------------------------------- legacy utilities
def compute_affinity_port(src, dst, temperature=1.0, normalize=True, softmax_dim=None, mask=None):
    """affinity_utils.py:6-30."""
    B, C = src.shape[:2]
    a, b = src.reshape(B, C, -1), dst.reshape(B, C, -1)
    if normalize:
        a, b = _unit(a, 1), _unit(b, 1)
    aff = torch.bmm(a.transpose(1, 2), b) / temperature
    if mask is not None:
        aff = aff.masked_fill(~mask.bool(), float("-inf"))
    if softmax_dim is not None:
        aff = aff.softmax(dim=softmax_dim)
    if mask is not None:
        aff = torch.nan_to_num(aff, nan=0.0) if aff.isnan().any() else aff
    return aff


def propagate_legacy_port(img, affinity, topk=None):
    """affinity_utils.py:33-50: subtract the k-th largest, clamp, L1-renormalise."""
    B, L, H, W = img.shape
    aff = affinity.clone()
    if topk is not None:
        kth = aff.topk(topk, dim=1)[0][:, topk - 1].view(B, 1, H * W)
        aff = (aff - kth).clamp(min=0)
        aff = aff / aff.sum(dim=1, keepdim=True).clamp(min=1e-12)
    return torch.bmm(img.reshape(B, L, -1), aff).reshape(B, L, H, W)


# -------------------------------------------------------------------------- TAP-Vid metrics
def tapvid_metrics_port(query_points, gt_occluded, gt_tracks, pred_occluded, pred_tracks, query_mode="first"):
    """numpy restatement of ``compute_tapvid_metrics`` (tapvid_evaluation_datasets.py:106-249):
    query_points [b,n,3] (t,y,x); *_occluded bool [b,n,T]; *_tracks [b,n,T,2] (x,y).  The query frame
    (and, for 'first', everything before the first visible frame) is not evaluated; occlusion accuracy
    divides by the evaluated points of the WHOLE batch, as the reference does."""
    T = gt_tracks.shape[2]
    qf = np.round(query_points[..., 0]).astype(np.int32)
    ev = np.eye(T)[qf] == 0
    if query_mode == "first":
        for i in range(gt_occluded.shape[0]):
            first = np.where(gt_occluded[i] == 0)[0][0]
            ev[i, :first] = False
    elif query_mode != "strided":
        raise ValueError("Unknown query mode " + query_mode)
    out = {"occlusion_accuracy": ((pred_occluded == gt_occluded) & ev).sum((1, 2)) / ev.sum()}
    vis, pvis = ~gt_occluded, ~pred_occluded
    d2 = ((pred_tracks - gt_tracks) ** 2).sum(-1)
    fr, ja = [], []
    for th in (1, 2, 4, 8, 16):
        within = d2 < th * th
        correct = within & vis
        n_vis = (vis & ev).sum((1, 2))
        f = (correct & ev).sum((1, 2)) / n_vis
        tp = (correct & pvis & ev).sum((1, 2))
        fp = ((((~vis) & pvis) | ((~within) & pvis)) & ev).sum((1, 2))
        j = tp / (n_vis + fp)
        out[f"pts_within_{th}"], out[f"jaccard_{th}"] = f, j
        fr.append(f)
        ja.append(j)
    out["average_jaccard"] = np.stack(ja, 1).mean(1)
    out["average_pts_within_thresh"] = np.stack(fr, 1).mean(1)
    return out


# ------------------------------------------------------------------- comparison helper
def compare_labels(got, exact, gap=None, tol=1e-3, gap_eps=2e-4):
    """Parity report used by the GPU tests.  ``got``/``exact`` [L,H,W]; ``gap`` [Nq] from
    ``propagate_exact``.  Queries whose k-th/(k+1)-th affinity gap is below ``gap_eps``
    (in affinity/temperature units) are tie-ambiguous: fp32 rounding in *any*
    implementation (the reference's own GEMM included) decides them."""
    got = torch.as_tensor(got).double().reshape(exact.shape[0], -1)
    ex = exact.reshape(exact.shape[0], -1)
    err = (got - ex).abs().amax(dim=0)
    rep = dict(max_abs=float(err.max()), frac_bad=float((err > tol).double().mean()))
    if gap is not None:
        # gap == 0 exactly only for a duplicated memory frame (same key, same label):
        # either pick gives the same output, so it is not an ambiguity
        clear = (gap > gap_eps) | (gap == 0)
        rep["n_ambiguous"] = int((~clear).sum())
        rep["max_abs_clear"] = float(err[clear].max()) if clear.any() else 0.0
    if got.shape[0] > 1:
        rep["argmax_agree"] = float((got.argmax(0) == ex.argmax(0)).double().mean())
    return rep
