"""Evaluation metrics and wire formats right after the path (SURVEY section 8f "next"), on torch tensors of any
device so the tracker's CUDA outputs need not leave the GPU.  Plain tensor ops: bookkeeping, not hot kernels.

  compute_tapvid_metrics  mmpt/datasets/tapvid_evaluation_datasets.py:106-249
  pck_distances / pck     JHMDB PCK, mmpt/datasets/jhmdb_dataset.py:143-256
  db_eval_iou / db_eval_boundary / db_statistics / jfm     DAVIS J&F, mmpt/core/evaluation/metrics.py:11-256
  tapvid_sample           TAP-Vid pickle record -> forward_test inputs, mmpt/datasets/tapvid.py:85-152


``compute_tapvid_metrics`` has the signature and semantics of the reference's
mmpt/datasets/tapvid_evaluation_datasets.py:106-249 (occlusion accuracy, points within
1/2/4/8/16 px, Jaccard, and their averages), on torch tensors of any device so the tracker's
CUDA outputs need not leave the GPU.  Plain tensor ops: bookkeeping, not a hot kernel.
"""
import torch


def compute_tapvid_metrics(query_points, gt_occluded, gt_tracks, pred_occluded, pred_tracks, query_mode,
                           additional_pck_thresholds=()):
    """query_points [b,n,3] (t,y,x); gt_occluded / pred_occluded bool [b,n,T]; gt_tracks / pred_tracks
    [b,n,T,2] (x,y) in raster coordinates.  Returns a dict of [b] float64 tensors."""
    gt_occluded, pred_occluded = gt_occluded.bool(), pred_occluded.bool()
    b, n, T = gt_occluded.shape
    dev = gt_tracks.device
    qf = torch.round(query_points[..., 0]).long().to(dev)
    ev = torch.ones(b, n, T, dtype=torch.bool, device=dev)
    ev.scatter_(2, qf.unsqueeze(-1), False)                       # the query frame is not evaluated
    if query_mode == "first":
        # reference quirk kept: np.where(gt_occluded[i] == 0)[0][0] is the index of the first POINT (row)
        # that has a visible frame, and the points before it are dropped from the evaluation
        for i in range(b):
            first = int(torch.nonzero(~gt_occluded[i])[0, 0]) if (~gt_occluded[i]).any() else 0
            ev[i, :first] = False
    elif query_mode != "strided":
        raise ValueError("Unknown query mode " + query_mode)
    f64 = torch.float64
    out = {"occlusion_accuracy": ((pred_occluded == gt_occluded) & ev).sum((1, 2)).to(f64) / ev.sum().to(f64)}
    vis, pvis = ~gt_occluded, ~pred_occluded
    d2 = ((pred_tracks.to(f64) - gt_tracks.to(f64)) ** 2).sum(-1)
    n_vis = (vis & ev).sum((1, 2)).to(f64)
    fracs, jacs = [], []
    for th in (1, 2, 4, 8, 16):
        within = d2 < th * th
        correct = within & vis
        frac = (correct & ev).sum((1, 2)).to(f64) / n_vis
        tp = (correct & pvis & ev).sum((1, 2)).to(f64)
        fp = ((((~vis) & pvis) | ((~within) & pvis)) & ev).sum((1, 2)).to(f64)
        out[f"pts_within_{th}"] = frac
        out[f"jaccard_{th}"] = tp / (n_vis + fp)
        fracs.append(frac)
        jacs.append(out[f"jaccard_{th}"])
    for th in additional_pck_thresholds:
        out[f"pts_within_{th}"] = (((d2 < th * th) & vis) & ev).sum((1, 2)).to(f64) / n_vis
    out["average_jaccard"] = torch.stack(jacs, 1).mean(1)
    out["average_pts_within_thresh"] = torch.stack(fracs, 1).mean(1)
    return out


# ------------------------------------------------------------------------------ JHMDB PCK
def pck_distances(pred_poses, gt_poses):
    """Normalised key-point errors of one video (jhmdb_dataset.py:201-233).  pred_poses / gt_poses [2,P,T] (x, y).
    A joint counts where its predicted x is > 0; per frame the error is divided by 0.6 x the diagonal of the box of
    the visible ground-truth joints.  Returns a list of P 1-D tensors (the errors of each joint over its frames)."""
    pred, gt = torch.as_tensor(pred_poses, dtype=torch.float64), torch.as_tensor(gt_poses, dtype=torch.float64)
    assert pred.shape == gt.shape and pred.shape[0] == 2
    vis = pred[0] > 0                                              # [P,T]
    hi = torch.where(vis[None], gt, torch.full_like(gt, -1.0)).amax(dim=1)        # [2,T]
    lo = torch.where(vis[None], gt, torch.full_like(gt, 1e6)).amin(dim=1)
    box = 0.6 * torch.linalg.norm(hi - lo, dim=0)                  # [T]
    dist = torch.linalg.norm(pred - gt, dim=0) / box[None]         # [P,T]
    return [dist[p][vis[p]] for p in range(pred.shape[1])]


def pck(dist_all, thresholds=(0.1, 0.2, 0.3, 0.4, 0.5)):
    """``compute_pck`` + the summary of ``pck_evaluate`` (jhmdb_dataset.py:143-152, 234-246): dist_all = per joint the
    concatenated errors of all videos.  Returns {'PCK@a': mean over joints of the % of errors <= a}."""
    out = {}
    for a in thresholds:
        per_joint = torch.stack([100.0 * (d <= a).double().mean() if d.numel() else torch.tensor(float("nan"),
                                 dtype=torch.float64) for d in dist_all])
        out[f"PCK@{a}"] = float(per_joint.mean())
    return out


# ------------------------------------------------------------------------------ DAVIS J & F
def db_eval_iou(annotation, segmentation, void_pixels=None):
    """Jaccard index per frame (metrics.py:11-42).  annotation / segmentation [..., H, W] binary."""
    a, s_ = torch.as_tensor(annotation).bool(), torch.as_tensor(segmentation).bool()
    assert a.shape == s_.shape
    keep = ~torch.as_tensor(void_pixels).bool() if void_pixels is not None else torch.ones_like(a)
    inter = ((s_ & a) & keep).sum(dim=(-2, -1)).double()
    union = ((s_ | a) & keep).sum(dim=(-2, -1)).double()
    return torch.where(union == 0, torch.ones_like(union), inter / union.clamp_min(1))


def _seg2bmap(seg):
    """1-pixel boundary map, offset by half a pixel towards the origin (metrics.py:127-183, same-size case)."""
    seg = seg.bool()
    e, s_, se = torch.zeros_like(seg), torch.zeros_like(seg), torch.zeros_like(seg)
    e[:, :-1] = seg[:, 1:]
    s_[:-1, :] = seg[1:, :]
    se[:-1, :-1] = seg[1:, 1:]
    b = (seg ^ e) | (seg ^ s_) | (seg ^ se)
    b[-1, :] = seg[-1, :] ^ e[-1, :]
    b[:, -1] = seg[:, -1] ^ s_[:, -1]
    b[-1, -1] = False
    return b


def _disk(radius, device):
    """skimage.morphology.disk: the pixels with dy^2 + dx^2 <= r^2"""
    r = int(radius)
    y, x = torch.meshgrid(torch.arange(-r, r + 1, device=device), torch.arange(-r, r + 1, device=device), indexing="ij")
    return (y * y + x * x <= r * r).float()


def f_measure(foreground_mask, gt_mask, void_pixels=None, bound_th=0.008):
    """Boundary F-measure of one frame (metrics.py:62-124): boundaries of both masks, each dilated by a disk of
    ceil(bound_th x image diagonal) pixels, precision / recall of the matches."""
    import math
    fg, gt = torch.as_tensor(foreground_mask).bool(), torch.as_tensor(gt_mask).bool()
    keep = ~torch.as_tensor(void_pixels).bool() if void_pixels is not None else torch.ones_like(fg)
    bound_pix = bound_th if bound_th >= 1 else math.ceil(bound_th * math.hypot(*fg.shape[-2:]))
    fb, gb = _seg2bmap(fg & keep), _seg2bmap(gt & keep)
    k = _disk(bound_pix, fg.device)[None, None]
    pad = k.shape[-1] // 2
    dil = lambda m: torch.nn.functional.conv2d(m[None, None].float(), k, padding=pad)[0, 0] > 0
    n_fg, n_gt = int(fb.sum()), int(gb.sum())
    if n_fg == 0 and n_gt > 0:
        precision, recall = 1.0, 0.0
    elif n_fg > 0 and n_gt == 0:
        precision, recall = 0.0, 1.0
    elif n_fg == 0 and n_gt == 0:
        precision, recall = 1.0, 1.0
    else:
        precision = float((fb & dil(gb)).sum()) / n_fg
        recall = float((gb & dil(fb)).sum()) / n_gt
    return 0.0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)


def db_eval_boundary(annotation, segmentation, void_pixels=None, bound_th=0.008):
    """F per frame for [T,H,W] (or one [H,W]) masks (metrics.py:45-59)."""
    a, s_ = torch.as_tensor(annotation), torch.as_tensor(segmentation)
    assert a.shape == s_.shape
    if a.ndim == 2:
        return f_measure(s_, a, void_pixels, bound_th)
    return torch.tensor([f_measure(s_[t], a[t], None if void_pixels is None else void_pixels[t], bound_th)
                         for t in range(a.shape[0])], dtype=torch.float64)


def db_statistics(per_frame_values):
    """mean, recall (> 0.5) and decay (first minus last quarter) of per-frame values (metrics.py:186-212)."""
    import numpy as np
    v = torch.as_tensor(per_frame_values, dtype=torch.float64)
    M = float(torch.nanmean(v))
    O = float(torch.nanmean((v > 0.5).double()))
    ids = (np.round(np.linspace(1, len(v), 5) + 1e-10) - 1).astype(np.uint8)
    bins = [v[int(ids[i]):int(ids[i + 1]) + 1] for i in range(4)]
    D = float(torch.nanmean(bins[0])) - float(torch.nanmean(bins[3]))
    return M, O, D


def jfm(all_gt_masks, all_res_masks, metric=("J", "F")):
    """``JFM`` (metrics.py:228-256): per object J / F mean, recall, decay.  Masks [objects, T, H, W] binary; missing
    result objects count as empty."""
    gt, res = torch.as_tensor(all_gt_masks).bool(), torch.as_tensor(all_res_masks).bool()
    if res.shape[0] > gt.shape[0]:
        raise ValueError("the results hold an index higher than the number of objects in the sequence")
    if res.shape[0] < gt.shape[0]:
        res = torch.cat([res, torch.zeros((gt.shape[0] - res.shape[0],) + tuple(res.shape[1:]), dtype=torch.bool,
                                          device=res.device)])
    out = {k: [] for k in ("JM", "JR", "JD", "FM", "FR", "FD")}
    for i in range(gt.shape[0]):
        j = db_eval_iou(gt[i], res[i]) if "J" in metric else torch.zeros(gt.shape[1], dtype=torch.float64)
        f = db_eval_boundary(gt[i], res[i]) if "F" in metric else torch.zeros(gt.shape[1], dtype=torch.float64)
        for key, val in zip(("JM", "JR", "JD"), db_statistics(j)):
            out[key].append(val)
        for key, val in zip(("FM", "FR", "FD"), db_statistics(f)):
            out[key].append(val)
    return out


# ------------------------------------------------------------------------------ TAP-Vid wire format
def tapvid_sample(sample, input_size, query_mode="first", stride=5):
    """One TAP-Vid pickle record -> the keyword arguments of ``forward_test`` (tapvid.py:85-152,
    tapvid_evaluation_datasets.py:297-401).  sample: {'video': [T,H,W,3] uint8 frames ALREADY resized to
    ``input_size`` (or a list of JPEG byte strings, decoded with PIL), 'points': [N,T,2] (x, y) in [0,1],
    'occluded': [N,T] bool}.  Frames are scaled to [-1, 1]; tracks to pixels of ``input_size`` = (h, w); the query of
    a track is its first visible frame ('first') or every ``stride``-th visible frame ('strided'); query points
    come out as (t, x, y) like ``preprocess_dataset_element`` makes them."""
    import io
    import numpy as np
    video = sample["video"]
    if len(video) and isinstance(video[0], (bytes, bytearray)):
        from PIL import Image
        video = np.stack([np.array(Image.open(io.BytesIO(f))) for f in video])
    video = np.asarray(video)
    frames = video.astype(np.float32) / 255.0 * 2.0 - 1.0
    pts = np.asarray(sample["points"], dtype=np.float64) * np.array([input_size[1], input_size[0]])
    occ = np.asarray(sample["occluded"]).astype(bool)
    if query_mode == "first":
        valid = (~occ).sum(axis=1) > 0
        pts, occ = pts[valid], occ[valid]
        first = np.argmax(~occ, axis=1)
        q = np.stack([first.astype(np.float64), pts[np.arange(len(pts)), first, 1], pts[np.arange(len(pts)), first, 0]], 1)
        tgt_pts, tgt_occ = pts, occ
    elif query_mode == "strided":
        qs, tp, to = [], [], []
        for t in range(0, occ.shape[1], stride):
            m = ~occ[:, t]
            qs.append(np.stack([np.full(m.sum(), t, dtype=np.float64), pts[m, t, 1], pts[m, t, 0]], 1))
            tp.append(pts[m])
            to.append(occ[m])
        q, tgt_pts, tgt_occ = np.concatenate(qs), np.concatenate(tp), np.concatenate(to)
    else:
        raise ValueError(f"Unknown query mode {query_mode}.")
    rgbs = torch.from_numpy(frames)[None].permute(0, 1, 4, 2, 3).contiguous()
    query_points = torch.from_numpy(q)[None][:, :, [0, 2, 1]].float()            # (t, y, x) -> (t, x, y)
    trajectories = torch.from_numpy(tgt_pts)[None].permute(0, 2, 1, 3).float()
    visibilities = ~torch.from_numpy(tgt_occ)[None].permute(0, 2, 1)
    return dict(rgbs=rgbs, query_points=query_points, trajectories=trajectories, visibilities=visibilities)
