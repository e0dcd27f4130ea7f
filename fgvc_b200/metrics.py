"""TAP-Vid metrics on the device (SURVEY section 8f "next": the step right after the path).

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
