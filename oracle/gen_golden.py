"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by executing the GENUINE
reference (loaded from /root/reference by oracle/ref_loader.py) on seeded inputs.

Run in the build container:  python -m oracle.gen_golden
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import warnings

import numpy as np
import torch

from . import ref_loader
from .inputs import coherent_feats as _coherent_feats, seeded_cfg3_inputs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_propagate(ref):
    cases = []
    specs = [
        # name, H, W, C, T, L, neighbor_range, topk, non_mask_len, coherent
        ("rand_small", 12, 14, 32, 3, 5, 8, 5, 0, False),
        ("rand_nonmask1", 10, 9, 32, 4, 3, 6, 10, 1, False),
        ("coh_dupframe0", 16, 20, 64, 4, 6, 10, 10, 0, True),
        ("coh_c256", 15, 27, 256, 6, 8, 24, 10, 0, True),
        ("tiny_fewcands", 5, 6, 32, 1, 2, 2, 10, 0, False),   # r=1: 1 in-mask key < k
    ]
    for i, (name, H, W, C, T, L, nr, k, nml, coh) in enumerate(specs):
        g = torch.Generator().manual_seed(100 + i)
        if coh:
            f = _coherent_feats(g, T + 1, C, H, W)
            q, kf = f[T][None], f[:T].permute(1, 0, 2, 3)[None].contiguous()
            if name == "coh_dupframe0":   # memory = [0, 0, 1, 2]: frame 0 twice
                kf = torch.cat([kf[:, :, :1], kf[:, :, :T - 1]], dim=2)
        else:
            q = torch.randn(1, C, H, W, generator=g)
            kf = torch.randn(1, C, T, H, W, generator=g)
        v = torch.rand(1, L, T, H, W, generator=g)
        if name == "coh_dupframe0":
            v = torch.cat([v[:, :, :1], v[:, :, :T - 1]], dim=2)
        mask = ref.spatial_neighbor(1, H, W, neighbor_range=nr, device="cpu", dtype=torch.float32)
        o1 = ref.masked_attention_efficient(q, kf, v, mask, temperature=0.07, topk=k, step=64,
                                            non_mask_len=nml)
        o2 = ref.masked_attention_efficient_v2(q, kf, v, nr // 2, temperature=0.07, topk=k, step=64)
        cases.append(name)
        np.savez_compressed(os.path.join(OUT, f"prop_{name}.npz"), q=q.numpy(), k=kf.numpy(),
                            v=v.numpy(), neighbor_range=nr, topk=k, non_mask_len=nml,
                            temperature=0.07, out_v1=o1.numpy(), out_v2=o2.numpy())
    return cases


def gen_cfg3_geometry(ref):
    """One frame at the reference eval geometry of BASELINE configs 3 / 5 (256^2 image, stride 2: 128 x 128,
    neighbor_range 30, 6 memory entries, C = 256) through the genuine masked_attention_efficient
    (local_attention.py:267).  Only the output and input checksums are stored."""
    q, kf, v = seeded_cfg3_inputs()
    mask = ref.spatial_neighbor(1, 128, 128, neighbor_range=30, device="cpu", dtype=torch.float32)
    o = ref.masked_attention_efficient(q, kf, v, mask, temperature=0.07, topk=10, step=512)
    np.savez_compressed(os.path.join(OUT, "prop_cfg3_128.npz"), seed=303, neighbor_range=30, topk=10,
                        temperature=0.07, q_sum=float(q.double().sum()), k_sum=float(kf.double().sum()),
                        v_sum=float(v.double().sum()), out=o.numpy())


def gen_dense(ref):
    """topk=None (dense soft-max / clamp^2 over all allowed candidates, local_attention.py:376-383)."""
    g = torch.Generator().manual_seed(311)
    H, W, C, T, L, nr = 11, 13, 32, 3, 5, 8
    f = _coherent_feats(g, T + 1, C, H, W)
    q, kf = f[T][None], f[:T].permute(1, 0, 2, 3)[None].contiguous()
    v = torch.rand(1, L, T, H, W, generator=g)
    mask = ref.spatial_neighbor(1, H, W, neighbor_range=nr, device="cpu", dtype=torch.float32)
    d = dict(q=q.numpy(), k=kf.numpy(), v=v.numpy(), neighbor_range=nr, temperature=0.07)
    d["softmax"] = ref.masked_attention_efficient(q, kf, v, mask, temperature=0.07, topk=None, step=64).numpy()
    d["softmax_nonmask1"] = ref.masked_attention_efficient(q, kf, v, mask, temperature=0.07, topk=None, step=64,
                                                           non_mask_len=1).numpy()
    d["cosine"] = ref.masked_attention_efficient(q, kf, v, mask, temperature=0.07, topk=None, step=64,
                                                 mode="cosine").numpy()
    d["l2"] = ref.masked_attention_efficient(q, kf, v, mask, temperature=0.07, topk=None, step=64,
                                             sim_mode="l2-distance").numpy()
    d["nomask"] = ref.masked_attention_efficient(q, kf, v, None, temperature=0.07, topk=None, step=64).numpy()
    np.savez_compressed(os.path.join(OUT, "prop_dense.npz"), **d)


def gen_masks(ref):
    d = {}
    for (H, W, nr) in [(6, 7, 4), (9, 5, 7), (12, 14, 8)]:
        d[f"circle_{H}_{W}_{nr}"] = ref.spatial_neighbor(1, H, W, nr, "cpu", torch.float32).numpy()
        d[f"square_{H}_{W}_{nr}"] = ref.spatial_neighbor(1, H, W, nr, "cpu", torch.float32,
                                                         mode="square")[0].bool().numpy()
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **d)


def gen_c2f(ref):
    g = torch.Generator().manual_seed(7)
    Hc, Wc, s, T, C, Cf, L, rf = 6, 7, 4, 3, 32, 16, 4, 3
    f = _coherent_feats(g, T + 1, C, Hc, Wc)
    ff = _coherent_feats(g, T + 1, Cf, Hc * s, Wc * s)
    q, k = f[T][None], f[:T].permute(1, 0, 2, 3)[None].contiguous()
    qf, kf = ff[T][None], ff[:T].permute(1, 0, 2, 3)[None].contiguous()
    v = torch.rand(1, L, T, Hc * s, Wc * s, generator=g)
    mask = ref.spatial_neighbor(1, Hc, Wc, neighbor_range=6, device="cpu", dtype=torch.float32)
    o = ref.masked_attention_efficient_c2f(q, k, qf, kf, v, mask, temperature=0.07, topk=10,
                                           step=16, radius_fine=rf)
    np.savez_compressed(os.path.join(OUT, "c2f_small.npz"), q=q.numpy(), k=k.numpy(),
                        qf=qf.numpy(), kf=kf.numpy(), v=v.numpy(), neighbor_range=6, topk=10,
                        temperature=0.07, radius_fine=rf, out=o.numpy())


def gen_c2f_driver(ref):
    """The coarse-to-fine clip loop (oracle.track_clip_c2f_port) around the GENUINE masked_attention_efficient_c2f."""
    from . import oracle as O
    g = torch.Generator().manual_seed(17)
    T, C, Cf, Hc, Wc, s = 5, 32, 64, 8, 10, 4
    fc = _coherent_feats(g, T, C, Hc, Wc)
    ff = _coherent_feats(g, T, Cf, Hc * s, Wc * s)
    h, w = Hc * s * 2, Wc * s * 2                       # fine stride 2, coarse stride 8
    pts = torch.tensor([[20.5, 17.25], [50.0, 40.0], [9.75, 55.5], [70.25, 8.0]])
    cfg = dict(precede_frames=2, topk=10, temperature=0.07, neighbor_range=8, radius_fine=4, with_first=True, step=16)
    mask = ref.spatial_neighbor(1, Hc, Wc, neighbor_range=8, device="cpu", dtype=torch.float32)
    outs, traj = O.track_clip_c2f_port(fc, ff, pts, (h, w), cfg, c2f=ref.masked_attention_efficient_c2f, mask=mask)
    np.savez_compressed(os.path.join(OUT, "c2f_driver.npz"), feats_c=fc.numpy(), feats_f=ff.numpy(), points=pts.numpy(),
                        image_hw=np.array([h, w]), precede_frames=2, topk=10, temperature=0.07, neighbor_range=8,
                        radius_fine=4, outs=torch.stack(outs).numpy(), traj=traj)


def gen_legacy(ref):
    g = torch.Generator().manual_seed(11)
    a = torch.randn(2, 16, 5, 6, generator=g)
    b = torch.randn(2, 16, 5, 6, generator=g)
    img = torch.rand(2, 3, 5, 6, generator=g)
    aff = ref.compute_affinity(a, b, temperature=0.07, softmax_dim=1)
    prop = ref.propagate(img, aff.clone(), topk=4)
    np.savez_compressed(os.path.join(OUT, "legacy.npz"), a=a.numpy(), b=b.numpy(), img=img.numpy(),
                        aff=aff.numpy(), prop=prop.numpy())


def gen_tracker():
    """Genuine VanillaTracker.forward_test + genuine random-init ResNet-18 (seed 0)."""
    trk = ref_loader.load_tracker()
    out = {}
    for tag, strides, hw, nr in [("s8", (1, 2, 2, 1), (64, 96), 8), ("s2", (1, 1, 1, 4), (32, 40), 12)]:
        cfg = trk.TestCfg(precede_frames=2, topk=10, temperature=0.07, neighbor_range=nr, step=64,
                          with_first=True, with_first_neighbor=True)
        torch.manual_seed(0)
        model = trk.VanillaTracker(backbone=dict(type="ResNet", depth=18, strides=strides,
                                                 out_indices=(2,), pool_type="none"),
                                   train_cfg=None, test_cfg=cfg)
        model.init_weights()
        model.eval()
        g = torch.Generator().manual_seed(5)
        T = 5
        base = torch.nn.functional.interpolate(torch.randn(1, 3, hw[0] // 4, hw[1] // 4, generator=g),
                                               size=hw, mode="bilinear", align_corners=False)[0]
        frames = [torch.roll(base, shifts=(t, 2 * t), dims=(1, 2)) + 0.02 * torch.randn(base.shape, generator=g)
                  for t in range(T)]
        rgbs = torch.stack(frames)[None]
        qp = torch.tensor([[[0, 10., 12.], [0, 30., 20.], [1, 22., 9.], [0, 5., 25.], [2, 17., 17.]]])
        P = qp.shape[1]
        with torch.no_grad():
            res = model(test_mode=True, rgbs=rgbs, query_points=qp,
                        trajectories=torch.zeros(1, T, P, 2), visibilities=torch.zeros(1, T, P))
            feats = model.extract_feat_test(rgbs[0])
        out[f"{tag}_rgbs"] = rgbs.numpy()
        out[f"{tag}_feats"] = feats.numpy().astype(np.float32)
        out[f"{tag}_query_points"] = qp.numpy()
        out[f"{tag}_traj_pred"] = res[2].numpy()
        out[f"{tag}_query_points_remap"] = res[4].numpy()
        out[f"{tag}_neighbor_range"] = nr
        out[f"{tag}_precede_frames"] = 2
        # img2coord on a propagated-looking map
    maps = np.random.RandomState(3).rand(3, 4, 9, 11).astype(np.float32)
    maps[1, 2] = 0
    out["i2c_maps"] = maps
    out["i2c_xy"] = trk.VanillaTracker.img2coord(None, maps, num_poses=4)
    np.savez_compressed(os.path.join(OUT, "tracker.npz"), **out)


def gen_tapvid_metrics():
    fn = ref_loader.load_tapvid_metrics()
    rs = np.random.RandomState(5)
    b, n, T = 2, 9, 12
    qp = np.stack([rs.randint(0, 5, (b, n)).astype(np.float64), rs.rand(b, n) * 256, rs.rand(b, n) * 256], -1)
    gt_occ = rs.rand(b, n, T) < 0.25
    for i in range(b):
        for j in range(n):
            gt_occ[i, j, int(qp[i, j, 0])] = False
    gt = rs.rand(b, n, T, 2) * 256
    pred = gt + rs.randn(b, n, T, 2) * rs.choice([0.5, 3.0, 12.0], (b, n, T, 1))
    pred_occ = gt_occ ^ (rs.rand(b, n, T) < 0.2)
    d = dict(qp=qp, gt_occ=gt_occ, gt=gt, pred_occ=pred_occ, pred=pred)
    for mode in ("first", "strided"):
        for k, v in fn(qp, gt_occ, gt, pred_occ, pred, mode).items():
            d[f"{mode}__{k}"] = v
    np.savez_compressed(os.path.join(OUT, "tapvid_metrics.npz"), **d)


def gen_eval_metrics():
    """DAVIS J & F, JHMDB PCK and the TAP-Vid query packaging from the genuine functions."""
    rs = np.random.RandomState(21)
    m = ref_loader.load_davis_metrics()
    T, H, W = 9, 48, 64
    gt, res = np.zeros((3, T, H, W), bool), np.zeros((2, T, H, W), bool)          # the 3rd object has no result
    for o in range(3):
        for t in range(T):
            y, x = 5 + 2 * t + 7 * o, 6 + 3 * t + 2 * o
            gt[o, t, y:y + 12, x:x + 16] = True
            if o < 2:
                d = t // 2                                                        # the result drifts: decay > 0
                res[o, t, y + d:y + 12 + d, x - d:x + 16 - d + o] = True
    res[1, 4] = False                                                             # an empty prediction
    out = m.JFM(gt.copy(), res.copy(), 3)
    d = dict(jf_gt=gt, jf_res=res)
    for k, v in out.items():
        d[f"jf_{k}"] = np.asarray(v, dtype=np.float64)
    d["jf_iou_obj0"] = m.db_eval_iou(gt[0].copy(), res[0].copy())
    d["jf_f_obj0"] = m.db_eval_boundary(gt[0].copy(), res[0].copy())
    cp, dist = ref_loader.load_jhmdb_pck()
    P, Tp = 15, 11
    gtp = rs.rand(2, P, Tp) * 200 + 20
    pred = gtp + rs.randn(2, P, Tp) * rs.choice([2.0, 10.0, 40.0], (1, P, Tp))
    pred[0, 3, 2] = -1                                                            # an invisible prediction
    pred[:, 7, :] = -1
    pred[:, 7, 5] = gtp[:, 7, 5] + 1.0
    dd = dist(pred.copy(), gtp.copy(), P)
    d.update(pck_pred=pred, pck_gt=gtp, pck_dist=np.array([np.asarray(a).ravel().sum() for a in dd]),
             pck_count=np.array([np.asarray(a).size for a in dd]),
             pck_values=np.array([np.mean(cp(dd, a)) for a in (0.1, 0.2, 0.3, 0.4, 0.5)]))
    first, strided = ref_loader.load_tapvid_query_samplers()
    N, Tq, h, w = 7, 12, 16, 20
    occ = rs.rand(N, Tq) < 0.4
    occ[2] = True                                                                 # a track that is never visible
    pts = rs.rand(N, Tq, 2)
    video = rs.randint(0, 256, (Tq, h, w, 3)).astype(np.uint8)
    frames = video.astype(np.float32) / 255.0 * 2 - 1
    pix = pts * np.array([w, h])
    a, b = first(occ.copy(), pix.copy(), frames), strided(occ.copy(), pix.copy(), frames, query_stride=5)
    d.update(tv_video=video, tv_points=pts, tv_occluded=occ, tv_hw=np.array([h, w]))
    for tag, r in (("first", a), ("strided", b)):
        d[f"tv_{tag}_query_points"] = r["query_points"]
        d[f"tv_{tag}_target_points"] = r["target_points"]
        d[f"tv_{tag}_occluded"] = r["occluded"]
    np.savez_compressed(os.path.join(OUT, "eval_metrics.npz"), **d)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_loader.load_functions()
    print("propagate:", gen_propagate(ref))
    gen_dense(ref)
    gen_cfg3_geometry(ref)
    gen_masks(ref)
    gen_c2f(ref)
    gen_c2f_driver(ref)
    gen_legacy(ref)
    gen_tracker()
    gen_tapvid_metrics()
    gen_eval_metrics()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
