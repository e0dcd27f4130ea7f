"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by the golden-fixture generator (oracle/gen_golden.py,
run in the build container next to the genuine reference) and the tests (run anywhere).  Nothing here touches
/root/reference."""
import torch


def coherent_feats(g, T, C, H, W, relu=True):
    """temporally coherent, spatially smooth features (near-ties like encoder output)."""
    base = torch.randn(C, H // 2 + 2, W // 2 + 2, generator=g)
    frames = []
    for t in range(T):
        base = base + 0.15 * torch.randn(base.shape, generator=g)
        f = torch.nn.functional.interpolate(base[None], size=(H, W), mode="bilinear",
                                            align_corners=False)[0]
        f = f + 0.05 * torch.randn(f.shape, generator=g)
        frames.append(f.relu() if relu else f)
    return torch.stack(frames)  # [T,C,H,W]


def seeded_cfg3_inputs(seed=303, H=128, W=128, C=256, L=4):
    """Inputs of the config-3 geometry fixture, regenerated from the seed on both sides (the tensors themselves
    are 100 MB): memory = [0, 0, 1, 2, 3, 4] as at t = 5 with precede_frames = 5 (vanilla_tracker.py:346-362)."""
    g = torch.Generator().manual_seed(seed)
    f = coherent_feats(g, 6, C, H, W)
    mem = [0, 0, 1, 2, 3, 4]
    q, kf = f[5][None], f[mem].permute(1, 0, 2, 3)[None].contiguous()
    v = torch.rand(1, L, 5, H, W, generator=g)[:, :, mem].contiguous()
    return q, kf, v
