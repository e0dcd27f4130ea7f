"""Seeded synthetic inputs shaped like the reference's evaluation data (there is no
network for datasets or checkpoints): temporally coherent video, random-init ResNet-18
features, Voronoi object masks, query points.  SURVEY.md section 8d."""
import torch

from .encoder import ResNetEncoder


def synthetic_video(T, h, w, seed=1000, device="cpu", drift=(1.0, 2.0)):
    """[T,3,h,w]: frame 0 = up-sampled low-res noise; frame t = frame 0 translated by a
    sub-pixel drift (bilinear) + 0.02 * noise."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(1, 3, h // 8 + 2, w // 8 + 2, generator=g)
    base = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    base = base + 0.1 * torch.randn(1, 3, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    frames = []
    for t in range(T):
        gx = xs - 2.0 * drift[1] * t / w
        gy = ys - 2.0 * drift[0] * t / h
        grid = torch.stack([gx, gy], dim=-1)[None]
        f = torch.nn.functional.grid_sample(base, grid, mode="bilinear", padding_mode="reflection",
                                            align_corners=False)[0]
        frames.append(f + 0.02 * torch.randn(3, h, w, generator=g))
    return torch.stack(frames).to(device)


def davis_encoder(stride=8, seed=0):
    """Random-init ResNet-18 with the reference's cfg: stride 8 = strides (1,2,2,1) without
    pooling (BASELINE configs 1-2); stride 2 = the shipped eval cfg strides (1,1,1,4)."""
    torch.manual_seed(seed)
    strides = {8: (1, 2, 2, 1), 2: (1, 1, 1, 4)}[stride]
    return ResNetEncoder(depth=18, strides=strides, out_indices=(2,), pool_type="none").eval()


@torch.no_grad()
def encode(encoder, frames, batch=8):
    out = []
    for s in range(0, frames.shape[0], batch):
        out.append(encoder(frames[s:s + batch]).float())
    return torch.cat(out)


def voronoi_mask(H, W, L, seed=0):
    """int64 [H,W] with labels 0..L-1: nearest of L seeded sites."""
    g = torch.Generator().manual_seed(seed)
    sy = torch.rand(L, generator=g) * H
    sx = torch.rand(L, generator=g) * W
    ys = torch.arange(H).view(-1, 1, 1).float()
    xs = torch.arange(W).view(1, -1, 1).float()
    d = (ys - sy.view(1, 1, -1)) ** 2 + (xs - sx.view(1, 1, -1)) ** 2
    return d.argmin(dim=-1)


def query_points(P, T, h, w, seed=0, first_frame_only=True):
    """[P,3] (t,x,y)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(P, generator=g) * (w - 1)
    y = torch.rand(P, generator=g) * (h - 1)
    t = torch.zeros(P) if first_frame_only else torch.randint(0, max(1, T // 2), (P,), generator=g).float()
    return torch.stack([t, x, y], dim=1)
