"""PyTorch feature encoder used by the tracker (OUT OF SCOPE for the CUDA work -- the
north star keeps the ResNet in PyTorch; it is here so the drop-in tracker runs without
mmcv).  Architecture contract of the reference backbone cfg
``dict(type='ResNet', depth=18, strides=..., out_indices=..., pool_type=...)``
(mmpt/models/backbones/resnet.py:329-640): 7x7/2 stem conv-bn-relu, optional 3x3/2
pooling, four stages of basic blocks with per-stage stride, features taken after the
stages in ``out_indices``.  Parameter names follow the reference's (``conv1.conv.weight``,
``layer3.0.downsample.bn.bias`` ...) so a released checkpoint loads with
``load_state_dict``.
"""
import torch
import torch.nn as nn

_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3)}


class ConvBN(nn.Module):
    def __init__(self, cin, cout, k, stride=1, padding=0, dilation=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, padding, dilation, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.act = act

    def forward(self, x):
        x = self.bn(self.conv(x))
        return torch.relu_(x) if self.act else x


class Basic(nn.Module):
    def __init__(self, cin, planes, stride, dilation, downsample):
        super().__init__()
        self.conv1 = ConvBN(cin, planes, 3, stride, dilation, dilation)
        self.conv2 = ConvBN(planes, planes, 3, 1, 1, 1, act=False)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        return torch.relu_(self.conv2(self.conv1(x)) + idt)


class ResNetEncoder(nn.Module):
    def __init__(self, depth=18, in_channels=3, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(3,),
                 pool_type="max", zero_init_residual=True, **unused):
        super().__init__()
        if depth not in _BLOCKS:
            raise NotImplementedError(f"ResNetEncoder supports depth 18/34, got {depth}")
        self.out_indices = tuple(out_indices)
        self.conv1 = ConvBN(in_channels, 64, 7, 2, 3)
        self.pool = {"max": nn.MaxPool2d(3, 2, 1), "mean": nn.AvgPool2d(3, 2, 1)}.get(pool_type)
        cin = 64
        self.res_layers = []
        for i, n in enumerate(_BLOCKS[depth]):
            planes = 64 * 2 ** i
            s, d = strides[i], dilations[i]
            ds = ConvBN(cin, planes, 1, s, act=False) if (s != 1 or cin != planes) else None
            blocks = [Basic(cin, planes, s, d if d == 1 else d // 2, ds)]
            blocks += [Basic(planes, planes, 1, d, None) for _ in range(1, n)]
            self.add_module(f"layer{i + 1}", nn.Sequential(*blocks))
            self.res_layers.append(f"layer{i + 1}")
            cin = planes
        self.zero_init_residual = zero_init_residual
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, Basic):
                    nn.init.constant_(m.conv2.bn.weight, 0)

    def forward(self, x):
        x = self.conv1(x)
        if self.pool is not None:
            x = self.pool(x)
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        return outs[0] if len(outs) == 1 else tuple(outs)


def build_backbone(cfg):
    """cfg: nn.Module, callable, or the reference's backbone dict (type='ResNet')."""
    if isinstance(cfg, nn.Module) or callable(cfg):
        return cfg
    cfg = dict(cfg)
    typ = cfg.pop("type", "ResNet")
    if typ != "ResNet":
        try:  # inside an mmpt installation: defer to its registry
            from mmpt.models.builder import build_backbone as _bb
            return _bb(dict(cfg, type=typ))
        except ImportError as e:
            raise NotImplementedError(f"backbone type {typ!r} needs mmpt") from e
    return ResNetEncoder(**cfg)
