"""TEST INFRASTRUCTURE ONLY -- loads the *genuine* FGVC reference modules from
/root/reference by file path so the oracle restatement (oracle/oracle.py) and the
golden fixtures (tests/golden/) can be pinned to what the reference itself computes.

Nothing here is shipped or measured: only ``tests/``, ``oracle/gen_golden.py`` and
the validation scripts import it, and only inside the build container
(``/root/reference`` does not exist on the GPU box).  No reference source is copied:
the files are executed where they lie.

The reference depends on mmcv-full 1.5.2 (not installable offline).  Two levels:

* ``load_functions()`` -- ``mmpt/models/common/{affinity_utils,local_attention}.py``
  only need ``torch`` plus one ``from mmpt.models.common import part_unfold`` line;
  three empty stub packages satisfy it.
* ``load_tracker()`` -- ``VanillaTracker`` + ``ResNet`` additionally need a handful of
  mmcv symbols (BaseModule, Registry, ConvModule, build_from_cfg ...); a small
  in-memory stand-in for those is registered in ``sys.modules`` (it re-implements
  only mmcv plumbing, none of the reference's arithmetic).
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("FGVC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "mmpt/models/common/local_attention.py"))


def _stub(name, is_pkg=True):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    if is_pkg:
        m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


def _load(modname, relpath):
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    parent, _, child = modname.rpartition(".")
    if parent:
        setattr(_stub(parent), child, mod)
    spec.loader.exec_module(mod)
    return mod


def _export(src, dst):
    names = getattr(src, "__all__", None) or [n for n in vars(src) if not n.startswith("_")]
    for n in names:
        setattr(dst, n, getattr(src, n))


_FUNCS = None


def load_functions():
    """Genuine ``masked_attention_efficient{,_v2,_c2f}``, ``spatial_neighbor`` etc."""
    global _FUNCS
    if _FUNCS is not None:
        return _FUNCS
    assert available(), f"reference not found under {REF_ROOT}"
    _stub("mmpt")
    _stub("mmpt.models")
    common = _stub("mmpt.models.common")
    if not hasattr(common, "part_unfold"):
        common.part_unfold = None  # only the *_correlation* variants touch it
    au = _load("mmpt.models.common.affinity_utils", "mmpt/models/common/affinity_utils.py")
    la = _load("mmpt.models.common.local_attention", "mmpt/models/common/local_attention.py")
    ns = types.SimpleNamespace(
        masked_attention_efficient=la.masked_attention_efficient,
        masked_attention_efficient_v2=la.masked_attention_efficient_v2,
        masked_attention_efficient_c2f=la.masked_attention_efficient_c2f,
        spatial_neighbor=au.spatial_neighbor,
        compute_affinity=au.compute_affinity,
        propagate=au.propagate,
        propagate_temporal=au.propagate_temporal,
        coords_grid=la.coords_grid,
        local_attention=la,
        affinity_utils=au,
    )
    _FUNCS = ns
    return ns


# --------------------------------------------------------------------------- mmcv stand-in
class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()


class _ConvModule(nn.Module):
    """conv -> norm -> act with the attribute names the reference ResNet touches."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias="auto", conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type="ReLU"), inplace=True, **kw):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding,
                              dilation, groups, bias=bias)
        if self.with_norm:
            self.bn = nn.BatchNorm2d(out_channels)
            for p in self.bn.parameters():
                p.requires_grad = norm_cfg.get("requires_grad", True)
        if self.with_activation:
            self.activate = nn.ReLU(inplace=act_cfg.get("inplace", inplace))

    @property
    def norm(self):
        return self.bn if self.with_norm else None

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.bn(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


def _kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    if hasattr(module, "weight") and module.weight is not None:
        if distribution == "uniform":
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _install_mmcv_standin():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_fgvc_standin", False):
        return
    mmcv = _stub("mmcv")
    mmcv._fgvc_standin = True
    mmcv.build_from_cfg = _build_from_cfg
    mmcv.mkdir_or_exist = lambda d, mode=0o777: os.makedirs(d, mode=mode, exist_ok=True)

    runner = _stub("mmcv.runner")
    runner.BaseModule = _BaseModule

    def auto_fp16(*a, **k):
        def deco(fn):
            return fn
        return deco
    runner.auto_fp16 = auto_fp16
    runner._load_checkpoint = lambda *a, **k: {}
    runner.load_checkpoint = lambda *a, **k: {}
    runner.get_dist_info = lambda: (0, 1)

    utils = _stub("mmcv.utils")
    utils.Registry = _Registry
    utils._BatchNorm = nn.modules.batchnorm._BatchNorm
    utils.get_logger = lambda name, log_file=None, log_level=None: __import__("logging").getLogger(name)

    cnn = _stub("mmcv.cnn")
    cnn.ConvModule = _ConvModule
    cnn.kaiming_init = _kaiming_init
    cnn.constant_init = _constant_init


class TestCfg(dict):
    """dict with attribute access + .get, like mmcv.ConfigDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


_TRACKER = None


def load_tracker():
    """Genuine ``VanillaTracker`` and ``ResNet`` classes behind the mmcv stand-in."""
    global _TRACKER
    if _TRACKER is not None:
        return _TRACKER
    assert available(), f"reference not found under {REF_ROOT}"
    _install_mmcv_standin()
    load_functions()
    import logging
    _stub("mmpt")
    u = _stub("mmpt.utils")
    u.get_root_logger = lambda *a, **k: logging.getLogger("mmpt")
    u.__all__ = ["get_root_logger"]
    _stub("mmpt.models")
    common = _stub("mmpt.models.common")
    bb = _stub("mmpt.models.backbones")
    _stub("mmpt.models.trackers")
    _load("mmpt.models.registry", "mmpt/models/registry.py")
    _load("mmpt.models.builder", "mmpt/models/builder.py")
    for name in ("corr_lookup", "part_unfold", "utils", "affinity_utils", "local_attention",
                 "correlation"):
        if name == "part_unfold":
            # drop the placeholder left by load_functions() so the real module binds
            if getattr(common, "part_unfold", None) is None and hasattr(common, "part_unfold"):
                delattr(common, "part_unfold")
        m = _load(f"mmpt.models.common.{name}", f"mmpt/models/common/{name}.py")
        _export(m, common)
    rn = _load("mmpt.models.backbones.resnet", "mmpt/models/backbones/resnet.py")
    bb.ResNet = rn.ResNet
    _load("mmpt.models.trackers.base", "mmpt/models/trackers/base.py")
    vt = _load("mmpt.models.trackers.vanilla_tracker", "mmpt/models/trackers/vanilla_tracker.py")
    if not torch.cuda.is_available():
        # the reference hard-codes .cuda(); on a CPU-only box make it a no-op
        torch.Tensor.cuda = lambda self, *a, **k: self
    _TRACKER = types.SimpleNamespace(VanillaTracker=vt.VanillaTracker, ResNet=rn.ResNet,
                                     TestCfg=TestCfg, module=vt)
    return _TRACKER


def load_tapvid_metrics():
    """Genuine ``compute_tapvid_metrics`` (mmpt/datasets/tapvid_evaluation_datasets.py:106-249).  The
    module imports tensorflow-era dependencies at the top, so only the function definition is taken
    from the file (by ast) and executed with numpy in scope; nothing is copied into the repo."""
    import ast
    from typing import Iterable, Mapping
    import numpy as np
    path = os.path.join(REF_ROOT, "mmpt/datasets/tapvid_evaluation_datasets.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_tapvid_metrics"][0]
    ns = dict(np=np, Iterable=Iterable, Mapping=Mapping)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["compute_tapvid_metrics"]


def load_davis_metrics():
    """Genuine DAVIS J & F functions (mmpt/core/evaluation/metrics.py).  The file imports cv2 (present) and mmcv
    (stand-in) and uses numpy aliases / skimage.morphology.disk that this container lacks: ``np.bool`` is re-created
    and a two-line ``disk`` (pixels with dy^2 + dx^2 <= r^2, skimage's definition) is provided as skimage.morphology."""
    import types
    import numpy as np
    if not hasattr(np, "bool"):
        np.bool = bool
    _install_mmcv_standin()
    if "skimage" not in sys.modules:
        sk, mo = types.ModuleType("skimage"), types.ModuleType("skimage.morphology")

        def disk(radius):
            r = int(radius)
            y, x = np.mgrid[-r:r + 1, -r:r + 1]
            return (y * y + x * x <= r * r).astype(np.uint8)
        mo.disk = disk
        sk.morphology = mo
        sys.modules["skimage"], sys.modules["skimage.morphology"] = sk, mo
    return _load("mmpt_ref_davis_metrics", "mmpt/core/evaluation/metrics.py")


def load_jhmdb_pck():
    """Genuine ``JhmdbVideoDataset.compute_pck`` and the distance loop of ``pck_evaluate``
    (mmpt/datasets/jhmdb_dataset.py:143-233), taken by ast (the class needs the dataset registry): returns
    (compute_pck, distances(pred_poses, gt_poses, n_keypoints) -> list of per-joint arrays)."""
    import ast
    import numpy as np
    path = os.path.join(REF_ROOT, "mmpt/datasets/jhmdb_dataset.py")
    src = open(path).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef)][0]
    fns = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}
    cp = fns["compute_pck"]
    cp.decorator_list = []
    ns = dict(np=np)
    exec(compile(ast.Module(body=[cp], type_ignores=[]), path, "exec"), ns)
    # the per-video distance loop: lines of pck_evaluate between 'joint_visible =' and the end of the double loop
    lines = src.splitlines()
    start = next(i for i, l in enumerate(lines) if "joint_visible = pred_poses[0] > 0" in l)
    end = next(i for i, l in enumerate(lines) if "dist_all[t] = np.append(dist_all[t], [[dist]])" in l)
    body = "\n".join(l[12:] for l in lines[start:end + 1])

    def distances(pred_poses, gt_poses, n_keypoints):
        class _S:
            NUM_KEYPOINTS = n_keypoints
        env = dict(np=np, self=_S, pred_poses=pred_poses, gt_poses=gt_poses, clip_len=gt_poses.shape[-1],
                   dist_all=[np.zeros((0, 0)) for _ in range(n_keypoints)])
        exec(body, env)
        return env["dist_all"]
    return ns["compute_pck"], distances


def load_tapvid_query_samplers():
    """Genuine ``sample_queries_first`` / ``sample_queries_strided`` (tapvid_evaluation_datasets.py:297-401), by ast."""
    import ast
    from typing import Iterable, Mapping, Optional
    import numpy as np
    path = os.path.join(REF_ROOT, "mmpt/datasets/tapvid_evaluation_datasets.py")
    tree = ast.parse(open(path).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("sample_queries_first",
                                                                               "sample_queries_strided")]
    ns = dict(np=np, Iterable=Iterable, Mapping=Mapping, Optional=Optional)
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return ns["sample_queries_first"], ns["sample_queries_strided"]
