"""The C-ABI library loads and exports every symbol include/fgvc_b200.h declares (CPU only:
no compute call is made)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fgvc_b200.h")).read()
    return re.findall(r"^FGVC_API\s+[\w\s\*]+?\b(fgvc_\w+)\s*\(", src, flags=re.M)


def test_header_declares_the_path():
    names = _declared()
    for must in ("fgvc_prep_features", "fgvc_affinity_topk", "fgvc_gather_labels", "fgvc_heatmap_coords",
                 "fgvc_c2f_propagate", "fgvc_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from fgvc_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in fgvc_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == sorted(_declared())
    assert lib.fgvc_version() == 100
    assert lib.fgvc_topk_bytes(3, 2, 100, 10) == 3 * 2 * 100 * 10 * 4


def test_errors_are_reported_not_thrown():
    from fgvc_b200 import _lib
    lib = _lib.load()
    # null pointers are rejected before any CUDA call
    rc = lib.fgvc_gather_labels(None, None, 10, 1, None, 0, 1, None, 16, 0.07, 0, None, 4, None)
    assert rc == -1
    assert b"null pointer" in lib.fgvc_last_error()
    with pytest.raises(_lib.FgvcError):
        _lib.check(rc)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import fgvc_b200
    q = torch.randn(1, 32, 8, 8)
    k = torch.randn(1, 32, 2, 8, 8)
    v = torch.rand(1, 3, 2, 8, 8)
    with pytest.raises(fgvc_b200.FgvcError):
        fgvc_b200.masked_attention_efficient_v2(q, k, v, 3, temperature=0.07, topk=5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fgvc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn
