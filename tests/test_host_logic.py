"""CPU-only checks: the C-ABI library loads and exports every declared symbol, host-side index maps, sharding."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import tta as otta
from pistoseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pistoseg_b200.h")).read()
    declared = set(re.findall(r"\b(pisto_[a-z0-9_]+)\s*\(", header))
    declared -= {"pisto_ctx"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in pistoseg_b200/_lib.py"
    assert lib.pisto_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.View) == 32
    assert ctypes.sizeof(_lib.FuseArgs) == 12 * 4 + 8 * 8
    assert ctypes.sizeof(_lib.TilePos) == 16
    assert ctypes.sizeof(_lib.MosaicQuad) == 64
    assert ctypes.sizeof(_lib.MosaicPlan) == 16 + 4 * 64
    assert ctypes.sizeof(_lib.MosaicCell) == 8


def test_no_cpu_fallback_without_device():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pistoseg_b200 import ops
    with pytest.raises(_lib.PistoError):
        ops.fuse_argmax_confusion([torch.zeros((1, 3, 28, 28))], [0], (224, 224))
    with pytest.raises(_lib.PistoError):
        _lib.handle(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pistoseg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
    for f in ("loss.py", "infer_pseudo_masks.py", "segmentation_test.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), re.M), f"{f} imports the oracle"


def test_d4_codes_and_affine_maps():
    assert [otta.deaug_code(h, a) for h, a in otta.D4_VIEWS] == [0, 3, 2, 1, 4, 7, 6, 5]
    for (h, w) in [(5, 7), (4, 4), (3, 8)]:
        y = torch.arange(h * w).reshape(1, 1, h, w)
        for code in range(8):
            d = otta.apply_code(y, code)[0, 0].numpy()
            ho, wo, base, si, sj = otta.code_affine(code, h, w)
            ii, jj = np.meshgrid(np.arange(ho), np.arange(wo), indexing="ij")
            assert d.shape == (ho, wo) and (d == base + ii * si + jj * sj).all()
    x = torch.randn(2, 3, 6, 6)
    for hf, ang in otta.D4_VIEWS:
        assert torch.equal(otta.apply_code(otta.augment(x, hf, ang), otta.deaug_code(hf, ang)), x)
