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
    assert ctypes.sizeof(_lib.FuseArgs) == 12 * 4 + 9 * 8
    assert ctypes.sizeof(_lib.TilePos) == 16
    assert ctypes.sizeof(_lib.MosaicQuad) == 64
    assert ctypes.sizeof(_lib.MosaicPlan) == 16 + 4 * 64
    assert ctypes.sizeof(_lib.MosaicCell) == 8
    assert ctypes.sizeof(_lib.ResizeDesc) == 4 * 4 + 4 * 8


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
    for f in ("loss.py", "infer_pseudo_masks.py", "segmentation_test.py", "OEEM/classification/prepare_seg_inputs.py"):
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


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors published with Random123 (kat_vectors)."""
    from pistoseg_b200 import philox
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32_10(*c, *k)
        assert tuple(int(v) for v in got) == want
    # vectorised evaluation == element-wise evaluation
    import numpy as np
    c0 = np.arange(7, dtype=np.uint64)
    vec = philox.philox4x32_10(c0, 5, 6, 7, 8, 9)
    for i in range(7):
        one = philox.philox4x32_10(int(c0[i]), 5, 6, 7, 8, 9)
        assert all(int(a[i]) == int(b) for a, b in zip(vec, one))


def test_mosaic_planner_is_a_pure_function_of_seed_and_index():
    import numpy as np
    from pistoseg_b200 import mosaic
    rng = np.random.default_rng(5)
    sizes = [(224, 224)] * 6 + [(50, 60), (70, 40)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    bgs = [(rng.random((h, w)) < 0.5).astype(np.uint8) * 255 for h, w in sizes]
    pool = mosaic.TilePool(imgs, rng.integers(0, 3, len(sizes)).astype(np.uint8), bgs, device="cpu")
    pl = mosaic.MosaicPlanner(pool, 4, 56, seed=2022, reject_bg=True)
    p, c = pl.plans(range(12))
    p2, c2 = pl.plans([11, 3, 2 ** 33 + 5])
    assert p2[0].tobytes() == p[11].tobytes() and p2[1].tobytes() == p[3].tobytes()
    assert c2[0].tobytes() == c[11].tobytes() and c2[1].tobytes() == c[3].tobytes()
    one_p, one_c = pl.plan(3)
    assert one_p.tobytes() == p[3].tobytes() and one_c.tobytes() == c[3].tobytes()
    assert mosaic.MosaicPlanner(pool, 4, 56, seed=2023, reject_bg=True).plans([3])[1].tobytes() != c[3].tobytes()
    # ranges and invariants of the decisions
    assert ((p["split_h"] % 2 == 0) & (p["split_h"] >= 44) & (p["split_h"] <= 180)).all()
    assert (c["tile"] >= 0).all() and (c["tile"] < len(sizes)).all()
    hw = pool.hw_host[c["tile"]]
    assert (c["cy"] + 56 <= np.maximum(hw[..., 0], 56)).all() and (c["cx"] + 56 <= np.maximum(hw[..., 1], 56)).all()
    flips = p["quad"]["flip"]
    assert set(np.unique(flips)) <= {0, 1, 2, 3} and 0.6 < (flips > 0).mean() <= 1.0
    # accepted cells satisfy the background criterion (or exhausted the tries); rejection really happens on this pool
    ioff, I = pool.integral_host(56)
    t, y, x = c["tile"].astype(np.int64), c["cy"].astype(np.int64), c["cx"].astype(np.int64)
    W1 = np.maximum(pool.hw_host[t, 1], 56).astype(np.int64) + 1
    n = (I[ioff[t] + (y + 56) * W1 + x + 56].astype(np.int64) - I[ioff[t] + y * W1 + x + 56] - I[ioff[t] + (y + 56) * W1 + x] + I[ioff[t] + y * W1 + x]) & 0xFFFF
    brute = np.array([(np.asarray(mosaic._pad_reflect101((bgs[tt] > 0).astype(np.int64), 56))[yy:yy + 56, xx:xx + 56]).sum()
                      for tt, yy, xx in zip(t.reshape(-1)[:50], y.reshape(-1)[:50], x.reshape(-1)[:50])])
    assert np.array_equal(n.reshape(-1)[:50], brute)
    free = mosaic.MosaicPlanner(pool, 4, 56, seed=2022, reject_bg=False).cells_host(range(12))
    assert free.tobytes() != c.tobytes()
    # the vectorised affine equals the scalar restatement of cv2 / albumentations bit for bit
    M = mosaic.shift_scale_rotate_batch(224, 224, np.array([12.5, -33.0]), np.array([0.9, 1.15]), np.array([0.01, -0.05]), np.array([0.03, 0.0]))
    inv = mosaic.invert_affine_batch(M)
    for k, (a, s, dx, dy) in enumerate([(12.5, 0.9, 0.01, 0.03), (-33.0, 1.15, -0.05, 0.0)]):
        ref = mosaic.invert_affine(mosaic.shift_scale_rotate_matrix(224, 224, a, s, dx, dy)).reshape(-1)
        assert ref.tobytes() == inv[k].tobytes()


def test_division_by_32bit_inverse_is_exact_below_65536():
    """The kernels replace x / d by umulhi(x, ceil(2^32 / d)) (fuse_filter.cuh g_inv, fuse_fullres.cu fastdiv, mosaic.cu inv_gpr)
    and guard the call sites to x < 2^16: the identity must hold for every such x and every divisor they can meet."""
    import numpy as np
    x = np.arange(65536, dtype=np.uint64)
    for d in list(range(2, 300)) + [448, 512, 1000, 2047, 4096, 65535]:
        inv = np.uint64(((1 << 32) + d - 1) // d)
        assert np.array_equal((x * inv) >> np.uint64(32), x // np.uint64(d)), d


def test_fixed_cos_sin_of_the_planners_is_accurate():
    """mosaic.cos_sin_deg (the polynomial both planners evaluate instead of libm, so that host and device agree to the bit) is within
    5e-16 of numpy over the ShiftScaleRotate angle range and exact at the quadrant angles."""
    from pistoseg_b200 import mosaic
    ang = np.concatenate([np.linspace(-180, 180, 100001), np.array([0.0, 45.0, -45.0, 90.0, -90.0, 180.0, 44.999999, -0.0])])
    co, si = mosaic.cos_sin_deg(ang)
    assert np.max(np.abs(co - np.cos(np.deg2rad(ang)))) < 5e-16 and np.max(np.abs(si - np.sin(np.deg2rad(ang)))) < 5e-16
    c90, s90 = mosaic.cos_sin_deg(np.array([90.0, 180.0, -90.0, 0.0]))
    assert c90.tolist() == [-0.0, -1.0, 0.0, 1.0] or np.array_equal(np.abs(c90), [0.0, 1.0, 0.0, 1.0])
    assert np.array_equal(np.abs(s90), [1.0, 0.0, 1.0, 0.0])
