"""Mosaic synthesis: plan in -> pixels out, bit-exact against the oracle (itself pinned to cv2) and the cv2 fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import mosaic as omosaic
from oracle import warp_affine as owa
from pistoseg_b200 import mosaic

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def to_oracle_plan(plan):
    quads = []
    for q in range(4):
        qd = plan["quad"][q]
        M = None
        if qd["warp"]:
            # the oracle takes the forward matrix; invert the stored inverse map back (exact round trip is not needed:
            # we hand the oracle the same inverse through a tiny shim below)
            M = np.asarray(qd["minv"], np.float64).reshape(2, 3)
        quads.append(dict(flip=int(qd["flip"]), warp=bool(qd["warp"]), crop_y=int(qd["crop_y"]), crop_x=int(qd["crop_x"]), Minv=M))
    return dict(split_h=int(plan["split_h"]), split_w=int(plan["split_w"]), quads=quads)


def oracle_synthesize(plan, cells, pool_imgs, pool_bgs, labels, pn, ps):
    """oracle.mosaic.synthesize, but driven by the INVERSE matrices the C ABI carries."""
    H = W = pn * ps
    p = to_oracle_plan(plan)
    h, w = p["split_h"], p["split_w"]
    sizes = [(h, w), (h, W - w), (H - h, w), (H - h, W - w)]
    cl = np.stack([cells["tile"], cells["cy"], cells["cx"]], -1).astype(np.int64)
    outs = []
    for q in range(4):
        img, msk = omosaic.create_one_image(cl[q], pool_imgs, pool_bgs, labels, pn, ps)
        qd = p["quads"][q]
        img = np.ascontiguousarray(omosaic._flip(img, qd["flip"])); msk = np.ascontiguousarray(omosaic._flip(msk, qd["flip"]))
        if qd["warp"]:
            img = _warp_with_inverse(img, qd["Minv"], False); msk = _warp_with_inverse(msk, qd["Minv"], True)
        hq, wq = sizes[q]
        outs.append((img[qd["crop_y"]:qd["crop_y"] + hq, qd["crop_x"]:qd["crop_x"] + wq], msk[qd["crop_y"]:qd["crop_y"] + hq, qd["crop_x"]:qd["crop_x"] + wq]))
    image = np.zeros((H, W, 3), np.uint8); mask = np.zeros((H, W), np.uint8)
    image[:h, :w] = outs[0][0]; image[:h, w:] = outs[1][0]; image[h:, :w] = outs[2][0]; image[h:, w:] = outs[3][0]
    mask[:h, :w] = outs[0][1]; mask[:h, w:] = outs[1][1]; mask[h:, :w] = outs[2][1]; mask[h:, w:] = outs[3][1]
    return image, mask


def _warp_with_inverse(img, Minv, nearest):
    orig = owa.invert_affine
    try:
        owa.invert_affine = lambda M: np.asarray(M, np.float64).reshape(2, 3)
        return owa.warp_affine_u8(img, Minv, nearest)
    finally:
        owa.invert_affine = orig


def make_pool(rng, P, ps, with_bg):
    sizes = [(224, 224)] * (P - 4) + [(ps - 5, ps + 3), (ps + 4, ps - 6), (ps - 3, ps - 2), (300, ps + 1)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    bgs = [(rng.random((h, w)) < 0.12).astype(np.uint8) * 255 for h, w in sizes] if with_bg else None
    labels = rng.integers(0, 3 if with_bg else 4, P).astype(np.uint8)
    return imgs, bgs, labels


@pytest.mark.parametrize("pn,ps,with_bg", [(4, 56, True), (2, 112, False), (7, 32, True)])
def test_mosaic_bit_exact_vs_oracle(cuda, pn, ps, with_bg):
    rng = np.random.default_rng(pn * 100 + ps)
    imgs, bgs, labels = make_pool(rng, 24, ps, with_bg)
    pool = mosaic.TilePool(imgs, labels, bgs, device=cuda)
    planner = mosaic.MosaicPlanner(pool, pn, ps, seed=2022, reject_bg=with_bg)
    idx = list(range(10))
    plans, cells = planner.plans(idx)
    img, msk = mosaic.synthesize(pool, plans, cells, pn, ps)
    img_p, msk_p = mosaic.synthesize(pool, plans, cells, pn, ps, packed=False)   # planar pools: same pixels
    assert torch.equal(img, img_p) and torch.equal(msk, msk_p)
    img, msk = img.cpu().numpy(), msk.cpu().numpy()
    for k in idx:
        ri, rm = oracle_synthesize(plans[k], cells[k], imgs, bgs, labels, pn, ps)
        assert np.array_equal(img[k], ri), k
        assert np.array_equal(msk[k], rm), k
    # determinism / shard independence: mosaic i depends only on (seed, i)
    p2, c2 = planner.plans([7, 3])
    i2, m2 = mosaic.synthesize(pool, p2, c2, pn, ps)
    assert np.array_equal(i2[0].cpu().numpy(), img[7]) and np.array_equal(m2[1].cpu().numpy(), msk[3])


@pytest.mark.parametrize("tag,pn,ps", [("luad", 4, 16), ("bcss", 2, 32)])
def test_mosaic_matches_cv2_fixtures(cuda, tag, pn, ps):
    """Fixtures were produced with the real cv2.flip / cv2.warpAffine / copyMakeBorder (tests/golden/make_golden.py)."""
    from tests.test_oracle_golden import unpack_pool
    z = np.load(os.path.join(G, "mosaic.npz"))
    pool_imgs, pool_bgs = unpack_pool(z, tag)
    labels = z[f"{tag}_labels"]
    pool = mosaic.TilePool(pool_imgs, labels, pool_bgs, device=cuda)
    n = z[f"{tag}_img"].shape[0]
    plans = np.zeros(n, mosaic.PLAN_DTYPE); cells = np.zeros((n, 4, pn * pn), mosaic.CELL_DTYPE)
    for t in range(n):
        row = z[f"{tag}_plan"][t]
        plans[t]["split_h"], plans[t]["split_w"] = row[0], row[1]
        for q in range(4):
            flip, warp, cy, cx = (int(v) for v in row[2 + 4 * q: 6 + 4 * q])
            qd = plans[t]["quad"][q]
            qd["flip"], qd["warp"], qd["crop_y"], qd["crop_x"] = flip, warp, cy, cx
            if warp:
                qd["minv"] = mosaic.invert_affine(z[f"{tag}_M{t}"][q]).reshape(-1)
        cl = z[f"{tag}_cells"][t]
        cells[t]["tile"], cells[t]["cy"], cells[t]["cx"] = cl[..., 0], cl[..., 1], cl[..., 2]
    img, msk = mosaic.synthesize(pool, plans, cells, pn, ps)
    assert np.array_equal(img.cpu().numpy(), z[f"{tag}_img"])
    assert np.array_equal(msk.cpu().numpy(), z[f"{tag}_mask"])


@pytest.mark.parametrize("pn,ps,with_bg", [(4, 56, True), (2, 112, False), (7, 32, True)])
def test_device_planner_equals_host_planner(cuda, pn, ps, with_bg):
    """pisto_mosaic_plan_cells / pisto_mosaic_bg_integral == their numpy restatements, bit for bit; sharded generation
    (first_index, stride) reproduces the same mosaics."""
    rng = np.random.default_rng(pn * 7 + ps)
    imgs, bgs, labels = make_pool(rng, 40, ps, with_bg)
    if with_bg:  # make the rejection loop bite: half of the tiles are mostly background
        for t in range(0, 40, 2):
            bgs[t] = (rng.random(bgs[t].shape) < 0.9).astype(np.uint8) * 255
    pool = mosaic.TilePool(imgs, labels, bgs, device=cuda)
    planner = mosaic.MosaicPlanner(pool, pn, ps, seed=77, reject_bg=with_bg)
    if with_bg:
        ioff, I = pool.integral_host(ps)
        ioff_d, I_d = pool.integral_device(ps)
        assert np.array_equal(I_d.cpu().numpy().view(np.uint16), I)
    N = 16
    host = planner.cells_host(range(5, 5 + 3 * N, 3))
    dev = planner.cells_device(5, 3, N).cpu().numpy().view(mosaic.CELL_DTYPE).reshape(N, 4, pn * pn)
    assert np.array_equal(dev, host)
    if with_bg:
        assert not np.array_equal(host, mosaic.MosaicPlanner(pool, pn, ps, seed=77, reject_bg=False).cells_host(range(5, 5 + 3 * N, 3)))
    img, msk = mosaic.synthesize_range(pool, planner, 5, 3, N)
    p, c = planner.plans(range(5, 5 + 3 * N, 3))
    img2, msk2 = mosaic.synthesize(pool, p, c, pn, ps)
    assert torch.equal(img, img2) and torch.equal(msk, msk2)
    # 64-bit mosaic indices
    big = planner.cells_device(2 ** 33 + 1, 1, 2).cpu().numpy().view(mosaic.CELL_DTYPE).reshape(2, 4, pn * pn)
    assert np.array_equal(big, planner.cells_host([2 ** 33 + 1, 2 ** 33 + 2]))


@pytest.mark.parametrize("pn,ps", [(4, 56), (7, 32)])
def test_device_quadrant_planner_equals_host_planner(cuda, pn, ps):
    """pisto_mosaic_plan_quads == MosaicPlanner.quad_plans byte for byte (split, flip, warp flag, float64 inverse affine, crop
    origin) over 4 096 strided indices and beyond 2^32."""
    rng = np.random.default_rng(3)
    imgs, bgs, labels = make_pool(rng, 12, ps, False)
    pool = mosaic.TilePool(imgs, labels, bgs, device=cuda)
    planner = mosaic.MosaicPlanner(pool, pn, ps, seed=2022)
    for first, stride, N in [(0, 1, 4096), (3, 8, 1000), (2 ** 33 + 5, 7, 64)]:
        host = planner.quad_plans([first + k * stride for k in range(N)])
        dev = planner.quads_device(first, stride, N).cpu().numpy().view(mosaic.PLAN_DTYPE)
        assert host.tobytes() == dev.tobytes(), (first, stride)
    assert planner.quads_device(0, 1, 0).numel() == 0


def test_rejection_loop_exhaustion_is_counted(cuda):
    """The reference's `while True` rejection loop (create_dataset.ipynb:303-309) never ends on a pool that is all background; the
    device planner accepts the draw number max_tries and counts the cell (pisto_filter_stats slot 0) instead of hiding it."""
    from pistoseg_b200 import _lib
    rng = np.random.default_rng(4)
    pn, ps, P = 4, 56, 6
    imgs = [rng.integers(0, 256, (224, 224, 3), dtype=np.uint8) for _ in range(P)]
    bgs = [np.full((224, 224), 255, np.uint8) for _ in range(P)]
    pool = mosaic.TilePool(imgs, rng.integers(0, 3, P).astype(np.uint8), bgs, device=cuda)
    planner = mosaic.MosaicPlanner(pool, pn, ps, seed=1, reject_bg=True, max_tries=3)
    _lib.filter_stats(cuda.index or 0, reset=True)
    N = 8
    planner.cells_device(0, 1, N)
    assert _lib.filter_stats(cuda.index or 0, reset=True)["mosaic_cells_exhausted"] == N * 4 * pn * pn
    bgs = [np.zeros((224, 224), np.uint8) for _ in range(P)]
    pool = mosaic.TilePool(imgs, rng.integers(0, 3, P).astype(np.uint8), bgs, device=cuda)
    mosaic.MosaicPlanner(pool, pn, ps, seed=1, reject_bg=True, max_tries=3).cells_device(0, 1, N)
    assert _lib.filter_stats(cuda.index or 0, reset=True)["mosaic_cells_exhausted"] == 0


def test_export_dataset_writes_the_notebook_layout(cuda, tmp_path):
    """create_dataset.ipynb:523-560: rank r writes img/<i>.png (RGB) and mask/<i>.png (mode P, dataset palette) for i = r (mod G);
    the union over ranks is every index once and the decoded pixels are the synthesised ones."""
    from PIL import Image
    rng = np.random.default_rng(5)
    pn, ps = 4, 16
    imgs, bgs, labels = make_pool(rng, 12, ps, True)
    pool = mosaic.TilePool(imgs, labels, bgs, device=cuda)
    planner = mosaic.MosaicPlanner(pool, pn, ps, seed=9, reject_bg=True)
    n_total, world = 21, 2
    wrote = [mosaic.export_dataset(pool, planner, str(tmp_path), n_total, rank=r, world=world, chunk=4) for r in range(world)]
    assert wrote == [11, 10]
    names = sorted(os.listdir(tmp_path / "img"))
    assert names == sorted(os.listdir(tmp_path / "mask")) == [f"{i:07d}.png" for i in range(n_total)]
    img, msk = mosaic.synthesize_range(pool, planner, 0, 1, n_total)
    for i in (0, 1, 10, 20):
        im = Image.open(tmp_path / "img" / f"{i:07d}.png"); mk = Image.open(tmp_path / "mask" / f"{i:07d}.png")
        assert im.mode == "RGB" and mk.mode == "P"
        assert np.array_equal(np.asarray(im), img[i].cpu().numpy()) and np.array_equal(np.asarray(mk), msk[i].cpu().numpy())
