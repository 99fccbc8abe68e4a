"""Parity of the other C-ABI entry points (confusion, bilinear resize, stitching, f64 argmax) against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bilinear as obil
from oracle import confusion as oconf
from oracle import stitch as ostitch
from pistoseg_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,n", [(3, 224 * 224 * 16), (4, 1_000_003), (3, 15), (3, 16), (6, 50_000), (1, 1000)])
def test_confusion_accumulate_exact(cuda, C, n):
    g = torch.Generator().manual_seed(C * 1000 + n % 97)
    pred = torch.randint(0, C, (n,), generator=g, dtype=torch.uint8)
    gt = torch.randint(0, C + 2, (n,), generator=g, dtype=torch.uint8)  # C, C+1 are ignored labels
    conf = ops.new_confusion(C, cuda)
    bad = ops.confusion_accumulate(pred.to(cuda), gt.to(cuda), conf)
    ops.confusion_accumulate(pred.to(cuda), gt.to(cuda), conf)  # accumulates
    ref = oconf.generate_matrix(pred.numpy(), gt.numpy(), C)
    assert np.array_equal(conf.cpu().numpy(), 2 * ref)
    assert int(bad.cpu()) == 0


def test_confusion_bad_pred_counted(cuda):
    pred = torch.tensor([0, 1, 2, 3, 7, 3], dtype=torch.uint8)
    gt = torch.tensor([0, 1, 2, 0, 1, 3], dtype=torch.uint8)
    conf = ops.new_confusion(3, cuda)
    bad = ops.confusion_accumulate(pred.to(cuda), gt.to(cuda), conf)
    assert int(bad.cpu()) == 2  # pred 3 / 7 at counted pixels; the last pixel is ignored (gt == 3)
    assert np.array_equal(conf.cpu().numpy(), np.eye(3, dtype=np.int64))


@pytest.mark.parametrize("C", [1, 2, 3, 4])
def test_confusion_all_byte_values(cuda, C):
    """Every byte value in both arrays (ignore labels 255 / C, predictions >= C), sizes around the 32-pixel grouping."""
    g = torch.Generator().manual_seed(40 + C)
    for n in (31, 32, 33, 95, 4096 + 17, 1 << 20):
        pred = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
        gt = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
        small = torch.rand(n, generator=g) < 0.7      # mostly plausible labels, some arbitrary bytes
        pred = torch.where(small, pred % (C + 1), pred)
        gt = torch.where(small, gt % (C + 2), gt)
        conf = ops.new_confusion(C, cuda)
        bad = ops.confusion_accumulate(pred.to(cuda), gt.to(cuda), conf)
        p64, g64 = pred.numpy().astype(np.int64), gt.numpy().astype(np.int64)
        ok = (g64 < C) & (p64 < C)
        ref = np.bincount(g64[ok] * C + p64[ok], minlength=C * C).reshape(C, C)
        assert np.array_equal(conf.cpu().numpy(), ref), (C, n)
        assert int(bad.cpu()) == int(((g64 < C) & (p64 >= C)).sum())


def test_confusion_full_size_property(cuda):
    """At BASELINE size (10k BCSS tiles = 5e8 px): total == number of valid px, row sums == gt histogram."""
    C, n = 4, 10_000 * 224 * 224
    g = torch.Generator(device=cuda).manual_seed(3)
    pred = torch.randint(0, C, (n,), generator=g, dtype=torch.uint8, device=cuda)
    gt = torch.randint(0, C + 1, (n,), generator=g, dtype=torch.uint8, device=cuda)
    conf = ops.new_confusion(C, cuda)
    ops.confusion_accumulate(pred, gt, conf)
    hist = torch.bincount(gt.to(torch.int64), minlength=C + 1)[:C]
    assert torch.equal(conf.sum(1), hist)
    assert int(conf.sum()) == int((gt < C).sum())


@pytest.mark.parametrize("shape,size,dtype", [
    ((2, 3, 28, 28), (224, 224), torch.float32), ((2, 3, 21, 21), (224, 224), torch.float32),
    ((2, 3, 35, 35), (224, 224), torch.float32), ((2, 3, 224, 224), (32, 32), torch.float32),
    ((1, 3, 224, 224), (200, 180), torch.float32), ((1, 2, 26, 31), (207, 250), torch.float32),
    ((1, 3, 300, 280), (240, 224), torch.float64), ((1, 3, 150, 130), (333, 261), torch.float64),
    ((1, 3, 64, 64), (64, 64), torch.float32), ((1, 1, 1, 1), (5, 7), torch.float32),
])
def test_upsample_bit_exact_vs_oracle(cuda, shape, size, dtype):
    g = torch.Generator().manual_seed(sum(shape) + size[0])
    x = (torch.randn(shape, generator=g) * 3).to(dtype)
    got = ops.upsample_bilinear(x.to(cuda), size).cpu()
    ref = torch.from_numpy(obil.bilinear_restated(x.numpy(), size))
    assert torch.equal(got, ref)
    # and the torch kernel running on this GPU (what the reference executes) agrees within the float gate
    tref = F.interpolate(x.to(cuda), size, mode="bilinear").cpu()
    assert float(((got - tref).abs() / tref.abs().clamp_min(1)).max()) <= 1e-5


def _make_tiles(g, n_img_hw, scale, C=3, P=224, stride=112):
    h, w = int(n_img_hw[0] * scale), int(n_img_hw[1] * scale)
    tiles = []
    for y in range(0, max(h - P, 0) + 1, stride):
        for x in range(0, max(w - P, 0) + 1, stride):
            oh, ow = min(P, h - y), min(P, w - x)
            tiles.append((torch.randn((C, P, P), generator=g) * 3, scale, (y, x), (oh, ow)))
    return tiles


def test_stitch_big_mask_path(cuda):
    """segmentation_test.py:141-215 for one image: stitched + normalised + resized + averaged f64 probabilities."""
    g = torch.Generator().manual_seed(31)
    H, W, C = 300, 420, 3
    scales = [1.0, 1.25, 1.5]
    tiles = []
    for s in scales:
        tiles += _make_tiles(g, (H, W), s)
    ref = ostitch.big_mask_fuse(tiles, (H, W))  # [H,W,C] f64
    total = torch.zeros((C, H, W), dtype=torch.float64, device=cuda)
    for s in scales:
        ts = [t for t in tiles if t[1] == s]
        hs, ws = int(H * s), int(W * s)
        canvas = torch.zeros((C, hs, ws), dtype=torch.float64, device=cuda)
        count = torch.zeros((hs, ws), dtype=torch.float64, device=cuda)
        stack = torch.stack([t[0] for t in ts]).to(cuda)
        pos = [[t[2][0], t[2][1], t[3][0], t[3][1]] for t in ts]
        ops.stitch_accumulate(stack, pos, canvas, count, softmax=True)
        ops.canvas_normalize(canvas, count)
        ops.canvas_axpy(total, ops.upsample_bilinear(canvas, (H, W)))
    ops.canvas_normalize(total, None, float(len(scales)))
    got = total.permute(1, 2, 0).cpu().numpy()
    # fp32 softmax (expf) differs between CPU and GPU by <= 1 ulp: float gate, not bit-exactness
    assert np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-5
    gt = torch.randint(0, 4, (H, W), generator=g, dtype=torch.uint8)
    pred_ref, lab_ref = ostitch.big_mask_labels(np.nan_to_num(ref), gt.numpy())
    conf = ops.new_confusion(3, cuda)
    out = ops.argmax_f64(torch.nan_to_num(total), gt=gt, bg_match=3, bg_label=3, conf=conf)
    assert (out["pred"].cpu().numpy() == pred_ref).mean() >= 0.9999
    assert (out["labels"].cpu().numpy() == lab_ref).mean() >= 0.9999
    assert np.array_equal(conf.cpu().numpy(), oconf.generate_matrix(out["pred"].cpu().numpy(), gt.numpy(), 3))


def test_stitch_exact_without_softmax(cuda):
    """With softmax off the whole f64 pipeline is bit-exact (OEEM sum_cam / counter path, prepare_seg_inputs.py:120-131)."""
    g = torch.Generator().manual_seed(41)
    C, side, w_, h_ = 3, 224, 300, 260
    pos = [(0, 0), (0, 36), (76, 0), (76, 36), (40, 20)]
    crops = torch.randn((len(pos), C, side, side), generator=g)
    sum_cam = np.zeros((C, w_, h_)); cnt = np.zeros_like(sum_cam)
    for k, (y, x) in enumerate(pos):
        sum_cam[:, y:y + side, x:x + side] += crops[k].numpy()[:, :w_ - y, :h_ - x]
        cnt[:, y:y + side, x:x + side] += 1
    cnt[cnt < 1] = 1
    ref = obil.bilinear_restated(sum_cam / cnt, (200, 180))
    canvas = torch.zeros((C, w_, h_), dtype=torch.float64, device=cuda)
    count = torch.zeros((w_, h_), dtype=torch.float64, device=cuda)
    ops.stitch_accumulate(crops.to(cuda), [[y, x, side, side] for y, x in pos], canvas, count, softmax=False)
    ops.canvas_normalize(canvas, count, 1.0)
    got = ops.upsample_bilinear(canvas, (200, 180)).cpu().numpy()
    assert np.array_equal(got, ref)


def test_stitch_many_tiles_keeps_the_reference_order(cuda):
    """More tiles than one position chunk (512), heavy overlap, ragged crops, canvas width not a multiple of the block: the
    float64 sums must equal the reference's sequential `canvas[y:y+h, x:x+w] += tile` loop bit for bit (order matters)."""
    rng = np.random.default_rng(8)
    C, side, H, W, n = 3, 24, 211, 309, 1300
    tiles = (rng.standard_normal((n, C, side, side)) * 1e3).astype(np.float32)   # large dynamic range: order-sensitive sums
    pos = []
    ref = np.zeros((C, H, W)); cnt = np.zeros((H, W))
    for k in range(n):
        y, x = int(rng.integers(0, H - 1)), int(rng.integers(0, W - 1))
        ch, cw = int(rng.integers(1, side + 1)), int(rng.integers(1, side + 1))
        hh, ww = min(ch, H - y), min(cw, W - x)
        ref[:, y:y + hh, x:x + ww] += tiles[k][:, :hh, :ww].astype(np.float64)
        cnt[y:y + hh, x:x + ww] += 1
        pos.append([y, x, ch, cw])
    canvas = torch.zeros((C, H, W), dtype=torch.float64, device=cuda)
    count = torch.zeros((H, W), dtype=torch.float64, device=cuda)
    ops.stitch_accumulate(torch.from_numpy(tiles).to(cuda), pos, canvas, count, softmax=False)
    assert np.array_equal(canvas.cpu().numpy(), ref)
    assert np.array_equal(count.cpu().numpy(), cnt)


def test_argmax_f64_present_mask(cuda):
    g = torch.Generator().manual_seed(51)
    e = torch.randn((4, 50, 60), generator=g, dtype=torch.float64)
    e[1, :10] = e[0, :10]  # ties -> lowest index
    ref = ostitch.cam_validation_labels(e.numpy(), [1, 0, 1, 1])
    out = ops.argmax_f64(e.to(cuda), present=[1, 0, 1, 1], want_labels=False)
    assert np.array_equal(out["pred"].cpu().numpy(), ref.astype(np.uint8))
    out = ops.argmax_f64(e.to(cuda), want_labels=False)
    assert np.array_equal(out["pred"].cpu().numpy(), e.numpy().argmax(0).astype(np.uint8))


@pytest.mark.parametrize("V", [3, 5, 6, 7, 10, 12])
def test_division_by_view_count_exhaustive(cuda, V):
    """The 3-instruction a / V used for the fused-score exports equals IEEE division for ALL 2^32 float inputs."""
    import ctypes
    from pistoseg_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device=cuda)
    _lib.check(_lib.load().pisto_selftest_div(_lib.handle(0), V, ctypes.c_void_p(bad.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert int(bad.cpu()) == 0
