"""Parity of the fused hot path (pisto_fuse_argmax_confusion) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import confusion as oconf
from oracle import fuse as ofuse
from pistoseg_b200 import ops, synthetic
from pistoseg_b200._lib import (DECIDE_RAW, DECIDE_SOFTMAX, FUSE_LOGIT_MEAN, FUSE_PROB_MEAN, IMPL_GENERIC, IMPL_STREAM,
                                MASK_FILL, MASK_MULTIPLY, MASK_NEG_INF, MASK_NONE)

pytestmark = pytest.mark.gpu

FLOAT_GATE = 1e-5  # |ours - ref| <= 1e-5 * max(|ref|, 1)   (BASELINE.json north_star)


def gate(ours, ref):
    ours = np.asarray(ours, np.float64); ref = np.asarray(ref, np.float64)
    return float((np.abs(ours - ref) / np.maximum(np.abs(ref), 1.0)).max())


def run(cfg, cuda, impl, **kw):
    views = [v.to(cuda) for v in cfg["views"]]
    return ops.fuse_argmax_confusion(views, cfg["codes"], (cfg["T"], cfg["T"]), impl=impl,
                                     present=cfg.get("present"), bg=cfg.get("bg"), gt=cfg.get("gt"), **kw)


def oracle_fused(cfg, mode=ofuse.LOGIT_MEAN):
    return ofuse.fuse_views(cfg["views"], cfg["codes"], (cfg["T"], cfg["T"]), mode)


@pytest.mark.parametrize("impl", [IMPL_GENERIC, IMPL_STREAM])
def test_cfg1_single_view(cuda, impl):
    """BASELINE config 1: labels, bg, confusion and fused scores are bit-exact vs the oracle."""
    cfg = synthetic.cfg1(N=24)
    out = run(cfg, cuda, impl, mask_mode=MASK_NONE, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, want_fused=True)
    fused = oracle_fused(cfg)
    assert torch.equal(out["fused"].cpu(), fused), "fused scores differ from the oracle bit pattern"
    pred = ofuse.miou_pred(fused).numpy()
    lab = pred.copy(); lab[cfg["bg"].numpy() == 1] = 3
    assert np.array_equal(out["labels"].cpu().numpy(), lab)
    cm = sum(oconf.generate_matrix(pred[n], cfg["gt"][n].numpy(), 3) for n in range(pred.shape[0]))
    assert np.array_equal(out["conf"].cpu().numpy(), cm)


@pytest.mark.parametrize("impl", [IMPL_GENERIC, IMPL_STREAM])
def test_cfg2_pseudo_masks(cuda, impl):
    """BASELINE config 2: 3 scales x flip, present vector with single-label shortcut, bg, 32x32 logit export."""
    cfg = synthetic.cfg2(N=32)
    out = run(cfg, cuda, impl, mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32), want_fused=True)
    fused = oracle_fused(cfg)
    assert torch.equal(out["fused"].cpu(), fused)
    assert torch.equal(out["lowres"].cpu(), ofuse.lowres_32(fused))
    lab = ofuse.pseudo_masks(fused, cfg["present"].numpy(), cfg["bg"].numpy())
    got = out["labels"].cpu().numpy()
    agree = (got == lab).mean()
    assert agree >= 0.9999, agree
    # without fused_out the single-label tiles take the no-scores path: same labels, same 32x32 logits
    out2 = run(cfg, cuda, impl, mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, lowres=(32, 32))
    assert torch.equal(out2["labels"], out["labels"])
    assert torch.equal(out2["lowres"], out["lowres"])


@pytest.mark.parametrize("impl", [IMPL_GENERIC, IMPL_STREAM])
def test_cfg3_bcss_confusion(cuda, impl):
    cfg = synthetic.cfg3(N=20)
    out = run(cfg, cuda, impl, decide=DECIDE_SOFTMAX, want_fused=True)
    fused = oracle_fused(cfg)
    assert torch.equal(out["fused"].cpu(), fused)
    pred = ofuse.miou_pred(fused).numpy()
    assert (out["labels"].cpu().numpy() == pred).mean() >= 0.9999
    # confusion must be exact GIVEN our labels (integer path), and within the label-agreement budget of the oracle's
    cm_ours = oconf.generate_matrix(out["labels"].cpu().numpy(), cfg["gt"].numpy(), 4)
    assert np.array_equal(out["conf"].cpu().numpy(), cm_ours)
    cm_ref = oconf.generate_matrix(pred, cfg["gt"].numpy(), 4)
    assert np.abs(cm_ours - cm_ref).sum() <= 2 * (1 - 0.9999) * pred.size


def test_cfg5_large_tile(cuda):
    cfg = synthetic.cfg5(N=2, T=512)
    outs = {}
    for impl in (IMPL_GENERIC, IMPL_STREAM):
        outs[impl] = run(cfg, cuda, impl, decide=DECIDE_RAW, want_fused=True)
    fused = oracle_fused(cfg)
    for impl in outs:
        assert torch.equal(outs[impl]["fused"].cpu(), fused), impl
        assert np.array_equal(outs[impl]["labels"].cpu().numpy(), fused.argmax(1).numpy().astype(np.uint8))
    assert torch.equal(outs[IMPL_GENERIC]["conf"], outs[IMPL_STREAM]["conf"])


def test_tie_stress_labels_exact(cuda):
    """Logits rounded to multiples of 0.5 produce many exact ties: lowest index must win, on both kernels, exactly."""
    cfg = synthetic.cfg1(N=8)
    cfg["views"] = [torch.round(v * 2) / 2 for v in cfg["views"]]
    fused = oracle_fused(cfg)
    for decide, ref in ((DECIDE_SOFTMAX, ofuse.miou_pred(fused)), (DECIDE_RAW, ofuse.miou_pred(fused, probs=True))):
        for impl in (IMPL_GENERIC, IMPL_STREAM):
            out = run(cfg, cuda, impl, decide=decide, bg_match=255)
            got = out["labels"].cpu().numpy()
            # ties in softmax space depend on expf rounding; raw-argmax ties do not
            if decide == DECIDE_RAW:
                assert np.array_equal(got, ref.numpy())
            else:
                assert (got == ref.numpy()).mean() >= 0.9999


def test_argmax_bit_exact_on_reference_scores(cuda):
    """north_star: 'argmax labels given identical fused scores must be bit-exact'.  Feed the oracle's fused scores as a
    single full-resolution view (identity resize) and compare with torch.argmax on the same scores."""
    cfg = synthetic.cfg2(N=16)
    fused = oracle_fused(cfg)
    for impl in (IMPL_GENERIC, IMPL_STREAM):
        out = ops.fuse_argmax_confusion([fused.to(cuda)], [0], (224, 224), decide=DECIDE_RAW, impl=impl)
        assert np.array_equal(out["labels"].cpu().numpy(), fused.argmax(1).numpy().astype(np.uint8))


def test_prob_mean_mode(cuda):
    cfg = synthetic.cfg2(N=8)
    ref = oracle_fused(cfg, ofuse.PROB_MEAN)
    for impl in (IMPL_GENERIC, IMPL_STREAM):
        out = run(cfg, cuda, impl, fuse_mode=FUSE_PROB_MEAN, mask_mode=MASK_NONE, decide=DECIDE_RAW, bg_match=255, want_fused=True)
        assert gate(out["fused"].cpu().numpy(), ref.numpy()) <= FLOAT_GATE
        assert (out["labels"].cpu().numpy() == ref.argmax(1).numpy()).mean() >= 0.9999


def test_d4_fullres_views(cuda):
    """Reference-literal mode: 8 dihedral full-resolution views (ttach d4), merged as mean."""
    from oracle import tta
    g = torch.Generator().manual_seed(5)
    N, C, T = 6, 3, 64
    views = [torch.randn((N, C, T, T), generator=g) * 3 for _ in tta.D4_VIEWS]
    codes = [tta.deaug_code(hf, ang) for hf, ang in tta.D4_VIEWS]
    ref = ofuse.fuse_views(views, codes, (T, T))
    for impl in (IMPL_GENERIC, IMPL_STREAM):
        out = ops.fuse_argmax_confusion([v.to(cuda) for v in views], codes, (T, T), decide=DECIDE_SOFTMAX, want_fused=True, impl=impl)
        assert torch.equal(out["fused"].cpu(), ref), impl
        assert (out["labels"].cpu().numpy() == ofuse.miou_pred(ref).numpy()).mean() >= 0.9999


def test_dihedral_lowres_nonsquare(cuda):
    """All 8 de-augmentation codes on non-square low-resolution views (generic kernel)."""
    g = torch.Generator().manual_seed(9)
    N, C = 3, 4
    T = (96, 120)
    views, codes = [], []
    for code in range(8):
        h, w = (12, 15) if code % 2 == 0 else (15, 12)
        views.append(torch.randn((N, C, h, w), generator=g) * 3); codes.append(code)
    ref = ofuse.fuse_views(views, codes, T)
    out = ops.fuse_argmax_confusion([v.to(cuda) for v in views], codes, T, decide=DECIDE_RAW, want_fused=True)
    assert torch.equal(out["fused"].cpu(), ref)


def test_mask_modes(cuda):
    g = torch.Generator().manual_seed(11)
    N, C, T = 8, 4, 64
    x = torch.randn((N, C, T, T), generator=g) * 2
    present = synthetic.make_present(N, C, 12, single_frac=0.25)
    # NEG_INF + raw argmax (generate_CAM.py:91-99)
    ref = x.clone().numpy()
    for n in range(N):
        for c in range(C):
            if present[n, c] == 0:
                ref[n, c] = -np.inf
    out = ops.fuse_argmax_confusion([x.to(cuda)], [0], (T, T), mask_mode=MASK_NEG_INF, decide=DECIDE_RAW, present=present)
    assert np.array_equal(out["labels"].cpu().numpy(), ref.argmax(1).astype(np.uint8))
    # MULTIPLY on a channel slice of [N, C+1, T, T] (infer_revise_masks.py:137-143)
    xx = torch.randn((N, C + 1, T, T), generator=g)
    label = torch.cat([torch.ones(N, 1), present.float()], 1)
    bgm = (torch.rand((N, T, T), generator=g) < 0.2).to(torch.uint8) * 255
    ref = ofuse.revise_masks(xx, label, bgm.numpy(), 3)
    xg = xx.to(cuda)
    out = ops.fuse_argmax_confusion([xg[:, 1:]], [0], (T, T), mask_mode=MASK_MULTIPLY, decide=DECIDE_RAW, present=present, bg=bgm,
                                    bg_match=255, bg_label=3)
    assert np.array_equal(out["labels"].cpu().numpy(), ref.astype(np.uint8))


def test_entropy_output(cuda):
    cfg = synthetic.cfg2(N=6)
    fused = oracle_fused(cfg)
    lab, ent = ofuse.pseudo_masks(fused, cfg["present"].numpy(), cfg["bg"].numpy(), want_entropy=True)
    out = run(cfg, cuda, IMPL_GENERIC, mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, bg_match=1, bg_label=3, want_entropy=True)
    assert (out["labels"].cpu().numpy() == lab).mean() >= 0.9999
    assert np.abs(out["entropy"].cpu().numpy() - ent).max() <= 2e-5


def test_matches_torch_cuda_interpolate(cuda):
    """The real reference runs F.interpolate on CUDA tensors: our fused scores equal it bit-for-bit on this GPU."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(21)
    for h in (21, 28, 35):
        x = (torch.randn((4, 3, h, h), generator=g) * 3).to(cuda)
        out = ops.fuse_argmax_confusion([x], [0], (224, 224), want_fused=True, want_labels=False)
        assert torch.equal(out["fused"], F.interpolate(x, (224, 224), mode="bilinear"))


def test_host_pipeline_equals_device_call(cuda):
    cfg = synthetic.cfg2(N=70)
    dev = run(cfg, cuda, 0, mask_mode=MASK_FILL, bg_match=1, bg_label=3, lowres=(32, 32))
    host = ops.fuse_argmax_confusion_host([v.pin_memory() for v in cfg["views"]], cfg["codes"], (224, 224), mask_mode=MASK_FILL,
                                          present=cfg["present"], bg=cfg["bg"].pin_memory(), bg_match=1, bg_label=3,
                                          lowres=(32, 32), chunk=32)
    assert torch.equal(host["labels"], dev["labels"].cpu())
    assert torch.equal(host["lowres"], dev["lowres"].cpu())


def test_empty_and_errors(cuda):
    from pistoseg_b200._lib import PistoError
    x = torch.zeros((0, 3, 28, 28), device=cuda)
    out = ops.fuse_argmax_confusion([x], [0], (224, 224))
    assert out["labels"].shape == (0, 224, 224)
    with pytest.raises(PistoError):
        ops.fuse_argmax_confusion([torch.zeros((1, 3, 28, 28))], [0], (224, 224))  # CPU tensor: no fallback
    with pytest.raises(PistoError):
        ops.fuse_argmax_confusion([torch.zeros((1, 9, 28, 28), device=cuda)], [0], (224, 224))  # C > 8


@pytest.mark.parametrize("C", [2, 3, 4, 5])
def test_identity_view_fast_path_equals_generic(cuda, C):
    """One full-resolution view (mIoUMask.forward / revise masks): the streaming identity kernel (automatic dispatch) and the
    generic kernel agree on labels and confusion for every mask / decide mode, with ties, NaN and a strided channel slice."""
    g = torch.Generator().manual_seed(100 + C)
    N, T = 7, 64
    x = torch.round(torch.randn((N, C + 1, T, T), generator=g) * 4) / 4   # quarter steps: plenty of exact ties
    x[1, 1, 3, 5] = float("nan"); x[2, 2, 0, 0] = float("inf")
    present = synthetic.make_present(N, C, 7, single_frac=0.3)
    gt = torch.randint(0, C + 1, (N, T, T), generator=g, dtype=torch.uint8)
    bg = (torch.rand((N, T, T), generator=g) < 0.2).to(torch.uint8)
    xg = x.to(cuda)
    for view in (xg[:, 1:], xg[:, :C].contiguous()):
        for mask in (MASK_NONE, MASK_FILL, MASK_NEG_INF, MASK_MULTIPLY):
            for decide in (DECIDE_RAW, DECIDE_SOFTMAX):
                kw = dict(mask_mode=mask, decide=decide, present=present if mask != MASK_NONE else None, bg=bg, bg_match=1, bg_label=C, gt=gt)
                a = ops.fuse_argmax_confusion([view], [0], (T, T), conf=ops.new_confusion(C, cuda), **kw)
                b = ops.fuse_argmax_confusion([view], [0], (T, T), conf=ops.new_confusion(C, cuda), impl=IMPL_GENERIC, **kw)
                assert torch.equal(a["labels"], b["labels"]), (C, mask, decide)
                assert torch.equal(a["conf"], b["conf"]), (C, mask, decide)


@pytest.mark.parametrize("T,C", [((224, 224), 3), ((96, 120), 4), ((70, 45), 3), ((100, 120), 3), ((64, 32), 2), ((224, 224), 4)])
def test_fullres_d4_kernel_equals_generic_and_oracle(cuda, T, C):
    """ttach d4 on full-resolution logits (the literal infer_pseudo_masks.py path): the streaming full-resolution kernel
    (automatic dispatch) == the generic kernel == the oracle, for fused scores, labels, 32x32 export and confusion."""
    from oracle import tta
    g = torch.Generator().manual_seed(31 + C)
    N = 5
    views, codes = [], []
    for hf, ang in tta.D4_VIEWS:
        code = tta.deaug_code(hf, ang)
        shape = (N, C, T[1], T[0]) if code % 2 else (N, C, T[0], T[1])
        views.append(torch.round(torch.randn(shape, generator=g) * 8) / 8); codes.append(code)
    present = synthetic.make_present(N, C, 5, single_frac=0.4)
    gt = torch.randint(0, C + 1, (N,) + T, generator=g, dtype=torch.uint8)
    bg = (torch.rand((N,) + T, generator=g) < 0.2).to(torch.uint8)
    vg = [v.to(cuda) for v in views]
    low = (32, 32) if T == (224, 224) else None
    kw = dict(mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, present=present, bg=bg, bg_match=1, bg_label=C, gt=gt, lowres=low, want_fused=True)
    a = ops.fuse_argmax_confusion(vg, codes, T, conf=ops.new_confusion(C, cuda), **kw)
    b = ops.fuse_argmax_confusion(vg, codes, T, conf=ops.new_confusion(C, cuda), impl=IMPL_GENERIC, **kw)
    ref = ofuse.fuse_views(views, codes, T)
    assert torch.equal(a["fused"].cpu(), ref)
    for k in ("fused", "labels", "conf") + (("lowres",) if low else ()):
        assert torch.equal(a[k], b[k]), k
    # without the score outputs (single-label tiles then skip the views entirely)
    kw.update(want_fused=False, lowres=None)
    a2 = ops.fuse_argmax_confusion(vg, codes, T, conf=ops.new_confusion(C, cuda), **kw)
    assert torch.equal(a2["labels"], b["labels"]) and torch.equal(a2["conf"], b["conf"])


@pytest.mark.parametrize("C", [3, 4])
def test_second_label_output_equals_a_separate_raw_pass(cuda, C):
    """label_raw_out (segmentation_test.py:137-139,182: softmax-argmax for the matrix, logit-argmax for the PNG, one pass over the logits):
    labels == the DECIDE_SOFTMAX call's, labels_raw == the DECIDE_RAW call's, the confusion matrix is the softmax one -- on the one-view
    full-resolution kernel and on the generic kernel, with near ties, exact ties, a presence vector and a background mask."""
    from pistoseg_b200 import _lib
    g = torch.Generator().manual_seed(70 + C)
    N, T = 6, 224
    x = torch.randn((N, C, T, T), generator=g) * 2
    x[1, 1] = x[1, 0] + torch.randn((T, T), generator=g) * 1e-7          # near ties
    x[2] = torch.round(x[2])                                                # exact ties
    x[3, :, 5, 7] = float("nan")
    gt = torch.randint(0, C + 1, (N, T, T), generator=g, dtype=torch.uint8)
    bg = (torch.rand((N, T, T), generator=g) < 0.1).to(torch.uint8)
    present = synthetic.make_present(N, C, 9, single_frac=0.3)
    for impl in (0, IMPL_GENERIC):
        for kw in (dict(), dict(present=present.to(cuda), mask_mode=MASK_FILL, bg=bg.to(cuda), bg_match=1, bg_label=C)):
            conf2 = ops.new_confusion(C, cuda)
            both = ops.fuse_argmax_confusion([x.to(cuda)], [0], (T, T), decide=DECIDE_SOFTMAX, gt=gt.to(cuda), conf=conf2, impl=impl,
                                             want_raw_labels=True, **kw)
            conf1 = ops.new_confusion(C, cuda)
            soft = ops.fuse_argmax_confusion([x.to(cuda)], [0], (T, T), decide=DECIDE_SOFTMAX, gt=gt.to(cuda), conf=conf1, impl=impl, **kw)
            raw = ops.fuse_argmax_confusion([x.to(cuda)], [0], (T, T), decide=DECIDE_RAW, impl=impl, **kw)
            assert torch.equal(both["labels"], soft["labels"]) and torch.equal(conf2, conf1)
            assert torch.equal(both["labels_raw"], raw["labels"])
    # several views: only the generic kernel serves the second output; a forced fast kernel refuses
    cfg = synthetic.cfg2(N=4)
    views = [v.to(cuda) for v in cfg["views"]]
    both = ops.fuse_argmax_confusion(views, cfg["codes"], (224, 224), decide=DECIDE_SOFTMAX, want_raw_labels=True)
    assert torch.equal(both["labels_raw"], ops.fuse_argmax_confusion(views, cfg["codes"], (224, 224), decide=DECIDE_RAW, impl=IMPL_GENERIC)["labels"])
    with pytest.raises(_lib.PistoError):
        ops.fuse_argmax_confusion(views, cfg["codes"], (224, 224), decide=DECIDE_SOFTMAX, want_raw_labels=True, impl=IMPL_STREAM)
