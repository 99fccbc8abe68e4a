"""The CPU oracle against fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import bilinear as obil
from oracle import confusion as oconf
from oracle import fuse as ofuse
from oracle import mosaic as omosaic

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name))


@pytest.mark.parametrize("tag,C", [("luad", 3), ("bcss", 4), ("empty", 3)])
def test_confusion_and_iou_match_reference_loss_py(tag, C):
    z = load("miou.npz")
    l1, l2 = torch.from_numpy(z[f"{tag}_logits1"]), torch.from_numpy(z[f"{tag}_logits2"])
    cm = oconf.generate_matrix(ofuse.miou_pred(l1).numpy(), z[f"{tag}_mask1"], C)
    ret1 = (oconf.mean_iou(cm), oconf.fw_iou(cm))
    assert np.array_equal(np.array(ret1), z[f"{tag}_ret1"])
    cm = cm + oconf.generate_matrix(ofuse.miou_pred(torch.softmax(l2, 1), probs=True).numpy(), z[f"{tag}_mask2"], C)
    assert np.array_equal(cm.astype(np.float64), z[f"{tag}_cm"])
    assert np.array_equal(oconf.tissue_iou(cm), z[f"{tag}_tissue"])
    assert oconf.mean_iou(cm) == float(z[f"{tag}_miou"])
    assert oconf.fw_iou(cm) == float(z[f"{tag}_fwiou"])
    assert np.array_equal(np.array([oconf.mean_iou(cm), oconf.fw_iou(cm)]), z[f"{tag}_ret2"])


def test_empty_class_iou_is_zero_and_skipped_in_fwiou():
    z = load("miou.npz")
    assert z["empty_tissue"][2] == 0.0
    assert z["empty_cm"][2].sum() == 0 and z["empty_cm"][:, 2].sum() == 0


def test_get_mask_pred_and_entropy_matches_reference_function():
    z = load("pmask.npz")
    i = 0
    while f"logit{i}" in z:
        logit = torch.from_numpy(z[f"logit{i}"].copy())
        lab = [int(v) for v in z[f"label{i}"]]
        mask, ent = ofuse.get_mask_pred_and_entropy(logit, z[f"tissue{i}"], lab)
        assert np.array_equal(np.asarray(mask), z[f"mask{i}"]), i
        assert np.array_equal(np.asarray(ent, dtype=np.float32), z[f"entropy{i}"]), i
        assert np.array_equal(logit.numpy(), z[f"mutated{i}"]), "in-place -1e10 fill must be visible to the caller"
        # batch driver (what the GPU tests use) == per-tile reference function
        fused = torch.from_numpy(z[f"logit{i}"].copy())[None]
        got = ofuse.pseudo_masks(fused, np.array([lab]), (z[f"tissue{i}"] == 0)[None])
        assert np.array_equal(got[0], z[f"mask{i}"].astype(np.uint8))
        # 56 -> 8 export is the gather [3::7, 3::7]
        assert np.array_equal(z[f"low{i}"], z[f"logit{i}"][:, 3::7, 3::7])
        assert np.array_equal(obil.bilinear_restated(z[f"logit{i}"], (8, 8)), z[f"low{i}"])
        i += 1
    assert i == 7


def test_bilinear_restatement_matches_torch_interpolate():
    z = load("bilinear.npz")
    for i in range(int(z["n"])):
        x, y, size = z[f"x{i}"], z[f"y{i}"], tuple(int(v) for v in z[f"size{i}"])
        got = obil.bilinear_restated(x, size)
        assert got.dtype == y.dtype
        # fixtures were produced by ATen's vectorised CPU kernel (and equal the CUDA kernel, profiles/r01/probe_torch.json)
        assert np.array_equal(got, y), (i, np.abs(got - y).max())


def test_bilinear_restatement_within_gate_of_live_torch():
    """Whatever code path this machine's torch takes (it depends on shape / thread count), the restatement stays within
    the 1e-5 float gate of it (in fact within 2 ulp)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    for shp, size in (((1, 3, 8, 8), (16, 16)), ((1, 3, 3, 5), (7, 11)), ((1, 3, 256, 256), (32, 32)), ((2, 3, 28, 28), (224, 224))):
        x = torch.randn(shp, generator=g) * 3
        ref = F.interpolate(x, size, mode="bilinear").numpy()
        got = obil.bilinear_restated(x.numpy(), size)
        assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1)).max() <= 1e-6


def _plan_from_row(row, Ms):
    quads = []
    for q in range(4):
        flip, warp, cy, cx = (int(v) for v in row[2 + 4 * q: 6 + 4 * q])
        quads.append(dict(flip=flip, warp=bool(warp), crop_y=cy, crop_x=cx, M=Ms[q] if warp else None))
    return dict(split_h=int(row[0]), split_w=int(row[1]), quads=quads)


def unpack_pool(z, tag):
    hw = z[f"{tag}_pool_hw"]
    flat = z[f"{tag}_pool"]
    bgflat = z[f"{tag}_bg"] if f"{tag}_bg" in z else None
    pool, bgs, o = [], [], 0
    for h, w in hw:
        pool.append(flat[3 * o: 3 * (o + h * w)].reshape(h, w, 3))
        if bgflat is not None:
            bgs.append(bgflat[o: o + h * w].reshape(h, w))
        o += h * w
    return pool, (bgs if bgflat is not None else None)


@pytest.mark.parametrize("tag,pn,ps", [("luad", 4, 16), ("bcss", 2, 32)])
def test_mosaic_oracle_matches_cv2_fixtures(tag, pn, ps):
    z = load("mosaic.npz")
    pool, bgs = unpack_pool(z, tag)
    for t in range(z[f"{tag}_img"].shape[0]):
        plan = _plan_from_row(z[f"{tag}_plan"][t], z[f"{tag}_M{t}"])
        img, msk = omosaic.synthesize(plan, z[f"{tag}_cells"][t], pool, bgs, z[f"{tag}_labels"], pn, ps, use_cv2=False)
        assert np.array_equal(img, z[f"{tag}_img"][t]), t
        assert np.array_equal(msk, z[f"{tag}_mask"][t]), t


def test_warp_affine_restatement_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import warp_affine as wa
    rng = np.random.default_rng(7)
    for trial in range(12):
        S = int(rng.choice([64, 96, 224]))
        img = rng.integers(0, 256, (S, S, 3), dtype=np.uint8); msk = rng.integers(0, 4, (S, S), dtype=np.uint8)
        M = wa.shift_scale_rotate_matrix(S, S, rng.uniform(-45, 45), rng.uniform(0.8, 1.2), rng.uniform(-.0625, .0625), rng.uniform(-.0625, .0625))
        assert np.array_equal(wa.warp_affine_u8(img, M), cv2.warpAffine(img, M, (S, S), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101))
        assert np.array_equal(wa.warp_affine_u8(msk, M, nearest=True), cv2.warpAffine(msk, M, (S, S), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_REFLECT_101))


def test_background_oracle_known_cases():
    """oracle/background.py: cv2 gray/threshold + restated remove_small_objects (4-connectivity, size < 50 removed)."""
    from oracle import background as obg
    img = np.zeros((40, 40, 3), np.uint8)
    img[0:7, 0:7] = 255            # 49 pixels: removed
    img[10:15, 10:20] = 255        # 50 pixels: kept
    img[15, 20] = 255              # diagonal neighbour of the 50-pixel block: not 4-connected, removed
    img[30:32, 0:40] = (201, 201, 201)  # gray 201 > 200: 80 pixels kept
    img[35:37, 0:40] = (200, 200, 200)  # gray 200: not background
    m = obg.get_background(img)
    assert m.dtype == np.uint8 and set(np.unique(m)) == {0, 255}
    assert m[0:7, 0:7].sum() == 0 and (m[10:15, 10:20] == 255).all() and m[15, 20] == 0
    assert (m[30:32] == 255).all() and m[35:37].sum() == 0


# ---- fixtures produced by executing the reference's SCRIPT bodies (make_golden.py: golden_bigmask / golden_oeem / golden_revise)
def _bigmask_tiles(z, key):
    """Tiles of image `key` in dataset order: (logits torch [3,P,P], scale, (y, x), (orig_h, orig_w))."""
    out = []
    for name, logit, hw in zip(z["names"], z["logits"], z["orig_hw"]):
        name = str(name)
        if name.split("_")[0] != key:
            continue
        out.append((torch.from_numpy(logit), float(name.split("_")[1]), (int(name.split("_")[2]), int(name.split("_")[3].split("-")[0])),
                    (int(hw[0]), int(hw[1]))))
    return out


@pytest.mark.parametrize("literal", [True, False])
def test_big_mask_fusion_matches_the_reference_script(literal):
    """oracle/stitch.py vs segmentation_test.py:141-215 run on synthetic tiles (softmax, overlap-add per scale, f64 resize, mean)."""
    from oracle import stitch as ostitch
    z = load("bigmask.npz")
    total = np.zeros((3, 3))
    for key in ("00", "01"):
        gt = z[f"gt_{key}"]
        fused = ostitch.big_mask_fuse(_bigmask_tiles(z, key), gt.shape, literal=literal)
        ref = z[f"fused_{key}"]
        assert fused.shape == ref.shape
        if literal:
            assert np.array_equal(fused, ref)               # the same torch calls: bit for bit
        else:
            assert np.abs(fused - ref).max() <= 1e-12       # restated float64 bilinear
        pred, lab = ostitch.big_mask_labels(ref, gt)
        assert np.array_equal(lab, z[f"png_{key}"]) and str(z[f"png_mode_{key}"]) == "P"
        assert list(z[f"palette_{key}"]) == [0, 64, 128, 64, 128, 0, 243, 152, 0, 255, 255, 255]
        total += oconf.generate_matrix(pred, gt, 3)
    assert np.array_equal(total, z["big_cm"])
    # patch-level matrix of the same run (loss.py:55-67 on every batch)
    pm = np.zeros((3, 3))
    for logit, mask in zip(z["logits"], z["masks"]):
        pm += oconf.generate_matrix(ofuse.miou_pred(torch.from_numpy(logit)[None]).numpy()[0], mask, 3)
    assert np.array_equal(pm, z["patch_cm"])
    assert any("MosaSegmentationic Test - Test tissue IoU (big mask)" in str(l) for l in z["log"])


def test_oeem_ensemble_matches_the_reference_script():
    """oracle/stitch.py::cam_ensemble / cam_to_32 vs prepare_seg_inputs.py:81-138 run with a stub forward_cam."""
    from oracle import stitch as ostitch
    z = load("oeem.npz")
    scales, side = [float(s) for s in z["scales"]], int(z["side"])
    for name in ("a.png", "b.png"):
        wh = tuple(int(v) for v in z[f"{name}_wh"])
        cams = [torch.from_numpy(z[f"{name}_cam{s}"]) for s in range(len(scales))]
        pos = [[tuple(int(v) for v in p) for p in z[f"{name}_pos{s}"]] for s in range(len(scales))]
        for literal in (True, False):
            ens = ostitch.cam_ensemble(cams, pos, scales, wh, side=side, literal=literal)
            got = ostitch.cam_to_32(ens, literal=literal)
            ref = z[f"{name}_ens32"]
            assert got.shape == ref.shape == (3, 32, 32)
            # literal = the same torch calls: bit for bit.  The restatement follows ATen's vectorised / CUDA kernel; the fixture's tiny
            # 4 x 4 CAMs take ATen's scalar CPU path (no fma), 1-2 float32 ulp away -- inside the 1e-5 gate of BASELINE.json
            assert np.array_equal(got, ref) if literal else np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


def test_tiling_positions_match_the_reference_function():
    """pistoseg_b200.oeem.online_cut_positions (host logic of the product) vs pyutils.online_cut_patches executed from the reference."""
    from pistoseg_b200 import oeem
    z = load("oeem.npz")
    for i, (h, w) in enumerate(z["tiling_cases"]):
        ref = [tuple(int(v) for v in p) for p in z[f"tiling_pos{i}"]]
        assert oeem.online_cut_positions(int(h), int(w), int(z["side"]), int(z["stride"])) == ref, (h, w)
    for name in ("a.png", "b.png"):
        w, h = (int(v) for v in z[f"{name}_wh"])
        got = oeem.multiscale_positions(w, h, int(z["side"]), int(z["stride"]), [float(s) for s in z["scales"]])
        for s, p in enumerate(got):
            assert p == [tuple(int(v) for v in q) for q in z[f"{name}_pos{s}"]]


def test_revise_masks_match_the_reference_script():
    """oracle/fuse.py::revise_masks(_to_original) vs infer_revise_masks.py:137-157: multiply, drop channel 0, argmax, mode-'P' resize
    (NEAREST) to the original size, background at the ORIGINAL resolution."""
    z = load("revise.npz")
    label = torch.from_numpy(z["label"])
    for head in ("pmask_rv", "pcam_rv", "cam_rv"):
        assert np.array_equal(ofuse.revise_masks(torch.from_numpy(z[head]), label), z[head + "_masks"])
    sizes = [tuple(int(v) for v in s) for s in z["sizes"]]
    bgs = [z[f"background{i}"] for i in range(len(sizes))]
    got = ofuse.revise_masks_to_original(torch.from_numpy(z["pmask_rv"]), label, sizes, bgs)
    for i, m in enumerate(got):
        assert np.array_equal(m, z[f"pmask_png{i}"]) and str(z[f"pmask_mode{i}"]) == "P"
    assert (z["pcam_rv_masks"][1] == np.argmax(np.concatenate([z["pcam_rv"][1, 1:]]) * z["label"][1, 1:, None, None], 0)).all()


def test_nearest_index_matches_live_pil():
    """The accumulated-double source index (oracle and product host code) vs Image.resize on mode-'P' images."""
    from PIL import Image
    from pistoseg_b200.postproc import pil_nearest_index
    for n_in in (48, 256, 31, 224):
        for n_out in list(range(1, 400, 7)) + [48, 256, 512, 1000]:
            a = (np.arange(n_in) % 251).astype(np.uint8)[None, :].repeat(2, 0)
            r = np.array(Image.fromarray(a, mode="P").resize((n_out, 2), resample=Image.BILINEAR))[0]
            assert np.array_equal(r, a[0][ofuse.nearest_resize_index(n_in, n_out)]), (n_in, n_out)
            assert np.array_equal(pil_nearest_index(n_in, n_out), ofuse.nearest_resize_index(n_in, n_out)), (n_in, n_out)
