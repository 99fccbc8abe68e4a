"""GPU product paths against fixtures made by EXECUTING the reference's script bodies (tests/golden/make_golden.py):
segmentation_test.py:125-227 (big-mask fusion + report) and infer_revise_masks.py:137-157 (revise-mask tail)."""
import os

import numpy as np
import pytest
import torch

from pistoseg_b200 import ops, postproc
from pistoseg_b200.metrics import mIoUMask
from pistoseg_b200.stitch import BigMaskFuser

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_big_mask_fuser_matches_the_reference_script(cuda):
    z = np.load(os.path.join(G, "bigmask.npz"))
    big = mIoUMask(num_classes=3)
    patch = mIoUMask(num_classes=3)
    B = int(z["batch"])
    names = [str(n) for n in z["names"]]
    for b0 in range(0, len(names), B):   # patch-level matrix, batch by batch as the script does (loss.py:55-67)
        patch(torch.from_numpy(z["logits"][b0:b0 + B]).to(cuda), torch.from_numpy(z["masks"][b0:b0 + B].astype(np.int64)).to(cuda))
    assert np.array_equal(patch.confusion_matrix, z["patch_cm"])
    agree, total = 0, 0
    for key in ("00", "01"):
        gt = z[f"gt_{key}"]
        fuser = BigMaskFuser(gt.shape, 3, cuda)
        by_scale = {}
        for name, logit, hw in zip(names, z["logits"], z["orig_hw"]):
            if name.split("_")[0] != key:
                continue
            s = float(name.split("_")[1])
            by_scale.setdefault(s, []).append((logit, (int(name.split("_")[2]), int(name.split("_")[3].split("-")[0])), (int(hw[0]), int(hw[1]))))
        for s, items in by_scale.items():
            fuser.add_tiles(torch.from_numpy(np.stack([i[0] for i in items])).to(cuda), s, [i[1] for i in items], [i[2] for i in items])
        fused = fuser.fused().cpu().numpy().transpose(1, 2, 0)
        ref = z[f"fused_{key}"]
        # float32 softmax per tile (CUDA expf vs the reference's CPU expf: 1 ulp), everything after it float64 in the reference's order
        assert np.abs(fused - ref).max() <= 1e-6
        out = fuser.finish(gt=torch.from_numpy(gt).to(cuda), conf=big._acc(cuda))
        lab = out["labels"].cpu().numpy()
        agree += int((lab == z[f"png_{key}"]).sum()); total += lab.size
    assert agree / total >= 0.9999
    assert np.abs(big.confusion_matrix - z["big_cm"]).sum() <= 2 * (total - agree)   # equal up to the (rare) near-tie pixels
    assert big.confusion_matrix.sum() == z["big_cm"].sum()


def test_revise_masks_to_original_matches_the_reference_script(cuda):
    z = np.load(os.path.join(G, "revise.npz"))
    label = torch.from_numpy(z["label"]).to(cuda)
    sizes = [tuple(int(v) for v in s) for s in z["sizes"]]
    bgs = [z[f"background{i}"] for i in range(len(sizes))]
    for head in ("pmask_rv", "pcam_rv", "cam_rv"):     # argmax of (x * label)[:, 1:] at the network resolution: integer work, exact
        got = postproc.revise_masks(torch.from_numpy(z[head]).to(cuda), label)
        assert np.array_equal(got.cpu().numpy(), z[head + "_masks"])
    got = postproc.revise_masks_to_original(torch.from_numpy(z["pmask_rv"]).to(cuda), label, sizes, bgs)
    for i, m in enumerate(got):
        assert m.dtype == torch.uint8 and tuple(m.shape) == sizes[i]
        assert np.array_equal(m.cpu().numpy(), z[f"pmask_png{i}"]), i
    # without backgrounds (BCSS branch, infer_revise_masks.py:189-206): resize only
    from oracle import fuse as ofuse
    got = postproc.revise_masks_to_original(torch.from_numpy(z["cam_rv"]).to(cuda), label, sizes, None)
    ref = ofuse.revise_masks_to_original(torch.from_numpy(z["cam_rv"]), torch.from_numpy(z["label"]), sizes, None)
    for a, b in zip(got, ref):
        assert np.array_equal(a.cpu().numpy(), b)


def test_revise_masks_to_png(cuda, tmp_path):
    from PIL import Image
    z = np.load(os.path.join(G, "revise.npz"))
    label = torch.from_numpy(z["label"]).to(cuda)
    sizes = [tuple(int(v) for v in s) for s in z["sizes"]]
    bgs = [z[f"background{i}"] for i in range(len(sizes))]
    heads = {"pmask": torch.from_numpy(z["pmask_rv"]).to(cuda), "pcam": torch.from_numpy(z["pcam_rv"]).to(cuda), "cam": torch.from_numpy(z["cam_rv"]).to(cuda)}
    names = [f"t{i}" for i in range(len(sizes))]
    paths = postproc.revise_masks_to_png(heads, label, names, sizes, str(tmp_path), bgs, dataset="wsss4luad")
    assert len(paths) == 15
    for i in range(len(sizes)):
        im = Image.open(os.path.join(str(tmp_path), "refine", "pmask", f"t{i}.png"))
        assert im.mode == "P" and im.size == (sizes[i][1], sizes[i][0])
        assert im.getpalette()[:12] == [0, 64, 128, 64, 128, 0, 243, 152, 0, 255, 255, 255]
        assert np.array_equal(np.array(im), z[f"pmask_png{i}"])
