"""Lightning validation hooks (models/mosaic_module.py:127-255, models/segmentation_module.py:117-232) through
pistoseg_b200.validation.BigMaskValidation, against the oracle's restatement of the same arithmetic."""
import types

import numpy as np
import pytest
import torch

from oracle import confusion as oconf
from oracle import stitch as ostitch
from pistoseg_b200.metrics import mIoUMask
from pistoseg_b200.validation import BigMaskValidation

pytestmark = pytest.mark.gpu


class _Module(BigMaskValidation):
    """Stands in for the LightningModule: forward() looks the tile's logits up, log() records."""

    def __init__(self, dataset, num_classes, device, sizes=None, gts=None):
        self.args = types.SimpleNamespace(dataset=dataset, val_data="/nowhere/val/img", mosaic_data="synthetic", log_path="-")
        self.valid_iou = mIoUMask(num_classes=num_classes, device=device)
        self.logged = {}
        self.sizes, self.gts = sizes, gts

    def __call__(self, x):
        return x  # the "images" of the synthetic batches are the logits

    def log(self, key, value, prog_bar=False):
        self.logged[key] = float(value)

    def _val_image_size(self, idx):
        return self.sizes[idx]

    def _val_gt_mask(self, idx):
        return self.gts[idx]


def _tiles(g, idx, hw, scale, P=256, stride=128, C=3):
    h, w = int(hw[0] * scale), int(hw[1] * scale)
    out = []
    for y in range(0, max(h - P, 0) + 1, stride):
        for x in range(0, max(w - P, 0) + 1, stride):
            oh, ow = min(P, h - y), min(P, w - x)
            out.append((f"{idx}_{scale}_{y}_{x}-[1, 0, 1].png", torch.randn((C, P, P), generator=g) * 3, scale, (y, x), (oh, ow)))
    return out


def test_wsss4luad_validation_epoch_matches_the_reference_arithmetic(cuda, capsys):
    g = torch.Generator().manual_seed(77)
    sizes = {"00": (300, 420), "01": (280, 350)}
    gts = {k: torch.randint(0, 4, hw, generator=g, dtype=torch.uint8).numpy() for k, hw in sizes.items()}
    tiles = []
    for idx, hw in sizes.items():
        for s in (1.0, 1.25):
            tiles += _tiles(g, idx, hw, s)
    perm = torch.randperm(len(tiles), generator=g).tolist()    # DataLoader order is not grouped by image
    tiles = [tiles[i] for i in perm]
    patch_gt = torch.randint(0, 4, (len(tiles), 256, 256), generator=g, dtype=torch.int64)

    mod = _Module("wsss4luad", 3, cuda, sizes, gts)
    mod.on_validation_epoch_start()
    B = 5
    for b0 in range(0, len(tiles), B):
        chunk = tiles[b0:b0 + B]
        batch = (torch.stack([t[1] for t in chunk]).to(cuda), patch_gt[b0:b0 + B].to(cuda), [t[0] for t in chunk],
                 torch.tensor([t[4][0] for t in chunk]), torch.tensor([t[4][1] for t in chunk]))
        mod.validation_step(batch, b0 // B)
    out = mod.validation_epoch_end([])
    text = capsys.readouterr().out
    assert "Validation Result (Patch)" in text and "Validation Result (Big Mask)" in text

    # patch metric: argmax(softmax(logits)) vs the patch masks, 0 <= gt < 3 (loss.py:55-67)
    cm_patch = np.zeros((3, 3))
    for (name, logits, *_), gt in zip(tiles, patch_gt):
        cm_patch += oconf.generate_matrix(torch.argmax(torch.softmax(logits, 0), 0).numpy(), gt.numpy(), 3)
    assert out["validation_miou_patch_epoch"] == oconf.mean_iou(cm_patch)
    assert out["validation_fwiou_patch_epoch"] == oconf.fw_iou(cm_patch)
    # big-mask metric: per image, stitched / normalised / resized / averaged float64 probabilities (mosaic_module.py:171-191)
    cm_big = np.zeros((3, 3))
    agree = []
    for idx, hw in sizes.items():
        ref = ostitch.big_mask_fuse([(t[1], t[2], t[3], t[4]) for t in tiles if t[0].startswith(idx + "_")], hw)
        pred = np.argmax(np.nan_to_num(ref), axis=2)
        cm_big += oconf.generate_matrix(pred, gts[idx], 3)
    got = mod.last_big_mask_confusion
    assert got.sum() == cm_big.sum()                                   # same pixels counted
    assert np.abs(got - cm_big).sum() <= 2e-4 * cm_big.sum()           # >= 99.99 % of the pixels agree (expf differs by 1 ulp)
    assert abs(out["validation_miou_mask_epoch"] - oconf.mean_iou(cm_big)) < 1e-3
    assert set(out) == {f"validation_{k}_{w}_epoch" for k in ("tiou", "siou", "niou", "miou", "fwiou") for w in ("patch", "mask")}
    assert mod.logged == {k: float(v) for k, v in out.items()}
    assert mod.valid_iou.confusion_matrix.sum() == 0                   # reset, as the reference does


def test_bcss_validation_epoch(cuda, capsys):
    g = torch.Generator().manual_seed(78)
    logits = torch.randn((6, 4, 224, 224), generator=g) * 2
    gt = torch.randint(0, 5, (6, 224, 224), generator=g, dtype=torch.int64)
    mod = _Module("bcss", 4, cuda)
    mod.on_validation_epoch_start()
    for b0 in (0, 3):
        mod.validation_step((logits[b0:b0 + 3].to(cuda), gt[b0:b0 + 3].to(cuda), ["a", "b", "c"], torch.zeros(3), torch.zeros(3)), b0)
    out = mod.validation_epoch_end([])
    cm = oconf.generate_matrix(torch.argmax(torch.softmax(logits, 1), 1).numpy(), gt.numpy(), 4)
    assert out["validation_miou_mask_epoch"] == oconf.mean_iou(cm)
    assert out["validation_fwiou_mask_epoch"] == oconf.fw_iou(cm)
    assert [out[f"validation_{k}_mask_epoch"] for k in ("tmr", "str", "lym", "nec")] == list(oconf.tissue_iou(cm))
    assert "Necrosis IoU" in capsys.readouterr().out
