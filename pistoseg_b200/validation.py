"""Validation hooks of the Lightning modules (``models/mosaic_module.py:127-255``, ``models/segmentation_module.py:117-232``)
on the GPU -- SURVEY.md 8(f) rank 2.

The reference's ``validation_step`` copies every tile's softmax to the host and overlap-adds it into per-(image, scale) numpy
canvases; ``validation_epoch_end`` normalises, resizes (float64 bilinear on the CPU), averages over scales and feeds the
result to a fresh ``mIoUMask``.  That is the arithmetic of ``segmentation_test.py:141-215``; here the same kernels serve it
(``stitch.BigMaskFuser``), the canvases stay in HBM and only the two C x C confusion matrices ever reach the host.

``BigMaskValidation`` is a mixin: a ``LightningModule`` (or any object with ``self.args``, ``self.valid_iou``, ``self.log`` and
``forward``) inherits the three hooks unchanged::

    class MosaicModule(BigMaskValidation, pl.LightningModule): ...

``pytorch_lightning`` itself is not needed (and not installed in the build image); the hooks only use what they are given.
"""
import os

import numpy as np
import torch

from .metrics import mIoUMask
from .stitch import BigMaskFuser


def parse_tile_name(name):
    """``{img}_{scale}_{y}_{x}-{label}.png`` (``mosaic_module.py:153-156``)."""
    parts = name.split("_")
    return parts[0], float(parts[1]), (int(parts[2]), int(parts[3].split("-")[0]))


class BigMaskValidation:
    """Hooks with the reference's names and signatures.  Image sizes and ground-truth masks are read the way the reference
    reads them (``<val_data>/../img/<idx>.png``, ``<val_data>/../mask/<idx>.png``); override ``_val_image_size`` /
    ``_val_gt_mask`` to feed them from elsewhere (tests do)."""

    # ---- file access, as mosaic_module.py:161,179,190 ----------------------------------------------------------------
    def _val_root(self):
        return "/".join(self.args.val_data.split("/")[:-1])

    def _val_image_size(self, image_idx):
        from PIL import Image
        w, h = Image.open(os.path.join(self._val_root(), "img", image_idx + ".png")).size
        return h, w

    def _val_gt_mask(self, image_idx):
        from PIL import Image
        return np.asarray(Image.open(os.path.join(self._val_root(), "mask", image_idx + ".png")))

    def _val_report_header(self):
        return getattr(self.args, "mosaic_data", None) or getattr(self.args, "pseudo_mask_dir", "")

    # ---- hooks ------------------------------------------------------------------------------------------------------
    def on_validation_epoch_start(self):
        if self.args.dataset == "wsss4luad":
            self._fusers = {}                      # image idx -> BigMaskFuser (device canvases, one per scale)

    def validation_step(self, batch, batch_idx):
        image_batch, mask_batch, name_batch, original_h_batch, original_w_batch = batch
        output = self(image_batch)
        self.valid_iou.update(output, mask_batch)  # = self.valid_iou(output, mask_batch) without the per-step host read-back of the two IoUs
        if self.args.dataset == "wsss4luad":
            groups = {}
            for j, name in enumerate(name_batch):
                idx, scale, pos = parse_tile_name(name)
                groups.setdefault((idx, scale), []).append((j, pos, (int(original_h_batch[j]), int(original_w_batch[j]))))
            for (idx, scale), items in groups.items():
                if idx not in self._fusers:
                    self._fusers[idx] = BigMaskFuser(self._val_image_size(idx), output.shape[1], output.device)
                sel = torch.tensor([t[0] for t in items], device=output.device)
                self._fusers[idx].add_tiles(output.index_select(0, sel).float(), scale, [t[1] for t in items], [t[2] for t in items])

    def validation_epoch_end(self, validation_step_outputs=None):
        names = ("Tumor", "Stroma", "Normal") if self.args.dataset == "wsss4luad" else ("Tumor", "Stroma", "Lymphocytic infiltrate", "Necrosis")
        big_mask_iou = None
        if self.args.dataset == "wsss4luad":
            big_mask_iou = mIoUMask()              # mosaic_module.py:187: default 3 classes
            dev = next(iter(self._fusers.values())).device if self._fusers else None
            for idx, fuser in self._fusers.items():
                gt = torch.from_numpy(np.ascontiguousarray(self._val_gt_mask(idx)).astype(np.uint8)).to(dev)
                # probs=True: argmax of the fused float64 probabilities, confusion over 0 <= gt < 3 (loss.py:17-24,55-67)
                fuser.finish(gt=gt, conf=big_mask_iou._acc(dev), bg_match=255, bg_label=0)
            self._fusers = {}
        tissue_iou = self.valid_iou.Tissue_Intersection_over_Union()
        bar = "\n" + "-" * 50
        print(bar)
        print("\nExperiment Settings")
        print(f"Dataset: \033[1;34m{self._val_report_header()}\033[0m")
        print(f"Log Path: \033[1;34m{getattr(self.args, 'log_path', '')}\033[0m")
        print(bar)
        print("\nValidation Result (Patch)" if big_mask_iou is not None else "\nValidation Result (Mask)")
        for n, v in zip(names, tissue_iou):
            print(f"{n} IoU: \033[1;35m{v:.4f}\033[0m")
        print(f"mIoU: \033[1;35m{self.valid_iou.Mean_Intersection_over_Union():.4f}\033[0m")
        print(f"fwIoU: \033[1;35m{self.valid_iou.Frequency_Weighted_Intersection_over_Union():.4f}\033[0m")
        print(bar)
        out = {}
        if big_mask_iou is not None:
            keys = ("tiou", "siou", "niou")
            for k, v in zip(keys, tissue_iou):
                out[f"validation_{k}_patch_epoch"] = v
            out["validation_miou_patch_epoch"] = self.valid_iou.Mean_Intersection_over_Union()
            out["validation_fwiou_patch_epoch"] = self.valid_iou.Frequency_Weighted_Intersection_over_Union()
            self.valid_iou.reset()
            big = big_mask_iou.Tissue_Intersection_over_Union()
            print(bar)
            print("\nValidation Result (Big Mask)")
            for n, v in zip(names, big):
                print(f"{n} IoU: \033[1;35m{v:.4f}\033[0m")
            print(f"mIoU: \033[1;35m{big_mask_iou.Mean_Intersection_over_Union():.4f}\033[0m")
            print(f"fwIoU: \033[1;35m{big_mask_iou.Frequency_Weighted_Intersection_over_Union():.4f}\033[0m")
            print(bar)
            for k, v in zip(keys, big):
                out[f"validation_{k}_mask_epoch"] = v
            out["validation_miou_mask_epoch"] = big_mask_iou.Mean_Intersection_over_Union()
            out["validation_fwiou_mask_epoch"] = big_mask_iou.Frequency_Weighted_Intersection_over_Union()
            self.last_big_mask_confusion = big_mask_iou.confusion_matrix
            big_mask_iou.reset()
        else:
            keys = ("tmr", "str", "lym", "nec")                     # mosaic_module.py:260-263
            for k, v in zip(keys, tissue_iou):
                out[f"validation_{k}_mask_epoch"] = v
            out["validation_miou_mask_epoch"] = self.valid_iou.Mean_Intersection_over_Union()
            out["validation_fwiou_mask_epoch"] = self.valid_iou.Frequency_Weighted_Intersection_over_Union()
            self.valid_iou.reset()
        prog = {"validation_miou_patch_epoch", "validation_fwiou_patch_epoch", "validation_miou_mask_epoch", "validation_fwiou_mask_epoch"}
        for k, v in out.items():
            self.log(k, v, prog_bar=k in prog)
        return out
