"""Stand-in for the part of ``ttach`` the reference uses (``infer_pseudo_masks.py:96``):
``tta.SegmentationTTAWrapper(model, tta.aliases.d4_transform(), merge_mode='mean')``.

The augmentation of the INPUT images stays in torch (it feeds the backbone, which stays in PyTorch); the
de-augmentation + merge of the model OUTPUTS is the library's fused kernel: nothing is flipped / rotated / summed as
separate tensors, the 8 raw outputs are read once through their de-augmentation index maps.
"""
import itertools

import torch

from . import _lib, ops


class _Aliases:
    @staticmethod
    def d4_transform():
        """Compose([HorizontalFlip(), Rotate90([0, 90, 180, 270])]) -> [(hflip, angle)] in ttach's product order."""
        return list(itertools.product([False, True], [0, 90, 180, 270]))

    @staticmethod
    def hflip_transform():
        return [(False, 0), (True, 0)]

    @staticmethod
    def multiscale_flip_transform(scales):
        """[(scale, hflip)] for scale in scales for hflip in (False, True) -- the BASELINE multi-scale + flip view set."""
        return [(s, f) for s in scales for f in (False, True)]


aliases = _Aliases()


def augment(image, hflip, angle):
    """ttach HorizontalFlip.apply_aug_image then Rotate90.apply_aug_image."""
    if hflip:
        image = image.flip(3)
    return torch.rot90(image, angle // 90, (2, 3))


def deaug_code(hflip, angle):
    """xform code (k + 4*hflip) undoing ``augment`` on the model output: rot90 by -angle, then hflip."""
    return ((360 - angle) % 360) // 90 + 4 * int(bool(hflip))


class SegmentationTTAWrapper(torch.nn.Module):
    def __init__(self, model, transforms, merge_mode="mean", output_mask_key=None):
        super().__init__()
        if merge_mode != "mean":
            raise _lib.PistoError("only merge_mode='mean' (what the reference uses) is implemented")
        self.model = model
        self.transforms = list(transforms)
        self.merge_mode = merge_mode
        self.output_key = output_mask_key

    def views(self, image, *args):
        """Raw model outputs for every augmented input + their de-augmentation codes (no merge yet)."""
        outs, codes = [], []
        for hflip, angle in self.transforms:
            y = self.model(augment(image, hflip, angle), *args)
            if self.output_key is not None:
                y = y[self.output_key]
            outs.append(y.float().contiguous())
            codes.append(deaug_code(hflip, angle))
        return outs, codes

    def forward(self, image, *args):
        outs, codes = self.views(image, *args)
        size = image.shape[-2:] if outs[0].shape[-2:] == image.shape[-2:] else None
        if size is None:
            k_odd = codes[0] & 1
            size = (outs[0].shape[-1], outs[0].shape[-2]) if k_odd else outs[0].shape[-2:]
        merged = ops.fuse_argmax_confusion(outs, codes, size, want_labels=False, want_fused=True)["fused"]
        if self.output_key is not None:
            return {self.output_key: merged}
        return merged
