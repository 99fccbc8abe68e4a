"""The reference's pseudo-mask post-processing loop on the CPU, literal torch calls -- used as the timed CPU baseline
(``bench.py`` ``cpu_baseline`` / ``--impl reference``) and as the end-to-end label oracle.

Follows ``infer_pseudo_masks.py:118-154`` after the backbone: TTA merge of the per-view logits (ttach Merger 'mean',
generalised to multi-scale views by ``interpolate_tensor``), then PER TILE (the reference's Python loop):
32x32 logit export (:126), label parsing result ``patch_label`` (:130-134), ``get_mask_pred_and_entropy`` (:137).
PNG encoding / torch.save (:127,143-154) are excluded on both sides.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .fuse import get_mask_pred_and_entropy
from .tta import apply_code


def interpolate_tensor(tensor, target_shape):
    return F.interpolate(tensor.unsqueeze(0), target_shape, mode='bilinear')[0]


def pseudo_mask_batch(views, codes, size, present, bg, batch_size=32):
    """views: list of CPU f32 [N,C,h,w]; present [N,C] u8; bg [N,T,T] u8 (1 = background) or None.
    Returns (labels u8 [N,T,T], lowres f32 [N,C,32,32])."""
    N, C = views[0].shape[:2]
    labels = np.empty((N, size[0], size[1]), np.uint8)
    lowres = torch.empty((N, C, 32, 32), dtype=torch.float32)
    with torch.no_grad():
        for b0 in range(0, N, batch_size):
            b1 = min(b0 + batch_size, N)
            acc = None
            for v, code in zip(views, codes):
                y = apply_code(v[b0:b1], code)
                if tuple(y.shape[-2:]) != tuple(size):
                    y = F.interpolate(y, size, mode='bilinear')
                acc = y if acc is None else acc + y
            logit_pred = acc / len(views)
            for j in range(b1 - b0):
                n = b0 + j
                patch_logit_pred = logit_pred[j]
                lowres[n] = interpolate_tensor(patch_logit_pred, (32, 32))
                patch_label = [int(t) for t in present[n]] if present is not None else [1] * C
                tissue = np.where(np.asarray(bg[n]) != 0, 0.0, 127.0) if bg is not None else np.full(size, 127.0)
                mask_pred, _ = get_mask_pred_and_entropy(patch_logit_pred, tissue, patch_label)
                labels[n] = np.uint8(mask_pred)
    return labels, lowres
