"""Multi-view logit fusion -> label masking -> argmax -> background apply  (the fused hot path).

Restates, for a batch of tiles:
  * the TTA merge (``infer_pseudo_masks.py:96,121``; oracle/tta.py) generalised to views of different
    resolution (stride-8 multi-scale logits): de-augment by index remap, bilinear upsample to the tile size
    (``interpolate_tensor``, ``infer_pseudo_masks.py:89-90``), sequential fp32 sum in view order, one divide;
  * ``get_mask_pred_and_entropy`` (``infer_pseudo_masks.py:69-87``);
  * the 32x32 logit export (``infer_pseudo_masks.py:126``);
  * the revise-mask variant (``infer_revise_masks.py:137-143,154-155``);
  * ``mIoUMask.forward``'s prediction (``loss.py:55-60``).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .bilinear import bilinear_restated
from .tta import apply_code

LOGIT_MEAN, PROB_MEAN = 0, 1


def upsample_view(v, code, size, literal=False):
    """v: torch [N,C,h,w] f32 (as the model produced it on the augmented input) -> [N,C,T_h,T_w]."""
    d = apply_code(v, code)
    if tuple(d.shape[-2:]) == tuple(size):
        return d
    if literal:
        return F.interpolate(d, size, mode='bilinear')
    return torch.from_numpy(bilinear_restated(d.numpy(), size))


def fuse_views(views, codes, size, mode=LOGIT_MEAN, literal=False):
    """Sequential left-to-right fp32 sum of the de-augmented, upsampled views, then ``/ V`` (ttach Merger 'mean';
    ``OEEM/classification/prepare_seg_inputs.py:134-136``).  PROB_MEAN takes a channel softmax of every view first
    (``segmentation_test.py:150,173``)."""
    acc = None
    for v, code in zip(views, codes):
        u = upsample_view(v, code, size, literal)
        if mode == PROB_MEAN:
            u = torch.softmax(u, dim=1)
        acc = u if acc is None else acc + u
    return acc / len(views)


def get_mask_pred_and_entropy(patch_logit_pred, tissue, patch_label):
    """Restatement of ``infer_pseudo_masks.py:69-87`` (including the in-place mutation of the logits)."""
    if sum(patch_label) == 1:
        mask_pred = np.full((patch_logit_pred.shape[-2], patch_logit_pred.shape[-1]), patch_label.index(1))
        entropy = np.zeros_like(mask_pred)
    else:
        for i in range(len(patch_label)):
            if patch_label[i] == 0:
                patch_logit_pred[i, :, :] = -1e10
        patch_pos_pred = torch.softmax(patch_logit_pred, dim=0)
        entropy = -torch.sum(patch_pos_pred * torch.log(patch_pos_pred + 1e-10), dim=0).cpu().numpy()
        mask_pred = torch.argmax(patch_pos_pred, dim=0)
        mask_pred = mask_pred.cpu().numpy()
    mask_pred[tissue == 0] = len(patch_label)
    return mask_pred, entropy


def lowres_32(fused, literal=False):
    """``interpolate_tensor(patch_logit_pred, (32, 32))`` for a batch [N,C,T,T] (``infer_pseudo_masks.py:126``)."""
    if literal:
        return F.interpolate(fused, (32, 32), mode='bilinear')
    return torch.from_numpy(bilinear_restated(fused.numpy(), (32, 32)))


def pseudo_masks(fused, present, tissue_is_bg, want_entropy=False):
    """Batch driver for ``get_mask_pred_and_entropy``.

    fused [N,C,T,T] f32 torch (NOT mutated here: a clone is passed per tile), present [N,C] 0/1,
    tissue_is_bg [N,T,T] bool/u8 (1 where ``tissue == 0``) or None.  Returns uint8 labels [N,T,T] (and f32 entropy).
    """
    N, C, Th, Tw = fused.shape
    labels = np.empty((N, Th, Tw), np.uint8)
    ent = np.zeros((N, Th, Tw), np.float32) if want_entropy else None
    for n in range(N):
        lab = [int(v) for v in present[n]]
        tissue = np.full((Th, Tw), 127.0) if tissue_is_bg is None else np.where(np.asarray(tissue_is_bg[n]) != 0, 0.0, 127.0)
        m, e = get_mask_pred_and_entropy(fused[n].clone(), tissue, lab)
        labels[n] = m.astype(np.uint8)
        if want_entropy:
            ent[n] = e
    return (labels, ent) if want_entropy else labels


def miou_pred(logits, probs=False):
    """``mIoUMask.forward``'s prediction (``loss.py:55-60``): argmax(softmax(logits,1),1).byte() (argmax only if probs)."""
    if probs:
        return torch.argmax(logits, dim=1).byte()
    return torch.argmax(F.softmax(logits, dim=1), dim=1).byte()


def revise_masks(x, label, background=None, bg_value=3):
    """``infer_revise_masks.py:137-143,154-155``: (x * label[B,C+1,1,1])[:, 1:] -> argmax(dim=1) -> mask[background>0]=3."""
    y = (x * label.view(label.shape[0], -1, 1, 1))[:, 1:, :, :]
    m = torch.argmax(y, dim=1).numpy()
    if background is not None:
        m = m.copy()
        m[np.asarray(background) > 0] = bg_value
    return m


def revise_masks_to_original(x, label, original_hw, background=None, bg_value=3):
    """``infer_revise_masks.py:137-143,152-155`` for ONE head of ONE tile batch: ``(x * label)[:, 1:]`` -> ``argmax(dim=1)`` ->
    PIL mode-P resize to the original ``(w, h)`` (``resample=Image.BILINEAR`` is asked for, but PIL resizes mode 'P' -- like '1' --
    with NEAREST whatever the argument) -> ``mask[background > 0] = 3`` AT THE ORIGINAL RESOLUTION.
    x [B,C+1,S,S] torch; original_hw: list of (h, w); background: list of [h,w] arrays or None.  Returns a list of uint8 [h,w]."""
    from PIL import Image
    m = revise_masks(x, label)
    out = []
    for j, (h, w) in enumerate(original_hw):
        r = np.array(Image.fromarray(np.uint8(m[j]), mode='P').resize((int(w), int(h)), resample=Image.BILINEAR))
        if background is not None:
            r[np.asarray(background[j]) > 0] = bg_value
        out.append(r)
    return out


def nearest_resize_index(n_in, n_out):
    """Source index of PIL's NEAREST resize along one axis (Pillow ``ImagingScaleAffine``, the path ``Image.resize`` takes for
    NEAREST without rotation): ``xo = 0.5 * a; for x: idx[x] = int(xo); xo += a`` with ``a = n_in / n_out`` in double -- the
    ACCUMULATED double, not ``(x + 0.5) * a`` (they differ where the product is an integer).  Checked against ``Image.resize``
    in tests/test_oracle_golden.py."""
    a = n_in / n_out
    xo = a * 0.5
    out = np.empty(n_out, np.int64)
    for x in range(n_out):
        out[x] = int(xo)
        xo += a
    return np.minimum(out, n_in - 1)
