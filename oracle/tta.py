"""Test-time-augmentation views: the ttach d4 group and its generalisation to (scale, flip) views.

Reference call site: ``infer_pseudo_masks.py:96,121``
(``tta.SegmentationTTAWrapper(model, tta.aliases.d4_transform(), merge_mode='mean')``).
``ttach==0.0.3`` (``environment.yaml:204``) is NOT vendored in the reference and is not installed here, so this
file restates its published behaviour -- PARITY UNPINNED for the d4 enumeration / de-augmentation order:

  d4_transform() = Compose([HorizontalFlip(), Rotate90([0, 90, 180, 270])])
  views          = itertools.product([False, True], [0, 90, 180, 270])      (in that order)
  augment        : x.flip(3) if hflip;  then torch.rot90(x, angle // 90, (2, 3))
  de-augment     : torch.rot90(y, ((360 - angle) % 360) // 90, (2, 3));  then y.flip(3) if hflip
  Merger('mean') : acc = y1; acc = acc + y2; ...; result = acc / n          (fp32, left to right)

View transform codes used across this repo (``pisto_view_t.xform`` in include/pistoseg_b200.h):
  code = k + 4 * hflip  meaning  deaug(y) = flip3_if_hflip(rot90(y, k, (2, 3)))
"""
import itertools
import numpy as np
import torch

D4_VIEWS = list(itertools.product([False, True], [0, 90, 180, 270]))


def augment(x, hflip, angle):
    if hflip:
        x = x.flip(3)
    return torch.rot90(x, angle // 90, (2, 3))


def deaug_code(hflip, angle):
    """The xform code that undoes ``augment(x, hflip, angle)`` on the model output."""
    k = ((360 - angle) % 360) // 90
    return k + 4 * int(bool(hflip))


def apply_code(y, code):
    """deaug(y) for an xform code; y is [..., H, W] torch tensor or numpy array."""
    k, hf = code & 3, (code >> 2) & 1
    if isinstance(y, np.ndarray):
        y = np.rot90(y, k, axes=(-2, -1))
        if hf:
            y = y[..., ::-1]
        return np.ascontiguousarray(y)
    y = torch.rot90(y, k, (-2, -1))
    if hf:
        y = y.flip(-1)
    return y.contiguous()


def code_affine(code, h_in, w_in):
    """(h_out, w_out, base, stride_i, stride_j): deaug(y)[i, j] == y.flat[base + i*stride_i + j*stride_j].

    This is the index map the CUDA kernels use instead of materialising flipped / rotated copies.
    """
    k, hf = code & 3, (code >> 2) & 1
    if k == 0:
        ho, wo, base, si, sj = h_in, w_in, 0, w_in, 1
    elif k == 1:
        ho, wo, base, si, sj = w_in, h_in, w_in - 1, -1, w_in
    elif k == 2:
        ho, wo, base, si, sj = h_in, w_in, (h_in - 1) * w_in + w_in - 1, -w_in, -1
    else:
        ho, wo, base, si, sj = w_in, h_in, (h_in - 1) * w_in, 1, -w_in
    if hf:
        base, sj = base + (wo - 1) * sj, -sj
    return ho, wo, base, si, sj


def d4_merge_mean(model, image):
    """SegmentationTTAWrapper.forward with merge_mode='mean' (ttach 0.0.3 semantics)."""
    acc = None
    for hflip, angle in D4_VIEWS:
        y = model(augment(image, hflip, angle))
        y = apply_code(y, deaug_code(hflip, angle))
        acc = y if acc is None else acc + y
    return acc / len(D4_VIEWS)


def scale_flip_views(scales, flips=(False, True)):
    """Generalised view list for the BASELINE configs: [(s, f) for s in scales for f in flips]."""
    return [(s, f) for s in scales for f in flips]
