"""OEEM multi-scale CAM ensemble: ``OEEM/classification/prepare_seg_inputs.py:80-138``,
``OEEM/classification/utils/generate_CAM.py:46-102``, tiling of ``OEEM/classification/utils/pyutils.py:14-69``.

The classifier's ``forward_cam`` stays in PyTorch; everything after it -- f32 upsample of the stride-8 CAMs, float64
overlap-add per scale, normalisation, float64 resize to the image, mean over scales, float64 resize to 32 x 32 or the
masked argmax -- runs on the GPU through the same entry points as the big-mask path (``stitch.cam_ensemble``)."""
import numpy as np
import torch

from . import ops
from .stitch import cam_ensemble


def online_cut_positions(h, w, im_size, stride):
    """Top-left corners of ``online_cut_patches`` (``pyutils.py:27-46``), in its order (rows outer, columns inner)."""
    def axis(n):
        if n < im_size:
            return [0]
        a = list(range(0, n - im_size + 1, stride))
        if n % stride != 0:
            a.append(n - im_size)
        return a
    return [(i, j) for i in axis(h) for j in axis(w)]


def multiscale_positions(h, w, im_size, stride, scales):
    """``multiscale_online_crop`` (``pyutils.py:49-69``): PIL resizes to (int(w*s), int(h*s)); positions per scale."""
    return [online_cut_positions(int(h * s), int(w * s), im_size, stride) for s in scales]


def ensemble_32(cams_per_scale, positions_per_scale, scales, image_wh, side=224):
    """``prepare_seg_inputs.py:96-138``: CUDA float64 [C,32,32] (what the reference ``np.save``s)."""
    ens = cam_ensemble(cams_per_scale, positions_per_scale, scales, image_wh, side)
    return ops.upsample_bilinear(ens, (32, 32))


def validation_labels(cams_per_scale, positions_per_scale, scales, image_wh, big_label=None, side=224):
    """``generate_CAM.py:46-102``: classes absent from ``big_label`` -> -inf, ``argmax(axis=0)``; CUDA uint8 [w,h]."""
    ens = cam_ensemble(cams_per_scale, positions_per_scale, scales, image_wh, side)
    return ops.argmax_f64(ens, present=big_label, want_labels=False)["pred"]
