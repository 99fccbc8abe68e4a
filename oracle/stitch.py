"""WSSS4LUAD big-mask path of ``segmentation_test.py:141-215`` and the OEEM CAM ensemble
(``OEEM/classification/prepare_seg_inputs.py:96-138``, ``OEEM/classification/utils/generate_CAM.py:46-102``), restated.

Float work after the per-tile fp32 softmax / fp32 upsample is float64 on the CPU in the reference; so it is here.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .bilinear import bilinear_restated


def _resize64(x, size, literal):
    """x: numpy f64 [C,h,w] -> [C,H,W]  (``interpolate_tensor`` on a float64 CPU tensor)."""
    if tuple(x.shape[-2:]) == tuple(size):
        return x.copy()
    if literal:
        return F.interpolate(torch.from_numpy(x)[None], size, mode='bilinear')[0].numpy()
    return bilinear_restated(x, size)


def big_mask_fuse(tiles, image_hw, literal=False):
    """``segmentation_test.py:141-204`` for ONE image.

    tiles: iterable of (logits [C,Hp,Wp] f32 torch, scale float, (y, x), (orig_h, orig_w)).
    Returns the fused probabilities [H,W,C] float64 (``mask_pred`` after ``/= cnt`` at ``:204``).
    """
    h, w = image_hw
    canv, cnt = {}, {}
    for logits, scale, (y, x), (oh, ow) in tiles:
        out_ = logits[:, :oh, :ow]                                   # :145
        probs = torch.softmax(out_, dim=0).numpy().transpose(1, 2, 0)  # :150-153
        key = float(scale)
        if key not in canv:
            canv[key] = np.zeros((int(h * scale), int(w * scale), probs.shape[2]))  # :168-171
            cnt[key] = np.zeros((int(h * scale), int(w * scale), 1))
        canv[key][y:y + out_.shape[1], x:x + out_.shape[2], :] += probs  # :173
        cnt[key][y:y + out_.shape[1], x:x + out_.shape[2], :] += 1       # :174
    pred = None
    n = 0
    for key, mask in canv.items():                                   # :187-199
        with np.errstate(invalid='ignore', divide='ignore'):
            mask = mask / cnt[key]
        m = _resize64(np.ascontiguousarray(mask.transpose(2, 0, 1)), (h, w), literal).transpose(1, 2, 0)
        pred = m.copy() if pred is None else pred + m
        n += 1
    return pred / n                                                  # :204


def big_mask_labels(mask_pred, gt):
    """``segmentation_test.py:207-211``: (prediction used for the confusion matrix, label map with bg applied)."""
    pred = np.argmax(mask_pred, axis=2)
    out = pred.copy()
    out[gt == 3] = 3
    return pred.astype(np.uint8), out.astype(np.uint8)


def cam_ensemble(cams_per_scale, positions_per_scale, scales, image_wh, side=224, literal=False):
    """``prepare_seg_inputs.py:96-136`` for ONE image.  NOTE the reference names the image dims (w, h) =
    ``orig_img.shape[:2]`` (i.e. w is the ROW count); kept as is.

    cams_per_scale[s]: torch f32 [n_s, C, 28, 28] raw ``forward_cam`` scores; positions_per_scale[s]: [(y, x), ...].
    Returns ensemble_cam [C, w, h] float64 (after ``/= len(scales)``).
    """
    w, h = image_wh
    C = cams_per_scale[0].shape[1]
    ensemble = np.zeros((C, w, h))
    for s, scale in enumerate(scales):
        w_, h_ = int(w * scale), int(h * scale)
        ix = side if w_ >= side else w_
        iy = side if h_ >= side else h_
        cam = cams_per_scale[s]
        if literal:
            cam_list = F.interpolate(cam, (ix, iy), mode='bilinear', align_corners=False).numpy()
        else:
            cam_list = bilinear_restated(cam.numpy(), (ix, iy))
        sum_cam = np.zeros((C, w_, h_))
        sum_counter = np.zeros_like(sum_cam)
        for k in range(cam_list.shape[0]):
            y, x = positions_per_scale[s][k]
            sum_cam[:, y:y + side, x:x + side] += cam_list[k]
            sum_counter[:, y:y + side, x:x + side] += 1
        sum_counter[sum_counter < 1] = 1
        norm_cam = sum_cam / sum_counter
        ensemble += _resize64(norm_cam, (w, h), literal)
    ensemble /= len(scales)
    return ensemble


def cam_to_32(ensemble, literal=False):
    """``prepare_seg_inputs.py:137``."""
    return _resize64(ensemble, (32, 32), literal)


def cam_validation_labels(ensemble, big_label=None):
    """``generate_CAM.py:91-99``: absent classes -> -inf, argmax(axis=0)."""
    e = ensemble.copy()
    if big_label is not None:
        for k in range(e.shape[0]):
            if big_label[k] == 0:
                e[k, :, :] = -np.inf
    return e.argmax(axis=0)
