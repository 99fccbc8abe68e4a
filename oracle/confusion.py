"""Confusion matrix and IoU report -- restates ``loss.py:8-67`` (class ``mIoUMask``) in numpy.

``tests/golden/make_golden.py`` imports the reference's own ``loss.py`` and stores its outputs; the tests check this
restatement (and the CUDA kernel) against those fixtures.
"""
import numpy as np


def generate_matrix(pred, gt, num_class, ignore_class=None):
    """``loss.py:17-24``: rows = ground truth, cols = prediction; pixels with gt outside [0, C) are dropped."""
    pred = np.asarray(pred); gt = np.asarray(gt)
    mask = (gt >= 0) & (gt < num_class)
    if ignore_class is not None:
        mask = mask & (gt != ignore_class)
    label = num_class * gt[mask].astype('int') + pred[mask]
    count = np.bincount(label, minlength=num_class ** 2)
    return count.reshape(num_class, num_class)


def tissue_iou(cm):
    """``loss.py:33-38``."""
    cm = np.asarray(cm, np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        iou = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
    iou[np.isnan(iou)] = 0
    return iou


def mean_iou(cm):
    """``loss.py:40-43``."""
    return np.mean(tissue_iou(cm))


def fw_iou(cm):
    """``loss.py:45-53``."""
    cm = np.asarray(cm, np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        freq = np.sum(cm, axis=1) / np.sum(cm)
        iu = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
    return (freq[freq > 0] * iu[freq > 0]).sum()
