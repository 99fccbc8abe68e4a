"""OEEM CAM ensemble (prepare_seg_inputs.py:96-138, generate_CAM.py:46-102) on the GPU vs the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import stitch as ostitch
from pistoseg_b200 import oeem

pytestmark = pytest.mark.gpu


def _literal_positions(h, w, im_size, stride):
    """pyutils.py:27-46, literally."""
    if h < im_size:
        h_ = np.array([0])
    else:
        h_ = np.arange(0, h - im_size + 1, stride)
        if h % stride != 0:
            h_ = np.append(h_, h - im_size)
    if w < im_size:
        w_ = np.array([0])
    else:
        w_ = np.arange(0, w - im_size + 1, stride)
        if w % stride != 0:
            w_ = np.append(w_, w - im_size)
    return [(int(i), int(j)) for i in h_ for j in w_]


def _case(g, wh, scales, C=3, side=224, stride=74):
    w, h = wh  # the reference's (rows, cols)
    pos = [_literal_positions(int(w * s), int(h * s), side, stride) for s in scales]
    cams = [torch.randn((len(p), C, 28, 28), generator=g) * 2 for p in pos]
    return cams, pos


@pytest.mark.parametrize("wh", [(300, 260), (150, 400), (224, 224)])
def test_cam_ensemble_32_and_labels(cuda, wh):
    g = torch.Generator().manual_seed(wh[0] + wh[1])
    scales = [1, 1.25, 1.5, 1.75, 2]                       # configuration_wsss4luad.yml:8
    cams, pos = _case(g, wh, scales)
    assert pos == [oeem.online_cut_positions(int(wh[0] * s), int(wh[1] * s), 224, 74) for s in scales]
    ref = ostitch.cam_ensemble(cams, pos, scales, wh)
    got = oeem.cam_ensemble([c.to(cuda) for c in cams], pos, scales, wh)
    # no softmax anywhere on this path: f32 upsample and f64 sums are evaluated in the reference's order -> bit-exact
    assert np.array_equal(got.cpu().numpy(), ref)
    assert np.array_equal(oeem.ensemble_32([c.to(cuda) for c in cams], pos, scales, wh).cpu().numpy(), ostitch.cam_to_32(ref))
    for big_label in (None, [1, 0, 1], [0, 0, 1]):
        lab = oeem.validation_labels([c.to(cuda) for c in cams], pos, scales, wh, big_label)
        assert np.array_equal(lab.cpu().numpy(), ostitch.cam_validation_labels(ref, big_label))


def test_small_image_uses_the_scaled_size(cuda):
    # w_ < side_length: the CAM is interpolated to the scaled image size instead of 224 (prepare_seg_inputs.py:99-105)
    g = torch.Generator().manual_seed(5)
    wh, scales = (120, 180), [1, 1.5, 2]
    cams, pos = _case(g, wh, scales)
    ref = ostitch.cam_ensemble(cams, pos, scales, wh)
    got = oeem.cam_ensemble([c.to(cuda) for c in cams], pos, scales, wh)
    assert np.array_equal(got.cpu().numpy(), ref)
