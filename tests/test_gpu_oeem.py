"""OEEM CAM ensemble (prepare_seg_inputs.py:96-138, generate_CAM.py:46-102) on the GPU vs the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import stitch as ostitch
from pistoseg_b200 import oeem

pytestmark = pytest.mark.gpu


def _case(g, wh, scales, C=3, side=224, stride=74):
    w, h = wh  # the reference's (rows, cols)
    # tile positions: the product's host code, pinned against the reference's own online_cut_patches by
    # tests/test_oracle_golden.py::test_tiling_positions_match_the_reference_function (fixture oeem.npz)
    pos = [oeem.online_cut_positions(int(w * s), int(h * s), side, stride) for s in scales]
    cams = [torch.randn((len(p), C, 28, 28), generator=g) * 2 for p in pos]
    return cams, pos


@pytest.mark.parametrize("wh", [(300, 260), (150, 400), (224, 224)])
def test_cam_ensemble_32_and_labels(cuda, wh):
    g = torch.Generator().manual_seed(wh[0] + wh[1])
    scales = [1, 1.25, 1.5, 1.75, 2]                       # configuration_wsss4luad.yml:8
    cams, pos = _case(g, wh, scales)
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


def test_ensemble_matches_the_reference_script_fixture(cuda):
    """GPU path vs prepare_seg_inputs.py:81-138 executed on synthetic CAMs (tests/golden/oeem.npz): f32 upsample of the CAMs, f64
    overlap-add per scale, f64 resize, mean over scales, f64 resize to 32 x 32."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "oeem.npz"))
    scales, side, stride = [float(s) for s in z["scales"]], int(z["side"]), int(z["stride"])
    for name in ("a.png", "b.png"):
        wh = tuple(int(v) for v in z[f"{name}_wh"])
        cams = [torch.from_numpy(z[f"{name}_cam{s}"]).to(cuda) for s in range(len(scales))]
        pos = oeem.multiscale_positions(wh[0], wh[1], side, stride, scales)
        for s, p in enumerate(pos):
            assert p == [tuple(int(v) for v in q) for q in z[f"{name}_pos{s}"]]
        got = oeem.ensemble_32(cams, pos, scales, wh, side=side).cpu().numpy()
        ref = z[f"{name}_ens32"]
        # the fixture's 4 x 4 CAMs go through ATen's scalar CPU bilinear (no fma); the CUDA kernel equals ATen's vectorised / CUDA path
        assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
        assert np.array_equal(got, ostitch.cam_to_32(ostitch.cam_ensemble([c.cpu() for c in cams], pos, scales, wh, side=side)))

