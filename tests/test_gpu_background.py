"""pisto_get_background vs the oracle (cv2 gray / threshold + restated remove_small_objects): bit-exact."""
import numpy as np
import pytest
import torch

from oracle import background as obg
from pistoseg_b200 import background, ops

pytestmark = pytest.mark.gpu


def _tissue_like(rng, h, w):
    """pinkish texture with white holes of many sizes (some below, some above 50 px), thin bridges and border blobs"""
    img = rng.integers(90, 235, (h, w, 3), dtype=np.uint8)
    for _ in range(40):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(1, 14)
        yy, xx = np.ogrid[:h, :w]
        img[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = rng.integers(215, 256, 3, dtype=np.uint8)
    for _ in range(6):  # one-pixel-wide lines: 4-connectivity matters
        L = min(30, w)
        y0, x0 = rng.integers(0, h), rng.integers(0, w - L + 1)
        for k in range(L):
            img[min(h - 1, y0 + k // 2 * (k % 2)), x0 + k] = 255 if k % 7 else 180
    return img


@pytest.mark.parametrize("hw", [(224, 224), (97, 131), (1, 60), (300, 17)])
def test_matches_the_reference_arithmetic(cuda, hw):
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    for _ in range(3):
        img = _tissue_like(rng, *hw)
        assert np.array_equal(background.get_background(img), obg.get_background(img))


def test_gray_thresholds_every_colour_like_cv2(cuda):
    # every (r, g, b) with r, g in steps of 3 around the decision boundary: gray == 200 / 201 must split exactly as cv2 does
    r, g, b = np.meshgrid(np.arange(150, 256, 3), np.arange(150, 256, 3), np.arange(0, 256), indexing="ij")
    img = np.stack([r, g, b], -1).astype(np.uint8).reshape(-1, 256, 3)
    got = ops.get_background(torch.from_numpy(img).to(cuda), min_size=0).cpu().numpy()
    import cv2
    assert np.array_equal(got, (cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) > 200).astype(np.uint8) * 255)


def test_batch_components_do_not_leak_across_images(cuda):
    # a 40-pixel blob at the bottom of image 0 and a 40-pixel blob at the top of image 1 must not merge into one 80-pixel one
    a = np.zeros((2, 32, 32, 3), np.uint8)
    a[0, 28:32, 0:10] = 255
    a[1, 0:4, 0:10] = 255
    a[1, 10:20, 10:20] = 255   # 100 pixels: kept
    got = background.get_background_batch(torch.from_numpy(a).to(cuda)).cpu().numpy()
    want = np.stack([obg.get_background(a[0]), obg.get_background(a[1])])
    assert np.array_equal(got, want) and got[0].sum() == 0 and got[1].sum() == 100 * 255


def test_full_white_and_snake(cuda):
    img = np.full((224, 224, 3), 255, np.uint8)
    assert np.array_equal(background.get_background(img), obg.get_background(img))
    snake = np.zeros((64, 64, 3), np.uint8)
    for y in range(0, 64, 2):
        snake[y, :] = 255
        snake[min(y + 1, 63), 63 if (y // 2) % 2 == 0 else 0] = 255   # one long 4-connected path
    assert np.array_equal(background.get_background(snake), obg.get_background(snake))
