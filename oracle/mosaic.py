"""Mosaic dataset synthesis, plan in -> pixels out.

Restates ``CropAndConcatDataset`` (``create_dataset.ipynb:249-374`` [cell 9]; BCSS variant
``create_dataset_bcss.ipynb:233-342`` [cell 8]) as the literal sequence of array operations the notebook performs once
every random decision has been fixed in a *plan*:

  create_one_image (:291-321)  patch_num x patch_num grid of patch_size crops; mask = tile label, bg>0 -> 3 (LUAD)
  create_mosaic    (:323-372)  per quadrant: Flip -> ShiftScaleRotate (cv2.warpAffine, oracle/warp_affine.py)
                               -> RandomCrop(hq, wq); paste the 4 quadrants.

The random stream itself (MT19937 + ``random`` consumed in albumentations-1.2.1 order) is not reproducible offline
(albumentations is not installed: PARITY UNPINNED for the sampling order); the contract is plan -> pixels, bit-exact.

Plan layout (mirrors ``pisto_mosaic_plan_t`` / ``pisto_mosaic_cell_t`` in include/pistoseg_b200.h):
  plan  : dict(split_h, split_w, quads=[4 x dict(flip, warp, crop_y, crop_x, M (forward 2x3 f64) or None)])
  cells : int array [4, patch_num*patch_num, 3] = (tile_id, crop_y, crop_x) in PADDED tile coordinates
  flip  : 0 none, 1 = cv2.flip code 0 (rows reversed), 2 = code 1 (cols reversed), 3 = code -1 (both)
"""
import numpy as np

from . import warp_affine as wa

BG_LABEL = 3


def pad_if_needed(tile, ps):
    """albumentations PadIfNeeded(min_height=ps, min_width=ps): centred BORDER_REFLECT_101 pad (``:264``)."""
    h, w = tile.shape[:2]
    top = int((ps - h) / 2.0) if h < ps else 0
    bottom = ps - h - top if h < ps else 0
    left = int((ps - w) / 2.0) if w < ps else 0
    right = ps - w - left if w < ps else 0
    if not (top or bottom or left or right):
        return tile
    yy = wa.reflect101(np.arange(-top, h + bottom), h)
    xx = wa.reflect101(np.arange(-left, w + right), w)
    return tile[yy][:, xx]


def create_one_image(cells_q, pool_imgs, pool_bgs, pool_labels, patch_num, ps):
    """``create_one_image`` with the tile choice and crop offsets taken from the plan."""
    H = W = patch_num * ps
    image = np.zeros((H, W, 3), np.uint8)
    mask = np.zeros((H, W), np.uint8)
    for i in range(patch_num):
        for j in range(patch_num):
            t, cy, cx = (int(v) for v in cells_q[i * patch_num + j])
            tile = pool_imgs[t]
            tile_mask = np.full(tile.shape[:2], pool_labels[t], np.uint8)
            if pool_bgs is not None:
                tile_mask[pool_bgs[t] > 0] = BG_LABEL
            tile = pad_if_needed(tile, ps)
            tile_mask = pad_if_needed(tile_mask, ps)
            image[i * ps:(i + 1) * ps, j * ps:(j + 1) * ps] = tile[cy:cy + ps, cx:cx + ps]
            mask[i * ps:(i + 1) * ps, j * ps:(j + 1) * ps] = tile_mask[cy:cy + ps, cx:cx + ps]
    return image, mask


def _flip(img, code):
    return img if code == 0 else wa.flip_u8(img, {1: 0, 2: 1, 3: -1}[code])


def synthesize(plan, cells, pool_imgs, pool_bgs, pool_labels, patch_num, ps, use_cv2=False):
    """One mosaic (image [H,W,3] u8, mask [H,W] u8) from a plan.  ``use_cv2`` swaps in the real cv2 calls."""
    H = W = patch_num * ps
    h, w = plan['split_h'], plan['split_w']
    sizes = [(h, w), (h, W - w), (H - h, w), (H - h, W - w)]
    outs = []
    for q in range(4):
        img, msk = create_one_image(cells[q], pool_imgs, pool_bgs, pool_labels, patch_num, ps)
        qd = plan['quads'][q]
        img = np.ascontiguousarray(_flip(img, qd['flip']))
        msk = np.ascontiguousarray(_flip(msk, qd['flip']))
        if qd['warp']:
            if use_cv2:
                import cv2
                M = np.asarray(qd['M'], np.float64)
                img = cv2.warpAffine(img, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
                msk = cv2.warpAffine(msk, M, (W, H), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_REFLECT_101)
            else:
                img = wa.warp_affine_u8(img, qd['M'], nearest=False)
                msk = wa.warp_affine_u8(msk, qd['M'], nearest=True)
        hq, wq = sizes[q]
        cy, cx = qd['crop_y'], qd['crop_x']
        outs.append((img[cy:cy + hq, cx:cx + wq], msk[cy:cy + hq, cx:cx + wq]))
    image = np.zeros((H, W, 3), np.uint8)
    mask = np.zeros((H, W), np.uint8)
    image[:h, :w] = outs[0][0]; image[:h, w:W] = outs[1][0]; image[h:H, :w] = outs[2][0]; image[h:H, w:W] = outs[3][0]
    mask[:h, :w] = outs[0][1]; mask[:h, w:W] = outs[1][1]; mask[h:H, :w] = outs[2][1]; mask[h:H, w:W] = outs[3][1]
    return image, mask
