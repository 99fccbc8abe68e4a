"""Mosaic dataset synthesis on the GPU: host-side plan generator + tile pool + the gather call.

Mirrors ``CropAndConcatDataset`` of the reference (``create_dataset.ipynb:249-374`` [cell 9], BCSS variant
``create_dataset_bcss.ipynb:233-342`` [cell 8]).  The reference draws its random decisions from MT19937 / ``random`` in
albumentations-1.2.1 call order, which cannot be reproduced without that library; here every decision of mosaic ``i``
comes from a counter-based Philox stream keyed on ``(seed, i)`` (the reference keys on ``2022 + 2022*i``,
``create_dataset.ipynb:274-275``), so mosaic ``i`` is identical for any GPU count and any batch split.  Pixels given a
plan are bit-exact with cv2 / the oracle (tests/test_gpu_mosaic.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops

FLIP_NONE, FLIP_ROWS, FLIP_COLS, FLIP_BOTH = 0, 1, 2, 3   # cv2.flip codes none / 0 / 1 / -1

PLAN_DTYPE = np.dtype([("split_h", "<i4"), ("split_w", "<i4"), ("reserved", "<i4", (2,)),
                       ("quad", [("flip", "<i4"), ("warp", "<i4"), ("crop_y", "<i4"), ("crop_x", "<i4"), ("minv", "<f8", (6,))], (4,))])
CELL_DTYPE = np.dtype([("tile", "<i4"), ("cy", "<i2"), ("cx", "<i2")])
assert PLAN_DTYPE.itemsize == C.sizeof(_lib.MosaicPlan) and CELL_DTYPE.itemsize == C.sizeof(_lib.MosaicCell)


def rotation_matrix(center, angle, scale):
    """cv2.getRotationMatrix2D in float64 (same operation order)."""
    a = angle * (np.pi / 180.0)
    alpha, beta = np.cos(a) * scale, np.sin(a) * scale
    cx, cy = center
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


def shift_scale_rotate_matrix(H, W, angle, scale, dx, dy):
    """albumentations 1.2.1 ShiftScaleRotate forward matrix (create_dataset.ipynb:327)."""
    M = rotation_matrix((W / 2 - 0.5, H / 2 - 0.5), angle, scale)
    M[0, 2] += dx * W
    M[1, 2] += dy * H
    return M


def invert_affine(M):
    """The float64 inversion cv::warpAffine applies to a forward matrix."""
    M = np.array(M, np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11; M[0, 1] *= -D; M[1, 0] *= -D; M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2] = b1; M[1, 2] = b2
    return M


class TilePool:
    """Single-label source tiles packed into flat device buffers (variable tile sizes allowed)."""

    def __init__(self, images, labels, bg_masks=None, device="cuda"):
        hw = np.array([im.shape[:2] for im in images], np.int32)
        off = np.zeros(len(images), np.int64)
        off[1:] = np.cumsum(hw[:-1, 0].astype(np.int64) * hw[:-1, 1])
        self.hw_host, self.off_host = hw, off
        self.labels_host = np.asarray(labels, np.uint8)
        self.bg_host = bg_masks
        self.dev = {
            "img": torch.from_numpy(np.concatenate([np.ascontiguousarray(im).reshape(-1) for im in images])).to(device),
            "bg": torch.from_numpy(np.concatenate([np.ascontiguousarray(b).reshape(-1) for b in bg_masks])).to(device) if bg_masks is not None else None,
            "off": torch.from_numpy(off).to(device), "hw": torch.from_numpy(hw).to(device),
            "label": torch.from_numpy(self.labels_host).to(device),
        }
        self._bg_integral = None

    def __len__(self):
        return len(self.hw_host)

    def bg_value_sum(self, t, cy, cx, ps, bg_label=3):
        """sum(tile_mask[tile_mask == 3]) of a crop (the reference's rejection statistic, create_dataset.ipynb:314),
        in padded coordinates, via per-tile integral images."""
        if self.bg_host is None:
            return 0
        if self._bg_integral is None:
            self._bg_integral = {}
        if t not in self._bg_integral:
            from_pad = _pad_reflect101((self.bg_host[t] > 0).astype(np.int64), ps)
            self._bg_integral[t] = np.pad(from_pad.cumsum(0).cumsum(1), ((1, 0), (1, 0)))
        I = self._bg_integral[t]
        n = I[cy + ps, cx + ps] - I[cy, cx + ps] - I[cy + ps, cx] + I[cy, cx]
        return int(n) * bg_label


def _pad_reflect101(a, ps):
    h, w = a.shape[:2]
    top = int((ps - h) / 2.0) if h < ps else 0
    bottom = ps - h - top if h < ps else 0
    left = int((ps - w) / 2.0) if w < ps else 0
    right = ps - w - left if w < ps else 0

    def refl(p, n):
        if n == 1:
            return np.zeros_like(p)
        period = 2 * (n - 1)
        p = np.abs(p) % period
        return np.where(p >= n, period - p, p)
    return a[refl(np.arange(-top, h + bottom), h)][:, refl(np.arange(-left, w + right), w)]


class MosaicPlanner:
    """Decision table generator: plan(i) depends only on (seed, i)."""

    def __init__(self, pool, patch_num, patch_size, seed=2022, reject_bg=False, bg_label=3,
                 p_flip=0.8, p_warp=0.8, shift_limit=0.0625, scale_limit=0.2, rotate_limit=45.0):
        self.pool, self.pn, self.ps, self.seed = pool, patch_num, patch_size, seed
        self.reject_bg, self.bg_label = reject_bg, bg_label
        self.p_flip, self.p_warp = p_flip, p_warp
        self.shift_limit, self.scale_limit, self.rotate_limit = shift_limit, scale_limit, rotate_limit

    def plan(self, i):
        rng = np.random.Generator(np.random.Philox(key=[self.seed, int(i)]))
        pn, ps = self.pn, self.ps
        H = W = pn * ps
        plan = np.zeros((), PLAN_DTYPE)
        cells = np.zeros((4, pn * pn), CELL_DTYPE)
        for q in range(4):  # create_one_image x 4 (create_dataset.ipynb:283)
            for c in range(pn * pn):
                while True:
                    t = int(rng.integers(0, len(self.pool)))
                    th, tw = (int(v) for v in self.pool.hw_host[t])
                    ph, pw = max(th, ps), max(tw, ps)
                    # albumentations RandomCrop: y1 = int((H - h + 1) * r)
                    cy = int((ph - ps + 1) * rng.random()); cx = int((pw - ps + 1) * rng.random())
                    # "background area is smaller than 80 %" -- the reference sums label VALUES (3 per bg pixel)
                    if not self.reject_bg or self.pool.bg_value_sum(t, cy, cx, ps, self.bg_label) < ps * ps * 0.8:
                        break
                cells[q, c] = (t, cy, cx)
        h = int(H * (rng.random() * 0.6 + 0.2)); w = int(W * (rng.random() * 0.6 + 0.2))  # create_dataset.ipynb:336
        h += h % 2; w += w % 2
        plan["split_h"], plan["split_w"] = h, w
        sizes = [(h, w), (h, W - w), (H - h, w), (H - h, W - w)]
        for q in range(4):
            qd = plan["quad"][q]
            qd["flip"] = int(rng.integers(1, 4)) if rng.random() < self.p_flip else FLIP_NONE  # albu.Flip: d in {-1, 0, 1}
            if rng.random() < self.p_warp:
                angle = rng.uniform(-self.rotate_limit, self.rotate_limit)
                scale = rng.uniform(1 - self.scale_limit, 1 + self.scale_limit)
                dx = rng.uniform(-self.shift_limit, self.shift_limit); dy = rng.uniform(-self.shift_limit, self.shift_limit)
                qd["warp"] = 1
                qd["minv"] = invert_affine(shift_scale_rotate_matrix(H, W, angle, scale, dx, dy)).reshape(-1)
            hq, wq = sizes[q]
            qd["crop_y"] = int((H - hq + 1) * rng.random()); qd["crop_x"] = int((W - wq + 1) * rng.random())
        return plan, cells

    def plans(self, indices):
        ps, cs = zip(*(self.plan(i) for i in indices))
        return np.stack(ps), np.stack(cs)


def synthesize(pool, plans, cells, patch_num, patch_size, bg_label=3):
    """plans [N] PLAN_DTYPE, cells [N,4,pn*pn] CELL_DTYPE (numpy) -> (img u8 [N,S,S,3], mask u8 [N,S,S]) CUDA tensors."""
    device = pool.dev["img"].device
    p = torch.from_numpy(np.ascontiguousarray(plans).view(np.uint8).reshape(-1)).to(device)
    c = torch.from_numpy(np.ascontiguousarray(cells).view(np.uint8).reshape(-1)).to(device)
    return ops.mosaic_gather(pool.dev, p, c, patch_num, patch_size, bg_label)


def shard_indices(n_total, rank, world):
    """Rank r generates mosaics i = r (mod world), like the reference's 12 worker processes (create_dataset.ipynb:554)."""
    return range(rank, n_total, world)
