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

from . import _lib, ops, philox

FLIP_NONE, FLIP_ROWS, FLIP_COLS, FLIP_BOTH = 0, 1, 2, 3   # cv2.flip codes none / 0 / 1 / -1

PLAN_DTYPE = np.dtype([("split_h", "<i4"), ("split_w", "<i4"), ("reserved", "<i4", (2,)),
                       ("quad", [("flip", "<i4"), ("warp", "<i4"), ("crop_y", "<i4"), ("crop_x", "<i4"), ("minv", "<f8", (6,))], (4,))])
CELL_DTYPE = np.dtype([("tile", "<i4"), ("cy", "<i2"), ("cx", "<i2")])
assert PLAN_DTYPE.itemsize == C.sizeof(_lib.MosaicPlan) and CELL_DTYPE.itemsize == C.sizeof(_lib.MosaicCell)


_COS_COEF = [1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5, 1.0]
_SIN_COEF = [1.0 / 355687428096000.0, -1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0,
             -1.0 / 6.0, 1.0]


def cos_sin_deg(angle):
    """cos and sin of an angle in DEGREES, the same fixed arithmetic as ``pisto_cos_sin_deg`` in csrc/mosaic_plan.cu (quadrant
    reduction in degrees, Taylor polynomials in Horner form with separate multiply and add): the device planner and this host
    planner agree bit for bit, which the platforms' libm cos / sin would not guarantee.  |error| < 5e-16 against libm."""
    angle = np.asarray(angle, np.float64)
    k = np.rint(angle / 90.0)
    r = angle + -(90.0 * k)
    a = r * (3.14159265358979323846 / 180.0)
    z = a * a
    pc = np.full_like(a, _COS_COEF[0])
    for c in _COS_COEF[1:]:
        pc = pc * z + c
    ps = np.full_like(a, _SIN_COEF[0])
    for c in _SIN_COEF[1:]:
        ps = ps * z + c
    ps = ps * a
    q = k.astype(np.int64) & 3
    co = np.where(q == 0, pc, np.where(q == 1, -ps, np.where(q == 2, -pc, ps)))
    si = np.where(q == 0, ps, np.where(q == 1, pc, np.where(q == 2, -ps, -pc)))
    return co, si


def rotation_matrix(center, angle, scale):
    """cv2.getRotationMatrix2D in float64 (same operation order; cos / sin from ``cos_sin_deg``)."""
    co, si = cos_sin_deg(angle)
    alpha, beta = float(co) * scale, float(si) * scale
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
        self._integral_host, self._integral_dev = {}, {}

    def __len__(self):
        return len(self.hw_host)

    def integral_offsets(self, ps):
        """Entry offset of every tile's (ph+1) x (pw+1) summed-area table, ph = max(h, ps)."""
        ph, pw = np.maximum(self.hw_host[:, 0], ps).astype(np.int64), np.maximum(self.hw_host[:, 1], ps).astype(np.int64)
        sz = (ph + 1) * (pw + 1)
        ioff = np.zeros(len(sz), np.int64)
        ioff[1:] = np.cumsum(sz[:-1])
        return ioff, int(sz.sum())

    def integral_host(self, ps):
        """16-bit summed-area tables of (bg > 0) over the PadIfNeeded-padded tiles (numpy restatement of
        bg_integral_kernel; modulo 2^16, which is exact for crops of fewer than 65536 pixels)."""
        if ps not in self._integral_host:
            ioff, total = self.integral_offsets(ps)
            flat = np.zeros(total, np.uint16)
            for t, b in enumerate(self.bg_host):
                padded = _pad_reflect101((np.asarray(b) > 0).astype(np.int64), ps)
                I = np.pad(padded.cumsum(0).cumsum(1), ((1, 0), (1, 0)))
                flat[ioff[t]:ioff[t] + I.size] = (I & 0xFFFF).astype(np.uint16).reshape(-1)
            self._integral_host[ps] = (ioff, flat)
        return self._integral_host[ps]

    def integral_device(self, ps):
        """The same tables built on the device by pisto_mosaic_bg_integral (cached per patch size)."""
        if ps not in self._integral_dev:
            ioff, total = self.integral_offsets(ps)
            device = self.dev["img"].device
            dev = device.index if device.index is not None else torch.cuda.current_device()
            ioff_d = torch.from_numpy(ioff).to(device)
            integral = torch.empty(total, dtype=torch.int16, device=device)
            lib = _lib.load()
            _lib.check(lib.pisto_mosaic_bg_integral(_lib.handle(dev), ops._ptr(self.dev["bg"]), ops._ptr(self.dev["off"]), ops._ptr(self.dev["hw"]),
                                                    ops._ptr(ioff_d), len(self), int(ps), ops._ptr(integral), ops._stream(dev)))
            self._integral_dev[ps] = (ioff_d, integral)
        return self._integral_dev[ps]


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


KEY_CELLS, KEY_QUADS = 0xC3110000, 0x51AD0000   # xor-ed into the high key word: independent streams per purpose


class MosaicPlanner:
    """Decision tables of ``CropAndConcatDataset`` (create_dataset.ipynb:273-372): plan(i) depends only on (seed, i).

    Per-cell decisions (source tile, crop origin, background rejection: 4 * patch_num^2 per mosaic) come from
    ``pisto_mosaic_plan_cells`` on the device (``cells_device``) or from the identical numpy arithmetic (``cells_host``);
    the per-quadrant decisions (split, flip, ShiftScaleRotate parameters -> inverse affine in float64, RandomCrop origin:
    16 numbers per mosaic) likewise: ``quads_device`` (``pisto_mosaic_plan_quads``) or ``quad_plans`` (numpy) -- the same Philox
    draws, the same float64 operation order and the same fixed cos / sin polynomial (``cos_sin_deg``), bit for bit.
    """

    def __init__(self, pool, patch_num, patch_size, seed=2022, reject_bg=False, bg_label=3,
                 p_flip=0.8, p_warp=0.8, shift_limit=0.0625, scale_limit=0.2, rotate_limit=45.0, max_tries=64):
        self.pool, self.pn, self.ps, self.seed = pool, patch_num, patch_size, int(seed)
        self.reject_bg, self.bg_label, self.max_tries = bool(reject_bg) and pool.bg_host is not None, bg_label, max_tries
        self.p_flip, self.p_warp = p_flip, p_warp
        self.shift_limit, self.scale_limit, self.rotate_limit = shift_limit, scale_limit, rotate_limit

    def _keys(self, purpose):
        return self.seed & 0xFFFFFFFF, ((self.seed >> 32) & 0xFFFFFFFF) ^ purpose

    # ---- per-cell decisions ---------------------------------------------------------------------------------------
    def cells_host(self, indices):
        """[N, 4, pn^2] CELL_DTYPE, numpy restatement of plan_cells_kernel."""
        idx = np.asarray(list(indices), np.uint64)
        N, pn2, ps, P = len(idx), self.pn * self.pn, self.ps, len(self.pool)
        k0, k1 = self._keys(KEY_CELLS)
        ilo = np.repeat(idx & np.uint64(0xFFFFFFFF), 4 * pn2)
        ihi = np.repeat(idx >> np.uint64(32), 4 * pn2)
        slot = np.tile(np.arange(4 * pn2, dtype=np.uint64), N)
        tile = np.zeros(N * 4 * pn2, np.int64); cy = np.zeros_like(tile); cx = np.zeros_like(tile)
        pending = np.arange(N * 4 * pn2)
        hw = self.pool.hw_host.astype(np.int64)
        if self.reject_bg:
            ioff, I = self.pool.integral_host(ps)
        for t in range(self.max_tries):
            r0, r1, r2, _ = philox.philox4x32_10(ilo[pending], ihi[pending], slot[pending], t, k0, k1)
            tl = philox.mulhi(r0, P)
            ph, pw = np.maximum(hw[tl, 0], ps), np.maximum(hw[tl, 1], ps)
            y, x = philox.mulhi(r1, ph - ps + 1), philox.mulhi(r2, pw - ps + 1)
            tile[pending], cy[pending], cx[pending] = tl, y, x
            if not self.reject_bg:
                break
            W1 = pw + 1
            base = ioff[tl]
            n = (I[base + (y + ps) * W1 + x + ps].astype(np.int64) - I[base + y * W1 + x + ps] - I[base + (y + ps) * W1 + x] + I[base + y * W1 + x]) & 0xFFFF
            pending = pending[10 * n * self.bg_label >= 8 * ps * ps]  # "background area is smaller than 80 %" fails: draw again
            if len(pending) == 0:
                break
        cells = np.zeros((N, 4, pn2), CELL_DTYPE)
        cells["tile"] = tile.reshape(N, 4, pn2); cells["cy"] = cy.reshape(N, 4, pn2); cells["cx"] = cx.reshape(N, 4, pn2)
        return cells

    def cells_device(self, first_index, index_stride, N):
        """CUDA uint8 view of MosaicCell[N, 4, pn^2] for mosaics first_index + k * index_stride (pisto_mosaic_plan_cells)."""
        device = self.pool.dev["img"].device
        dev = device.index if device.index is not None else torch.cuda.current_device()
        cells = torch.empty(N * 4 * self.pn * self.pn * CELL_DTYPE.itemsize, dtype=torch.uint8, device=device)
        ioff_d, integral = self.pool.integral_device(self.ps) if self.reject_bg else (None, None)
        lib = _lib.load()
        _lib.check(lib.pisto_mosaic_plan_cells(_lib.handle(dev), self.seed & 0xFFFFFFFFFFFFFFFF, int(first_index), int(index_stride), int(N), self.pn,
                                               self.ps, len(self.pool), ops._ptr(self.pool.dev["hw"]), ops._ptr(integral), ops._ptr(ioff_d),
                                               int(self.bg_label), int(self.max_tries), ops._ptr(cells), ops._stream(dev)))
        return cells

    def quads_device(self, first_index, index_stride, N):
        """CUDA uint8 view of MosaicPlan[N] for mosaics first_index + k * index_stride (pisto_mosaic_plan_quads)."""
        device = self.pool.dev["img"].device
        dev = device.index if device.index is not None else torch.cuda.current_device()
        plans = torch.empty(N * PLAN_DTYPE.itemsize, dtype=torch.uint8, device=device)
        lib = _lib.load()
        _lib.check(lib.pisto_mosaic_plan_quads(_lib.handle(dev), self.seed & 0xFFFFFFFFFFFFFFFF, int(first_index), int(index_stride), int(N), self.pn, self.ps,
                                               float(self.p_flip), float(self.p_warp), float(self.shift_limit), float(self.scale_limit),
                                               float(self.rotate_limit), ops._ptr(plans), ops._stream(dev)))
        return plans

    # ---- per-quadrant decisions -----------------------------------------------------------------------------------
    def quad_plans(self, indices):
        """[N] PLAN_DTYPE: split, and per quadrant flip code, warp flag + inverse affine, crop origin."""
        idx = np.asarray(list(indices), np.uint64)
        N, H = len(idx), self.pn * self.ps
        W = H
        k0, k1 = self._keys(KEY_QUADS)
        ilo, ihi = idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32)
        plans = np.zeros(N, PLAN_DTYPE)
        r = philox.philox4x32_10(ilo, ihi, 4, 0, k0, k1)
        h = (H * (philox.u53(r[0], r[1]) * 0.6 + 0.2)).astype(np.int64)   # create_dataset.ipynb:336
        w = (W * (philox.u53(r[2], r[3]) * 0.6 + 0.2)).astype(np.int64)
        h += h % 2; w += w % 2
        plans["split_h"], plans["split_w"] = h, w
        sizes = [(h, w), (h, W - w), (H - h, w), (H - h, W - w)]
        for q in range(4):
            rA = philox.philox4x32_10(ilo, ihi, q, 0, k0, k1)
            rB = philox.philox4x32_10(ilo, ihi, q, 1, k0, k1)
            rC = philox.philox4x32_10(ilo, ihi, q, 2, k0, k1)
            rD = philox.philox4x32_10(ilo, ihi, q, 3, k0, k1)
            flip = np.where(philox.u32(rA[0]) < self.p_flip, 1 + philox.mulhi(rA[1], 3), FLIP_NONE)   # albu.Flip: d in {-1, 0, 1}
            warp = philox.u32(rA[2]) < self.p_warp
            angle = -self.rotate_limit + 2 * self.rotate_limit * philox.u53(rB[0], rB[1])
            scale = (1 - self.scale_limit) + 2 * self.scale_limit * philox.u53(rB[2], rB[3])
            dx = -self.shift_limit + 2 * self.shift_limit * philox.u53(rC[0], rC[1])
            dy = -self.shift_limit + 2 * self.shift_limit * philox.u53(rC[2], rC[3])
            hq, wq = sizes[q]
            qd = plans["quad"][:, q]
            qd["flip"], qd["warp"] = flip, warp
            qd["crop_y"] = ((H - hq + 1) * philox.u53(rD[0], rD[1])).astype(np.int64)   # albumentations RandomCrop
            qd["crop_x"] = ((W - wq + 1) * philox.u53(rD[2], rD[3])).astype(np.int64)
            minv = invert_affine_batch(shift_scale_rotate_batch(H, W, angle, scale, dx, dy))
            qd["minv"] = np.where(warp[:, None], minv, 0.0)
            plans["quad"][:, q] = qd
        return plans

    def plans(self, indices):
        """(plans [N] PLAN_DTYPE, cells [N,4,pn^2] CELL_DTYPE) on the host."""
        indices = list(indices)
        return self.quad_plans(indices), self.cells_host(indices)

    def plan(self, i):
        p, c = self.plans([i])
        return p[0], c[0]


def shift_scale_rotate_batch(H, W, angle, scale, dx, dy):
    """``shift_scale_rotate_matrix`` over arrays, same float64 operation order -> [N, 2, 3]."""
    co, si = cos_sin_deg(angle)
    alpha, beta = co * scale, si * scale
    cx, cy = W / 2 - 0.5, H / 2 - 0.5
    M = np.empty((len(alpha), 2, 3), np.float64)
    M[:, 0, 0] = alpha; M[:, 0, 1] = beta; M[:, 0, 2] = (1 - alpha) * cx - beta * cy
    M[:, 1, 0] = -beta; M[:, 1, 1] = alpha; M[:, 1, 2] = beta * cx + (1 - alpha) * cy
    M[:, 0, 2] += dx * W
    M[:, 1, 2] += dy * H
    return M


def invert_affine_batch(M):
    """``invert_affine`` over [N, 2, 3], same float64 operation order -> [N, 6]."""
    M = M.copy()
    D = M[:, 0, 0] * M[:, 1, 1] - M[:, 0, 1] * M[:, 1, 0]
    with np.errstate(divide="ignore"):
        D = np.where(D != 0, 1.0 / D, 0.0)
    A11, A22 = M[:, 1, 1] * D, M[:, 0, 0] * D
    M[:, 0, 0] = A11; M[:, 0, 1] *= -D; M[:, 1, 0] *= -D; M[:, 1, 1] = A22
    b1 = -M[:, 0, 0] * M[:, 0, 2] - M[:, 0, 1] * M[:, 1, 2]
    b2 = -M[:, 1, 0] * M[:, 0, 2] - M[:, 1, 1] * M[:, 1, 2]
    M[:, 0, 2] = b1; M[:, 1, 2] = b2
    return M.reshape(len(M), 6)


def synthesize(pool, plans, cells, patch_num, patch_size, bg_label=3, packed=True):
    """plans [N] PLAN_DTYPE (numpy), cells [N,4,pn*pn] CELL_DTYPE (numpy) or the CUDA uint8 tensor of ``cells_device``
    -> (img u8 [N,S,S,3], mask u8 [N,S,S]) CUDA tensors."""
    device = pool.dev["img"].device
    p = plans if torch.is_tensor(plans) else torch.from_numpy(np.ascontiguousarray(plans).view(np.uint8).reshape(-1)).to(device)
    c = cells if torch.is_tensor(cells) else torch.from_numpy(np.ascontiguousarray(cells).view(np.uint8).reshape(-1)).to(device)
    return ops.mosaic_gather(pool.dev, p, c, patch_num, patch_size, bg_label, packed=packed)


def synthesize_range(pool, planner, first_index, index_stride, N, bg_label=3, host_quads=False):
    """Mosaics first_index + k * index_stride, k < N: quadrant plans and cells both planned on the device (host_quads=True: the
    quadrant plans from the numpy planner instead -- identical bits)."""
    if host_quads:
        quads = planner.quad_plans([first_index + k * index_stride for k in range(N)])
    else:
        quads = planner.quads_device(first_index, index_stride, N)
    return synthesize(pool, quads, planner.cells_device(first_index, index_stride, N), planner.pn, planner.ps, bg_label)


def export_dataset(pool, planner, out_dir, n_total, rank=0, world=1, chunk=4096, dataset="wsss4luad", writer=None, name_fmt="{:07d}.png"):
    """The export driver of create_dataset.ipynb:523-560 (``func(indexes)``): rank r of ``world`` synthesises the mosaics
    ``i = r (mod world)`` and writes ``<out_dir>/img/<i>.png`` (RGB) and ``<out_dir>/mask/<i>.png`` (mode 'P' with the dataset's
    palette), chunk by chunk: the GPU gathers chunk k + 1 while the host threads encode chunk k.  Returns the number written."""
    import os
    from . import io as pio
    os.makedirs(os.path.join(out_dir, "img"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "mask"), exist_ok=True)
    palette = pio.palette_for(dataset)
    own = writer is None
    writer = writer or pio.AsyncWriter()
    mine = len(range(rank, n_total, world))
    done = 0
    for k0 in range(0, mine, chunk):
        n = min(chunk, mine - k0)
        img, mask = synthesize_range(pool, planner, rank + k0 * world, world, n)
        img_h, mask_h = img.cpu().numpy(), mask.cpu().numpy()
        for k in range(n):
            i = rank + (k0 + k) * world
            writer.submit(pio.save_rgb_png, img_h[k], os.path.join(out_dir, "img", name_fmt.format(i)))
            writer.submit(pio.save_mask_png, mask_h[k], os.path.join(out_dir, "mask", name_fmt.format(i)), palette)
        done += n
    if own:
        writer.close()
    return done


def shard_indices(n_total, rank, world):
    """Rank r generates mosaics i = r (mod world), like the reference's 12 worker processes (create_dataset.ipynb:554)."""
    return range(rank, n_total, world)
