"""Fixed-point ``cv2.warpAffine`` (u8, INTER_LINEAR / INTER_NEAREST, BORDER_REFLECT_101) restated in numpy.

Call site in the reference: albumentations ``ShiftScaleRotate`` inside ``create_mosaic``
(``create_dataset.ipynb:323-330`` [cell 9]; same in ``create_dataset_bcss.ipynb`` [cell 8]); albumentations 1.2.1 calls
``cv2.getRotationMatrix2D((W/2-0.5, H/2-0.5), angle, scale)``, adds ``(dx*W, dy*H)`` to the translation and then
``cv2.warpAffine(img, M, (W, H), flags=INTER_LINEAR|INTER_NEAREST, borderMode=BORDER_REFLECT_101)``.

OpenCV is a third-party dependency of the reference (``opencv-python-headless==4.6.0.66``, ``environment.yaml:148``);
cv2 4.13.0 is installed in the build container and ``tests/test_oracle_warp.py`` pins this restatement to it
bit-for-bit.
"""
import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS


def get_rotation_matrix_2d(center, angle, scale):
    """cv2.getRotationMatrix2D in float64."""
    a = angle * (np.pi / 180.0)  # cv: angle *= CV_PI/180
    alpha = np.cos(a) * scale
    beta = np.sin(a) * scale
    cx, cy = center
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy],
                     [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


def shift_scale_rotate_matrix(H, W, angle, scale, dx, dy):
    """albumentations 1.2.1 ``F.shift_scale_rotate`` forward matrix."""
    M = get_rotation_matrix_2d((W / 2 - 0.5, H / 2 - 0.5), angle, scale)
    M[0, 2] += dx * W
    M[1, 2] += dy * H
    return M


def invert_affine(M):
    """The in-place inversion ``cv::warpAffine`` performs when WARP_INVERSE_MAP is not set (float64)."""
    M = np.array(M, np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2] = b1
    M[1, 2] = b2
    return M


def linear_weight_table():
    """The 32x32 table of 2x2 int16 bilinear weights OpenCV builds once (``initInterTab2D`` for INTER_LINEAR):
    float32 outer product * 32768, rounded half-to-even, each 2x2 block then forced to sum to 32768."""
    tab1 = np.empty((INTER_TAB_SIZE, 2), np.float32)
    sc = np.float32(1.0 / INTER_TAB_SIZE)
    for i in range(INTER_TAB_SIZE):
        x = np.float32(i) * sc
        tab1[i, 0] = np.float32(1.0) - x
        tab1[i, 1] = x
    itab = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, 2, 2), np.int16)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            isum = 0
            blk = np.empty((2, 2), np.int32)
            for k1 in range(2):
                vy = tab1[i, k1]
                for k2 in range(2):
                    v = np.float32(vy * tab1[j, k2])
                    # saturate_cast<short>(float): round half to even, then clamp (1.0*32768 -> 32767)
                    blk[k1, k2] = min(32767, max(-32768, int(np.rint(np.float32(v * np.float32(INTER_REMAP_COEF_SCALE))))))
                    isum += blk[k1, k2]
            if isum != INTER_REMAP_COEF_SCALE:
                # OpenCV's fix-up scans k1,k2 in [ksize/2, ksize/2+2) = [1,3): for ksize == 2 that is element [1][1]
                # of this block plus three slots of the NEXT (not yet written, zero) block.  Net effect: a short sum
                # (diff < 0) is added to [1][1] (it is the maximum unless it is negative, which cannot happen); an
                # excess (diff > 0) would be taken from a zero slot that is overwritten afterwards -- no effect.
                diff = isum - INTER_REMAP_COEF_SCALE
                if diff < 0:
                    blk[1, 1] -= diff
                elif blk[1, 1] <= 0:
                    blk[1, 1] -= diff
            itab[i, j] = blk.astype(np.int16)
    return itab


_ITAB = None


def _itab():
    global _ITAB
    if _ITAB is None:
        _ITAB = linear_weight_table()
    return _ITAB


def reflect101(p, n):
    """BORDER_REFLECT_101 index fold for arbitrary overshoot (``cv::borderInterpolate``)."""
    p = np.asarray(p, np.int64)
    if n == 1:
        return np.zeros_like(p)
    period = 2 * (n - 1)
    p = np.abs(p) % period
    return np.where(p >= n, period - p, p)


def fixed_point_coords(Minv, H, W, nearest):
    """Per-pixel fixed-point source coordinates (X, Y) as OpenCV's WarpAffineInvoker computes them."""
    Mi = np.asarray(Minv, np.float64).reshape(2, 3)
    x = np.arange(W, dtype=np.float64)
    y = np.arange(H, dtype=np.float64)
    adelta = np.rint(Mi[0, 0] * x * AB_SCALE).astype(np.int64)
    bdelta = np.rint(Mi[1, 0] * x * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // 2 if nearest else AB_SCALE // INTER_TAB_SIZE // 2
    X0 = np.rint((Mi[0, 1] * y + Mi[0, 2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((Mi[1, 1] * y + Mi[1, 2]) * AB_SCALE).astype(np.int64) + rd
    X = X0[:, None] + adelta[None, :]
    Y = Y0[:, None] + bdelta[None, :]
    return X, Y


def warp_affine_u8(img, M, nearest=False):
    """img: uint8 [H,W] or [H,W,ch]; M: forward 2x3 matrix; output size == input size."""
    img = np.asarray(img)
    H, W = img.shape[:2]
    X, Y = fixed_point_coords(invert_affine(M), H, W, nearest)
    if nearest:
        sx = np.clip(X >> AB_BITS, -32768, 32767)
        sy = np.clip(Y >> AB_BITS, -32768, 32767)
        return img[reflect101(sy, H), reflect101(sx, W)]
    X >>= (AB_BITS - INTER_BITS)
    Y >>= (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = X & (INTER_TAB_SIZE - 1)
    fy = Y & (INTER_TAB_SIZE - 1)
    w = _itab()[fy, fx].astype(np.int64)  # [H,W,2,2]
    x0 = reflect101(sx, W); x1 = reflect101(sx + 1, W)
    y0 = reflect101(sy, H); y1 = reflect101(sy + 1, H)
    src = img.astype(np.int64)
    if src.ndim == 2:
        src = src[..., None]
    acc = (src[y0, x0] * w[..., 0, 0, None] + src[y0, x1] * w[..., 0, 1, None]
           + src[y1, x0] * w[..., 1, 0, None] + src[y1, x1] * w[..., 1, 1, None])
    out = np.clip((acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS, 0, 255).astype(np.uint8)
    return out.reshape(img.shape)


def flip_u8(img, code):
    """cv2.flip: code 0 = around the x-axis (rows reversed), 1 = around the y-axis (cols reversed), -1 = both."""
    if code == 0:
        return img[::-1]
    if code == 1:
        return img[:, ::-1]
    return img[::-1, ::-1]
