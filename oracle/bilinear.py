"""Bilinear resize, ``align_corners=False``, no antialias -- what ``interpolate_tensor`` does.

Reference: ``infer_pseudo_masks.py:89-90`` / ``segmentation_test.py:88-89``
(``F.interpolate(tensor.unsqueeze(0), target_shape, mode='bilinear')[0]``), the same call inline at
``OEEM/classification/prepare_seg_inputs.py:116,131,137`` and ``OEEM/classification/utils/generate_CAM.py:71,86``.

Two implementations:
  * :func:`interpolate_tensor` -- the literal reference call (torch-CPU ATen kernel);
  * :func:`bilinear_restated` -- the arithmetic written out (index / lambda computation of ATen's
    ``area_pixel_compute_source_index`` + ``compute_source_index_and_lambda``, and the FMA association the
    CPU kernel evaluates), which the CUDA kernels reproduce operation for operation.
``tests/test_oracle_bilinear.py`` checks that the two agree bit-for-bit.
"""
import numpy as np
import torch
import torch.nn.functional as F


def interpolate_tensor(tensor, target_shape):
    """Literal reference function (``infer_pseudo_masks.py:89-90``)."""
    return F.interpolate(tensor.unsqueeze(0), target_shape, mode='bilinear')[0]


def _fma(a, b, c):
    """Fused multiply-add with a single rounding, for float32 and float64 numpy arrays."""
    a = np.asarray(a); b = np.asarray(b); c = np.asarray(c)
    if a.dtype == np.float32:
        # The product of two float32 is exact in float64 (48 bits); the f64 sum is rounded once to 53 bits
        # and then to 24.  That double rounding differs from a true fma only when the f64 value sits exactly
        # on a float32 rounding midpoint (low 29 mantissa bits == 1 followed by 28 zeros): those (rare)
        # elements are redone in exact rational arithmetic.
        p = a.astype(np.float64) * b.astype(np.float64)
        c64 = c.astype(np.float64)
        r64 = p + c64
        r = r64.astype(np.float32)
        bits = np.ascontiguousarray(r64).view(np.uint64)
        hazard = ((bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)) & np.isfinite(r64)
        if hazard.any():
            # two-sum: was the f64 addition itself exact?  (then the single f64->f32 rounding is already right)
            t = r64 - p
            err = (p - (r64 - t)) + (c64 - t)
            hazard &= (err != 0)
        if hazard.any():
            from fractions import Fraction
            aa, bb, cc = (np.broadcast_to(v, r64.shape) for v in (a, b, c))
            for t in np.argwhere(hazard):
                t = tuple(t)
                exact = Fraction(float(aa[t])) * Fraction(float(bb[t])) + Fraction(float(cc[t]))
                r[t] = _round_fraction_f32(exact)
        return r
    return _fma64(a, b, c)


def _round_fraction_f32(fr):
    """Round an exact rational to float32, round-half-to-even."""
    from fractions import Fraction
    f = np.float32(float(fr))
    cands = {float(np.nextafter(f, np.float32(-np.inf))), float(f), float(np.nextafter(f, np.float32(np.inf)))}
    best = min(cands, key=lambda v: (abs(Fraction(v) - fr), int(np.float32(v).view(np.uint32)) & 1))
    return np.float32(best)


def _two_prod(a, b):
    # Veltkamp split + Dekker product: p + e == a*b exactly (no overflow assumed)
    p = a * b
    s = 134217729.0  # 2^27 + 1
    ta = s * a; ah = ta - (ta - a); al = a - ah
    tb = s * b; bh = tb - (tb - b); bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def _fma64(a, b, c):
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64))
    p, e = _two_prod(a, b)
    # two-sum p + c
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    # result = s + (err + e): correct to within the last-bit tie cases (sufficient: cross-checked vs ATen)
    return s + (err + e)


def source_index_and_lambda(in_size, out_size, dtype):
    """Per-output-index (i0, i1, lambda0, lambda1) exactly as ATen computes them (UpSample.h)."""
    t = np.dtype(dtype).type
    dst = np.arange(out_size)
    if in_size == out_size:
        i0 = dst.copy(); i1 = dst.copy()
        return i0, i1, np.ones(out_size, dtype), np.zeros(out_size, dtype)
    scale = t(t(in_size) / t(out_size))
    src = _fma(np.full(out_size, scale, dtype), (dst.astype(dtype) + t(0.5)), np.full(out_size, t(-0.5), dtype))
    src = np.maximum(src, t(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = np.clip(src - i0.astype(dtype), t(0), t(1)).astype(dtype)
    l0 = (t(1) - l1).astype(dtype)
    return i0, i1, l0, l1


def bilinear_restated(x, size):
    """x: numpy [..., H, W] float32/float64 -> [..., size[0], size[1]].

    out = fma(hl0, fma(wl0, a, wl1*b), hl1 * fma(wl0, c, wl1*d))   
    i.e. horizontal lerp on the two source rows first, then the vertical lerp.
    """
    x = np.asarray(x)
    dt = x.dtype
    H, W = x.shape[-2:]
    oh, ow = size
    hi0, hi1, hl0, hl1 = source_index_and_lambda(H, oh, dt)
    wi0, wi1, wl0, wl1 = source_index_and_lambda(W, ow, dt)
    # horizontal pass on every source row: R[..., r, ox]
    a = x[..., :, wi0]; b = x[..., :, wi1]
    R = _fma(np.broadcast_to(wl0, a.shape), a, (wl1 * b).astype(dt))
    r0 = R[..., hi0, :]; r1 = R[..., hi1, :]
    h0 = hl0[:, None]; h1 = hl1[:, None]
    return _fma(np.broadcast_to(h0, r0.shape), r0, (h1 * r1).astype(dt))
