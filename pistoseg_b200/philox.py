"""Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), vectorised with numpy.

The same integer arithmetic as ``philox4x32_10`` in csrc/mosaic_plan.cu: the host evaluates the per-quadrant decisions of
a mosaic plan with it, the device the per-cell decisions, and ``tests/test_host_logic.py`` pins it to the published
known-answer vectors.  Counter-based: the 128 output bits are a pure function of (counter[4], key[2]).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable arrays / ints holding 32-bit values.  Returns four uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(*(np.asarray(v, np.uint64) & _MASK for v in (c0, c1, c2, c3, k0, k1)))
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2                    # 32 x 32 -> 64 bit products
        c0, c1, c2, c3 = (p1 >> _S32) ^ c1 ^ k0, p1 & _MASK, (p0 >> _S32) ^ c3 ^ k1, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def mulhi(r, n):
    """floor(r * n / 2^32): a uniform integer in [0, n) from 32 random bits (n < 2^32, array or scalar)."""
    return ((np.asarray(r, np.uint64) * np.asarray(n, np.uint64)) >> _S32).astype(np.int64)


def u53(a, b):
    """A double in [0, 1) with 53 random bits from two 32-bit words (the construction numpy's generators use)."""
    return ((np.asarray(a, np.uint64) >> np.uint64(5)) * np.uint64(67108864) + (np.asarray(b, np.uint64) >> np.uint64(6))).astype(np.float64) / 9007199254740992.0


def u32(a):
    """A double in [0, 1) from one 32-bit word (exact)."""
    return np.asarray(a, np.float64) / 4294967296.0
