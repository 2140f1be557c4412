"""Coefficient tables of PIL's antialiased LANCZOS resize for 8-bit images, re-derived from the algorithm
Pillow implements (libImaging/Resample.c: ``precompute_coeffs`` + ``normalize_coeffs_8bpc``), so the
depth-network feed image (deep_models.py:195-198 ``pil.resize(..., LANCZOS)``) can be produced on the
device bit-for-bit.  The tables depend only on (in_size, out_size); the resampling itself is integer
arithmetic in the CUDA kernels of csrc/depth_ops.cu.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
SUPPORT = 3.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def coeffs(in_size, out_size):
    """Returns (bounds int32 [out,2] = (first input index, tap count), kk int32 [out,ksize], ksize)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size        # (double)(in1 - in0) / outSize with float box
    filterscale = max(scale, 1.0)
    support = SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize
