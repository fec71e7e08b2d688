"""Synthetic images of SURVEY.md section 8(d): smooth-plus-noise so that bit planes are neither empty nor
incompressible, plus uniform-random and constant stress variants."""
import numpy as np


def synthetic_planes(width, height, numcomps, prec, seed, kind="smooth", amp=None, sigma=None):
    """-> list of numcomps int32 [height,width] arrays with values in [0, 2^prec)."""
    rng = np.random.default_rng(seed)
    maxv = (1 << prec) - 1
    mid = (maxv + 1) / 2.0
    if amp is None:
        amp = 0.235 * (maxv + 1)
    if sigma is None:
        sigma = max(1.0, 0.023 * (maxv + 1)) if prec <= 8 else 0.0046 * (maxv + 1)
    out = []
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    for c in range(numcomps):
        if kind == "smooth":
            v = mid + amp * np.sin(x / (37.0 + 11 * c) + c) * np.cos(y / (53.0 + 7 * c))
            v = v + rng.standard_normal((height, width), dtype=np.float32) * sigma
        elif kind == "random":
            v = rng.integers(0, maxv + 1, (height, width)).astype(np.float32)
        else:
            v = np.full((height, width), mid, np.float32)
        out.append(np.clip(np.rint(v), 0, maxv).astype(np.int32))
    return out
