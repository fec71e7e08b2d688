"""Host-side parameter derivation for the tile batch handed to libgrok_b200.so.

The kernels CONSUME quantisation constants (step size, inverse step, band bit depth, R/D weight);
in a Grok host they come from the codec (Quantizer.cpp:65-105, HTParams.cpp:164-249) and are passed
through unchanged.  This module provides the stand-alone equivalents used by
bench.py, the smoke test and the round-trip tests: the step-size formula of ISO 15444-1 E.1.1 and
synthesis-gain based weights computed numerically from the filter taps (not from the reference's
rounded tables, so R/D weights differ in the 4th digit from the reference's; parity tests pass the
reference's own numbers instead).
"""
import functools
import math

import numpy as np

from .binding import CompParams, MAX_RES


def _ceil_div_pow2(v, n):
    return (v + (1 << n) - 1) >> n


@functools.lru_cache(maxsize=None)
def synthesis_gains(reversible, levels):
    """L2 norms of the 1-D synthesis waveforms: (low[d], high[d]) for d = 0..levels.
    low[d]  : a low-pass sample after d levels of synthesis (low[0] = 1)
    high[d] : a high-pass sample of decomposition d+1 brought back to full resolution"""
    n = 1 << (levels + 6)

    def synth(lo, hi):
        x = np.zeros(2 * len(lo))
        x[0::2], x[1::2] = lo, hi
        m = len(x)
        e = np.arange(0, m, 2)
        o = np.arange(1, m, 2)
        if reversible:
            x[e] -= (x[(e - 1) % m] + x[(e + 1) % m]) / 4.0
            x[o] += (x[(o - 1) % m] + x[(o + 1) % m]) / 2.0
        else:
            K = 1.230174105
            x[e] *= K
            x[o] *= HIGH_SCALE
            for idx, c in ((e, -0.443506852), (o, -0.882911075), (e, 0.052980118), (o, 1.586134342)):
                x[idx] += c * (x[(idx - 1) % m] + x[(idx + 1) % m])
        return x

    low, high = [1.0], []
    for d in range(levels + 1):
        # high-pass impulse at decomposition d+1, then d low-pass synthesis steps
        size = n >> (d + 1)
        lo = np.zeros(size)
        hi = np.zeros(size)
        hi[size // 2] = 1.0
        x = synth(lo, hi)
        for _ in range(d):
            x = synth(x, np.zeros(len(x)))
        high.append(float(np.sqrt((x * x).sum())))
        lo[size // 2] = 1.0
        hi[:] = 0
        x = synth(lo, hi)
        for _ in range(d):
            x = synth(x, np.zeros(len(x)))
        low.append(float(np.sqrt((x * x).sum())))
    return tuple(low), tuple(high)


def dwt_norm(level, orient, reversible, maxlevels=32):
    low, high = synthesis_gains(bool(reversible), min(maxlevels, 12))
    level = min(level, len(high) - 1)
    if orient == 0:
        return low[level] * low[level]
    if orient in (1, 2):
        return low[level + 1] * high[level]
    return high[level] * high[level]


HIGH_SCALE = 2.0 / 1.230174105  # synthesis scaling of the high-pass branch

MCT_NORMS_REV = (1.732, 0.8292, 0.8292)   # norms of the RCT / ICT synthesis basis vectors
MCT_NORMS_IRREV = (1.732, 1.805, 1.573)


def band_quant(numres, prec, reversible, guard_bits=2, mct_norm=1.0, encoder=True, compno=0, ht=False):
    """Per band (host order): stepsize, inv_step, numbps, rd_weight for the stand-alone defaults.
    ht: the HTJ2K block coder -- one guard bit (j2k.cpp:1834), and on the decoder side the irreversible step size is scaled
    down to the MSB-aligned magnitudes the HT decoder returns (Quantizer.cpp:98-104)"""
    if ht:
        guard_bits = 1
    nb = 3 * numres - 2
    step = np.ones(nb, np.float32)
    inv = np.zeros(nb, np.uint32)
    nbps = np.zeros(nb, np.uint32)
    rdw = np.zeros(nb, np.float64)
    decomps = numres - 1
    for b in range(nb):
        resno = 0 if b == 0 else (b + 2) // 3
        orient = 0 if b == 0 else (b - 1) % 3 + 1
        level = decomps - resno  # dwt level of the band (T1Part1.cpp:114)
        if reversible:
            gain = 0 if orient == 0 else (1 if orient < 3 else 2)
            expn = prec + gain + 1
            mant = 0
            numbps_nominal = prec + gain
        else:
            g = dwt_norm(level, orient, False)
            delta = (1.0 / (1 << prec)) / g
            expn = 0
            while delta < 1.0:
                expn += 1
                delta *= 2.0
            mant = min(int(round(delta * 2048.0)) - 2048, 0x7FF)
            numbps_nominal = prec
        s = np.float32((1.0 + mant / 2048.0) * math.pow(2.0, numbps_nominal - expn))
        if not encoder:
            s = np.float32(s * np.float32(0.5))
        step[b] = s
        inv[b] = np.uint32(int(8192.0 / float(np.float32((1.0 + mant / 2048.0) * math.pow(2.0, numbps_nominal - expn))) + 0.5))
        nbps[b] = max(expn + guard_bits - 1, 1)
        if ht and not encoder and not reversible:
            step[b] = np.float32(step[b] / np.float32(1 << (30 - int(nbps[b]))))
        rdw[b] = (mct_norm * dwt_norm(level, orient, reversible)) * float(step[b])
    return step, inv, nbps, rdw


def comp_params(x0, y0, x1, y1, numres, reversible, prec, sgnd=0, cblk_expn=(6, 6), prc_expn=15, quant=None,
                mct_norm=1.0, encoder=True, guard_bits=2, ht=False):
    p = CompParams()
    p.cblk_sty = 0x40 if ht else 0
    p.x0, p.y0, p.x1, p.y1 = x0, y0, x1, y1
    p.numres = numres
    p.cblkw_expn, p.cblkh_expn = cblk_expn
    for r in range(MAX_RES):
        pe = prc_expn[r] if isinstance(prc_expn, (list, tuple)) else prc_expn
        if isinstance(pe, (list, tuple)):
            p.prcw_expn[r], p.prch_expn[r] = pe
        else:
            p.prcw_expn[r] = p.prch_expn[r] = pe
    p.qmfbid = 1 if reversible else 0
    p.prec = prec
    p.sgnd = sgnd
    p.dc_shift = 0 if sgnd else 1 << (prec - 1)
    step, inv, nbps, rdw = quant if quant is not None else band_quant(numres, prec, reversible, guard_bits, mct_norm, encoder, ht=ht)
    for b in range(3 * numres - 2):
        p.stepsize[b] = float(step[b])
        p.inv_step[b] = int(inv[b])
        p.band_numbps[b] = int(nbps[b])
        p.rd_weight[b] = float(rdw[b])
    return p


def image_tiles(width, height, numcomps, prec, reversible, tile=(None, None), numres=6, mct=None, rate_control=False,
                cblk_expn=(6, 6), sgnd=0, encoder=True, numres_decode=0, prc_expn=15, ht=False):
    """Tile grid of an image (origin 0,0, no sub-sampling) -> list of tile dicts for binding.Plan, in raster order."""
    tw = tile[0] or width
    th = tile[1] or height
    if mct is None:
        mct = 1 if numcomps >= 3 else 0
    norms = MCT_NORMS_REV if reversible else MCT_NORMS_IRREV
    tiles = []
    for ty in range(0, height, th):
        for tx in range(0, width, tw):
            comps = []
            for c in range(numcomps):
                mn = norms[c] if (mct and c < 3) else 1.0
                comps.append(comp_params(tx, ty, min(tx + tw, width), min(ty + th, height), numres, reversible, prec, sgnd,
                                         cblk_expn, prc_expn, mct_norm=mn, encoder=encoder, ht=ht))
            tiles.append({"comps": comps, "mct": mct, "rate_control": int(rate_control), "numres_decode": numres_decode})
    return tiles


def split_planes(image_planes, width, height, tile):
    """image_planes: list of [H,W] int32 arrays -> tile-major, component-minor list of contiguous planes."""
    tw = tile[0] or width
    th = tile[1] or height
    out = []
    for ty in range(0, height, th):
        for tx in range(0, width, tw):
            for pl in image_planes:
                out.append(np.ascontiguousarray(pl[ty:ty + th, tx:tx + tw], np.int32))
    return out


def join_planes(tile_planes, width, height, numcomps, tile):
    tw = tile[0] or width
    th = tile[1] or height
    out = [np.zeros((height, width), np.int32) for _ in range(numcomps)]
    i = 0
    for ty in range(0, height, th):
        for tx in range(0, width, tw):
            for c in range(numcomps):
                out[c][ty:ty + th, tx:tx + tw] = tile_planes[i]
                i += 1
    return out
