"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile.ref
from /root/reference).  Run here (where /root/reference exists); the fixtures are committed so that the
oracle can be checked against the reference's own outputs wherever the reference is absent.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _libs import (aligned, ref, ref_t1_encode, ref_t1_decode, ref_encode_image, ref_decode_image, ref_t1_encode_sty,  # noqa: E402
                   ref_t1_decode_segs, segments_from_passes, random_pass_tables)
from grokimagecompression_b200.synth import synthetic_planes  # noqa: E402


def t1_vectors():
    rng = np.random.default_rng(20261018)
    out = {}
    shapes = [(64, 64), (32, 32), (17, 13), (5, 4), (1, 1), (64, 7), (3, 64), (33, 64)]
    for i, (w, h) in enumerate(shapes * 3):
        amp = [0.7, 25.0, 3000.0][i // len(shapes)]
        orient = i % 4
        q = (np.rint(rng.laplace(0, amp, (h, w))).astype(np.int32) * 64 + rng.integers(0, 64, (h, w))).astype(np.int32)
        norms = np.array([1.732, 1.805, 1.573])
        step, lvl, comp = 0.03125 * (1 + i % 3), i % 5, i % 3
        data, numbps, rates, dists = ref_t1_encode(q, orient, comp, lvl, 0, step, norms, True)
        w2 = ref().ref_dwt_norm(lvl, orient, 0)
        out[f"blk{i}_q"] = q
        out[f"blk{i}_meta"] = np.array([orient, numbps, len(rates)], np.int64)
        out[f"blk{i}_wbase"] = np.array([(norms[comp] * w2) * step])
        out[f"blk{i}_data"] = np.frombuffer(data, np.uint8)
        out[f"blk{i}_rates"] = rates
        out[f"blk{i}_dists"] = dists
        if len(rates):
            k = max(1, len(rates) // 2)
            out[f"blk{i}_dec_full"] = ref_t1_decode(data, len(rates), numbps, orient, w, h)
            out[f"blk{i}_dec_half"] = ref_t1_decode(data[:int(rates[k - 1])], k, numbps, orient, w, h)
    out["count"] = np.array([len(shapes) * 3])
    np.savez_compressed(os.path.join(HERE, "t1_blocks.npz"), **out)


def t1_style_vectors():
    """code-block style switches: the reference's bytes, rates, termination flags and segment-wise decodes"""
    rng = np.random.default_rng(20261019)
    out = {}
    styles = [1, 2, 4, 8, 16, 32, 1 | 4, 1 | 16, 1 | 4 | 16, 2 | 8 | 32, 63, 1 | 2 | 8, 4 | 16, 1 | 32]
    shapes = [(64, 64), (32, 32), (17, 13), (64, 7), (33, 64)]
    n = 0
    for j, sty in enumerate(styles):
        for (w, h) in (shapes[j % len(shapes)], shapes[(j + 2) % len(shapes)]):
            amp = [25.0, 3000.0, 60000.0][n % 3]
            orient = n % 4
            q = (np.rint(rng.laplace(0, amp, (h, w))).astype(np.int64) * 64 + rng.integers(0, 64, (h, w))).astype(np.int32)
            norms = np.array([1.732, 1.805, 1.573])
            step, lvl, comp = 0.03125 * (1 + n % 3), n % 5, n % 3
            data, numbps, rates, dists, terms = ref_t1_encode_sty(q, orient, sty, comp, lvl, 0, step, norms, True)
            out[f"blk{n}_q"] = q
            out[f"blk{n}_meta"] = np.array([orient, numbps, len(rates), sty], np.int64)
            out[f"blk{n}_wbase"] = np.array([(norms[comp] * ref().ref_dwt_norm(lvl, orient, 0)) * step])
            out[f"blk{n}_data"] = np.frombuffer(data, np.uint8)
            out[f"blk{n}_rates"], out[f"blk{n}_dists"], out[f"blk{n}_terms"] = rates, dists, terms
            if len(rates):
                for tag, k in (("full", len(rates)), ("half", max(1, len(rates) // 2))):
                    sl, sp = segments_from_passes(rates, terms, k)
                    out[f"blk{n}_dec_{tag}"] = ref_t1_decode_segs(data[:int(sl.sum())], sl, sp, numbps, orient, sty, w, h)
            n += 1
    out["count"] = np.array([n])
    np.savez_compressed(os.path.join(HERE, "t1_style_blocks.npz"), **out)


def transform_vectors():
    rng = np.random.default_rng(7)
    out = {}
    geoms = [(0, 0, 64, 48, 4), (3, 5, 40, 41, 6), (1, 1, 2, 2, 3), (7, 0, 8, 33, 5), (5, 3, 133, 97, 6), (0, 0, 1, 1, 2)]
    R = ref()
    for i, (x0, y0, x1, y1, nr) in enumerate(geoms):
        for rev in (1, 0):
            d = rng.integers(-2 ** 11, 2 ** 11, (y1 - y0, x1 - x0)).astype(np.int32) * (1 if rev else 2048)
            f = aligned(d)
            R.ref_dwt_encode(f.ravel(), x0, y0, x1, y1, nr, rev)
            out[f"dwt{i}_{rev}_geom"] = np.array([x0, y0, x1, y1, nr])
            out[f"dwt{i}_{rev}_in"] = d
            out[f"dwt{i}_{rev}_fwd"] = np.array(f)
            for nd in (nr, max(1, nr - 2)):
                top = nr - nd
                cd = lambda v: (v + (1 << top) - 1) >> top
                ww, hh = cd(x1) - cd(x0), cd(y1) - cd(y0)
                c = rng.integers(-2 ** 10, 2 ** 10, (hh, ww)).astype(np.int32) if rev else \
                    (rng.standard_normal((hh, ww)) * 50).astype(np.float32).view(np.int32)
                g = aligned(c)
                R.ref_dwt_decode(g.ravel(), x0, y0, x1, y1, nr, nd, rev)
                out[f"dwt{i}_{rev}_inv{nd}_in"] = c
                out[f"dwt{i}_{rev}_inv{nd}_out"] = np.array(g)
    n = 4099
    a = [rng.integers(-2 ** 19, 2 ** 19, n).astype(np.int32) for _ in range(3)]
    for name, fn in (("rct_fwd", R.ref_mct_encode_rev), ("rct_inv", R.ref_mct_decode_rev), ("ict_fwd", R.ref_mct_encode_irrev)):
        b = [aligned(x) for x in a]
        fn(*b, n)
        out[name] = np.stack([np.array(x) for x in b])
    f = [(rng.standard_normal(n) * 200).astype(np.float32) for _ in range(3)]
    b = [aligned(x) for x in f]
    R.ref_mct_decode_irrev(*b, n)
    out["mct_in"] = np.stack(a)
    out["ict_inv_in"] = np.stack(f)
    out["ict_inv"] = np.stack([np.array(x) for x in b])
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **out)


def codestream_vectors():
    """small whole-codec cases: the reference's codestream and decoded pixels for seeded synthetic images"""
    out = {}
    cases = [("gray53", 1, 8, True, (0, 0), ()), ("rgb53_tiled", 3, 8, True, (64, 64), ()), ("rgb97_layers", 3, 8, False, (64, 48), (20, 8, 3))]
    for name, nc, prec, rev, tile, rates in cases:
        img = synthetic_planes(160, 112, nc, prec, seed=len(name))
        cs = ref_encode_image(img, prec, tile=tile, numres=4, cblk=(32, 32), irreversible=not rev, rates=rates, rc_algorithm=1)
        dec = ref_decode_image(cs, nc, 160, 112)
        out[name + "_cs"] = np.frombuffer(cs, np.uint8)
        out[name + "_dec"] = np.stack(dec)
    np.savez_compressed(os.path.join(HERE, "codestreams.npz"), **out)


def rd_vectors():
    """RateControl::convexHull on pass tables: random ones (zero-length passes, flat and negative distortion steps, equal slopes)
    and the real tables of the reference's Tier-1 on synthetic blocks"""
    rng = np.random.default_rng(4242)
    R = ref()
    tables = random_pass_tables(rng, 120)
    z = np.load(os.path.join(HERE, "t1_blocks.npz"))
    for i in range(int(z["count"][0])):
        rates = z[f"blk{i}_rates"].astype(np.int64)
        if len(rates):
            tables.append((np.diff(np.concatenate([[0], rates])).astype(np.uint32), z[f"blk{i}_dists"].astype(np.float64)))
    out = {"count": np.array([len(tables)])}
    for i, (lens, dist) in enumerate(tables):
        slopes = np.zeros(len(lens), np.uint16)
        R.ref_rd_convex_hull(np.ascontiguousarray(lens), np.ascontiguousarray(dist), len(lens), slopes)
        out[f"t{i}_len"], out[f"t{i}_dist"], out[f"t{i}_slope"] = lens, dist, slopes
    np.savez_compressed(os.path.join(HERE, "rd_slopes.npz"), **out)


if __name__ == "__main__":
    rd_vectors()
    t1_vectors()
    t1_style_vectors()
    transform_vectors()
    codestream_vectors()
    print("golden fixtures written to", HERE)
