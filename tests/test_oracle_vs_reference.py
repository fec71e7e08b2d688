"""CPU: the oracle against the LIVE compiled reference (oracle/_ref) on seeded random inputs.
Skipped where oracle/_ref has not been built (it is built by __graft_entry__.build() when
/root/reference is present and travels to the GPU box with the snapshot)."""
import ctypes as C

import numpy as np
import pytest

from _libs import (aligned, have_ref, oracle, oracle_t1_decode, oracle_t1_encode, ref, ref_t1_decode, ref_t1_encode,
                   ref_encode_image, ref_decode_image, random_pass_tables)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


def test_mct_live():
    rng = np.random.default_rng(1)
    O, R = oracle(), ref()
    for n in (1, 7, 10007):
        for fo, fr in (("gbo_rct_fwd", "ref_mct_encode_rev"), ("gbo_rct_inv", "ref_mct_decode_rev"), ("gbo_ict_fwd", "ref_mct_encode_irrev")):
            a = [rng.integers(-2 ** 20, 2 ** 20, n).astype(np.int32) for _ in range(3)]
            b = [aligned(x) for x in a]
            getattr(O, fo)(*a, n)
            getattr(R, fr)(*b, n)
            assert all((x == y).all() for x, y in zip(a, b)), fo
        a = [(rng.standard_normal(n) * 300).astype(np.float32) for _ in range(3)]
        b = [aligned(x) for x in a]
        O.gbo_ict_inv(*a, n)
        R.ref_mct_decode_irrev(*b, n)
        assert all((x.view(np.int32) == y.view(np.int32)).all() for x, y in zip(a, b))


@pytest.mark.parametrize("rev", [1, 0])
def test_dwt_live(rev):
    rng = np.random.default_rng(2 + rev)
    O, R = oracle(), ref()
    geoms = [(0, 0, 64, 64, 6), (0, 0, 37, 53, 4), (3, 5, 40, 41, 6), (1, 1, 2, 2, 3), (7, 0, 8, 33, 5), (0, 0, 1, 1, 2),
             (5, 3, 300, 211, 6), (1, 0, 3, 1, 3), (0, 0, 2, 1, 2), (1, 1, 3, 2, 2), (0, 0, 1024, 112, 6)]
    for (x0, y0, x1, y1, nr) in geoms:
        d = rng.integers(-2 ** 15, 2 ** 15, (y1 - y0, x1 - x0)).astype(np.int32) * (1 if rev else 8)
        a, b = d.copy(), aligned(d)
        O.gbo_dwt_fwd(a.ravel(), x0, y0, x1, y1, nr, rev)
        R.ref_dwt_encode(b.ravel(), x0, y0, x1, y1, nr, rev)
        assert (a == b).all(), (x0, y0, x1, y1, nr)
        for nd in sorted({nr, max(1, nr - 2), 1}):
            top = nr - nd
            cd = lambda v: (v + (1 << top) - 1) >> top
            ww, hh = cd(x1) - cd(x0), cd(y1) - cd(y0)
            c = rng.integers(-2 ** 12, 2 ** 12, (hh, ww)).astype(np.int32) if rev else \
                (rng.standard_normal((hh, ww)) * 100).astype(np.float32).view(np.int32)
            p, q = c.copy(), aligned(c)
            O.gbo_dwt_inv(p.ravel(), x0, y0, x1, y1, nr, nd, rev)
            R.ref_dwt_decode(q.ravel(), x0, y0, x1, y1, nr, nd, rev)
            assert (p == q).all(), (x0, y0, x1, y1, nr, nd)


def test_t1_live_random_blocks():
    rng = np.random.default_rng(7)
    nsym = 0
    for it in range(120):
        w = int(rng.choice([64, 32, 17, 5, 1, 64, 64, 33]))
        h = int(rng.choice([64, 32, 13, 4, 1, 7, 64, 64]))
        kind = rng.choice(["lap", "uni", "sparse"])
        amp = float(rng.choice([0.4, 2, 20, 300, 5000, 60000]))
        v = rng.laplace(0, amp, (h, w)) if kind != "uni" else rng.integers(-int(amp) - 1, int(amp) + 2, (h, w))
        if kind == "sparse":
            v = v * (rng.random((h, w)) < 0.05)
        q = (np.rint(v).astype(np.int64) * 64 + rng.integers(0, 64, (h, w)) * (it % 2)).astype(np.int32)
        orient, do_rd = int(rng.integers(0, 4)), bool(it % 2)
        step, lvl, comp, qm = float(rng.choice([1.0, 0.03125, 0.0123, 2.0])), int(rng.integers(0, 5)), int(rng.integers(0, 3)), int(rng.integers(0, 2))
        norms = np.array([1.732, 1.805, 1.573]) if qm == 0 else np.array([1.732, .8292, .8292])
        rb, rnb, rr, rd = ref_t1_encode(q, orient, comp, lvl, qm, step, norms, do_rd)
        wbase = (norms[comp] * ref().ref_dwt_norm(lvl, orient, qm)) * step
        ob, onb, orr, od, ns = oracle_t1_encode(q, orient, do_rd, wbase)
        assert rb == ob and rnb == onb and (rr == orr).all() and (rd == od).all(), it
        nsym += ns
        if len(rb):
            for k in sorted({len(rr), max(1, len(rr) // 2), 1}):
                ln = int(rr[k - 1])
                assert (ref_t1_decode(rb[:ln], k, rnb, orient, w, h) == oracle_t1_decode(rb[:ln], k, rnb, orient, w, h)).all(), (it, k)
    assert nsym > 500000


def test_t1_code_block_styles_against_live_reference():
    """every code-block style switch (LAZY, RESET, TERMALL, VSC, PTERM, SEGSYM) and random combinations of them:
    bytes, bit planes, rates, termination flags and fp64 distortions of the restatement equal the reference's, and the
    segment-wise decode of full and truncated streams agrees"""
    from _libs import oracle_t1_encode_sty, ref_t1_encode_sty, segments_from_passes, oracle_t1_decode_segs, ref_t1_decode_segs
    rng = np.random.default_rng(77)
    stys = [1, 2, 4, 8, 16, 32, 1 | 4, 1 | 16, 4 | 16, 2 | 8 | 32, 63] + [int(v) for v in rng.integers(1, 64, 40)]
    for it, sty in enumerate(stys):
        w = int(rng.choice([64, 32, 17, 5, 64, 33]))
        h = int(rng.choice([64, 32, 13, 4, 7, 64]))
        amp = float(rng.choice([2, 20, 300, 5000, 60000]))
        v = rng.laplace(0, amp, (h, w))
        if it % 4 == 3:
            v = v * (rng.random((h, w)) < 0.1)
        q = (np.rint(v).astype(np.int64) * 64 + rng.integers(0, 64, (h, w)) * (it % 2)).astype(np.int32)
        orient, do_rd = int(rng.integers(0, 4)), bool(it % 2)
        step, lvl, comp, qm = float(rng.choice([1.0, 0.03125, 2.0])), int(rng.integers(0, 5)), int(rng.integers(0, 3)), int(rng.integers(0, 2))
        norms = np.array([1.732, 1.805, 1.573]) if qm == 0 else np.array([1.732, .8292, .8292])
        rb, rnb, rr, rd, rt = ref_t1_encode_sty(q, orient, sty, comp, lvl, qm, step, norms, do_rd)
        wbase = (norms[comp] * ref().ref_dwt_norm(lvl, orient, qm)) * step
        ob, onb, orr, od, ot, ns = oracle_t1_encode_sty(q, orient, sty, do_rd, wbase)
        assert rnb == onb and (rr == orr).all() and (rt == ot).all() and rb == ob and (rd == od).all(), (it, sty)
        if len(rr):
            for k in sorted({len(rr), max(1, len(rr) // 2), 1}):
                sl, sp = segments_from_passes(rr, rt, k)
                ln = int(sl.sum())
                a = ref_t1_decode_segs(rb[:ln], sl, sp, rnb, orient, sty, w, h)
                b = oracle_t1_decode_segs(rb[:ln], sl, sp, rnb, orient, sty, w, h)
                assert (a == b).all(), (it, sty, k)


def test_tables_and_quantiser_constants():
    O, R = oracle(), ref()
    # step size / numbps / inv_step formula (Quantizer.cpp:65-105) against the E.1.1 restatement used by params.py
    for (expn, mant, orient, qm, prec) in [(10, 0, 0, 1, 8), (11, 0, 3, 1, 8), (13, 1234, 1, 0, 8), (9, 2047, 2, 0, 12), (18, 0, 3, 1, 16)]:
        st, nb, inv = C.c_float(), C.c_uint32(), C.c_uint32()
        R.ref_band_stepsize(expn, mant, 1 if orient else 0, max(orient - 1, 0), orient, qm, 2, prec, 1.0, C.byref(st), C.byref(nb), C.byref(inv))
        gain = 0 if (qm == 0 or orient == 0) else (1 if orient < 3 else 2)
        want = np.float32((1.0 + mant / 2048.0) * 2.0 ** (prec + gain - expn))
        assert st.value == want and nb.value == expn + 2 - 1
        assert inv.value == int(8192.0 / float(want) + 0.5)


def test_whole_codec_golden_is_current():
    """the committed codestream fixtures are what the reference produces today"""
    import os
    from grokimagecompression_b200.synth import synthetic_planes
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "codestreams.npz"))
    img = synthetic_planes(160, 112, 1, 8, seed=len("gray53"))
    cs = ref_encode_image(img, 8, numres=4, cblk=(32, 32), rc_algorithm=1)
    assert cs == z["gray53_cs"].tobytes()
    assert (np.stack(ref_decode_image(cs, 1, 160, 112)) == z["gray53_dec"]).all()
    assert (z["gray53_dec"][0] == img[0]).all()


def test_rd_convex_hull_matches_the_reference():
    """RateControl::convexHull (t2/RateControl.cpp:31-118) incl. the 8.8 log-slope conversion (:159-168)"""
    rng = np.random.default_rng(77)
    O, R = oracle(), ref()
    for lens, dist in random_pass_tables(rng, 600):
        a = np.zeros(len(lens), np.uint16)
        b = np.zeros(len(lens), np.uint16)
        O.gbo_rd_convex_hull(lens, dist, len(lens), a)
        R.ref_rd_convex_hull(lens, dist, len(lens), b)
        assert (a == b).all(), (lens, dist, a, b)
