"""CPU: the oracle (oracle/gb_oracle.c) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  Runs anywhere gcc is present; needs no GPU and no /root/reference."""
import os

import numpy as np

from _libs import oracle, oracle_t1_encode, oracle_t1_decode

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_t1_blocks_against_reference_vectors():
    z = np.load(os.path.join(G, "t1_blocks.npz"))
    n = int(z["count"][0])
    coded = 0
    for i in range(n):
        q = z[f"blk{i}_q"]
        orient, numbps, npass = (int(v) for v in z[f"blk{i}_meta"])
        data, nb, rates, dists, nsym = oracle_t1_encode(q, orient, True, float(z[f"blk{i}_wbase"][0]))
        assert nb == numbps and len(rates) == npass, i
        assert (rates == z[f"blk{i}_rates"]).all(), i
        assert (dists == z[f"blk{i}_dists"]).all(), i
        assert data == z[f"blk{i}_data"].tobytes(), i
        if npass:
            coded += 1
            h, w = q.shape
            assert (oracle_t1_decode(data, npass, numbps, orient, w, h) == z[f"blk{i}_dec_full"]).all(), i
            k = max(1, npass // 2)
            assert (oracle_t1_decode(data[:int(rates[k - 1])], k, numbps, orient, w, h) == z[f"blk{i}_dec_half"]).all(), i
    assert coded >= 16


def test_t1_style_blocks_against_reference_vectors():
    """LAZY / RESET / TERMALL / VSC / PTERM / SEGSYM: the restatement against vectors the reference produced"""
    from _libs import oracle_t1_encode_sty, oracle_t1_decode_segs, segments_from_passes
    z = np.load(os.path.join(G, "t1_style_blocks.npz"))
    n = int(z["count"][0])
    assert n >= 24
    for i in range(n):
        q = z[f"blk{i}_q"]
        orient, numbps, npass, sty = (int(v) for v in z[f"blk{i}_meta"])
        data, nb, rates, dists, terms, _ = oracle_t1_encode_sty(q, orient, sty, True, float(z[f"blk{i}_wbase"][0]))
        assert nb == numbps and len(rates) == npass, (i, sty)
        assert (rates == z[f"blk{i}_rates"]).all() and (terms == z[f"blk{i}_terms"]).all(), (i, sty)
        assert (dists == z[f"blk{i}_dists"]).all(), (i, sty)
        assert data == z[f"blk{i}_data"].tobytes(), (i, sty)
        if npass:
            h, w = q.shape
            for tag, k in (("full", npass), ("half", max(1, npass // 2))):
                sl, sp = segments_from_passes(rates, terms, k)
                got = oracle_t1_decode_segs(data[:int(sl.sum())], sl, sp, numbps, orient, sty, w, h)
                assert (got == z[f"blk{i}_dec_{tag}"]).all(), (i, sty, tag)


def test_transforms_against_reference_vectors():
    z = np.load(os.path.join(G, "transforms.npz"))
    O = oracle()
    i = 0
    while f"dwt{i}_1_geom" in z:
        for rev in (1, 0):
            x0, y0, x1, y1, nr = (int(v) for v in z[f"dwt{i}_{rev}_geom"])
            a = z[f"dwt{i}_{rev}_in"].copy()
            O.gbo_dwt_fwd(a.ravel(), x0, y0, x1, y1, nr, rev)
            assert (a == z[f"dwt{i}_{rev}_fwd"]).all(), (i, rev)
            for nd in (nr, max(1, nr - 2)):
                c = z[f"dwt{i}_{rev}_inv{nd}_in"].copy()
                O.gbo_dwt_inv(c.ravel(), x0, y0, x1, y1, nr, nd, rev)
                assert (c == z[f"dwt{i}_{rev}_inv{nd}_out"]).all(), (i, rev, nd)
        i += 1
    assert i >= 6
    n = z["mct_in"].shape[1]
    for name, fn in (("rct_fwd", O.gbo_rct_fwd), ("rct_inv", O.gbo_rct_inv), ("ict_fwd", O.gbo_ict_fwd)):
        a = [np.ascontiguousarray(r) for r in z["mct_in"]]
        fn(*a, n)
        assert (np.stack(a) == z[name]).all(), name
    f = [np.ascontiguousarray(r) for r in z["ict_inv_in"]]
    O.gbo_ict_inv(*f, n)
    assert (np.stack(f).view(np.int32) == z["ict_inv"].view(np.int32)).all()


def test_bench_dwt_check_semantics():
    # the reference's only value-pinning DWT test (bench_dwt.cpp:138-279, -check): values (i % 511) - 256 on an
    # odd origin / odd size must survive inverse-then-forward 5/3 unchanged
    O = oracle()
    x0, y0, x1, y1, nr = 3, 1, 3 + 257, 1 + 131, 6
    i = np.arange((y1 - y0) * (x1 - x0))
    d = ((i % 511) - 256).astype(np.int32).reshape(y1 - y0, x1 - x0)
    b = d.copy()
    O.gbo_dwt_inv(b.ravel(), x0, y0, x1, y1, nr, nr, 1)
    assert not (b == d).all()
    O.gbo_dwt_fwd(b.ravel(), x0, y0, x1, y1, nr, 1)
    assert (b == d).all()


def test_degenerate_and_empty_blocks():
    # zero-pass blocks: max|q| < 64 codes nothing (t1.cpp:1202-1212)
    q = np.full((8, 8), 63, np.int32)
    data, nb, rates, dists, nsym = oracle_t1_encode(q, 0)
    assert (data, nb, len(rates), nsym) == (b"", 0, 0, 0)
    # a single coefficient
    q = np.array([[-(5 << 6)]], np.int32)
    data, nb, rates, dists, nsym = oracle_t1_encode(q, 3)
    assert nb == 3 and len(rates) == 7
    assert oracle_t1_decode(data, 7, 3, 3, 1, 1)[0, 0] == -(5 * 2 + 1)


def test_rd_convex_hull_against_reference_vectors():
    """feasible truncation points and 8.8 log slopes (RateControl.cpp:31-168) of 120 random and the real Tier-1 pass tables"""
    z = np.load(os.path.join(G, "rd_slopes.npz"))
    O = oracle()
    n = int(z["count"][0])
    feasible = 0
    for i in range(n):
        lens, dist = np.ascontiguousarray(z[f"t{i}_len"]), np.ascontiguousarray(z[f"t{i}_dist"])
        got = np.zeros(len(lens), np.uint16)
        O.gbo_rd_convex_hull(lens, dist, len(lens), got)
        assert (got == z[f"t{i}_slope"]).all(), i
        feasible += int((got != 0).sum())
    assert n >= 130 and feasible > 500
