"""GPU parity: every stage kernel, through the C ABI, against the CPU oracle on the same inputs.
Integer paths must match bit for bit; the float paths (inverse 9/7, inverse ICT) as well, because the
kernels reproduce the reference's operation order without fused multiply-add."""
import numpy as np
import pytest

import grokimagecompression_b200 as gb
from _libs import oracle, oracle_t1_encode, oracle_t1_decode

pytestmark = pytest.mark.gpu


def test_mct_and_dc_shift(ctx):
    rng = np.random.default_rng(11)
    O = oracle()
    for n in (1, 3, 4, 1000, 65537, 1 << 20):
        a = [rng.integers(-2 ** 20, 2 ** 20, n).astype(np.int32) for _ in range(3)]
        for fo, fg in ((O.gbo_rct_fwd, ctx.mct_encode_rev), (O.gbo_rct_inv, ctx.mct_decode_rev),
                       (O.gbo_ict_fwd, ctx.mct_encode_irrev)):
            x = [v.copy() for v in a]
            y = [v.copy() for v in a]
            fo(*x, n)
            fg(*y)
            assert all((p == q).all() for p, q in zip(x, y))
        f = [(rng.standard_normal(n) * 300).astype(np.float32) for _ in range(3)]
        x = [v.copy() for v in f]
        y = [v.copy() for v in f]
        O.gbo_ict_inv(*x, n)
        ctx.mct_decode_irrev(*y)
        assert all((p.view(np.int32) == q.view(np.int32)).all() for p, q in zip(x, y))
        for rev in (1, 0):
            v = rng.integers(0, 4096, n).astype(np.int32)
            x, y = v.copy(), v.copy()
            O.gbo_dc_shift_fwd(x, n, 2048, rev)
            ctx.dc_shift_encode(y, 2048, rev)
            assert (x == y).all()
            if rev:
                w = rng.integers(-5000, 5000, n).astype(np.int32)
            else:  # halves exercise round-half-to-even
                w = (rng.integers(-10000, 10000, n) / 2.0).astype(np.float32).view(np.int32)
            x, y = w.copy(), w.copy()
            O.gbo_dc_shift_inv(x, n, 2048, rev, 0, 4095)
            ctx.dc_shift_decode(y, 2048, rev, 0, 4095)
            assert (x == y).all()


GEOMS = [(0, 0, 64, 64, 6), (0, 0, 37, 53, 4), (3, 5, 40, 41, 6), (1, 1, 2, 2, 3), (7, 0, 8, 33, 5), (0, 0, 1, 1, 2),
         (5, 3, 300, 211, 6), (1, 0, 3, 1, 3), (0, 0, 1024, 112, 6), (1024, 2048, 2048, 2160, 6), (0, 0, 129, 65, 3),
         (63, 63, 64 + 130, 64 + 67, 4), (0, 0, 5, 1, 4), (0, 0, 1, 7, 4), (9, 9, 10, 200, 3)]


@pytest.mark.parametrize("rev", [1, 0])
def test_dwt_forward(ctx, rev):
    rng = np.random.default_rng(5)
    O = oracle()
    for (x0, y0, x1, y1, nr) in GEOMS:
        d = rng.integers(-2 ** 15, 2 ** 15, (y1 - y0, x1 - x0)).astype(np.int32)
        if not rev:
            d = d * 8
        a, b = d.copy(), d.copy()
        O.gbo_dwt_fwd(a.ravel(), x0, y0, x1, y1, nr, rev)
        ctx.dwt_encode(b, x0, y0, x1, y1, nr, rev)
        assert (a == b).all(), (x0, y0, x1, y1, nr)


@pytest.mark.parametrize("rev", [1, 0])
def test_dwt_inverse_and_reduced(ctx, rev):
    rng = np.random.default_rng(6)
    O = oracle()
    for (x0, y0, x1, y1, nr) in GEOMS:
        for nd in sorted({nr, max(1, nr - 1), max(1, nr - 3), 1}):
            top = nr - nd
            cd = lambda v: (v + (1 << top) - 1) >> top
            ww, hh = cd(x1) - cd(x0), cd(y1) - cd(y0)
            if rev:
                c = rng.integers(-2 ** 12, 2 ** 12, (hh, ww)).astype(np.int32)
            else:
                c = (rng.standard_normal((hh, ww)) * 100).astype(np.float32).view(np.int32)
            a, b = c.copy(), c.copy()
            O.gbo_dwt_inv(a.ravel(), x0, y0, x1, y1, nr, nd, rev)
            ctx.dwt_decode(b, x0, y0, x1, y1, nr, nd, rev)
            assert (a == b).all(), (x0, y0, x1, y1, nr, nd)


def test_dwt53_perfect_reconstruction(ctx):
    # bench_dwt -check semantics (bench_dwt.cpp:138-279): odd origin and size, 5/3 inverse o forward = identity
    x0, y0, x1, y1, nr = 3, 5, 3 + 1021, 5 + 767, 6
    i = np.arange((y1 - y0) * (x1 - x0), dtype=np.int64)
    d = ((i % 511) - 256).astype(np.int32).reshape(y1 - y0, x1 - x0)
    b = d.copy()
    ctx.dwt_encode(b, x0, y0, x1, y1, nr, 1)
    assert not (b == d).all()
    ctx.dwt_decode(b, x0, y0, x1, y1, nr, nr, 1)
    assert (b == d).all()


def _random_blocks(rng, count):
    blocks = []
    for it in range(count):
        w = int(rng.choice([64, 64, 64, 32, 17, 5, 1, 33, 64]))
        h = int(rng.choice([64, 64, 32, 13, 4, 1, 7, 64, 3]))
        kind = rng.choice(["lap", "uni", "sparse", "zero"])
        amp = float(rng.choice([0.4, 2, 20, 300, 5000, 60000]))
        if kind == "lap":
            v = rng.laplace(0, amp, (h, w))
        elif kind == "uni":
            v = rng.integers(-int(amp) - 1, int(amp) + 2, (h, w))
        elif kind == "sparse":
            v = rng.laplace(0, amp, (h, w)) * (rng.random((h, w)) < 0.05)
        else:
            v = np.zeros((h, w))
        blocks.append(np.rint(v).astype(np.int32))
    return blocks


def _layout(blocks):
    """place blocks side by side in one plane"""
    H = max(b.shape[0] for b in blocks)
    W = sum(b.shape[1] for b in blocks)
    plane = np.zeros((H, W), np.int32)
    desc = np.zeros(len(blocks), gb.T1_BLOCK_DTYPE)
    x = 0
    for i, b in enumerate(blocks):
        h, w = b.shape
        plane[:h, x:x + w] = b
        desc[i]["x"], desc[i]["y"], desc[i]["w"], desc[i]["h"] = x, 0, w, h
        x += w
    return plane, desc


@pytest.mark.parametrize("rev,rd", [(1, False), (1, True), (0, True), (0, False)])
def test_t1_encode_blocks(ctx, rev, rd):
    rng = np.random.default_rng(100 + rev * 2 + rd)
    blocks = _random_blocks(rng, 96)
    plane, desc = _layout(blocks)
    for i in range(len(blocks)):
        desc[i]["orient"] = rng.integers(0, 4)
        desc[i]["qmfbid"] = rev
        desc[i]["inv_step"] = 8192 if rev else int(rng.choice([8192, 16384, 4096 * 3, 77777, 1000]))
        desc[i]["stepsize"] = 1.0
        desc[i]["rd_weight"] = float(rng.choice([1.0, 0.0123, 3.7]))
    res, rates, dists, data = ctx.t1_encode_blocks(plane, desc, rate_control=rd, max_passes=100)
    O = oracle()
    total_dec = 0
    for i, b in enumerate(blocks):
        h, w = b.shape
        q = np.zeros((h, w), np.int32)
        O.gbo_quantise_block(b.ctypes.data, w, w, h, rev, int(desc[i]["inv_step"]), q.ravel())
        ob, onb, orr, od, ns = oracle_t1_encode(q, int(desc[i]["orient"]), rd, float(desc[i]["rd_weight"]))
        r = res[i]
        assert r["numbps"] == onb, i
        assert r["numpasses"] == len(orr), i
        assert (rates[i, :len(orr)] == orr).all(), i
        assert r["data_len"] == len(ob), i
        got = bytes(data[int(r["data_offset"]):int(r["data_offset"]) + int(r["data_len"])])
        assert got == ob, i
        assert r["decisions"] == ns, i
        if rd:
            assert (dists[i, :len(orr)] == od).all(), i
        total_dec += ns
    assert total_dec > 100000


STYLES = [1, 2, 4, 8, 16, 32, 1 | 4, 1 | 16, 4 | 16, 1 | 4 | 16, 2 | 8 | 32, 63, 1 | 2 | 8, 1 | 32, 4 | 8]


@pytest.mark.parametrize("rd", [False, True])
def test_t1_encode_blocks_with_style_switches(ctx, rd):
    """LAZY / RESET / TERMALL / VSC / PTERM / SEGSYM and combinations: bytes, rates and distortions against the oracle"""
    from _libs import oracle_t1_encode_sty
    rng = np.random.default_rng(500 + rd)
    blocks = _random_blocks(rng, 120)
    plane, desc = _layout(blocks)
    for i in range(len(blocks)):
        desc[i]["orient"] = rng.integers(0, 4)
        desc[i]["qmfbid"] = 1
        desc[i]["inv_step"] = 8192
        desc[i]["stepsize"] = 1.0
        desc[i]["rd_weight"] = float(rng.choice([1.0, 0.0123, 3.7]))
        desc[i]["cblk_sty"] = STYLES[i % len(STYLES)] if i % 7 else int(rng.integers(1, 64))
    res, rates, dists, data = ctx.t1_encode_blocks(plane, desc, rate_control=rd, max_passes=100)
    for i, b in enumerate(blocks):
        sty = int(desc[i]["cblk_sty"])
        q = (b.astype(np.int64) * 64).astype(np.int32)
        ob, onb, orr, od, ot, ns = oracle_t1_encode_sty(q, int(desc[i]["orient"]), sty, rd, float(desc[i]["rd_weight"]))
        r = res[i]
        assert r["numbps"] == onb and r["numpasses"] == len(orr), (i, sty)
        assert (rates[i, :len(orr)] == orr).all(), (i, sty, rates[i, :len(orr)], orr)
        got = bytes(data[int(r["data_offset"]):int(r["data_offset"]) + int(r["data_len"])])
        assert got == ob, (i, sty)
        assert r["decisions"] == ns, (i, sty)
        if rd:
            assert (dists[i, :len(orr)] == od).all(), (i, sty)


@pytest.mark.parametrize("rev", [1, 0])
def test_t1_decode_blocks_with_style_switches(ctx, rev):
    """segment-wise decode of LAZY / RESET / TERMALL / VSC / PTERM / SEGSYM streams (full and truncated at a pass
    boundary) against the oracle"""
    from _libs import oracle_t1_encode_sty, oracle_t1_decode_segs, segments_from_passes
    rng = np.random.default_rng(700 + rev)
    blocks = _random_blocks(rng, 120)
    _, desc = _layout(blocks)
    O = oracle()
    inputs = np.zeros(len(blocks), gb.CBLK_DEC_DTYPE)
    chunks, expect, seg_start, segs = [], [], [0], []
    off = 0
    for i, b in enumerate(blocks):
        h, w = b.shape
        orient = int(rng.integers(0, 4))
        sty = STYLES[i % len(STYLES)] if i % 7 else int(rng.integers(1, 64))
        desc[i]["orient"], desc[i]["qmfbid"], desc[i]["cblk_sty"] = orient, rev, sty
        desc[i]["stepsize"] = 1.0 if rev else float(np.float32(rng.choice([0.5, 0.0371, 1.9])))
        ob, onb, orr, _, ot, _ = oracle_t1_encode_sty((b.astype(np.int64) * 64).astype(np.int32), orient, sty)
        npass = len(orr)
        # every fifth block belongs to a component with a max-shift ROI: the stream is decoded roishift planes higher and
        # samples at or above 2^roishift are shifted back (T1Part1.cpp:184-186, 230-252)
        roishift = int(rng.integers(1, 8)) if (i % 5 == 4 and onb + 8 < 30) else 0
        desc[i]["roishift"] = roishift
        k = npass if (i % 3 == 0 or npass == 0) else int(rng.integers(1, npass + 1))
        sl, sp = segments_from_passes(orr, ot, k) if k else (np.zeros(0, np.uint32), np.zeros(0, np.uint32))
        ln = int(sl.sum())
        inputs[i]["numbps"], inputs[i]["numpasses"], inputs[i]["data_len"], inputs[i]["data_offset"] = onb, k, ln, off
        chunks.append(ob[:ln])
        off += ln
        for a, c in zip(sl, sp):
            segs.append((int(a), int(c)))
        seg_start.append(len(segs))
        dec = oracle_t1_decode_segs(ob[:ln], sl, sp, onb, orient, sty, w, h, roishift) if k else np.zeros((h, w), np.int32)
        out = np.zeros((h, w), np.int32)
        O.gbo_dequantise_block(dec.ravel(), w, h, rev, float(desc[i]["stepsize"]), out.ctypes.data, w)
        expect.append(out)
    data = np.frombuffer(b"".join(chunks), np.uint8)
    H = max(b.shape[0] for b in blocks)
    W = sum(b.shape[1] for b in blocks)
    plane = ctx.t1_decode_blocks((H, W), desc, inputs, data, np.array(seg_start, np.uint32), np.array(segs, gb.CBLK_SEG_DTYPE))
    for i, e in enumerate(expect):
        h, w = e.shape
        x = int(desc[i]["x"])
        assert (plane[:h, x:x + w] == e).all(), (i, int(desc[i]["cblk_sty"]))


@pytest.mark.parametrize("rev", [1, 0])
def test_t1_decode_blocks(ctx, rev):
    rng = np.random.default_rng(300 + rev)
    blocks = _random_blocks(rng, 96)
    _, desc = _layout(blocks)
    O = oracle()
    inputs = np.zeros(len(blocks), gb.CBLK_DEC_DTYPE)
    chunks = []
    off = 0
    expect = []
    for i, b in enumerate(blocks):
        h, w = b.shape
        orient = int(rng.integers(0, 4))
        desc[i]["orient"] = orient
        desc[i]["qmfbid"] = rev
        desc[i]["stepsize"] = 1.0 if rev else float(np.float32(rng.choice([0.5, 0.0371, 1.9])))
        ob, onb, orr, _, _ = oracle_t1_encode((b * 64).astype(np.int32), orient)
        npass = len(orr)
        # decode all passes, or a truncation at a pass boundary (what a lower quality layer delivers)
        k = npass if (i % 3 == 0 or npass == 0) else int(rng.integers(1, npass + 1))
        ln = int(orr[k - 1]) if k else 0
        inputs[i]["numbps"], inputs[i]["numpasses"], inputs[i]["data_len"], inputs[i]["data_offset"] = onb, k, ln, off
        chunks.append(ob[:ln])
        off += ln
        dec = oracle_t1_decode(ob[:ln], k, onb, orient, w, h) if ln else np.zeros((h, w), np.int32)
        out = np.zeros((h, w), np.int32)
        O.gbo_dequantise_block(dec.ravel(), w, h, rev, float(desc[i]["stepsize"]), out.ctypes.data, w)
        expect.append(out)
    data = np.frombuffer(b"".join(chunks), np.uint8)
    H = max(b.shape[0] for b in blocks)
    W = sum(b.shape[1] for b in blocks)
    plane = ctx.t1_decode_blocks((H, W), desc, inputs, data)
    for i, e in enumerate(expect):
        h, w = e.shape
        x = int(desc[i]["x"])
        assert (plane[:h, x:x + w] == e).all(), i
