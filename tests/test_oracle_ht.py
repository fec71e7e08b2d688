"""CPU: the HTJ2K cleanup-pass restatement (oracle/gb_oracle_ht.c) against the LIVE compiled reference (t1/t1_ht/coding,
through oracle/ref_driver.cpp) on seeded random code blocks: byte streams equal, decoded sign-magnitude samples equal."""
import numpy as np
import pytest

from _libs import have_ref, oracle, ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


def random_block(rng, w, h, planes, kind):
    """sign-magnitude samples as T1HT::preEncode leaves them: magnitude MSB aligned below bit 31"""
    if kind == "sparse":
        mag = (rng.random((h, w)) < 0.08) * rng.integers(1, 1 << planes, (h, w))
    elif kind == "dense":
        mag = rng.integers(0, 1 << planes, (h, w))
    elif kind == "smooth":
        y, x = np.mgrid[0:h, 0:w]
        mag = np.abs((np.sin(x / 5.0) * np.cos(y / 7.0) * (1 << planes) * 0.9 + rng.normal(0, 1.5, (h, w)))).astype(np.int64)
        mag = np.minimum(mag, (1 << planes) - 1)
    else:
        mag = np.zeros((h, w), np.int64)
    sign = rng.integers(0, 2, (h, w)).astype(np.int64)
    return mag.astype(np.int64), sign


@pytest.mark.parametrize("seed", range(6))
def test_ht_encode_and_decode_match_reference(seed):
    rng = np.random.default_rng(100 + seed)
    O, R = oracle(), ref()
    n = 0
    for it in range(60):
        w = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 13, 16, 31, 32, 33, 63, 64]))
        h = int(rng.choice([1, 2, 3, 4, 5, 8, 15, 16, 32, 47, 64]))
        planes = int(rng.integers(1, 17))
        missing = int(rng.integers(0, 31 - planes))
        kind = ["sparse", "dense", "smooth", "zero"][it % 4]
        mag, sign = random_block(rng, w, h, planes, kind)
        shift = 31 - (missing + planes)  # top coded plane at bit 30 - missing
        sm = ((sign << 31) | (mag << shift)).astype(np.uint32).view(np.int32)
        sm = np.ascontiguousarray(sm.reshape(-1))
        cap = w * h * 8 + 4096
        a, b = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        la = R.ref_ht_encode_block(sm, w, h, w, missing, a, cap)
        lb = O.gbo_ht_encode_block(sm, w, h, w, missing, b, cap)
        assert la > 0 and la == lb, (it, w, h, planes, missing, kind, la, lb)
        assert bytes(a[:la]) == bytes(b[:lb]), (it, w, h, planes, missing, kind)
        da, db = np.zeros(w * h, np.int32), np.zeros(w * h, np.int32)
        R.ref_ht_decode_block(a[:la].copy(), la, missing, w, h, w, da)
        assert O.gbo_ht_decode_block(b[:lb].copy(), lb, missing, w, h, w, db) == 0
        assert (da == db).all(), (it, w, h, planes, missing, kind)
        # the cleanup pass carries every magnitude bit from the top coded plane down: decoding gives the samples back with
        # the half-bit of the reconstruction point set below them
        p = 30 - missing
        want = np.where(mag >> (p - shift) if p >= shift else mag << (shift - p), 1, 0)
        assert ((db.view(np.uint32) & 0x7FFFFFFF != 0) == (want.reshape(-1) != 0)).all()
        n += 1
    assert n == 60
