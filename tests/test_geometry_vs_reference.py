"""CPU: the block table the product derives from the coding parameters (gb200_enumerate_blocks / gb200_precinct_grid, pure host
code of libgrok_b200.so) against the tile structure the UNMODIFIED reference builds (TileComponent::init), block by block,
through the tap on the reference's tile coder (oracle/ref_tap.cpp).  Covers what the GPU-less suite can say about SURVEY 8
a-0: default and explicit precincts (also smaller than the nominal code block), the DCI 2K cinema profile, ragged tiles."""
import os
import subprocess
import sys

import pytest

from _libs import ORACLE_DIR

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libgrkref_tap.so")), reason="oracle/_ref not built")

CHILD = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, _libs
T = _libs.tap()   # before the reference driver: its stage calls then resolve to the tap
from grokimagecompression_b200.synth import synthetic_planes
cases = [
    # w, h, comps, prec, reversible, tile, numres, cblk, kwargs
    (2048, 1080, 3, 12, False, (0, 0), 6, (32, 32), dict(cinema2k_fps=24)),
    (300, 217, 3, 8, True, (128, 96), 5, (32, 32), dict(precincts=[(64, 64)])),
    (512, 384, 3, 8, False, (256, 256), 6, (64, 64), dict(precincts=[(256, 256), (128, 128)], rates=(20, 5))),
    (211, 157, 1, 8, True, (0, 0), 4, (64, 64), dict(precincts=[(32, 32), (16, 16)])),
    (400, 300, 3, 8, False, (0, 0), 5, (32, 32), dict(precincts=[(128, 64), (64, 64), (32, 64)], progression=2, rates=(30, 8))),
    (333, 127, 1, 16, True, (97, 61), 6, (16, 64), dict(precincts=[(64, 128), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)])),
    (5, 3, 1, 8, True, (0, 0), 3, (64, 64), dict()),
    (1, 1, 3, 8, True, (0, 0), 2, (64, 64), dict()),
    (640, 480, 3, 8, True, (200, 200), 6, (64, 64), dict()),
    (1024, 768, 1, 8, False, (0, 0), 8, (4, 1024 // 4 // 4), dict(precincts=[(512, 512), (256, 256), (128, 128), (64, 64)], rates=(10,))),
]
total = 0
for w, h, nc, prec, rev, tile, numres, cblk, kw in cases:
    img = synthetic_planes(w, h, nc, prec, seed=w + h)
    before = [T.ref_tap_geometry(i) for i in range(3)]
    cs = _libs.ref_encode_image(img, prec, tile=tile, numres=numres, cblk=cblk, irreversible=not rev, rc_algorithm=1, **kw)
    after = [T.ref_tap_geometry(i) for i in range(3)]
    assert after[0] > before[0], ("the tap saw no tile", w, h)
    assert after[2] == before[2], ("geometry mismatch", w, h, kw, after)
    total += after[1] - before[1]
assert T.ref_tap_calls(3) > 0 and T.ref_tap_seconds(3) > 0   # t1_encode went through the tap
print("blocks compared:", total)
"""


def test_block_tables_equal_the_reference_tile_structure():
    out = subprocess.check_output([sys.executable, "-c", CHILD % (HERE, ROOT)], text=True, timeout=600)
    n = int(out.strip().split(":")[-1])
    assert n > 10000, out
