"""GPU: the narrow-sample boundary (gb200_plan_set_sample_bytes, gb200_encode_tiles_packed / gb200_decode_tiles_packed).
Packed uint8 / uint16 / int8 / int16 planes must give exactly what the int32 planes of the reference's tile-buffer contract
give: the same code-block bytes, pass tables and distortions on the way in, the same pixels on the way out
(TileProcessor.cpp:1201-1258 copy-in, 1691-1921 copy-out: the reference widens / narrows on the host)."""
import numpy as np
import pytest

import grokimagecompression_b200 as gb
from grokimagecompression_b200 import params as P
from grokimagecompression_b200.synth import synthetic_planes

pytestmark = pytest.mark.gpu

CASES = [
    # width, height, comps, prec, sgnd, reversible, tile, numres, sample_bytes
    (256, 200, 3, 8, 0, False, (128, 112), 6, 1),     # configs[1] style: 8-bit RGB, 9/7 + ICT, tiled (ragged last tiles)
    (300, 217, 3, 8, 0, True, (128, 128), 4, 1),      # 8-bit RGB, 5/3 + RCT
    (190, 133, 3, 16, 0, True, (96, 64), 3, 2),       # configs[2] style: 16-bit, 5/3 + RCT
    (130, 70, 3, 12, 0, False, (None, None), 6, 2),   # configs[3] style: 12-bit in uint16, 9/7 + ICT
    (201, 150, 1, 8, 0, True, (None, None), 6, 1),    # gray, no MCT
    (77, 45, 1, 8, 1, True, (None, None), 3, 1),      # signed 8-bit: no level shift, int8 planes
    (99, 64, 3, 10, 1, False, (64, 64), 4, 2),        # signed 10-bit in int16, 9/7 + ICT
    (64, 64, 2, 8, 0, True, (None, None), 3, 2),      # two components (no MCT), 8-bit carried in uint16
    (37, 3, 1, 8, 0, False, (None, None), 2, 1),      # fewer samples than one vector
]


@pytest.mark.parametrize("case", CASES)
def test_packed_planes_equal_int32_planes(ctx, case):
    width, height, nc, prec, sgnd, rev, tile, numres, sb = case
    img = synthetic_planes(width, height, nc, prec, seed=width + 3 * height)
    if sgnd:
        img = [p - (1 << (prec - 1)) for p in img]
    rc = not rev
    tiles = P.image_tiles(width, height, nc, prec, rev, tile, numres, rate_control=rc, sgnd=sgnd)
    planes = P.split_planes(img, width, height, tile)
    wide = gb.Plan(ctx, tiles, encoder=True)
    narrow = gb.Plan(ctx, tiles, encoder=True, sample_bytes=sb)
    a = wide.encode(planes)
    packed = [np.ascontiguousarray(p.astype(narrow.sample_dtype(i))) for i, p in enumerate(planes)]
    for p, q in zip(planes, packed):
        assert (p == q).all()  # the narrow type holds every sample
    b = narrow.encode(packed)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all() and bytes(a[3]) == bytes(b[3])
    # a second run on the same plan (the packed planes are staged in a ping-pong buffer the transform overwrites)
    b2 = narrow.encode(packed)
    assert (b2[0] == b[0]).all() and bytes(b2[3]) == bytes(b[3])
    # the int32 entry point refuses a packed plan instead of misreading the buffers
    with pytest.raises(gb.GrokB200Error):
        gb.binding.check(gb.lib().gb200_encode_upload(narrow._h, narrow._ptr_array(planes)))
    inp = np.zeros(len(a[0]), gb.CBLK_DEC_DTYPE)
    for k in ("numbps", "numpasses", "data_len", "data_offset"):
        inp[k] = a[0][k]
    for nd in (0, max(1, numres - 1)):
        tiles_d = P.image_tiles(width, height, nc, prec, rev, tile, numres, sgnd=sgnd, encoder=False, numres_decode=nd)
        keep = np.array([(nd == 0) or (wide.blocks[i]["resno"] < nd) for i in range(wide.num_blocks)], bool)
        dw = gb.Plan(ctx, tiles_d, encoder=False)
        dn = gb.Plan(ctx, tiles_d, encoder=False, sample_bytes=sb)
        got_w = dw.decode(inp[keep], a[3])
        got_n = dn.decode(inp[keep], a[3])
        for i, (x, y) in enumerate(zip(got_w, got_n)):
            assert y.dtype == dn.sample_dtype(i) and x.shape == y.shape
            assert (x == y.astype(np.int32)).all()
        if rev and nd == 0:
            for x, y in zip(P.join_planes([g.astype(np.int32) for g in got_n], width, height, nc, tile), img):
                assert (x == y).all()


def test_packed_plan_rejects_precision_that_does_not_fit(ctx):
    tiles = P.image_tiles(64, 64, 1, 12, True, (None, None), 3)
    with pytest.raises(gb.GrokB200Error):
        gb.Plan(ctx, tiles, encoder=True, sample_bytes=1)


def test_two_contexts_share_the_tables(ctx):
    """a second context (own stream) created after the first has run Tier-1 must see initialised constant tables"""
    img = synthetic_planes(128, 128, 1, 8, seed=4)
    tiles = P.image_tiles(128, 128, 1, 8, True, (None, None), 4)
    planes = P.split_planes(img, 128, 128, (None, None))
    a = gb.Plan(ctx, tiles, encoder=True).encode(planes)
    c2 = gb.Context(0)
    b = gb.Plan(c2, tiles, encoder=True).encode(planes)
    assert bytes(a[3]) == bytes(b[3]) and (a[0] == b[0]).all()
    c2.close()
