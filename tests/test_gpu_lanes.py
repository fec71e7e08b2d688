"""GPU: the Tier-1 kernels pick how many code blocks share a warp from the size of the launch (t1_enc.cu / t1_dec.cu launch code).
Every instantiation must give the same bytes and the same samples: the measurement knobs GB200_T1_MQ_LANES / GB200_T1_DEC_LANES
force each one on the same image (the default choice is checked against the oracle elsewhere)."""
import os

import numpy as np
import pytest

import grokimagecompression_b200 as gb
from grokimagecompression_b200 import params as P
from grokimagecompression_b200.synth import synthetic_planes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rev", [True, False])
def test_every_lane_count_gives_the_same_result(ctx, rev):
    w, h, nc = 640, 512, 3
    img = synthetic_planes(w, h, nc, 8, seed=11)
    tiles = P.image_tiles(w, h, nc, 8, rev, (256, 256), 5, rate_control=not rev, cblk_expn=(5, 5))
    planes = P.split_planes(img, w, h, (256, 256))
    tiles_d = P.image_tiles(w, h, nc, 8, rev, (256, 256), 5, cblk_expn=(5, 5), encoder=False)
    try:
        ref = gb.Plan(ctx, tiles, encoder=True).encode(planes)
        assert len(ref[0]) > 1000
        inp = np.zeros(len(ref[0]), gb.CBLK_DEC_DTYPE)
        for k in ("numbps", "numpasses", "data_len", "data_offset"):
            inp[k] = ref[0][k]
        ref_dec = gb.Plan(ctx, tiles_d, encoder=False).decode(inp, ref[3])
        for lanes in (1, 4, 8, 16, 32):
            os.environ["GB200_T1_MQ_LANES"] = str(lanes)
            got = gb.Plan(ctx, tiles, encoder=True).encode(planes)
            assert (got[0] == ref[0]).all() and (got[1] == ref[1]).all() and (got[2] == ref[2]).all() and bytes(got[3]) == bytes(ref[3]), lanes
        for lanes in (1, 2, 4, 8):
            os.environ["GB200_T1_DEC_LANES"] = str(lanes)
            os.environ["GB200_T1_DEC_UNIFORM"] = "0"
            got = gb.Plan(ctx, tiles_d, encoder=False).decode(inp, ref[3])
            for a, b in zip(got, ref_dec):
                assert (a == b).all(), lanes
        os.environ["GB200_T1_DEC_UNIFORM"] = "1"  # the warp-uniform decoder (one warp per block)
        got = gb.Plan(ctx, tiles_d, encoder=False).decode(inp, ref[3])
        for a, b in zip(got, ref_dec):
            assert (a == b).all(), "uniform"
    finally:
        os.environ.pop("GB200_T1_MQ_LANES", None)
        os.environ.pop("GB200_T1_DEC_LANES", None)
        os.environ.pop("GB200_T1_DEC_UNIFORM", None)


def test_contexts_on_two_threads_with_different_launch_shapes():
    """Two host threads, each with its own context, decode images of different sizes at the same time: the decoder kernel is
    launched with different amounts of dynamic shared memory (85 KB and 24 KB), and neither launch may disturb the other (the
    opt-in limit is a property of the kernel function, shared by every context of the process)."""
    import threading
    jobs = []
    for side, seed in ((4096, 5), (2048, 6)):
        img = synthetic_planes(side, side, 1, 8, seed=seed)
        tiles = P.image_tiles(side, side, 1, 8, True, (None, None), 6)
        tiles_d = P.image_tiles(side, side, 1, 8, True, (None, None), 6, encoder=False)
        planes = P.split_planes(img, side, side, (None, None))
        c = gb.Context(0)
        res, rates, dists, data = gb.Plan(c, tiles, encoder=True).encode(planes)
        inp = np.zeros(len(res), gb.CBLK_DEC_DTYPE)
        for k in ("numbps", "numpasses", "data_len", "data_offset"):
            inp[k] = res[k]
        jobs.append((c, tiles_d, inp, data, planes))
    errors = []

    def run(job):
        c, tiles_d, inp, data, planes = job
        try:
            plan = gb.Plan(c, tiles_d, encoder=False)
            for _ in range(25):
                got = plan.decode(inp, data)
            for a, b in zip(got, planes):
                assert (a == b).all()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ths = [threading.Thread(target=run, args=(j,)) for j in jobs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
