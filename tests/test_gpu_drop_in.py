"""GPU: drop-in parity.  The unmodified reference codec is run twice on the same seeded images -- pure CPU, and
with the TCD stage seam bound to libgrok_b200.so (integration/grok_tcd_shim.cpp).  Tier-2, PCRD and codestream
writing are the reference's own code in both runs, so equal codestreams mean the GPU path handed it identical
code-block bytes, pass counts, rates and distortions; equal decoded pixels mean the GPU decode path is exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libgrok_b200_tcd.so")), reason="oracle/_ref not built")]


def _run(mode, out, cases):
    subprocess.check_call([sys.executable, os.path.join(HERE, "drop_in_runner.py"), mode, out] + cases, timeout=900)
    return np.load(out)


@pytest.mark.parametrize("cases", [["gray53", "rgb53_tiled", "rgb97_layers", "rgb16_53"], ["c1_full"], ["c2_crop", "c4_frame"],
                                   ["random53", "constant53", "random97", "ragged53", "ragged97", "tiny", "onepixel"],
                                   ["sweep53", "sweep97"],
                                   ["lazy53", "termall97", "resetvsc53", "allmodes53", "lazyterm97", "segsympterm16"],
                                   ["roi53", "roi97"]])
def test_codestreams_and_pixels_identical(tmp_path, cases):
    pure = _run("pure", str(tmp_path / "pure.npz"), cases)
    shim = _run("shim", str(tmp_path / "shim.npz"), cases)
    calls = shim["calls"]
    assert calls[0] > 0 and calls[3] > 0 and calls[4] > 0 and calls[5] > 0, calls  # the seam really was taken
    # ... including the rate allocator's per-block RateControl::convexHull, answered with the slopes computed on the device
    # (a single lossless layer needs no slopes: the host does not ask, TileProcessor.cpp:407)
    if any("97" in n or n in ("c2_crop", "c4_frame") for n in cases):
        assert int(shim["hulls"][0]) > 0
    for name in cases:
        lossless = name in ("gray53", "rgb53_tiled", "rgb16_53", "c1_full", "random53", "constant53", "ragged53", "tiny", "onepixel", "sweep53",
                            "lazy53", "resetvsc53", "allmodes53", "segsympterm16")  # (with -ROI the reference itself is not lossless: its encoder
        # declares the shift without applying it; the seam must reproduce exactly that)
        assert pure[name + "_cs"].tobytes() == shim[name + "_cs"].tobytes(), f"{name}: codestream differs"
        assert (pure[name + "_dec"] == shim[name + "_dec"]).all(), f"{name}: decoded pixels differ"
        for key in [k for k in pure.files if k.startswith(name + "_dec_")]:
            assert (pure[key] == shim[key]).all(), f"{key}: pixels differ"
        if lossless:
            assert (shim[name + "_dec"] == shim[name + "_img"]).all(), f"{name}: not lossless"
