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
REF = os.path.join(ROOT, "integration", "_build")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libgrok_b200_tcd.so")), reason="integration/_build not built")]


def _run(mode, out, cases):
    subprocess.check_call([sys.executable, os.path.join(HERE, "drop_in_runner.py"), mode, out] + cases, timeout=1800)
    return np.load(out)


@pytest.mark.parametrize("cases", [["gray53", "rgb53_tiled", "rgb97_layers", "rgb16_53"], ["c1_full"], ["c2_crop", "c4_frame"],
                                   ["random53", "constant53", "random97", "ragged53", "ragged97", "tiny", "onepixel"],
                                   ["sweep53", "sweep97"],
                                   ["lazy53", "termall97", "resetvsc53", "allmodes53", "lazyterm97", "segsympterm16"],
                                   ["roi53", "roi97"],
                                   # non-default precincts, the DCI 2K cinema profile (configs[3] as BASELINE states it)
                                   ["cinema2k", "prc53_64", "prc97_mixed", "prc53_clip", "prc97_rpcl"],
                                   # the HTJ2K block coder (-M 64)
                                   ["ht53", "ht53_gray16", "ht97", "ht97_12"],
                                   # BASELINE configs[1], [2] and [4] at their full size (compared by digest)
                                   ["c2_full"], ["c3_full"], ["c5_53"], ["c5_97"]])
def test_codestreams_and_pixels_identical(tmp_path, cases):
    pure = _run("pure", str(tmp_path / "pure.npz"), cases)
    shim = _run("shim", str(tmp_path / "shim.npz"), cases)
    calls = shim["calls"]
    assert calls[0] > 0 and calls[3] > 0 and calls[4] > 0 and calls[5] > 0, calls  # the seam really was taken
    # ... including the rate allocator's per-block RateControl::convexHull, answered with the slopes computed on the device
    # (a single lossless layer needs no slopes: the host does not ask, TileProcessor.cpp:407)
    if any(("97" in n and not n.startswith("ht")) or n in ("c2_crop", "c4_frame", "c2_full", "cinema2k") for n in cases):
        assert int(shim["hulls"][0]) > 0
    for name in cases:
        lossless = name in ("gray53", "rgb53_tiled", "rgb16_53", "c1_full", "random53", "constant53", "ragged53", "tiny", "onepixel", "sweep53",
                            "lazy53", "resetvsc53", "allmodes53", "segsympterm16", "prc53_64", "prc53_clip", "c3_full", "c5_53", "ht53", "ht53_gray16")  # (with -ROI the reference itself is not lossless: its encoder
        # declares the shift without applying it; the seam must reproduce exactly that)
        assert pure[name + "_cs"].tobytes() == shim[name + "_cs"].tobytes(), f"{name}: codestream differs"
        assert (pure[name + "_dec"] == shim[name + "_dec"]).all(), f"{name}: decoded pixels differ"
        for key in [k for k in pure.files if k.startswith(name + "_dec_")]:
            assert (pure[key] == shim[key]).all(), f"{key}: pixels differ"
        if lossless and name + "_lossless" in shim.files:
            assert bool(shim[name + "_lossless"][0]), f"{name}: not lossless"
        elif lossless:
            assert (shim[name + "_dec"] == shim[name + "_img"]).all(), f"{name}: not lossless"
