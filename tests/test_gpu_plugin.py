"""GPU: drop-in parity through Grok's official plugin ABI (SURVEY.md 8(b) "B1").  The unmodified reference is asked
to encode a PNM file the way `grk_compress -g <dir>` does: grk_plugin_load finds integration/_build/libgrok_plugin.so
(integration/grok_plugin_b200.cpp), plugin_encode runs DC shift + MCT + DWT + quantisation + Tier-1 on the B200 and hands
the host a grk_plugin_tile; the host's own PCRD, Tier-2 and codestream writer finish the job.  The codestream must be
byte-identical to a pure CPU run of the reference on the same pixels."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "integration", "_build")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libgrok_plugin.so")), reason="integration/_build not built")]

CASES = {
    # name: (width, height, comps, prec, reversible, numres, cblk, rates)
    "gray53": (200, 150, 1, 8, True, 5, (32, 32), ()),
    "rgb53": (260, 200, 3, 8, True, 6, (64, 64), ()),
    "rgb16_53": (130, 90, 3, 16, True, 3, (64, 64), ()),
    "rgb97_layers": (256, 200, 3, 8, False, 6, (64, 64), (20, 8, 3)),
    "gray12_97": (173, 131, 1, 12, False, 6, (32, 16), (12, 4)),
    "c1_full": (2048, 2048, 1, 8, True, 6, (64, 64), ()),
}

RUNNER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.dirname(sys.argv[1]))
import numpy as np, _libs
from grokimagecompression_b200.synth import synthetic_planes
from test_gpu_plugin import CASES
out = {}
for name in sys.argv[4:]:
    w, h, nc, prec, rev, numres, cblk, rates = CASES[name]
    img = synthetic_planes(w, h, nc, prec, seed=len(name) + w)
    if sys.argv[2] == "plugin":
        path = os.path.join(os.path.dirname(sys.argv[3]), name + (".pgm" if nc == 1 else ".ppm"))
        _libs.write_pnm(path, img, prec)
        cs = _libs.ref_plugin_encode_file(path, w * h * nc, numres=numres, cblk=cblk, irreversible=not rev, rates=rates)
        assert not isinstance(cs, int), f"{name}: plugin path status {cs}"
    else:
        cs = _libs.ref_encode_image(img, prec, numres=numres, cblk=cblk, irreversible=not rev, rates=rates, rc_algorithm=1)
    out[name] = np.frombuffer(cs, np.uint8)
    if sys.argv[2] == "plugin":
        # decode through the plugin ABI as well: host T2 -> device T1 / IDWT / MCT -> host stores the image
        for key, kw in (("_dec", {}), ("_dec_r1", {"reduce": 1}), ("_dec_l1", {"layers": 1})):
            dec = _libs.ref_plugin_decode(cs, nc, w, h, **kw)
            assert not isinstance(dec, int), f"{name}{key}: plugin decode status {dec}"
            out[name + key] = np.stack(dec)
    else:
        out[name + "_dec"] = np.stack(_libs.ref_decode_image(cs, nc, w, h))
        out[name + "_dec_r1"] = np.stack(_libs.ref_decode_image(cs, nc, w, h, reduce=1))
        out[name + "_dec_l1"] = np.stack(_libs.ref_decode_image(cs, nc, w, h, layers=1))
    out[name + "_img"] = np.stack(img)
np.savez_compressed(sys.argv[3], **out)
"""


def _run(mode, out, cases):
    subprocess.check_call([sys.executable, "-c", RUNNER, HERE, mode, out] + cases, timeout=900)
    return np.load(out)


@pytest.mark.parametrize("cases", [["gray53", "rgb53", "rgb16_53"], ["rgb97_layers", "gray12_97"], ["c1_full"]])
def test_plugin_encode_and_decode_match_the_reference(tmp_path, cases):
    pure = _run("pure", str(tmp_path / "pure.npz"), cases)
    plug = _run("plugin", str(tmp_path / "plugin.npz"), cases)
    for name in cases:
        assert pure[name].tobytes() == plug[name].tobytes(), f"{name}: codestream differs"
        for key in ("_dec", "_dec_r1", "_dec_l1"):
            assert (pure[name + key] == plug[name + key]).all(), f"{name}{key}: decoded pixels differ"
        if CASES[name][4]:
            assert (plug[name + "_dec"] == plug[name + "_img"]).all(), f"{name}: not lossless"


def test_multi_tile_request_is_declined(tmp_path):
    """one grk_plugin_tile describes the whole image (j2k.cpp:2069): the adapter must say no, not mis-encode"""
    code = RUNNER.replace('cs = _libs.ref_plugin_encode_file(path, w * h * nc, numres=numres',
                          'cs = _libs.ref_plugin_encode_file(path, w * h * nc, tile=(64, 64), numres=numres')
    code = code.replace('assert not isinstance(cs, int), f"{name}: plugin path status {cs}"', 'assert cs == -2, cs; sys.exit(0)')
    subprocess.check_call([sys.executable, "-c", code, HERE, "plugin", str(tmp_path / "x.npz"), "gray53"], timeout=300)


BATCH_RUNNER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.dirname(sys.argv[1]))
import numpy as np, _libs
from grokimagecompression_b200.synth import synthetic_planes
d = sys.argv[2]
w, h, nc, prec, n = 320, 240, 3, 12, 7
frames = [synthetic_planes(w, h, nc, prec, seed=1000 + f) for f in range(n)]
for f, img in enumerate(frames):
    _libs.write_pnm(os.path.join(d, "frame%03d.ppm" % f), img, prec)
open(os.path.join(d, "notes.txt"), "w").write("not an image")
got = _libs.ref_plugin_batch_encode(d, n + 2, w * h * nc, numres=5, cblk=(32, 32), irreversible=True, rates=(12, 4))
assert not isinstance(got, int), got
assert len(got) == n, len(got)
for f, img in enumerate(frames):
    want = _libs.ref_encode_image(img, prec, numres=5, cblk=(32, 32), irreversible=True, rates=(12, 4), rc_algorithm=1)
    assert got[f] == want, f
print("batch ok", n)
"""


def test_plugin_batch_encode_owns_the_frame_loop(tmp_path):
    """plugin_batch_encode / plugin_is_batch_complete / plugin_stop_batch_encode: every PNM frame of a directory goes through
    the device while the host's callback finishes the previous one; each codestream equals the pure reference's"""
    out = subprocess.check_output([sys.executable, "-c", BATCH_RUNNER, HERE, str(tmp_path)], timeout=600, text=True)
    assert "batch ok 7" in out


BATCH_DEC_RUNNER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.dirname(sys.argv[1]))
import numpy as np, _libs
from grokimagecompression_b200.synth import synthetic_planes
os.environ["GRK_REF_TRACE"] = "1"  # step markers on stderr (shown by pytest only when the runner fails)
d = sys.argv[2]
specs = [(320, 240, 3, 8, True, ()), (200, 150, 1, 12, False, (10,)), (256, 256, 3, 8, False, (20, 5)), (130, 90, 3, 16, True, ()),
         (320, 240, 3, 8, True, ())]
streams = []
for f, (w, h, nc, prec, rev, rates) in enumerate(specs):
    img = synthetic_planes(w, h, nc, prec, seed=2000 + f)
    cs = _libs.ref_encode_image(img, prec, numres=5, cblk=(32, 32), irreversible=not rev, rates=rates, rc_algorithm=1)
    open(os.path.join(d, "frame%03d.j2k" % f), "wb").write(cs)
    streams.append(cs)
open(os.path.join(d, "notes.txt"), "w").write("not a codestream")
for reduce in (0, 1):
    got = _libs.ref_plugin_batch_decode(d, len(specs) + 2, 320 * 256, reduce=reduce)
    assert not isinstance(got, int), got
    assert len(got) == len(specs), len(got)
    for f, (w, h, nc, prec, rev, rates) in enumerate(specs):
        want = np.stack(_libs.ref_decode_image(streams[f], nc, w, h, reduce=reduce))
        assert got[f].shape == want.shape and (got[f] == want).all(), (f, reduce)
print("batch decode ok", len(specs))
"""


def test_plugin_batch_decode_walks_the_directory(tmp_path):
    """plugin_init_batch_decode / plugin_batch_decode / plugin_stop_batch_decode: every codestream of a directory (mixed
    geometry, precision and wavelet) is decoded on the device and handed to the host's callback; pixels equal the pure reference's"""
    out = subprocess.check_output([sys.executable, "-X", "faulthandler", "-c", BATCH_DEC_RUNNER, HERE, str(tmp_path)], timeout=600, text=True)
    assert "batch decode ok 5" in out


ALL_DEVICES_TAIL = r"""
import ctypes, torch
P = ctypes.CDLL(os.path.join(os.path.dirname(sys.argv[1]), "integration", "_build", "libgrok_plugin.so"))
P.grok_b200_plugin_stat.restype = ctypes.c_uint64
P.grok_b200_plugin_stat.argtypes = [ctypes.c_int]
per_device = [int(P.grok_b200_plugin_stat(4 + k)) for k in range(16)]
ndev = torch.cuda.device_count()
print("per device", per_device[:ndev], "devices", ndev)
assert sum(per_device) >= WANT_FRAMES, per_device
assert sum(1 for v in per_device if v) == min(ndev, WANT_FRAMES), per_device   # every GPU of the box got frames
assert max(per_device) - min(per_device[:ndev]) <= 1 * REPEATS                    # dealt round robin
"""


def test_plugin_batch_on_all_devices(tmp_path):
    """grk_compress / grk_decompress -G -1 ("all devices", grk_compress.cpp:423-426): plugin_init takes every GPU of the box, the
    batch entry points deal frame i to device i mod N, the host still sees the frames in file-name order and every codestream /
    image equals the pure reference's.  On a one-GPU box this runs the same code with N = 1."""
    env = dict(os.environ, GROK_B200_DEVICE="-1")
    (tmp_path / "enc").mkdir()
    (tmp_path / "dec").mkdir()
    code = BATCH_RUNNER + ALL_DEVICES_TAIL.replace("WANT_FRAMES", "7").replace("REPEATS", "1")
    out = subprocess.check_output([sys.executable, "-c", code, HERE, str(tmp_path / "enc")], timeout=600, text=True, env=env)
    assert "batch ok 7" in out
    code = BATCH_DEC_RUNNER + ALL_DEVICES_TAIL.replace("WANT_FRAMES", "10").replace("REPEATS", "2")
    out = subprocess.check_output([sys.executable, "-X", "faulthandler", "-c", code, HERE, str(tmp_path / "dec")], timeout=600, text=True, env=env)
    assert "batch decode ok 5" in out
