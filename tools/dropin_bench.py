#!/usr/bin/env python
"""What an unmodified Grok gains from the TCD stage seam: the reference codec (oracle/_ref/libgrok_ref.so, public grk_* API,
memory streams, all host threads) encodes and decodes a workload image pure, and again with integration/grok_tcd_shim.cpp
loaded in front of it, so that DC shift / MCT / DWT / quantisation / Tier-1 run on the B200 while Tier-2, PCRD and the
codestream stay in Grok.  Prints wall-clock Mpixel/s of both and the host-side remainder.

    python tools/dropin_bench.py [c1|c2|c3|c4] [reps]          (spawns itself once per mode)
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(mode, name, reps):
    shim = None
    if mode == "shim":
        shim = C.CDLL(os.path.join(ROOT, "integration", "_build", "libgrok_b200_tcd.so"), mode=C.RTLD_GLOBAL)
    import _libs
    import bench
    from grokimagecompression_b200.synth import synthetic_planes
    w = bench.WORKLOADS[name]
    img = synthetic_planes(w["width"], w["height"], w["comps"], w["prec"], seed=1000)
    cb = (1 << w["cblk"][0], 1 << w["cblk"][1])
    kw = dict(tile=(w["tile"][0] or 0, w["tile"][1] or 0), numres=w["numres"], cblk=cb, irreversible=not w["reversible"],
              rates=w["rates"], rc_algorithm=1)
    if w.get("cinema"):
        kw["cinema2k_fps"] = w["cinema"]
    if w.get("ht"):
        kw["cblk_sty"] = 64
    te, td = [], []
    cs = None
    for i in range(reps + 1):
        t0 = time.perf_counter()
        cs = _libs.ref_encode_image(img, w["prec"], **kw)
        t1 = time.perf_counter()
        _libs.ref_decode_image(cs, w["comps"], w["width"], w["height"])
        t2 = time.perf_counter()
        if i:  # the first pass warms thread pool, CUDA context and tables
            te.append(t1 - t0)
            td.append(t2 - t1)
        elif shim is not None:
            shim.grok_b200_shim_reset_seconds()
    import hashlib
    secs = None
    if shim is not None:
        shim.grok_b200_shim_seconds.restype = C.c_double
        n = reps
        secs = [shim.grok_b200_shim_seconds(i) / n for i in range(8)]
    print(json.dumps(dict(mode=mode, secs=secs, enc_s=min(te), dec_s=min(td), bytes=len(cs), sha=hashlib.sha1(cs).hexdigest(),
                          pixels=w["width"] * w["height"], cores=os.cpu_count())))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("pure", "shim"):
        child(sys.argv[1], sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    as_json = "--json" in sys.argv
    argv = [a for a in sys.argv if a != "--json"]
    name = argv[1] if len(argv) > 1 else "c2"
    reps = int(argv[2]) if len(argv) > 2 else 3
    res = {}
    for mode in ("pure", "shim"):
        out = subprocess.check_output([sys.executable, os.path.abspath(__file__), mode, name, str(reps)], text=True)
        res[mode] = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    p, s = res["pure"], res["shim"]
    mp = p["pixels"] / 1e6
    if as_json:  # one line for bench.py's `drop_in` object
        t = s.get("secs") or [0.0] * 8
        print(json.dumps({"what": "unmodified reference codec, public grk API, memory streams: pure CPU vs TCD stage seam bound to libgrok_b200 (wall clock)",
                          "workload": name, "host_cores": p["cores"], "codestreams_identical": p["sha"] == s["sha"],
                          "encode_x": round(p["enc_s"] / s["enc_s"], 2), "decode_x": round(p["dec_s"] / s["dec_s"], 2),
                          "pure_encode_ms": round(p["enc_s"] * 1e3, 1), "seam_encode_ms": round(s["enc_s"] * 1e3, 1),
                          "pure_decode_ms": round(p["dec_s"] * 1e3, 1), "seam_decode_ms": round(s["dec_s"] * 1e3, 1),
                          "inside_seam_encode_ms": round(1e3 * (t[0] + t[3]), 1), "inside_seam_decode_ms": round(1e3 * (t[4] + t[6]), 1)}))
        sys.exit(0)
    print(f"{name}: codestreams identical: {p['sha'] == s['sha']} ({p['bytes']} bytes), host cores {p['cores']}")
    print(f"  encode  pure {mp / p['enc_s']:8.1f} Mpixel/s ({p['enc_s'] * 1e3:7.1f} ms)   with the seam on the B200 {mp / s['enc_s']:8.1f} Mpixel/s ({s['enc_s'] * 1e3:7.1f} ms)   x{p['enc_s'] / s['enc_s']:.1f}")
    if s.get("secs"):
        t = s["secs"]
        print(f"  inside the seam, mean per image: encode {1e3 * (t[0] + t[3]):.1f} ms (device path {1e3 * t[0]:.1f} + hand-over {1e3 * t[3]:.1f}), "
              f"decode {1e3 * (t[4] + t[6]):.1f} ms (collect {1e3 * t[4]:.1f} + device path {1e3 * t[6]:.1f}); the rest is Grok's own Tier-2 / PCRD / codestream")
    print(f"  decode  pure {mp / p['dec_s']:8.1f} Mpixel/s ({p['dec_s'] * 1e3:7.1f} ms)   with the seam on the B200 {mp / s['dec_s']:8.1f} Mpixel/s ({s['dec_s'] * 1e3:7.1f} ms)   x{p['dec_s'] / s['dec_s']:.1f}")
