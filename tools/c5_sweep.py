"""configs[4]-style decode sweep: a tiled 3-component image is encoded once (5/3 lossless and 9/7), then decoded at
full and reduced resolutions; prints device-resident decode throughput per reduction (Mpixel/s of the FULL image)
and checks the properties that do not need an oracle at this size: 5/3 full-resolution decode is lossless, and
every reduced decode equals the LL band the encoder produced at that level after inverse MCT / level shift of a
re-encode-free reference (done at small size in tests/test_gpu_drop_in.py::sweep*).
usage: python tools/c5_sweep.py [size]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grokimagecompression_b200 as gb
from grokimagecompression_b200 import params as P
from grokimagecompression_b200.synth import synthetic_planes

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
tile = (1024, 1024)
ctx = gb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
img = synthetic_planes(size, size, 3, 8, seed=5)
planes = P.split_planes(img, size, size, tile)
for rev in (True, False):
    te = P.image_tiles(size, size, 3, 8, rev, tile, 6, rate_control=not rev)
    ep = gb.Plan(ctx, te, True)
    res, rates, dists, data = ep.encode(planes)
    inp = np.zeros(len(res), gb.CBLK_DEC_DTYPE)
    for k in ("numbps", "numpasses", "data_len", "data_offset"):
        inp[k] = res[k]
    print(f"{'5/3 lossless' if rev else '9/7 lossy'}: {size}x{size}x3, {ep.num_blocks} blocks, {len(data)/1e6:.1f} MB of code-block bytes")
    resno = ep.blocks["resno"].copy()
    ep.close()
    for red in (0, 1, 2, 3):
        nd = 6 - red
        td = P.image_tiles(size, size, 3, 8, rev, tile, 6, encoder=False, numres_decode=nd)
        dp = gb.Plan(ctx, td, False)
        mask = resno < nd  # blocks of the decoded resolutions, in the encoder's (= the host's) order
        dinp = inp[mask]
        dp.decode_upload(dinp, data)
        dp.decode_run(); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            dp.decode_run()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out = dp.decode_download()
        ok = ""
        if rev and red == 0:
            full = P.join_planes(out, size, size, 3, tile)
            ok = "lossless" if all((a == b).all() for a, b in zip(full, img)) else "MISMATCH"
        print(f"   reduce {red}: {ms:8.2f} ms  {size*size/ms/1e3:9.1f} Mpixel/s (full-image pixels)  out {out[0].shape} {ok}")
        dp.close()
