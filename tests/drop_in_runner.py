"""Runs the UNMODIFIED reference codec (oracle/_ref/libgrok_ref.so via ref_driver) on seeded synthetic images,
either pure (CPU) or with integration/grok_tcd_shim.cpp loaded in front of it, in which case Grok's TCD stage
calls land in libgrok_b200.so.  Writes codestreams and decoded pixels to an .npz for the caller to compare.

    python tests/drop_in_runner.py {pure|shim} out.npz [case ...]
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

CASES = {
    # name: (width, height, comps, prec, reversible, tile, numres, cblk, rates, reduce)
    "gray53": (160, 112, 1, 8, True, (0, 0), 4, (32, 32), (), 0),
    "rgb53_tiled": (200, 150, 3, 8, True, (64, 64), 4, (32, 32), (), 1),
    "rgb97_layers": (256, 200, 3, 8, False, (128, 112), 6, (64, 64), (20, 8, 3), 2),
    "rgb16_53": (130, 90, 3, 16, True, (0, 0), 3, (64, 64), (), 0),
    "c1_full": (2048, 2048, 1, 8, True, (0, 0), 6, (64, 64), (), 0),
    "c2_crop": (2048, 1080, 3, 8, False, (1024, 1024), 6, (64, 64), (40, 20, 10, 5), 0),
    "c4_frame": (2048, 1080, 3, 12, False, (0, 0), 6, (32, 32), (10,), 0),
    # stress variants: uniform random = worst case for Tier-1, constant = zero-pass blocks
    "random53": (300, 200, 3, 8, True, (128, 128), 5, (64, 64), (), 1, "random"),
    "constant53": (256, 256, 1, 8, True, (0, 0), 6, (64, 64), (), 0, "constant"),
    "random97": (256, 192, 3, 8, False, (0, 0), 5, (32, 32), (30, 5), 0, "random"),
    # ragged geometry: prime sizes, tiles that do not divide the image, tiny images
    "ragged53": (211, 157, 3, 8, True, (97, 61), 4, (16, 32), (), 1),
    "ragged97": (173, 131, 1, 12, False, (80, 80), 6, (32, 16), (12, 4), 2),
    "tiny": (5, 3, 1, 8, True, (0, 0), 3, (64, 64), (), 0),
    "onepixel": (1, 1, 3, 8, True, (0, 0), 2, (64, 64), (), 0),
    # configs[4]-style decode sweep: tiled 5/3 and 9/7 streams at every reduction, and a layer-limited decode
    "sweep53": (1536, 1280, 3, 8, True, (512, 512), 6, (64, 64), (), "sweep"),
    "sweep97": (1536, 1280, 3, 8, False, (512, 512), 6, (64, 64), (10,), "sweep"),
    # code-block style switches (grk_compress -M): LAZY 1, RESET 2, TERMALL 4, VSC 8, PTERM 16, SEGSYM 32
    "lazy53": (300, 200, 3, 8, True, (128, 128), 5, (64, 64), (), 1),
    "termall97": (256, 200, 3, 8, False, (128, 112), 6, (32, 32), (20, 8, 3), 0),
    "resetvsc53": (211, 157, 1, 12, True, (0, 0), 4, (16, 32), (), 0),
    "allmodes53": (256, 256, 3, 8, True, (0, 0), 6, (64, 64), (), 2),
    "lazyterm97": (640, 480, 3, 8, False, (0, 0), 6, (64, 64), (30, 10, 4), 1),
    "segsympterm16": (130, 90, 3, 16, True, (0, 0), 3, (64, 64), (), 0),
    # max-shift region of interest on one component (-ROI c=1,U=5 / c=0,U=3)
    "roi53": (200, 150, 3, 8, True, (64, 64), 4, (32, 32), (), 1),
    "roi97": (256, 192, 1, 8, False, (0, 0), 5, (64, 64), (20, 5), 0),
}
# non-default precincts / profiles (TileComponent.cpp:303-328, 437-489; the cinema profile forces 32x32 blocks,
# 256x256 precincts (128x128 at the lowest resolution), CPRL and a byte budget: j2kprofile.cpp:941-1080)
CASES.update({
    "cinema2k": (2048, 1080, 3, 12, False, (0, 0), 6, (32, 32), (), 0),
    "prc53_64": (300, 217, 3, 8, True, (128, 96), 5, (32, 32), (), 1),
    "prc97_mixed": (512, 384, 3, 8, False, (256, 256), 6, (64, 64), (20, 5), 0),
    "prc53_clip": (211, 157, 1, 8, True, (0, 0), 4, (64, 64), (), 1),        # precincts smaller than the nominal code block
    "prc97_rpcl": (400, 300, 3, 8, False, (0, 0), 5, (32, 32), (30, 8), 2),
    # BASELINE.json configs at their full size, compared by digest (the arrays would be gigabytes)
    "c2_full": (4096, 2160, 3, 8, False, (1024, 1024), 6, (64, 64), (40, 20, 10, 5), 0),
    "c3_full": (8192, 8192, 3, 16, True, (1024, 1024), 6, (64, 64), (), 0),
    "c5_53": (16384, 16384, 1, 8, True, (1024, 1024), 6, (64, 64), (), "sweep3"),
    "c5_97": (16384, 16384, 1, 8, False, (1024, 1024), 6, (64, 64), (10,), "sweep3"),
})
EXTRA = {
    "cinema2k": dict(cinema2k_fps=24),
    "prc53_64": dict(precincts=[(64, 64)]),
    "prc97_mixed": dict(precincts=[(256, 256), (128, 128)]),
    "prc53_clip": dict(precincts=[(32, 32), (16, 16)]),
    "prc97_rpcl": dict(precincts=[(128, 64), (64, 64), (32, 64)], progression=2),
}
# the HTJ2K block coder (grk_compress -M 64): the reference's T1HT replaced by the device's HT cleanup-pass kernels
CASES.update({
    "ht53": (300, 217, 3, 8, True, (128, 128), 5, (64, 64), (), 1),
    "ht53_gray16": (130, 90, 1, 16, True, (0, 0), 3, (32, 32), (), 0),
    "ht97": (256, 200, 3, 8, False, (128, 112), 6, (64, 64), (), 2),
    "ht97_12": (173, 131, 1, 12, False, (0, 0), 6, (32, 16), (), 0),
})
DIGEST_ONLY = ("c2_full", "c3_full", "c5_53", "c5_97")
ROI = {"roi53": (1, 5), "roi97": (0, 3)}
# region (window) decodes, grk_decompress -d x0,y0,x1,y1 (full-resolution image coordinates), optionally reduced
WINDOWS = {"rgb53_tiled": [((40, 30, 150, 120), 0), ((70, 10, 131, 75), 1)], "rgb97_layers": [((10, 20, 200, 180), 0), ((64, 64, 192, 160), 2)],
           "sweep53": [((300, 200, 1100, 900), 0)], "lazy53": [((100, 50, 260, 190), 1)]}
STYLES = {"ht53": 64, "ht53_gray16": 64, "ht97": 64, "ht97_12": 64, "lazy53": 1, "termall97": 4, "resetvsc53": 2 | 8, "allmodes53": 63, "lazyterm97": 1 | 4 | 16, "segsympterm16": 16 | 32}


def main():
    mode, out = sys.argv[1], sys.argv[2]
    names = sys.argv[3:] or ["gray53", "rgb53_tiled", "rgb97_layers", "rgb16_53"]
    shim = None
    if mode == "shim":
        shim = C.CDLL(os.path.join(ROOT, "integration", "_build", "libgrok_b200_tcd.so"), mode=C.RTLD_GLOBAL)
    import _libs
    from grokimagecompression_b200.synth import synthetic_planes
    res = {}
    for name in names:
        case = CASES[name]
        w, h, nc, prec, rev, tile, numres, cblk, rates, reduce = case[:10]
        kind = case[10] if len(case) > 10 else "smooth"
        if name.startswith("c5_"):  # 16K x 16K: a 4096 x 4096 synthetic image repeated, every 1024 x 1024 tile shifted differently
            base = synthetic_planes(4096, 4096, nc, prec, seed=16)
            ty, tx = np.arange(h, dtype=np.int32)[:, None] // 1024, np.arange(w, dtype=np.int32)[None, :] // 1024
            img = [((np.tile(b, (h // 4096, w // 4096)) + 5 * ty + 3 * tx) % (1 << prec)).astype(np.int32) for b in base]
            del base, ty, tx
        else:
            img = synthetic_planes(w, h, nc, prec, seed=len(name) + w, kind=kind)
        # rate-control algorithm 1 so that a single lossless layer is formed from the synced pass data
        cs = _libs.ref_encode_image(img, prec, tile=tile, numres=numres, cblk=cblk, irreversible=not rev, rates=rates, rc_algorithm=1, cblk_sty=STYLES.get(name, 0), roi=ROI.get(name, (-1, 0)),
                                    **EXTRA.get(name, {}))
        if name in DIGEST_ONLY:
            import hashlib
            digest = lambda a: np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)
            res[name + "_cs"] = digest(np.frombuffer(cs, np.uint8))
            res[name + "_cslen"] = np.array([len(cs)])
            dec = _libs.ref_decode_image(cs, nc, w, h)
            res[name + "_dec"] = digest(np.stack(dec))
            res[name + "_lossless"] = np.array([all((d == i).all() for d, i in zip(dec, img))])
            del dec
            if reduce == "sweep3":
                for r in (1, 2, 3):
                    res[name + f"_dec_r{r}"] = digest(np.stack(_libs.ref_decode_image(cs, nc, w, h, reduce=r)))
            continue
        res[name + "_cs"] = np.frombuffer(cs, np.uint8)
        res[name + "_dec"] = np.stack(_libs.ref_decode_image(cs, nc, w, h))
        if reduce == "sweep":
            for r in (1, 2, 3, 4):
                res[name + f"_dec_r{r}"] = np.stack(_libs.ref_decode_image(cs, nc, w, h, reduce=r))
            if len(rates) > 0 or True:
                res[name + "_dec_l1"] = np.stack(_libs.ref_decode_image(cs, nc, w, h, layers=1))
        elif reduce:
            res[name + "_dec_r"] = np.stack(_libs.ref_decode_image(cs, nc, w, h, reduce=reduce))
        for k, (win, red) in enumerate(WINDOWS.get(name, [])):
            res[name + f"_dec_w{k}"] = np.stack(_libs.ref_decode_image(cs, nc, w, h, reduce=red, window=win))
        res[name + "_img"] = np.stack(img)
    if shim is not None:
        shim.grok_b200_shim_calls.restype = C.c_uint64
        res["calls"] = np.array([shim.grok_b200_shim_calls(i) for i in range(8)], np.uint64)
        shim.grok_b200_shim_hulls.restype = C.c_uint64
        res["hulls"] = np.array([shim.grok_b200_shim_hulls()], np.uint64)
    np.savez_compressed(out, **res)


if __name__ == "__main__":
    main()
