"""ctypes loaders for the three native libraries the tests talk to.

oracle()  -> oracle/libgb_oracle.so       the CPU restatement (checker)
ref()     -> oracle/_ref/libgrkref_driver.so   the compiled, unmodified reference (checker of the checker)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PLUGIN_DIR = os.path.join(ROOT, "integration", "_build")  # where grk_plugin_load finds libgrok_plugin.so (`-g <dir>`)

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")

_oracle = None
_ref = None


def aligned(a, align=64):
    """Copy of `a` whose data pointer is `align`-byte aligned (the reference's AVX2 loops use
    aligned loads on tile buffers)."""
    a = np.ascontiguousarray(a)
    raw = np.empty(a.nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


class GboBlock(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("resno", "orient", "precno", "cblkno", "x0", "y0", "x1", "y1", "off_x", "off_y")]


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    so = os.path.join(ORACLE_DIR, "libgb_oracle.so")
    src = os.path.join(ORACLE_DIR, "gb_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ORACLE_DIR, "gb_oracle_ht.c"))):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libgb_oracle.so"])
    L = C.CDLL(so)
    L.gbo_dc_shift_fwd.argtypes = [i32p, C.c_uint64, C.c_int32, C.c_int]
    L.gbo_dc_shift_inv.argtypes = [i32p, C.c_uint64, C.c_int32, C.c_int, C.c_int32, C.c_int32]
    for n in ("gbo_rct_fwd", "gbo_rct_inv", "gbo_ict_fwd"):
        getattr(L, n).argtypes = [i32p, i32p, i32p, C.c_uint64]
    L.gbo_ict_inv.argtypes = [f32p, f32p, f32p, C.c_uint64]
    L.gbo_dwt_fwd.argtypes = [i32p] + [C.c_uint32] * 5 + [C.c_int]
    L.gbo_dwt_inv.argtypes = [i32p] + [C.c_uint32] * 6 + [C.c_int]
    L.gbo_quantise_block.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int32, i32p]
    L.gbo_quantise_block.restype = C.c_uint32
    L.gbo_dequantise_block.argtypes = [i32p, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_void_p, C.c_uint32]
    L.gbo_t1_encode_block.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                      C.POINTER(C.c_uint32), u32p, f64p, C.POINTER(C.c_uint64)]
    L.gbo_t1_decode_block.argtypes = [u8p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.gbo_rd_convex_hull.argtypes = [u32p, f64p, C.c_uint32, u16p]
    L.gbo_rd_convex_hull.restype = None
    L.gbo_t1_encode_block_sty.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                          C.POINTER(C.c_uint32), u32p, f64p, u8p, C.POINTER(C.c_uint64)]
    L.gbo_t1_decode_block_segs.argtypes = [u8p, u32p, u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.gbo_t1_decode_block_roi.argtypes = [u8p, u32p, u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.gbo_nmsedec_tables.argtypes = [i16p] * 4
    L.gbo_context_tables.argtypes = [u8p, u8p, u8p]
    L.gbo_enumerate_blocks.argtypes = [C.c_uint32] * 7 + [u32p, C.c_void_p]
    L.gbo_ht_encode_block.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
    L.gbo_ht_decode_block.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.gbo_ht_quantise_block.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_uint32, i32p]
    L.gbo_ht_quantise_block.restype = C.c_uint32
    L.gbo_ht_dequantise_block.argtypes = [i32p, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_uint32, C.c_void_p, C.c_uint32]
    _oracle = L
    return L


_tap = None


def tap():
    """oracle/_ref/libgrkref_tap.so (oracle/ref_tap.cpp): must be loaded BEFORE ref() so that the reference's stage calls
    resolve to the tap's forwarding definitions"""
    global _tap
    if _tap is None:
        assert _ref is None, "load the tap before the reference driver"
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libgrkref_tap.so"), mode=C.RTLD_GLOBAL)
        L.ref_tap_seconds.argtypes = [C.c_int]
        L.ref_tap_seconds.restype = C.c_double
        L.ref_tap_calls.argtypes = [C.c_int]
        L.ref_tap_calls.restype = C.c_uint64
        L.ref_tap_geometry.argtypes = [C.c_int]
        L.ref_tap_geometry.restype = C.c_uint64
        _tap = L
    return _tap


TAP_STAGES = ("dc_enc", "mct_enc", "dwt_enc", "t1_enc", "t1_dec", "dwt_dec", "mct_dec", "dc_dec", "encode_tile", "decode_tile", "simulate")


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libgrkref_driver.so"))


def ref():
    global _ref
    if _ref is not None:
        return _ref
    L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libgrkref_driver.so"))
    L.ref_init.argtypes = [C.c_uint32]
    for n in ("ref_mct_encode_rev", "ref_mct_decode_rev", "ref_mct_encode_irrev"):
        getattr(L, n).argtypes = [i32p, i32p, i32p, C.c_uint64]
    L.ref_mct_decode_irrev.argtypes = [f32p, f32p, f32p, C.c_uint64]
    L.ref_dwt_norm.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.ref_dwt_norm.restype = C.c_double
    L.ref_mct_norm.argtypes = [C.c_uint32, C.c_int]
    L.ref_mct_norm.restype = C.c_double
    L.ref_dwt_encode.argtypes = [i32p] + [C.c_uint32] * 5 + [C.c_int]
    L.ref_dwt_decode.argtypes = [i32p] + [C.c_uint32] * 6 + [C.c_int]
    L.ref_t1_encode_cblk.argtypes = [i32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_double, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, u8p,
                                     C.POINTER(C.c_uint32), u32p, u32p, f64p, C.POINTER(C.c_double)]
    L.ref_rd_convex_hull.argtypes = [u32p, f64p, C.c_uint32, u16p]
    L.ref_rd_convex_hull.restype = None
    L.ref_t1_decode_cblk.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, i32p]
    L.ref_ht_encode_block.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
    L.ref_ht_decode_block.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.ref_set_cblk_sty.argtypes = [C.c_uint32]
    L.ref_set_precincts.argtypes = [C.c_uint32, u32p, u32p]
    L.ref_set_progression.argtypes = [C.c_int]
    L.ref_set_roi.argtypes = [C.c_int32, C.c_uint32]
    L.ref_set_decode_area.argtypes = [C.c_uint32] * 4
    L.ref_t1_want_terms.argtypes = [u8p]
    L.ref_t1_decode_cblk_segs.argtypes = [u8p, u32p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_uint32, i32p]
    L.ref_qcd_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.c_int, u32p, u32p]
    L.ref_band_stepsize.argtypes = [C.c_uint32] * 5 + [C.c_int, C.c_uint32, C.c_uint32, C.c_float,
                                                       C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                                       C.POINTER(C.c_uint32)]
    L.ref_encode_image.argtypes = [C.c_uint32] * 5 + [C.POINTER(C.c_void_p)] + [C.c_uint32] * 5 + [C.c_int, C.c_uint32,
                                   C.c_void_p, C.c_int, C.c_uint32, u8p, C.c_uint64]
    L.ref_encode_image.restype = C.c_int64
    L.ref_decode_image.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint64, u32p]
    L.ref_plugin_encode_file.argtypes = [C.c_char_p, C.c_char_p] + [C.c_uint32] * 5 + [C.c_int, C.c_uint32, C.c_void_p, C.c_uint32,
                                         u8p, C.c_uint64]
    L.ref_plugin_encode_file.restype = C.c_int64
    L.ref_plugin_batch_encode.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p,
                                          C.c_uint32, u8p, C.c_uint64, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), C.c_uint32]
    L.ref_plugin_batch_encode.restype = C.c_int32
    L.ref_plugin_batch_decode.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"), C.c_uint64,
                                          np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.c_uint32]
    L.ref_plugin_batch_decode.restype = C.c_int32
    L.ref_plugin_decode.argtypes = [C.c_char_p, u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint64, u32p]
    L.ref_init(int(os.environ.get("GRK_REF_THREADS", "0")) or (os.cpu_count() or 1))
    _ref = L
    return L


def write_pnm(path, planes, prec):
    """binary PGM / PPM (8 or 16 bit big endian), what grk_compress -i reads"""
    h, w = planes[0].shape
    nc = len(planes)
    assert nc in (1, 3)
    a = np.stack([np.asarray(p) for p in planes], axis=-1)
    with open(path, "wb") as f:
        f.write(b"P%d\n%d %d\n%d\n" % (5 if nc == 1 else 6, w, h, (1 << prec) - 1))
        f.write(a.astype(">u2" if prec > 8 else np.uint8).tobytes())


def ref_plugin_encode_file(infile, area, tile=(0, 0), numres=6, cblk=(64, 64), irreversible=False, rates=(), rc_algorithm=1):
    """`grk_compress -g integration/_build -i infile`: grk_plugin_load / init / encode with libgrok_plugin.so (integration/
    grok_plugin_b200.cpp).  Returns the codestream bytes, or the negative status of ref_plugin_encode_file."""
    L = ref()
    cap = area * 4 * 3 + (1 << 20)
    out = np.zeros(cap, np.uint8)
    r = np.ascontiguousarray(rates, np.float64)
    n = L.ref_plugin_encode_file(PLUGIN_DIR.encode(), infile.encode(), tile[0] or 0, tile[1] or 0, numres,
                                 cblk[0], cblk[1], int(irreversible), len(r), r.ctypes.data if len(r) else None, rc_algorithm, out, cap)
    return bytes(out[:n]) if n > 0 else int(n)


def ref_encode_image(planes, prec, sgnd=0, tile=(0, 0), numres=6, cblk=(64, 64), irreversible=False, rates=(),
                     cinema2k_fps=0, rc_algorithm=0, cblk_sty=0, roi=(-1, 0), precincts=(), progression=-1):
    """planes: list of int32 [h,w] -> J2K codestream bytes produced by the unmodified reference (cblk_sty = grk_compress -M,
    roi = (component, shift) = -ROI c=..,U=.., precincts = [(w, h), ...] = -c, highest resolution first, progression = -p)"""
    L = ref()
    L.ref_set_cblk_sty(cblk_sty)
    L.ref_set_roi(roi[0], roi[1])
    pw = np.ascontiguousarray([p[0] for p in precincts] or [0], np.uint32)
    ph = np.ascontiguousarray([p[1] for p in precincts] or [0], np.uint32)
    L.ref_set_precincts(len(precincts), pw, ph)
    L.ref_set_progression(progression)
    h, w = planes[0].shape
    keep = [aligned(np.ascontiguousarray(p, np.int32)) for p in planes]
    pa = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
    cap = w * h * len(planes) * 4 + (1 << 20)
    out = np.zeros(cap, np.uint8)
    r = np.ascontiguousarray(rates, np.float64)
    n = L.ref_encode_image(len(planes), w, h, prec, sgnd, pa, tile[0] or 0, tile[1] or 0, numres, cblk[0], cblk[1],
                           int(irreversible), len(r), r.ctypes.data if len(r) else None, cinema2k_fps, rc_algorithm, out, cap)
    assert n > 0, "reference encode failed"
    return bytes(out[:n])


def ref_decode_image(cs, numcomps, width, height, reduce=0, layers=0, window=None):
    """window = (x0, y0, x1, y1) in full-resolution image coordinates: grk_decompress -d"""
    L = ref()
    buf = np.frombuffer(cs, np.uint8).copy()
    cd = lambda v: (v + (1 << reduce) - 1) >> reduce
    if window is None:
        L.ref_set_decode_area(0, 0, 0, 0)
        planes = [np.zeros((cd(height), cd(width)), np.int32) for _ in range(numcomps)]
    else:
        L.ref_set_decode_area(*window)
        planes = [np.zeros((cd(window[3]) - cd(window[1]), cd(window[2]) - cd(window[0])), np.int32) for _ in range(numcomps)]
    pa = (C.c_void_p * numcomps)(*[p.ctypes.data for p in planes])
    dims = np.zeros(4, np.uint32)
    rc = L.ref_decode_image(buf, len(buf), reduce, layers, pa, planes[0].size, dims)
    assert rc == 0, f"reference decode failed rc={rc}"
    assert (dims[0], dims[1], dims[2]) == (planes[0].shape[1], planes[0].shape[0], numcomps), dims
    return planes


def ref_plugin_batch_encode(in_dir, max_frames, frame_area, numres=6, cblk=(64, 64), irreversible=False, rates=(), rc_algorithm=1):
    """`grk_compress -g integration/_build -y in_dir`: the plugin owns the frame loop.  Returns the list of codestreams in file-name
    order, or the negative status of ref_plugin_batch_encode."""
    L = ref()
    cap = max_frames * (frame_area * 4 * 3 + (1 << 20))
    out = np.zeros(cap, np.uint8)
    lens = np.zeros(max_frames, np.uint64)
    r = np.ascontiguousarray(rates, np.float64)
    n = L.ref_plugin_batch_encode(PLUGIN_DIR.encode(), in_dir.encode(), numres, cblk[0], cblk[1], int(irreversible),
                                  len(r), r.ctypes.data if len(r) else None, rc_algorithm, out, cap, lens, max_frames)
    if n < 0:
        return int(n)
    res, off = [], 0
    for k in range(n):
        res.append(bytes(out[off:off + int(lens[k])]))
        off += int(lens[k])
    return res


def ref_plugin_batch_decode(in_dir, max_frames, plane_area, reduce=0):
    """`grk_decompress -g integration/_build -y in_dir`: the plugin walks the codestreams of the directory.  Returns a list of
    (comps, h, w) int32 arrays in file-name order, or the negative status of ref_plugin_batch_decode."""
    L = ref()
    out = np.zeros(max_frames * 3 * plane_area, np.int32)
    dims = np.zeros(max_frames * 3, np.uint32)
    n = L.ref_plugin_batch_decode(PLUGIN_DIR.encode(), in_dir.encode(), reduce, out, plane_area, dims, max_frames)
    if n < 0:
        return int(n)
    res = []
    for k in range(n):
        w, h, nc = (int(v) for v in dims[3 * k:3 * k + 3])
        res.append(np.stack([out[(3 * k + c) * plane_area:(3 * k + c) * plane_area + w * h].reshape(h, w) for c in range(nc)]))
    return res


def ref_plugin_decode(cs, numcomps, width, height, reduce=0, layers=0):
    """`grk_decompress -g integration/_build`: grk_plugin_load / init / decode.  Returns the planes, or the negative status."""
    L = ref()
    buf = np.frombuffer(cs, np.uint8).copy()
    cd = lambda v: (v + (1 << reduce) - 1) >> reduce
    planes = [np.zeros((cd(height), cd(width)), np.int32) for _ in range(numcomps)]
    pa = (C.c_void_p * numcomps)(*[p.ctypes.data for p in planes])
    dims = np.zeros(4, np.uint32)
    rc = L.ref_plugin_decode(PLUGIN_DIR.encode(), buf, len(buf), reduce, layers, pa, planes[0].size, dims)
    if rc:
        return int(rc)
    assert (dims[0], dims[1], dims[2]) == (planes[0].shape[1], planes[0].shape[0], numcomps), dims
    return planes


# ---- convenience wrappers ------------------------------------------------------------------

def oracle_t1_encode(blk, orient, do_rd=False, wbase=0.0):
    """blk: int32 [h,w] with 6 fractional bits. -> (bytes, numbps, rates, dists, nsym)"""
    L = oracle()
    h, w = blk.shape
    buf = np.zeros(w * h * 4 + 16, np.uint8)
    rates = np.zeros(128, np.uint32)
    dists = np.zeros(128, np.float64)
    nb = C.c_uint32()
    ns = C.c_uint64()
    n = L.gbo_t1_encode_block(np.ascontiguousarray(blk, np.int32).ravel(), w, h, orient, int(do_rd), wbase,
                              buf.ctypes.data + 1, C.byref(nb), rates, dists, C.byref(ns))
    assert n >= 0
    total = int(rates[n - 1]) if n else 0
    return bytes(buf[1:1 + total]), nb.value, rates[:n].copy(), dists[:n].copy(), ns.value


def ref_t1_encode(blk, orient, compno=0, level=0, qmfbid=1, stepsize=1.0, mct_norms=None, do_rd=False):
    L = ref()
    h, w = blk.shape
    buf = np.zeros(w * h * 4 + 64, np.uint8)
    rates = np.zeros(128, np.uint32)
    lens = np.zeros(128, np.uint32)
    dists = np.zeros(128, np.float64)
    nb = C.c_uint32()
    td = C.c_double()
    if mct_norms is not None:
        mn = np.ascontiguousarray(mct_norms, np.float64)
        mp, nm = mn.ctypes.data, len(mn)
    else:
        mp, nm = None, 0
    n = L.ref_t1_encode_cblk(np.ascontiguousarray(blk, np.int32).ravel(), w, h, orient, compno, level, qmfbid,
                             stepsize, 0, mp, nm, int(do_rd), buf, C.byref(nb), rates, lens, dists, C.byref(td))
    assert n >= 0
    total = int(rates[n - 1]) if n else 0
    return bytes(buf[:total]), nb.value, rates[:n].copy(), dists[:n].copy()


def oracle_t1_encode_sty(blk, orient, sty, do_rd=False, wbase=0.0):
    """as oracle_t1_encode with a code-block style byte -> (bytes, numbps, rates, dists, terms, nsym)"""
    L = oracle()
    h, w = blk.shape
    buf = np.zeros(w * h * 4 + 1024, np.uint8)  # terminated passes add flush bytes: tiny blocks outgrow 4wh
    rates = np.zeros(128, np.uint32)
    dists = np.zeros(128, np.float64)
    terms = np.zeros(128, np.uint8)
    nb = C.c_uint32()
    ns = C.c_uint64()
    n = L.gbo_t1_encode_block_sty(np.ascontiguousarray(blk, np.int32).ravel(), w, h, orient, sty, int(do_rd), wbase,
                                  buf.ctypes.data + 2, C.byref(nb), rates, dists, terms, C.byref(ns))
    assert n >= 0
    total = int(rates[n - 1]) if n else 0
    return bytes(buf[2:2 + total]), nb.value, rates[:n].copy(), dists[:n].copy(), terms[:n].copy(), ns.value


def ref_t1_encode_sty(blk, orient, sty, compno=0, level=0, qmfbid=1, stepsize=1.0, mct_norms=None, do_rd=False):
    """-> (bytes, numbps, rates, dists, terms) from the unmodified reference"""
    L = ref()
    h, w = blk.shape
    buf = np.zeros(w * h * 4 + 1024, np.uint8)
    rates = np.zeros(128, np.uint32)
    lens = np.zeros(128, np.uint32)
    dists = np.zeros(128, np.float64)
    terms = np.zeros(128, np.uint8)
    nb = C.c_uint32()
    td = C.c_double()
    if mct_norms is not None:
        mn = np.ascontiguousarray(mct_norms, np.float64)
        mp, nm = mn.ctypes.data, len(mn)
    else:
        mp, nm = None, 0
    L.ref_t1_want_terms(terms)
    n = L.ref_t1_encode_cblk(np.ascontiguousarray(blk, np.int32).ravel(), w, h, orient, compno, level, qmfbid,
                             stepsize, sty, mp, nm, int(do_rd), buf, C.byref(nb), rates, lens, dists, C.byref(td))
    assert n >= 0
    total = int(rates[n - 1]) if n else 0
    return bytes(buf[:total]), nb.value, rates[:n].copy(), dists[:n].copy(), terms[:n].copy()


def segments_from_passes(rates, terms, npasses=None):
    """codeword segments (len, passes) of the first `npasses` coding passes: a segment ends at every terminated pass"""
    n = len(rates) if npasses is None else npasses
    lens, cnts, start, cnt = [], [], 0, 0
    for i in range(n):
        cnt += 1
        if terms[i] or i == n - 1:
            lens.append(int(rates[i]) - start)
            cnts.append(cnt)
            start, cnt = int(rates[i]), 0
    return np.array(lens, np.uint32), np.array(cnts, np.uint32)


def oracle_t1_decode_segs(data, seg_len, seg_passes, numbps, orient, sty, w, h, roishift=0):
    out = np.zeros((h, w), np.int32)
    b = np.frombuffer(bytes(data) + b"\0\0", np.uint8).copy()
    rc = oracle().gbo_t1_decode_block_roi(b, np.ascontiguousarray(seg_len, np.uint32), np.ascontiguousarray(seg_passes, np.uint32),
                                          len(seg_len), numbps, roishift, orient, sty, w, h, out.ravel())
    assert rc == 0
    return out


def ref_t1_decode_segs(data, seg_len, seg_passes, numbps, orient, sty, w, h):
    out = np.zeros((h, w), np.int32)
    b = np.frombuffer(bytes(data) + b"\0\0", np.uint8).copy()
    rc = ref().ref_t1_decode_cblk_segs(b, np.ascontiguousarray(seg_len, np.uint32), np.ascontiguousarray(seg_passes, np.uint32),
                                       len(seg_len), numbps, orient, 0, sty, w, h, out.ravel())
    assert rc == 0
    return out


def oracle_t1_decode(data, numpasses, numbps, orient, w, h):
    out = np.zeros((h, w), np.int32)
    b = np.frombuffer(bytes(data) + b"\0\0", np.uint8).copy()
    rc = oracle().gbo_t1_decode_block(b, len(data), numpasses, numbps, orient, w, h, out.ravel())
    assert rc == 0
    return out


def ref_t1_decode(data, numpasses, numbps, orient, w, h):
    out = np.zeros((h, w), np.int32)
    b = np.frombuffer(bytes(data) + b"\0\0", np.uint8).copy()
    rc = ref().ref_t1_decode_cblk(b, len(data), numpasses, numbps, orient, 0, 0, w, h, out.ravel())
    assert rc == 0
    return out


def random_pass_tables(rng, count):
    """pass tables the way Tier-1 leaves them: byte counts with zeros, cumulative distortions that mostly decay, with flat
    stretches, equal slopes and the occasional negative step (the cases the feasibility tests branch on)"""
    out = []
    for it in range(count):
        n = int(rng.integers(1, 92))
        lens = rng.integers(0, 400, n).astype(np.uint32)
        lens[rng.random(n) < 0.15] = 0
        step = rng.random(n) * (2.0 ** rng.integers(-8, 40)) * np.exp(-np.arange(n) * rng.random() * 0.3)
        step[rng.random(n) < 0.1] = 0.0
        if it % 5 == 0:
            step[rng.random(n) < 0.1] *= -1.0
        if it % 7 == 0:  # proportional rate and distortion: equal slopes
            step = lens.astype(np.float64) * 3.0
        out.append((lens, np.cumsum(step)))
    return out
