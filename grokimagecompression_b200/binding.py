"""ctypes binding of libgrok_b200.so (include/grok_b200.h).

The shared library IS the product; this module only marshals numpy buffers into its C ABI.
There is no CPU path here: if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

MAX_RES = 33
ABI_VERSION = 3  # GB200_ABI_VERSION of include/grok_b200.h
MAX_BANDS = 3 * MAX_RES - 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GB200_LIB") or os.path.join(_HERE, "libgrok_b200.so")  # GB200_LIB: A/B builds of the same ABI


class GrokB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgrok_b200 error {code}: {msg}")
        self.code = code


class CompParams(C.Structure):
    _fields_ = [
        ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
        ("numres", C.c_uint32), ("cblkw_expn", C.c_uint32), ("cblkh_expn", C.c_uint32),
        ("prcw_expn", C.c_uint32 * MAX_RES), ("prch_expn", C.c_uint32 * MAX_RES),
        ("qmfbid", C.c_uint32), ("prec", C.c_uint32), ("sgnd", C.c_uint32), ("dc_shift", C.c_int32),
        ("cblk_sty", C.c_uint32), ("roishift", C.c_uint32),
        ("stepsize", C.c_float * MAX_BANDS), ("inv_step", C.c_uint32 * MAX_BANDS),
        ("band_numbps", C.c_uint32 * MAX_BANDS), ("rd_weight", C.c_double * MAX_BANDS),
    ]


class TileParams(C.Structure):
    _fields_ = [("numcomps", C.c_uint32), ("mct", C.c_uint32), ("rate_control", C.c_uint32),
                ("numres_decode", C.c_uint32), ("comps", C.POINTER(CompParams))]


class CblkInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("tileno", "compno", "resno", "bandno", "precno", "cblkno",
                                          "x0", "y0", "x1", "y1", "band_index", "pass_offset", "max_passes")]


CBLK_ENC_DTYPE = np.dtype([("numbps", np.uint32), ("numpasses", np.uint32), ("data_len", np.uint32),
                           ("decisions", np.uint32), ("data_offset", np.uint64)])
CBLK_DEC_DTYPE = np.dtype([("numbps", np.uint32), ("numpasses", np.uint32), ("data_len", np.uint32),
                           ("reserved", np.uint32), ("data_offset", np.uint64)])
CBLK_SEG_DTYPE = np.dtype([("len", np.uint32), ("numpasses", np.uint32)])
CBLK_INFO_DTYPE = np.dtype([(n, np.uint32) for n, _ in CblkInfo._fields_])
T1_BLOCK_DTYPE = np.dtype([("x", np.uint32), ("y", np.uint32), ("w", np.uint32), ("h", np.uint32),
                           ("orient", np.uint32), ("qmfbid", np.uint32), ("inv_step", np.uint32),
                           ("stepsize", np.float32), ("rd_weight", np.float64), ("cblk_sty", np.uint32), ("roishift", np.uint32)],
                          align=True)

# every symbol include/grok_b200.h declares
SYMBOLS = [
    "gb200_abi_version", "gb200_last_error", "gb200_device_count", "gb200_create", "gb200_destroy", "gb200_launch_count", "gb200_stream",
    "gb200_plan_create", "gb200_plan_destroy", "gb200_plan_num_blocks", "gb200_plan_num_pass_slots",
    "gb200_plan_num_samples", "gb200_plan_blocks", "gb200_plan_data_capacity", "gb200_precinct_grid", "gb200_enumerate_blocks",
    "gb200_plan_set_sample_bytes", "gb200_encode_tiles_packed", "gb200_decode_tiles_packed", "gb200_encode_upload_packed",
    "gb200_decode_download_packed",
    "gb200_encode_tiles", "gb200_decode_tiles", "gb200_encode_upload", "gb200_encode_run", "gb200_encode_download",
    "gb200_decode_upload", "gb200_decode_run", "gb200_decode_download", "gb200_sync", "gb200_encode_slopes",
    "gb200_encode_stash", "gb200_encode_restore",
    "gb200_encode_run_stage", "gb200_decode_run_stage", "gb200_encode_get_coefficients",
    "gb200_decode_set_coefficients", "gb200_decode_set_segments", "gb200_t1_decode_blocks_segs",
    "gb200_mct_encode_rev", "gb200_mct_decode_rev", "gb200_mct_encode_irrev", "gb200_mct_decode_irrev",
    "gb200_dc_shift_encode", "gb200_dc_shift_decode", "gb200_dwt_encode", "gb200_dwt_decode",
    "gb200_t1_encode_blocks", "gb200_t1_decode_blocks",
]

_lib = None


def lib():
    """Load libgrok_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C grokimagecompression_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    L.gb200_abi_version.restype = C.c_int
    L.gb200_last_error.restype = C.c_char_p
    L.gb200_device_count.restype = C.c_int
    L.gb200_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.gb200_destroy.argtypes = [vp]
    L.gb200_destroy.restype = None
    L.gb200_launch_count.argtypes = [vp]
    L.gb200_launch_count.restype = u64
    L.gb200_stream.argtypes = [vp]
    L.gb200_stream.restype = vp
    L.gb200_sync.argtypes = [vp]
    L.gb200_plan_create.argtypes = [vp, u32, C.POINTER(TileParams), C.c_int, C.POINTER(vp)]
    L.gb200_plan_destroy.argtypes = [vp]
    L.gb200_plan_destroy.restype = None
    for n in ("gb200_plan_num_blocks", "gb200_plan_num_pass_slots", "gb200_plan_num_samples", "gb200_plan_data_capacity"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = u64
    L.gb200_plan_blocks.argtypes = [vp]
    L.gb200_plan_blocks.restype = C.POINTER(CblkInfo)
    L.gb200_enumerate_blocks.argtypes = [C.POINTER(CompParams), u32, vp, u64]
    L.gb200_enumerate_blocks.restype = u64
    L.gb200_precinct_grid.argtypes = [C.POINTER(CompParams), u32, C.POINTER(u32), C.POINTER(u32)]
    L.gb200_plan_set_sample_bytes.argtypes = [vp, u32]
    L.gb200_encode_tiles_packed.argtypes = [vp, C.POINTER(vp), vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.gb200_decode_tiles_packed.argtypes = [vp, vp, vp, u64, C.POINTER(vp)]
    L.gb200_encode_upload_packed.argtypes = [vp, C.POINTER(vp)]
    L.gb200_decode_download_packed.argtypes = [vp, C.POINTER(vp)]
    L.gb200_encode_tiles.argtypes = [vp, C.POINTER(vp), vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.gb200_decode_tiles.argtypes = [vp, vp, vp, u64, C.POINTER(vp)]
    L.gb200_encode_upload.argtypes = [vp, C.POINTER(vp)]
    L.gb200_encode_run.argtypes = [vp]
    L.gb200_encode_download.argtypes = [vp, vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.gb200_encode_slopes.argtypes = [vp, vp]
    L.gb200_decode_upload.argtypes = [vp, vp, vp, u64]
    L.gb200_decode_run.argtypes = [vp]
    L.gb200_decode_download.argtypes = [vp, C.POINTER(vp)]
    L.gb200_encode_stash.argtypes = [vp]
    L.gb200_encode_restore.argtypes = [vp]
    L.gb200_encode_run_stage.argtypes = [vp, C.c_int]
    L.gb200_decode_run_stage.argtypes = [vp, C.c_int]
    L.gb200_encode_get_coefficients.argtypes = [vp, u32, u32, vp]
    L.gb200_decode_set_coefficients.argtypes = [vp, u32, u32, vp]
    for n in ("gb200_mct_encode_rev", "gb200_mct_decode_rev", "gb200_mct_encode_irrev", "gb200_mct_decode_irrev"):
        getattr(L, n).argtypes = [vp, vp, vp, vp, u64]
    L.gb200_dc_shift_encode.argtypes = [vp, vp, u64, i32, C.c_int]
    L.gb200_dc_shift_decode.argtypes = [vp, vp, u64, i32, C.c_int, i32, i32]
    L.gb200_dwt_encode.argtypes = [vp, vp, u32, u32, u32, u32, u32, C.c_int]
    L.gb200_dwt_decode.argtypes = [vp, vp, u32, u32, u32, u32, u32, u32, C.c_int]
    L.gb200_t1_encode_blocks.argtypes = [vp, vp, u32, u32, u32, vp, C.c_int, u32, vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.gb200_t1_decode_blocks.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp, u64]
    L.gb200_t1_decode_blocks_segs.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp, vp, vp, u64]
    L.gb200_decode_set_segments.argtypes = [vp, vp, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise GrokB200Error(rc, lib().gb200_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Context:
    """One CUDA device + stream (gb200_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().gb200_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            lib().gb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def stream(self):
        return lib().gb200_stream(self._h)

    def launch_count(self):
        return int(lib().gb200_launch_count(self._h))

    def sync(self):
        check(lib().gb200_sync(self._h))

    # ---- stage-level calls (in place on numpy arrays) ----------------------------------------
    def mct_encode_rev(self, c0, c1, c2):
        check(lib().gb200_mct_encode_rev(self._h, _ptr(c0), _ptr(c1), _ptr(c2), c0.size))

    def mct_decode_rev(self, c0, c1, c2):
        check(lib().gb200_mct_decode_rev(self._h, _ptr(c0), _ptr(c1), _ptr(c2), c0.size))

    def mct_encode_irrev(self, c0, c1, c2):
        check(lib().gb200_mct_encode_irrev(self._h, _ptr(c0), _ptr(c1), _ptr(c2), c0.size))

    def mct_decode_irrev(self, c0, c1, c2):
        check(lib().gb200_mct_decode_irrev(self._h, _ptr(c0), _ptr(c1), _ptr(c2), c0.size))

    def dc_shift_encode(self, x, shift, qmfbid):
        check(lib().gb200_dc_shift_encode(self._h, _ptr(x), x.size, shift, qmfbid))

    def dc_shift_decode(self, x, shift, qmfbid, lo, hi):
        check(lib().gb200_dc_shift_decode(self._h, _ptr(x), x.size, shift, qmfbid, lo, hi))

    def dwt_encode(self, buf, x0, y0, x1, y1, numres, qmfbid):
        check(lib().gb200_dwt_encode(self._h, _ptr(buf), x0, y0, x1, y1, numres, qmfbid))

    def dwt_decode(self, buf, x0, y0, x1, y1, numres, numres_decode, qmfbid):
        check(lib().gb200_dwt_decode(self._h, _ptr(buf), x0, y0, x1, y1, numres, numres_decode, qmfbid))

    def t1_encode_blocks(self, plane, blocks, rate_control=False, max_passes=100):
        """plane: int32 [H,W]; blocks: T1_BLOCK_DTYPE array -> (results, rates[n,max_passes], dists, data bytes)"""
        plane = np.ascontiguousarray(plane, np.int32)
        blocks = np.ascontiguousarray(blocks, T1_BLOCK_DTYPE)
        n = len(blocks)
        res = np.zeros(n, CBLK_ENC_DTYPE)
        rates = np.zeros((n, max_passes), np.uint32)
        dists = np.zeros((n, max_passes), np.float64)
        cap = int(sum(int(b["w"]) * int(b["h"]) * 4 + 32 for b in blocks)) + 64
        data = np.zeros(cap, np.uint8)
        dl = C.c_uint64()
        check(lib().gb200_t1_encode_blocks(self._h, _ptr(plane), plane.shape[1], plane.shape[0], n, _ptr(blocks),
                                            int(rate_control), max_passes, _ptr(res), _ptr(rates), _ptr(dists),
                                            _ptr(data), cap, C.byref(dl)))
        return res, rates, dists, data[:dl.value]

    def t1_decode_blocks(self, shape, blocks, inputs, data, seg_start=None, segs=None):
        """seg_start (len(blocks) + 1 prefix offsets) / segs (CBLK_SEG_DTYPE): codeword segments of TERMALL / LAZY blocks"""
        plane = np.zeros(shape, np.int32)
        blocks = np.ascontiguousarray(blocks, T1_BLOCK_DTYPE)
        inputs = np.ascontiguousarray(inputs, CBLK_DEC_DTYPE)
        data = np.ascontiguousarray(data, np.uint8)
        if seg_start is None:
            check(lib().gb200_t1_decode_blocks(self._h, _ptr(plane), shape[1], shape[0], len(blocks), _ptr(blocks),
                                                _ptr(inputs), _ptr(data) if data.size else None, data.size))
        else:
            seg_start = np.ascontiguousarray(seg_start, np.uint32)
            segs = np.ascontiguousarray(segs, CBLK_SEG_DTYPE)
            check(lib().gb200_t1_decode_blocks_segs(self._h, _ptr(plane), shape[1], shape[0], len(blocks), _ptr(blocks),
                                                     _ptr(inputs), _ptr(seg_start), _ptr(segs) if segs.size else _ptr(np.zeros(1, CBLK_SEG_DTYPE)),
                                                     _ptr(data) if data.size else None, data.size))
        return plane


class Plan:
    """Geometry + block table + device buffers of a batch of tiles (gb200_plan)."""

    def __init__(self, ctx, tiles, encoder=True, sample_bytes=4):
        """tiles: list of dicts {numcomps, mct, rate_control, numres_decode, comps: [CompParams]};
        sample_bytes 1 / 2: the host planes are packed uint8 / uint16 (int8 / int16 for signed components)"""
        self.ctx = ctx
        self.sample_bytes = int(sample_bytes)
        self.comp_signed = [bool(cp.sgnd) for t in tiles for cp in t["comps"]]
        self.encoder = bool(encoder)
        self._keep = []
        arr = (TileParams * len(tiles))()
        self.comp_shapes = []
        for i, t in enumerate(tiles):
            comps = (CompParams * len(t["comps"]))(*t["comps"])
            self._keep.append(comps)
            arr[i].numcomps = len(t["comps"])
            arr[i].mct = int(t.get("mct", 0))
            arr[i].rate_control = int(t.get("rate_control", 0))
            arr[i].numres_decode = int(t.get("numres_decode", 0))
            arr[i].comps = comps
            for cp in t["comps"]:
                nd = cp.numres if encoder or not arr[i].numres_decode else min(arr[i].numres_decode, cp.numres)
                top = cp.numres - nd
                cd = lambda v: (v + (1 << top) - 1) >> top
                self.comp_shapes.append((cd(cp.y1) - cd(cp.y0), cd(cp.x1) - cd(cp.x0)))
        self._h = C.c_void_p()
        check(lib().gb200_plan_create(ctx.handle, len(tiles), arr, int(self.encoder), C.byref(self._h)))
        L = lib()
        if self.sample_bytes != 4:
            check(L.gb200_plan_set_sample_bytes(self._h, self.sample_bytes))
        self.num_blocks = int(L.gb200_plan_num_blocks(self._h))
        self.num_pass_slots = int(L.gb200_plan_num_pass_slots(self._h))
        self.num_samples = int(L.gb200_plan_num_samples(self._h))
        self.data_capacity = int(L.gb200_plan_data_capacity(self._h))
        if self.num_blocks:
            p = L.gb200_plan_blocks(self._h)
            buf = (CblkInfo * self.num_blocks).from_address(C.addressof(p.contents))
            self.blocks = np.frombuffer(buf, CBLK_INFO_DTYPE).copy()
        else:
            self.blocks = np.zeros(0, CBLK_INFO_DTYPE)

    def close(self):
        if self._h:
            lib().gb200_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr_array(arrays):
        pa = (C.c_void_p * len(arrays))()
        for i, a in enumerate(arrays):
            pa[i] = a.ctypes.data if isinstance(a, np.ndarray) else int(a)
        return pa

    def alloc_encode_outputs(self, data_capacity=None):
        cap = self.data_capacity if data_capacity is None else int(data_capacity)
        return (np.zeros(self.num_blocks, CBLK_ENC_DTYPE), np.zeros(max(self.num_pass_slots, 1), np.uint32),
                np.zeros(max(self.num_pass_slots, 1), np.float64), np.zeros(max(cap, 16), np.uint8))

    # whole path on host buffers
    def encode(self, planes, outputs=None):
        """planes: list of int32 arrays (tile-major, component-minor). -> (blocks, rates, dists, data)"""
        res, rates, dists, data = outputs if outputs is not None else self.alloc_encode_outputs()
        dl = C.c_uint64()
        pa = self._ptr_array(self._check_planes(planes))
        f = lib().gb200_encode_tiles if self.sample_bytes == 4 else lib().gb200_encode_tiles_packed
        check(f(self._h, pa, _ptr(res), _ptr(rates), _ptr(dists), _ptr(data), data.size, C.byref(dl)))
        return res, rates, dists, data[:dl.value]

    def sample_dtype(self, i=0):
        """numpy dtype of host plane i (tile-major, component-minor)"""
        if self.sample_bytes == 4:
            return np.dtype(np.int32)
        return np.dtype({(1, False): np.uint8, (1, True): np.int8, (2, False): np.uint16, (2, True): np.int16}[(self.sample_bytes, self.comp_signed[i])])

    def _check_planes(self, planes):
        for i, a in enumerate(planes):
            if isinstance(a, np.ndarray):
                assert a.dtype == self.sample_dtype(i) and a.flags["C_CONTIGUOUS"], (i, a.dtype, self.sample_dtype(i))
        return planes

    def _alloc_planes(self):
        return [np.zeros(s, self.sample_dtype(i)) for i, s in enumerate(self.comp_shapes)]

    def encode_upload(self, planes):
        f = lib().gb200_encode_upload if self.sample_bytes == 4 else lib().gb200_encode_upload_packed
        check(f(self._h, self._ptr_array(self._check_planes(planes))))

    def encode_stash(self):
        check(lib().gb200_encode_stash(self._h))

    def encode_restore(self):
        check(lib().gb200_encode_restore(self._h))

    def encode_run(self):
        check(lib().gb200_encode_run(self._h))

    def encode_run_stage(self, stage):
        check(lib().gb200_encode_run_stage(self._h, stage))

    def encode_download(self, outputs=None):
        res, rates, dists, data = outputs if outputs is not None else self.alloc_encode_outputs()
        dl = C.c_uint64()
        check(lib().gb200_encode_download(self._h, _ptr(res), _ptr(rates), _ptr(dists), _ptr(data), data.size, C.byref(dl)))
        return res, rates, dists, data[:dl.value]

    def encode_slopes(self):
        """feasible truncation points of the last encode run: uint16 ln(slope) in 8.8 fixed point per pass slot, 0 = none
        (RateControl::convexHull on the device)"""
        out = np.zeros(max(self.num_pass_slots, 1), np.uint16)
        check(lib().gb200_encode_slopes(self._h, _ptr(out)))
        return out[:self.num_pass_slots]

    def coefficients(self, tileno, compno):
        idx = self._plane_index(tileno, compno)
        out = np.zeros(self.comp_shapes[idx], np.int32)
        check(lib().gb200_encode_get_coefficients(self._h, tileno, compno, _ptr(out)))
        return out

    def _plane_index(self, tileno, compno):
        # planes are tile-major; tiles may differ in component count only in theory
        n = 0
        for t, comps in enumerate(self._keep):
            if t == tileno:
                return n + compno
            n += len(comps)
        raise IndexError(tileno)

    def decode(self, inputs, data, out=None):
        inputs = np.ascontiguousarray(inputs, CBLK_DEC_DTYPE)
        data = np.ascontiguousarray(data, np.uint8)
        planes = self._check_planes(out) if out is not None else self._alloc_planes()
        f = lib().gb200_decode_tiles if self.sample_bytes == 4 else lib().gb200_decode_tiles_packed
        check(f(self._h, _ptr(inputs), _ptr(data) if data.size else None, data.size, self._ptr_array(planes)))
        return planes

    def decode_upload(self, inputs, data):
        inputs = np.ascontiguousarray(inputs, CBLK_DEC_DTYPE)
        data = np.ascontiguousarray(data, np.uint8)
        self._dec_keep = (inputs, data)
        check(lib().gb200_decode_upload(self._h, _ptr(inputs), _ptr(data) if data.size else None, data.size))

    def set_segments(self, seg_start=None, segs=None):
        """codeword segments for the following decode calls (None = single-segment blocks)"""
        if seg_start is None:
            check(lib().gb200_decode_set_segments(self._h, None, None))
            return
        seg_start = np.ascontiguousarray(seg_start, np.uint32)
        segs = np.ascontiguousarray(segs, CBLK_SEG_DTYPE)
        if segs.size == 0:
            segs = np.zeros(1, CBLK_SEG_DTYPE)
        check(lib().gb200_decode_set_segments(self._h, _ptr(seg_start), _ptr(segs)))

    def decode_run(self):
        check(lib().gb200_decode_run(self._h))

    def decode_run_stage(self, stage):
        check(lib().gb200_decode_run_stage(self._h, stage))

    def decode_download(self, out=None):
        planes = self._check_planes(out) if out is not None else self._alloc_planes()
        f = lib().gb200_decode_download if self.sample_bytes == 4 else lib().gb200_decode_download_packed
        check(f(self._h, self._ptr_array(planes)))
        return planes

    def set_coefficients(self, tileno, compno, arr):
        arr = np.ascontiguousarray(arr, np.int32)
        check(lib().gb200_decode_set_coefficients(self._h, tileno, compno, _ptr(arr)))
        self.ctx.sync()
