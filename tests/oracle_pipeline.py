"""The whole tile path on the CPU oracle (oracle/gb_oracle.c), tile by tile, block by block, in the
order the reference's TileProcessor runs it.  Test infrastructure only."""
import ctypes as C

import numpy as np

from _libs import GboBlock, oracle, oracle_t1_decode, oracle_t1_encode


def enumerate_blocks(cp, numres_limit=None):
    O = oracle()
    prc = np.zeros(2 * cp.numres, np.uint32)
    for r in range(cp.numres):
        prc[2 * r], prc[2 * r + 1] = cp.prcw_expn[r], cp.prch_expn[r]
    n = O.gbo_enumerate_blocks(cp.x0, cp.y0, cp.x1, cp.y1, cp.numres, cp.cblkw_expn, cp.cblkh_expn, prc, None)
    arr = (GboBlock * max(n, 1))()
    O.gbo_enumerate_blocks(cp.x0, cp.y0, cp.x1, cp.y1, cp.numres, cp.cblkw_expn, cp.cblkh_expn, prc, C.addressof(arr))
    out = [arr[i] for i in range(n)]
    if numres_limit is not None:
        out = [b for b in out if b.resno < numres_limit]
    return out


def band_index(b):
    return 0 if b.resno == 0 else 3 * b.resno - 2 + (b.orient - 1)


def encode_tiles(tiles, planes, count_only=False):
    """tiles: list of tile dicts (params.image_tiles); planes: tile-major list of int32 [h,w].
    Returns (per-block list of dicts, coefficient planes)."""
    O = oracle()
    results, coeffs = [], []
    i = 0
    for t, tile in enumerate(tiles):
        comps = tile["comps"]
        work = [np.ascontiguousarray(planes[i + c], np.int32).copy() for c in range(len(comps))]
        i += len(comps)
        for c, cp in enumerate(comps):
            O.gbo_dc_shift_fwd(work[c].ravel(), work[c].size, cp.dc_shift, int(cp.qmfbid == 1))
        if tile.get("mct"):
            n = work[0].size
            (O.gbo_rct_fwd if comps[0].qmfbid == 1 else O.gbo_ict_fwd)(work[0].ravel(), work[1].ravel(), work[2].ravel(), n)
        for c, cp in enumerate(comps):
            if work[c].size:
                O.gbo_dwt_fwd(work[c].ravel(), cp.x0, cp.y0, cp.x1, cp.y1, cp.numres, int(cp.qmfbid == 1))
            coeffs.append(work[c])
            stride = cp.x1 - cp.x0
            for b in enumerate_blocks(cp):
                w, h = b.x1 - b.x0, b.y1 - b.y0
                bi = band_index(b)
                q = np.zeros((h, w), np.int32)
                base = work[c].ctypes.data + 4 * (b.off_y * stride + b.off_x)
                if cp.cblk_sty & 0x40:  # HTJ2K block coder: one cleanup pass, numbps 1 (T1HT.cpp:104-133)
                    if w == 0 or h == 0:
                        results.append(dict(tileno=t, compno=c, resno=b.resno, orient=b.orient, x0=b.x0, y0=b.y0, x1=b.x1, y1=b.y1,
                                            data=b"", numbps=0, rates=np.zeros(0, np.uint32), dists=np.zeros(0), nsym=0))
                        continue
                    k_msbs = int(cp.band_numbps[bi])
                    O.gbo_ht_quantise_block(base, stride, w, h, int(cp.qmfbid == 1), float(cp.stepsize[bi]), k_msbs, q.ravel())
                    buf = np.zeros(w * h * 8 + 8192, np.uint8)
                    n = O.gbo_ht_encode_block(q.ravel(), w, h, w, k_msbs, buf, len(buf))
                    assert n > 0
                    results.append(dict(tileno=t, compno=c, resno=b.resno, orient=b.orient, x0=b.x0, y0=b.y0, x1=b.x1, y1=b.y1,
                                        data=bytes(buf[:n]), numbps=1, rates=np.array([n], np.uint32), dists=np.zeros(1), nsym=0))
                    continue
                O.gbo_quantise_block(base, stride, w, h, int(cp.qmfbid == 1), int(cp.inv_step[bi]), q.ravel())
                data, numbps, rates, dists, nsym = oracle_t1_encode(q, b.orient, bool(tile.get("rate_control")), cp.rd_weight[bi])
                results.append(dict(tileno=t, compno=c, resno=b.resno, orient=b.orient, x0=b.x0, y0=b.y0, x1=b.x1, y1=b.y1,
                                    data=data, numbps=numbps, rates=rates, dists=dists, nsym=nsym))
    return results, coeffs


def decode_tiles(tiles, block_inputs):
    """block_inputs: per block (in plan order) dict(data=bytes, numbps, numpasses). Returns decoded planes."""
    O = oracle()
    out = []
    k = 0
    for tile in tiles:
        comps = tile["comps"]
        work = []
        for cp in comps:
            nd = tile.get("numres_decode") or cp.numres
            nd = min(nd, cp.numres)
            top = cp.numres - nd
            cd = lambda v: (v + (1 << top) - 1) >> top
            w, h = cd(cp.x1) - cd(cp.x0), cd(cp.y1) - cd(cp.y0)
            plane = np.zeros((h, w), np.int32)
            for b in enumerate_blocks(cp, nd):
                bw, bh = b.x1 - b.x0, b.y1 - b.y0
                inp = block_inputs[k]
                k += 1
                if not inp["numpasses"] or not len(inp["data"]):
                    continue
                base = plane.ctypes.data + 4 * (b.off_y * w + b.off_x)
                if cp.cblk_sty & 0x40:
                    k_msbs = int(cp.band_numbps[band_index(b)]) - int(inp["numbps"])
                    dec = np.zeros((bh, bw), np.int32)
                    d = np.frombuffer(inp["data"], np.uint8).copy()
                    if O.gbo_ht_decode_block(d, len(d), k_msbs, bw, bh, bw, dec.ravel()) == 0:
                        O.gbo_ht_dequantise_block(dec.ravel(), bw, bh, int(cp.qmfbid == 1), float(cp.stepsize[band_index(b)]), k_msbs, base, w)
                    continue
                dec = oracle_t1_decode(inp["data"], inp["numpasses"], inp["numbps"], b.orient, bw, bh)
                O.gbo_dequantise_block(dec.ravel(), bw, bh, int(cp.qmfbid == 1), float(cp.stepsize[band_index(b)]), base, w)
            if plane.size:
                O.gbo_dwt_inv(plane.ravel(), cp.x0, cp.y0, cp.x1, cp.y1, cp.numres, nd, int(cp.qmfbid == 1))
            work.append(plane)
        if tile.get("mct") and work[0].size:
            n = work[0].size
            if comps[0].qmfbid == 1:
                O.gbo_rct_inv(work[0].ravel(), work[1].ravel(), work[2].ravel(), n)
            else:
                f = [w.view(np.float32) for w in work[:3]]
                O.gbo_ict_inv(f[0].ravel(), f[1].ravel(), f[2].ravel(), n)
        for c, cp in enumerate(comps):
            lo, hi = (-(1 << (cp.prec - 1)), (1 << (cp.prec - 1)) - 1) if cp.sgnd else (0, (1 << cp.prec) - 1)
            O.gbo_dc_shift_inv(work[c].ravel(), work[c].size, cp.dc_shift, int(cp.qmfbid == 1), lo, hi)
            out.append(work[c])
    return out
