#!/usr/bin/env python
"""bench.py -- encode/decode throughput of the tile-coding hot path on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (rank 0 only)

One step = one pass of the hot path over one batch of synthetic input: the workload image is encoded
(DC shift + ICT/RCT + DWT + quantise + Tier-1) and the resulting code blocks are decoded again
(Tier-1 + de-quantise + inverse DWT + inverse MCT + clamp).  Every pixel therefore crosses the path
twice per step and the headline metric is  2 * pixels / step time.

Workload at N=1: BASELINE.json configs[1] -- 4096x2160 RGB 8-bit, irreversible 9/7 + ICT, quantised,
rate control on (per-pass distortion for 4 quality layers), 1024x1024 tiles.  The main line at N>1 is weak
scaling (every rank codes its own frame of the same shape, no collective); the `strong` object beside it
holds the partitions north_star names -- configs[2] dealt by tile, configs[3] by frame, configs[4] (decode)
by tile -- with the results gathered on rank 0 in unit order and compared with a one-GPU run of the same
work.

The reference arm (--impl reference, and `cpu_baseline` of the default run) times the UNMODIFIED reference
codec on the host cores with a tap on its tile coder (oracle/ref_tap.cpp): `value` is its HOT PATH ONLY
(level shift, MCT, DWT, Tier-1 and their inverses -- the same work the GPU arm times), the whole codec
(with its host-side PCRD / Tier-2 / codestream writer) is reported beside it.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: width, height, comps, prec, reversible, tile, numres, cblk exponents, layers (compression ratios)
    "c1": dict(width=2048, height=2048, comps=1, prec=8, reversible=True, tile=(None, None), numres=6, cblk=(6, 6), rates=()),
    "c2": dict(width=4096, height=2160, comps=3, prec=8, reversible=False, tile=(1024, 1024), numres=6, cblk=(6, 6),
               rates=(40, 20, 10, 5)),
    "c3": dict(width=8192, height=8192, comps=3, prec=16, reversible=True, tile=(1024, 1024), numres=6, cblk=(6, 6), rates=()),
    "c4": dict(width=2048, height=1080, comps=3, prec=12, reversible=False, tile=(None, None), numres=6, cblk=(5, 5), rates=(10,),
               prc=[7] + [8] * 32, cinema=24),
    # 30 of the 240 frames of configs[3]: what one GPU of eight gets when the batch is sharded by frame; one plan, one launch per stage
    "c4x30": dict(width=2048, height=1080, comps=3, prec=12, reversible=False, tile=(None, None), numres=6, cblk=(5, 5), rates=(10,), frames=30,
                  prc=[7] + [8] * 32, cinema=24),
}
# the same images through the HTJ2K block coder (grk_compress -M 64): no rate control in the reference's HT path
WORKLOADS["c2ht"] = dict(WORKLOADS["c2"], rates=(), ht=True)
WORKLOADS["c3ht"] = dict(WORKLOADS["c3"], ht=True)
WORKLOAD_TEXT = {
    "c1": "configs[0]: 2048x2048 8-bit gray, lossless 5/3, 1 tile, 64x64 blocks, 5 levels",
    "c2": "configs[1]: 4096x2160 RGB 8-bit, irreversible 9/7 + ICT, quantised, 4 quality layers, 1024x1024 tiles",
    "c3": "configs[2]: 8192x8192 3x16-bit, lossless 5/3 + RCT, 1024x1024 tiles",
    "c4": "configs[3]: one DCI 2K frame 2048x1080 3x12-bit, 9/7 + ICT, 32x32 blocks, 128/256 precincts",
    "c4x30": "configs[3]: batch of 30 DCI 2K frames 2048x1080 3x12-bit (240 frames over 8 GPUs), 9/7 + ICT, 32x32 blocks, 128/256 precincts",
    "c2ht": "configs[1] image with the HTJ2K block coder (-M 64): 4096x2160 RGB 8-bit, 9/7 + ICT, 1024x1024 tiles",
    "c3ht": "configs[2] image with the HTJ2K block coder (-M 64): 8192x8192 3x16-bit, 5/3 + RCT, 1024x1024 tiles",
}
STRONG_WORKERS = 3  # host workers per rank in the strong-scaling encode partitions (own context, plans, pinned buffers)
T1_SOURCES = ("t1_enc.cu", "t1_dec.cu", "t1_tables.cuh", "common.cuh")
DWT_SOURCES = ("dwt_stream.cuh", "dwt.cu", "dwt_plane.h")


def source_sha1(names):
    """digest of the kernel sources an ncu-derived constant in profiles/ belongs to: a changed kernel invalidates it"""
    h = hashlib.sha1()
    for n in names:
        with open(os.path.join(ROOT, "grokimagecompression_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def sample_bytes_of(prec):
    return 1 if prec <= 8 else 2


def make_workload(name, seed):
    from grokimagecompression_b200 import params as P
    from grokimagecompression_b200.synth import synthetic_planes
    w = WORKLOADS[name]
    rc = len(w["rates"]) > 0
    prc = w.get("prc", 15)
    img, tiles_e, tiles_d, planes = None, [], [], []
    for f in range(w.get("frames", 1)):  # frames of a batch are independent images: more tiles of the same plan
        fimg = synthetic_planes(w["width"], w["height"], w["comps"], w["prec"], seed=seed + f)
        img = fimg if img is None else img  # the correctness gate looks at the first frame
        tiles_e += P.image_tiles(w["width"], w["height"], w["comps"], w["prec"], w["reversible"], w["tile"], w["numres"],
                                 rate_control=rc, cblk_expn=w["cblk"], prc_expn=prc, ht=w.get("ht", False))
        tiles_d += P.image_tiles(w["width"], w["height"], w["comps"], w["prec"], w["reversible"], w["tile"], w["numres"],
                                 cblk_expn=w["cblk"], encoder=False, prc_expn=prc, ht=w.get("ht", False))
        planes += P.split_planes(fimg, w["width"], w["height"], w["tile"])
    return w, img, tiles_e, tiles_d, planes


def dwt_algorithmic_bytes(tiles):
    """8 bytes per sample of the LL region of every level (one read + one write), SURVEY.md section 8(d)"""
    total, launches = 0, 0
    for t in tiles:
        for cp in t["comps"]:
            for lvl in range(cp.numres - 1):
                cd = lambda v: (v + (1 << lvl) - 1) >> lvl
                total += 8 * (cd(cp.x1) - cd(cp.x0)) * (cd(cp.y1) - cd(cp.y0))
            launches = max(launches, cp.numres - 1)
    return total, launches


# ---- the reference's CPU path -------------------------------------------------------------------------------------

def cpu_reference_run(name, steps, warmup, budget_s=150.0, seed=1):
    """Times the unmodified reference (oracle/_ref) on the host cores: encode + decode through its public API with memory
    streams, with the tap (oracle/ref_tap.cpp) in front of its tile coder.  Returns the wall clock of the whole codec and
    the time spent inside the hot-path stage calls only.  The sample is the full image unless that would blow the time
    budget, in which case a crop of whole tile rows is used."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    from grokimagecompression_b200.synth import synthetic_planes
    if not _libs.have_ref():
        return None
    cores = os.cpu_count() or 1
    os.environ["GRK_REF_THREADS"] = str(cores)
    tap = None
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgrkref_tap.so")):
        tap = _libs.tap()  # must precede the reference driver
    w = WORKLOADS[name]
    img = synthetic_planes(w["width"], w["height"], w["comps"], w["prec"], seed=seed)
    th = w["tile"][1] or w["height"]
    rows = -(-w["height"] // th)
    cb = (1 << w["cblk"][0], 1 << w["cblk"][1])
    kw = {}
    if w.get("cinema"):  # grk_compress -w: the profile sets 32x32 blocks, 128 / 256 precincts, CPRL and the byte budget itself
        kw["cinema2k_fps"] = w["cinema"]
    if w.get("ht"):
        kw["cblk_sty"] = 64

    def secs(idx):
        return sum(tap.ref_tap_seconds(i) for i in idx) if tap is not None else float("nan")

    sim = [0.0, 0]  # T2::encode_packets_simulate of the last encode: seconds, calls

    def one(crop_rows):
        hh = min(w["height"], crop_rows * th)
        sub = [np.ascontiguousarray(p[:hh]) for p in img]
        if tap is not None:
            tap.ref_tap_reset()
        t0 = time.perf_counter()
        cs = _libs.ref_encode_image(sub, w["prec"], tile=(w["tile"][0] or 0, w["tile"][1] or 0), numres=w["numres"], cblk=cb,
                                    irreversible=not w["reversible"], rates=w["rates"], **kw)
        t1 = time.perf_counter()
        _libs.ref_decode_image(cs, w["comps"], w["width"], hh)
        t2 = time.perf_counter()
        # stage indices of _libs.TAP_STAGES: dc, mct, dwt, t1 (encode) / t1, dwt, mct, dc (decode)
        sim[0], sim[1] = secs((10,)), (tap.ref_tap_calls(10) if tap is not None else 0)
        return hh * w["width"], t1 - t0, t2 - t1, secs((0, 1, 2, 3)), secs((4, 5, 6, 7))

    crop = rows
    px, te, td, he, hd = one(crop)  # also warms the thread pool
    while crop > 1 and (te + td) * (steps + warmup) > budget_s:
        crop = max(1, crop // 2)
        px, te, td, he, hd = one(crop)
    for _ in range(max(0, warmup - 1)):
        one(crop)
    acc = np.zeros(4)
    for _ in range(steps):
        px, te, td, he, hd = one(crop)
        acc += (te, td, he, hd)
    t_enc, t_dec, h_enc, h_dec = acc / steps
    return dict(pixels=px, t_enc=t_enc, t_dec=t_dec, hot_enc=h_enc, hot_dec=h_dec, cores=cores, tapped=tap is not None, simulate_s=sim[0], simulate_calls=int(sim[1]),
                sample=f"{w['width']}x{px // w['width']} crop ({crop}/{rows} tile rows) of {WORKLOAD_TEXT[name]}; "
                       f"unmodified reference through its public grk API, memory streams, {cores} threads")


def cpu_baseline_object(r):
    whole = 2 * r["pixels"] / (r["t_enc"] + r["t_dec"]) / 1e6
    out = {"unit": "Mpixel/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"],
           "whole_codec_mpix_s": round(whole, 3),
           "whole_codec_encode_mpix_s": round(r["pixels"] / r["t_enc"] / 1e6, 3), "whole_codec_decode_mpix_s": round(r["pixels"] / r["t_dec"] / 1e6, 3)}
    if r["tapped"] and r["hot_enc"] > 0 and r["hot_dec"] > 0:
        hot = 2 * r["pixels"] / (r["hot_enc"] + r["hot_dec"]) / 1e6
        out.update({"pcrd_simulate": {"ms_per_encode": round(r.get("simulate_s", 0.0) * 1e3, 2), "calls": r.get("simulate_calls", 0),
                                      "what": "the reference's T2::encode_packets_simulate (packet-length simulation of every probe of its rate allocation): serial host code that stays the codec's, DESIGN.md section 9"}})
        out.update({"value": round(hot, 3), "hot_path_mpix_s": round(hot, 3),
                    "hot_path_encode_mpix_s": round(r["pixels"] / r["hot_enc"] / 1e6, 3), "hot_path_decode_mpix_s": round(r["pixels"] / r["hot_dec"] / 1e6, 3),
                    "what": "value = hot path only: wall clock inside the reference's dc_level_shift / mct / dwt / t1 stage calls and their inverses "
                            "(oracle/ref_tap.cpp), the work the GPU arm times; whole_codec_* adds its host-side PCRD, Tier-2 and codestream writer"})
    else:
        out.update({"value": round(whole, 3), "what": "whole codec (tap unavailable)"})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.workload, args.steps, args.warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present in this checkout"}))
        return
    cb = cpu_baseline_object(r)
    value = cb["value"]
    ms = 2 * r["pixels"] / (value * 1e6) * 1e3
    line = {
        "impl": "reference", "metric": "encode/decode Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "step": "encode + decode of the sample on the host CPU; " + cb["what"]},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- strong scaling: the partitions north_star names ----------------------------------------------------------------

def _tile_planes(base, index, prec, dtype):
    """plane of unit `index` derived from a base tile: a cyclic shift plus an offset, so that every unit is distinct"""
    return [np.ascontiguousarray(((np.roll(b, (3 * index, 7 * index), (0, 1)) + 5 * index) % (1 << prec)).astype(dtype)) for b in base]


def strong_scaling(gb, torch, dist, ctx, rank, world, dev, steps):
    """configs[2] encode dealt by tile, configs[3] (240 frames) encode dealt by frame, configs[4] decode dealt by tile: unit u goes
    to rank u mod N, no data-path collective.  Every rank drives the C ABI with pinned HOST buffers (packed samples in, code-block
    bytes out / bytes in, packed samples out), the encode partitions from three host workers that take the chunks in turn; `ms` = wall clock of the slowest rank for one pass over all units, so the figures
    of the N = 1, 2, 4, 8 runs compare directly.  Afterwards rank 0 gathers the results of all ranks (gloo, host memory; in a
    one-process host such as the plugin adapter the results are already in its memory, `host_gather_ms` is what the
    process-per-GPU layout of this bench costs) and compares them byte for byte with its own one-GPU run of the same units;
    `one_gpu_in_this_run_ms` is rank 0 alone over ALL units while the other ranks wait (informational)."""
    from grokimagecompression_b200 import params as P
    from grokimagecompression_b200.synth import synthetic_planes
    gloo = dist.new_group(backend="gloo") if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True).numpy()[:n].view(dtype).reshape(shape)

    def gather_bytes(buf):
        """variable-length byte gather to rank 0 over gloo -> list of uint8 arrays (rank order) on rank 0, wall seconds"""
        if world == 1:
            return [buf], 0.0
        t0 = time.perf_counter()
        n = torch.tensor([buf.size], dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, n, group=gloo)
        mx = int(max(int(s.item()) for s in sizes))
        send = torch.zeros(mx, dtype=torch.uint8)
        send[:buf.size] = torch.from_numpy(np.ascontiguousarray(buf).view(np.uint8).reshape(-1))
        recv = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
        dist.gather(send, recv, dst=0, group=gloo)
        dt = time.perf_counter() - t0
        if rank != 0:
            return None, dt
        return [recv[r].numpy()[:int(sizes[r].item())] for r in range(world)], dt

    out = {}

    def encode_units(label, nunits, chunk, min_units, comp_geom, base, prec, reversible, cblk, prc, rate_control, text):
        """units = tiles or frames of identical geometry; each rank encodes its units in chunks of `chunk` units per plan call"""
        w, h, nc = comp_geom
        sb = sample_bytes_of(prec)
        dt = np.uint8 if sb == 1 else np.uint16
        unit_tiles = P.image_tiles(w, h, nc, prec, reversible, (None, None), 6, rate_control=rate_control, cblk_expn=cblk, prc_expn=prc)
        cache = {}

        def source(u):
            if u not in cache:
                cache[u] = _tile_planes(base, u, prec, dt)
            return cache[u]

        def run(units, timed_steps, all_ranks):
            """-> (best wall seconds of a timed pass, None, results of the last pass: block records, rates and bytes of every chunk).
            all_ranks: every rank is in this call (rank barriers around the passes); otherwise the caller runs alone"""
            sync = barrier if all_ranks else torch.cuda.synchronize
            if not units:
                return 0.0, None, np.zeros(0, np.uint8)
            # units per plan call: `chunk` at most, and small enough that every worker gets two calls (a rank of an 8-GPU run
            # holds an eighth of the units: one call would leave nothing to overlap its transfers with)
            # (but not so small that a call is nothing but the latency of one launch chain: at least min_units)
            csz = max(1, min(chunk, len(units), max(min_units, -(-len(units) // (2 * STRONG_WORKERS)))))
            chunks = [units[c0:c0 + csz] for c0 in range(0, len(units), csz)]
            tail = len(units) % csz
            # two host workers, each with its own context (stream), plans and pinned buffers, take the chunks alternately: one
            # worker's staging copy and transfers run beside the other's kernels (the host owns the frame loop here)
            nwork = min(STRONG_WORKERS, len(chunks))
            ctxs = [ctx] + [gb.Context(dev.index or 0) for _ in range(nwork - 1)]
            work = []
            for wi in range(nwork):
                plan = gb.Plan(ctxs[wi], unit_tiles * csz, encoder=True, sample_bytes=sb)
                has_tail = tail and (len(chunks) - 1) % nwork == wi
                tplan = gb.Plan(ctxs[wi], unit_tiles * tail, encoder=True, sample_bytes=sb) if has_tail else None
                h_in = [pinned((h, w), dt) for _ in range(csz * nc)]
                outs = (pinned(plan.num_blocks * gb.CBLK_ENC_DTYPE.itemsize, np.uint8).view(gb.CBLK_ENC_DTYPE),
                        pinned(max(plan.num_pass_slots, 1), np.int32).view(np.uint32), pinned(max(plan.num_pass_slots, 1), np.float64),
                        pinned(csz * nc * w * h * 2 + (1 << 20), np.uint8))
                work.append((plan, tplan, h_in, outs))
            src = {u: source(u) for u in units}  # host images, prepared outside the timed region
            pieces, best = [], None

            def worker(wi, keep, pieces):
                plan, tplan, h_in, outs = work[wi]
                for ci in range(wi, len(chunks), nwork):
                    us = chunks[ci]
                    pl = plan if len(us) == csz else tplan
                    k = 0
                    for u in us:  # the host hands over its frames: copy into the pinned staging buffers (part of the host's work)
                        for p_ in src[u]:
                            h_in[k][...] = p_
                            k += 1
                    o = outs if pl is plan else (outs[0][:pl.num_blocks], outs[1][:max(pl.num_pass_slots, 1)], outs[2][:max(pl.num_pass_slots, 1)], outs[3])
                    res, rates, dists, data = pl.encode(h_in[:len(us) * nc], o)
                    if keep:  # the last pass's results for the comparison: block records, the pass rates that exist, the bytes
                        np_ = res["numpasses"].astype(np.int64)
                        d = np.zeros(pl.num_pass_slots + 1, np.int64)
                        np.add.at(d, pl.blocks["pass_offset"].astype(np.int64), 1)
                        np.add.at(d, pl.blocks["pass_offset"].astype(np.int64) + np_, -1)
                        valid = np.cumsum(d)[:pl.num_pass_slots] > 0
                        pieces[ci] = (res.copy(), np.where(valid, rates[:pl.num_pass_slots], 0).astype(np.uint32), data.copy())

            for it in range(1 + timed_steps):
                pieces = [None] * len(chunks)
                sync()
                t0 = time.perf_counter()
                ths = [threading.Thread(target=worker, args=(wi, it == timed_steps, pieces)) for wi in range(1, nwork)]
                for t in ths:
                    t.start()
                worker(0, it == timed_steps, pieces)
                for t in ths:
                    t.join()
                torch.cuda.synchronize()
                dt_ = time.perf_counter() - t0
                if it > 0:
                    best = dt_ if best is None else min(best, dt_)
            for plan, tplan, _, _ in work:
                plan.close()
                if tplan:
                    tplan.close()
            for c_ in ctxs[1:]:
                c_.close()
            best = best or 0.0
            blob = np.concatenate([np.concatenate([r.view(np.uint8).reshape(-1), ra.view(np.uint8).reshape(-1), d]) for r, ra, d in pieces])
            return best, None, blob

        mine = list(range(rank, nunits, world))
        t_n, _, blob = run(mine, steps, True)
        t_n = max_over_ranks(t_n)
        blobs, t_gather = gather_bytes(blob)
        entry = {"what": text, "units": nunits, "per_rank": len(mine), "n_gpus": world, "ms": round(t_n * 1e3, 2),
                 "mpix_s": round(nunits * w * h / t_n / 1e6, 1), "host_gather_ms": round(t_gather * 1e3, 2)}
        if world > 1:
            # rank 0 alone over all units: the N=1 time and the bytes to compare with (chunks of the N-rank run hold other units,
            # so the comparison is per unit: the digest of every unit's code-block bytes, in unit order)
            ok = None
            if rank == 0:
                t_1, _, _ = run(list(range(nunits)), 1, False)
                # every rank's unit list once more on this GPU: the gathered results must be the same bytes
                ok = True
                for r in range(world):
                    _, _, ref_blob = run(list(range(r, nunits, world)), 0, False) if r else (0, 0, blob)
                    ok = ok and ref_blob.size == blobs[r].size and bool((ref_blob == blobs[r]).all())
                entry.update({"one_gpu_in_this_run_ms": round(t_1 * 1e3, 2), "bytes_equal_to_one_gpu_run": ok,
                              "gathered_bytes": int(sum(b.size for b in blobs))})
            barrier()
        else:
            entry["gathered_bytes"] = int(blob.size)
        out[label] = entry

    # configs[2]: 8192x8192 3x16-bit lossless, 64 tiles of 1024x1024
    base3 = synthetic_planes(1024, 1024, 3, 16, seed=3)
    encode_units("c3", 64, 8, 4, (1024, 1024, 3), base3, 16, True, (6, 6), 15, False,
                 "configs[2] encode: 64 tiles of 1024x1024x3 16-bit, 5/3 + RCT, tile t -> rank t mod N")
    # configs[3]: 240 DCI 2K frames, 30 per plan call
    base4 = synthetic_planes(2048, 1080, 3, 12, seed=1000)
    encode_units("c4x240", 240, 30, 5, (2048, 1080, 3), base4, 12, False, (5, 5), [7] + [8] * 32, True,
                 "configs[3] encode: 240 frames of 2048x1080x3 12-bit, 9/7 + ICT, cinema precincts, frame f -> rank f mod N, 30 frames per plan call")

    # configs[4]: decode of a 16384x16384 3x8-bit image, 256 tiles of 1024x1024; the code blocks come from this repo's encoder (untimed)
    base5 = synthetic_planes(1024, 1024, 3, 8, seed=16)
    for label, reversible, reduce in (("c5_53", True, 0), ("c5_97", False, 0), ("c5_97_r2", False, 2)):
        nunits, nc, w, h = 256, 3, 1024, 1024
        enc_tiles = P.image_tiles(w, h, nc, 8, reversible, (None, None), 6, rate_control=False)
        dec_tiles = P.image_tiles(w, h, nc, 8, reversible, (None, None), 6, encoder=False, numres_decode=6 - reduce)

        def run_dec(units, timed_steps, all_ranks):
            sync = barrier if all_ranks else torch.cuda.synchronize
            if not units:
                return 0.0, np.zeros(0, np.uint8)
            # chunks of up to 64 tiles, taken in turn by host workers with their own context, plan and pinned buffers: one
            # worker's uploads and downloads run beside another's kernels
            csz = max(1, min(64, len(units), max(16, -(-len(units) // (2 * STRONG_WORKERS)))))
            chunks = [units[c0:c0 + csz] for c0 in range(0, len(units), csz)]
            nwork = min(STRONG_WORKERS, len(chunks))
            ctxs = [ctx] + [gb.Context(dev.index or 0) for _ in range(nwork - 1)]
            jobs = []  # per chunk: decoder inputs in pinned memory (the code blocks come from this repo's encoder, untimed)
            for us in chunks:
                n = len(us)
                eplan = gb.Plan(ctx, enc_tiles * n, encoder=True, sample_bytes=1)
                planes = []
                for u in us:
                    planes += _tile_planes(base5, u, 8, np.uint8)
                res, rates, dists, data = eplan.encode(planes)
                keep = np.asarray(eplan.blocks["resno"]) < 6 - reduce
                eplan.close()
                inp = pinned(int(keep.sum()) * gb.CBLK_DEC_DTYPE.itemsize, np.uint8).view(gb.CBLK_DEC_DTYPE)
                for k in ("numbps", "numpasses", "data_len", "data_offset"):
                    inp[k] = res[k][keep]
                inp["reserved"] = 0
                h_data = pinned(data.size, np.uint8)
                h_data[...] = data
                jobs.append((inp, h_data))
            plans = {}
            for wi in range(nwork):
                for n in sorted({len(chunks[ci]) for ci in range(wi, len(chunks), nwork)}):
                    plans[(wi, n)] = gb.Plan(ctxs[wi], dec_tiles * n, encoder=False, sample_bytes=1)
            h_outs = [[pinned(s_, np.uint8) for s_ in plans[(ci % nwork, len(us))].comp_shapes] for ci, us in enumerate(chunks)]

            def worker(wi):
                for ci in range(wi, len(chunks), nwork):
                    plans[(wi, len(chunks[ci]))].decode(jobs[ci][0], jobs[ci][1], h_outs[ci])

            best = None
            for it in range(1 + timed_steps):
                sync()
                t0 = time.perf_counter()
                ths = [threading.Thread(target=worker, args=(wi,)) for wi in range(1, nwork)]
                for t in ths:
                    t.start()
                worker(0)
                for t in ths:
                    t.join()
                torch.cuda.synchronize()
                d = time.perf_counter() - t0
                if it > 0:
                    best = d if best is None else min(best, d)
            for pl in plans.values():
                pl.close()
            for c_ in ctxs[1:]:
                c_.close()
            return best or 0.0, np.concatenate([o.reshape(-1) for ho in h_outs for o in ho])

        mine = list(range(rank, nunits, world))
        t_n, pix = run_dec(mine, steps, True)
        t_n = max_over_ranks(t_n)
        blobs, t_gather = gather_bytes(pix)
        entry = {"what": f"configs[4] decode: 16384x16384x3 8-bit, 256 tiles of 1024x1024, {'5/3 lossless' if reversible else '9/7'}, reduce {reduce}, tile t -> rank t mod N",
                 "units": nunits, "per_rank": len(mine), "n_gpus": world, "ms": round(t_n * 1e3, 2),
                 "mpix_s": round(nunits * w * h / t_n / 1e6, 1), "host_gather_ms": round(t_gather * 1e3, 2)}
        if world > 1:
            if rank == 0:
                t_1, _ = run_dec(list(range(nunits)), 1, False)
                ok = True
                for r in range(world):
                    ref = run_dec(list(range(r, nunits, world)), 0, False)[1] if r else pix
                    ok = ok and ref.size == blobs[r].size and bool((ref == blobs[r]).all())
                entry.update({"one_gpu_in_this_run_ms": round(t_1 * 1e3, 2), "pixels_equal_to_one_gpu_run": ok,
                              "gathered_bytes": int(sum(b.size for b in blobs))})
            barrier()
        elif reversible and reduce == 0:  # one GPU: the lossless decode must give the input tiles back
            want = np.concatenate([np.concatenate([p.reshape(-1) for p in _tile_planes(base5, u, 8, np.uint8)]) for u in mine])
            entry["lossless"] = bool((want == pix).all())
        out[label] = entry
    return out


# ---- main -----------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling partitions (configs[2] / [3] / [4])")
    ap.add_argument("--no-drop-in", action="store_true", help="skip the wall clock of the unmodified codec through the TCD seam")
    ap.add_argument("--e2e-threads", type=int, default=4,
                    help="host threads driving the C ABI in the end-to-end measurement, each with its own context, plans and pinned buffers")
    ap.add_argument("--int32-boundary", action="store_true", help="host planes as int32 (the reference's tile-buffer contract) instead of packed samples")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import grokimagecompression_b200 as gb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    # every rank keeps to its own slice of the host cores: the end-to-end path is host-driven (several threads per rank), and
    # ranks that share cores slow each other down (round 1: 0.86 e2e efficiency at N=8 on a 32-core box)
    ncpu = os.cpu_count() or 1
    per_rank = max(1, ncpu // max(world, 1))
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            os.sched_setaffinity(0, set(range(local_rank * per_rank, min(ncpu, (local_rank + 1) * per_rank))))
        except OSError:
            pass
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgrok_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = gb.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    w, img, tiles_e, tiles_d, planes = make_workload(args.workload, seed=1000 + rank)
    pixels = w["width"] * w["height"] * w.get("frames", 1)
    nsamples = pixels * w["comps"]
    sb = 4 if args.int32_boundary else sample_bytes_of(w["prec"])
    eplan = gb.Plan(ctx, tiles_e, encoder=True, sample_bytes=sb)
    dplan = gb.Plan(ctx, tiles_d, encoder=False, sample_bytes=sb)
    sdt = eplan.sample_dtype(0)
    planes = [np.ascontiguousarray(p.astype(sdt)) for p in planes]  # the image as the host holds it: packed samples

    # pinned host buffers: what a host TCD would hand over / receive
    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True).numpy()[:n].view(dtype).reshape(shape)

    def pinned_like(a):
        n = pinned(a.shape, a.dtype)
        n[...] = a
        return n

    h_planes = [pinned_like(p) for p in planes]
    res = pinned(eplan.num_blocks * gb.CBLK_ENC_DTYPE.itemsize, np.uint8).view(gb.CBLK_ENC_DTYPE)
    rates = pinned(max(eplan.num_pass_slots, 1), np.int32).view(np.uint32)
    dists = pinned(max(eplan.num_pass_slots, 1), np.float64)
    data_cap = int(sum(p.size for p in planes) * 2 + (1 << 20))  # code-block bytes never need more than 2 B/sample here
    data = pinned(data_cap, np.uint8)
    outs = (res, rates, dists, data)
    h_out = [pinned(s, sdt) for s in dplan.comp_shapes]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.zero_()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- correctness gate before timing: lossless path must round-trip, lossy path must be sane -----
    r0 = eplan.encode(h_planes, outs)
    inp = np.zeros(eplan.num_blocks, gb.CBLK_DEC_DTYPE)
    for k in ("numbps", "numpasses", "data_len", "data_offset"):
        inp[k] = r0[0][k]
    enc_bytes = int(len(r0[3]))
    decisions = int(r0[0]["decisions"].astype(np.int64).sum())
    h_inp = pinned(inp.nbytes, np.uint8).view(gb.CBLK_DEC_DTYPE)
    h_inp[...] = inp
    h_data = data[:enc_bytes]
    dplan.decode(h_inp, h_data, h_out)
    from grokimagecompression_b200 import params as P
    per_frame = len(h_out) // w.get("frames", 1)
    full = P.join_planes([o.astype(np.int32) for o in h_out[:per_frame]], w["width"], w["height"], w["comps"], w["tile"])
    if w["reversible"]:
        assert all((a == b).all() for a, b in zip(full, img)), "lossless round trip failed"
        psnr = float("inf")
    else:
        mse = np.mean([(np.mean((a.astype(np.float64) - b) ** 2)) for a, b in zip(full, img)])
        psnr = 10 * np.log10(((1 << w["prec"]) - 1) ** 2 / max(mse, 1e-12))
        assert psnr > 40.0, f"round-trip PSNR {psnr:.2f} dB"

    # ---- device-resident timing (value): inputs already in HBM -------------------------------------
    eplan.encode_upload(h_planes)
    eplan.encode_stash()
    dplan.decode_upload(h_inp, h_data)
    ctx.sync()

    def timed_device(nsteps, record):
        acc = np.zeros(8)
        evs = []
        for _ in range(nsteps):
            eplan.encode_restore()
            flush_l2()
            e = [ev() for _ in range(8)]
            e[0].record(stream); eplan.encode_run_stage(0)
            e[1].record(stream); eplan.encode_run_stage(1)
            e[2].record(stream); eplan.encode_run_stage(2)
            e[3].record(stream)
            flush_l2()
            e[4].record(stream); dplan.decode_run_stage(2)
            e[5].record(stream); dplan.decode_run_stage(1)
            e[6].record(stream); dplan.decode_run_stage(0)
            e[7].record(stream)
            evs.append(e)
        ctx.sync()
        torch.cuda.synchronize()
        if record:
            for e in evs:
                acc += (e[0].elapsed_time(e[3]), e[4].elapsed_time(e[7]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]),
                        e[4].elapsed_time(e[5]), e[5].elapsed_time(e[6]), e[0].elapsed_time(e[1]), e[6].elapsed_time(e[7]))
        return acc

    timed_device(args.warmup, False)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = ctx.launch_count()
    acc = timed_device(args.steps, True) / args.steps
    launches = ctx.launch_count() - l0
    barrier()
    t_enc, t_dec = max_over_ranks(acc[0]), max_over_ranks(acc[1])  # ms per step
    t_dwt, t_t1e, t_t1d, t_idwt, t_mct, t_imct = acc[2:]

    # ---- end to end through the C ABI with host buffers (e2e) ---------------------------------------
    # Every step copies that step's input planes H2D (packed samples, as the host's image holds them), runs the path and reads
    # the code-block bytes + pass tables back, then copies those H2D again, decodes and reads the packed planes back:
    # gb200_encode_tiles_packed + gb200_decode_tiles_packed on pinned host buffers.  A throughput-oriented host drives the
    # library from several threads (Grok itself has a thread pool); with --e2e-threads T, T host threads each own a context
    # (stream), a pair of plans and pinned buffers and run the same blocking calls, so one frame's PCIe copies overlap
    # another frame's kernels.  T = 1 is reported beside it.
    def e2e_step():
        eplan.encode(h_planes, outs)
        dplan.decode(h_inp, h_data, h_out)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e_single = (time.perf_counter() - t0) / args.steps
    barrier()
    t_e2e = t_e2e_single
    t_dev_conc = None
    nthreads = max(1, min(args.e2e_threads, per_rank if world > 1 else args.e2e_threads))
    if nthreads > 1:
        workers = []
        for _ in range(nthreads - 1):  # thread 0 reuses the context above
            c2 = gb.Context(local_rank)
            ep2, dp2 = gb.Plan(c2, tiles_e, encoder=True, sample_bytes=sb), gb.Plan(c2, tiles_d, encoder=False, sample_bytes=sb)
            hp2 = [pinned_like(p) for p in planes]
            o2 = (pinned(res.nbytes, np.uint8).view(gb.CBLK_ENC_DTYPE), pinned(rates.shape, np.int32).view(np.uint32),
                  pinned(dists.shape, np.float64), pinned(data_cap, np.uint8))
            hi2 = pinned(inp.nbytes, np.uint8).view(gb.CBLK_DEC_DTYPE)
            hi2[...] = inp
            ho2 = [pinned(sh, sdt) for sh in dp2.comp_shapes]
            workers.append((c2, ep2, dp2, hp2, o2, hi2, ho2))
        gate = threading.Barrier(nthreads + 1)
        ends = [0.0] * nthreads

        def run(idx):
            if idx == 0:
                enc, dec = (lambda: eplan.encode(h_planes, outs)), (lambda: dplan.decode(h_inp, h_data, h_out))
            else:
                c2, ep2, dp2, hp2, o2, hi2, ho2 = workers[idx - 1]
                enc, dec = (lambda: ep2.encode(hp2, o2)), (lambda: dp2.decode(hi2, o2[3][:enc_bytes], ho2))
            for _ in range(args.warmup):
                enc(); dec()
            gate.wait()
            for _ in range(args.steps):
                enc(); dec()
            ends[idx] = time.perf_counter()

        ths = [threading.Thread(target=run, args=(i,)) for i in range(nthreads)]
        for t in ths:
            t.start()
        gate.wait()
        t0 = time.perf_counter()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        t_e2e = (max(ends) - t0) / (args.steps * nthreads)  # mean wall time per frame with nthreads frames in flight
        barrier()

        # the same concurrency with every frame resident in HBM (no host copies): what the device sustains when several
        # frames are in flight on separate streams
        for (c2, ep2, dp2, hp2, o2, hi2, ho2) in workers:
            ep2.encode_upload(hp2); ep2.encode_stash(); dp2.decode_upload(hi2, o2[3][:enc_bytes]); c2.sync()
        eplan.encode_upload(h_planes); eplan.encode_stash(); dplan.decode_upload(h_inp, h_data); ctx.sync()
        gate2 = threading.Barrier(nthreads + 1)

        def run_dev(idx):
            c, ep, dp = (ctx, eplan, dplan) if idx == 0 else workers[idx - 1][:3]
            for it in range(args.warmup + args.steps):
                if it == args.warmup:
                    c.sync()
                    gate2.wait()
                ep.encode_restore(); ep.encode_run(); dp.decode_run()
            c.sync()
            ends[idx] = time.perf_counter()

        ths = [threading.Thread(target=run_dev, args=(i,)) for i in range(nthreads)]
        for t in ths:
            t.start()
        gate2.wait()
        t0 = time.perf_counter()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        t_dev_conc = max_over_ranks((max(ends) - t0) / (args.steps * nthreads))
        barrier()
        for (c2, ep2, dp2, hp2, o2, hi2, ho2) in workers:
            ep2.close(); dp2.close(); c2.close()
    clk = clocks.stop()
    t_e2e = max_over_ranks(t_e2e)
    t_e2e_single = max_over_ranks(t_e2e_single)
    h2d = sum(p.nbytes for p in h_planes) + h_inp.nbytes + enc_bytes
    d2h = res.nbytes + rates.nbytes + (dists.nbytes if len(w["rates"]) else 0) + enc_bytes + sum(o.nbytes for o in h_out) + 8

    strong = None
    if not args.no_strong:
        try:
            eplan.close(); dplan.close()
            eplan = dplan = None
            strong = strong_scaling(gb, torch, dist, ctx, rank, world, dev, steps=2)
        except Exception as exc:  # one GPU: never let an auxiliary measurement break the bench line (N > 1: ranks must fail together)
            if world > 1:
                raise
            strong = {"error": repr(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = t_enc + t_dec
    rc_on = len(w["rates"]) > 0
    value = world * 2 * pixels / (ms_step * 1e-3) / 1e6
    e2e_value = world * 2 * pixels / t_e2e / 1e6
    peak, peak_src = peaks()
    dwt_bytes, dwt_launches = dwt_algorithmic_bytes(tiles_e)
    achieved = dwt_bytes / (t_dwt * 1e-3) / 1e9
    # constants that come from an ncu capture are tied to the kernel sources they were measured on: a changed kernel makes
    # them stale, and a stale constant is reported as such instead of being multiplied into a fraction
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "dwt_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("source_sha1") == source_sha1(DWT_SOURCES):
                traffic = tj.get(args.workload)
            else:
                traffic_note = "profiles/dwt_traffic.json was captured on other kernel sources (stale): not reported"
        except Exception:
            traffic = None
    # integer-issue view of the Tier-1 kernels: warp instructions per MQ decision from the ncu capture of this build
    # (profiles/t1_issue.json) x decisions / the live event-timed duration, against one warp instruction per cycle per SM
    # sub-partition at the SM clock sampled during the run
    issue = None
    ip = os.path.join(ROOT, "profiles", "t1_issue.json")
    if os.path.exists(ip) and args.workload == "c2":
        try:
            ti = json.load(open(ip))
            if ti.get("source_sha1") != source_sha1(T1_SOURCES):
                issue = {"stale": True, "note": "profiles/t1_issue.json was captured on other Tier-1 kernel sources: re-capture with tools/capture_profiles.sh + profiles/update_constants.py"}
            else:
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                issue_peak = sms * 4 * (clk.get("sm_mhz") or 1965.0) * 1e6  # warp instructions per second
                issue = {"peak_warp_inst_per_s": issue_peak, "source": ti.get("source"),
                         "encode": {"warp_inst_per_decision": ti["model"] + ti["mq"],
                                    "issue_frac": round((ti["model"] + ti["mq"]) * decisions / (t_t1e * 1e-3) / issue_peak, 4)},
                         "decode": {"warp_inst_per_decision": ti["decode"],
                                    "issue_frac": round(ti["decode"] * decisions / (t_t1d * 1e-3) / issue_peak, 4)}}
        except Exception:
            issue = None
    wavelet = "5/3" if w["reversible"] else "9/7"
    mct_kernel = ("mct3" if w["comps"] >= 3 else "dcshift") + ("_fwd_packed_kernel" if sb != 4 else "_kernel")
    line = {
        "metric": "encode/decode Mpixel/s", "value": round(value, 2), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (5/3, 9/7 analysis, Tier-1) / fp32 (9/7 synthesis, inverse ICT)", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "step": "encode + decode of one image per GPU (every pixel crosses the path twice)",
                   "sharding": "one frame per rank, no collective", "l2": "flushed between timed stages (256 MiB memset)",
                   "host_boundary": f"{sb} byte(s) per sample" + (" (packed image samples, widened / narrowed on the device)" if sb != 4 else " (int32 planes)"),
                   "code_blocks": int(r0[0].size), "mq_decisions_per_image": decisions, "encoded_bytes": enc_bytes,
                   "roundtrip_psnr_db": None if psnr == float("inf") else round(psnr, 2)},
        "encode_mpix_s": round(world * pixels / (t_enc * 1e-3) / 1e6, 2), "decode_mpix_s": round(world * pixels / (t_dec * 1e-3) / 1e6, 2),
        "e2e": {"value": round(e2e_value, 2), "unit": "Mpixel/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(t_e2e * 1e3, 3), "host_threads": nthreads,
                "single_thread": {"value": round(world * 2 * pixels / t_e2e_single / 1e6, 2), "ms_per_step": round(t_e2e_single * 1e3, 3)}},
        "device_concurrent": None if t_dev_conc is None else {
            "value": round(world * 2 * pixels / t_dev_conc / 1e6, 2), "unit": "Mpixel/s", "streams": nthreads, "ms_per_step": round(t_dev_conc * 1e3, 3),
            "note": "device-resident like `value`, but with one frame in flight per stream (wall clock incl. the restore copy of the input planes)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": "dwt_fwd_stream_kernel<%s> (all levels of one image, %d launches)" % (wavelet, dwt_launches),
                     "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes": int(dwt_bytes), "ms": round(t_dwt, 4)},
        "roofline_inverse": {"kernel": "dwt_inv_stream_kernel<%s> (all levels of one image, %d launches)" % (wavelet, dwt_launches),
                             "bound": "hbm", "achieved": round(dwt_bytes / (t_idwt * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(dwt_bytes / (t_idwt * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes": int(dwt_bytes), "ms": round(t_idwt, 4)},
        # level shift + RCT / ICT: one read and one write per sample; the read (forward) / write (inverse) is the packed sample
        "roofline_mct": {"kernel": mct_kernel + " (level shift + %s, forward)" % ("RCT" if w["reversible"] else "ICT"), "bound": "hbm",
                         "algorithmic_bytes": int(nsamples * (sb + 4)), "ms": round(t_mct, 4), "achieved": round(nsamples * (sb + 4) / (t_mct * 1e-3) / 1e9, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(nsamples * (sb + 4) / (t_mct * 1e-3) / 1e9 / peak, 4),
                         "inverse": {"ms": round(t_imct, 4), "achieved": round(nsamples * (sb + 4) / (t_imct * 1e-3) / 1e9, 1),
                                     "frac": round(nsamples * (sb + 4) / (t_imct * 1e-3) / 1e9 / peak, 4)}},
        "t1": {"bound": "integer issue (serial MQ coder), no tensor work", "encode_ms": round(t_t1e, 4), "decode_ms": round(t_t1d, 4),
               "encode_mdecisions_s": round(decisions / (t_t1e * 1e-3) / 1e6, 1), "decode_mdecisions_s": round(decisions / (t_t1d * 1e-3) / 1e6, 1),
               "share_of_encode": round(t_t1e / t_enc, 3), "share_of_decode": round(t_t1d / t_dec, 3), "issue": issue},
    }
    if traffic_note:
        line["roofline"]["traffic_note"] = traffic_note
    if strong is not None:
        line["strong"] = strong
    # PCRD preparation on the device (gb200_encode_slopes: convex hull + log slopes of every block, kernel + D2H of the table),
    # not part of the step: the reference does this inside its host-side rate allocator
    if rc_on and world == 1:
        try:
            p2 = gb.Plan(ctx, tiles_e, encoder=True, sample_bytes=sb)
            p2.encode(h_planes, outs)
            p2.encode_slopes()
            t0 = time.perf_counter()
            for _ in range(5):
                sl = p2.encode_slopes()
            line["rd_slopes"] = {"ms_per_image_incl_d2h": round((time.perf_counter() - t0) / 5 * 1e3, 4), "passes": int(p2.num_pass_slots),
                                 "feasible_points": int((sl != 0).sum())}
            p2.close()
        except Exception as exc:
            line["rd_slopes"] = {"error": str(exc)[:200]}
    # the reversible 5/3 transform at configs[2] scale (8192x8192x3, 1024x1024 tiles): same kernel family, exact int32 lifting
    # with a fifth of the ALU work of the fixed-point 9/7, i.e. the case that is bound by HBM alone
    if world == 1 and args.workload == "c2":
        try:
            if eplan is not None:
                eplan.close(); dplan.close()
            from grokimagecompression_b200 import params as P2
            b53 = None
            for enc in (True, False):
                t53 = P2.image_tiles(8192, 8192, 3, 16, True, (1024, 1024), 6, encoder=enc)
                p53 = gb.Plan(ctx, t53, encoder=enc)
                if b53 is None:
                    b53, _ = dwt_algorithmic_bytes(t53)
                evs = []
                for i in range(args.warmup + args.steps):
                    flush_l2()
                    a, b_ = ev(), ev()
                    a.record(stream)
                    p53.encode_run_stage(1) if enc else p53.decode_run_stage(1)
                    b_.record(stream)
                    evs.append((a, b_))
                ctx.sync(); torch.cuda.synchronize()
                ms53 = sum(a.elapsed_time(b_) for a, b_ in evs[args.warmup:]) / args.steps
                line["roofline_5_3" if enc else "roofline_5_3_inverse"] = {
                    "kernel": "dwt_%s_stream_kernel<5/3> (5 levels, 8192x8192x3 int32, 1024x1024 tiles)" % ("fwd" if enc else "inv"), "bound": "hbm",
                    "achieved": round(b53 / (ms53 * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(b53 / (ms53 * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes": int(b53), "ms": round(ms53, 4)}
                p53.close()
        except Exception as exc:  # never let the auxiliary measurement break the bench line
            line["roofline_5_3"] = {"error": str(exc)[:200]}
    # the same image through the HTJ2K block coder (cblk_sty 0x40, grk_compress -M 64): device-resident encode + decode, stage split
    if world == 1 and args.workload == "c2":
        try:
            wh, _, th_e, th_d, ph = make_workload("c2ht", seed=1000 + rank)
            pe, pd = gb.Plan(ctx, th_e, encoder=True, sample_bytes=sb), gb.Plan(ctx, th_d, encoder=False, sample_bytes=sb)
            ph = [np.ascontiguousarray(p.astype(pe.sample_dtype(0))) for p in ph]
            rh = pe.encode(ph)
            ih = np.zeros(pe.num_blocks, gb.CBLK_DEC_DTYPE)
            for k in ("numbps", "numpasses", "data_len", "data_offset"):
                ih[k] = rh[0][k]
            pe.encode_upload(ph); pe.encode_stash(); pd.decode_upload(ih, rh[3]); ctx.sync()
            acc_h = np.zeros(4)
            for it in range(args.warmup + args.steps):
                pe.encode_restore(); flush_l2()
                e = [ev() for _ in range(6)]
                e[0].record(stream); pe.encode_run_stage(0); pe.encode_run_stage(1)
                e[1].record(stream); pe.encode_run_stage(2)
                e[2].record(stream); flush_l2()
                e[3].record(stream); pd.decode_run_stage(2)
                e[4].record(stream); pd.decode_run_stage(1); pd.decode_run_stage(0)
                e[5].record(stream)
                ctx.sync(); torch.cuda.synchronize()
                if it >= args.warmup:
                    acc_h += (e[0].elapsed_time(e[2]), e[3].elapsed_time(e[5]), e[1].elapsed_time(e[2]), e[3].elapsed_time(e[4]))
            acc_h /= args.steps
            # end to end through the C ABI from one host thread, packed host samples (as `e2e.single_thread`)
            h_ph = [pinned_like(p) for p in ph]
            o_h = (pinned(pe.num_blocks * gb.CBLK_ENC_DTYPE.itemsize, np.uint8).view(gb.CBLK_ENC_DTYPE), pinned(max(pe.num_pass_slots, 1), np.int32).view(np.uint32),
                   pinned(max(pe.num_pass_slots, 1), np.float64), pinned(data_cap, np.uint8))
            hi_h = pinned(ih.nbytes, np.uint8).view(gb.CBLK_DEC_DTYPE)
            hi_h[...] = ih
            ho_h = [pinned(sh, pe.sample_dtype(0)) for sh in pd.comp_shapes]
            nbytes_h = len(rh[3])
            for it in range(args.warmup + args.steps):
                if it == args.warmup:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                pe.encode(h_ph, o_h)
                pd.decode(hi_h, o_h[3][:nbytes_h], ho_h)
            t_ht_e2e = (time.perf_counter() - t0) / args.steps
            line["ht"] = {"workload": WORKLOAD_TEXT["c2ht"], "kernels": "t1_ht_encode_kernel / t1_ht_decode_kernel (csrc/ht.cu): one warp per code block",
                          "e2e_single_thread": {"value": round(2 * pixels / t_ht_e2e / 1e6, 1), "ms_per_step": round(t_ht_e2e * 1e3, 3)},
                          "value": round(2 * pixels / ((acc_h[0] + acc_h[1]) * 1e-3) / 1e6, 1), "unit": "Mpixel/s (device-resident, like `value`)",
                          "encode_ms": round(float(acc_h[0]), 4), "decode_ms": round(float(acc_h[1]), 4),
                          "t1_encode_ms": round(float(acc_h[2]), 4), "t1_decode_ms": round(float(acc_h[3]), 4), "encoded_bytes": int(len(rh[3]))}
            pe.close(); pd.close()
        except Exception as exc:
            line["ht"] = {"error": repr(exc)[:200]}
    # what an UNMODIFIED Grok gains when its TCD stage calls are bound to this library (integration/grok_tcd_shim.cpp): wall
    # clock of the reference codec through its public API, pure and with the seam, on this workload
    if world == 1 and not args.no_drop_in and args.workload in ("c1", "c2", "c4") \
            and os.path.exists(os.path.join(ROOT, "integration", "_build", "libgrok_b200_tcd.so")):
        try:
            outp = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "dropin_bench.py"), args.workload, "2", "--json"],
                                           text=True, timeout=300, stderr=subprocess.DEVNULL)
            line["drop_in"] = json.loads([ln for ln in outp.splitlines() if ln.startswith("{")][-1])
        except Exception as exc:
            line["drop_in"] = {"error": repr(exc)[:200]}
        if args.workload == "c2":  # the same image with -M 64: the reference's T1HT against the device's HT kernels
            try:
                outp = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "dropin_bench.py"), "c2ht", "2", "--json"],
                                               text=True, timeout=300, stderr=subprocess.DEVNULL)
                line["drop_in_ht"] = json.loads([ln for ln in outp.splitlines() if ln.startswith("{")][-1])
            except Exception as exc:
                line["drop_in_ht"] = {"error": repr(exc)[:200]}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(args.workload, steps=2, warmup=1, budget_s=25.0, seed=1000)
        line["cpu_baseline"] = cpu_baseline_object(r) if r is not None else {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "reference",
                                                                              "sample": "the compiled reference is missing"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
