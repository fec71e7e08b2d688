#!/usr/bin/env python
"""bench.py -- encode/decode throughput of the tile-coding hot path on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (rank 0 only)

One step = one pass of the hot path over one batch of synthetic input: the workload image is encoded
(DC shift + ICT/RCT + DWT + quantise + Tier-1) and the resulting code blocks are decoded again
(Tier-1 + de-quantise + inverse DWT + inverse MCT + clamp).  Every pixel therefore crosses the path
twice per step and the headline metric is  2 * pixels / step time.

Workload at N=1: BASELINE.json configs[1] -- 4096x2160 RGB 8-bit, irreversible 9/7 + ICT, quantised,
rate control on (per-pass distortion for 4 quality layers), 1024x1024 tiles.  Multi-GPU: the path
shards by tile/frame with no collective, every rank codes its own frame of the same shape (weak
scaling) and rank 0 reports the aggregate over the max-over-ranks time.
"""
import argparse
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
    "c4": dict(width=2048, height=1080, comps=3, prec=12, reversible=False, tile=(None, None), numres=6, cblk=(5, 5), rates=(10,)),
    # 30 of the 240 frames of configs[3]: what one GPU of eight gets when the batch is sharded by frame; one plan, one launch per stage
    "c4x30": dict(width=2048, height=1080, comps=3, prec=12, reversible=False, tile=(None, None), numres=6, cblk=(5, 5), rates=(10,), frames=30),
}
WORKLOAD_TEXT = {
    "c1": "configs[0]: 2048x2048 8-bit gray, lossless 5/3, 1 tile, 64x64 blocks, 5 levels",
    "c2": "configs[1]: 4096x2160 RGB 8-bit, irreversible 9/7 + ICT, quantised, 4 quality layers, 1024x1024 tiles",
    "c3": "configs[2]: 8192x8192 3x16-bit, lossless 5/3 + RCT, 1024x1024 tiles",
    "c4": "configs[3]: one DCI 2K frame 2048x1080 3x12-bit, 9/7 + ICT, 32x32 blocks",
    "c4x30": "configs[3]: batch of 30 DCI 2K frames 2048x1080 3x12-bit (240 frames over 8 GPUs), 9/7 + ICT, 32x32 blocks",
}


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


def make_workload(name, seed):
    from grokimagecompression_b200 import params as P
    from grokimagecompression_b200.synth import synthetic_planes
    w = WORKLOADS[name]
    rc = len(w["rates"]) > 0
    img, tiles_e, tiles_d, planes = None, [], [], []
    for f in range(w.get("frames", 1)):  # frames of a batch are independent images: more tiles of the same plan
        fimg = synthetic_planes(w["width"], w["height"], w["comps"], w["prec"], seed=seed + f)
        img = fimg if img is None else img  # the correctness gate looks at the first frame
        tiles_e += P.image_tiles(w["width"], w["height"], w["comps"], w["prec"], w["reversible"], w["tile"], w["numres"],
                                 rate_control=rc, cblk_expn=w["cblk"])
        tiles_d += P.image_tiles(w["width"], w["height"], w["comps"], w["prec"], w["reversible"], w["tile"], w["numres"],
                                 cblk_expn=w["cblk"], encoder=False)
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


def cpu_reference_run(name, steps, warmup, budget_s=150.0, seed=1):
    """Times the unmodified reference (oracle/_ref) on the host cores: encode + decode through its public
    API with memory streams.  The sample is the full image unless that would blow the time budget, in which
    case a crop of whole tile rows is used."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _libs
    from grokimagecompression_b200.synth import synthetic_planes
    if not _libs.have_ref():
        return None
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    os.environ["GRK_REF_THREADS"] = str(cores)
    img = synthetic_planes(w["width"], w["height"], w["comps"], w["prec"], seed=seed)
    th = w["tile"][1] or w["height"]
    rows = -(-w["height"] // th)
    cb = (1 << w["cblk"][0], 1 << w["cblk"][1])

    def one(crop_rows):
        hh = min(w["height"], crop_rows * th)
        sub = [np.ascontiguousarray(p[:hh]) for p in img]
        t0 = time.perf_counter()
        cs = _libs.ref_encode_image(sub, w["prec"], tile=(w["tile"][0] or 0, w["tile"][1] or 0), numres=w["numres"], cblk=cb,
                                    irreversible=not w["reversible"], rates=w["rates"])
        t1 = time.perf_counter()
        _libs.ref_decode_image(cs, w["comps"], w["width"], hh)
        t2 = time.perf_counter()
        return hh * w["width"], t1 - t0, t2 - t1

    crop = rows
    px, te, td = one(crop)  # also warms the thread pool
    while crop > 1 and (te + td) * (steps + warmup) > budget_s:
        crop = max(1, crop // 2)
        px, te, td = one(crop)
    for _ in range(max(0, warmup - 1)):
        one(crop)
    tes, tds = [], []
    for _ in range(steps):
        px, te, td = one(crop)
        tes.append(te)
        tds.append(td)
    t_enc, t_dec = sum(tes) / steps, sum(tds) / steps
    return dict(pixels=px, t_enc=t_enc, t_dec=t_dec, cores=cores,
                sample=f"{w['width']}x{px // w['width']} crop ({crop}/{rows} tile rows) of {WORKLOAD_TEXT[name]}; "
                       f"reference grk public API, memory streams, includes its host-side PCRD/T2/codestream")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.workload, args.steps, args.warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present in this checkout"}))
        return
    ms = (r["t_enc"] + r["t_dec"]) * 1e3
    value = 2 * r["pixels"] / (r["t_enc"] + r["t_dec"]) / 1e6
    line = {
        "impl": "reference", "metric": "encode/decode Mpixel/s", "value": round(value, 3), "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32/fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "step": "encode + decode of the sample on the host CPU"},
        "encode_mpix_s": round(r["pixels"] / r["t_enc"] / 1e6, 3), "decode_mpix_s": round(r["pixels"] / r["t_dec"] / 1e6, 3),
        "cpu_baseline": {"value": round(value, 3), "unit": "Mpixel/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
        "e2e": {"value": round(value, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-threads", type=int, default=4,
                    help="host threads driving the C ABI in the end-to-end measurement, each with its own context, plans and pinned buffers")
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
    eplan = gb.Plan(ctx, tiles_e, encoder=True)
    dplan = gb.Plan(ctx, tiles_d, encoder=False)

    # pinned host buffers: what a host TCD would hand over / receive
    def pinned_like(a):
        t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)), pin_memory=True)
        n = t.numpy()
        n[...] = a
        return n

    h_planes = [pinned_like(p) for p in planes]
    res = torch.empty(eplan.num_blocks * gb.CBLK_ENC_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(gb.CBLK_ENC_DTYPE)
    rates = torch.empty(max(eplan.num_pass_slots, 1), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    dists = torch.empty(max(eplan.num_pass_slots, 1), dtype=torch.float64, pin_memory=True).numpy()
    data_cap = int(sum(p.size for p in planes) * 2 + (1 << 20))  # int32 samples never need more than 2 B/sample here
    data = torch.empty(data_cap, dtype=torch.uint8, pin_memory=True).numpy()
    outs = (res, rates, dists, data)
    h_out = [torch.empty(s, dtype=torch.int32, pin_memory=True).numpy() for s in dplan.comp_shapes]

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
    h_inp = torch.empty(inp.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(gb.CBLK_DEC_DTYPE)
    h_inp[...] = inp
    h_data = data[:enc_bytes]
    dplan.decode(h_inp, h_data, h_out)
    from grokimagecompression_b200 import params as P
    per_frame = len(h_out) // w.get("frames", 1)
    full = P.join_planes(h_out[:per_frame], w["width"], w["height"], w["comps"], w["tile"])
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
        t_enc = t_dec = t_dwt = t_t1e = t_t1d = t_idwt = 0.0
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
                t_enc += e[0].elapsed_time(e[3]); t_dec += e[4].elapsed_time(e[7])
                t_dwt += e[1].elapsed_time(e[2]); t_t1e += e[2].elapsed_time(e[3]); t_t1d += e[4].elapsed_time(e[5])
                t_idwt += e[5].elapsed_time(e[6])
        return t_enc, t_dec, t_dwt, t_t1e, t_t1d, t_idwt

    timed_device(args.warmup, False)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = ctx.launch_count()
    t_enc, t_dec, t_dwt, t_t1e, t_t1d, t_idwt = timed_device(args.steps, True)
    launches = ctx.launch_count() - l0
    barrier()
    t_enc, t_dec = max_over_ranks(t_enc / args.steps), max_over_ranks(t_dec / args.steps)  # ms per step
    t_dwt, t_t1e, t_t1d, t_idwt = t_dwt / args.steps, t_t1e / args.steps, t_t1d / args.steps, t_idwt / args.steps

    # ---- end to end through the C ABI with host buffers (e2e) ---------------------------------------
    # Every step copies that step's input planes H2D, runs the path and reads the code-block bytes + pass tables back, then
    # copies those H2D again, decodes and reads the planes back: gb200_encode_tiles + gb200_decode_tiles on pinned host
    # buffers.  A throughput-oriented host drives the library from several threads (Grok itself has a thread pool); with
    # --e2e-threads T, T host threads each own a context (stream), a pair of plans and pinned buffers and run the same
    # blocking calls, so one frame's PCIe copies overlap another frame's kernels.  T = 1 is reported beside it.
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
    nthreads = max(1, args.e2e_threads)
    if nthreads > 1:
        workers = []
        for _ in range(nthreads - 1):  # thread 0 reuses the context above
            c2 = gb.Context(local_rank)
            ep2, dp2 = gb.Plan(c2, tiles_e, encoder=True), gb.Plan(c2, tiles_d, encoder=False)
            hp2 = [pinned_like(p) for p in planes]
            o2 = (torch.empty(res.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(gb.CBLK_ENC_DTYPE),
                  torch.empty(rates.shape, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
                  torch.empty(dists.shape, dtype=torch.float64, pin_memory=True).numpy(),
                  torch.empty(data_cap, dtype=torch.uint8, pin_memory=True).numpy())
            hi2 = torch.empty(inp.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(gb.CBLK_DEC_DTYPE)
            hi2[...] = inp
            ho2 = [torch.empty(sh, dtype=torch.int32, pin_memory=True).numpy() for sh in dp2.comp_shapes]
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
    clk = clocks.stop()
    t_e2e = max_over_ranks(t_e2e)
    t_e2e_single = max_over_ranks(t_e2e_single)
    h2d = sum(p.nbytes for p in h_planes) + h_inp.nbytes + enc_bytes
    d2h = res.nbytes + rates.nbytes + dists.nbytes + enc_bytes + sum(o.nbytes for o in h_out)

    if rank != 0:
        return
    ms_step = t_enc + t_dec
    rc_on = len(w["rates"]) > 0
    value = world * 2 * pixels / (ms_step * 1e-3) / 1e6
    e2e_value = world * 2 * pixels / t_e2e / 1e6
    peak, peak_src = peaks()
    dwt_bytes, dwt_launches = dwt_algorithmic_bytes(tiles_e)
    achieved = dwt_bytes / (t_dwt * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dwt_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
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
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            issue_peak = sms * 4 * (clk.get("sm_mhz") or 1965.0) * 1e6  # warp instructions per second
            issue = {"peak_warp_inst_per_s": issue_peak, "source": ti.get("source"),
                     "encode": {"warp_inst_per_decision": ti["model"] + ti["mq"],
                                "issue_frac": round((ti["model"] + ti["mq"]) * decisions / (t_t1e * 1e-3) / issue_peak, 4)},
                     "decode": {"warp_inst_per_decision": ti["decode"],
                                "issue_frac": round(ti["decode"] * decisions / (t_t1d * 1e-3) / issue_peak, 4)}}
        except Exception:
            issue = None
    line = {
        "metric": "encode/decode Mpixel/s", "value": round(value, 2), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (5/3, 9/7 analysis, Tier-1) / fp32 (9/7 synthesis, inverse ICT)", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "step": "encode + decode of one image per GPU (every pixel crosses the path twice)",
                   "sharding": "one frame per rank, no collective", "l2": "flushed between timed stages (256 MiB memset)",
                   "code_blocks": eplan.num_blocks, "mq_decisions_per_image": decisions, "encoded_bytes": enc_bytes,
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
        "roofline": {"kernel": "dwt_fwd_stream_kernel<%s> (all levels of one image, %d launches)" % ("5/3" if w["reversible"] else "9/7", dwt_launches),
                     "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes": int(dwt_bytes), "ms": round(t_dwt, 4)},
        "roofline_inverse": {"kernel": "dwt_inv_stream_kernel<%s> (all levels of one image, %d launches)" % ("5/3" if w["reversible"] else "9/7", dwt_launches),
                             "bound": "hbm", "achieved": round(dwt_bytes / (t_idwt * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(dwt_bytes / (t_idwt * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes": int(dwt_bytes), "ms": round(t_idwt, 4)},
        "t1": {"bound": "integer issue (serial MQ coder), no tensor work", "encode_ms": round(t_t1e, 4), "decode_ms": round(t_t1d, 4),
               "encode_mdecisions_s": round(decisions / (t_t1e * 1e-3) / 1e6, 1), "decode_mdecisions_s": round(decisions / (t_t1d * 1e-3) / 1e6, 1),
               "share_of_encode": round(t_t1e / t_enc, 3), "share_of_decode": round(t_t1d / t_dec, 3), "issue": issue},
    }
    # PCRD preparation on the device (gb200_encode_slopes: convex hull + log slopes of every block, kernel + D2H of the table),
    # not part of the step: the reference does this inside its host-side rate allocator
    if rc_on and world == 1:
        try:
            eplan.encode_restore(); eplan.encode_run(); ctx.sync()
            eplan.encode_slopes()
            t0 = time.perf_counter()
            for _ in range(5):
                sl = eplan.encode_slopes()
            line["rd_slopes"] = {"ms_per_image_incl_d2h": round((time.perf_counter() - t0) / 5 * 1e3, 4), "passes": int(eplan.num_pass_slots),
                                 "feasible_points": int((sl != 0).sum())}
        except Exception as exc:
            line["rd_slopes"] = {"error": str(exc)[:200]}
    # the reversible 5/3 transform at configs[2] scale (8192x8192x3, 1024x1024 tiles): same kernel family, exact int32 lifting
    # with a fifth of the ALU work of the fixed-point 9/7, i.e. the case that is bound by HBM alone
    if world == 1 and args.workload == "c2":
        try:
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
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(args.workload, steps=2, warmup=1, budget_s=25.0, seed=1000)
        if r is not None:
            v = 2 * r["pixels"] / (r["t_enc"] + r["t_dec"]) / 1e6
            line["cpu_baseline"] = {"value": round(v, 3), "unit": "Mpixel/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                                    "encode_mpix_s": round(r["pixels"] / r["t_enc"] / 1e6, 3), "decode_mpix_s": round(r["pixels"] / r["t_dec"] / 1e6, 3)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
