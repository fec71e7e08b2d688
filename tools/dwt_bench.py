"""DWT-only sweep on one GPU: forward and inverse transform of one image's planes, event-timed on the library's stream with
L2 flushed between runs, for the first-generation (shared-memory) kernels and the streaming kernels with their tuning knobs
(rows per work item, prefetch depth, fill threshold).  Prints one line per variant: ms, algorithmic GB/s, fraction of the
measured HBM peak.

    python tools/dwt_bench.py [c2|c3|c1|c4|c4x30] [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
import grokimagecompression_b200 as gb
from grokimagecompression_b200 import params as P


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    warm = int(os.environ.get("DWT_BENCH_WARMUP", "3"))
    variants = sys.argv[3:]
    w = bench.WORKLOADS[name]
    frames = w.get("frames", 1)
    tiles_e, tiles_d = [], []
    for _ in range(frames):
        tiles_e += P.image_tiles(w["width"], w["height"], w["comps"], w["prec"], w["reversible"], w["tile"], w["numres"], cblk_expn=w["cblk"])
        tiles_d += P.image_tiles(w["width"], w["height"], w["comps"], w["prec"], w["reversible"], w["tile"], w["numres"], cblk_expn=w["cblk"],
                                 encoder=False)
    nbytes, launches = bench.dwt_algorithmic_bytes(tiles_e)
    peak, _ = bench.peaks()
    ctx = gb.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if not variants:
        variants = ["rows=0,unroll=2", "rows=0,unroll=1", "rows=0,unroll=4", "rows=32,unroll=2", "rows=64,unroll=2", "rows=128,unroll=2",
                    "rows=256,unroll=2", "rows=0,unroll=2,fill=8", "rows=0,unroll=2,fill=32"]
    print(f"workload {name}: {nbytes / 1e6:.1f} MB algorithmic per direction, {launches} level launches, peak {peak} GB/s")
    out = []
    for v in variants:
        for k in ("GB200_DWT_ROWS", "GB200_DWT_UNROLL", "GB200_DWT_FILL", "GB200_DWT_HL", "GB200_DWT_ONLY", "GB200_DWT_RING", "GB200_DWT_MINROWS"):
            os.environ.pop(k, None)
        for kv in v.split(","):
            k, x = kv.split("=")
            os.environ["GB200_DWT_" + k.upper()] = x
        res = {"variant": v}
        for enc, tiles in ((True, tiles_e), (False, tiles_d)):
            plan = gb.Plan(ctx, tiles, encoder=enc)
            run = (lambda: plan.encode_run_stage(1)) if enc else (lambda: plan.decode_run_stage(1))
            evs = []
            for i in range(warm + steps):
                with torch.cuda.stream(stream):
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); run(); b.record(stream)
                evs.append((a, b))
            ctx.sync(); torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in evs[warm:])
            ms = sum(ts) / len(ts)
            res["fwd" if enc else "inv"] = {"ms": round(ms, 4), "min_ms": round(ts[0], 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 3)}
            plan.close()
        print(json.dumps(res), flush=True)
        out.append(res)
    return out


if __name__ == "__main__":
    main()
