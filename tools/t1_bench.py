#!/usr/bin/env python
"""Times only the Tier-1 stages of one workload (default configs[1]) on cuda:0: tools/t1_bench.py [workload] [iters]
The library is taken from $GB200_LIB when set (see tools/ab.sh)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import grokimagecompression_b200 as gb
import bench as B

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = gb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
w, img, tiles_e, tiles_d, planes = B.make_workload(name, seed=1000)
eplan, dplan = gb.Plan(ctx, tiles_e, encoder=True), gb.Plan(ctx, tiles_d, encoder=False)
res, rates, dists, data = eplan.encode(planes)
inp = np.zeros(eplan.num_blocks, gb.CBLK_DEC_DTYPE)
for k in ("numbps", "numpasses", "data_len", "data_offset"):
    inp[k] = res[k]
dec = int(res["decisions"].astype(np.int64).sum())
eplan.encode_upload(planes); eplan.encode_stash(); dplan.decode_upload(inp, data); ctx.sync()
te, td = [], []
for i in range(iters + 2):
    eplan.encode_restore(); eplan.encode_run_stage(0); eplan.encode_run_stage(1)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(stream); eplan.encode_run_stage(2); e[1].record(stream)
    e[2].record(stream); dplan.decode_run_stage(2); e[3].record(stream)
    ctx.sync(); torch.cuda.synchronize()
    if i >= 2:
        te.append(e[0].elapsed_time(e[1])); td.append(e[2].elapsed_time(e[3]))
print(f"{os.environ.get('GB200_LIB', 'default')}: {name} blocks {eplan.num_blocks} decisions {dec}  T1 encode {min(te):.3f} ms  T1 decode {min(td):.3f} ms")
