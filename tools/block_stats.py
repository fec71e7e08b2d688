import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, grokimagecompression_b200 as gb
w, img, te, td, planes = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2", 1000)
ctx = gb.Context(0); plan = gb.Plan(ctx, te, True)
res, rates, dists, data = plan.encode(planes)
d = res["decisions"].astype(np.int64)
print("blocks", len(d), "total", d.sum(), "mean", d.mean(), "max", d.max(), "p99", np.percentile(d, 99), "p90", np.percentile(d, 90))
for r in range(6):
    m = plan.blocks["resno"] == r
    if m.any(): print("res", r, "n", m.sum(), "mean", d[m].mean(), "max", d[m].max(), "numbps max", res["numbps"][m].max(), "w", (plan.blocks["x1"][m]-plan.blocks["x0"][m]).max())
# per-warp imbalance: blocks in groups of 32
g = [d[i:i+32] for i in range(0, len(d), 32)]
print("sum of per-warp max", sum(x.max() for x in g), "vs total/32", d.sum() / 32, "global max", d.max())
