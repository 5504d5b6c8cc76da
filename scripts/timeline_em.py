"""Per-warp phase timeline of one iteration of the persistent EM kernel (dev helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq
C = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
eq, proj, eff, uniq = synth_eq(seed=1, C=C, M=C // 2, total_count=40 * C)
ctx = EMContext(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
p = default_params(min_iter=60, max_iter=60)
ctx.upload(eq, proj, eff, uniq); ctx.prepare(p)
ctx.run()
nw = ctx.arm_timeline(40)
r = ctx.run()
t = ctx.read_timeline(nw).astype(np.int64)
t0 = t[:, 0].min()
names = ["P1 start", "P1 end", "bar1 end", "P2 start", "P2 end", "reduce end", "bar2 end"]
sizes = (eq.off[1:] - eq.off[:-1]).astype(np.int64)
tc = np.bincount(eq.tids[np.repeat(sizes > 1, sizes)], minlength=eq.n_txps)
print("classes >32:", int((sizes > 32).sum()), " txps >32:", int((tc > 32).sum()), ">256:", int((tc > 256).sum()), ">2048:", int((tc > 2048).sum()), "max", int(tc.max()), " entries in txp rows >32:", int(tc[tc > 32].sum()))
print("C", C, "loop us/iter", r.loop_kernel_ms / 60 * 1e3, "warps", nw)
for i, n in enumerate(names):
    col = t[:, i] - t0
    print(f"{n:11s} min {col.min()/1e3:8.2f}  p50 {np.median(col)/1e3:8.2f}  p90 {np.percentile(col,90)/1e3:8.2f}  max {col.max()/1e3:8.2f} us")
d1 = (t[:, 1] - t[:, 0]) / 1e3; d2 = (t[:, 4] - t[:, 3]) / 1e3
print("P1 duration per warp: p50 %.2f p90 %.2f p99 %.2f max %.2f us" % (np.median(d1), np.percentile(d1, 90), np.percentile(d1, 99), d1.max()))
print("P2 duration per warp: p50 %.2f p90 %.2f p99 %.2f max %.2f us" % (np.median(d2), np.percentile(d2, 90), np.percentile(d2, 99), d2.max()))
d3 = (t[:, 7] - t[:, 3]) / 1e3
print("P2 SELL part per warp: p50 %.2f p90 %.2f max %.2f us ; long-row part p50 %.2f max %.2f" % (np.median(d3), np.percentile(d3, 90), d3.max(), np.median(d2 - d3), (d2 - d3).max()))
w1 = np.argsort(-d1)[:5]; w2 = np.argsort(-d2)[:5]
print("slowest P1 warps", w1, d1[w1]); print("slowest P2 warps", w2, d2[w2])
