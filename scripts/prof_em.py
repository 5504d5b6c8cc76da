"""Profiling driver (run under ncu): EM loop on config 2.
usage: prof_em.py iters variant cfg keep_cm keep_tm [C M]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq
iters = int(sys.argv[1]); variant = int(sys.argv[2]); cfg = int(sys.argv[3])
k1 = int(sys.argv[4]) if len(sys.argv) > 4 else 100
k2 = int(sys.argv[5]) if len(sys.argv) > 5 else 50
C = int(sys.argv[6]) if len(sys.argv) > 6 else 500000
M = int(sys.argv[7]) if len(sys.argv) > 7 else 250000
eq, proj, eff, uniq = synth_eq(seed=1, C=C, M=M, total_count=40 * C)
ctx = EMContext(0)
ctx.set_option("variant", variant); ctx.set_option("config", cfg)
ctx.set_option("l2_keep_cm", k1); ctx.set_option("l2_keep_tm", k2)
p = default_params(min_iter=iters, max_iter=iters)
ctx.upload(eq, proj, eff, uniq); ctx.prepare(p)
r = ctx.run()
r = ctx.run()
print("iters", r.iters, "loop us/iter", r.loop_kernel_ms / iters * 1e3, "nnz", eq.nnz)
