"""Profiling driver (run under ncu): EM loop on config 2 with the chosen variant/config."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq
iters = int(sys.argv[1]); variant = int(sys.argv[2]); cfg = int(sys.argv[3])
eq, proj, eff, uniq = synth_eq(seed=1)
ctx = EMContext(0)
ctx.set_option("variant", variant)
p = default_params(min_iter=iters, max_iter=iters)
ctx.upload(eq, proj, eff, uniq); ctx.prepare(p)
r = ctx.run()
print("iters", r.iters, "loop us/iter", r.loop_kernel_ms / iters * 1e3)
