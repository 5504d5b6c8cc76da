"""Quick device timing of the EM loop on BASELINE config 2 (dev helper, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eq, proj, eff, uniq = synth_eq(seed=1)
ctx = EMContext(0)
p = default_params(min_iter=iters, max_iter=iters)
t = time.time(); ctx.upload(eq, proj, eff, uniq); print("upload s", time.time() - t)
for variant in (1, 0):
    for bps in (0, 1, 2, 3):
        ctx.set_option("variant", variant)
        ctx.set_option("blocks_per_sm", bps)
        st = ctx.prepare(p)
        for rep in range(2):
            r = ctx.run()
        print(f"variant {variant} blocks/sm {bps}: prepare {st.prepare_ms:.2f} ms, run {r.run_ms:.2f} ms, "
              f"loop {r.loop_kernel_ms:.2f} ms -> {r.loop_kernel_ms / iters * 1e3:.2f} us/iter, "
              f"{iters / (r.run_ms / 1e3):.0f} iters/s; multi classes {st.n_multi_classes} nnz {st.nnz_multi} active {st.n_active_txps}")
    if variant == 0:
        break
a, s, ok = ctx.download()
print("alpha sum", s, ok)
