"""Quick device timing of the EM loop on BASELINE config 2 (dev helper, not the bench)."""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
configs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
ovh = [tuple(int(y) for y in x.split(":")) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [(6, 20)]
variants = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
eq, proj, eff, uniq = synth_eq(seed=1)
ctx = EMContext(0)
p = default_params(min_iter=iters, max_iter=iters)
t = time.time(); ctx.upload(eq, proj, eff, uniq); print("upload s", time.time() - t)
ref = None
for cfg, (o1, o2), variant in itertools.product(configs, ovh, variants):
    ctx.set_option("config", cfg); ctx.set_option("variant", variant)
    ctx.set_option("overhead_p1", o1); ctx.set_option("overhead_p2", o2)
    st = ctx.prepare(p)
    for rep in range(2):
        r = ctx.run()
    a, s, ok = ctx.download()
    if ref is None: ref = a
    print(f"cfg {cfg} ovh {o1}:{o2} variant {variant}: prepare {st.prepare_ms:.2f} ms, run {r.run_ms:.2f} ms, "
          f"loop {r.loop_kernel_ms / iters * 1e3:.2f} us/iter, {iters / (r.run_ms / 1e3):.0f} iters/s; "
          f"maxdiff vs first {np.max(np.abs(a - ref) / np.maximum(ref, 1e-6)):.1e}", flush=True)
print("multi classes", st.n_multi_classes, "nnz", st.nnz_multi, "active", st.n_active_txps, "alpha sum", s, ok)
