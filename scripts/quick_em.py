"""Quick device timing of the EM loop on BASELINE config 2 (dev helper, not the bench).
usage: quick_em.py iters cfgs keep_cm:keep_tm,... [variants] [vbem]"""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
configs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
keeps = [tuple(int(y) for y in x.split(":")) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [(100, 50)]
variants = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
vbem = int(sys.argv[5]) if len(sys.argv) > 5 else 1
lmaxs = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [96]
eq, proj, eff, uniq = synth_eq(seed=1)
ctx = EMContext(0)
p = default_params(min_iter=iters, max_iter=iters, use_vbem=vbem)
t = time.time(); ctx.upload(eq, proj, eff, uniq); print("upload s", time.time() - t)
ref = None
for cfg, (k1, k2), variant, lmax in itertools.product(configs, keeps, variants, lmaxs):
    ctx.set_option("lmax", lmax); ctx.set_option("config", cfg); ctx.set_option("variant", variant)
    ctx.set_option("l2_keep_cm", k1); ctx.set_option("l2_keep_tm", k2)
    st = ctx.prepare(p)
    for rep in range(2):
        ctx.flush_l2()
        r = ctx.run()
    a, s, ok = ctx.download()
    if ref is None: ref = a
    print(f"lmax {lmax} cfg {cfg} keep {k1}:{k2} variant {variant}: prepare {st.prepare_ms:.2f} ms, run {r.run_ms:.2f} ms, "
          f"loop {r.loop_kernel_ms / iters * 1e3:.2f} us/iter, {iters / (r.run_ms / 1e3):.0f} iters/s; "
          f"maxdiff vs first {np.max(np.abs(a - ref) / np.maximum(ref, 1e-6)):.1e}", flush=True)
print("multi classes", st.n_multi_classes, "nnz", st.nnz_multi, "active", st.n_active_txps, "alpha sum", s, ok)
