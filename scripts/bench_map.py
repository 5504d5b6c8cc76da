"""Stage A dev timing: synthetic transcriptome -> index -> map batches -> finish.
usage: bench_map.py n_genes n_reads batch [reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200._capi import Index, MapContext, map_default_params, pin
from salmon_b200.synth import synth_txome, synth_reads_fast, flatten_txome

n_genes = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t0 = time.time(); txps, _ = synth_txome(seed=44, n_genes=n_genes); flat = flatten_txome(txps)
print(f"txome: {len(txps)} txps, {flat[1].shape[0]/1e6:.1f} Mb, {time.time()-t0:.1f}s", flush=True)
t0 = time.time(); left, right, truth = synth_reads_fast(txps, seed=7, n=n_reads, flat=flat)
print(f"reads: {n_reads} pairs {time.time()-t0:.1f}s", flush=True)
t0 = time.time(); idx = Index(txps); print("index:", idx.info(), f"{time.time()-t0:.1f}s", flush=True)
p = map_default_params()
pin(left); pin(right)
for rep in range(reps):
    ctx = MapContext(idx, p, batch_cap=batch, max_read_len=left.shape[1])
    t0 = time.time(); dev = 0.0; full = 0
    for s in range(0, n_reads, batch):
        st = ctx.map_batch(left[s:s + batch], right[s:s + batch]); dev += st.device_ms; full += st.full_dp
    t1 = time.time(); res = ctx.finish(); t2 = time.time()
    c = res["counters"]
    print(f"rep {rep}: map wall {t1-t0:.3f}s dev {dev:.1f}ms -> {n_reads/dev/1e3:.3f} Mreads/s (device), "
          f"{n_reads/(t2-t0)/1e6:.3f} Mreads/s (wall incl finish {t2-t1:.3f}s); classes {len(res['counts'])} "
          f"full_dp {full} mapped {c['n_mapped']} lookups {c['lookups']} postings {c['postings']} cands {c['candidates']} kept {c['kept']}",
          flush=True)
    ctx.close()
