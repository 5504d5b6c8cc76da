"""Stage A chunk-size sweep at human-transcriptome scale (dev helper).
usage: sweep_map.py n_genes n_reads batch chunks(comma)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from salmon_b200._capi import Index, MapContext, map_default_params, pin
from salmon_b200.synth import synth_txome, synth_reads_fast, flatten_txome

n_genes = int(sys.argv[1]); n_reads = int(sys.argv[2]); batch = int(sys.argv[3])
chunks = [int(x) for x in sys.argv[4].split(",")]
txps, _ = synth_txome(seed=44, n_genes=n_genes); flat = flatten_txome(txps)
left, right, truth = synth_reads_fast(txps, seed=7, n=n_reads, flat=flat)
idx = Index(txps); print("index:", idx.info(), flush=True)
p = map_default_params()
pin(left); pin(right)
ref = None
for chunk in chunks:
    for rep in range(2):
        ctx = MapContext(idx, p, batch_cap=batch, max_read_len=left.shape[1])
        ctx.set_option("chunk", min(chunk, batch))
        t0 = time.time(); dev = 0.0
        for s in range(0, n_reads, batch):
            st = ctx.map_batch(left[s:s + batch], right[s:s + batch]); dev += st.device_ms
        t1 = time.time(); res = ctx.finish(); t2 = time.time()
        ctx.close()
    key = (len(res["counts"]), int(res["counts"].sum()), int(res["tids"].sum()))
    if ref is None: ref = key
    print(f"chunk {chunk:7d}: device {dev:.1f} ms -> {n_reads/dev/1e3:.2f} Mreads/s (H2D inside), wall incl finish {n_reads/(t2-t0)/1e6:.2f} Mreads/s; "
          f"same classes as first: {key == ref}", flush=True)
