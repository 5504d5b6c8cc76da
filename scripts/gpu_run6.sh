#!/bin/bash
mkdir -p gpurun_out
SB_MAP_OVERLAP=1 timeout 600 python -m pytest tests/test_map_gpu.py tests/test_pipeline_gpu.py -m gpu -q --maxfail=5 > gpurun_out/tests_gpu6_overlap.txt 2>&1
tail -4 gpurun_out/tests_gpu6_overlap.txt
cat > /tmp/ovl.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from salmon_b200._capi import Index, MapContext, map_default_params, pin
from salmon_b200.synth import synth_txome, synth_reads_fast, flatten_txome
import torch
txps, _ = synth_txome(seed=44, n_genes=60000); flat = flatten_txome(txps)
n, batch = 1048576, 262144
left, right, truth = synth_reads_fast(txps, seed=7, n=n, flat=flat)
idx = Index(txps); p = map_default_params(); pin(left); pin(right)
dl = torch.from_numpy(left).cuda(); dr = torch.from_numpy(right).cuda()
L = left.shape[1]
ref = None
for ovl, dev in ((0, 0), (1, 0), (0, 1), (1, 1), (0, 1), (1, 1)):
    ctx = MapContext(idx, p, batch_cap=batch, max_read_len=L)
    ctx.set_option("overlap_assign", ovl); ctx.set_option("input_on_device", dev)
    for rep in range(2):
        ctx.reset(); t0 = time.time(); dms = 0.0
        for s in range(0, n, batch):
            st = (ctx.map_batch_ptr(dl.data_ptr() + s * L, dr.data_ptr() + s * L, batch, L) if dev else ctx.map_batch(left[s:s + batch], right[s:s + batch]))
            dms += st.device_ms
        res = ctx.finish(); wall = time.time() - t0
    key = (len(res["counts"]), int(res["counts"].sum()), int(res["tids"].astype(np.int64).sum()), float(res["weights"].sum()))
    if ref is None: ref = key
    print(f"overlap {ovl} input_on_device {dev}: device {dms:.1f} ms = {n/dms/1e3:.2f} Mreads/s; wall incl finish {n/wall/1e6:.2f} Mreads/s; identical classes: {key == ref}", flush=True)
    ctx.close()
PY
timeout 600 python /tmp/ovl.py > gpurun_out/overlap_ab.txt 2>&1; tail -8 gpurun_out/overlap_ab.txt
