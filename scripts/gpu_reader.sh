#!/bin/bash
# where does the host reader's time go on the GPU box?  (4 M pairs of configs[2]-like reads, small index)
D=/dev/shm/sb_rd; mkdir -p $D/idx gpurun_out
python - <<PY
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from salmon_b200._capi import Index
from salmon_b200.synth import synth_txome, synth_reads_fast, flatten_txome
txps, _ = synth_txome(seed=44, n_genes=8000); flat = flatten_txome(txps)
idx = Index(txps, names=[f"T{i}" for i in range(len(txps))]); idx.save("$D/idx/sb_index.bin")
lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
left, right, _ = synth_reads_fast(txps, seed=7, n=6_000_000, read_len=100, flat=flat)
for tag, codes in (("1", left), ("2", right)):
    m, L = codes.shape
    rec = np.empty((m, 3 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8); rec[:, 3:3 + L] = lut[codes]
    rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8); rec[:, 6 + L:6 + 2 * L] = ord("I"); rec[:, 6 + 2 * L] = 10
    rec.tofile(f"$D/r_{tag}.fq")
PY
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread"
export SB_READS_PROFILE=1
gzip -1 -c $D/r_1.fq > $D/g_1.fq.gz & gzip -1 -c $D/r_2.fq > $D/g_2.fq.gz; wait
Q="salmon_b200/sb_salmon quant -i $D/idx -l IU -o $D/out --maxReadLen 128"
G="sb_reads|sb_quant|mapping|index loaded|done "
echo "== plain p32"; $Q -1 $D/r_1.fq -2 $D/r_2.fq -p 32 2>&1 | grep -E "$G"
echo "== plain p32 again"; $Q -1 $D/r_1.fq -2 $D/r_2.fq -p 32 2>&1 | grep -E "$G"
echo "== gz p32"; $Q -1 $D/g_1.fq.gz -2 $D/g_2.fq.gz -p 32 2>&1 | grep -E "$G"
echo "== gz p64"; $Q -1 $D/g_1.fq.gz -2 $D/g_2.fq.gz -p 64 2>&1 | grep -E "$G"
echo "== gz p32 one inflate thread per file"; SB_READS_INFLATERS=1 $Q -1 $D/g_1.fq.gz -2 $D/g_2.fq.gz -p 32 2>&1 | grep -E "$G"
echo "== reader only, gz, p32"; python scripts/bench_reader.py 4000000 100 32 gz 2>&1 | tail -4
echo "== reader only, gz, p64"; python scripts/bench_reader.py 4000000 100 64 gz 2>&1 | tail -4
rm -rf $D /dev/shm/rb
