#!/bin/bash
# BASELINE.json configs[3] and [4] on N GPUs of one box through the C++ driver (`sb_salmon quant --gpus N`), at a size the
# GPU budget of a development round allows (stated in the output): decoy-aware index, 2x150 bp pairs read-sharded over
# the GPUs with the alpha exchange inside the EM kernel; 100 Gibbs samples on human-scale classes split over the GPUs.
set -e
N=${1:-8}
GENES=${2:-15000}
PAIRS=${3:-4000000}
D=/dev/shm/sb_cfg34; mkdir -p $D/idx gpurun_out
python - <<PY
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from salmon_b200 import _capi
from salmon_b200._capi import Index
from salmon_b200.synth import synth_txome, synth_reads_fast, flatten_txome, synth_eq
t0 = time.time()
txps, _ = synth_txome(seed=44, n_genes=$GENES)
n_real = len(txps)
# decoys: "chromosomes" of random sequence with transcript pieces embedded, so that decoy hits occur
rng = np.random.default_rng(99)
decoys = []
for c in range(24):
    chrom = rng.integers(0, 4, size=${DECOY_LEN:-2000000}, dtype=np.uint8)
    for _ in range(400):
        t = txps[int(rng.integers(n_real))]
        a = int(rng.integers(0, max(1, len(t) - 300))); piece = t[a:a + 300]
        o = int(rng.integers(0, len(chrom) - 400)); chrom[o:o + len(piece)] = piece
    decoys.append(chrom)
allref = txps + decoys
names = [f"ENST{i:08d}.1" for i in range(n_real)] + [f"chr{c+1}" for c in range(24)]
idx = Index(allref, names=names, first_decoy=n_real); idx.save("$D/idx/sb_index.bin")
print(f"decoy-aware index: {n_real} transcripts + 24 decoy sequences, {sum(len(x) for x in allref)/1e6:.0f} Mb, {time.time()-t0:.0f} s", flush=True)
flat = flatten_txome(txps)
lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
left, right, _ = synth_reads_fast(txps, seed=7, n=$PAIRS, read_len=150, flat=flat)
for tag, codes in (("1", left), ("2", right)):
    m, L = codes.shape
    rec = np.empty((m, 3 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8); rec[:, 3:3 + L] = lut[codes]
    rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8); rec[:, 6 + L:6 + 2 * L] = ord("I"); rec[:, 6 + 2 * L] = 10
    rec.tofile(f"$D/r_{tag}.fq")
print(f"reads: $PAIRS 2x150 bp pairs, {time.time()-t0:.0f} s", flush=True)
eq, proj, eff, uniq = synth_eq(seed=1)
_capi.write_eq_classes("$D/eq.txt.gz", [f"t{i}" for i in range(eq.n_txps)], eq.off, eq.tids, eq.counts, eq.weights)
print(f"classes: {eq.n_classes} classes / {eq.n_txps} transcripts written, {time.time()-t0:.0f} s", flush=True)
PY
EXE=salmon_b200/sb_salmon
tm() { local t0=$(date +%s.%N); "$@"; local rc=$?; echo "$(python -c "import time; print(round(time.time() - $t0, 2))") s wall"; return $rc; }
echo "== configs[3]: $N GPUs, read-sharded"; timeout 90 $EXE quant -i $D/idx -l IU -1 $D/r_1.fq -2 $D/r_2.fq -o $D/c3_multi -p 8 --maxReadLen 160 --gpus $N 2>&1 | grep -v NCCL | tail -4
echo "== configs[3]: 1 GPU";  timeout 90 $EXE quant -i $D/idx -l IU -1 $D/r_1.fq -2 $D/r_2.fq -o $D/c3_one -p 32 --maxReadLen 160 2>&1 | tail -4
echo "== configs[4]: 100 Gibbs samples, $N GPUs"; tm timeout 90 $EXE quant -e $D/eq.txt.gz -o $D/c4_multi --numGibbsSamples 100 --seed 5 --gpus $N 2>&1 | grep -v NCCL | tail -4
echo "== configs[4]: 100 Gibbs samples, 1 GPU"; tm timeout 90 $EXE quant -e $D/eq.txt.gz -o $D/c4_one --numGibbsSamples 100 --seed 5 2>&1 | tail -4
python - <<PY
import gzip, json, numpy as np
D = "$D"
def sf(p):
    rows = open(p).read().splitlines()[1:]
    return np.array([float(r.split("\t")[4]) for r in rows]), np.array([float(r.split("\t")[3]) for r in rows])
a1, t1 = sf(D + "/c3_one/quant.sf"); aN, tN = sf(D + "/c3_multi/quant.sf")
m1 = json.load(open(D + "/c3_one/aux_info/meta_info.json")); mN = json.load(open(D + "/c3_multi/aux_info/meta_info.json"))
print(f"configs[3]: rows {len(a1)} (decoys dropped: {m1['num_decoy_targets']}), mapped {m1['num_mapped']} vs {mN['num_mapped']} of {mN['num_processed']} "
      f"on {mN['sb_num_gpus']} GPUs, corr(NumReads) {np.corrcoef(a1, aN)[0,1]:.6f}, TPM rel diff > 1e-4 on TPM > 1: "
      f"{int(((np.abs(t1 - tN) / np.maximum(t1, 1e-9) > 1e-4) & (t1 > 1)).sum())} of {int((t1 > 1).sum())}")
def boots(p, n):
    raw = gzip.open(p + "/aux_info/bootstrap/bootstraps.gz", "rb").read()
    return np.frombuffer(raw, dtype=np.float64).reshape(n, -1)
g1, gN = boots(D + "/c4_one", 100), boots(D + "/c4_multi", 100)
eff = np.full(g1.shape[1], 100.0)
tp = lambda g: ((g / eff) / (g / eff).sum(axis=1, keepdims=True)).mean(axis=0)
print(f"configs[4]: 100 samples x {g1.shape[1]} transcripts; posterior-mean TPM fraction max |{mN['sb_num_gpus']} GPUs - 1 GPU| = "
      f"{np.abs(tp(g1) - tp(gN)).max():.2e} (criterion 1e-3); totals {g1.sum(axis=1).mean():.1f} vs {gN.sum(axis=1).mean():.1f}")
PY
rm -rf $D
