"""Diagnostic for tests/test_map_gpu.py::test_long_reads_chunks_and_variants: which option set differs from the oracle,
on which reads, and is it stable from run to run."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as oracle
from salmon_b200._capi import Index, MapContext, map_default_params
from salmon_b200.synth import synth_reads, synth_txome

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
only = sys.argv[3] if len(sys.argv) > 3 else None
txps, _ = synth_txome(seed=21, n_genes=120)
left, right, _ = synth_reads(txps, seed=22, n=3000, read_len=150, frag_mean=320, frag_sd=30, indel_rate=0.003)
left[11, 5] = 4
left, right = left[:n], right[:n]
p = map_default_params()
idx = Index(txps)
ref = oracle.map_reads(oracle.MapIndex(txps), oracle.map_params(), left, right, 0)
asc = np.frombuffer(b"ACGTN", dtype=np.uint8)
for name, opts in (("default", {}), ("chunk700", dict(chunk=700)), ("variant0", dict(variant=0)), ("fast_dp0", dict(fast_dp=0)),
                   ("ascii", dict(ascii_reads=1))):
    if only and name != only:
        continue
    for rep in range(reps):
        ctx = MapContext(idx, p, batch_cap=4096, max_read_len=150)
        for k, v in opts.items():
            ctx.set_option(k, v)
        st = ctx.map_batch(asc[left], asc[right]) if opts.get("ascii_reads") else ctx.map_batch(left, right)
        got = ctx.last_alignments()
        bad = np.nonzero(got["n_aln"] != ref["n_aln"])[0]
        m = np.arange(p.max_read_occ)[None, :] < np.minimum(got["n_aln"], ref["n_aln"])[:, None]
        bad_s = np.nonzero(((got["score"] != ref["score"]) & m).any(axis=1))[0]
        bad_t = np.nonzero(((got["tid"] != ref["tid"]) & m).any(axis=1))[0]
        print(f"{name} rep {rep}: n_aln differs on {bad.size} reads {bad[:8].tolist()} got {got['n_aln'][bad[:8]].tolist()} "
              f"want {ref['n_aln'][bad[:8]].tolist()}; score differs on {bad_s.size} {bad_s[:8].tolist()}; tid differs on {bad_t.size}; "
              f"mapped {st.mapped} vs {ref['counters']['mapped']}", flush=True)
        ctx.close()
