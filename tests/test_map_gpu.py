"""GPU parity of Stage A (mapping -> scoring -> filtering -> labels -> equivalence classes) against
the oracle: per-read alignments, scores, probabilities, labels, weights bit-exact; class labels and
counts bit-exact; class weights bit-exact within one batch (same summation order) and 1e-12 across
batches (batch-wise association)."""
import numpy as np
import pytest

from salmon_b200 import EMContext, default_params
from salmon_b200._capi import EqClasses, Index, MapContext, map_default_params
from salmon_b200.synth import synth_reads, synth_txome
from test_map_host import compare

pytestmark = pytest.mark.gpu


def canon(res):
    """classes as a list sorted by full label (tids ++ bins)."""
    off = res["off"].astype(np.int64)
    out = []
    for c in range(len(res["counts"])):
        t = tuple(res["tids"][off[c]:off[c + 1]].tolist())
        b = tuple(res["bins"][off[c]:off[c + 1]].tolist()) if res.get("bins") is not None else ()
        out.append((t + b, t, res["weights"][off[c]:off[c + 1]].copy(), int(res["counts"][c])))
    out.sort(key=lambda x: (x[0]))
    return out


def oracle_classes(oracle, m, cap, binned=True):
    e = oracle.eq_aggregate(m, cap, binned)
    # recover bins from the first fragment of each class is not needed: orc_eq_aggregate sorts by full label
    return e


def check_classes(got, ref_e, exact_weights=True):
    g = canon(got)
    assert len(g) == len(ref_e["counts"])
    off = ref_e["off"].astype(np.int64)
    for c, (lab, t, w, cnt) in enumerate(g):
        assert list(t) == ref_e["tids"][off[c]:off[c + 1]].tolist()
        assert cnt == int(ref_e["counts"][c])
        rw = ref_e["weights"][off[c]:off[c + 1]]
        if exact_weights:
            assert np.array_equal(w.view(np.uint64), rw.view(np.uint64))
        else:
            np.testing.assert_allclose(w, rw, rtol=1e-12, atol=0)


@pytest.mark.parametrize("frag_counter_reads", [0, 1])
def test_single_batch_bit_exact(oracle, frag_counter_reads):
    txps, _ = synth_txome(seed=3, n_genes=200)
    left, right, truth = synth_reads(txps, seed=5, n=6000, indel_rate=0.002)
    p = map_default_params(num_pre_burnin=0 if frag_counter_reads else 5000)
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=8192, max_read_len=100)
    st = ctx.map_batch(left, right)
    got = ctx.last_alignments()
    oix = oracle.MapIndex(txps)
    op = oracle.map_params(num_pre_burnin=0 if frag_counter_reads else 5000)
    ref = oracle.map_reads(oix, op, left, right, 0)
    compare(got, ref, p.max_read_occ)
    for k in ("lookups", "postings", "seeds", "kept", "label_entries", "mapped"):
        assert getattr(st, k) == ref["counters"][k], k
    res = ctx.finish()
    check_classes(res, oracle.eq_aggregate(ref, p.max_read_occ, True), exact_weights=True)
    assert res["counters"]["n_mapped"] == ref["counters"]["mapped"]
    assert st.gpu_launches > 0
    ctx.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def check_state(got, ref):
    for k in ("assigned", "frags_seen", "timestep", "burned_in", "min_val"):
        assert got[k] == ref[k], (k, got[k], ref[k])
    assert np.array_equal(bits(got["mass"]), bits(ref["mass"]))          # integer accumulation: order-free, bit-exact
    assert np.array_equal(bits(got["hist"]), bits(ref["hist"]))
    assert bits(np.array([got["tot"]]))[0] == bits(np.array([ref["tot"]]))[0]
    assert np.array_equal(bits(got["log_eff"]), bits(ref["log_eff"]))


def test_multi_batch_online_state_and_regimes(oracle):
    """three batches: masses and the fragment-length distribution evolve (batched semantics), the fragment counter
    crosses numPreBurninFrags and numBurninFrags between batches, effective lengths and the cached FLD tables are
    built at burn-in; then normalizeAlphas.  State bit-exact after every batch."""
    txps, _ = synth_txome(seed=4, n_genes=120)
    left, right, _ = synth_reads(txps, seed=6, n=9000)
    over = dict(num_pre_burnin=2500, num_burnin=5500)
    p = map_default_params(mini_batch=1000, seed=7, **over)
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=3000, max_read_len=100)
    oix = oracle.MapIndex(txps)
    on = oracle.Online(oix, oracle.map_params(**over), seed=7, mini_batch=1000)
    check_state(ctx.online_state(), on.state())
    parts = []
    for b in range(3):
        sl = slice(3000 * b, 3000 * (b + 1))
        ctx.map_batch(left[sl], right[sl])
        ref = on.batch(left[sl], right[sl])
        compare(ctx.last_alignments(), ref, p.max_read_occ)
        check_state(ctx.online_state(), on.state())
        parts.append(ref)
    st = on.state()
    assert st["burned_in"] == 1 and st["timestep"] == 9 and st["min_val"] < 1000
    res = ctx.finish()
    cap = p.max_read_occ
    merged = {k: np.concatenate([q[k] for q in parts]) for k in ("n_aln", "label", "weight")}
    ref_e = oracle.eq_aggregate(merged, cap, True)
    check_classes(res, ref_e, exact_weights=False)
    fin = on.finish(res["off"], res["tids"], res["counts"])
    assert np.array_equal(res["unique_counts"], fin["unique_counts"])
    assert np.array_equal(res["total_counts"], fin["total_counts"])
    assert np.array_equal(bits(res["eff_len"]), bits(fin["eff_len"]))
    np.testing.assert_allclose(res["projected_counts"], fin["projected_counts"], rtol=1e-9, atol=1e-9)
    assert abs(res["projected_counts"].sum() - st["assigned"]) < 1e-6 * st["assigned"]
    ctx.close()


def test_never_burned_in_effective_lengths(oracle):
    """fewer fragments than numBurninFrags: effective lengths come from the observed FLD at finish."""
    txps, _ = synth_txome(seed=14, n_genes=60)
    left, right, _ = synth_reads(txps, seed=16, n=4000)
    p = map_default_params(num_pre_burnin=1000)
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=2048, max_read_len=100)
    on = oracle.Online(oracle.MapIndex(txps), oracle.map_params(num_pre_burnin=1000), seed=42, mini_batch=5000)
    for sl in (slice(0, 2048), slice(2048, 4000)):
        ctx.map_batch(left[sl], right[sl])
        compare(ctx.last_alignments(), on.batch(left[sl], right[sl]), p.max_read_occ)
    check_state(ctx.online_state(), on.state())
    res = ctx.finish()
    fin = on.finish(res["off"], res["tids"], res["counts"])
    assert np.array_equal(bits(res["eff_len"]), bits(fin["eff_len"]))
    lens = np.array([t.shape[0] for t in txps], dtype=np.float64)
    assert np.all(res["eff_len"] <= lens) and np.all(res["eff_len"] >= 1.0)
    np.testing.assert_allclose(res["projected_counts"], fin["projected_counts"], rtol=1e-9, atol=1e-9)
    ctx.close()


def test_edge_cases_gpu(oracle):
    rng = np.random.default_rng(9)
    unit = rng.integers(0, 4, size=400, dtype=np.uint8)
    txps = [unit.copy() for _ in range(90)]
    txps += [np.concatenate([unit[:200], rng.integers(0, 4, size=150, dtype=np.uint8)]) for _ in range(130)]
    txps += [rng.integers(0, 4, size=20, dtype=np.uint8)]
    withn = rng.integers(0, 4, size=500, dtype=np.uint8); withn[100:103] = 4
    txps += [withn]
    left, right, _ = synth_reads(txps, seed=2, n=500, read_len=75, frag_mean=180, frag_sd=15, random_frac=0.1)
    left[5, 10] = 4; right[7, 60] = 4
    for over in (dict(), dict(max_read_occ=50, max_occs_per_hit=64, stride=3, range_bins=0),
                 dict(first_decoy=200, decoy_threshold=0.9), dict(hard_filter=1)):
        p = map_default_params(**over)
        idx = Index(txps)
        ctx = MapContext(idx, p, batch_cap=512, max_read_len=75)
        ctx.map_batch(left, right)
        ref = oracle.map_reads(oracle.MapIndex(txps), oracle.map_params(**over), left, right, 0)
        compare(ctx.last_alignments(), ref, p.max_read_occ)
        res = ctx.finish()
        check_classes(res, oracle.eq_aggregate(ref, p.max_read_occ, p.range_bins > 0), exact_weights=True)
        ctx.close()
    # an empty batch and a batch of unmappable reads
    p = map_default_params()
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=64, max_read_len=75)
    junk = rng.integers(0, 4, size=(32, 75), dtype=np.uint8)
    st = ctx.map_batch(junk, junk)
    assert st.mapped == 0
    res = ctx.finish()
    assert len(res["counts"]) == 0
    ctx.close()


def test_map_then_em_end_to_end(oracle):
    """config-1-style plumbing: reads -> eq classes (GPU) -> VBEM (GPU) vs the oracle chain."""
    txps, _ = synth_txome(seed=11, n_genes=150)
    left, right, _ = synth_reads(txps, seed=12, n=8000)
    p = map_default_params()
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=8192, max_read_len=100)
    ctx.map_batch(left, right)
    res = ctx.finish()
    ctx.close()
    M = len(txps)
    eq = EqClasses(M, res["off"], res["tids"], res["weights"], res["counts"])
    lens = np.array([t.shape[0] for t in txps], dtype=np.float64)
    eff = np.maximum(1.0, lens - 250.0 + 1.0)
    sizes = (eq.off[1:] - eq.off[:-1]).astype(np.int64)
    proj = np.bincount(eq.tids, weights=np.repeat(eq.counts.astype(float), sizes) * eq.weights, minlength=M)
    uniq = np.bincount(eq.tids[np.repeat(sizes == 1, sizes)], weights=eq.counts[sizes == 1].astype(float),
                       minlength=M).astype(np.uint64)
    em = EMContext(0)
    alpha, st, ok = em.optimize(eq, default_params(), proj, eff, uniq)
    ref, rst = oracle.em_optimize(eq, proj, eff, uniq, default_params())
    assert ok and st.iters == rst.iters
    np.testing.assert_allclose(alpha, ref, rtol=1e-9, atol=1e-9)
    assert abs(alpha.sum() - float(eq.counts.sum())) < 1e-6 * float(eq.counts.sum())
    em.close()


def test_long_reads_chunks_and_variants(oracle):
    """2x150 bp (the 4-word coverage / 8-word read paths), a batch cut into several pipeline chunks, and
    every kernel variant: all bit-identical to the oracle."""
    txps, _ = synth_txome(seed=21, n_genes=120)
    left, right, _ = synth_reads(txps, seed=22, n=3000, read_len=150, frag_mean=320, frag_sd=30, indel_rate=0.003)
    left[11, 5] = 4
    p = map_default_params()
    idx = Index(txps)
    oix = oracle.MapIndex(txps)
    ref = oracle.map_reads(oix, oracle.map_params(), left, right, 0)
    ref_e = oracle.eq_aggregate(ref, p.max_read_occ, True)
    asc = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for opts in (dict(), dict(chunk=700), dict(variant=0), dict(fast_dp=0), dict(ascii_reads=1)):
        ctx = MapContext(idx, p, batch_cap=4096, max_read_len=150)
        for k, v in opts.items():
            ctx.set_option(k, v)
        st = ctx.map_batch(asc[left], asc[right]) if opts.get("ascii_reads") else ctx.map_batch(left, right)
        compare(ctx.last_alignments(), ref, p.max_read_occ)
        for k in ("lookups", "postings", "seeds", "kept", "label_entries", "mapped"):
            assert getattr(st, k) == ref["counters"][k], (opts, k)
        if opts.get("variant", 1) == 1 and opts.get("fast_dp", 1) == 1:
            assert st.full_dp < st.candidates          # the ungapped shortcut resolved some alignments
        check_classes(ctx.finish(), ref_e, exact_weights=True)
        ctx.close()


def test_variants_agree_at_scale():
    """no oracle (too slow at this size): warp kernels == serial-form kernels on 150k pairs, two batches."""
    from salmon_b200.synth import synth_reads_fast
    txps, _ = synth_txome(seed=31, n_genes=1500)
    left, right, _ = synth_reads_fast(txps, seed=32, n=150_000, indel_rate=0.001)
    p = map_default_params()
    idx = Index(txps)
    outs = []
    for variant in (1, 0):
        ctx = MapContext(idx, p, batch_cap=100_000, max_read_len=100)
        ctx.set_option("variant", variant)
        stats = [ctx.map_batch(left[:100_000], right[:100_000])]
        a = ctx.last_alignments()
        stats.append(ctx.map_batch(left[100_000:], right[100_000:]))
        res = ctx.finish()
        outs.append((a, canon(res), [(s.mapped, s.lookups, s.postings, s.seeds, s.candidates, s.kept) for s in stats]))
        ctx.close()
    compare(outs[0][0], outs[1][0], p.max_read_occ)
    assert outs[0][2] == outs[1][2]
    assert len(outs[0][1]) == len(outs[1][1])
    for x, y in zip(outs[0][1], outs[1][1]):
        assert x[0] == y[0] and x[3] == y[3]
        assert np.array_equal(x[2].view(np.uint64), y[2].view(np.uint64))


def test_quant_pipeline_end_to_end(tmp_path):
    """config-0 style plumbing + accuracy: reads with known origin -> map -> classes -> normalizeAlphas -> VBEM -> TPM;
    estimated read counts track the simulated truth, files are written in the reference's layouts."""
    from salmon_b200.quant import quant_reads
    txps, _ = synth_txome(seed=41, n_genes=150)
    left, right, truth = synth_reads(txps, seed=42, n=30000)
    idx = Index(txps)
    out = quant_reads(idx, left, right, batch=8192, out_dir=str(tmp_path), dump_eq_weights=True)
    M = len(txps)
    true_counts = np.bincount(truth["tid"][truth["tid"] >= 0], minlength=M).astype(float)
    assert abs(out["alpha"].sum() - out["n_mapped"]) < 1e-6 * out["n_mapped"]
    assert out["n_mapped"] >= 0.97 * (truth["tid"] >= 0).sum()
    r = np.corrcoef(out["alpha"], true_counts)[0, 1]
    assert r > 0.97, r
    assert abs(out["tpm"].sum() - 1e6) < 1e-3
    lines = (tmp_path / "quant.sf").read_text().splitlines()
    assert lines[0].split("\t") == ["Name", "Length", "EffectiveLength", "TPM", "NumReads"] and len(lines) == M + 1
    import gzip
    eqt = gzip.open(tmp_path / "aux_info" / "eq_classes.txt.gz", "rt").read().splitlines()
    assert int(eqt[0]) == M and int(eqt[1]) == len(out["classes"]["counts"])


@pytest.mark.parametrize("L,cap_len", [(60, 128), (100, 128), (31, 100), (140, 256)])
def test_read_length_below_context_capacity(oracle, L, cap_len):
    """One context serves batches of different read lengths (quant_files groups reads by length): a batch of L-base
    reads in a context created for cap_len bases is bit-exact against the oracle, like a context created for L."""
    txps, _ = synth_txome(seed=3, n_genes=100)
    left, right, _ = synth_reads(txps, seed=15, n=2000, read_len=L, frag_mean=max(200, 2 * L), frag_sd=20)
    p = map_default_params()
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=2048, max_read_len=cap_len)
    st = ctx.map_batch(left, right)
    got = ctx.last_alignments()
    ref = oracle.map_reads(oracle.MapIndex(txps), oracle.map_params(), left, right, 0)
    compare(got, ref, p.max_read_occ)
    for k in ("lookups", "postings", "seeds", "kept", "label_entries", "mapped"):
        assert getattr(st, k) == ref["counters"][k], k
    # and a second batch of another length through the same context
    left2, right2, _ = synth_reads(txps, seed=16, n=1000, read_len=max(31, L - 9), frag_mean=max(200, 2 * L), frag_sd=20)
    ctx.map_batch(left2, right2)
    got2 = ctx.last_alignments()
    ref2 = oracle.map_reads(oracle.MapIndex(txps), oracle.map_params(), left2, right2, 0)
    assert np.array_equal(got2["n_aln"], ref2["n_aln"])
    m = np.arange(p.max_read_occ)[None, :] < got2["n_aln"][:, None]
    assert np.array_equal(got2["tid"][m], ref2["tid"][m]) and np.array_equal(got2["score"][m], ref2["score"][m])
    ctx.close()


def test_decoy_aware_batch_bit_exact(oracle):
    """BASELINE configs[3] semantics on the GPU: decoys last in the index, first_decoy set -- no decoy alignment is
    reported, fragments that map best to a decoy are dropped, everything bit-exact against the oracle."""
    rng = np.random.default_rng(17)
    txps, _ = synth_txome(seed=8, n_genes=40)
    M = len(txps)
    decoys = []
    for g in range(6):
        parts = [rng.integers(0, 4, size=300, dtype=np.uint8)]
        for t in rng.choice(M, size=5, replace=False):
            parts += [txps[t], rng.integers(0, 4, size=200, dtype=np.uint8)]
        decoys.append(np.concatenate(parts))
    allseq = txps + decoys
    l1, r1, _ = synth_reads(txps, seed=18, n=2000)
    l2, r2, _ = synth_reads(decoys, seed=19, n=1000, expressed_frac=1.0)
    left, right = np.concatenate([l1, l2]), np.concatenate([r1, r2])
    p = map_default_params(first_decoy=M)
    ctx = MapContext(Index(allseq), p, batch_cap=4096, max_read_len=100)
    st = ctx.map_batch(left, right)
    got = ctx.last_alignments()
    ref = oracle.map_reads(oracle.MapIndex(allseq), oracle.map_params(first_decoy=M), left, right, 0)
    compare(got, ref, p.max_read_occ)
    assert st.mapped == ref["counters"]["mapped"]
    sel = np.arange(p.max_read_occ)[None, :] < got["n_aln"][:, None]
    assert (got["tid"][sel] < M).all() and (got["n_aln"][2000:] == 0).sum() > 100
    check_classes(ctx.finish(), oracle.eq_aggregate(ref, p.max_read_occ, True), exact_weights=True)
    ctx.close()


def test_gpu_vs_oracle_at_bench_scale(oracle):
    """VERDICT r1 next #3b: the CUDA path against the ORACLE (not against itself) at the bench's human-scale index
    (synth_txome(seed=44, n_genes=60000): ~293 k transcripts / 398 Mb / 113 M distinct 31-mers) on >= 1 M read pairs,
    several batches through sb_map_batch + sb_map_finish: per-read alignment counts, labels (transcripts + bins) and
    the merged class table (labels, counts) bit-exact; class weights 1e-12 (batch-wise association).  The auxiliary
    model stays in its first regime (num_pre_burnin above the read count) so that the labels do not depend on the online
    state and the stateless, OpenMP-parallel oracle applies to every batch.  Then the same reads through the oracle on
    the PLATFORM libm: every read whose label differs must be an exact-boundary case (tests/test_math_pinning.py).
    SB_SCALE_GENES overrides the transcriptome size (development)."""
    import os
    from salmon_b200.synth import synth_reads_fast
    from test_math_pinning import label_flips
    import oracle_lib as O
    n_genes = int(os.environ.get("SB_SCALE_GENES", "60000"))
    n = 1_048_576
    txps, _ = synth_txome(seed=44, n_genes=n_genes)
    left, right, _ = synth_reads_fast(txps, seed=7, n=n)
    over = dict(num_pre_burnin=10 ** 9, num_burnin=2 * 10 ** 9)
    p = map_default_params(**over)
    idx = Index(txps)
    ctx = MapContext(idx, p, batch_cap=262144, max_read_len=100)
    cap = p.max_read_occ
    got_naln, got_label = [], []
    for s in range(0, n, 262144):
        ctx.map_batch(left[s:s + 262144], right[s:s + 262144])
        a = ctx.last_alignments()
        got_naln.append(a["n_aln"].copy()); got_label.append(a["label"].copy())
        if s == 0:
            first = {k: v.copy() for k, v in a.items() if isinstance(v, np.ndarray)}
    res = ctx.finish()
    ctx.close()
    got_naln = np.concatenate(got_naln); got_label = np.concatenate(got_label)
    oix = oracle.MapIndex(txps)
    assert idx.info()["n_kmers"] == oix.n_kmers
    ref = oracle.map_reads(oix, oracle.map_params(**over), left, right, 0)          # fdlibm mode: bit-exact
    assert np.array_equal(got_naln, ref["n_aln"])
    m2 = np.arange(2 * cap)[None, :] < 2 * got_naln[:, None]
    assert np.array_equal(got_label[m2], ref["label"][m2])
    compare(first, {k: (v[:262144] if isinstance(v, np.ndarray) else v) for k, v in ref.items()}, cap)   # first batch: every field
    assert int(res["counts"].sum()) == int((got_naln > 0).sum()) == ref["counters"]["mapped"]
    check_classes(res, oracle.eq_aggregate(ref, cap, True), exact_weights=False)
    assert (got_naln > 1).sum() > 0.3 * n
    # ---- the platform libm
    with O.math_mode("libm"):
        lm = oracle.map_reads(oix, oracle.map_params(**over), left, right, 0)
    flips = label_flips(ref, lm, cap)
    assert len(flips) <= 0.02 * n, len(flips)
    print(f"bench-scale parity: {n} pairs, {len(res['counts'])} classes, {int((got_naln > 1).sum())} multi-mapping reads; "
          f"labels on the platform libm differ on {len(flips)} reads (exact bin boundaries)")


def test_join_policy_knobs_gpu_vs_oracle(oracle):
    """the warp kernels' lane-parallel form of the join policy (preMerge / postMerge / orphan thresholds, allowDovetail,
    discardOrphans) against the oracle: alignments, labels bit-exact for every knob setting"""
    from test_map_host import JOIN_VARIANTS, join_reads
    txps, left, right = join_reads()
    idx = Index(txps)
    oix = oracle.MapIndex(txps)
    for over in JOIN_VARIANTS:
        p = map_default_params(**over)
        for variant in (1, 0):
            ctx = MapContext(idx, p, batch_cap=8192, max_read_len=100)
            ctx.set_option("variant", variant)
            ctx.map_batch(left, right)
            got = ctx.last_alignments()
            ctx.close()
            ref = oracle.map_reads(oix, oracle.map_params(**over), left, right, 0)
            compare(got, ref, p.max_read_occ)


def test_library_types_and_single_end_gpu_vs_oracle(oracle):
    """row a1: expected library formats IU / ISF / ISR (paired) and U / SF / SR (single-end reads: sb_map_batch with
    right = None) on the CUDA path against the oracle -- alignments and labels bit-exact, class tables equal; the
    per-format fragment counts against a recount from the alignments"""
    from test_map_host import LIB, stranded_reads
    txps, left, right, flip = stranded_reads()
    idx = Index(txps)
    oix = oracle.MapIndex(txps)
    absent = np.full_like(right, 4)
    for lib, se in (("IU", False), ("ISF", False), ("ISR", False), ("U", True), ("SF", True), ("SR", True)):
        over = dict(lib_type=LIB[lib])
        if se:
            over["pre_merge_thresh"] = 1.0
        p = map_default_params(**over)
        ctx = MapContext(idx, p, batch_cap=8192, max_read_len=100)
        ctx.map_batch(left, None if se else right)
        got = ctx.last_alignments()
        res = ctx.finish()
        ctx.close()
        ref = oracle.map_reads(oix, oracle.map_params(**over), left, absent if se else right, 0)
        compare(got, ref, p.max_read_occ)
        check_classes(res, oracle.eq_aggregate(ref, p.max_read_occ, True))
        # observed formats per fragment (SalmonQuantify.cpp:765,1000-1002)
        cap = p.max_read_occ
        valid = np.arange(cap)[None, :] < got["n_aln"][:, None]
        st = (got["flags"] >> 2) & 3; fw = (got["flags"] & 1) == 1
        want = dict(ISF=int((valid & (st == 0) & fw).any(axis=1).sum()), ISR=int((valid & (st == 0) & ~fw).any(axis=1).sum()),
                    SF=int((valid & (st != 0) & fw).any(axis=1).sum()), SR=int((valid & (st != 0) & ~fw).any(axis=1).sum()))
        assert res["lib_format_counts"] == want, (lib, res["lib_format_counts"], want)
        if se:
            assert want["ISF"] == want["ISR"] == 0
    # a paired type without the second mate (and the reverse) is refused
    ctx = MapContext(idx, map_default_params(), batch_cap=64, max_read_len=100)
    with pytest.raises(Exception, match="both mates"):
        ctx.map_batch(left[:8], None)
    ctx.close()
