"""CPU tests of the oracle (oracle/em_oracle.c) against golden vectors and an
independent pure-Python restatement.  No GPU needed."""
import json
import os

import numpy as np
import pytest

from salmon_b200._capi import EqClasses, sb_em_params

HERE = os.path.dirname(os.path.abspath(__file__))


def mkparams(**kw):
    d = dict(use_vbem=1, per_txp_prior=1, init_uniform=0, eq_class_mode=0, no_rich_eq=0,
             no_length_correction=0, alt_init=0, vb_prior=1e-2, tol=0.01, num_required_frags=5e7,
             min_iter=100, max_iter=10000)
    d.update(kw)
    return d


def random_problem(rng, C, M, max_label=6, zero_frac=0.0):
    sizes = rng.integers(1, max_label + 1, size=C)
    sizes = np.minimum(sizes, M)
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    tids = np.concatenate([np.sort(rng.choice(M, size=s, replace=False)) for s in sizes]).astype(np.uint32)
    w = rng.random(int(off[-1])) + 0.01
    if zero_frac:
        w[rng.random(w.shape[0]) < zero_frac] = 0.0
    s = np.add.reduceat(w, off[:-1].astype(np.int64))
    s[s == 0] = 1.0
    w = w / np.repeat(s, sizes)
    counts = rng.integers(1, 50, size=C).astype(np.uint64)
    eq = EqClasses(M, off, tids, w, counts)
    eff = rng.uniform(0.5, 3000.0, size=M)
    cnt_e = np.repeat(counts.astype(np.float64), sizes)
    proj = np.bincount(tids, weights=cnt_e * w, minlength=M)
    uniq = np.bincount(tids[np.repeat(sizes == 1, sizes)], weights=counts[sizes == 1].astype(float),
                       minlength=M).astype(np.uint64)
    return eq, proj, eff, uniq


def as_classes(eq):
    out = []
    for c in range(eq.n_classes):
        b, e = int(eq.off[c]), int(eq.off[c + 1])
        out.append((eq.tids[b:e].tolist(), eq.weights[b:e].tolist(), int(eq.counts[c])))
    return out


def test_digamma_golden(oracle):
    import mpmath
    mpmath.mp.dps = 40
    rows = json.load(open(os.path.join(HERE, "golden", "digamma_golden.json")))["rows"]
    assert len(rows) > 400
    for r in rows:
        x = float.fromhex(r["x"])
        want = mpmath.mpf(r["digamma"])
        got = oracle.digamma(x)
        if abs(x - 1.4616321449683623) < 0.05:   # near the root: absolute
            assert abs(got - want) < 4e-16
        else:
            assert abs((got - want) / want) < 2e-15, (x, got, want)


def test_philox_kat(oracle):
    # Known-answer vectors of Philox-4x32-10 (Random123 kat_vectors)
    assert oracle.philox((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert oracle.philox((f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_tiny_hand_case(oracle):
    # Two transcripts, one ambiguous class with equal weights + one unique class.
    # EM fixed point: alpha0 = 10 + x, alpha1 = 10 - x ... checked against closed form:
    # classes: {0}:count 10 ; {0,1}: count 10, weights .5/.5, equal effLen.
    # plain EM: a0' = 10 + 10*a0/(a0+a1), a1' = 10*a1/(a0+a1)  -> converges to (20, 0).
    eq = EqClasses(2, np.array([0, 1, 3]), np.array([0, 0, 1]), np.array([1.0, 0.5, 0.5]), np.array([10, 10]))
    p = mkparams(use_vbem=0, min_iter=2000, max_iter=2000)
    a, st = oracle.em_optimize(eq, np.array([15.0, 5.0]), np.array([100.0, 100.0]),
                               np.array([10, 0], dtype=np.uint64), p)
    assert st.iters == 2000 and st.ok == 1
    assert abs(a.sum() - 20.0) < 1e-9
    assert a[0] > 19.9 and a[1] < 0.1
    # one-step closed form (use min_iter=max_iter=1): includes the +1.0 the reference's
    # EM path carries on its first iteration (alphasPrime starts at 1.0, :812/:821).
    p1 = mkparams(use_vbem=0, min_iter=1, max_iter=1)
    a1, _ = oracle.em_optimize(eq, np.array([15.0, 5.0]), np.array([100.0, 100.0]),
                               np.array([10, 0], dtype=np.uint64), p1)
    tw = 20.0
    frac = min(0.999, tw / 5e7)
    a0 = [15.0 * frac + (tw / 2) * (1 - frac), 5.0 * frac + (tw / 2) * (1 - frac)]
    want0 = 1.0 + 10 + 10 * a0[0] / (a0[0] + a0[1])
    want1 = 1.0 + 10 * a0[1] / (a0[0] + a0[1])
    assert abs(a1[0] - want0) < 1e-12 and abs(a1[1] - want1) < 1e-12


@pytest.mark.parametrize("kw", [
    dict(), dict(use_vbem=0), dict(per_txp_prior=0, vb_prior=1e-5), dict(init_uniform=1),
    dict(eq_class_mode=1, init_uniform=1), dict(no_rich_eq=1), dict(alt_init=1),
])
def test_oracle_vs_python_restatement(oracle, kw):
    import py_ref
    rng = np.random.default_rng(5)
    eq, proj, eff, uniq = random_problem(rng, C=60, M=25, zero_frac=0.05)
    p = mkparams(min_iter=30, max_iter=200, **kw)
    a, st, cw, valid = oracle.em_optimize(eq, proj, eff, uniq, p, want_combined=True)
    pk = {k: p[k] for k in ("use_vbem", "per_txp_prior", "init_uniform", "eq_class_mode", "no_rich_eq",
                            "alt_init", "vb_prior", "tol", "num_required_frags", "min_iter", "max_iter")}
    ra, rit, rconv, rmax, rcomb, rvalid = py_ref.optimize(as_classes(eq), eq.n_txps, proj, eff, uniq, **pk)
    assert st.iters == rit
    assert bool(st.converged) == rconv
    assert [bool(v) for v in valid] == rvalid
    np.testing.assert_allclose(cw, np.concatenate(rcomb), rtol=1e-15, atol=0)
    np.testing.assert_allclose(a, np.array(ra), rtol=1e-10, atol=1e-12)
    assert abs(st.max_rel_diff - rmax) <= 1e-9 * max(1.0, abs(rmax))


def test_mass_conservation_and_mt(oracle):
    rng = np.random.default_rng(11)
    eq, proj, eff, uniq = random_problem(rng, C=5000, M=800, max_label=12)
    p = mkparams(min_iter=50, max_iter=50)
    a, st = oracle.em_optimize(eq, proj, eff, uniq, p)
    # every valid class hands out exactly its count
    assert st.n_degenerate == 0
    assert abs(a.sum() - float(eq.counts.sum())) / float(eq.counts.sum()) < 1e-9
    a2, st2 = oracle.em_optimize(eq, proj, eff, uniq, p, mt=True, n_threads=4)
    np.testing.assert_allclose(a2, a, rtol=1e-9, atol=1e-9)
    t = oracle.tpm(a, eff)
    assert abs(t.sum() - 1e6) < 1e-3


def test_degenerate_classes_are_dropped(oracle):
    # a class whose only members have zero combined weight is marked degenerate
    eq = EqClasses(3, np.array([0, 2, 3]), np.array([0, 1, 2]), np.array([0.0, 0.0, 1.0]), np.array([7, 5]))
    p = mkparams(min_iter=5, max_iter=5)
    a, st = oracle.em_optimize(eq, np.array([1.0, 1.0, 5.0]), np.array([100.0, 100.0, 100.0]),
                               np.zeros(3, dtype=np.uint64), p)
    assert st.n_degenerate == 1
    assert a[2] == 5.0 and a[0] == 0.0 and a[1] == 0.0



def test_online_oracle_properties():
    """Stage A online phase of the oracle: every assigned fragment adds exactly its forgetting mass (spread over its
    alignments), clusters keep their hit counts through normalizeAlphas, effective lengths stay in (1, length]."""
    import oracle_lib as O
    from salmon_b200.synth import synth_reads, synth_txome
    txps, _ = synth_txome(seed=5, n_genes=40)
    left, right, _ = synth_reads(txps, seed=8, n=2400)
    p = O.map_params(num_pre_burnin=500, num_burnin=1500)
    oix = O.MapIndex(txps)
    on = O.Online(oix, p, seed=3, mini_batch=400)
    fm = [0.0]
    for j in range(1, 12):
        fm.append(fm[-1] + 0.65 * np.log(j) - np.log((j + 1) ** 0.65 - 1))
    expect, parts = 0.0, []
    for b in range(3):
        sl = slice(800 * b, 800 * (b + 1))
        m = on.batch(left[sl], right[sl])
        parts.append(m)
        mapped = m["n_aln"] > 0
        step = 2 * b + np.arange(800) // 400
        expect += np.exp(np.array(fm)[step][mapped]).sum()
    st = on.state()
    assert st["timestep"] == 6 and st["frags_seen"] == 2400 and st["burned_in"] == 1
    tot_mass = np.exp(st["mass"][np.isfinite(st["mass"])]).sum()
    assert abs(tot_mass - expect) < 1e-6 * expect
    assert abs(np.exp(st["hist"] - st["tot"]).sum() - 1.0) < 1e-9      # the FLD stays a distribution
    cap = p.max_read_occ
    merged = {k: np.concatenate([q[k] for q in parts]) for k in ("n_aln", "label", "weight")}
    e = O.eq_aggregate(merged, cap, True)
    fin = on.finish(e["off"], e["tids"], e["counts"])
    assert abs(fin["projected_counts"].sum() - st["assigned"]) < 1e-6 * st["assigned"]
    assert np.all(fin["projected_counts"] <= fin["total_counts"] + 1e-6)
    assert np.all(fin["projected_counts"] >= fin["unique_counts"] - 1e-6)
    lens = np.array([t.shape[0] for t in txps], dtype=np.float64)
    assert np.all(fin["eff_len"] >= 1.0) and np.all(fin["eff_len"] <= lens)
    assert fin["eff_len"].mean() < lens.mean() - 100          # roughly length - mean fragment length


def test_map_oracle_vs_python_restatement_of_the_aux_model(oracle):
    """Second, independent restatement (tests/py_ref_map.py, written from the reference source) of
    filterAndCollectAlignments + the auxiliary-probability / label arithmetic of processMiniBatch, checked against the
    oracle read by read: paired alignments, left / right orphans, multi-mapping reads with range-factorised labels."""
    import py_ref_map as R
    from salmon_b200.synth import synth_reads, synth_txome
    txps, _ = synth_txome(seed=12, n_genes=120)
    left, right, truth = synth_reads(txps, seed=13, n=4000, indel_rate=0.002, sub_rate=0.02)
    rng = np.random.default_rng(2)
    kill = rng.random(4000) < 0.15                      # make orphans: one mate becomes unmappable noise
    right[kill] = rng.integers(0, 4, size=(int(kill.sum()), right.shape[1]), dtype=np.uint8)
    kill2 = rng.random(4000) < 0.08
    left[kill2 & ~kill] = rng.integers(0, 4, size=(int((kill2 & ~kill).sum()), left.shape[1]), dtype=np.uint8)
    p = oracle.map_params()
    m = oracle.map_reads(oracle.MapIndex(txps), p, left, right, 0)
    tx_len = np.array([t.shape[0] for t in txps])
    cmf = R.quirk_cmf(p.max_frag_len)
    cap = p.max_read_occ
    n_checked = n_multi = n_orphan = 0
    for r in range(left.shape[0]):
        k = R.check_read(m["n_aln"][r], m["tid"][r], m["score"][r], m["prob"][r], m["pos"][r], m["flags"][r],
                         m["label"][r], m["weight"][r], left.shape[1], tx_len, cmf, score_exp=p.score_exp,
                         min_aln_prob=p.min_aln_prob, range_bins=p.range_bins, max_read_occ=cap)
        n_checked += k > 0
        n_multi += k > 1
        n_orphan += k > 0 and ((int(m["flags"][r][0]) >> 2) & 3) != 0
    assert n_checked > 3000 and n_multi > 500 and n_orphan > 300, (n_checked, n_multi, n_orphan)


def test_normalize_alphas_vs_python_restatement(oracle):
    """normalizeAlphas + projectToPolytope restated in Python from the reference (tests/py_ref_map.py) against the
    oracle's orc_online_finish on the state three mapped batches leave behind."""
    import py_ref_map as R
    from salmon_b200.synth import synth_reads, synth_txome
    txps, _ = synth_txome(seed=14, n_genes=100)
    left, right, _ = synth_reads(txps, seed=15, n=6000)
    p = oracle.map_params(num_pre_burnin=1500, num_burnin=4000)
    on = oracle.Online(oracle.MapIndex(txps), p, seed=3, mini_batch=1000)
    outs = [on.batch(left[s:s + 2000], right[s:s + 2000]) for s in range(0, 6000, 2000)]
    cap = p.max_read_occ
    # classes over all batches (transcript part of the labels)
    agg = {}
    for m in outs:
        for r in range(m["n_aln"].shape[0]):
            k = int(m["n_aln"][r])
            if k:
                key = tuple(int(x) for x in m["tid"][r, :k])
                agg[key] = agg.get(key, 0) + 1
    keys = sorted(agg)
    off = np.concatenate([[0], np.cumsum([len(k) for k in keys])]).astype(np.uint64)
    tids = np.array([t for k in keys for t in k], dtype=np.uint32)
    counts = np.array([agg[k] for k in keys], dtype=np.uint64)
    st = on.state()
    fin = on.finish(off, tids, counts)
    proj, uniq, total = R.normalize_alphas([float(x) for x in st["mass"]], [(list(k), agg[k]) for k in keys])
    assert np.array_equal(fin["unique_counts"], np.array(uniq, dtype=np.uint64))
    assert np.array_equal(fin["total_counts"], np.array(total, dtype=np.uint64))
    np.testing.assert_allclose(fin["projected_counts"], np.array(proj), rtol=1e-9, atol=1e-9)
    assert abs(fin["projected_counts"].sum() - counts.sum()) < 1e-6 * counts.sum()
    assert (np.array(proj) > 0).sum() > 50


def test_fld_prior_tables_vs_python_restatement(oracle):
    """The fragment-length prior (FragmentLengthDistribution ctor, pmf, cmf) restated in Python against orc_fld_tables:
    where the normal mass is resolvable (within ~7 sd) to 1e-9 in log space; farther out both carry the rounding noise
    of a cdf difference near 1 (as the reference's own tables do), so only the floor is compared there."""
    import ctypes as C
    import py_ref_map as R
    lib = oracle.load()
    for mean, sd, mx in ((250.0, 25.0, 1000), (180.0, 15.0, 800), (400.0, 80.0, 1000)):
        pmf = np.zeros(mx + 1); cmf = np.zeros(mx + 1)
        lib.orc_fld_tables(C.c_double(mean), C.c_double(sd), C.c_uint32(mx), pmf.ctypes.data_as(C.c_void_p),
                           cmf.ctypes.data_as(C.c_void_p))
        rp, rc = R.fld_prior_tables(mean, sd, mx)
        rp, rc = np.array(rp), np.array(rc)
        core = np.abs(np.arange(mx + 1) - mean) < 7 * sd
        np.testing.assert_allclose(pmf[core], rp[core], rtol=0, atol=1e-9)
        np.testing.assert_allclose(cmf[core], rc[core], rtol=0, atol=1e-9)
        assert abs(np.exp(cmf[-1])) - 1 < 1e-9 and np.all(np.diff(cmf) >= -1e-12)     # a distribution, monotone
        far = np.arange(mx + 1) > mean + 12 * sd          # the normal mass underflows: the LOG_EPSILON floor, exactly
        if far.any():
            np.testing.assert_allclose(pmf[far], rp[far], rtol=0, atol=1e-12)


def test_eq_aggregate_vs_python_dict(oracle):
    """EquivalenceClassBuilder::addGroup / finish (EquivalenceClassBuilder.hpp:165-181,237-250; TGValue :114-123)
    restated with a Python dict: count += 1 and weights[i] += w_i per fragment under the full label (transcripts ++ range
    bins), finish() scales a class's weights to sum 1."""
    from salmon_b200.synth import synth_reads, synth_txome
    txps, _ = synth_txome(seed=16, n_genes=80)
    left, right, _ = synth_reads(txps, seed=17, n=3000)
    p = oracle.map_params()
    m = oracle.map_reads(oracle.MapIndex(txps), p, left, right, 0)
    cap = p.max_read_occ
    agg = {}
    for r in range(3000):
        k = int(m["n_aln"][r])
        if not k:
            continue
        key = tuple(int(x) for x in m["label"][r, :2 * k])
        cnt, ws = agg.get(key, (0, [0.0] * k))
        agg[key] = (cnt + 1, [a + float(b) for a, b in zip(ws, m["weight"][r, :k])])
    e = oracle.eq_aggregate(m, cap, True)
    off = e["off"].astype(np.int64)
    assert len(e["counts"]) == len(agg)
    for c, key in enumerate(sorted(agg)):                         # the oracle emits classes sorted by full label
        cnt, ws = agg[key]
        k = len(ws)
        assert e["tids"][off[c]:off[c + 1]].tolist() == list(key[:k]) and int(e["counts"][c]) == cnt
        tot = sum(ws)
        np.testing.assert_allclose(e["weights"][off[c]:off[c + 1]], [w / tot for w in ws], rtol=1e-12)


def test_bootstrap_and_gibbs_oracle_statistics(oracle):
    """The samplers' RNG is substituted (Philox for mt19937 / pcg32), so the oracle is pinned on what does not depend on
    the generator: a bootstrap replicate is a multinomial resample of the class counts (every replicate draws exactly N
    fragments, replicate means and variances per class are those of Multinomial(N, count/N),
    CollapsedEMOptimizer.cpp:398-552), and the Gibbs chain conserves the fragments and centres on the EM estimate
    (CollapsedGibbsSampler.cpp:317-508)."""
    from salmon_b200._capi import default_params
    from salmon_b200.synth import synth_eq
    from test_sampling_gpu import boot_inputs, unique_labels
    eq, proj, eff, uniq = synth_eq(seed=21, C=800, M=300, total_count=40_000)
    eq = unique_labels(eq, proj, eff, uniq)
    N = float(eq.counts.sum())
    p_main = default_params(use_vbem=1)
    alpha, st = oracle.em_optimize(eq, proj, eff, uniq, p_main)
    cw, valid, valid_boot, prior, active = boot_inputs(oracle, eq, proj, eff, uniq, p_main, N)
    p_boot = default_params(use_vbem=1, min_iter=50)
    n_boot = 120
    alphas, samp, rc = oracle.bootstrap(eq, cw, valid_boot, prior, active, p_boot, n_boot, 11)
    assert rc == 0
    assert np.all(samp.sum(axis=1) == int(N))                       # N draws per replicate
    np.testing.assert_allclose(alphas.sum(axis=1), N, rtol=1e-6)
    pc = eq.counts / N
    mean, var = samp.mean(axis=0), samp.var(axis=0, ddof=1)
    big = eq.counts >= 50
    z = (mean[big] - eq.counts[big]) / np.sqrt(N * pc[big] * (1 - pc[big]) / n_boot)
    assert np.abs(z).max() < 5 and abs(z.mean()) < 0.5, (np.abs(z).max(), z.mean())
    ratio = var[big] / (N * pc[big] * (1 - pc[big]))
    assert 0.8 < np.median(ratio) < 1.25, np.median(ratio)
    # replicates differ from each other and scatter around the point estimate
    assert np.abs(alphas - alphas[0]).max() > 1.0
    hi = alpha > 200
    assert np.all(np.abs(alphas[:, hi].mean(axis=0) - alpha[hi]) < 0.15 * alpha[hi] + 30)
    # Gibbs: thinning 4, with and without the gamma draw
    for no_gamma in (1, 0):
        g = oracle.gibbs(eq, valid, eff, alpha, 1, 1, p_main.vb_prior, 40, 4, no_gamma, N, 5)
        np.testing.assert_allclose(g.sum(axis=1), N, rtol=1e-9)
        assert (g >= 0).all() and np.abs(g - g[0]).max() > 1.0
        assert np.all(np.abs(g[:, hi].mean(axis=0) - alpha[hi]) < 0.2 * alpha[hi] + 30)
