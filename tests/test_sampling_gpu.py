"""GPU parity of the bootstrap and Gibbs samplers against the oracle (same counter-RNG
streams): resampled class counts and no-gamma Gibbs counts are bit-exact; bootstrap alphas
within 1e-9; Gibbs with gamma draws statistically (means over samples)."""
import numpy as np
import pytest

from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = EMContext(0)
    yield c
    c.close()


def unique_labels(eq, proj, eff, uniq):
    """Drop duplicate labels (several classes on the same transcript set) so that per-class
    sampled counts are well defined for the comparison."""
    seen, keep = set(), []
    for c in range(eq.n_classes):
        k = eq.tids[int(eq.off[c]):int(eq.off[c + 1])].tobytes()
        if k not in seen:
            seen.add(k); keep.append(c)
    keep = np.array(keep)
    from salmon_b200._capi import EqClasses
    sizes = (eq.off[1:] - eq.off[:-1]).astype(np.int64)[keep]
    off = np.concatenate(([0], np.cumsum(sizes)))
    idx = np.concatenate([np.arange(int(eq.off[c]), int(eq.off[c + 1])) for c in keep])
    return EqClasses(eq.n_txps, off, eq.tids[idx], eq.weights[idx], eq.counts[keep])


def boot_inputs(oracle, eq, proj, eff, uniq, p_main, nmapped):
    """valid flags + priors exactly as gatherBootstraps sees them."""
    _, _, cw, valid = oracle.em_optimize(eq, proj, eff, uniq, p_main, want_combined=True)
    M = eq.n_txps
    active = np.zeros(M, dtype=np.uint8)
    active[eq.tids] = 1
    unif = (1.0 / active.sum()) * nmapped
    sizes = (eq.off[1:] - eq.off[:-1]).astype(np.int64)
    v = np.where(active[eq.tids] > 0, unif, 0.0) * cw
    v[np.isnan(v)] = 0.0
    denom = np.add.reduceat(v, eq.off[:-1].astype(np.int64))
    denom[sizes == 0] = 0.0
    valid_boot = (valid.astype(bool) & (denom > np.finfo(float).tiny)).astype(np.uint8)
    prior = np.full(M, p_main.vb_prior) if p_main.per_txp_prior else p_main.vb_prior * eff
    return cw, valid, valid_boot, prior, active


@pytest.mark.parametrize("vbem", [1, 0])
def test_bootstrap_parity(ctx, oracle, vbem):
    eq, proj, eff, uniq = synth_eq(seed=12, C=6000, M=1500, total_count=150_000)
    eq = unique_labels(eq, proj, eff, uniq)
    nmapped = float(eq.counts.sum())
    p_main = default_params(use_vbem=vbem)
    alpha, st, ok = ctx.optimize(eq, p_main, proj, eff, uniq)
    cw, valid, valid_boot, prior, active = boot_inputs(oracle, eq, proj, eff, uniq, p_main, nmapped)
    p_boot = default_params(use_vbem=vbem, min_iter=50, max_iter=10000)
    n_boot, seed = 3, 0x1234ABCD5678
    got, okb = ctx.bootstrap(p_boot, nmapped, n_boot, seed)
    counts_last = ctx.bootstrap_last_counts()
    ref, samp, rc = oracle.bootstrap(eq, cw, valid_boot, prior, active, p_boot, n_boot, seed)
    assert okb and rc == 0 and got.shape == ref.shape
    assert np.array_equal(counts_last, samp[-1])                     # bit-exact resampled counts
    assert int(counts_last.sum()) == int(eq.counts[valid_boot > 0].sum())
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_gibbs_no_gamma_bit_exact_counts(ctx, oracle):
    eq, proj, eff, uniq = synth_eq(seed=13, C=5000, M=1200, total_count=80_000)
    p_main = default_params()
    alpha, st, ok = ctx.optimize(eq, p_main, proj, eff, uniq)
    _, _, cw, valid = oracle.em_optimize(eq, proj, eff, uniq, p_main, want_combined=True)
    nmapped = float(eq.counts.sum())
    got = ctx.gibbs(alpha, 1, 1, 1e-2, n_samples=6, thinning=4, no_gamma_draw=1, num_mapped_frags=nmapped, seed=99)
    ref = oracle.gibbs(eq, valid, eff, alpha, 1, 1, 1e-2, 6, 4, 1, nmapped, 99)
    assert got.shape == ref.shape == (6, eq.n_txps)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    assert abs(got[-1].sum() - nmapped) / nmapped < 1e-9


def test_gibbs_gamma_posterior_mean(ctx, oracle):
    """config 5 criterion: posterior-mean TPM of the GPU chain within 1e-3 (absolute, in
    TPM/1e6 units = fraction) of the CPU chain driven by the same streams."""
    eq, proj, eff, uniq = synth_eq(seed=14, C=4000, M=800, total_count=200_000)
    p_main = default_params()
    alpha, st, ok = ctx.optimize(eq, p_main, proj, eff, uniq)
    _, _, cw, valid = oracle.em_optimize(eq, proj, eff, uniq, p_main, want_combined=True)
    nmapped = float(eq.counts.sum())
    got = ctx.gibbs(alpha, 1, 1, 1e-2, n_samples=40, thinning=4, no_gamma_draw=0, num_mapped_frags=nmapped, seed=7)
    ref = oracle.gibbs(eq, valid, eff, alpha, 1, 1, 1e-2, 40, 4, 0, nmapped, 7)
    tg = np.array([oracle.tpm(s, eff) for s in got]).mean(0) / 1e6
    tr = np.array([oracle.tpm(s, eff) for s in ref]).mean(0) / 1e6
    assert np.max(np.abs(tg - tr)) < 1e-3
    assert np.isfinite(got).all() and (got >= 0).all()


@pytest.mark.parametrize("vbem", [1, 0])
def test_optimize_after_bootstrap_on_a_larger_problem(ctx, oracle, vbem):
    """the bootstrap driver's override buffers outlive it; a later optimize on the same context --
    here on a LARGER table -- must not see the last replicate's resampled counts / base / uniform start."""
    eq, proj, eff, uniq = synth_eq(seed=21, C=3000, M=900, total_count=90_000)
    p = default_params(use_vbem=vbem, min_iter=30, max_iter=30)
    ctx.optimize(eq, p, proj, eff, uniq)
    got, okb = ctx.bootstrap(default_params(use_vbem=vbem, min_iter=50, max_iter=200), float(eq.counts.sum()), 2, 77)
    assert okb
    eq2, proj2, eff2, uniq2 = synth_eq(seed=22, C=9000, M=2500, total_count=400_000)
    alpha, st, ok = ctx.optimize(eq2, p, proj2, eff2, uniq2)
    ref, rst = oracle.em_optimize(eq2, proj2, eff2, uniq2, p)
    assert ok and st.iters == rst.iters == 30
    np.testing.assert_allclose(alpha, ref, rtol=1e-9, atol=1e-9)
    # and the staged path (run without a new prepare) on the same context
    ctx.upload(eq2, proj2, eff2, uniq2); ctx.prepare(p); ctx.run()
    a2, _, ok2 = ctx.download()
    np.testing.assert_allclose(a2, ref, rtol=1e-9, atol=1e-9)
