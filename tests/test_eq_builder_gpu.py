"""B2 seam (sb_eq_create / add_batch / from_host / finish) against the oracle's class aggregation
(orc_eq_aggregate = EquivalenceClassBuilder::addGroup + finish, EquivalenceClassBuilder.hpp:165-181,237-250) and a
Python dict restatement; labels, counts bit-exact, weights 1e-12 (sums over batches in batch order)."""
import numpy as np
import pytest

import oracle_lib as O
from salmon_b200._capi import EqBuilder, EqClasses, EMContext
from salmon_b200 import default_params
from salmon_b200.synth import synth_eq

pytestmark = pytest.mark.gpu


def random_groups(rng, n, M, cap, binned):
    """n addGroup calls: sorted transcript ids, softmax-like weights, bins = int(w * (sqrt(k) + 4)) when binned"""
    n_aln = rng.integers(1, cap + 1, n).astype(np.uint32)
    label = np.zeros((n, 2 * cap), dtype=np.uint32); weight = np.zeros((n, cap))
    pool = rng.integers(0, M, (200, cap))       # few distinct transcript sets, so that labels repeat
    for i in range(n):
        k = int(n_aln[i])
        t = np.unique(pool[rng.integers(0, 200)][:k])
        k = len(t); n_aln[i] = k
        w = rng.dirichlet(np.ones(k) * 0.7) if rng.random() < 0.5 else np.full(k, 1.0 / k)
        label[i, :k] = t
        weight[i, :k] = w
        if binned:
            label[i, k:2 * k] = (w * (int(np.sqrt(k)) + 4)).astype(np.uint32)
    return n_aln, label, weight


def to_csr(n_aln, label, weight, binned):
    ll = n_aln.astype(np.uint64) * (2 if binned else 1)
    loff = np.concatenate(([0], np.cumsum(ll))).astype(np.uint64)
    woff = np.concatenate(([0], np.cumsum(n_aln.astype(np.uint64)))).astype(np.uint64)
    labs = np.concatenate([label[i, :int(ll[i])] for i in range(len(n_aln))]) if len(n_aln) else np.zeros(0, np.uint32)
    ws = np.concatenate([weight[i, :int(n_aln[i])] for i in range(len(n_aln))]) if len(n_aln) else np.zeros(0)
    return loff, labs.astype(np.uint32), woff, ws


@pytest.mark.parametrize("binned", [True, False])
def test_add_batch_and_finish_vs_oracle(binned):
    rng = np.random.default_rng(31 + binned)
    M, cap, n = 500, 12, 20000
    n_aln, label, weight = random_groups(rng, n, M, cap, binned)
    b = EqBuilder(M)
    try:
        for s in range(0, n, 6000):          # several addGroup batches
            b.add_batch(*to_csr(n_aln[s:s + 6000], label[s:s + 6000], weight[s:s + 6000], binned))
        got = b.finish()
    finally:
        b.close()
    want = O.eq_aggregate(dict(n_aln=n_aln, label=label, weight=weight), cap, binned)
    assert got["n_groups"] == n and int(got["counts"].sum()) == n
    # canonical order for the comparison: by full label
    def keyed(off, ntx, labs_of, counts, w):
        d = {}
        for c in range(len(counts)):
            d[labs_of(c)] = (int(counts[c]), w[int(off[c]):int(off[c + 1])])
        return d
    g = keyed(got["off"], got["ntx"], lambda c: got["labels"][int(got["label_off"][c]):int(got["label_off"][c + 1])].tobytes(),
              got["counts"], got["weights"])
    # the oracle returns the transcript part and (binned) the classes keyed by the full label in label order
    full = {}
    for i in range(n):
        k = int(n_aln[i]); key = label[i, :k * (2 if binned else 1)].tobytes()
        cnt, ws = full.get(key, (0, np.zeros(k)))
        full[key] = (cnt + 1, ws + weight[i, :k])
    assert set(g) == set(full) and len(g) == len(want["counts"])
    for key, (cnt, ws) in full.items():
        assert g[key][0] == cnt
        np.testing.assert_allclose(g[key][1], ws / ws.sum(), rtol=1e-12, atol=0)
    assert sorted(int(x) for x in got["counts"]) == sorted(int(x) for x in want["counts"])


def test_from_host_feeds_the_optimiser_like_the_table_itself(oracle):
    """--eqclasses path: a finished table (distinct labels) through sb_eq_from_host -> sb_eq_finish -> sb_em_optimize
    gives what the table itself gives"""
    from test_sampling_gpu import unique_labels
    eq, proj, eff, uniq = synth_eq(seed=41, C=8000, M=2000, total_count=300000)
    eq = unique_labels(eq, proj, eff, uniq)
    b = EqBuilder(eq.n_txps)
    try:
        b.from_host(eq)
        t = b.finish()
    finally:
        b.close()
    assert len(t["counts"]) == eq.n_classes and int(t["counts"].sum()) == int(eq.counts.sum())
    eq2 = EqClasses(eq.n_txps, t["off"], t["tids"], t["weights"], t["counts"])
    p = default_params(min_iter=20, max_iter=20)
    ctx = EMContext(0)
    try:
        a2, _, ok2 = ctx.optimize(eq2, p, proj, eff, uniq)
    finally:
        ctx.close()
    ref, _ = oracle.em_optimize(eq, proj, eff, uniq, p)
    np.testing.assert_allclose(a2, ref, rtol=1e-9, atol=1e-9)      # class order differs, the classes do not


def test_argument_checks():
    b = EqBuilder(10)
    try:
        with pytest.raises(Exception, match="out of range"):
            b.add_batch([0, 2], [3, 11], [0, 2], [0.5, 0.5])
        with pytest.raises(Exception, match="weights"):
            b.add_batch([0, 3], [1, 2, 3], [0, 2], [0.5, 0.5])
        assert b.finish()["off"].tolist() == [0]
    finally:
        b.close()
