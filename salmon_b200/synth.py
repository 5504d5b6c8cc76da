"""Seeded synthetic workloads.  No real data exists on the box.

synth_eq(seed, C, M): equivalence classes shaped like a human GENCODE run:
label sizes 1 + Geometric (mean ~6, capped at 200 = maxReadOcc), members drawn from
gene-family blocks of neighbouring transcript ids, counts Zipf(1.3) scaled to the
requested total, weights Dirichlet(1) per class (normalised, as finish() leaves them).
"""
from __future__ import annotations

import numpy as np

from ._capi import EqClasses


def synth_eq(seed=1, C=500_000, M=250_000, total_count=20_000_000, mean_extra=5.0, max_label=200,
             expressed_frac=0.4):
    rng = np.random.default_rng(seed)
    # label sizes: 1 + Geometric(failures), mean 1 + mean_extra
    p = 1.0 / (1.0 + mean_extra)
    sizes = 1 + (rng.geometric(p, size=C) - 1)
    sizes = np.minimum(sizes, min(max_label, M)).astype(np.int64)
    # expression: log-normal over the expressed subset
    expr = np.zeros(M)
    n_expr = max(1, int(M * expressed_frac))
    expressed = rng.choice(M, size=n_expr, replace=False)
    expr[expressed] = rng.lognormal(0.0, 2.0, size=n_expr)
    cdf = np.cumsum(expr)
    anchors = np.searchsorted(cdf, rng.random(C) * cdf[-1], side="right").clip(0, M - 1)
    # window = gene-family block around the anchor, at least 2x the label size
    win = np.maximum(16, 1 << np.ceil(np.log2(np.maximum(2 * sizes, 2))).astype(np.int64))
    win = np.minimum(win, M)
    start = (anchors // 16) * 16
    start = np.minimum(start, M - win)
    # choose `size` distinct offsets in the window: random keys, keep the smallest
    tot = int(win.sum())
    cls_of = np.repeat(np.arange(C), win)
    woff = np.concatenate(([0], np.cumsum(win)))
    pos = np.arange(tot) - np.repeat(woff[:-1], win)
    keys = rng.random(tot)
    # the anchor itself is always a member
    keys[woff[:-1] + (anchors - start)] = -1.0
    order = np.lexsort((keys, cls_of))
    rank = np.arange(tot) - np.repeat(woff[:-1], win)
    keep = rank < np.repeat(sizes, win)
    sel = order[keep]
    tids = (np.repeat(start, win)[sel] + pos[sel]).astype(np.uint32)
    cls_sel = cls_of[sel]
    # sort members ascending inside each class (TranscriptGroup keeps txps sorted)
    o2 = np.lexsort((tids, cls_sel))
    tids = tids[o2]
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    nnz = int(off[-1])
    # weights ~ Dirichlet(1) per class
    e = rng.exponential(1.0, size=nnz)
    s = np.add.reduceat(e, off[:-1].astype(np.int64))
    weights = e / np.repeat(s, sizes)
    # counts ~ Zipf(1.3), scaled
    raw = np.minimum(rng.zipf(1.3, size=C).astype(np.float64), 1e6)
    counts = np.maximum(1, np.rint(raw * (total_count / raw.sum()))).astype(np.uint64)
    eq = EqClasses(M, off, tids, weights, counts)
    # per-transcript inputs
    lens = np.maximum(250.0, rng.lognormal(np.log(1500.0), 0.8, size=M))
    eff_len = np.maximum(1.0, lens - 200.0)
    cnt_e = np.repeat(counts.astype(np.float64), sizes)
    projected = np.bincount(tids, weights=cnt_e * weights, minlength=M).astype(np.float64)
    single = sizes == 1
    unique = np.bincount(tids[np.repeat(single, sizes)], weights=counts[single].astype(np.float64),
                         minlength=M).astype(np.uint64)
    return eq, projected, eff_len, unique


def shard_classes(eq: EqClasses, rank: int, nranks: int) -> EqClasses:
    """Round-robin shard of the classes (stand-in for per-GPU eq-class tables)."""
    C = eq.n_classes
    sel = np.arange(rank, C, nranks)
    sizes = (eq.off[1:] - eq.off[:-1]).astype(np.int64)
    s = sizes[sel]
    off = np.concatenate(([0], np.cumsum(s))).astype(np.uint64)
    starts = eq.off[:-1].astype(np.int64)[sel]
    idx = np.repeat(starts - off[:-1].astype(np.int64), s) + np.arange(int(off[-1]))
    return EqClasses(eq.n_txps, off, eq.tids[idx], eq.weights[idx], eq.counts[sel])


# --------------------------------------------------------------------------------------
# Stage A: synthetic transcriptome and paired-end reads
# --------------------------------------------------------------------------------------
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)   # A<->T, C<->G in 2-bit codes (A=0,C=1,G=2,T=3)


def synth_txome(seed=44, n_genes=2000, max_isoforms=12, exon_mu=5.0, exon_sigma=0.8, min_len=300):
    """Genes own a pool of exons; isoforms are ordered subsets of the pool, so isoforms of a
    gene share >= k-mers (realistic multi-mapping).  Returns a list of uint8 arrays (2-bit codes)."""
    rng = np.random.default_rng(seed)
    txps = []
    gene_of = []
    for g in range(n_genes):
        n_ex = int(rng.integers(3, 16))
        exons = [rng.integers(0, 4, size=max(40, int(rng.lognormal(exon_mu, exon_sigma))), dtype=np.uint8)
                 for _ in range(n_ex)]
        n_iso = int(min(max_isoforms, 1 + rng.zipf(1.6)))
        for _ in range(n_iso):
            keep = rng.random(n_ex) < 0.7
            keep[rng.integers(0, n_ex)] = True
            seq = np.concatenate([e for e, k in zip(exons, keep) if k])
            if seq.shape[0] < min_len:
                seq = np.concatenate([seq, rng.integers(0, 4, size=min_len - seq.shape[0], dtype=np.uint8)])
            txps.append(seq)
            gene_of.append(g)
    return txps, np.array(gene_of)


def revcomp(codes):
    return _COMP[codes[::-1]]


def synth_reads(txps, seed=7, n=10000, read_len=100, frag_mean=250.0, frag_sd=25.0, sub_rate=0.005,
                indel_rate=0.0001, random_frac=0.03, expressed_frac=0.4):
    """IU (inward, unstranded) pairs.  Returns (left [n,L] uint8, right [n,L] uint8, truth dict)."""
    rng = np.random.default_rng(seed)
    N = len(txps)
    lens = np.array([t.shape[0] for t in txps])
    expr = np.zeros(N)
    ex = rng.choice(N, size=max(1, int(N * expressed_frac)), replace=False)
    expr[ex] = rng.lognormal(0.0, 2.0, size=ex.shape[0])
    w = expr * np.maximum(lens - frag_mean, 1.0)
    cdf = np.cumsum(w)
    left = np.zeros((n, read_len), dtype=np.uint8)
    right = np.zeros((n, read_len), dtype=np.uint8)
    t_tid = np.full(n, -1, dtype=np.int64)
    t_pos = np.zeros(n, dtype=np.int64)
    t_flen = np.zeros(n, dtype=np.int64)

    def mutate(r):
        r = r.copy()
        m = rng.random(r.shape[0]) < sub_rate
        r[m] = (r[m] + rng.integers(1, 4, size=int(m.sum()))) % 4
        if indel_rate > 0 and rng.random() < indel_rate * r.shape[0]:
            p = int(rng.integers(5, r.shape[0] - 5))
            if rng.random() < 0.5:   # deletion in the read: shift left, pad with random base
                r = np.concatenate([r[:p], r[p + 1:], rng.integers(0, 4, size=1, dtype=np.uint8)])
            else:                    # insertion
                r = np.concatenate([r[:p], rng.integers(0, 4, size=1, dtype=np.uint8), r[p:-1]])
        return r.astype(np.uint8)

    for i in range(n):
        if rng.random() < random_frac:
            left[i] = rng.integers(0, 4, size=read_len); right[i] = rng.integers(0, 4, size=read_len)
            continue
        t = int(np.searchsorted(cdf, rng.random() * cdf[-1], side="right"))
        t = min(t, N - 1)
        L = lens[t]
        fl = int(np.clip(round(rng.normal(frag_mean, frag_sd)), read_len, min(1000, L)))
        pos = int(rng.integers(0, L - fl + 1))
        frag = txps[t][pos:pos + fl]
        a = frag[:read_len]
        b = revcomp(frag[-read_len:])
        if rng.random() < 0.5:       # unstranded: fragment from the reverse strand
            a, b = b, a
        left[i] = mutate(a); right[i] = mutate(b)
        t_tid[i] = t; t_pos[i] = pos; t_flen[i] = fl
    return left, right, dict(tid=t_tid, pos=t_pos, flen=t_flen)


def flatten_txome(txps):
    lens = np.array([t.shape[0] for t in txps], dtype=np.uint64)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    codes = np.ascontiguousarray(np.concatenate(txps).astype(np.uint8)) if len(txps) else np.zeros(0, np.uint8)
    return off, codes


def synth_reads_fast(txps, seed=7, n=1_000_000, read_len=100, frag_mean=250.0, frag_sd=25.0, sub_rate=0.005,
                     indel_rate=0.0001, random_frac=0.03, expressed_frac=0.4, flat=None):
    """Vectorised synth_reads (same model, different random stream): IU pairs, substitutions everywhere,
    one indel in a fraction indel_rate*read_len of the mates.  Returns (left, right, truth)."""
    rng = np.random.default_rng(seed)
    off, codes = flat if flat is not None else flatten_txome(txps)
    N = off.shape[0] - 1
    lens = (off[1:] - off[:-1]).astype(np.int64)
    expr = np.zeros(N)
    ex = rng.choice(N, size=max(1, int(N * expressed_frac)), replace=False)
    expr[ex] = rng.lognormal(0.0, 2.0, size=ex.shape[0])
    w = expr * np.maximum(lens - frag_mean, 1.0)
    cdf = np.cumsum(w)
    t = np.minimum(np.searchsorted(cdf, rng.random(n) * cdf[-1], side="right"), N - 1)
    L = lens[t]
    fl = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=n)).astype(np.int64), read_len, np.minimum(1000, L))
    pos = (rng.random(n) * (L - fl + 1)).astype(np.int64)
    g0 = off[:-1].astype(np.int64)[t] + pos
    ar = np.arange(read_len, dtype=np.int64)
    a = codes[g0[:, None] + ar[None, :]]                                   # fragment start, forward
    b = _COMP[codes[(g0 + fl - 1)[:, None] - ar[None, :]]]                 # fragment end, reverse-complemented
    swap = rng.random(n) < 0.5
    left = np.where(swap[:, None], b, a)
    right = np.where(swap[:, None], a, b)
    for r in (left, right):
        m = rng.random(r.shape) < sub_rate
        r[m] = (r[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) % 4
        if indel_rate > 0:
            rows = np.nonzero(rng.random(n) < indel_rate * read_len)[0]
            for i in rows:
                p = int(rng.integers(5, read_len - 5))
                x = r[i]
                if rng.random() < 0.5:
                    r[i] = np.concatenate([x[:p], x[p + 1:], rng.integers(0, 4, size=1, dtype=np.uint8)])
                else:
                    r[i] = np.concatenate([x[:p], rng.integers(0, 4, size=1, dtype=np.uint8), x[p:-1]])
    rnd = rng.random(n) < random_frac
    nr = int(rnd.sum())
    left[rnd] = rng.integers(0, 4, size=(nr, read_len), dtype=np.uint8)
    right[rnd] = rng.integers(0, 4, size=(nr, read_len), dtype=np.uint8)
    t = np.where(rnd, -1, t)
    return np.ascontiguousarray(left, dtype=np.uint8), np.ascontiguousarray(right, dtype=np.uint8), \
        dict(tid=t, pos=pos, flen=fl)
