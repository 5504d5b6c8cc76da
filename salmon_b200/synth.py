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
