"""world_size-2 gloo test (CPU) of the N>1 host logic: classes shard round-robin
across ranks, each rank produces its partial alpha' for the same alpha, and the
all-reduced sum equals the unsharded update."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from salmon_b200.synth import shard_classes, synth_eq
    eq, proj, eff, uniq = synth_eq(seed=3, C=4000, M=900, total_count=100000)
    sh = shard_classes(eq, rank, world)
    # uid-style broadcast plumbing: rank 0 owns a 128-byte token, everyone must agree
    tok = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        tok = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(tok, 0)
    assert tok.tolist() == list(range(128))
    rng = np.random.default_rng(0)
    alpha = rng.random(eq.n_txps) * 10
    prior = np.full(eq.n_txps, 1e-2)
    sizes = (sh.off[1:] - sh.off[:-1]).astype(np.int64)
    cw = sh.weights  # already normalised per class
    valid = np.ones(sh.n_classes, dtype=np.uint8)
    # EM step on the shard (theta = alpha): partial alpha'
    part, _ = O.em_step(sh, cw, valid, prior, alpha, vbem=0)
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    full, _ = O.em_step(eq, eq.weights, np.ones(eq.n_classes, dtype=np.uint8), prior, alpha, vbem=0)
    ok = np.allclose(t.numpy(), full, rtol=1e-12, atol=1e-12)
    cnt = torch.tensor([float(sh.counts.sum())], dtype=torch.float64)
    dist.all_reduce(cnt)
    ok = ok and cnt.item() == float(eq.counts.sum())
    q.put((rank, bool(ok), int(sizes.sum())))
    dist.destroy_process_group()


def test_sharded_update_equals_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res


def _make_partial(rank, M=500, nf=101):
    rng = np.random.default_rng(100 + rank)
    prior = np.log(np.full(nf, 1.0 / nf)) + 0.0
    contrib = rng.random(nf) * (rng.random(nf) < 0.4)                 # this rank's FLD additions (linear)
    hist = np.log(np.exp(prior) + contrib)
    mass = np.where(rng.random(M) < 0.6, np.log(rng.random(M) * 50 + 1e-3), np.inf)   # +inf = LOG_0
    return dict(mass=mass, fld_hist=hist, fld_tot=float(np.log(np.exp(hist).sum())), fld_prior_hist=prior,
                fld_prior_tot=float(np.log(np.exp(prior).sum())), fld_min=int(40 + 7 * rank),
                unique_counts=rng.integers(0, 50, M).astype(np.uint64), total_counts=rng.integers(50, 90, M).astype(np.uint64),
                cluster_hits=rng.integers(0, 30, M).astype(np.uint64),
                cluster_root=(np.arange(M) // (3 + rank) * (3 + rank)).astype(np.uint32), assigned=1000 + rank), contrib


def _worker_partials(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from salmon_b200.dist import reduce_partials
    p, _ = _make_partial(rank)
    g, roots = reduce_partials(p, dist, "cpu")
    parts = [_make_partial(r) for r in range(world)]
    lin = sum(np.where(np.isfinite(pp["mass"]), np.exp(pp["mass"]), 0.0) for pp, _ in parts)
    ok = np.allclose(np.where(np.isfinite(g["mass"]), np.exp(g["mass"]), 0.0), lin, rtol=1e-12)
    ok &= bool(np.all(np.isinf(g["mass"]) == (lin == 0)))
    hist_lin = np.exp(parts[0][0]["fld_prior_hist"]) + sum(c for _, c in parts)
    ok &= np.allclose(np.exp(g["fld_hist"]), hist_lin, rtol=1e-10)
    ok &= abs(np.exp(g["fld_tot"]) - hist_lin.sum()) < 1e-9 * hist_lin.sum()
    for k in ("unique_counts", "total_counts", "cluster_hits"):
        ok &= np.array_equal(g[k], sum(pp[k] for pp, _ in parts))
    ok &= g["fld_min"] == 40 and g["assigned"] == sum(1000 + r for r in range(world))
    ok &= roots.shape == (world, 500) and all(np.array_equal(roots[r], parts[r][0]["cluster_root"]) for r in range(world))
    q.put((rank, bool(ok), 0))
    dist.destroy_process_group()


def test_end_of_mapping_reductions_gloo():
    """Stage A, N > 1: the once-per-run reduction of masses / FLD / counts / cluster roots (salmon_b200.dist)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_partials, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
