"""torchrun --nproc-per-node N: (1) EM: per-rank class shards, alpha all-reduced per iteration; rank 0 compares with
the oracle on the union of the shards.  (2) Stage A + quant: reads sharded over the ranks, end-of-mapping reductions
(salmon_b200.dist), sharded VBEM; rank 0 compares with a single-GPU run over all reads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
from salmon_b200 import EMContext, default_params, _capi
from salmon_b200._capi import EqClasses, Index
from salmon_b200.synth import synth_eq, shard_classes, synth_txome, synth_reads_fast
from salmon_b200.quant import quant_reads
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", rank=rank, world_size=world)
eq, proj, eff, uniq = synth_eq(seed=5, C=80000, M=30000, total_count=3_000_000)
sh = shard_classes(eq, rank, world)
ctx = EMContext(local)
uid = [_capi.nccl_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
ok_all = True
for vbem in (1, 0):
    for k in (1, 3, 40):
        p = default_params(use_vbem=vbem, min_iter=k, max_iter=k)
        alpha, st, ok = ctx.optimize(sh, p, proj, eff, uniq)
        if rank == 0:
            import oracle_lib as O
            ref, rst = O.em_optimize(eq, proj, eff, uniq, p)
            err = np.max(np.abs(alpha - ref) / np.maximum(np.abs(ref), 1e-9) * (np.abs(alpha - ref) > 1e-9))
            good = np.allclose(alpha, ref, rtol=1e-9, atol=1e-9) and st.iters == rst.iters
            ok_all &= bool(good)
            print(f"world {world} vbem {vbem} k {k}: iters {st.iters} max rel err {err:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
p = default_params()
alpha, st, ok = ctx.optimize(sh, p, proj, eff, uniq)
if rank == 0:
    import oracle_lib as O
    ref, rst = O.em_optimize(eq, proj, eff, uniq, p)
    good = st.iters == rst.iters and np.allclose(alpha, ref, rtol=1e-9, atol=1e-9)
    ok_all &= bool(good)
    print(f"converged run: iters {st.iters} vs {rst.iters} -> {'OK' if good else 'FAIL'}", flush=True)
ctx.close()
# ---- the same with alpha all-reduced INSIDE the persistent kernel over peer memory (CUDA IPC / NVLink)
ctx = EMContext(local)
ctx.peer_setup(dist, eq.n_txps)
for vbem in (1, 0):
    for k in (1, 3, 40):
        p = default_params(use_vbem=vbem, min_iter=k, max_iter=k)
        alpha, st, ok = ctx.optimize(sh, p, proj, eff, uniq)
        if rank == 0:
            ref, rst = O.em_optimize(eq, proj, eff, uniq, p)
            good = np.allclose(alpha, ref, rtol=1e-9, atol=1e-9) and st.iters == rst.iters and st.loop_kernel_launches == 1
            ok_all &= bool(good)
            print(f"fused world {world} vbem {vbem} k {k}: iters {st.iters} launches {st.loop_kernel_launches} -> {'OK' if good else 'FAIL'}", flush=True)
p = default_params()
alpha, st, ok = ctx.optimize(sh, p, proj, eff, uniq)
t = torch.from_numpy(alpha.copy()).cuda(); t0 = t.clone(); dist.broadcast(t0, 0)
same = bool(torch.equal(t, t0))       # every rank holds bit-identical alpha
if rank == 0:
    ref, rst = O.em_optimize(eq, proj, eff, uniq, p)
    good = st.iters == rst.iters and np.allclose(alpha, ref, rtol=1e-9, atol=1e-9) and same
    ok_all &= bool(good)
    print(f"fused converged run: iters {st.iters} vs {rst.iters}, {st.loop_kernel_ms / st.iters * 1e3:.1f} us/iter, "
          f"identical on all ranks {same} -> {'OK' if good else 'FAIL'}", flush=True)
ctx.close()
# ---- Stage A + quant, reads sharded
txps, _ = synth_txome(seed=51, n_genes=800)
left, right, truth = synth_reads_fast(txps, seed=52, n=120_000)
idx = Index(txps)
sl = slice(rank, None, world)
out = quant_reads(idx, left[sl], right[sl], device=local, batch=32768, dist=dist)
if rank == 0:
    one = quant_reads(idx, left, right, device=local, batch=32768)
    M = len(txps)
    g1 = out["n_mapped"] == one["n_mapped"]
    g2 = np.array_equal(out["unique_counts"], one["unique_counts"])
    g3 = abs(out["alpha"].sum() - one["alpha"].sum()) < 1e-6 * one["alpha"].sum()
    r = np.corrcoef(out["alpha"], one["alpha"])[0, 1]
    rel = np.abs(out["tpm"] - one["tpm"]) / np.maximum(one["tpm"], 1.0)
    true_counts = np.bincount(truth["tid"][truth["tid"] >= 0], minlength=M).astype(float)
    rt = np.corrcoef(out["alpha"], true_counts)[0, 1]
    rt1 = np.corrcoef(one["alpha"], true_counts)[0, 1]
    good = g1 and g2 and g3 and r > (0.9999 if world <= 2 else 0.9995) and rt > rt1 - 0.01   # shard-local burn-in: more shards, more deviation
    ok_all &= bool(good)
    print(f"stage A sharded over {world}: mapped equal {g1}, unique counts equal {g2}, sum alpha equal {g3}, "
          f"corr(alpha sharded, alpha single) {r:.6f}, median rel TPM diff {np.median(rel):.2e}, corr vs truth {rt:.4f} "
          f"-> {'OK' if good else 'FAIL'}; ALL {'OK' if ok_all else 'FAIL'}", flush=True)
dist.destroy_process_group()
