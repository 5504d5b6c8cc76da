"""torchrun -> per-rank class shards, alpha all-reduced per iteration; rank 0 compares the
result with the oracle run on the union of the shards (same iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, torch.distributed as dist
from salmon_b200 import EMContext, default_params, _capi
from salmon_b200._capi import EqClasses
from salmon_b200.synth import synth_eq, shard_classes
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
# run to convergence
p = default_params()
alpha, st, ok = ctx.optimize(sh, p, proj, eff, uniq)
if rank == 0:
    import oracle_lib as O
    ref, rst = O.em_optimize(eq, proj, eff, uniq, p)
    good = st.iters == rst.iters and np.allclose(alpha, ref, rtol=1e-9, atol=1e-9)
    ok_all &= bool(good)
    print(f"converged run: iters {st.iters} vs {rst.iters} -> {'OK' if good else 'FAIL'}; ALL {'OK' if ok_all else 'FAIL'}", flush=True)
dist.destroy_process_group()
