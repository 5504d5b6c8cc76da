"""torchrun --nproc-per-node N: per-stage timeline of one iteration of the fused multi-GPU EM kernel (dev helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from salmon_b200 import EMContext, default_params
from salmon_b200.synth import synth_eq
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", rank=rank, world_size=world)
eq, proj, eff, uniq = synth_eq(seed=1 + rank)
tp = torch.from_numpy(proj).cuda(); dist.all_reduce(tp); proj = tp.cpu().numpy()
tu = torch.from_numpy(uniq.astype(np.int64)).cuda(); dist.all_reduce(tu); uniq = tu.cpu().numpy().astype(np.uint64)
ctx = EMContext(local)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
ctx.peer_setup(dist, eq.n_txps)
p = default_params(min_iter=60, max_iter=60)
ctx.upload(eq, proj, eff, uniq); ctx.prepare(p)
dist.barrier(); ctx.run()
nw = ctx.arm_timeline(40)
dist.barrier(); r = ctx.run()
t = ctx.read_timeline(nw).astype(np.int64)
if rank == 0:
    names = ["P1 start", "P1 end", "gs1 end", "P2p end", "owner end", "unpack end", "gs2 end", "iter end"]
    print("world", world, "loop us/iter", r.loop_kernel_ms / 60 * 1e3)
    # per-warp differences (same SM clock); medians over warps
    for i in range(1, 8):
        d = (t[:, i] - t[:, i - 1]) / 1e3
        print(f"{names[i-1]:10s} -> {names[i]:10s}  p10 {np.percentile(d,10):7.2f}  p50 {np.median(d):7.2f}  p90 {np.percentile(d,90):7.2f}  max {d.max():7.2f} us")
    d = (t[:, 7] - t[:, 0]) / 1e3
    print("P1 start -> iteration end  p50 %.2f" % np.median(d))
dist.barrier()
ctx.close()
dist.destroy_process_group()
