#!/bin/bash
# multi-GPU check (run with gpurun --gpus N): parity of the sharded EM / Stage A, then the EM bench fused vs NCCL
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/check_multigpu.py > gpurun_out/mg_check_n$N.txt 2>&1; tail -16 gpurun_out/mg_check_n$N.txt | grep -v "^W\|^\[W"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-stage-a > gpurun_out/mg_bench_fused_n$N.json 2> gpurun_out/mg_bench_fused_n$N.err; tail -c 1500 gpurun_out/mg_bench_fused_n$N.json | head -c 600; echo
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-stage-a --nccl > gpurun_out/mg_bench_nccl_n$N.json 2> gpurun_out/mg_bench_nccl_n$N.err; head -c 400 gpurun_out/mg_bench_nccl_n$N.json; echo
