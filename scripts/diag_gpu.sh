#!/bin/bash
# One diagnostic GPU call: EM per-warp timeline, ncu full+source capture of the persistent EM kernel,
# ncu full+source capture of the Stage A top kernels at human-transcriptome scale.
mkdir -p gpurun_out
timeout 300 python scripts/timeline_em.py 500000 > gpurun_out/timeline.txt 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_em_persistent -s 1 -c 1 -f -o gpurun_out/em_full \
  python scripts/prof_em.py 100 1 1 100 30 > gpurun_out/ncu_em.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_seed_chain_w|k_assign|k_dp_pair|k_dp_classify" -s 4 -c 4 -f -o gpurun_out/stagea_full \
  python scripts/bench_map.py 60000 262144 262144 1 > gpurun_out/ncu_stagea.log 2>&1
ls -la gpurun_out
