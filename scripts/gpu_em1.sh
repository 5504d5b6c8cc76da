#!/bin/bash
# round-2 EM kernel check: parity tests, configuration sweep, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_em_gpu.py tests/test_sampling_gpu.py -x -q -m gpu > gpurun_out/r2_em_tests.txt 2>&1
tail -5 gpurun_out/r2_em_tests.txt
timeout 600 python scripts/sweep_em.py 300 "${1:-config=0:rebalance=0,config=0:rebalance=3,config=5:rebalance=0,config=5:rebalance=3,config=6:rebalance=0,config=6:rebalance=3,config=7:rebalance=0,config=7:rebalance=3,config=8:rebalance=0,config=8:rebalance=3,config=7:rebalance=3:lwarp=256,config=7:rebalance=3:lmax=32}" > gpurun_out/r2_em_sweep.txt 2>&1
cat gpurun_out/r2_em_sweep.txt
