#!/bin/bash
# round-2 EM kernel check: parity tests, configuration sweep, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_em_gpu.py tests/test_sampling_gpu.py -x -q -m gpu > gpurun_out/r2_em_tests.txt 2>&1
tail -5 gpurun_out/r2_em_tests.txt
timeout 600 python scripts/sweep_em.py 300 "${1:-config=0:rebalance=0,config=0:rebalance=1,config=1:rebalance=1,config=2:rebalance=1,config=3:rebalance=1}" > gpurun_out/r2_em_sweep.txt 2>&1
cat gpurun_out/r2_em_sweep.txt
