#!/bin/bash
mkdir -p gpurun_out
B="96:2048:0:3:12:100:30:1024:1024"
timeout 150 python scripts/sweep_em.py 300 "1:$B,11:$B,12:$B,11:96:2048:0:2:6:100:30:1024:1024" > gpurun_out/sweep_dyn.txt 2>&1
cat gpurun_out/sweep_dyn.txt
SB_EM_CONFIG=11 timeout 120 python -m pytest tests/test_em_gpu.py tests/test_sampling_gpu.py -m gpu -x -q > gpurun_out/tests_dyn.txt 2>&1
tail -3 gpurun_out/tests_dyn.txt
SB_EM_CONFIG=11 timeout 60 python scripts/timeline_em.py 500000 > gpurun_out/timeline_dyn.txt 2>&1
cat gpurun_out/timeline_dyn.txt
