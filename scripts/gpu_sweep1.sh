#!/bin/bash
mkdir -p gpurun_out
B="1:96:2048:0:3:12:100:30"
S="$B,6:96:2048:0:3:12:100:30,6:96:512:0:3:12:100:30,6:96:512:150:3:12:100:30,6:96:512:150:2:4:100:30,6:96:512:150:1:2:100:30,6:32:512:150:2:4:100:30,6:64:512:150:2:4:100:30,6:192:512:150:2:4:100:30,6:512:2048:150:2:4:100:30,7:96:512:150:2:4:100:30,8:96:512:150:2:4:100:30,9:96:512:150:2:4:100:30,10:96:512:150:2:4:100:30,9:96:512:150:2:4:100:100,9:192:512:150:2:4:100:30,9:96:512:300:2:4:100:30"
timeout 500 python scripts/sweep_em.py 300 "$S" > gpurun_out/sweep1.txt 2>&1
SB_EM_CONFIG=6 SB_EM_LWARP=512 SB_EM_BALANCE=150 timeout 300 python -m pytest tests/test_em_gpu.py tests/test_sampling_gpu.py -m gpu -x -q > gpurun_out/tests_cfg6.txt 2>&1
SB_EM_CONFIG=9 SB_EM_LWARP=512 SB_EM_BALANCE=150 timeout 300 python -m pytest tests/test_em_gpu.py tests/test_sampling_gpu.py -m gpu -x -q > gpurun_out/tests_cfg9.txt 2>&1
SB_EM_CONFIG=6 SB_EM_LWARP=512 SB_EM_BALANCE=150 timeout 200 python scripts/timeline_em.py 500000 > gpurun_out/timeline_cfg6.txt 2>&1
SB_EM_CONFIG=9 SB_EM_LWARP=512 SB_EM_BALANCE=150 timeout 200 python scripts/timeline_em.py 500000 > gpurun_out/timeline_cfg9.txt 2>&1
tail -3 gpurun_out/tests_cfg6.txt gpurun_out/tests_cfg9.txt; cat gpurun_out/sweep1.txt
