#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/tests_gpu3.txt 2>&1
tail -25 gpurun_out/tests_gpu3.txt
SB_MAP_CHUNK=262144 timeout 500 python scripts/sweep_map.py 60000 1048576 262144 65536,131072,262144 > gpurun_out/sweep_map.txt 2>&1
tail -8 gpurun_out/sweep_map.txt
