#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/tests_gpu5.txt 2>&1
tail -15 gpurun_out/tests_gpu5.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke5.txt 2>&1; tail -3 gpurun_out/smoke5.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5_n1.json 2> gpurun_out/bench5_n1.err; tail -c 1500 gpurun_out/bench5_n1.json; tail -3 gpurun_out/bench5_n1.err
