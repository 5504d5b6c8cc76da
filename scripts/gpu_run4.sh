#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_long_reads.py 3000 3 > gpurun_out/diag_long.txt 2>&1
cat gpurun_out/diag_long.txt
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/diag_long_reads.py 600 1 default > gpurun_out/racecheck.txt 2>&1
grep -c . gpurun_out/racecheck.txt; grep -i "race\|hazard\|error" gpurun_out/racecheck.txt | head -20; tail -5 gpurun_out/racecheck.txt
