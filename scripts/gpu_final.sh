#!/bin/bash
# round-end style validation: GPU tests, smoke, bench line, then the ncu launch list of the same bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/tests_gpu_final.txt 2>&1
tail -6 gpurun_out/tests_gpu_final.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; tail -2 gpurun_out/smoke_final.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_final_n1.json').read().strip().splitlines()[-1])
    sa = d.get('stage_a', {})
    print('EM', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
    print('StageA', sa.get('value'), 'e2e', sa.get('e2e', {}).get('value'), 'clocks', sa.get('clocks'))
    print('from_files', sa.get('from_files'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_final_n1.err').read()[-1500:])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-files --cpu-budget 1 > gpurun_out/ncu_bench_final.log 2>&1
wc -l gpurun_out/launches_final.csv
