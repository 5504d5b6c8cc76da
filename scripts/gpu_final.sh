#!/bin/bash
# round-end style validation: smoke, bench line (both arms), the ncu launch list of the same bench command, one ncu --set full
# capture of the dominant kernels (EM persistent kernel at the bench's launch shape; Stage A seed kernel)
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; tail -2 gpurun_out/smoke_final.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_final_n1.json').read().strip().splitlines()[-1])
    sa = d.get('stage_a', {})
    print('EM', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'], 'cpu', d.get('cpu_baseline'))
    print('StageA', sa.get('value'), 'e2e', sa.get('e2e', {}).get('value'), 'roof', sa.get('roofline', {}).get('frac'), 'clocks', sa.get('clocks'))
    print('from_files', sa.get('from_files'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_final_n1.err').read()[-1500:])
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --no-stage-a > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; head -c 300 gpurun_out/bench_final_ref.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-files --cpu-budget 1 > gpurun_out/ncu_bench_final.log 2>&1
wc -l gpurun_out/launches_final.csv
# EM persistent kernel, 1000 iterations per launch as in the bench (prepare: 1 re-balancing launch, then 2 runs)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_em_persistent -s 2 -c 1 -o gpurun_out/em_r2_final python scripts/prof_em.py 1000 1 0 100 30 > gpurun_out/ncu_em_final.log 2>&1; tail -1 gpurun_out/ncu_em_final.log
# Stage A: seed + assign + dp kernels of one chunk at human scale
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seed_chain_w|k_assign|k_dp_pair|k_dp_classify" -s 8 -c 4 -o gpurun_out/stagea_r2_final python scripts/bench_map.py 60000 524288 262144 1 > gpurun_out/ncu_stagea_final.log 2>&1; tail -1 gpurun_out/ncu_stagea_final.log
