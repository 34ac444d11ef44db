#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dynamics_device_gpu.py tests/test_graph_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r13_f1f2.log 2>&1
echo "f1/f2 rc=$?"; grep -E "passed|failed|steps|FIRE|NVE|device graph build|Error|error|assert" gpurun_out/r13_f1f2.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r13_bench_c3.json 2> gpurun_out/r13_bench_c3.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r13_bench_c3.json') if l.startswith('{')][0])
    print('c3 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['breakdown'], 'c4 ms', d['c4']['ms_per_step'], 'c4 e2e', d['c4']['e2e']['ms_per_step'])
    print('roofline', json.dumps(d['roofline'])[:1500])
    print('md', json.dumps(d['c4'].get('md'))[:1200])
    for k,v in list(d['kernel_shares'].items())[:8]: print('  ', k, v)
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r13_bench_c3.err').read()[-1500:])
PY
# ---- profiles ----
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4 > gpurun_out/r13_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gated_ws_(fwd|bwd)" -s 7 -c 4 -o gpurun_out/prof_gated_ws_r2 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r13_ncu_gated.log 2>&1
echo "ncu gated rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:segment_sum -s 3 -c 2 -o gpurun_out/prof_scatter_c4_r2 python bench.py --workload c4 --scatter-only > gpurun_out/r13_ncu_scatter_c4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:segment_sum -s 3 -c 2 -o gpurun_out/prof_scatter_c3_r2 python bench.py --workload c3 --scatter-only > gpurun_out/r13_ncu_scatter_c3.log 2>&1
echo "ncu scatter rc=$?"
ls -la gpurun_out/*.ncu-rep
