#!/bin/bash
# round 2, run 31: validation of the final tree
mkdir -p gpurun_out
bash tools/gpu_sanitize_r2.sh | tail -4
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r31_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -2 gpurun_out/r31_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r31_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python bench.py > gpurun_out/r31_bench_default.json 2> gpurun_out/r31_bench_default.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r31_bench_default.json') if l.startswith('{')][0])
print('default', d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'parity', d['parity'].get('ok'), 'binding', d.get('cpu_binding'))
print('roofline', {k: v for k, v in d['roofline'].items() if k in ('frac', 'us_per_launch', 'achieved')}, 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'replay', (d.get('graph_replay') or {}).get('ms_per_step'))
c4 = d.get('c4', {})
print('c4', c4.get('ms_per_step'), (c4.get('e2e') or {}).get('ms_per_step'), (c4.get('e2e_from_structure') or {}).get('ms_per_step'), 'roofline', (c4.get('roofline') or {}).get('frac'))
print('md', {k: (v.get('ms_per_step') if isinstance(v, dict) else v) for k, v in (c4.get('md') or {}).items()})
print('cpu', d.get('cpu_baseline'))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r31_ref.json 2> gpurun_out/r31_ref.err; echo "reference arm rc=$?"; grep '^{' gpurun_out/r31_ref.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('reference', d['value'], d['unit'], d['ms_per_step'], d['cpu_baseline']['sample'][:160], d.get('product_so_mapped'))"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c3_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4 > gpurun_out/r31_launches_c3.log 2>&1
echo "launch list c3 rc=$?"
