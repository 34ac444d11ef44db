#!/bin/bash
# round 2, run 24: validation of the final tree: sanitizer, full gpu suite, smoke, benches (c3 default incl. c4 + MD, c2, c1, c5), launch list
mkdir -p gpurun_out
bash tools/gpu_sanitize_r2.sh
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r24_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -3 gpurun_out/r24_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r24_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r24_smoke.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/r24_bench_default.json 2> gpurun_out/r24_bench_default.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r24_bench_default.json') if l.startswith('{')][0])
print('default', d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'parity', d['parity'].get('ok'))
print('roofline', {k: v for k, v in d['roofline'].items() if k in ('frac', 'us_per_launch', 'achieved')}, 'launches', d['gpu_launches'], 'clocks', d['clocks'])
c4 = d.get('c4', {})
print('c4', c4.get('ms_per_step'), (c4.get('e2e') or {}).get('ms_per_step'), (c4.get('e2e_from_structure') or {}).get('ms_per_step'), 'roofline', (c4.get('roofline') or {}).get('frac'))
print('md', {k: (v.get('ms_per_step') if isinstance(v, dict) else v) for k, v in (c4.get('md') or {}).items()})
print('cpu', d.get('cpu_baseline'))
PY
for wl in c2 c1 c5; do
timeout 900 python bench.py --workload $wl --no-c4 > gpurun_out/r24_bench_$wl.json 2> gpurun_out/r24_bench_$wl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r24_bench_$wl.json') if l.startswith('{')][0])
    print('$wl', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))
except Exception as e:
    print('$wl parse failed', e)
PY
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c3_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4 > gpurun_out/r24_launches_c3.log 2>&1
echo "launch list c3 rc=$?"
