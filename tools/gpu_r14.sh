#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_sanitize_r2.sh
timeout 900 python -m pytest tests/test_dynamics_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r14_md.log 2>&1
echo "md rc=$?"; grep -E "passed|failed|steps|Error|error|assert" gpurun_out/r14_md.log | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r14_bench_c3.json 2> gpurun_out/r14_bench_c3.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r14_bench_c3.json') if l.startswith('{')][0])
    print('c3 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'c4 ms', d['c4']['ms_per_step'], 'c4 e2e', d['c4']['e2e']['ms_per_step'])
    print('c4 from structure', d['c4'].get('e2e_from_structure'))
    print('md', json.dumps(d['c4'].get('md'))[:1800])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r14_bench_c3.err').read()[-1500:])
PY
for wl in c2 c1; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-c4 > gpurun_out/r14_bench_$wl.json 2> gpurun_out/r14_bench_$wl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r14_bench_$wl.json') if l.startswith('{')][0])
    print('$wl', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'parity', d.get('parity'))
except Exception as e:
    print('$wl parse failed', e)
PY
done
