#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r10_graphdev.log 2>&1
echo "graph device rc=$?"; grep -E "passed|failed|device graph build|Error|error" gpurun_out/r10_graphdev.log | tail -15
timeout 900 python -m pytest tests/test_dynamics_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r10_md.log 2>&1
echo "md rc=$?"; grep -E "passed|failed|steps|FIRE|NVE|Error|error|assert" gpurun_out/r10_md.log | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench_c3.json 2> gpurun_out/r10_bench_c3.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r10_bench_c3.json') if l.startswith('{')][0])
    print('c3 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'c4 ms', d['c4']['ms_per_step'], 'c4 e2e', d['c4']['e2e']['ms_per_step'])
    print('md', json.dumps(d['c4'].get('md'))[:1200])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r10_bench_c3.err').read()[-1500:])
PY
