#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r11_graphdev.log 2>&1
echo "graph device rc=$?"; grep -E "passed|failed|device graph build|Error|error|assert" gpurun_out/r11_graphdev.log | tail -15
timeout 900 python -m pytest tests/test_dynamics_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r11_md.log 2>&1
echo "md rc=$?"; grep -E "passed|failed|steps|FIRE|NVE|Error|error|assert" gpurun_out/r11_md.log | tail -12
timeout 300 python tools/time_build_batch.py c3 > gpurun_out/r11_time_bb_c3.log 2>&1; head -25 gpurun_out/r11_time_bb_c3.log
timeout 300 python tools/time_build_batch.py c4 > gpurun_out/r11_time_bb_c4.log 2>&1; head -12 gpurun_out/r11_time_bb_c4.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench_c3.json 2> gpurun_out/r11_bench_c3.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r11_bench_c3.json') if l.startswith('{')][0])
    print('md', json.dumps(d['c4'].get('md'))[:900])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r11_bench_c3.err').read()[-1500:])
PY
