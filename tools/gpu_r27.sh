#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -m gpu -k "static_evaluator or predict" > gpurun_out/r27_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r27_tests.log
for wl in c1 c2 c3; do
for rp in "" "--graph-replay"; do
timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-c4 $rp > gpurun_out/r27_bench.json 2> gpurun_out/r27_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r27_bench.json') if l.startswith('{')][0])
    print('$wl [$rp]', 'ms', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r27_bench.err').read()[-1500:])
PY
done
done
