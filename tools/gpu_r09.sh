#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "defaults" -p no:cacheprovider > gpurun_out/r09_spec.log 2>&1
echo "spec (fwd+bwd ws) rc=$?"; tail -5 gpurun_out/r09_spec.log
timeout 600 python -m pytest tests/test_graph_device_gpu.py -q -p no:cacheprovider -s > gpurun_out/r09_graphdev.log 2>&1
echo "graph device rc=$?"; tail -25 gpurun_out/r09_graphdev.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_configs_gpu.py -q -p no:cacheprovider > gpurun_out/r09_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r09_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r09_bench_c3.json 2> gpurun_out/r09_bench_c3.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r09_bench_c3.json') if l.startswith('{')][0])
print('c3 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['breakdown'], 'c4 ms', d['c4']['ms_per_step'], 'c4 e2e', d['c4']['e2e']['ms_per_step'])
for k,v in list(d['kernel_shares'].items())[:8]: print('  ', k, v)
PY
