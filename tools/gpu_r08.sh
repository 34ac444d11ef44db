#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "defaults" -p no:cacheprovider > gpurun_out/r08_fused_spec.log 2>&1
echo "fused spec rc=$?"; tail -5 gpurun_out/r08_fused_spec.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_parity_configs_gpu.py -q -p no:cacheprovider > gpurun_out/r08_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r08_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r08_bench_c3.json 2> gpurun_out/r08_bench_c3.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r08_bench_c3.json') if l.startswith('{')][0])
print('c3 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['breakdown'], 'c4 ms', d['c4']['ms_per_step'], 'c4 e2e', d['c4']['e2e']['ms_per_step'])
for k,v in list(d['kernel_shares'].items())[:8]: print('  ', k, v)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gated_ws_fwd -s 2 -c 2 -o gpurun_out/prof_fused_r08 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r08_ncu.log 2>&1
echo "ncu rc=$?"
