#!/bin/bash
# GPU run: fused gated kernel bring-up (tests first, in a guarded process), then A/B benches
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "defaults" -p no:cacheprovider > gpurun_out/r06_fused_spec.log 2>&1
echo "fused spec rc=$?" | tee -a gpurun_out/r06_fused_spec.log
tail -25 gpurun_out/r06_fused_spec.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r06_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r06_tests.log
tail -40 gpurun_out/r06_tests.log
for impl in 3 0; do
  CHG_GATED_IMPL=$impl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r06_bench_c3_g$impl.json 2> gpurun_out/r06_bench_c3_g$impl.err
  echo "bench gated_impl=$impl rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r06_bench_c3_g$impl.json') if l.startswith('{')][0])
    print('c3 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'c4 ms', d['c4']['ms_per_step'], 'c4 e2e', d['c4']['e2e']['ms_per_step'])
    for k,v in list(d['kernel_shares'].items())[:8]: print('  ', k, v)
except Exception as e: print('parse failed', e)
PY
done
