#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/h2d_probe2.py > gpurun_out/r19_probe2.log 2>&1; cat gpurun_out/r19_probe2.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -q -p no:cacheprovider -k "wgrad or train" > gpurun_out/r19_kern.log 2>&1
echo "wgrad/train tests rc=$?"; tail -3 gpurun_out/r19_kern.log
for impl in 1 0; do
CHG_WGRAD_IMPL=$impl timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r19_bench_c5_w$impl.json 2> gpurun_out/r19_bench_c5_w$impl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r19_bench_c5_w$impl.json') if l.startswith('{')][0])
    print('c5 wgrad_impl=$impl', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
    for k,v in list(d['kernel_shares'].items())[:3]: print('  ', k, v)
except Exception as e:
    print('c5 parse failed', e); print(open('gpurun_out/r19_bench_c5_w$impl.err').read()[-1500:])
PY
done
