#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -m gpu > gpurun_out/r29_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r29_tests.log
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-md > gpurun_out/r29_bench.json 2> gpurun_out/r29_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r29_bench.json') if l.startswith('{')][0])
ks=d['kernel_shares']
print('c3 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'c4', d['c4']['ms_per_step'], {k: ks[k]['ms'] for k in ('atom_conv_bwd','bond_conv_bwd','atom_conv_fused','bond_conv_fused','segment_sum','linear')})
PY
done
