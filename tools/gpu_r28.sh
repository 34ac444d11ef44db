#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -m gpu -k "defaults or conv" > gpurun_out/r28_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r28_tests.log
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-c4 > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r28_bench.json') if l.startswith('{')][0])
ks=d['kernel_shares']
print('c3 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'replay', d.get('graph_replay'), {k: ks[k]['ms'] for k in ('atom_conv_bwd','bond_conv_bwd','atom_conv_fused','bond_conv_fused','segment_sum','linear')})
PY
done
