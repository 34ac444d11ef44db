#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "wgrad or defaults or training" > gpurun_out/r15_kern.log 2>&1
echo "kernel tests rc=$?"; tail -12 gpurun_out/r15_kern.log
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_parity_configs_gpu.py -q -p no:cacheprovider -k "train or c5 or trainer or forward_in_training" -s > gpurun_out/r15_train.log 2>&1
echo "train tests rc=$?"; grep -E "passed|failed|worst|Error|assert" gpurun_out/r15_train.log | tail -10
for impl in 1 0; do
CHG_WGRAD_IMPL=$impl timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r15_bench_c5_w$impl.json 2> gpurun_out/r15_bench_c5_w$impl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r15_bench_c5_w$impl.json') if l.startswith('{')][0])
    print('c5 wgrad_impl=$impl', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['breakdown'])
    for k,v in list(d['kernel_shares'].items())[:6]: print('  ', k, v)
except Exception as e:
    print('c5 parse failed', e); print(open('gpurun_out/r15_bench_c5_w$impl.err').read()[-1500:])
PY
done
timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r15_bench_c1.json 2> gpurun_out/r15_bench_c1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r15_bench_c1.json') if l.startswith('{')][0]); print('c1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
cat > /tmp/ps.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, cProfile, pstats
from chgnet_b200 import graphgen
from chgnet_b200.model import CHGNet
m = CHGNet.from_file("tests/golden/chgnet_0.3.0_weights.npz", version="0.3.0").to("cuda")
z, frac, lat = graphgen.limno2_structure((10, 5, 25), 0.02, 4000)
for _ in range(3): m.predict_structure((z, frac, lat), task="efs")
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0=time.perf_counter()
for _ in range(5): m.predict_structure((z, frac, lat), task="efs")
torch.cuda.synchronize()
print('predict_structure ms', (time.perf_counter()-t0)/5*1e3)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(12)
PY
timeout 300 python /tmp/ps.py > gpurun_out/r15_ps_profile.log 2>&1; head -30 gpurun_out/r15_ps_profile.log
