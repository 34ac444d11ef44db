#!/bin/bash
# compute-sanitizer memcheck over one full forward + reverse pass on a small batch (every kernel),
# for both implementations of the GEMM-bearing kernels.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from chgnet_b200 import graphgen
from chgnet_b200.model import CHGNet
from chgnet_b200._lib import CudaKernels
m = CHGNet.from_file("tests/golden/chgnet_0.3.0_weights.npz", version="0.3.0").to("cuda")
gs = graphgen.random_graphs(2, 8, 12, 9300)
K = CudaKernels()
for lin, gat in ((1, 0), (2, 1), (0, 2)):
    K.set_option("linear_impl", lin); K.set_option("gated_impl", gat)
    out = m.predict_graph(gs, task="efsm", return_site_energies=True, return_crystal_feas=True)
    torch.cuda.synchronize()
    print("impl", lin, gat, [float(o["e"]) for o in out])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
tail -15 gpurun_out/sanitizer_memcheck.log
timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.json 2>gpurun_out/bench_c1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_c1.json').read().strip().splitlines()[-1])
print('c1', d['value'], d['ms_per_step'], d['e2e'])"
