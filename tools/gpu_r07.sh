#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_configs_gpu.py -q -p no:cacheprovider -k "csr or c5 or c2_full" > gpurun_out/r07_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r07_tests.log
tail -15 gpurun_out/r07_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r07_bench_c3.json 2> gpurun_out/r07_bench_c3.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r07_bench_c3.json') if l.startswith('{')][0])
print('c3 ms', d['ms_per_step'], 'e2e', d['e2e'])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gated_ws_fwd -s 2 -c 2 -o gpurun_out/prof_fused_r07 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r07_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r07_ncu.log
ls -la gpurun_out/*.ncu-rep
