#!/bin/bash
# round 2, run 16: wgrad A/B after hoisting the stage loads, Trainer in-place repack, full gpu suite, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "wgrad" > gpurun_out/r16_kern.log 2>&1
echo "wgrad tests rc=$?"; tail -3 gpurun_out/r16_kern.log
for impl in 1 0; do
CHG_WGRAD_IMPL=$impl timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r16_bench_c5_w$impl.json 2> gpurun_out/r16_bench_c5_w$impl.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r16_bench_c5_w$impl.json') if l.startswith('{')][0])
    print('c5 wgrad_impl=$impl', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['breakdown'])
    for k,v in list(d['kernel_shares'].items())[:6]: print('  ', k, v)
except Exception as e:
    print('c5 parse failed', e); print(open('gpurun_out/r16_bench_c5_w$impl.err').read()[-1500:])
PY
done
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r16_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -5 gpurun_out/r16_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r16_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r16_smoke.log
timeout 1200 python bench.py > gpurun_out/r16_bench_default.json 2> gpurun_out/r16_bench_default.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r16_bench_default.json') if l.startswith('{')][0])
print('default', d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'))
print('roofline', d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('c4', {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ('ms_per_step','atoms_per_s','steps_per_s')}) for k, v in d.get('c4', {}).items()})
print('collective', d.get('collective'))
PY
