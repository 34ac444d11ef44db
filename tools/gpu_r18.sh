#!/bin/bash
# round 2, run 18: compact wire format + two-phase pack/copy overlap: parity with the full format, e2e A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_wire.py tests/test_parity_configs_gpu.py -q -p no:cacheprovider -m gpu -k "wire or csr" > gpurun_out/r18_wire_tests.log 2>&1
echo "wire tests rc=$?"; tail -4 gpurun_out/r18_wire_tests.log
for w in 1 0; do
CHGNET_B200_WIRE=$w timeout 300 python tools/h2d_probe.py > gpurun_out/r18_probe_w$w.log 2>&1; grep -E "build_batch|cudaHostAlloc" gpurun_out/r18_probe_w$w.log
done
for wl in c3 c2; do
for w in 1 0; do
CHGNET_B200_WIRE=$w timeout 900 python bench.py --workload $wl --no-cpu-baseline --no-md > gpurun_out/r18_bench_${wl}_w$w.json 2> gpurun_out/r18_bench_${wl}_w$w.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r18_bench_${wl}_w$w.json') if l.startswith('{')][0])
    print('$wl wire=$w', 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e'].get('breakdown'), 'c4', (d.get('c4') or {}).get('ms_per_step'), ((d.get('c4') or {}).get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r18_bench_${wl}_w$w.err').read()[-1500:])
PY
done
done
