#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_wire.py tests/test_model_gpu.py -q -p no:cacheprovider -m gpu > gpurun_out/r20_tests.log 2>&1
echo "wire + model tests rc=$?"; tail -3 gpurun_out/r20_tests.log
timeout 300 python tools/h2d_probe.py 2>&1 | grep build_batch
timeout 300 python tools/time_build_batch.py c3 > gpurun_out/r20_time_bb_c3.log 2>&1; head -22 gpurun_out/r20_time_bb_c3.log
timeout 300 python tools/time_convert_many.py 2>&1 | tail -3
for wl in c3 c2; do
timeout 900 python bench.py --workload $wl --no-cpu-baseline --no-md > gpurun_out/r20_bench_${wl}.json 2> gpurun_out/r20_bench_${wl}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r20_bench_${wl}.json') if l.startswith('{')][0])
    print('$wl', 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e'].get('breakdown'), 'c4', (d.get('c4') or {}).get('ms_per_step'), ((d.get('c4') or {}).get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r20_bench_${wl}.err').read()[-1500:])
PY
done
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r20_bench_c5.json 2> gpurun_out/r20_bench_c5.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r20_bench_c5.json') if l.startswith('{')][0]); print('c5', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
