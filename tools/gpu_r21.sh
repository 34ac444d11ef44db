#!/bin/bash
mkdir -p gpurun_out
for st in wc pinned; do
echo "== staging $st"
CHGNET_B200_STAGING=$st timeout 300 python -m pytest tests/test_batch_wire.py -q -p no:cacheprovider -m gpu 2>&1 | tail -1
CHGNET_B200_STAGING=$st timeout 300 python tools/h2d_probe.py 2>&1 | grep build_batch
CHGNET_B200_STAGING=$st timeout 300 python tools/time_build_batch.py c3 2>&1 | grep -E "build_batch|pack_wire"
CHGNET_B200_STAGING=$st timeout 900 python bench.py --workload c3 --no-cpu-baseline --no-md > gpurun_out/r21_bench_c3_$st.json 2> gpurun_out/r21_bench_c3_$st.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r21_bench_c3_$st.json') if l.startswith('{')][0])
    print('c3 $st', 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('breakdown'), 'c4', (d.get('c4') or {}).get('ms_per_step'), ((d.get('c4') or {}).get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r21_bench_c3_$st.err').read()[-1500:])
PY
done
