#!/bin/bash
mkdir -p gpurun_out
nproc; uptime
for i in 1 2; do
timeout 1500 python bench.py > gpurun_out/r25_bench_default_$i.json 2> gpurun_out/r25_bench_default_$i.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r25_bench_default_$i.json') if l.startswith('{')][0])
c4=d.get('c4') or {}
print('default run $i: ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e']['breakdown'], 'c4', c4.get('ms_per_step'), (c4.get('e2e') or {}).get('ms_per_step'), (c4.get('e2e_from_structure') or {}).get('ms_per_step'), 'md', ((c4.get('md') or {}).get('device_driver') or {}).get('ms_per_step'))
PY
uptime
done
timeout 600 python tools/time_md_small.py 2>&1 | grep -v Warn | tail -5
