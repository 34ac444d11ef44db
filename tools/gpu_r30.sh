#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8
lscpu | grep -i "numa\|socket\|model name" | head -8
for b in 1 0; do
PROBE_BIND=$b timeout 300 python tools/h2d_probe.py 2>&1 | grep -E "binding|build_batch|H2D pinned 35"
CHGNET_BENCH_BIND=$b timeout 600 python bench.py --no-cpu-baseline --no-md > gpurun_out/r30_bench_b$b.json 2> gpurun_out/r30_bench_b$b.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r30_bench_b$b.json') if l.startswith('{')][0])
    print('bind=$b', d.get('cpu_binding'), '| c3 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e']['breakdown'], 'c4', d['c4']['ms_per_step'], d['c4']['e2e']['ms_per_step'], (d['c4'].get('e2e_from_structure') or {}).get('ms_per_step'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r30_bench_b$b.err').read()[-1200:])
PY
done
