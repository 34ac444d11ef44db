#!/bin/bash
mkdir -p gpurun_out
for t in 8 16; do
CHG_PACK_THREADS=$t timeout 300 python tools/h2d_probe.py 2>&1 | grep "build_batch" | sed "s/^/threads=$t /"
done
CHG_PACK_THREADS=8 timeout 300 python bench.py --no-cpu-baseline --no-c4 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('threads=8 c3 e2e', d['e2e']['ms_per_step'], d['e2e']['breakdown'])"
