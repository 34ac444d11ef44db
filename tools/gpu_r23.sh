#!/bin/bash
mkdir -p gpurun_out
for wl in c3 c4; do
for v in "4 0" "8 0" "8 2" "4 2" "8 4"; do
set -- $v
CHG_SEGSUM_UNROLL=$1 CHG_SEGSUM_S=$2 timeout 300 python bench.py --workload $wl --scatter-only 2> /dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', d)
        print('$wl unroll=$1 S=$2', 'frac', r.get('frac'), 'us', r.get('us_per_launch'), 'w64', (r.get('forward_message_sum_w64') or {}).get('frac'), 'at10k', (r.get('at_10k_atoms') or {}).get('frac'))
"
done
done
