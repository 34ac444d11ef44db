#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_batch_wire.py -q -p no:cacheprovider -m gpu 2>&1 | tail -1
timeout 120 python bench.py --no-cpu-baseline --no-c4 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('c3 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['breakdown'], d['parity'])"
