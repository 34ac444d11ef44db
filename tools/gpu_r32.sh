#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dynamics_device_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider -m gpu -k "dynamics or md or fire or static_evaluator or trajectory or device" > gpurun_out/r32_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r32_tests.log
timeout 300 python bench.py --workload c1 --no-cpu-baseline --no-c4 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('c1', d['ms_per_step'], d['graph_replay'])"
