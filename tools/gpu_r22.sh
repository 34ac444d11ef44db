#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_device_gpu.py tests/test_batch_wire.py -q -p no:cacheprovider -m gpu > gpurun_out/r22_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r22_tests.log
timeout 300 python tools/time_convert_many.py 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/r22_convert_many.log
