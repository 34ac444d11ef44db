#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/linear_ab.py 2>&1 | tee gpurun_out/linear_ab.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -5
