#!/bin/bash
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none -k regex:segment_sum -s 3 -c 2 -o gpurun_out/prof_scatter_c3_r2b python bench.py --workload c3 --scatter-only > gpurun_out/r34_ncu_scatter_c3.log 2>&1
echo "ncu scatter rc=$?"; tail -2 gpurun_out/r34_ncu_scatter_c3.log | cut -c1-200
