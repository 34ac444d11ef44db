#!/bin/bash
# round 2, run 17: H2D probe, launch lists with DRAM bytes (c3, c4), ncu captures of the wgrad tcgen05 and bwd2 kernels
mkdir -p gpurun_out
timeout 300 python tools/h2d_probe.py > gpurun_out/r17_h2d_probe.log 2>&1; cat gpurun_out/r17_h2d_probe.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c3_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4 > gpurun_out/r17_launches_c3.log 2>&1
echo "launch list c3 rc=$?"
timeout 900 ncu --metrics $M --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c4_r2.csv python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-md > gpurun_out/r17_launches_c4.log 2>&1
echo "launch list c4 rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:wgrad_tc -s 12 -c 3 -o gpurun_out/prof_wgrad_tc_r2 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r17_ncu_wgrad.log 2>&1
echo "ncu wgrad rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bwd2 -s 2 -c 3 -o gpurun_out/prof_bwd2_r2 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r17_ncu_bwd2.log 2>&1
echo "ncu bwd2 rc=$?"
ls -la gpurun_out | tail -8
