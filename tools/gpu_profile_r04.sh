#!/bin/bash
# Round-1 final profile pass (runs on the GPU box via gpurun): benches, reference arms, ncu launch
# lists, full captures of the scatter kernel, the warp-specialised linear, wgrad and a second-order kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r04}
for wl in c2 c4 c3 c5; do
  extra="--no-cpu-baseline"; [ "$wl" = "c2" ] && extra=""
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 $extra > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  tail -c 300 gpurun_out/bench_${wl}_${TAG}.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c2_${TAG}.json 2>> gpurun_out/bench_ref.err
timeout 600 python bench.py --impl reference --workload c5 --steps 2 --warmup 1 --cpu-sample 4 > gpurun_out/bench_ref_c5_${TAG}.json 2>> gpurun_out/bench_ref.err
for wl in c2 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_${wl}_${TAG}.csv python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/ncu_launch_${wl}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:segment_sum -s 3 -c 2 \
  -o gpurun_out/prof_scatter_${TAG} -f python bench.py --workload c4 --scatter-only > gpurun_out/ncu_scatter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:segment_sum -s 3 -c 2 \
  -o gpurun_out/prof_scatter_c2_${TAG} -f python bench.py --workload c2 --scatter-only > gpurun_out/ncu_scatter_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_ws -s 20 -c 4 \
  -o gpurun_out/prof_linearws_${TAG} -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_linearws.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 40 -c 4 \
  -o gpurun_out/prof_wgrad_${TAG} -f python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gated_bwd2 -s 2 -c 3 \
  -o gpurun_out/prof_bwd2_${TAG} -f python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bwd2.log 2>&1
ls -la gpurun_out | tail -30
