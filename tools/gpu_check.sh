#!/bin/bash
# quick GPU validation + bench (no ncu)
set -u
mkdir -p gpurun_out
TAG=${1:-r01b}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_${TAG}.log
for wl in c2 c4 c3; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  tail -c 400 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
print("${wl}", round(d["value"],1), "struct/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],1), d["e2e"].get("breakdown"), "| roofline", d["roofline"]["frac"], d["roofline"]["us_per_launch"], "| cpu", d["cpu_baseline"])
for k,v in list(d["kernel_shares"].items())[:8]: print("   ",k,v)
PY
done
