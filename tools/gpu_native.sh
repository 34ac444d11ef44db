#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-nat}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_${TAG}.log
for wl in c2 c1 c4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  tail -c 400 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
print("${wl}", round(d["value"],1), "struct/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "ms", d["e2e"].get("breakdown"), "| launches", d["gpu_launches"], "| roofline", d["roofline"]["frac"])
PY
done
