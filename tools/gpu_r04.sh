#!/bin/bash
# full GPU test suite + inference bench (c2, c4) + training bench (c5)
mkdir -p gpurun_out
TAG=${1:-r04}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_${TAG}.log
for wl in c2 c4 c5; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  tail -c 600 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
print("${wl}", d["metric"], round(d["value"],1), "struct/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "ms | launches", d["gpu_launches"], "| roofline", d["roofline"]["frac"], d.get("last_report"), d.get("breakdown"))
for k,v in list(d["kernel_shares"].items())[:10]: print("   ",k,v)
PY
done
