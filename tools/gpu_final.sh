#!/bin/bash
# final measurement pass of the round: GPU tests, default bench (with CPU baseline), c4/c3/c1, reference arm
set -u
mkdir -p gpurun_out
TAG=${1:-r03}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err
for wl in c4 c3; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c2_${TAG}.json 2>> gpurun_out/bench_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_${TAG}.log
for f in gpurun_out/bench_c2_${TAG}.json gpurun_out/bench_c4_${TAG}.json gpurun_out/bench_c3_${TAG}.json; do python - <<PY
import json
d=json.loads(open("$f").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), d["e2e"]["breakdown"], "roof", d["roofline"]["frac"], d["roofline"].get("at_10k_atoms",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
done
