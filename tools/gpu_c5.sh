#!/bin/bash
# training kernels tests + c5 bench
mkdir -p gpurun_out
TAG=${1:-c5}
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -x -q -k "training or trainer or second_order" 2>&1 | tail -4
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_${TAG}.json 2> gpurun_out/bench_c5_${TAG}.err
tail -c 600 gpurun_out/bench_c5_${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c5_${TAG}.json").read().strip().splitlines()[-1])
print("c5", d["metric"], round(d["value"],1), "struct/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "ms | launches", d["gpu_launches"], d.get("breakdown"))
for k,v in list(d["kernel_shares"].items())[:14]: print("   ",k,v)
PY
