#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "limno2_known or batched_equals or native_forward" 2>&1 | tail -2
timeout 120 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_pack.json 2> gpurun_out/bench_c3_pack.err
tail -c 200 gpurun_out/bench_c3_pack.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c3_pack.json").read().strip().splitlines()[-1])
print("c3", round(d["value"],1), round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "ms", d["e2e"]["breakdown"])
PY
