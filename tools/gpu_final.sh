#!/bin/bash
# what the driver runs at round end: smoke(), pytest -m gpu, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "initialized with" | tail -4 | tee gpurun_out/smoke_final.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err
tail -c 300 gpurun_out/bench_default_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default_final.json").read().strip().splitlines()[-1])
print("default", d["metric"], round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],1), "| launches", d["gpu_launches"], "| roofline", d["roofline"]["frac"], "| cpu", d["cpu_baseline"]["value"], "| torch_cuda", d.get("torch_cuda_baseline"))
PY
