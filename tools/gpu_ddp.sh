#!/bin/bash
# 2-GPU check of the data-parallel training step (NCCL): tools/ddp_check.py + bench c5 --gpus 2
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1
grep -v "initialized with" gpurun_out/ddp_check.log | grep -E "DDP CHECK|ranks_identical|Error|assert" | tail -6
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_c5_2gpu.err | tail -1 > gpurun_out/bench_c5_2gpu.json
tail -c 300 gpurun_out/bench_c5_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c5_2gpu.json").read().strip().splitlines()[-1])
print("c5 x2", d["metric"], round(d["value"],1), "struct/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"],1))
PY
