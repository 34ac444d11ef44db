#!/bin/bash
# A/B of the gated implementations on the bench workloads (no CPU baseline leg)
mkdir -p gpurun_out
for impl in 0 2 1; do
for wl in c2 c4; do
  CHG_GATED_IMPL=$impl timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_gated_${impl}.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_gated_${impl}.json").read().strip().splitlines()[-1])
ks=d["kernel_shares"]
print("${impl} ${wl}", round(d["ms_per_step"],3), "ms |", " ".join(f"{k}={v['ms']:.2f}" for k,v in ks.items() if k in ("atom_conv_fwd","atom_conv_bwd","bond_conv_fwd","bond_conv_bwd","angle_update_fwd","angle_update_bwd","linear","segment_sum")))
PY
done; done
