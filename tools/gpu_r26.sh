#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29526 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r26_bench_c3_2gpu.json 2> gpurun_out/r26_bench_c3_2gpu.err
echo "2gpu bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r26_bench_c3_2gpu.json') if l.startswith('{')][0])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['breakdown'])
    print('config', d['config']); print('rank0_share', d['rank0_share'])
    print('collective', d['collective'])
    print('parity', d['parity']); print('cpu', d['cpu_baseline'])
    print('c4', {k:v for k,v in d['c4'].items() if k in ('ms_per_step','atoms_per_s')}, d['c4']['e2e'])
    print('md', json.dumps(d['c4'].get('md'))[:1000])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r26_bench_c3_2gpu.err').read()[-3000:])
PY
tail -5 gpurun_out/r26_bench_c3_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r26_ref_2gpu.json 2> gpurun_out/r26_ref_2gpu.err
echo "reference arm (2 ranks) rc=$?"; grep '^{' gpurun_out/r26_ref_2gpu.json | cut -c1-600
