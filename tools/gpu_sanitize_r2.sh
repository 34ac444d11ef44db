#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels: default path (fused tcgen05 forward + ws reverse, device CSR),
# device graph builder, device MD step, on small inputs.
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from chgnet_b200 import graphgen
from chgnet_b200.model import CHGNet
m = CHGNet.from_file("tests/golden/chgnet_0.3.0_weights.npz", version="0.3.0").to("cuda")
gs = graphgen.random_graphs(3, 8, 14, 9300)
out = m.predict_graph(gs, task="efsm", return_site_energies=True, return_crystal_feas=True)
torch.cuda.synchronize()
print("default impl", [float(o["e"]) for o in out])
z, frac, lat = graphgen.limno2_structure((2, 2, 1), 0.03, 4200)
p = m.predict_structure((z, frac, lat), task="efs")  # device graph builder + device CSR
print("predict_structure (device graph)", float(p["e"]))
from chgnet_b200.dynamics_device import DeviceMD
md = DeviceMD(m, z, frac @ lat, lat, timestep=1.0, skin=0.0)
md.set_temperature(300.0, seed=1)
md.run(2, log_every=0)
md2 = DeviceMD(m, z, frac @ lat, lat, timestep=1.0, skin=0.5, use_cuda_graph=False)
md2.set_temperature(300.0, seed=1)
md2.run(2, log_every=0)
torch.cuda.synchronize()
print("md ok", md.potential_energy, md2.potential_energy)
# list of structures: chg_graph_build_many -> chg_pack_batch_wire (two-phase copies, expand_image / derive_angle_columns)
structs = [graphgen.random_structure(n, 9400 + n) for n in (9, 14, 11)]
ps = m.predict_structure(structs, task="efs", batch_size=2)
print("predict_structure (list)", [float(p["e"]) for p in ps])
# one training step whose angle-level weight gradients take the tcgen05 wgrad kernel (>= 4096 reduction rows)
from chgnet_b200.trainer import Trainer
graphs = graphgen.random_graphs(2, 20, 24, 4243)
base = m.predict_graph(graphs, task="efsm", batch_size=2)
lab = {"e": torch.tensor([float(p["e"]) + 0.05 for p in base]), "f": [torch.as_tensor(p["f"]) + 0.02 for p in base],
       "s": [torch.as_tensor(p["s"]) - 0.05 for p in base], "m": [torch.as_tensor(p["m"]) + 0.03 for p in base]}
tr = Trainer(m, targets="efsm", criterion="MSE", learning_rate=1e-5)
print("train_step", tr.train_step(graphs, lab), "angles", sum(int(g.bond_graph.shape[0]) for g in graphs))
torch.cuda.synchronize()
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san2.py > gpurun_out/sanitizer_memcheck_r2.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck_r2.log
tail -12 gpurun_out/sanitizer_memcheck_r2.log
