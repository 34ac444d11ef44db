"""MD steps/s of small cells (the common CHGNet use: relaxations / MD of 16-500 atoms): device driver with a Verlet skin
(one CUDA-graph replay per step) vs the host-driven calculator loop of the reference's structure (dynamics.py:129-181).
Run under gpurun."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chgnet_b200 import graphgen
from chgnet_b200.model import CHGNet
from chgnet_b200.dynamics import Atoms, CHGNetCalculator, VelocityVerlet
from chgnet_b200.dynamics_device import DeviceMD

model = CHGNet.from_file("tests/golden/chgnet_0.3.0_weights.npz", version="0.3.0").to("cuda")
out = []
for reps, steps in (((1, 1, 2), 400), ((2, 2, 2), 400), ((3, 3, 4), 200), ((5, 4, 6), 100)):
    z, frac, lat = graphgen.limno2_structure(reps, 0.02, 77)
    pos = frac @ lat
    rec = {"atoms": int(len(z))}
    for label, kw in (("device_skin0.5_cuda_graph", dict(skin=0.5)), ("device_skin0_rebuild_every_step", dict(skin=0.0))):
        md = DeviceMD(model, z, pos, lat, timestep=2.0, **kw)
        md.set_temperature(300.0, seed=1)
        md.run(20, log_every=0)
        torch.cuda.synchronize()
        b0 = md.n_builds
        t0 = time.perf_counter()
        md.run(steps, log_every=0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        rec[label] = {"ms_per_step": round(dt * 1e3, 4), "steps_per_s": round(1 / dt, 1), "graph_rebuilds": md.n_builds - b0, "steps": steps}
    host = VelocityVerlet(Atoms(z, pos, lat), CHGNetCalculator(model=model, on_isolated_atoms="ignore"), timestep=2.0)
    host.set_temperature(300.0, seed=1)
    host.run(5)
    hs = max(10, steps // 10)
    t0 = time.perf_counter()
    host.run(hs)
    dt = (time.perf_counter() - t0) / hs
    rec["host_calculator_loop"] = {"ms_per_step": round(dt * 1e3, 4), "steps_per_s": round(1 / dt, 1), "steps": hs}
    print(json.dumps(rec), flush=True)
    out.append(rec)
json.dump(out, open("gpurun_out/md_small_r2.json", "w"), indent=1)
