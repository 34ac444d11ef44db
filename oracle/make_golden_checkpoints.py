"""TEST INFRASTRUCTURE — goldens for the other two shipped checkpoints (0.2.0, r2scan).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_checkpoints

For each of the reference's pretrained models besides 0.3.0 (reference
chgnet/model/model.py:718-736: ``0.2.0`` = 9 radial / 9 angular basis functions, no
LayerNorm, mlp_out bias, 5 A atom-graph cutoff, AtomRef "MPtrj_e"; ``r2scan`` = the
0.3.0 architecture with the "MP-r2SCAN" AtomRef) this writes

* ``tests/golden/chgnet_<name>_weights.npz`` — the checkpoint's ``state_dict`` as plain arrays plus a
  JSON copy of its ``model_args`` (entry ``__model_args__``), loadable by
  ``chgnet_b200.CHGNet.load(model_name=<name>)`` where /root/reference does not exist (the GPU box);
* ``tests/golden/chgnet_checkpoints_golden.npz`` — outputs of the UNMODIFIED reference
  ``CHGNet.predict_graph`` (fp32, CPU, task "efsm" + site energies) on LiMnO2 mp-18767 and on a seeded
  random batch (3 cells, 10..16 atoms), built with the checkpoint's own cutoffs, and the fp64 oracle
  outputs on the same graphs.

It also asserts that ``oracle/chgnet_oracle.py`` run with each checkpoint's weights and arguments agrees
with the live reference — which pins the oracle for these two architectures as make_golden.py does
for 0.3.0.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chgnet_b200 import graphgen  # noqa: E402
from oracle import chgnet_oracle as orc  # noqa: E402
from oracle.make_golden import GOLD, maxabs, pack_pred, to_ref_graph  # noqa: E402
from oracle.ref_import import load_reference_model  # noqa: E402

NAMES = ("0.2.0", "r2scan")
ORACLE_KEYS = ("num_radial", "num_angular", "gMLP_norm", "readout_norm", "mlp_out_bias", "cutoff_coeff",
               "atom_graph_cutoff", "bond_graph_cutoff", "n_conv", "is_intensive")


def oracle_args(model_args: dict) -> dict:
    a = {k: model_args[k] for k in ORACLE_KEYS if k in model_args}
    a["atom_graph_cutoff"] = float(a["atom_graph_cutoff"])
    a["bond_graph_cutoff"] = float(a["bond_graph_cutoff"])
    return a


def main() -> None:
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    fix: dict = {}
    kw = dict(return_site_energies=True)
    for name in NAMES:
        model = load_reference_model(name)
        sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        margs = {k: v for k, v in model.model_args.items() if isinstance(v, (int, float, str, bool, list, type(None)))}
        np.savez_compressed(os.path.join(GOLD, f"chgnet_{name}_weights.npz"), __model_args__=np.array(json.dumps(margs)), **sd)
        args = oracle_args(model.model_args)
        cut = dict(atom_graph_cutoff=args["atom_graph_cutoff"], bond_graph_cutoff=args["bond_graph_cutoff"])
        z, frac, lat = graphgen.limno2_structure()
        graphs = [graphgen.make_crystal_graph(z, frac, lat, graph_id="mp-18767", backend="numpy", **cut)]
        graphs += graphgen.random_graphs(3, 10, 16, 7900, backend="numpy", **cut)
        ref = model.predict_graph([to_ref_graph(g) for g in graphs], task="efsm", batch_size=len(graphs), **kw)
        o32 = orc.predict_graph(sd, graphs, "efsm", batch_size=len(graphs), args=args, **kw)
        o64 = orc.predict_graph(sd, graphs, "efsm", batch_size=len(graphs), dtype=torch.float64, args=args, **kw)
        for i, (r, a, b) in enumerate(zip(ref, o32, o64)):
            pack_pred(f"{name}.{i}.ref32", r, fix)
            pack_pred(f"{name}.{i}.oracle64", b, fix)
            print(f"{name}[{i}] n={len(r['m'])} oracle32 vs ref32:", {k: f"{maxabs(a[k], r[k]):.2e}" for k in r})
            print(f"{name}[{i}]      ref32 vs oracle64:", {k: f"{maxabs(b[k], r[k]):.2e}" for k in r})
            assert maxabs(a["e"], r["e"]) < 5e-6 and maxabs(a["f"], r["f"]) < 2e-4
            assert maxabs(a["s"], r["s"]) < 3e-3 and maxabs(a["m"], r["m"]) < 2e-5
        print(f"{name}: {len(sd)} tensors, {sum(v.size for v in sd.values())} params, e(LiMnO2) = {float(ref[0]['e']):.6f}")
    np.savez_compressed(os.path.join(GOLD, "chgnet_checkpoints_golden.npz"), **fix)
    print("wrote", os.path.join(GOLD, "chgnet_checkpoints_golden.npz"), len(fix), "arrays")


if __name__ == "__main__":
    main()
