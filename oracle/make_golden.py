"""TEST INFRASTRUCTURE — writes tests/golden/* from the LIVE reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

What it does
1. exports the pretrained 0.3.0 state_dict to ``tests/golden/chgnet_0.3.0_weights.npz``
   (pure data; the kernels and the oracle both read this file);
2. builds the LiMnO2 mp-18767 CrystalGraph with ``chgnet_b200.graphgen`` and checks
   it row-for-row against the reference's own ``Graph`` class
   (reference chgnet/graph/graph.py:132-328);
3. runs the UNMODIFIED reference ``CHGNet.predict_graph`` (fp32, CPU) on LiMnO2 and
   on a seeded random batch and stores its outputs; also stores the oracle's fp64
   outputs (error-budget truth) and per-layer fp32 intermediates;
4. asserts that ``oracle/chgnet_oracle.py`` (fp32) agrees with the live reference
   and that the reference reproduces the known answers of reference
   tests/test_model.py:60-119 — this is what pins the oracle.
"""
from __future__ import annotations

import hashlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chgnet_b200 import graphgen  # noqa: E402
from oracle import chgnet_oracle as orc  # noqa: E402
from oracle.ref_import import import_reference, load_reference_model  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# reference tests/test_model.py:68-119
KNOWN_E = -7.36769
KNOWN_F_Z = [2.38135569e-02, -2.38130391e-02, 9.25870836e-02, -9.25877392e-02,
             -2.43449211e-03, -1.30698681e-02, 1.30702555e-02, 2.43446976e-03]
KNOWN_S_DIAG = [-3.0366361e-01, 2.2305478e-01, -1.0736181e-01]
KNOWN_M = [3.0495524e-03, 3.0494630e-03, 3.8694179e00, 3.8694181e00,
           4.4136152e-02, 3.8622141e-02, 3.8622111e-02, 4.4136211e-02]
KNOWN_SITE_E = [-3.6264274, -3.6264274, -9.634681, -9.634682,
                -8.024935, -8.184724, -8.184724, -8.024935]


def graph_digest(g) -> str:
    h = hashlib.sha256()
    for name in ("atomic_number", "atom_graph", "directed2undirected", "undirected2directed", "bond_graph"):
        h.update(np.ascontiguousarray(getattr(g, name).numpy()).tobytes())
    h.update(np.ascontiguousarray(g.neighbor_image.numpy().astype(np.int8)).tobytes())
    return h.hexdigest()[:16]


def to_ref_graph(g):
    from chgnet.graph.crystalgraph import CrystalGraph as RefGraph

    return RefGraph(**g.to_dict())


def pack_pred(prefix: str, pred: dict, out: dict) -> None:
    for k, v in pred.items():
        out[f"{prefix}.{k}"] = np.asarray(v)


def maxabs(a, b) -> float:
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))) if np.size(a) else 0.0


def main() -> None:
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    os.makedirs(GOLD, exist_ok=True)
    import_reference()
    from chgnet.graph.graph import Graph, Node

    model = load_reference_model("0.3.0")
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, "chgnet_0.3.0_weights.npz"), **sd)
    print(f"weights: {len(sd)} tensors, {sum(v.size for v in sd.values())} params")

    # ---- LiMnO2: graph parity with the reference builder ----
    z, frac, lat = graphgen.limno2_structure()
    c, n, img, d = graphgen.neighbor_list(frac, lat, 6.0)
    ag, d2u, u2d, bg = graphgen.build_graph_arrays(c, n, img, d, 3.0)
    rg = Graph([Node(index=i) for i in range(len(z))])
    for ii, jj, im, dd in zip(c, n, img, d):
        rg.add_edge(center_index=ii, neighbor_index=jj, image=im, distance=dd)
    rag, rd2u = rg.adjacency_list()
    rbg, ru2d = rg.line_graph_adjacency_list(cutoff=3.0)
    assert np.array_equal(np.array(rag), ag) and np.array_equal(np.array(rd2u), d2u)
    assert np.array_equal(np.array(ru2d), u2d) and np.array_equal(np.array(rbg), bg)
    assert (len(ag), len(bg), len(u2d)) == (672, 744, 336)
    g1 = graphgen.make_crystal_graph(z, frac, lat, graph_id="mp-18767")

    fix: dict = {}
    kw = dict(return_site_energies=True, return_atom_feas=True, return_crystal_feas=True)
    ref = model.predict_graph(to_ref_graph(g1), task="efsm", **kw)
    # the live reference must reproduce its own published known answers
    assert abs(float(ref["e"]) - KNOWN_E) < 1e-4
    assert maxabs(ref["f"][:, 2], KNOWN_F_Z) < 1e-4
    assert maxabs(np.diag(ref["s"]), KNOWN_S_DIAG) < 5e-3 * 0.31 + 1e-4
    assert maxabs(ref["m"], KNOWN_M) < 1e-4 + 1e-3 * 3.9
    assert maxabs(ref["site_energies"], KNOWN_SITE_E) < 1e-3
    assert abs(ref["crystal_fea"].mean() - 0.26999) < 1e-4 and abs(ref["atom_fea"].mean() + 0.09668) < 1e-4
    pack_pred("limno2.ref32", ref, fix)
    o32 = orc.predict_graph(sd, g1, "efsm", **kw)
    o64 = orc.predict_graph(sd, g1, "efsm", dtype=torch.float64, **kw)
    pack_pred("limno2.oracle64", o64, fix)
    print("LiMnO2  oracle32 vs ref32:", {k: f"{maxabs(o32[k], ref[k]):.2e}" for k in ref})
    print("LiMnO2  ref32 vs oracle64:", {k: f"{maxabs(o64[k], ref[k]):.2e}" for k in ref})
    assert maxabs(o32["e"], ref["e"]) < 2e-6 and maxabs(o32["f"], ref["f"]) < 5e-5
    assert maxabs(o32["s"], ref["s"]) < 1.5e-3 and maxabs(o32["m"], ref["m"]) < 1e-5
    for name in ("atomic_number", "atom_frac_coord", "atom_graph", "neighbor_image",
                 "directed2undirected", "undirected2directed", "bond_graph", "lattice"):
        fix[f"limno2.graph.{name}"] = getattr(g1, name).detach().numpy()
    inter = orc.forward(sd, [g1], "e", return_intermediates=True)["intermediates"]
    for k, v in inter.items():
        if v is not None:
            fix[f"limno2.inter32.{k}"] = v.numpy()

    # ---- seeded random batch: 4 cells, 12..20 atoms ----
    graphs = graphgen.random_graphs(4, 12, 20, 7000)
    fix["rand4.digest"] = np.array([graph_digest(g) for g in graphs])
    refb = model.predict_graph([to_ref_graph(g) for g in graphs], task="efsm", batch_size=4, **kw)
    o32b = orc.predict_graph(sd, graphs, "efsm", batch_size=4, **kw)
    o64b = orc.predict_graph(sd, graphs, "efsm", batch_size=4, dtype=torch.float64, **kw)
    for i, (r, a, b) in enumerate(zip(refb, o32b, o64b)):
        pack_pred(f"rand4.{i}.ref32", r, fix)
        pack_pred(f"rand4.{i}.oracle64", b, fix)
        print(f"rand4[{i}] n={len(r['m'])} oracle32 vs ref32:", {k: f"{maxabs(a[k], r[k]):.2e}" for k in r})
        print(f"rand4[{i}]      ref32 vs oracle64:", {k: f"{maxabs(b[k], r[k]):.2e}" for k in r})
        assert maxabs(a["e"], r["e"]) < 5e-6 and maxabs(a["f"], r["f"]) < 2e-4 and maxabs(a["s"], r["s"]) < 3e-3

    np.savez_compressed(os.path.join(GOLD, "chgnet_0.3.0_golden.npz"), **fix)
    print("wrote", os.path.join(GOLD, "chgnet_0.3.0_golden.npz"), len(fix), "arrays")


if __name__ == "__main__":
    main()
