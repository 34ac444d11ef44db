"""CPU: the oracle (oracle/chgnet_oracle.py) against the committed outputs of the LIVE
reference (tests/golden/, written by oracle/make_golden.py) and against the known
answers of reference tests/test_model.py:60-119."""
import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen
from oracle import chgnet_oracle as orc


def _maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def test_oracle_reproduces_reference_known_answers(weights030, limno2_graph, golden):
    out = orc.predict_graph(weights030, limno2_graph, "efsm", return_site_energies=True, return_atom_feas=True,
                            return_crystal_feas=True)
    assert float(out["e"]) == pytest.approx(-7.36769, rel=1e-4, abs=1e-4)
    fz = [2.38135569e-02, -2.38130391e-02, 9.25870836e-02, -9.25877392e-02, -2.43449211e-03, -1.30698681e-02,
          1.30702555e-02, 2.43446976e-03]
    assert out["f"][:, 2] == pytest.approx(np.array(fz), rel=1e-3, abs=1e-4)
    assert np.diag(out["s"]) == pytest.approx(np.array([-3.0366361e-01, 2.2305478e-01, -1.0736181e-01]), rel=5e-3, abs=1e-4)
    assert out["m"][2] == pytest.approx(3.8694179, rel=1e-3)
    assert out["crystal_fea"].mean() == pytest.approx(0.26999, rel=1e-4, abs=1e-4)
    assert out["atom_fea"].mean() == pytest.approx(-0.09668, rel=1e-4, abs=1e-4)
    for k, tol in (("e", 2e-6), ("f", 5e-5), ("s", 1.5e-3), ("m", 1e-5), ("site_energies", 1e-5)):
        assert _maxabs(out[k], golden[f"limno2.ref32.{k}"]) < tol, k


def test_oracle_fp64_matches_committed_truth(weights030, limno2_graph, golden):
    out = orc.predict_graph(weights030, limno2_graph, "efsm", dtype=torch.float64)
    for k in "efsm":
        assert _maxabs(out[k], golden[f"limno2.oracle64.{k}"]) < 1e-9, k


def test_oracle_on_seeded_random_batch(weights030, golden):
    graphs = graphgen.random_graphs(4, 12, 20, 7000)
    preds = orc.predict_graph(weights030, graphs, "efsm", batch_size=4)
    for i, p in enumerate(preds):
        for k, tol in (("e", 5e-6), ("f", 2e-4), ("s", 3e-3), ("m", 1e-4)):
            assert _maxabs(p[k], golden[f"rand4.{i}.ref32.{k}"]) < tol, (i, k)


def test_graph_builder_counts():
    """reference tests/test_crystal_graph.py:22-42, 256-278"""
    z, frac, lat = graphgen.limno2_structure()
    g = graphgen.make_crystal_graph(z, frac, lat, atom_graph_cutoff=5.0, bond_graph_cutoff=3.0)
    assert (len(g.atom_graph), len(g.bond_graph), len(g.undirected2directed)) == (384, 744, 192)
    z, frac, lat = graphgen.limno2_structure((2, 3, 4))
    g = graphgen.make_crystal_graph(z, frac, lat, atom_graph_cutoff=5.0, bond_graph_cutoff=3.0)
    assert (len(g.atom_graph), len(g.bond_graph), len(g.undirected2directed)) == (9216, 17856, 4608)
    assert torch.all(g.atom_graph[1:, 0] >= g.atom_graph[:-1, 0]) and torch.all(g.bond_graph[1:, 1] >= g.bond_graph[:-1, 1])


def test_limno2_fixture_matches_builder(limno2_graph):
    z, frac, lat = graphgen.limno2_structure()
    g = graphgen.make_crystal_graph(z, frac, lat)
    for name in ("atom_graph", "directed2undirected", "undirected2directed", "bond_graph", "neighbor_image"):
        assert torch.equal(getattr(g, name), getattr(limno2_graph, name)), name
