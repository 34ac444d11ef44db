"""CPU: packed graph store round trip, and that its zero-copy graphs feed the batch
builder like freshly built ones."""
import warnings

import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen
from chgnet_b200.batch import build_batch
from chgnet_b200.graphfile import GraphStore, load_graphs, save_graphs

FIELDS = ("atomic_number", "atom_frac_coord", "atom_graph", "neighbor_image", "directed2undirected",
          "undirected2directed", "bond_graph", "lattice")


def test_round_trip_and_batching(tmp_path):
    graphs = graphgen.random_graphs(5, 4, 14, 6100)
    graphs.append(graphgen.make_crystal_graph([3], np.zeros((1, 3)), np.eye(3) * 20.0, graph_id="isolated"))  # no edges
    graphs[1].mp_id, graphs[2].composition = "mp-1", "Li2O"
    path = save_graphs(str(tmp_path / "set.chgpack"), graphs)
    store = GraphStore(path)
    assert len(store) == 6 and store[-1].graph_id == "isolated" and store[1].mp_id == "mp-1" and store[2].composition == "Li2O"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch warns about non-writable numpy views
        back = load_graphs(path)
        for a, b in zip(graphs, back):
            for f in FIELDS:
                x, y = getattr(a, f), getattr(b, f)
                assert x.dtype == y.dtype and torch.equal(x.reshape(y.shape), y), f
            assert (a.atom_graph_cutoff, a.bond_graph_cutoff, a.graph_id) == (b.atom_graph_cutoff, b.bond_graph_cutoff, b.graph_id)
        b1, b2 = build_batch(graphs, "cpu"), build_batch(store.batch([0, 1, 2, 3, 4, 5]), "cpu")
    for name in ("z", "frac", "center", "nbr", "image", "d2u", "u2d", "ang_atom", "ang_is", "ang_js", "ptr_c", "short_ids", "lattice"):
        assert torch.equal(getattr(b1, name), getattr(b2, name)), name
    with pytest.raises(IndexError):
        store[6]
    (tmp_path / "bad").write_bytes(b"not a pack")
    with pytest.raises(ValueError):
        GraphStore(str(tmp_path / "bad"))
