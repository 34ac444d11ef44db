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


def test_native_host_packer_equals_torch_packing():
    """chg_pack_batch_host (one C pass: concatenation, offsets, owners, sortedness flags) builds the same
    DeviceBatch as the tensor-op path, also for unsorted graphs, reshaped empty graphs and fp64 inputs
    (which fall back to the tensor-op path)."""
    import dataclasses

    from chgnet_b200.graph import CrystalGraph

    graphs = graphgen.random_graphs(6, 4, 12, 6200)
    graphs.append(graphgen.make_crystal_graph([3], np.zeros((1, 3)), np.eye(3) * 20.0))
    g = graphs[2]
    perm = torch.randperm(len(g.atom_graph), generator=torch.Generator().manual_seed(0))  # unsorted edges
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(len(perm))
    bgp = g.bond_graph.clone()
    bgp[:, 2], bgp[:, 4] = inv[bgp[:, 2].long()].int(), inv[bgp[:, 4].long()].int()
    graphs[2] = CrystalGraph(atomic_number=g.atomic_number, atom_frac_coord=g.atom_frac_coord, atom_graph=g.atom_graph[perm].contiguous(),
                             atom_graph_cutoff=6.0, neighbor_image=g.neighbor_image[perm].contiguous(),
                             directed2undirected=g.directed2undirected[perm].contiguous(),
                             undirected2directed=inv[g.undirected2directed.long()].int(), bond_graph=bgp,
                             bond_graph_cutoff=3.0, lattice=g.lattice)

    def same(gs):
        a, b = build_batch(gs, "cpu", native_pack=True), build_batch(gs, "cpu", native_pack=False)
        for f in dataclasses.fields(a):
            x, y = getattr(a, f.name), getattr(b, f.name)
            if torch.is_tensor(x):
                assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), f.name
            elif f.name != "h2d_bytes":
                assert x == y, f.name
        return a

    same(graphs)
    same(graphs[-1:])  # only an isolated atom
    g64 = graphs[0]
    g64 = CrystalGraph(**{**g64.to_dict(), "atom_frac_coord": g64.atom_frac_coord.double(), "lattice": g64.lattice.double()})
    assert same([g64, graphs[1]]).frac.dtype == torch.float32  # not packable by memcpy -> converted by the tensor-op path
