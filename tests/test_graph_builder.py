"""CPU: the host-native C++ graph builder (chg_graph_build, csrc/graph_builder.cu) against the numpy
restatement of the reference's converter + Graph class (graphgen.py, itself pinned against the live
reference in tests/test_oracle_golden.py): every index array identical, row for row."""
import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen

FIELDS = ("atom_graph", "neighbor_image", "directed2undirected", "undirected2directed", "bond_graph")


def _same(z, frac, lat, **cut):
    a = graphgen.make_crystal_graph(z, frac, lat, backend="native", **cut)
    b = graphgen.make_crystal_graph(z, frac, lat, backend="numpy", **cut)
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), f
    return a


@pytest.mark.parametrize("n,seed", [(1, 11), (2, 12), (9, 13), (33, 14), (120, 15)])
def test_random_cells(n, seed):
    z, frac, lat = graphgen.random_structure(n, seed)
    g = _same(z, frac, lat)
    assert len(g.directed2undirected) == 2 * len(g.undirected2directed)  # crystalgraph.py:96-100


def test_reference_counts_and_cutoffs():
    z, frac, lat = graphgen.limno2_structure()
    g = _same(z, frac, lat, atom_graph_cutoff=5.0, bond_graph_cutoff=3.0)
    assert (len(g.atom_graph), len(g.bond_graph), len(g.undirected2directed)) == (384, 744, 192)  # tests/test_crystal_graph.py:22-42
    g = _same(z, frac, lat)
    assert (len(g.atom_graph), len(g.bond_graph), len(g.undirected2directed)) == (672, 744, 336)
    z, frac, lat = graphgen.limno2_structure((3, 2, 2), 0.03, 7)
    _same(z, frac, lat, atom_graph_cutoff=4.5, bond_graph_cutoff=2.2)


def test_edge_cases():
    # isolated atom (no edges), no bond graph, fractional coordinates outside [0, 1), triclinic cell that
    # needs several images per axis, an atom bonded to its own images
    g = _same([3], np.zeros((1, 3)), np.eye(3) * 20.0)
    assert len(g.atom_graph) == 0 and len(g.bond_graph) == 0
    g = _same([3, 8], np.array([[0.0, 0, 0], [0.5, 0.5, 0.5]]), np.eye(3) * 5.5)
    assert len(g.bond_graph) == 0 and len(g.atom_graph) > 0
    frac = np.array([[1.3, -0.2, 0.5], [0.1, 0.9, 2.2]])
    lat = np.array([[2.1, 0.1, 0.0], [0.3, 2.4, 0.2], [0.0, 0.5, 2.9]])
    g = _same([3, 8], frac, lat)
    assert (g.atom_graph[:, 0] == g.atom_graph[:, 1]).any()  # self-image bonds exist in a 2 A cell
    _same([26], np.array([[0.25, 0.25, 0.25]]), np.eye(3) * 2.5)


def test_bad_input_is_reported():
    from chgnet_b200._lib import ChgnetB200Error

    with pytest.raises(ChgnetB200Error):
        graphgen.native_graph_arrays(np.zeros((1, 3)), np.zeros((3, 3)), 6.0, 3.0)  # singular lattice


def _random_structures(n, seed, lo=5, hi=40):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        k = int(rng.integers(lo, hi + 1))
        a = (k / 0.1) ** (1 / 3)
        out.append((rng.integers(1, 90, k), rng.random((k, 3)), np.eye(3) * a + rng.normal(0, 0.05, (3, 3))))
    return out


def test_many_structures_at_once_equal_the_converter_loop():
    """chg_graph_build_many + chg_graph_views + the batch packers reading the builder's memory directly
    (CHGNet.structures_to_batch) against converting every structure (converter.py:102-190) and batching the
    CrystalGraphs: every field of the batch descriptor."""
    import dataclasses

    from chgnet_b200.batch import build_batch
    from chgnet_b200.model import CHGNet

    model = CHGNet.from_file("tests/golden/chgnet_0.3.0_weights.npz", version="0.3.0")
    model.graph_converter.on_isolated_atoms = "ignore"
    structs = _random_structures(37, 3) + [([3], np.zeros((1, 3)), np.eye(3) * 20.0)]  # + one isolated atom, no edges
    a = model.structures_to_batch(structs)
    b = build_batch([model.graph_converter(s) for s in structs], "cpu")
    n = 0
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if isinstance(x, torch.Tensor):
            assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), f.name
            n += 1
        elif f.name != "h2d_bytes":
            assert x == y, f.name
    assert n >= 30 and a.n_graphs == 38
    # isolated atoms are reported like the converter does (converter.py:160-174)
    model.graph_converter.on_isolated_atoms = "error"
    with pytest.raises(ValueError, match="isolated atom"):
        model.structures_to_batch(structs)
    # a failing structure fails the call with the builder's message
    from chgnet_b200._lib import ChgnetB200Error

    with pytest.raises(ChgnetB200Error, match="singular lattice"):
        model.structures_to_batch(structs[:3] + [([3], np.zeros((1, 3)), np.zeros((3, 3)))])


def test_convert_many_equals_the_loop():
    from chgnet_b200.model import GraphConverter

    gc = GraphConverter()
    structs = _random_structures(12, 5, lo=12)
    for x, y in zip(gc.convert_many(structs, n_threads=4), [gc(s) for s in structs]):
        assert torch.equal(x.atom_graph, y.atom_graph) and torch.equal(x.bond_graph, y.bond_graph)
        assert torch.equal(x.neighbor_image, y.neighbor_image) and torch.equal(x.undirected2directed, y.undirected2directed)


# ---- known answers of the reference's own converter tests (tests/test_crystal_graph.py; cutoffs 5 A / 3 A) -------------
_KNOWN = [
    # strain per lattice vector (pymatgen apply_strain), (edges, angles, bonds), {(column, atom): count} for atom_graph
    ((0.0, 0.0, 0.0), (384, 744, 192), {(0, 0): 48, (1, 0): 48, (0, 4): 48, (0, 7): 48}),     # :22-42
    ((0.1, 0.1, 0.1), (264, 288, 132), {(0, 3): 34, (1, 3): 34, (0, 7): 32}),                 # :174-211
    ((0.2, -0.3, 0.5), (336, 256, 168), {(0, 3): 42, (1, 3): 42, (0, 7): 42}),                # :214-253
]


@pytest.mark.parametrize("backend", ["native", "numpy"])
@pytest.mark.parametrize("strain,sizes,counts", _KNOWN)
def test_reference_known_answers_strained_cells(backend, strain, sizes, counts):
    z, frac, lat = graphgen.limno2_structure()
    cell = lat * (1.0 + np.asarray(strain))[:, None]  # apply_strain scales lattice vector i by 1 + strain_i
    g = graphgen.make_crystal_graph(z, frac, cell, atom_graph_cutoff=5.0, bond_graph_cutoff=3.0, backend=backend)
    assert g.atomic_number.tolist() == [3, 3, 25, 25, 8, 8, 8, 8]
    assert (len(g.atom_graph), len(g.bond_graph), len(g.undirected2directed)) == sizes
    assert len(g.directed2undirected) == sizes[0] and list(g.lattice.shape) == [3, 3] and list(g.atom_frac_coord.shape) == [8, 3]
    for (col, atom), want in counts.items():
        assert int((g.atom_graph[:, col] == atom).sum()) == want, (col, atom)
    if strain == (0.0, 0.0, 0.0):
        assert int((g.bond_graph[:, 0] == 1).sum()) == 72  # :36


def test_reference_stability_invariants_under_large_perturbations():
    """tests/test_crystal_graph.py:306-335: 2x2x2 supercells with every atom displaced by 0.5 A still give complete
    undirected bonds (the reference's converter falls back to the legacy algorithm when they do not)."""
    rng = np.random.default_rng(2024)
    for trial in range(20):
        z, frac, lat = graphgen.limno2_structure((2, 2, 2), 0.0, 0)
        step = rng.standard_normal((len(z), 3))
        step *= 0.5 / np.linalg.norm(step, axis=1, keepdims=True)  # Structure.perturb(distance=0.5)
        g = _same(z, frac + step @ np.linalg.inv(lat), lat, atom_graph_cutoff=5.0, bond_graph_cutoff=3.0)
        assert g.directed2undirected.shape[0] == 2 * g.undirected2directed.shape[0]
        assert g.atom_graph.shape[0] == g.directed2undirected.shape[0]
