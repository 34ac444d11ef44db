"""Host logic: the kernel schedule (chgnet_b200/engine.py) + batch builder, executed
with the torch kernel specifications on the CPU, must reproduce the oracle's
autograd results.  This validates every analytic reverse formula the CUDA kernels
transcribe, with no GPU involved."""
import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen
from chgnet_b200.batch import build_batch
from chgnet_b200.engine import EV_A3_TO_GPA, Engine
from chgnet_b200.weights import pack_weights
from oracle import chgnet_oracle as orc
from oracle.kernel_specs import SpecKernels


def run_engine(weights, graphs, dtype, compact=True, model_args=None, **kw):
    sd = {k: torch.as_tensor(v) for k, v in weights.items()}
    pw = pack_weights(sd, model_args, device="cpu", dtype=dtype)
    b = build_batch(graphs, "cpu", compact_bonds=compact)
    for name in ("frac", "image", "lattice", "volume"):
        setattr(b, name, getattr(b, name).to(dtype))
    if dtype == torch.float64:  # fp64 truth needs fp64 geometry inputs
        b.frac = torch.cat([g.atom_frac_coord.detach().double() for g in graphs])
        b.lattice = torch.stack([g.lattice.detach().double().reshape(9) for g in graphs])
        L = b.lattice.view(-1, 3, 3)
        b.volume = (L[:, 0] * torch.linalg.cross(L[:, 1], L[:, 2])).sum(dim=1)
    return b, Engine(pw, SpecKernels()).run(b, **kw)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_schedule_matches_oracle_autograd(weights030, dtype, tol):
    graphs = graphgen.random_graphs(3, 8, 14, 9100)
    b, out = run_engine(weights030, graphs, dtype, need_grad=True, need_magmom=True, need_atom_fea=True,
                        need_crystal_fea=True, keep_intermediates=True)
    ref = orc.forward(weights030, graphs, "efsm", dtype=dtype, return_site_energies=True,
                      return_atom_feas=True, return_crystal_feas=True, return_intermediates=True)
    n = torch.tensor(b.atoms_per_graph)
    e = (out.energy + out.e_ref) / n
    # AtomRef is evaluated in fp32 by the reference even for an fp64 model (composition_model.py:191)
    assert torch.allclose(e, ref["e"].double(), atol=max(tol * 10, 2e-6), rtol=0)
    e_model = torch.zeros(len(graphs), dtype=torch.float64).index_add_(
        0, b.owner.long(), ref["intermediates"]["site_e_model"].double())
    assert torch.allclose(out.energy, e_model, atol=tol * 100, rtol=0)
    f_ref = torch.cat(ref["f"]).double()
    assert torch.allclose(out.force, f_ref, atol=tol * 10, rtol=0), (out.force - f_ref).abs().max()
    s = out.virial.view(-1, 3, 3) * (EV_A3_TO_GPA / b.volume.double())[:, None, None]
    s_ref = torch.stack(ref["s"]).double()
    assert torch.allclose(s, s_ref, atol=tol * 100, rtol=0), (s - s_ref).abs().max()
    assert torch.allclose(out.magmom.double(), torch.cat(ref["m"]).double(), atol=tol * 10)
    assert torch.allclose(out.atom_fea.double(), torch.cat(ref["atom_fea"]).double(), atol=tol * 10)
    assert torch.allclose(out.crystal_fea.double(), ref["crystal_fea"].double(), atol=tol * 100)
    for k, v in ref["intermediates"].items():
        if v is not None and k in out.extras and out.extras[k] is not None:
            assert torch.allclose(out.extras[k].double(), v.double(), atol=tol * 10, rtol=tol * 10), k


def test_unsorted_graph_is_reordered(weights030):
    """CrystalGraph does not guarantee center-sorted edges / i-sorted angles."""
    g = graphgen.random_graphs(1, 10, 10, 9200)[0]
    rng = np.random.default_rng(0)
    pe = torch.from_numpy(rng.permutation(len(g.atom_graph)))
    inv = torch.empty_like(pe)
    inv[pe] = torch.arange(len(pe))
    pa = torch.from_numpy(rng.permutation(len(g.bond_graph)))
    bg = g.bond_graph[pa].clone()
    bg[:, 2] = inv[bg[:, 2].long()].int()
    bg[:, 4] = inv[bg[:, 4].long()].int()
    from chgnet_b200.graph import CrystalGraph

    g2 = CrystalGraph(
        atomic_number=g.atomic_number, atom_frac_coord=g.atom_frac_coord, atom_graph=g.atom_graph[pe],
        atom_graph_cutoff=6.0, neighbor_image=g.neighbor_image[pe], directed2undirected=g.directed2undirected[pe],
        undirected2directed=inv[g.undirected2directed.long()].int(), bond_graph=bg, bond_graph_cutoff=3.0,
        lattice=g.lattice)
    _, o1 = run_engine(weights030, [g], torch.float64, need_grad=True)
    _, o2 = run_engine(weights030, [g2], torch.float64, need_grad=True)
    assert torch.allclose(o1.energy, o2.energy, atol=1e-10)
    assert torch.allclose(o1.force, o2.force, atol=1e-10)
    assert torch.allclose(o1.virial, o2.virial, atol=1e-9)


def test_no_angles_and_isolated_atom(weights030):
    """Empty bond graph (model.py:438,460) and an atom with no edges (model.py:841-843)."""
    z, frac, lat = [3, 8], np.array([[0.0, 0, 0], [0.5, 0.5, 0.5]]), np.eye(3) * 5.5
    g = graphgen.make_crystal_graph(z, frac, lat)  # nearest distance 4.76 A > 3 A: no angles
    assert len(g.bond_graph) == 0 and len(g.atom_graph) > 0
    _, out = run_engine(weights030, [g], torch.float64, need_grad=True, need_magmom=True)
    ref = orc.forward(weights030, [g], "efsm", dtype=torch.float64)
    assert torch.allclose((out.energy + out.e_ref) / 2, ref["e"], atol=2e-6)
    assert torch.allclose(out.force, torch.cat(ref["f"]), atol=1e-9)
    g_iso = graphgen.make_crystal_graph([3], np.zeros((1, 3)), np.eye(3) * 20.0)
    assert len(g_iso.atom_graph) == 0
    _, out = run_engine(weights030, [g_iso, g], torch.float64, need_grad=True)
    ref = orc.forward(weights030, [g_iso, g], "efs", dtype=torch.float64)
    assert torch.allclose((out.energy + out.e_ref) / torch.tensor([1, 2]), ref["e"], atol=2e-6)
    assert torch.allclose(out.force, torch.cat(ref["f"]), atol=1e-9)


def test_identity_compaction_gives_same_result(weights030):
    graphs = graphgen.random_graphs(2, 8, 12, 9500)
    b1, o1 = run_engine(weights030, graphs, torch.float64, compact=True, need_grad=True)
    b2, o2 = run_engine(weights030, graphs, torch.float64, compact=False, need_grad=True)
    assert b1.n_short < b2.n_short == b2.n_bonds
    assert torch.allclose(o1.energy, o2.energy, atol=1e-11) and torch.allclose(o1.force, o2.force, atol=1e-11)
    assert torch.allclose(o1.virial, o2.virial, atol=1e-10)


def test_v020_shaped_model_no_layernorm_bias_small_basis():
    """0.2.0 architecture: 9 radial / 9 angular functions, no LayerNorm anywhere, mlp_out bias,
    two readout hidden layers, cutoff_coeff 5, atom-graph cutoff 5 A."""
    args = dict(num_radial=9, num_angular=9, gMLP_norm=None, readout_norm=None, mlp_out_bias=True, cutoff_coeff=5,
                atom_graph_cutoff=5.0)
    w = orc.random_weights(3, args)
    w = {k: v for k, v in w.items() if not k.startswith("mlp.layers.4.")}
    w["mlp.layers.5.weight"], w["mlp.layers.5.bias"] = w.pop("mlp.layers.7.weight"), w.pop("mlp.layers.7.bias")
    graphs = graphgen.random_graphs(2, 8, 12, 9400, atom_graph_cutoff=5.0)
    b, out = run_engine(w, graphs, torch.float64, compact=False, model_args=dict(atom_graph_cutoff=5.0, cutoff_coeff=5),
                        need_grad=True, need_magmom=True)
    ref = orc.forward(w, graphs, "efsm", dtype=torch.float64, args=args)
    assert torch.allclose(out.force, torch.cat(ref["f"]), atol=1e-10)
    assert torch.allclose(out.magmom, torch.cat(ref["m"]), atol=1e-10)
    s = out.virial.view(-1, 3, 3) * (EV_A3_TO_GPA / b.volume.double())[:, None, None]
    assert torch.allclose(s, torch.stack(ref["s"]), atol=1e-9)
