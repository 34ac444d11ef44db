"""CPU: the training reverse pass (parameter gradients) of the kernel schedule, with the torch
kernel specifications injected, against autograd through the oracle in fp64
(reference trainer.py:398-411: prediction -> CombinedLoss -> loss.backward())."""
import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen
from chgnet_b200.batch import build_batch
from chgnet_b200.engine import Engine
from chgnet_b200.weights import pack_weights, unpack_grads
from oracle import chgnet_oracle as orc
from oracle.kernel_specs import SpecKernels


def _oracle_param_grads(weights, graphs, ce, cm, args=None):
    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight")
         for k, v in weights.items()}
    out = orc.forward(P, graphs, "em", dtype=torch.float64, train=True, args=args)
    n = out["atoms_per_graph"].double()
    comp = (out["e"] - 0).detach() * 0  # AtomRef shift is constant: drops out of the gradient
    e_tot = out["e"] * n if (args or {}).get("is_intensive", True) else out["e"]
    loss = (e_tot * ce).sum() + (torch.cat(out["m"]) * cm).sum() + comp.sum()
    names = [k for k, v in P.items() if v.requires_grad]
    gr = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gr)}


def _engine_param_grads(weights, graphs, ce, cm, args=None, compact=True):
    sd = {k: torch.as_tensor(np.asarray(v)).double() for k, v in weights.items()}
    pw = pack_weights(sd, args, device="cpu", dtype=torch.float64)
    eng = Engine(pw, SpecKernels())
    b = build_batch(graphs, "cpu", compact_bonds=compact)
    b.frac, b.lattice, b.image = b.frac.double(), b.lattice.double(), b.image.double()
    out = eng.run(b, need_grad=True, need_magmom=True, train=True)
    G = eng.param_grads(out, ce, cm)
    return unpack_grads(G, sd)


@pytest.mark.parametrize("compact", [True, False])
def test_param_grads_match_oracle_autograd(weights030, compact):
    graphs = graphgen.random_graphs(3, 6, 10, 8100)
    n_atoms = sum(g.atomic_number.shape[0] for g in graphs)
    gen = torch.Generator().manual_seed(3)
    ce = torch.randn(len(graphs), generator=gen, dtype=torch.float64)
    cm = torch.randn(n_atoms, generator=gen, dtype=torch.float64)
    want = _oracle_param_grads(weights030, graphs, ce, cm)
    got = _engine_param_grads(weights030, graphs, ce, cm, compact=compact)
    assert set(want) <= set(got)
    for k, w in want.items():
        scale = max(float(w.abs().max()), 1e-12)
        err = float((got[k] - w).abs().max())
        assert err <= 1e-9 * max(scale, 1.0), (k, err, scale)
    # the dead AngleUpdate really has zero gradient in the reference too
    assert float(want["angle_layers.2.twoBody_bond.mlp_core.layers.1.weight"].abs().max()) == 0.0
