"""-m gpu: the training step (reference trainer.py:398-411) on the CUDA path: parameter gradients
through the C ABI against fp64 autograd through the oracle, the autograd bridge of
``CHGNet.forward`` in training mode, CombinedLoss + fused Adam in ``Trainer.train_step``."""
import os

import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen

pytestmark = pytest.mark.gpu

# fp32 kernels vs fp64 truth, per parameter tensor: |err| <= GRAD_RTOL * max|grad of that tensor| (+ tiny)
GRAD_RTOL = 2e-3


def _new_model(**kw):
    from chgnet_b200.model import CHGNet

    path = os.path.join(os.path.dirname(__file__), "golden", "chgnet_0.3.0_weights.npz")
    return CHGNet.from_file(path, version="0.3.0", **kw).to("cuda")


def _oracle_grads(weights, graphs, loss_fn, args=None):
    from oracle import chgnet_oracle as orc

    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight")
         for k, v in weights.items()}
    out = orc.forward(P, graphs, "em", dtype=torch.float64, train=True, args=args)
    loss = loss_fn(out["e"], torch.cat(out["m"]))
    names = [k for k, v in P.items() if v.requires_grad]
    gr = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    return float(loss.detach()), {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gr)}


def _check(got: dict, want: dict, rtol=GRAD_RTOL):
    worst = {}
    for k, w in want.items():
        scale = float(w.abs().max())
        err = float((got[k].double().cpu() - w).abs().max())
        worst[k] = err / max(scale, 1e-30) if scale > 0 else err
        assert err <= rtol * scale + 1e-7, (k, err, scale)
    return max(worst.values())


def test_forward_in_training_mode_backpropagates_to_parameters(weights030):
    model = _new_model()
    model.train()
    graphs = graphgen.random_graphs(4, 10, 18, 8300)
    gen = torch.Generator().manual_seed(1)
    n_atoms = sum(g.atomic_number.shape[0] for g in graphs)
    ce = torch.randn(len(graphs), generator=gen, dtype=torch.float64)
    cm = torch.randn(n_atoms, generator=gen, dtype=torch.float64)

    def loss_fn(e, m):
        return (e * ce.to(e.device, e.dtype)).sum() + ((m - 0.3) ** 2 * cm.to(m.device, m.dtype)).sum()

    pred = model(graphs, task="em")
    assert pred["e"].requires_grad and pred["m"][0].requires_grad
    loss = loss_fn(pred["e"], torch.cat(pred["m"]))
    loss.backward()
    want_loss, want = _oracle_grads(weights030, graphs, loss_fn)
    assert float(loss) == pytest.approx(want_loss, rel=1e-4, abs=1e-4)
    got = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
    assert model.get_parameter("composition_model.fc.weight").grad is None  # frozen (composition_model.py:127-131)
    print("worst relative gradient error:", _check(got, want))
    # eval mode / no_grad: plain tensors, no history
    model.eval()
    assert not model(graphs, task="e")["e"].requires_grad


def test_trainer_step_matches_reference_loss_and_torch_adam(weights030):
    from chgnet_b200.trainer import Trainer

    model = _new_model()
    graphs = graphgen.random_graphs(6, 10, 18, 8400)
    gen = torch.Generator().manual_seed(2)
    base = model.predict_graph(graphs, task="em", batch_size=6)
    e_t = torch.tensor([float(p["e"]) for p in base]) + 0.05 * torch.randn(6, generator=gen)
    e_t[4] = float("nan")
    m_t = [torch.as_tensor(p["m"]) + 0.1 * torch.randn(len(p["m"]), generator=gen) for p in base]
    m_t[1] = None

    crit = torch.nn.HuberLoss(delta=0.1)

    def loss_fn(e, m):
        valid = ~torch.isnan(e_t)
        keep = torch.cat([torch.full((len(b["m"]),), t is not None) for b, t in zip(base, m_t)])
        mt = torch.cat([t if t is not None else torch.zeros(len(b["m"])) for b, t in zip(base, m_t)]).double()
        return crit(e_t.double()[valid], e[valid]) + 0.1 * crit(mt[keep], m[keep])

    want_loss, want = _oracle_grads(weights030, graphs, loss_fn)
    trainer = Trainer(model, targets="em", criterion="Huber", learning_rate=1e-3)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    report = trainer.train_step(graphs, {"e": e_t, "m": m_t})
    assert report["loss"] == pytest.approx(want_loss, rel=2e-3, abs=1e-6)
    assert report["e_MAE_size"] == 5
    names = trainer.names
    got = {n: g.clone() for n, g in trainer.grads_by_name().items()}
    _check(got, want, rtol=5e-3)
    # the fused Adam step == torch.optim.Adam on the same gradients
    ref_params = [before[n].clone().requires_grad_(True) for n in names]
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    for p, n in zip(ref_params, names):
        p.grad = got[n].view(p.shape).clone()
    opt.step()
    sd = model.state_dict()
    for p, n in zip(ref_params, names):
        assert float((sd[n] - p.detach()).abs().max()) < 1e-6, n
    # predict_* sees the new weights
    after = model.predict_graph(graphs, task="e", batch_size=6)
    assert abs(float(after[0]["e"]) - float(base[0]["e"])) > 1e-6
    # small steps on the same batch: the loss goes down
    slow = Trainer(_new_model(), targets="em", criterion="Huber", learning_rate=1e-5)
    losses = [slow.train_step(graphs, {"e": e_t, "m": m_t})["loss"] for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
    with pytest.raises(ValueError):
        Trainer(model, targets="fx")


def test_v020_shaped_architecture_parameter_gradients():
    """no LayerNorm, mlp_out bias (identity bond compaction), 9 radial / 9 angular functions."""
    from chgnet_b200.model import CHGNet

    torch.manual_seed(5)
    model = CHGNet(num_radial=9, num_angular=9, gMLP_norm=None, readout_norm=None, mlp_hidden_dims=[64, 64],
                   cutoff_coeff=5, atom_graph_cutoff=5, mlp_out_bias=True, composition_model="MPtrj").to("cuda")
    model.train()
    w = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    args = dict(num_radial=9, num_angular=9, gMLP_norm=None, readout_norm=None, mlp_out_bias=True, cutoff_coeff=5,
                atom_graph_cutoff=5.0)
    graphs = graphgen.random_graphs(3, 8, 14, 8500, atom_graph_cutoff=5.0)

    def loss_fn(e, m):
        return (e**2).sum() + 0.1 * (m**2).sum()

    pred = model(graphs, task="em")
    loss_fn(pred["e"], torch.cat(pred["m"])).backward()
    _, want = _oracle_grads(w, graphs, loss_fn, args=args)
    got = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
    print("worst relative gradient error:", _check(got, want, rtol=5e-3))


def _oracle_grads_efsm(weights, graphs, loss_fn, args=None):
    from oracle import chgnet_oracle as orc

    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight")
         for k, v in weights.items()}
    out = orc.forward(P, graphs, "efsm", dtype=torch.float64, train=True, args=args)
    loss = loss_fn(out["e"], torch.cat(out["m"]), torch.cat(out["f"]), torch.stack(out["s"]))
    names = [k for k, v in P.items() if v.requires_grad]
    gr = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    return float(loss.detach()), {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gr)}


def test_force_and_stress_losses_backpropagate_through_the_second_order_pass(weights030):
    """loss on e, f, s, m through CHGNet.forward in training mode == autograd's double backward (fp64)."""
    model = _new_model()
    model.train()
    graphs = graphgen.random_graphs(4, 10, 18, 8700)
    gen = torch.Generator().manual_seed(4)
    n_atoms = sum(g.atomic_number.shape[0] for g in graphs)
    cf = torch.randn(n_atoms, 3, generator=gen, dtype=torch.float64)
    cs = torch.randn(len(graphs), 3, 3, generator=gen, dtype=torch.float64)

    def loss_fn(e, m, f, s):
        d = dict(device=e.device, dtype=e.dtype)
        return (e**2).sum() + 0.1 * (m**2).sum() + ((f - 0.05 * cf.to(**d)) ** 2).mean() + 0.1 * ((s - 0.1 * cs.to(**d)) ** 2).mean()

    pred = model(graphs, task="efsm")
    assert pred["f"][0].requires_grad and pred["s"][0].requires_grad
    loss = loss_fn(pred["e"], torch.cat(pred["m"]), torch.cat(pred["f"]), torch.stack(pred["s"]))
    loss.backward()
    want_loss, want = _oracle_grads_efsm(weights030, graphs, loss_fn)
    assert float(loss) == pytest.approx(want_loss, rel=1e-3)
    got = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
    print("worst relative gradient error:", _check(got, want, rtol=1e-2))


def test_trainer_efsm_step_matches_reference_combined_loss(weights030):
    from chgnet_b200.trainer import Trainer

    model = _new_model()
    graphs = graphgen.random_graphs(5, 10, 18, 8800)
    gen = torch.Generator().manual_seed(6)
    base = model.predict_graph(graphs, task="efsm", batch_size=5)
    noisy = lambda v, a: torch.as_tensor(np.asarray(v), dtype=torch.float32) + a * torch.randn(np.asarray(v).shape, generator=gen)  # noqa: E731
    lab = {"e": noisy([float(p["e"]) for p in base], 0.05), "f": [noisy(p["f"], 0.02) for p in base],
           "s": [noisy(p["s"], 0.05) for p in base], "m": [noisy(p["m"], 0.05) for p in base]}
    lab["m"][3] = None
    crit = torch.nn.MSELoss()

    def loss_fn(e, m, f, s):
        keep = torch.cat([torch.full((len(b["m"]),), t is not None) for b, t in zip(base, lab["m"])])
        mt = torch.cat([t if t is not None else torch.zeros(len(b["m"])) for b, t in zip(base, lab["m"])]).double()
        return (crit(lab["e"].double(), e) + crit(torch.cat(lab["f"]).double(), f)
                + 0.1 * crit(torch.stack(lab["s"]).double(), s) + 0.1 * crit(mt[keep], m[keep]))

    want_loss, want = _oracle_grads_efsm(weights030, graphs, loss_fn)
    trainer = Trainer(model, targets="efsm", criterion="MSE", learning_rate=1e-5)
    report = trainer.train_step(graphs, lab)
    assert report["loss"] == pytest.approx(want_loss, rel=5e-3, abs=1e-7)
    assert report["f_MAE_size"] == 3 * sum(len(b["m"]) for b in base) and report["s_MAE_size"] == 45
    _check({n: g.clone() for n, g in trainer.grads_by_name().items()}, want, rtol=1e-2)
    losses = [report["loss"]] + [trainer.train_step(graphs, lab)["loss"] for _ in range(6)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
