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


# ---------------------------------------------------------------------------------------------
# CombinedLoss + data-parallel gradient: 2 ranks over gloo vs the reference loss on the full batch
# ---------------------------------------------------------------------------------------------
def _labels(graphs, seed):
    gen = torch.Generator().manual_seed(seed)
    e_t = torch.randn(len(graphs), generator=gen, dtype=torch.float64)
    e_t[1] = float("nan")  # a missing energy label
    m_t = [torch.rand(g.atomic_number.shape[0], generator=gen, dtype=torch.float64) for g in graphs]
    m_t[2] = None  # a structure without magmom labels (trainer.py:846-851)
    return e_t, m_t


def _reference_loss_grads(weights, graphs, e_t, m_t, criterion):
    """CombinedLoss semantics (trainer.py:797-867, allow_missing_labels=True) on the oracle."""
    crit = {"MSE": torch.nn.MSELoss(), "MAE": torch.nn.L1Loss(), "Huber": torch.nn.HuberLoss(delta=0.1)}[criterion]
    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight")
         for k, v in weights.items()}
    out = orc.forward(P, graphs, "em", dtype=torch.float64, train=True)
    valid = ~torch.isnan(e_t)
    loss = 1.0 * crit(e_t[valid], out["e"][valid])
    mp_, mt_ = [], []
    for mp, mt in zip(out["m"], m_t):
        if mt is not None:
            mp_.append(mp), mt_.append(mt)
    loss = loss + 0.1 * crit(torch.cat(mt_), torch.cat(mp_))
    names = [k for k, v in P.items() if v.requires_grad]
    gr = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    return float(loss.detach()), {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gr)}


def _rank_grads(weights, graphs, e_t, m_t, criterion, group=None):
    from chgnet_b200.trainer import LossConfig, loss_and_grads

    sd = {k: torch.as_tensor(np.asarray(v)).double() for k, v in weights.items()}
    eng = Engine(pack_weights(sd, None, device="cpu", dtype=torch.float64), SpecKernels())
    b = build_batch(graphs, "cpu")
    b.frac, b.lattice, b.image = b.frac.double(), b.lattice.double(), b.image.double()
    m_flat = torch.cat([torch.full((g.atomic_number.shape[0],), float("nan"), dtype=torch.float64) if m is None else m
                        for g, m in zip(graphs, m_t)])
    report, G = loss_and_grads(eng, b, LossConfig("em", criterion), {"e": e_t, "m": m_flat}, True, group)
    return report, unpack_grads(G, sd)


@pytest.mark.parametrize("criterion", ["MSE", "MAE", "Huber"])
def test_combined_loss_gradients_match_reference_loss(weights030, criterion):
    graphs = graphgen.random_graphs(4, 6, 9, 8200)
    e_t, m_t = _labels(graphs, 5)
    want_loss, want = _reference_loss_grads(weights030, graphs, e_t, m_t, criterion)
    report, got = _rank_grads(weights030, graphs, e_t, m_t, criterion)
    # the oracle (like the reference, composition_model.py:191) rounds the AtomRef energy to fp32
    assert report["loss"] == pytest.approx(want_loss, rel=1e-6)
    assert report["e_MAE_size"] == 3 and report["m_MAE_size"] == sum(g.atomic_number.shape[0] for g, m in zip(graphs, m_t) if m is not None)
    for k, w in want.items():
        assert float((got[k] - w).abs().max()) <= 1e-6 * max(float(w.abs().max()), 1.0), k


def _ddp_worker(rank, world, port, q):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from chgnet_b200.batch import graph_cost, partition_graphs

        w = orc.load_weights_npz(os.path.join(os.path.dirname(__file__), "golden", "chgnet_0.3.0_weights.npz"))
        graphs = graphgen.random_graphs(4, 6, 9, 8200)
        e_t, m_t = _labels(graphs, 5)
        mine = partition_graphs([graph_cost(g) for g in graphs], world)[rank]
        report, grads = _rank_grads(w, [graphs[i] for i in mine], e_t[mine], [m_t[i] for i in mine], "MSE")
        flat = torch.cat([grads[k].reshape(-1) for k in sorted(grads) if k != "composition_model.fc.weight"])
        dist.all_reduce(flat)  # the one collective of a training step
        q.put((rank, report["loss"], flat.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_global_batch(weights030):
    import os

    import torch.multiprocessing as mp

    graphs = graphgen.random_graphs(4, 6, 9, 8200)
    e_t, m_t = _labels(graphs, 5)
    want_loss, want = _reference_loss_grads(weights030, graphs, e_t, m_t, "MSE")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=600) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    keys = sorted(want)
    flat_want = torch.cat([want[k].reshape(-1) for k in keys]).numpy()
    for _, loss, flat in got:
        assert loss == pytest.approx(want_loss, rel=1e-6)  # global-batch means on every rank
        assert flat.shape == flat_want.shape
        assert np.abs(flat - flat_want).max() <= 1e-6 * max(np.abs(flat_want).max(), 1.0)


# ---------------------------------------------------------------------------------------------
# force / stress losses: the second-order pass against autograd's double backward (fp64)
# ---------------------------------------------------------------------------------------------
def _oracle_efsm_grads(weights, graphs, ce, cm, cf, cs, args=None):
    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight")
         for k, v in weights.items()}
    out = orc.forward(P, graphs, "efsm", dtype=torch.float64, train=True, args=args)
    n = out["atoms_per_graph"].double()
    loss = ((out["e"] * n * ce).sum() + (torch.cat(out["m"]) * cm).sum() + (torch.cat(out["f"]) * cf).sum()
            + (torch.stack(out["s"]) * cs).sum())
    names = [k for k, v in P.items() if v.requires_grad]
    gr = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gr)}


@pytest.mark.parametrize("compact", [True, False])
def test_force_and_stress_loss_gradients_match_double_backward(weights030, compact):
    graphs = graphgen.random_graphs(3, 6, 10, 8600)
    n_atoms = sum(g.atomic_number.shape[0] for g in graphs)
    gen = torch.Generator().manual_seed(7)
    ce = torch.randn(len(graphs), generator=gen, dtype=torch.float64)
    cm = torch.randn(n_atoms, generator=gen, dtype=torch.float64)
    cf = torch.randn(n_atoms, 3, generator=gen, dtype=torch.float64)
    cs = torch.randn(len(graphs), 3, 3, generator=gen, dtype=torch.float64)
    want = _oracle_efsm_grads(weights030, graphs, ce, cm, cf, cs)

    sd = {k: torch.as_tensor(np.asarray(v)).double() for k, v in weights030.items()}
    eng = Engine(pack_weights(sd, None, device="cpu", dtype=torch.float64), SpecKernels())
    b = build_batch(graphs, "cpu", compact_bonds=compact)
    b.frac, b.lattice, b.image = b.frac.double(), b.lattice.double(), b.image.double()
    out = eng.run(b, need_grad=True, need_magmom=True, train=True)
    eng.input_grads(out, record=True)
    got = unpack_grads(eng.param_grads(out, ce, cm, cf, cs), sd)
    worst = 0.0
    for k, w in want.items():
        scale = max(float(w.abs().max()), 1.0)
        err = float((got[k] - w).abs().max())
        worst = max(worst, err / scale)
        # second derivatives near collinear angles amplify rounding (acos' ~ 1/sqrt(1-u^2) up to 700)
        assert err <= 1e-6 * scale, (k, err, scale)
    print("worst relative error", worst)


def test_training_gradients_without_angles_and_with_isolated_atom(weights030):
    """Empty bond graph (model.py:438, 460) + an atom without edges in the batch, targets "ef" (no magmom):
    both the first-order and the second-order pass against autograd."""
    g_noang = graphgen.make_crystal_graph([3, 8], np.array([[0.0, 0, 0], [0.5, 0.5, 0.5]]), np.eye(3) * 5.5)
    g_iso = graphgen.make_crystal_graph([3], np.zeros((1, 3)), np.eye(3) * 20.0)
    graphs = [g_iso, g_noang]
    assert len(g_noang.bond_graph) == 0 and len(g_iso.atom_graph) == 0
    gen = torch.Generator().manual_seed(9)
    ce = torch.randn(2, generator=gen, dtype=torch.float64)
    cf = torch.randn(3, 3, generator=gen, dtype=torch.float64)
    cs = torch.randn(2, 3, 3, generator=gen, dtype=torch.float64)
    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight")
         for k, v in weights030.items()}
    out = orc.forward(P, graphs, "efs", dtype=torch.float64, train=True)
    n = out["atoms_per_graph"].double()
    names = [k for k, v in P.items() if v.requires_grad]
    loss1 = (out["e"] * n * ce).sum()
    loss2 = loss1 + (torch.cat(out["f"]) * cf).sum() + (torch.stack(out["s"]) * cs).sum()
    want1 = dict(zip(names, torch.autograd.grad(loss1, [P[k] for k in names], allow_unused=True, retain_graph=True)))
    want2 = dict(zip(names, torch.autograd.grad(loss2, [P[k] for k in names], allow_unused=True)))

    sd = {k: torch.as_tensor(np.asarray(v)).double() for k, v in weights030.items()}
    eng = Engine(pack_weights(sd, None, device="cpu", dtype=torch.float64), SpecKernels())

    def batch():
        b = build_batch(graphs, "cpu")
        b.frac, b.lattice, b.image = b.frac.double(), b.lattice.double(), b.image.double()
        return b

    o = eng.run(batch(), need_grad=True, train=True)
    got1 = unpack_grads(eng.param_grads(o, ce), sd)
    o = eng.run(batch(), need_grad=True, train=True)
    eng.input_grads(o, record=True)
    got2 = unpack_grads(eng.param_grads(o, ce, None, cf, cs), sd)
    for want, got in ((want1, got1), (want2, got2)):
        for k in names:
            w = want[k] if want[k] is not None else torch.zeros_like(P[k])
            assert float((got[k] - w).abs().max()) <= 1e-7 * max(float(w.abs().max()), 1.0), k


# ---------------------------------------------------------------------------------------------
# Trainer host logic on the CPU (flat parameter buffer, label packing, fused Adam, engine re-pack) with the
# torch kernel specifications injected in place of the CUDA library
# ---------------------------------------------------------------------------------------------
def test_trainer_step_host_logic_with_spec_kernels(monkeypatch):
    import os

    from chgnet_b200.model import CHGNet
    from chgnet_b200.trainer import Trainer

    path = os.path.join(os.path.dirname(__file__), "golden", "chgnet_0.3.0_weights.npz")
    model = CHGNet.from_file(path, version="0.3.0")  # CPU parameters: the product refuses to run, the test injects the spec
    with pytest.raises(RuntimeError):
        model._get_engine()
    calls = {"n": 0}

    def spec_engine():
        key = tuple(int(v._version) for v in model.state_dict().values()) + (model._engine_key,)
        if model._engine is None or model._engine_key is None:
            model._engine = Engine(pack_weights(model.state_dict(), model.model_args, device="cpu"), SpecKernels())
            model._engine_key = key
            calls["n"] += 1
        return model._engine

    monkeypatch.setattr(model, "_get_engine", spec_engine)
    graphs = graphgen.random_graphs(3, 5, 8, 8900)
    w = orc.load_weights_npz(path)
    base = orc.predict_graph(w, graphs, "efsm", batch_size=3)
    lab = {"e": torch.tensor([float(p["e"]) + 0.05 for p in base]), "f": [torch.as_tensor(p["f"]) + 0.02 for p in base],
           "s": [torch.as_tensor(p["s"]) - 0.05 for p in base], "m": [torch.as_tensor(p["m"]) + 0.03 for p in base]}
    lab["f"][1] = None  # a structure without force labels
    trainer = Trainer(model, targets="efsm", criterion="MSE", learning_rate=1e-3)
    # parameters are views of ONE 64-byte-aligned flat buffer, padding is zero
    lo, hi = trainer.flat.data_ptr(), trainer.flat.data_ptr() + trainer.flat.numel() * 4
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert lo <= p.data_ptr() < hi and (p.data_ptr() - lo) % 64 == 0, n
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    report = trainer.train_step(graphs, lab)

    # reference CombinedLoss on the oracle (fp64) for the same labels
    P = {k: torch.as_tensor(np.asarray(v)).double().requires_grad_(k != "composition_model.fc.weight") for k, v in w.items()}
    o = orc.forward(P, graphs, "efsm", dtype=torch.float64, train=True)
    mse = torch.nn.MSELoss()
    keep = [i for i, f in enumerate(lab["f"]) if f is not None]
    loss = (mse(lab["e"].double(), o["e"]) + mse(torch.cat([lab["f"][i] for i in keep]).double(), torch.cat([o["f"][i] for i in keep]))
            + 0.1 * mse(torch.stack(lab["s"]).double(), torch.stack(o["s"])) + 0.1 * mse(torch.cat(lab["m"]).double(), torch.cat(o["m"])))
    names = [k for k, v in P.items() if v.requires_grad]
    want = dict(zip(names, torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)))
    assert report["loss"] == pytest.approx(float(loss.detach()), rel=1e-3)
    assert report["f_MAE_size"] == 3 * sum(g.atomic_number.shape[0] for i, g in enumerate(graphs) if i in keep)
    got = {k: v.clone() for k, v in trainer.grads_by_name().items()}
    # the one-gather flattening (weights.GradFlattenMap) is the same re-arrangement as unpack_grads + per-parameter copies
    eng = model._get_engine()
    b_chk = build_batch(graphs, "cpu")
    from chgnet_b200.trainer import loss_and_grads

    _, G_chk = loss_and_grads(eng, b_chk, trainer.cfg, trainer._targets(lab, b_chk.atoms_per_graph, "cpu"), model.is_intensive, False)
    slow = {k: v.clone() for k, v in unpack_grads(G_chk, model.state_dict()).items()}
    fast = trainer.flatten_packed_grads(G_chk).clone()
    for n_, o_, sz_ in zip(trainer.names, trainer.offsets, trainer.sizes):
        assert torch.equal(fast[o_:o_ + sz_], slow[n_].reshape(-1).float()), n_
    for k in names:
        wk = want[k] if want[k] is not None else torch.zeros_like(P[k])
        assert float((got[k].double() - wk).abs().max()) <= 2e-2 * float(wk.abs().max()) + 1e-5, k  # fp32 spec vs fp64
    # the fused Adam kernel == torch.optim.Adam on the same gradients; padding untouched; engine re-packed
    ref_params = [before[n].clone().requires_grad_(True) for n in trainer.names]
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    for p, n in zip(ref_params, trainer.names):
        p.grad = got[n].clone()
    opt.step()
    sd = model.state_dict()
    for p, n in zip(ref_params, trainer.names):
        assert float((sd[n] - p.detach()).abs().max()) < 1e-6, n
    mask = torch.ones_like(trainer.flat, dtype=torch.bool)
    for o_, sz in zip(trainer.offsets, trainer.sizes):
        mask[o_:o_ + sz] = False
    assert float(trainer.flat[mask].abs().max()) == 0.0
    assert torch.equal(sd["composition_model.fc.weight"], before["composition_model.fc.weight"])  # frozen
    n_before = calls["n"]
    # checkpoint / resume: a trainer restored from the file continues bit for bit
    import tempfile

    ckpt = os.path.join(tempfile.mkdtemp(), "ckpt.pth.tar")
    trainer.scheduler_step()
    trainer.save(ckpt)
    report2 = trainer.train_step(graphs, lab)
    # the new weights were used, and WITHOUT another Python re-pack: the packed kernel weights are refreshed in place from
    # the flat buffer (weights.RepackMap) and must equal a fresh pack of the current state_dict, tensor for tensor
    assert calls["n"] == n_before and report2["loss"] != report["loss"]
    from chgnet_b200.weights import packed_tensors, _get

    fresh = pack_weights(model.state_dict(), model.model_args, device="cpu")
    live = model._engine.pw
    for (o1, k1), (o2, k2) in zip(packed_tensors(live), packed_tensors(fresh)):
        assert k1 == k2 and torch.equal(_get(o1, k1), _get(o2, k2)), k1
    assert live.b_last == pytest.approx(fresh.b_last, rel=1e-7) and live.b_mag == pytest.approx(fresh.b_mag, rel=1e-7)
    resumed = Trainer.load(ckpt, device="cpu")
    assert resumed.step_count == 1 and resumed.lr == pytest.approx(trainer.lr) and resumed.lr < 1e-3
    m2 = resumed.model
    m2._engine, m2._engine_key = None, None

    def spec_engine2():
        if m2._engine is None or m2._engine_key is None:
            m2._engine = Engine(pack_weights(m2.state_dict(), m2.model_args, device="cpu"), SpecKernels())
            m2._engine_key = ("spec",)
        return m2._engine

    monkeypatch.setattr(m2, "_get_engine", spec_engine2)
    report2b = resumed.train_step(graphs, lab)
    assert report2b["loss"] == pytest.approx(report2["loss"], rel=1e-6)
    for (n1, p1), (n2, p2) in zip(model.named_parameters(), m2.named_parameters()):
        assert n1 == n2 and float((p1.detach() - p2.detach()).abs().max()) < 1e-7, n1
    with pytest.raises(ValueError):
        Trainer(model, targets="fx")


def test_lr_schedules_match_torch():
    from torch.optim.lr_scheduler import CosineAnnealingLR, ExponentialLR, MultiStepLR

    from chgnet_b200.trainer import LRSchedule

    epochs, lr = 3, 1e-3
    for kind, make in (("CosLR", lambda o: CosineAnnealingLR(o, T_max=10 * epochs, eta_min=1e-2 * lr)),
                       ("Exp", lambda o: ExponentialLR(o, gamma=0.98)),
                       ("MultiStepLR", lambda o: MultiStepLR(o, milestones=[4 * epochs, 6 * epochs, 8 * epochs, 9 * epochs], gamma=0.3))):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=lr)
        ref, mine = make(opt), LRSchedule(kind, lr, epochs)
        for _ in range(10 * epochs):
            opt.step()
            ref.step()
            assert mine.step() == pytest.approx(opt.param_groups[0]["lr"], rel=1e-9), kind
    with pytest.raises(NotImplementedError):
        LRSchedule("nope", lr, epochs)
