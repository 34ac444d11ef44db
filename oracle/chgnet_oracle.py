"""TEST INFRASTRUCTURE — CPU oracle for the CHGNet hot path (not product code).

A from-equations restatement of the reference's
forward + autograd backward, written as pure functions over a flat weight
dictionary (the reference ``state_dict`` names).  Runs in fp32 (the reference's
arithmetic) or fp64 (truth for error budgets).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs import this module; the product path (chgnet_b200/) never does.

Parity status: PINNED.  ``oracle/make_golden.py`` checks this file against the
live reference (imported from /root/reference in the build container) on the
LiMnO2 golden structure of reference tests/test_model.py:60-119 and on random
batches, and the fixtures it writes under tests/golden/ are re-checked by
tests/test_oracle_golden.py on every run.

Reference lines restated by each function are cited in its docstring.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
from torch import Tensor

EV_A3_TO_GPA = 160.21766208  # reference chgnet/model/model.py:533


# --------------------------------------------------------------------------
# weights / hyper-parameters
# --------------------------------------------------------------------------
DEFAULT_ARGS = dict(  # pretrained 0.3.0
    num_radial=31,
    num_angular=31,
    n_conv=4,
    atom_graph_cutoff=6.0,
    bond_graph_cutoff=3.0,
    cutoff_coeff=8,
    gMLP_norm="layer",
    readout_norm="layer",
    is_intensive=True,
    mlp_out_bias=False,
)


def load_weights_npz(path: str) -> dict[str, np.ndarray]:
    with np.load(path) as f:
        return {k: f[k] for k in f.files}


def random_weights(seed: int = 0, args: dict | None = None) -> dict[str, np.ndarray]:
    """Random-init weights with the 0.3.0 shapes (scaled so activations stay O(1))."""
    a = {**DEFAULT_ARGS, **(args or {})}
    rng = np.random.default_rng(seed)
    R, NA = a["num_radial"], a["num_angular"]
    w: dict[str, np.ndarray] = {}

    def lin(name, out_d, in_d, bias=True, scale=1.0):
        w[f"{name}.weight"] = (rng.standard_normal((out_d, in_d)) * scale / math.sqrt(in_d)).astype(np.float32)
        if bias:
            w[f"{name}.bias"] = (rng.standard_normal(out_d) * 0.1).astype(np.float32)

    def ln(name):
        w[f"{name}.weight"] = (1.0 + 0.1 * rng.standard_normal(64)).astype(np.float32)
        w[f"{name}.bias"] = (0.1 * rng.standard_normal(64)).astype(np.float32)

    w["composition_model.fc.weight"] = rng.uniform(-10, 0, (1, 94)).astype(np.float32)
    w["atom_embedding.embedding.weight"] = rng.standard_normal((94, 64)).astype(np.float32)
    w["bond_basis_expansion.rbf_expansion_ag.frequencies"] = (
        np.pi * np.arange(1, R + 1) * rng.uniform(0.9, 1.1, R)
    ).astype(np.float32)
    w["bond_basis_expansion.rbf_expansion_bg.frequencies"] = (
        np.pi * np.arange(1, R + 1) * rng.uniform(0.9, 1.1, R)
    ).astype(np.float32)
    w["angle_basis_expansion.fourier_expansion.frequencies"] = (
        np.arange(1, (NA - 1) // 2 + 1) * rng.uniform(0.9, 1.1, (NA - 1) // 2)
    ).astype(np.float32)
    lin("bond_embedding", 64, R, bias=False)
    lin("bond_weights_ag", 64, R, bias=False)
    lin("bond_weights_bg", 64, R, bias=False)
    lin("angle_embedding", 64, NA, bias=False)
    for t in range(a["n_conv"]):
        p = f"atom_conv_layers.{t}.twoBody_atom"
        for br in ("mlp_core", "mlp_gate"):
            lin(f"{p}.{br}.layers.0", 64, 192)
            lin(f"{p}.{br}.layers.3", 64, 64)
        if a["gMLP_norm"] == "layer":
            ln(f"{p}.bn1"), ln(f"{p}.bn2")
        lin(f"atom_conv_layers.{t}.mlp_out.layers.1", 64, 64, bias=a["mlp_out_bias"], scale=0.3)
    for t in range(a["n_conv"] - 1):
        p = f"bond_conv_layers.{t}.twoBody_bond"
        for br in ("mlp_core", "mlp_gate"):
            lin(f"{p}.{br}.layers.0", 64, 256)
            lin(f"{p}.{br}.layers.3", 64, 64)
        if a["gMLP_norm"] == "layer":
            ln(f"{p}.bn1"), ln(f"{p}.bn2")
        lin(f"bond_conv_layers.{t}.mlp_out.layers.1", 64, 64, bias=a["mlp_out_bias"], scale=0.3)
        p = f"angle_layers.{t}.twoBody_bond"
        for br in ("mlp_core", "mlp_gate"):
            lin(f"{p}.{br}.layers.1", 64, 256)
        if a["gMLP_norm"] == "layer":
            ln(f"{p}.bn1"), ln(f"{p}.bn2")
    lin("site_wise", 1, 64)
    if a["readout_norm"] == "layer":
        ln("readout_norm")
    lin("mlp.layers.0", 64, 64)
    lin("mlp.layers.2", 64, 64)
    lin("mlp.layers.4", 64, 64)
    lin("mlp.layers.7", 1, 64)
    return w


def _t(w: dict, dtype, device=None) -> dict[str, Tensor]:
    return {k: torch.as_tensor(np.asarray(v)).to(device=device, dtype=dtype) for k, v in w.items()}


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def radial_bessel(d: Tensor, freq: Tensor, cutoff: float, p: int) -> Tensor:
    """sqrt(2/rc) sin(w d/rc)/d * env(d/rc)  (reference basis.py:108-116, 184-205)."""
    d = d[:, None]
    x = d / cutoff
    out = math.sqrt(2.0 / cutoff) * torch.sin(freq * x) / d
    if p != 0:
        a = -(p + 1) * (p + 2) / 2
        b = p * (p + 2)
        c = -p * (p + 1) / 2
        env = 1 + a * x**p + b * x ** (p + 1) + c * x ** (p + 2)
        env = torch.where(x < 1, env, torch.zeros_like(x))
        out = env * out
    return out


def fourier(theta: Tensor, freq: Tensor) -> Tensor:
    """[1/sqrt2, sin(w th), cos(w th)]/sqrt(pi)  (reference basis.py:33-40)."""
    arg = theta[:, None] * freq[None, :]
    const = torch.full((theta.shape[0], 1), 1.0 / math.sqrt(2.0), dtype=theta.dtype, device=theta.device)
    return torch.cat([const, torch.sin(arg), torch.cos(arg)], dim=1) / math.sqrt(math.pi)


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def silu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def linear(x: Tensor, w: dict, name: str) -> Tensor:
    y = x @ w[f"{name}.weight"].T
    if f"{name}.bias" in w:
        y = y + w[f"{name}.bias"]
    return y


def gated_mlp(z: Tensor, w: dict, prefix: str, hidden: bool) -> Tensor:
    """silu(LN1(core(z))) * sigmoid(LN2(gate(z)))  (reference functions.py:168-183)."""
    def branch(br):
        if hidden:
            h = silu(linear(z, w, f"{prefix}.{br}.layers.0"))
            return linear(h, w, f"{prefix}.{br}.layers.3")
        return linear(z, w, f"{prefix}.{br}.layers.1")

    core, gate = branch("mlp_core"), branch("mlp_gate")
    if f"{prefix}.bn1.weight" in w:
        core = layer_norm(core, w[f"{prefix}.bn1.weight"], w[f"{prefix}.bn1.bias"])
        gate = layer_norm(gate, w[f"{prefix}.bn2.weight"], w[f"{prefix}.bn2.bias"])
    return silu(core) * torch.sigmoid(gate)


def scatter_sum(data: Tensor, owners: Tensor, n: int) -> Tensor:
    """reference functions.py:10-40 with average=False."""
    out = data.new_zeros((n, data.shape[1]))
    return out.index_add_(0, owners, data)


# --------------------------------------------------------------------------
# the path
# --------------------------------------------------------------------------
def forward(
    weights: dict,
    graphs: Sequence,
    task: str = "efsm",
    *,
    dtype=torch.float32,
    args: dict | None = None,
    return_site_energies: bool = False,
    return_atom_feas: bool = False,
    return_crystal_feas: bool = False,
    return_intermediates: bool = False,
    train: bool = False,
    device=None,
) -> dict:
    """CHGNet.forward restated (reference model.py:330-542, 792-913).

    ``train=True`` is the reference's training mode: ``weights`` is a dict of torch tensors
    (leaf parameters), forces/stresses are taken with ``create_graph=True`` (model.py:518, 529)
    and nothing is detached, so a loss on e/f/s/m back-propagates to the parameters.

    Returns the reference's dict: e [B] (eV/atom if intensive), f list[n_i,3],
    s list[3,3] GPa, m list[n_i], atoms_per_graph, plus optional extras.
    """
    a = {**DEFAULT_ARGS, **(args or {})}
    w = dict(weights) if train else _t(weights, dtype, device)
    if device is not None:  # stock PyTorch on an accelerator: same ops, graph tensors moved like CrystalGraph.to()
        graphs = [g.to(device) for g in graphs]
    R = a["num_radial"]
    p = int(a["cutoff_coeff"])
    n_conv = a["n_conv"]
    want_f, want_s, want_m = "f" in task, "s" in task, "m" in task

    z_all, pos_list, strain_list, vols = [], [], [], []
    bases_ag, bases_bg, bases_ang = [], [], []
    ag_list, d2u_list, bg_list, owners = [], [], [], []
    atom_off = und_off = 0
    for gi, g in enumerate(graphs):
        n = g.atomic_number.shape[0]
        lat0 = g.lattice.detach().to(dtype)
        if want_s:  # model.py:826-830
            strain = torch.zeros(3, 3, dtype=dtype, device=lat0.device, requires_grad=True)
            lat = lat0 @ (torch.eye(3, dtype=dtype, device=lat0.device) + strain)
        else:
            strain, lat = None, lat0
        vols.append(torch.dot(lat[0], torch.linalg.cross(lat[1], lat[2])))  # 834-836
        strain_list.append(strain)
        frac = g.atom_frac_coord.detach().to(dtype)
        cart = frac @ lat  # 840
        if want_f and not cart.requires_grad:
            cart.requires_grad_(True)
        ag = g.atom_graph.long().reshape(-1, 2)
        img = g.neighbor_image.detach().to(dtype).reshape(-1, 3)
        # encoders.py:98-110
        nb = cart[ag[:, 1]] + img @ lat
        vec = cart[ag[:, 0]] - nb
        dist = torch.norm(vec, dim=1)
        unit = vec / dist[:, None]
        du = dist[g.undirected2directed.long()]
        bases_ag.append(radial_bessel(du, w["bond_basis_expansion.rbf_expansion_ag.frequencies"], a["atom_graph_cutoff"], p))
        bases_bg.append(radial_bessel(du, w["bond_basis_expansion.rbf_expansion_bg.frequencies"], a["bond_graph_cutoff"], p))
        pos_list.append(cart)
        ag_list.append(ag + atom_off)
        d2u_list.append(g.directed2undirected.long() + und_off)
        bg = g.bond_graph.long().reshape(-1, 5)
        if len(bg):  # model.py:863-877, encoders.py:144-146
            cosij = (unit[bg[:, 2]] * unit[bg[:, 4]]).sum(dim=1) * (1 - 1e-6)
            bases_ang.append(fourier(torch.acos(cosij), w["angle_basis_expansion.fourier_expansion.frequencies"]))
            bg_list.append(torch.stack([bg[:, 0] + atom_off, bg[:, 1] + und_off, bg[:, 3] + und_off], dim=1))
        z_all.append(g.atomic_number.long())
        owners.append(torch.full((n,), gi, dtype=torch.long, device=lat0.device))
        atom_off += n
        und_off += len(du)

    z_all = torch.cat(z_all)
    owners = torch.cat(owners)
    n_atoms = atom_off
    B = len(graphs)
    atoms_per_graph = torch.bincount(owners, minlength=B)
    ag = torch.cat(ag_list)
    d2u = torch.cat(d2u_list)
    b_ag = torch.cat(bases_ag)
    b_bg = torch.cat(bases_bg)
    has_angles = len(bases_ang) != 0
    inter: dict = {}

    # embeddings (model.py:432-439)
    x = w["atom_embedding.embedding.weight"][z_all - 1]
    e = b_ag @ w["bond_embedding.weight"].T
    w_ag = b_ag @ w["bond_weights_ag.weight"].T
    w_bg = b_bg @ w["bond_weights_bg.weight"].T
    if has_angles:
        ang = torch.cat(bases_ang) @ w["angle_embedding.weight"].T
        bgr = torch.cat(bg_list)
    if return_intermediates:
        inter.update(x0=x, e0=e, w_ag=w_ag, w_bg=w_bg, a0=ang if has_angles else None)

    def atom_conv(t, x, e):  # layers.py:113-132
        zc = torch.cat([x[ag[:, 0]], e[d2u], x[ag[:, 1]]], dim=1)
        msg = gated_mlp(zc, w, f"atom_conv_layers.{t}.twoBody_atom", True) * w_ag[d2u]
        agg = scatter_sum(msg, ag[:, 0], n_atoms)
        return linear(agg, w, f"atom_conv_layers.{t}.mlp_out.layers.1") + x

    out: dict = {"atoms_per_graph": atoms_per_graph}
    for t in range(n_conv - 1):
        x = atom_conv(t, x, e)
        if has_angles:
            # BondConv, layers.py:238-260
            zc = torch.cat([e[bgr[:, 1]], e[bgr[:, 2]], ang, x[bgr[:, 0]]], dim=1)
            upd = gated_mlp(zc, w, f"bond_conv_layers.{t}.twoBody_bond", True)
            upd = upd * w_bg[bgr[:, 1]] * w_bg[bgr[:, 2]]
            agg = scatter_sum(upd, bgr[:, 1], e.shape[0])
            e = linear(agg, w, f"bond_conv_layers.{t}.mlp_out.layers.1") + e
            # AngleUpdate, layers.py:348-360 (the last one is dead compute; kept for fidelity)
            zc = torch.cat([e[bgr[:, 1]], e[bgr[:, 2]], ang, x[bgr[:, 0]]], dim=1)
            ang = gated_mlp(zc, w, f"angle_layers.{t}.twoBody_bond", False) + ang
        if return_intermediates:
            inter[f"x{t + 1}"] = x
            inter[f"e{t + 1}"] = e
            inter[f"a{t + 1}"] = ang if has_angles else None
        if t == n_conv - 2:  # model.py:477-487
            if return_atom_feas:
                out["atom_fea"] = list(torch.split(x, atoms_per_graph.tolist()))
            if want_m:
                mag = torch.abs(linear(x, w, "site_wise")).view(-1)
                out["m"] = list(torch.split(mag, atoms_per_graph.tolist()))
    x = atom_conv(n_conv - 1, x, e)
    if "readout_norm.weight" in w:
        x = layer_norm(x, w["readout_norm.weight"], w["readout_norm.bias"])
    # readout MLP (model.py:497-509; functions.py:81-92)
    h = x
    idx = 0
    while f"mlp.layers.{idx}.weight" in w and w[f"mlp.layers.{idx}.weight"].shape[0] != 1:
        h = silu(linear(h, w, f"mlp.layers.{idx}"))
        idx += 2
    last = max(int(k.split(".")[2]) for k in w if k.startswith("mlp.layers.") and k.endswith(".weight"))
    site_e = linear(h, w, f"mlp.layers.{last}").view(-1)
    energy = torch.zeros(B, dtype=dtype, device=site_e.device).index_add_(0, owners, site_e)
    if return_crystal_feas:
        out["crystal_fea"] = scatter_sum(x, owners, B)
    if return_intermediates:
        inter["x_readout"] = x
        inter["site_e_model"] = site_e

    if want_f:  # model.py:517-524
        grads = torch.autograd.grad(energy.sum(), pos_list, retain_graph=want_s or train, create_graph=train)
        out["f"] = [-gr for gr in grads]
    if want_s:  # model.py:527-535
        grads = torch.autograd.grad(energy.sum(), strain_list, retain_graph=train, create_graph=train)
        out["s"] = [gr * (EV_A3_TO_GPA / v.detach()) for gr, v in zip(grads, vols)]

    # AtomRef (composition_model.py:175-205) + intensive normalisation (model.py:538-540)
    wref = w["composition_model.fc.weight"][0].float()
    comp = torch.stack(
        [
            (torch.bincount(g.atomic_number.long() - 1, minlength=94) / g.atomic_number.shape[0]).float()
            if a["is_intensive"]
            else torch.bincount(g.atomic_number.long() - 1, minlength=94).float()
            for g in graphs
        ]
    )
    comp_e = (comp @ wref).to(dtype)
    e_out = energy if train else energy.detach()
    if a["is_intensive"]:
        e_out = e_out / atoms_per_graph
    out["e"] = e_out + comp_e
    if return_site_energies:
        shift = wref[z_all - 1].to(dtype)
        out["site_energies"] = list(torch.split(site_e.detach() + shift, atoms_per_graph.tolist()))
    if not train:
        for k in ("m", "atom_fea"):
            if k in out:
                out[k] = [v.detach() for v in out[k]]
        if "crystal_fea" in out:
            out["crystal_fea"] = out["crystal_fea"].detach()
    if return_intermediates:
        out["intermediates"] = {k: (v.detach() if v is not None else None) for k, v in inter.items()}
    return out


def predict_graph(weights, graphs, task="efsm", *, batch_size=16, dtype=torch.float32, args=None, **kw):
    """``CHGNet.predict_graph`` restated (reference model.py:593-665): numpy outputs."""
    single = not isinstance(graphs, (list, tuple))
    gl = [graphs] if single else list(graphs)
    preds: list[dict] = []
    for s in range(0, len(gl), batch_size):
        chunk = gl[s : s + batch_size]
        o = forward(weights, chunk, task, dtype=dtype, args=args, **kw)
        for i in range(len(chunk)):
            d = {}
            for key in ("e", "f", "s", "m", "site_energies", "atom_fea", "crystal_fea"):
                if key in o:
                    d[key] = o[key][i].detach().cpu().numpy()
            preds.append(d)
    return preds[0] if single else preds
