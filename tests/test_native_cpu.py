"""CPU: the whole-path C entry points that need no GPU — host weight packing (chg_pack_weights_host)
against weights.py::pack_weights, and the native schedule (chg_forward_plan: call list + workspace
size) against the Python engine's call sequence."""
import numpy as np
import pytest
import torch

from chgnet_b200 import graphgen, native
from chgnet_b200.batch import build_batch
from chgnet_b200.engine import Engine
from chgnet_b200.weights import pack_weights

V020 = dict(num_radial=9, num_angular=9, gMLP_norm=None, readout_norm=None, mlp_out_bias=True, cutoff_coeff=5,
            atom_graph_cutoff=5.0)


def _v020_weights():
    from oracle import chgnet_oracle as orc

    w = orc.random_weights(3, V020)
    w = {k: v for k, v in w.items() if k not in ("mlp.layers.4.weight", "mlp.layers.4.bias")}
    w["mlp.layers.5.weight"], w["mlp.layers.5.bias"] = w.pop("mlp.layers.7.weight"), w.pop("mlp.layers.7.bias")
    return w


@pytest.mark.parametrize("which", ["0.3.0", "0.2.0-shaped"])
def test_host_packing_matches_python_packing(weights030, which):
    w, args = (weights030, None) if which == "0.3.0" else (_v020_weights(), V020)
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in w.items()}
    pw = pack_weights(sd, args, device="cpu")
    hps, blob, hp = native.pack_weights_native(sd, args)
    lay = native.packed_layout(hps)
    assert lay["__total__"][0] == blob.numel()

    def piece(name):
        off, n = lay[name]
        return blob[off:off + n]

    def same(name, t):
        if t is None:
            assert name not in lay
            return
        assert torch.equal(piece(name), t.reshape(-1).float()), name

    for name in ("emb", "freq_ag", "freq_bg", "freq_ang", "w3t", "w3", "wang_t", "wang", "mlp_wt", "mlp_w", "mlp_b", "w_last",
                 "w_mag", "atom_ref"):
        same(name, getattr(pw, name))
    same("readout_ln", pw.readout_ln)
    assert hps.b_last == pytest.approx(pw.b_last) and hps.b_mag == pytest.approx(pw.b_mag)
    for kind, packs in (("atom", pw.atom), ("bond", pw.bond), ("angle", pw.angle)):
        for t, gp in enumerate(packs):
            k = f"{kind}.{t}"
            if kind != "angle":
                same(f"{k}.w2t", gp.w2t), same(f"{k}.w2", gp.w2), same(f"{k}.b2", gp.b2)
                same(f"{k}.wo_t", gp.extra["wo_t"]), same(f"{k}.wo", gp.extra["wo"]), same(f"{k}.bo", gp.extra["bo"])
            same(f"{k}.ln", gp.ln)
            names = ("wcn_t", "we_t", "b1", "wcn_b", "we_b") if kind == "atom" else \
                ("wij_t", "bij", "wx_t", "w1a_t", "wij_b", "wx_b", "w1a_b")
            for name in names:
                key = name
                same(f"{k}.{name}", gp.extra[key])
    assert (hp.n_conv, hp.num_radial) == ((4, 31) if which == "0.3.0" else (4, 9))


@pytest.mark.parametrize("flags", [dict(need_grad=True, need_magmom=True), dict(need_grad=False, need_crystal_fea=True, need_atom_fea=True),
                                   dict(need_grad=True)])
def test_native_schedule_is_the_python_schedule(weights030, flags):
    from kernel_replay import RecordingKernels

    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in weights030.items()}
    graphs = graphgen.random_graphs(3, 8, 12, 9900)
    b = build_batch(graphs, "cpu")
    rec = RecordingKernels()
    Engine(pack_weights(sd, None, device="cpu"), rec).run(b, **flags)
    want = [name for name, _, _ in rec.calls]

    hps, _, _ = native.pack_weights_native(sd, None)
    sizes = native.batch_struct(b)
    wanted = native.Outputs(energy=1, e_ref=1, site_e=1, magmom=1 if flags.get("need_magmom") else None,
                            atom_fea=1 if flags.get("need_atom_fea") else None,
                            crystal_fea=1 if flags.get("need_crystal_fea") else None,
                            force=1 if flags["need_grad"] else None, virial=1 if flags["need_grad"] else None)
    need, got = native.plan(hps, sizes, wanted, want_trace=True)
    assert got == want
    # the workspace covers at least the buffers saved for the reverse pass, and is finite
    saved = (b.n_edges * 128 * 4 + b.n_angles * 128 * 8) * 4 if flags["need_grad"] else 0  # p of 4 AtomConv, pre + p of 3 BondConv, p of 2 AngleUpdate
    assert need >= saved and need < 4 * saved + 64 * (b.n_edges + b.n_angles + b.n_atoms) * 64 * 4
    # no angles / no edges still plan
    g0 = graphgen.make_crystal_graph([3], np.zeros((1, 3)), np.eye(3) * 20.0)
    b0 = build_batch([g0], "cpu")
    need0, got0 = native.plan(hps, native.batch_struct(b0), wanted, want_trace=True)
    rec0 = RecordingKernels()
    Engine(pack_weights(sd, None, device="cpu"), rec0).run(b0, **flags)
    assert got0 == [name for name, _, _ in rec0.calls] and need0 > 0
