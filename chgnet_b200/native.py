"""ctypes face of the whole-path C entry points (include/chgnet_b200.h: ``chg_pack_weights_host``,
``chg_forward_plan``, ``chg_forward``).

``NativeForward`` is what ``CHGNet`` uses for inference: the reference ``state_dict`` is packed by the
C library on the host, uploaded once, and every prediction is ONE C call that runs the whole kernel
schedule (forward + force / stress reverse pass) on a cached workspace — the boundary a non-Python host
would bind.  The Python ``Engine`` (engine.py) runs the same kernels call by call; it stays the
schedule used for training and for the CPU specification tests.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

import torch
from torch import Tensor

from chgnet_b200._lib import ChgnetB200Error, load_library
from chgnet_b200.batch import DeviceBatch
from chgnet_b200.weights import HyperParams, infer_hyper_params

MAX_CONV = 8
FP = POINTER(c_float)


class HParams(Structure):
    _fields_ = [(n, c_int32) for n in ("num_radial", "num_angular", "n_conv", "cutoff_coeff", "n_readout_hidden", "use_ln",
                                       "readout_ln", "has_mlp_out_bias")] + \
               [(n, c_float) for n in ("atom_graph_cutoff", "bond_graph_cutoff", "b_last", "b_mag")]


class GatedSD(Structure):
    _fields_ = [(n, FP) for n in ("core_w1", "core_b1", "gate_w1", "gate_b1", "core_w2", "core_b2", "gate_w2", "gate_b2",
                                  "ln1_w", "ln1_b", "ln2_w", "ln2_b", "out_w", "out_b")]


class StateDict(Structure):
    _fields_ = [(n, FP) for n in ("atom_embedding", "freq_ag", "freq_bg", "freq_ang", "bond_embedding", "bond_weights_ag",
                                  "bond_weights_bg", "angle_embedding")] + \
               [("atom", GatedSD * MAX_CONV), ("bond", GatedSD * MAX_CONV), ("angle", GatedSD * MAX_CONV),
                ("readout_ln_w", FP), ("readout_ln_b", FP), ("mlp_w", FP * 4), ("mlp_b", FP * 4),
                ("mlp_last_w", FP), ("mlp_last_b", c_float), ("site_wise_w", FP), ("site_wise_b", c_float), ("atom_ref", FP)]


_BATCH_PTRS = ("z", "frac", "owner", "lattice", "center", "nbr", "image", "d2u", "u2d", "ptr_c", "perm_n", "ptr_n", "perm_u",
               "ptr_u", "ang_atom", "ang_di", "ang_dj", "ang_is", "ang_js", "ptr_is", "perm_js", "ptr_js", "perm_x", "ptr_x",
               "short_ids", "graph_ptr")


class Batch(Structure):
    _fields_ = [(n, c_int32) for n in ("n_atoms", "n_edges", "n_bonds", "n_angles", "n_graphs", "n_short")] + \
               [(n, c_void_p) for n in _BATCH_PTRS]


class Outputs(Structure):
    _fields_ = [(n, c_void_p) for n in ("energy", "e_ref", "site_e", "magmom", "atom_fea", "crystal_fea", "force", "virial")]


def _bind(lib):
    if getattr(lib, "_native_bound", False):
        return lib
    lib.chg_packed_floats.restype = c_int64
    lib.chg_packed_floats.argtypes = [POINTER(HParams)]
    lib.chg_pack_weights_host.restype = c_int32
    lib.chg_pack_weights_host.argtypes = [POINTER(HParams), POINTER(StateDict), FP]
    lib.chg_forward_plan.restype = c_int32
    lib.chg_forward_plan.argtypes = [POINTER(HParams), POINTER(Batch), POINTER(Outputs), POINTER(c_size_t), c_char_p, c_size_t]
    lib.chg_forward.restype = c_int32
    lib.chg_forward.argtypes = [POINTER(HParams), c_void_p, POINTER(Batch), POINTER(Outputs), c_void_p, c_size_t, c_void_p]
    lib._native_bound = True
    return lib


def _check(lib, rc: int, what: str) -> None:
    if rc != 0:
        raise ChgnetB200Error(f"{what} failed ({rc}): {lib.chg_last_error().decode()}")


def hparams_struct(hp: HyperParams, has_bias: bool) -> HParams:
    return HParams(num_radial=hp.num_radial, num_angular=hp.num_angular, n_conv=hp.n_conv, cutoff_coeff=hp.cutoff_coeff,
                   n_readout_hidden=hp.n_readout_hidden, use_ln=int(hp.use_ln), readout_ln=int(hp.readout_ln),
                   has_mlp_out_bias=int(has_bias), atom_graph_cutoff=hp.atom_graph_cutoff, bond_graph_cutoff=hp.bond_graph_cutoff)


def pack_weights_native(state_dict: dict, model_args: dict | None = None) -> tuple[HParams, Tensor, HyperParams]:
    """state_dict (reference names) -> (hyper-parameter struct, packed fp32 blob on the HOST, HyperParams).
    The packing itself runs in the C library (``chg_pack_weights_host``)."""
    lib = _bind(load_library())
    sd = {k: v.detach().to(device="cpu", dtype=torch.float32).contiguous() for k, v in state_dict.items()
          if torch.is_tensor(v) and torch.is_floating_point(v)}
    hp = infer_hyper_params(sd, model_args)
    has_bias = "atom_conv_layers.0.mlp_out.layers.1.bias" in sd
    hps = hparams_struct(hp, has_bias)
    keep = []  # keeps the host arrays alive during the call

    def ptr(name: str | None):
        if name is None or name not in sd:
            return None
        keep.append(sd[name])
        return ctypes.cast(sd[name].data_ptr(), FP)

    def gated(prefix: str, first: str, second: str | None, out: str | None) -> GatedSD:
        g = GatedSD(core_w1=ptr(f"{prefix}.mlp_core.{first}.weight"), core_b1=ptr(f"{prefix}.mlp_core.{first}.bias"),
                    gate_w1=ptr(f"{prefix}.mlp_gate.{first}.weight"), gate_b1=ptr(f"{prefix}.mlp_gate.{first}.bias"),
                    ln1_w=ptr(f"{prefix}.bn1.weight"), ln1_b=ptr(f"{prefix}.bn1.bias"),
                    ln2_w=ptr(f"{prefix}.bn2.weight"), ln2_b=ptr(f"{prefix}.bn2.bias"))
        if second is not None:
            g.core_w2, g.core_b2 = ptr(f"{prefix}.mlp_core.{second}.weight"), ptr(f"{prefix}.mlp_core.{second}.bias")
            g.gate_w2, g.gate_b2 = ptr(f"{prefix}.mlp_gate.{second}.weight"), ptr(f"{prefix}.mlp_gate.{second}.bias")
        if out is not None:
            g.out_w, g.out_b = ptr(f"{out}.weight"), ptr(f"{out}.bias")
        return g

    s = StateDict(atom_embedding=ptr("atom_embedding.embedding.weight"),
                  freq_ag=ptr("bond_basis_expansion.rbf_expansion_ag.frequencies"),
                  freq_bg=ptr("bond_basis_expansion.rbf_expansion_bg.frequencies"),
                  freq_ang=ptr("angle_basis_expansion.fourier_expansion.frequencies"),
                  bond_embedding=ptr("bond_embedding.weight"), bond_weights_ag=ptr("bond_weights_ag.weight"),
                  bond_weights_bg=ptr("bond_weights_bg.weight"), angle_embedding=ptr("angle_embedding.weight"),
                  readout_ln_w=ptr("readout_norm.weight"), readout_ln_b=ptr("readout_norm.bias"),
                  site_wise_w=ptr("site_wise.weight"), site_wise_b=float(sd["site_wise.bias"].reshape(-1)[0]),
                  atom_ref=ptr("composition_model.fc.weight"))
    for t in range(hp.n_conv):
        s.atom[t] = gated(f"atom_conv_layers.{t}.twoBody_atom", "layers.0", "layers.3", f"atom_conv_layers.{t}.mlp_out.layers.1")
    for t in range(hp.n_conv - 1):
        s.bond[t] = gated(f"bond_conv_layers.{t}.twoBody_bond", "layers.0", "layers.3", f"bond_conv_layers.{t}.mlp_out.layers.1")
        s.angle[t] = gated(f"angle_layers.{t}.twoBody_bond", "layers.1", None, None)
    hidden = sorted(int(k.split(".")[2]) for k in sd if k.startswith("mlp.layers.") and k.endswith(".weight") and sd[k].shape[0] == 64)
    last = max(int(k.split(".")[2]) for k in sd if k.startswith("mlp.layers.") and k.endswith(".weight"))
    for l, i in enumerate(hidden):
        s.mlp_w[l], s.mlp_b[l] = ptr(f"mlp.layers.{i}.weight"), ptr(f"mlp.layers.{i}.bias")
    s.mlp_last_w, s.mlp_last_b = ptr(f"mlp.layers.{last}.weight"), float(sd[f"mlp.layers.{last}.bias"].reshape(-1)[0])
    n = int(lib.chg_packed_floats(ctypes.byref(hps)))
    if n <= 0:
        raise ChgnetB200Error(f"chg_packed_floats: unsupported hyper-parameters ({lib.chg_last_error().decode()})")
    blob = torch.zeros(n, dtype=torch.float32)
    _check(lib, lib.chg_pack_weights_host(ctypes.byref(hps), ctypes.byref(s), ctypes.cast(blob.data_ptr(), FP)), "chg_pack_weights_host")
    return hps, blob, hp


def batch_struct(b: DeviceBatch, graph_ptr: Tensor | None = None) -> Batch:
    bs = Batch(n_atoms=b.n_atoms, n_edges=b.n_edges, n_bonds=b.n_bonds, n_angles=b.n_angles, n_graphs=b.n_graphs,
               n_short=b.n_short)
    for name in _BATCH_PTRS:
        t = graph_ptr if name == "graph_ptr" else getattr(b, name, None)
        setattr(bs, name, None if t is None or t.numel() == 0 else t.data_ptr())
    return bs


def plan(hps: HParams, sizes: Batch, wanted: Outputs, want_trace: bool = False) -> tuple[int, list[str]]:
    """(workspace bytes, kernel call list) for these sizes — no GPU needed."""
    lib = _bind(load_library())
    need = c_size_t(0)
    buf = ctypes.create_string_buffer(1 << 16) if want_trace else None
    _check(lib, lib.chg_forward_plan(ctypes.byref(hps), ctypes.byref(sizes), ctypes.byref(wanted), ctypes.byref(need), buf,
                                     (1 << 16) if want_trace else 0), "chg_forward_plan")
    return int(need.value), (buf.value.decode().split() if want_trace else [])


class NativeForward:
    """Packed weights on the device + a grow-only workspace; ``__call__`` = one ``chg_forward``."""

    def __init__(self, state_dict: dict, model_args: dict | None, device: torch.device) -> None:
        self.lib = _bind(load_library())
        if device.type != "cuda":
            raise ChgnetB200Error("chgnet_b200 has no CPU path: NativeForward needs a CUDA device")
        self.hps, blob, self.hp = pack_weights_native(state_dict, model_args)
        self.device = device
        self.weights = blob.to(device)
        self.workspace = torch.empty(0, dtype=torch.uint8, device=device)
        self.atom_ref = self.weights.new_zeros(94)
        if "composition_model.fc.weight" in state_dict:
            self.atom_ref = state_dict["composition_model.fc.weight"].detach().reshape(-1).to(device=device, dtype=torch.float32)
        self.calls = 0
        self._ws_version = 0          # bumped whenever the workspace is reallocated (captured graphs point into it)
        self._graphs: dict = {}       # (id(batch), flags) -> captured CUDA graph of one chg_forward
        self._seen: dict = {}

    def reserve(self, b: DeviceBatch, *, need_grad: bool = True) -> None:
        """Grow the workspace for this batch NOW (e.g. before a CUDA-graph capture, where it must not be reallocated)."""
        outs = Outputs(energy=1, e_ref=1, site_e=1, force=1 if need_grad else None, virial=1 if need_grad else None)
        need, _ = plan(self.hps, batch_struct(b, None), outs)
        if self.workspace.numel() < need + 256:
            self.workspace = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=self.device)
            self._ws_version += 1

    def replay(self, b: DeviceBatch, *, need_grad: bool, need_magmom: bool = False, need_atom_fea: bool = False) -> dict[str, Tensor]:
        """``chg_forward`` for a batch descriptor that is evaluated again and again (same ``DeviceBatch`` object: fixed
        topology, coordinates / lattice updated in place): the second call captures the launches of one forward into a
        CUDA graph, later calls replay it (one graph launch instead of ~130 kernel launches - small systems are
        launch-bound).  The returned tensors are REUSED by every replay: consume them before the next call."""
        flags = (bool(need_grad), bool(need_magmom), bool(need_atom_fea))
        key = (id(b), flags)
        ent = self._graphs.get(key)
        if ent is not None and ent["batch"] is b and ent["ws"] == self._ws_version:
            ent["graph"].replay()
            self.calls += 1
            return ent["res"]
        kw = dict(need_grad=need_grad, need_magmom=need_magmom, need_atom_fea=need_atom_fea)
        res = self(b, **kw)  # eager: sizes the workspace, sets the kernels' one-time attributes
        seen = self._seen.get(key)
        if seen is None or seen[0] is not b:
            self._seen = {key: (b, 1)}  # one candidate at a time
            return res
        dev = self.device
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # thread_local: CUDA calls of other threads (an NCCL watchdog, a clock sampler) must not invalidate this capture
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            captured = self(b, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        if len(self._graphs) >= 8:  # small cache: drop the oldest entry
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = {"graph": graph, "res": captured, "batch": b, "ws": self._ws_version}
        self._seen = {}
        return res

    def __call__(self, b: DeviceBatch, *, need_grad: bool, need_magmom: bool = False, need_atom_fea: bool = False,
                 need_crystal_fea: bool = False) -> dict[str, Tensor]:
        dev, N, B = self.device, b.n_atoms, b.n_graphs
        f32, f64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.float64, device=dev)
        res: dict[str, Tensor] = {"energy": torch.empty(B, **f64), "e_ref": torch.empty(B, **f64), "site_e": torch.empty(N, **f32)}
        if need_magmom:
            res["magmom"] = torch.empty(N, **f32)
        if need_atom_fea:
            res["atom_fea"] = torch.empty(N, 64, **f32)
        gptr = None
        if need_crystal_fea:
            res["crystal_fea"] = torch.empty(B, 64, **f32)
            gptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            gptr[1:] = torch.cumsum(torch.tensor(b.atoms_per_graph, device=dev), 0)
        if need_grad:
            res["force"], res["virial"] = torch.empty(N, 3, **f64), torch.empty(B, 9, **f64)
        outs = Outputs(**{k: v.data_ptr() for k, v in res.items()})
        bs = batch_struct(b, gptr)
        need, _ = plan(self.hps, bs, outs)
        if self.workspace.numel() < need + 256:
            self.workspace = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=dev)
            self._ws_version += 1
        base = (self.workspace.data_ptr() + 255) // 256 * 256
        room = self.workspace.numel() - (base - self.workspace.data_ptr())
        with torch.cuda.device(dev):  # launches go to the model's device and its current stream
            rc = self.lib.chg_forward(ctypes.byref(self.hps), self.weights.data_ptr(), ctypes.byref(bs), ctypes.byref(outs), base, room,
                                      torch.cuda.current_stream(dev).cuda_stream)
        _check(self.lib, rc, "chg_forward")
        self.calls += 1
        return res


def packed_layout(hps: HParams) -> dict[str, tuple[int, int]]:
    """name -> (offset, floats) of the packed blob; mirrors ``walk`` in csrc/native_engine.cu (tests only)."""
    R, NA, L = hps.num_radial, hps.num_angular, hps.n_readout_hidden
    F = (NA - 1) // 2
    out: dict[str, tuple[int, int]] = {}
    off = 0

    def take(name: str, n: int) -> None:
        nonlocal off
        out[name] = (off, n)
        off += (n + 15) // 16 * 16

    for name, n in (("emb", 94 * 64), ("freq_ag", R), ("freq_bg", R), ("freq_ang", F), ("w3t", 3 * R * 64), ("w3", 3 * 64 * R),
                    ("wang_t", NA * 64), ("wang", 64 * NA)):
        take(name, n)

    def second(k: str) -> None:
        take(f"{k}.w2t", 64 * 128), take(f"{k}.w2", 128 * 64), take(f"{k}.b2", 128)

    for t in range(hps.n_conv):
        k = f"atom.{t}"
        second(k)
        if hps.use_ln:
            take(f"{k}.ln", 256)
        take(f"{k}.wcn_t", 64 * 256), take(f"{k}.we_t", 64 * 128), take(f"{k}.b1", 128)
        take(f"{k}.wcn_b", 256 * 64), take(f"{k}.we_b", 128 * 64), take(f"{k}.wo_t", 4096), take(f"{k}.wo", 4096)
        if hps.has_mlp_out_bias:
            take(f"{k}.bo", 64)
    for t in range(hps.n_conv - 1):
        for kind in ("bond", "angle"):
            k = f"{kind}.{t}"
            if kind == "bond":
                second(k)
            if hps.use_ln:
                take(f"{k}.ln", 256)
            take(f"{k}.wij_t", 64 * 256), take(f"{k}.bij", 256), take(f"{k}.wx_t", 64 * 128), take(f"{k}.w1a_t", 64 * 128)
            take(f"{k}.wij_b", 256 * 64), take(f"{k}.wx_b", 128 * 64), take(f"{k}.w1a_b", 128 * 64)
            if kind == "bond":
                take(f"{k}.wo_t", 4096), take(f"{k}.wo", 4096)
                if hps.has_mlp_out_bias:
                    take(f"{k}.bo", 64)
    if hps.readout_ln:
        take("readout_ln", 128)
    take("mlp_wt", L * 4096), take("mlp_w", L * 4096), take("mlp_b", L * 64), take("w_last", 64), take("w_mag", 64)
    take("atom_ref", 94)
    out["__total__"] = (off, 0)
    return out


__all__ = ["NativeForward", "pack_weights_native", "plan", "packed_layout", "HParams", "Batch", "Outputs", "batch_struct",
           "hparams_struct"]
