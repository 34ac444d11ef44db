"""Structure -> batch descriptor entirely on the device.

``build_batch_device`` is the GPU replacement of the reference's per-MD-step host pipeline
``CrystalGraphConverter.forward`` (reference chgnet/graph/converter.py:102-190: pymatgen neighbour list ->
``create_graph.c`` / ``cygraph.pyx`` -> ``Graph.line_graph_adjacency_list``) + ``BatchedGraph.from_graphs``
(model.py:792-913) for ONE structure: fractional coordinates that already live on the device go through
``chg_graph_build_device`` (csrc/graph_device.cu: cell list, neighbour search, edge pairing, bond graph) and
``chg_build_csr`` (csrc/batch_csr.cu) and come out as a :class:`chgnet_b200.batch.DeviceBatch` - no host graph,
no H2D copy of index arrays.  The integer arrays are bit-identical to the host builders
(tests/test_graph_device_gpu.py).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import Tensor

from chgnet_b200._lib import ChgnetB200Error, load_library
from chgnet_b200.batch import MAX_Z, DeviceBatch, _csr_device, _pack_lib

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = load_library()
        P, I, D = ctypes.c_void_p, ctypes.c_int32, ctypes.c_double
        lib.chg_graph_device_scratch_bytes.restype = ctypes.c_int64
        lib.chg_graph_device_scratch_bytes.argtypes = [I, I]
        lib.chg_graph_build_device.restype = I
        lib.chg_graph_build_device.argtypes = [P, P, I, D, D, I, I] + [P] * 13 + [P]
        lib.chg_bond_graph_count.restype = I
        lib.chg_bond_graph_count.argtypes = [P, P, I, I, P, P, P]
        _LIB = lib
    return _LIB


class DeviceGraphBuilder:
    """Grow-only device buffers + the two C calls; one instance per model / MD run."""

    def __init__(self, device: torch.device | str, atom_graph_cutoff: float = 6.0, bond_graph_cutoff: float = 3.0) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ChgnetB200Error("DeviceGraphBuilder needs a CUDA device (the host builders are chg_graph_build / graphgen)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.r_atom, self.r_bond = float(atom_graph_cutoff), float(bond_graph_cutoff)
        self.cap_edges = self.cap_angles = 0
        self.buf: dict[str, Tensor] = {}
        self.scratch: Tensor | None = None
        self.last_sizes = (0, 0, 0)

    def _ensure(self, n_atoms: int, cap_edges: int, cap_angles: int) -> None:
        dev = self.device
        if cap_edges > self.cap_edges or "ptr_c" not in self.buf or self.buf["ptr_c"].numel() < n_atoms + 1:
            self.cap_edges = max(cap_edges, self.cap_edges)
            i32 = dict(dtype=torch.int32, device=dev)
            self.buf.update(center=torch.empty(self.cap_edges, **i32), nbr=torch.empty(self.cap_edges, **i32),
                            d2u=torch.empty(self.cap_edges, **i32), u2d=torch.empty(self.cap_edges // 2 + 1, **i32),
                            image=torch.empty(self.cap_edges, 3, dtype=torch.float32, device=dev),
                            ptr_c=torch.empty(n_atoms + 1, **i32))
            need = int(_lib().chg_graph_device_scratch_bytes(n_atoms, self.cap_edges))
            self.scratch = torch.empty(need + 8192, dtype=torch.uint8, device=dev)
        if cap_angles > self.cap_angles or "ang_atom" not in self.buf:
            self.cap_angles = max(cap_angles, self.cap_angles)
            for k in ("ang_atom", "ang_i", "ang_di", "ang_j", "ang_dj"):
                self.buf[k] = torch.empty(self.cap_angles, dtype=torch.int32, device=dev)

    def graph_arrays(self, frac64: Tensor, lattice) -> dict[str, Tensor]:
        """frac64 [N,3] fp64 on the device, lattice 3x3 (host array-like) -> views of the edge / angle arrays."""
        lib = _lib()
        dev = self.device
        n = int(frac64.shape[0])
        if frac64.dtype != torch.float64 or frac64.device != dev or not frac64.is_contiguous():
            raise ChgnetB200Error("frac must be a contiguous fp64 tensor on the builder's device")
        lat = np.ascontiguousarray(np.asarray(lattice, dtype=np.float64).reshape(3, 3))
        vol = abs(float(np.linalg.det(lat)))
        if n and self.cap_edges == 0:  # first guess from the number density; a too small capacity is reported with the need
            rho = n / max(vol, 1e-9)
            self._ensure(n, int(n * (4.19 * self.r_atom ** 3 * rho) * 1.3) + 4096,
                         int(n * max(4.19 * self.r_bond ** 3 * rho, 1.0) ** 2 * 1.5) + 4096)
        else:
            self._ensure(n, self.cap_edges, self.cap_angles)
        sizes = (ctypes.c_int32 * 4)()
        for _ in range(3):
            b = self.buf
            with torch.cuda.device(dev):
                rc = lib.chg_graph_build_device(frac64.data_ptr(), lat.ctypes.data, n, self.r_atom, self.r_bond, self.cap_edges,
                                                self.cap_angles, b["center"].data_ptr(), b["nbr"].data_ptr(), b["image"].data_ptr(),
                                                b["d2u"].data_ptr(), b["u2d"].data_ptr(), b["ptr_c"].data_ptr(),
                                                b["ang_atom"].data_ptr(), b["ang_i"].data_ptr(), b["ang_di"].data_ptr(),
                                                b["ang_j"].data_ptr(), b["ang_dj"].data_ptr(), self.scratch.data_ptr(), sizes,
                                                torch.cuda.current_stream(dev).cuda_stream)
            if rc == 0:
                break
            msg = lib.chg_last_error().decode()
            if "capacity" not in msg:
                raise (ValueError if "not complete" in msg else ChgnetB200Error)(msg)
            self._ensure(n, int(max(sizes[0], self.cap_edges) * 1.25) + 1024, int(max(sizes[2], self.cap_angles, 1) * 1.25) + 1024)
        else:
            raise ChgnetB200Error("chg_graph_build_device: capacity negotiation did not converge")
        ed, eu, an = int(sizes[0]), int(sizes[1]), int(sizes[2])
        self.last_sizes = (ed, eu, an)
        b = self.buf
        out = {k: b[k][:ed] for k in ("center", "nbr", "d2u")}
        out["image"] = b["image"][:ed]
        out["u2d"] = b["u2d"][:eu]
        out["ptr_c"] = b["ptr_c"][: n + 1]
        for k in ("ang_atom", "ang_i", "ang_di", "ang_j", "ang_dj"):
            out[k] = b[k][:an]
        return out

    def build_batch(self, atomic_numbers, frac64: Tensor, lattice, *, with_reverse: bool = True,
                    compact_bonds: bool = True) -> DeviceBatch:
        """One structure -> :class:`DeviceBatch` (a 1-graph batch).  ``atomic_numbers``: host array-like or device int32
        tensor; ``frac64`` [N,3] fp64 on the device; ``lattice`` 3x3 on the host."""
        dev = self.device
        n = int(frac64.shape[0])
        if torch.is_tensor(atomic_numbers) and atomic_numbers.device == dev:
            z = atomic_numbers.to(torch.int32).contiguous()
            zmin, zmax = (int(z.min()), int(z.max())) if n else (1, 1)
        else:
            zh = np.asarray(atomic_numbers, dtype=np.int64).reshape(-1)
            zmin, zmax = (int(zh.min()), int(zh.max())) if n else (1, 1)
            z = torch.as_tensor(zh, dtype=torch.int32).to(dev)
        if zmin < 1 or zmax > MAX_Z:
            raise IndexError(f"index out of range in self: atomic numbers span [{zmin}, {zmax}], outside [1, {MAX_Z}]")
        g = self.graph_arrays(frac64, lattice)
        ed, eu, an = self.last_sizes
        lib = _lib()
        _pack_lib()  # binds chg_build_csr
        es = -1
        if an and compact_bonds:
            cnt = ctypes.c_int32(0)
            scr = torch.empty(2 * (eu + 1) + 4096 + 64, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                rc = lib.chg_bond_graph_count(g["ang_i"].data_ptr(), g["ang_j"].data_ptr(), an, eu, scr.data_ptr(), ctypes.byref(cnt),
                                              torch.cuda.current_stream(dev).cuda_stream)
            if rc != 0:
                raise ChgnetB200Error(lib.chg_last_error().decode())
            es = int(cnt.value)
        # the batch owns copies of the index arrays (the builder's buffers are reused by the next call)
        own = {k: v.clone() for k, v in g.items()}
        v = _csr_device(dev, n, ed, eu, an, es, with_reverse, own["center"], own["nbr"], own["d2u"], own["ang_atom"], own["ang_i"],
                        own["ang_j"])
        i32 = dict(dtype=torch.int32, device=dev)
        if es >= 0:
            short_ids, ang_is, ang_js, ptr_is, perm_js, ptr_js, n_short = (v["short_ids"], v["ang_is"], v["ang_js"], v["ptr_is"],
                                                                           v["perm_j"], v["ptr_js"], es)
        else:
            short_ids = torch.arange(eu, **i32)
            ang_is, ang_js, ptr_is, perm_js, ptr_js, n_short = own["ang_i"], own["ang_j"], v["ptr_i"], v["perm_j"], v["ptr_j"], eu
        lat = torch.as_tensor(np.asarray(lattice, dtype=np.float64).reshape(1, 9), dtype=torch.float32).to(dev)
        L = lat.view(1, 3, 3)
        volume = (L[:, 0] * torch.linalg.cross(L[:, 1], L[:, 2])).sum(dim=1)
        return DeviceBatch(
            n_graphs=1, n_atoms=n, n_edges=ed, n_bonds=eu, n_angles=an, atoms_per_graph=[n],
            z=z, frac=frac64.to(torch.float32).contiguous(), owner=torch.zeros(n, **i32), lattice=lat, volume=volume,
            center=own["center"], nbr=own["nbr"], image=own["image"], d2u=own["d2u"], u2d=own["u2d"],
            ptr_c=v["ptr_c"], perm_n=v["perm_n"], ptr_n=v["ptr_n"], perm_u=v["perm_u"], ptr_u=v["ptr_u"],
            ang_atom=own["ang_atom"], ang_i=own["ang_i"], ang_j=own["ang_j"], ang_di=own["ang_di"], ang_dj=own["ang_dj"],
            ptr_i=v["ptr_i"], perm_j=v["perm_j"], ptr_j=v["ptr_j"], perm_x=v["perm_x"], ptr_x=v["ptr_x"],
            n_short=n_short, short_ids=short_ids, ang_is=ang_is, ang_js=ang_js, ptr_is=ptr_is, perm_js=perm_js, ptr_js=ptr_js,
            h2d_bytes=4 * n + 36,
        )


def crystal_graph_from_device(builder: DeviceGraphBuilder, atomic_numbers, frac, lattice, graph_id=None):
    """A host ``CrystalGraph`` built by the DEVICE builder (for comparison with the host builders and for callers that
    want the reference's input type)."""
    from chgnet_b200.graph import TORCH_DTYPE, CrystalGraph

    f64 = torch.as_tensor(np.ascontiguousarray(np.asarray(frac, dtype=np.float64).reshape(-1, 3))).to(builder.device).contiguous()
    g = builder.graph_arrays(f64, lattice)
    torch.cuda.synchronize(builder.device)
    host = {k: v.cpu() for k, v in g.items()}
    if host["ang_atom"].numel():
        bond_graph = torch.stack([host["ang_atom"], host["ang_i"], host["ang_di"], host["ang_j"], host["ang_dj"]], dim=1)
    else:
        bond_graph = torch.zeros(0, 5, dtype=torch.int32)
    return CrystalGraph(
        atomic_number=torch.as_tensor(np.asarray(atomic_numbers), dtype=torch.int32),
        atom_frac_coord=torch.as_tensor(np.asarray(frac), dtype=TORCH_DTYPE),
        atom_graph=torch.stack([host["center"], host["nbr"]], dim=1).reshape(-1, 2),
        atom_graph_cutoff=builder.r_atom, neighbor_image=host["image"].reshape(-1, 3),
        directed2undirected=host["d2u"], undirected2directed=host["u2d"], bond_graph=bond_graph.reshape(-1, 5).contiguous(),
        bond_graph_cutoff=builder.r_bond, lattice=torch.as_tensor(np.asarray(lattice), dtype=TORCH_DTYPE), graph_id=graph_id)


__all__ = ["DeviceGraphBuilder", "crystal_graph_from_device"]
