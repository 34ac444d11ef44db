"""Batch descriptor: list[CrystalGraph] -> one SoA on the device.

Replaces the per-graph Python loop of the reference's
``BatchedGraph.from_graphs`` (reference chgnet/model/model.py:792-913): the
host concatenates every field of every graph into ONE int32 and ONE fp32 pinned
staging buffer (two H2D copies per batch, instead of 8 per graph,
crystalgraph.py:102-118), index offsets are applied on the device, and the
CSR row pointers / gather permutations that make every scatter-add of the model
an atomics-free segmented reduction are built with a handful of device sorts.

Segment structures (all int32, device):
  ptr_c   [N+1]   directed edges grouped by center (edges are center-sorted)
  perm_n  [Ed], ptr_n [N+1]     directed edges grouped by NEIGHBOUR
  perm_u  [Ed]    directed edges grouped by undirected bond (2 per bond)
  ptr_u   [Eu+1]  = 2*arange
  ptr_i   [Eu+1]  angles grouped by bond i (angles are i-sorted)
  perm_j  [A], ptr_j [Eu+1]     angles grouped by bond j
  perm_x  [A], ptr_x [N+1]      angles grouped by center atom
  short_ids [Es], ang_is/ang_js [A], ptr_is, perm_js/ptr_js [Es+1]
                  the same angle groupings in the COMPACT index space of the bonds that
                  appear in the bond graph (about 1/8 of all bonds at 6 A / 3 A cutoffs)
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch
from torch import Tensor


@dataclass
class DeviceBatch:
    n_graphs: int
    n_atoms: int
    n_edges: int  # directed
    n_bonds: int  # undirected
    n_angles: int
    atoms_per_graph: list[int]
    # atoms
    z: Tensor
    frac: Tensor
    owner: Tensor
    lattice: Tensor  # [B, 9]
    volume: Tensor  # [B]
    # directed edges (center-sorted)
    center: Tensor
    nbr: Tensor
    image: Tensor
    d2u: Tensor
    u2d: Tensor
    ptr_c: Tensor
    perm_n: Tensor
    ptr_n: Tensor
    perm_u: Tensor
    ptr_u: Tensor
    # angles (bond-i sorted)
    ang_atom: Tensor
    ang_i: Tensor
    ang_j: Tensor
    ang_di: Tensor
    ang_dj: Tensor
    ptr_i: Tensor
    perm_j: Tensor
    ptr_j: Tensor
    perm_x: Tensor
    ptr_x: Tensor
    # bond-graph bonds ("short" bonds, d < bond_graph_cutoff): the only bonds BondConv /
    # AngleUpdate read or write.  Angles address them through compact slots.
    n_short: int = 0
    short_ids: Tensor | None = None  # [Es] undirected bond index of each slot (ascending)
    ang_is: Tensor | None = None  # [A] slot of bond i   (non-decreasing)
    ang_js: Tensor | None = None  # [A] slot of bond j
    ptr_is: Tensor | None = None  # [Es+1] angles by slot i
    perm_js: Tensor | None = None  # [A], ptr_js [Es+1]: angles by slot j
    ptr_js: Tensor | None = None
    h2d_bytes: int = 0


MAX_Z = 94  # rows of the atom embedding / AtomRef tables (reference encoders.py:22-32)

_STAGING: dict = {}  # (dtype, pinned) -> [buffer, event of the last H2D copy that read it]


def _staging_buffer(n: int, dtype, pin: bool) -> Tensor:
    """Grow-only host staging buffer (pinned when the target is a CUDA device), reused across
    batches so that cudaHostAlloc is not paid per call."""
    key = (dtype, pin)
    slot = _STAGING.get(key)
    if slot is not None and slot[1] is not None:
        slot[1].synchronize()  # the previous batch's copy out of this buffer has finished
        slot[1] = None
    if slot is None or slot[0].numel() < n:
        cap = max(n, int(1.5 * slot[0].numel()) if slot is not None else n)
        slot = [torch.empty(cap, dtype=dtype, pin_memory=pin), None]
        _STAGING[key] = slot
    return slot[0][:n]


_DEFAULT_STAGING = "pinned"  # "wc": write-combined staging for the wire packer (env CHGNET_B200_STAGING overrides)
_WC_STAGING: dict = {}  # name -> [pointer, capacity in bytes, event of the last copy out of it]


def _wc_staging(name: str, n_bytes: int) -> int:
    """Grow-only WRITE-COMBINED pinned staging buffer (chg_host_alloc), returned as a raw pointer: the wire packer's
    threads only write it and the copy engine reads it at full PCIe rate (ordinary pinned memory written by many
    cores copies several times slower: its lines sit dirty in those cores' caches).  Waits for the previous batch's
    copy out of the buffer before handing it out again."""
    slot = _WC_STAGING.get(name)
    if slot is not None and slot[2] is not None:
        slot[2].synchronize()
        slot[2] = None
    if slot is None or slot[1] < n_bytes:
        lib = _pack_lib()
        cap = max(n_bytes, int(1.5 * slot[1]) if slot is not None else n_bytes, 4096)
        ptr = ctypes.c_void_p()
        if lib.chg_host_alloc(cap, 1, ctypes.byref(ptr)) != 0:
            raise RuntimeError(f"chg_host_alloc failed: {lib.chg_last_error().decode()}")
        if slot is not None:
            lib.chg_host_free(ctypes.c_void_p(slot[0]))
        slot = [ptr.value, cap, None]
        _WC_STAGING[name] = slot
    return slot[0]


def _mark_wc_in_flight(device: torch.device) -> None:
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    for slot in _WC_STAGING.values():
        slot[2] = ev


def _mark_staging_in_flight(dtype, pin: bool, device: torch.device) -> None:
    if device.type == "cuda":
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        _STAGING[(dtype, pin)][1] = ev


def _is_sorted(t: Tensor) -> bool:
    if t.numel() < 2:
        return True
    if t.device.type == "cpu":
        a = t.numpy()
        return bool(np.all(a[1:] >= a[:-1]))
    return bool((t[1:] >= t[:-1]).all())


def _csr_ptr(sorted_keys: Tensor, n_rows: int) -> Tensor:
    edges = torch.arange(n_rows + 1, device=sorted_keys.device, dtype=sorted_keys.dtype)
    return torch.searchsorted(sorted_keys.contiguous(), edges).to(torch.int32)


def _group(keys: Tensor, n_rows: int) -> tuple[Tensor, Tensor]:
    """(perm, ptr) such that rows perm[ptr[r]:ptr[r+1]] are the items with key r."""
    if keys.numel() == 0:
        z = torch.zeros(n_rows + 1, dtype=torch.int32, device=keys.device)
        return keys.to(torch.int32), z
    sk, perm = torch.sort(keys, stable=True)
    return perm.to(torch.int32), _csr_ptr(sk, n_rows)


_PACK_LIB = None


class _CsrIn(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_atoms", "n_edges", "n_bonds", "n_angles", "n_short", "with_reverse")] + \
               [(n, ctypes.c_void_p) for n in ("center", "nbr", "d2u", "ang_atom", "ang_i", "ang_j")]


_CSR_OUT = ("ptr_c", "perm_n", "ptr_n", "perm_u", "ptr_u", "ptr_i", "perm_j", "ptr_j", "perm_x", "ptr_x",
            "short_ids", "ang_is", "ang_js", "ptr_is", "ptr_js")


class _CsrOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _CSR_OUT]


_CSR_SCRATCH: dict = {}  # device -> grow-only int32 scratch of chg_build_csr


def _csr_device(device, N, Ed, Eu, A, Es, with_reverse, center, nbr, d2u, ang_atom, ang_i, ang_j) -> dict:
    """All segment structures of the batch in ONE C call (csrc/batch_csr.cu): boundary searches for the sorted
    keys, counting sorts for the transposed groupings, the compact bond-graph slots - no ATen sorts, no host
    synchronisation.  ``Es`` < 0: no compaction (every bond keeps its index)."""
    import ctypes

    lib = _pack_lib()
    sizes = {"ptr_c": N + 1, "perm_n": Ed, "ptr_n": N + 1, "perm_u": Ed, "ptr_u": Eu + 1, "ptr_i": Eu + 1, "perm_j": A,
             "ptr_j": Eu + 1, "perm_x": A, "ptr_x": N + 1, "short_ids": max(Es, 0), "ang_is": A if Es >= 0 else 0,
             "ang_js": A if Es >= 0 else 0, "ptr_is": Es + 1 if Es >= 0 else 0, "ptr_js": Es + 1 if Es >= 0 else 0}
    if not with_reverse:
        for k in ("perm_n", "perm_u", "perm_j", "perm_x", "ptr_js"):
            sizes[k] = 0
    pad = lambda n: (n + 3) // 4 * 4  # noqa: E731  16-byte aligned pieces
    buf = torch.empty(sum(pad(n) for n in sizes.values()) + 4, dtype=torch.int32, device=device)
    views, off = {}, 0
    for k in _CSR_OUT:
        views[k] = buf[off : off + sizes[k]]
        off += pad(sizes[k])
    need = int(lib.chg_build_csr_scratch_ints(N, Ed, Eu, A))
    scr = _CSR_SCRATCH.get(device)
    if scr is None or scr.numel() < need:
        scr = torch.empty(int(need * 1.25) + 16, dtype=torch.int32, device=device)
        _CSR_SCRATCH[device] = scr
    ptr = lambda t: t.data_ptr() if t is not None and t.numel() else None  # noqa: E731
    cin = _CsrIn(n_atoms=N, n_edges=Ed, n_bonds=Eu, n_angles=A, n_short=Es, with_reverse=int(with_reverse),
                 center=ptr(center), nbr=ptr(nbr), d2u=ptr(d2u), ang_atom=ptr(ang_atom), ang_i=ptr(ang_i), ang_j=ptr(ang_j))
    # zero-size outputs still get a valid (never written) address so that the C side can tell "absent" from "empty"
    cout = _CsrOut(**{k: (views[k].data_ptr() if sizes[k] else buf[-4:].data_ptr()) for k in _CSR_OUT})
    with torch.cuda.device(device):
        rc = lib.chg_build_csr(ctypes.byref(cin), ctypes.byref(cout), scr.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    if rc != 0:
        raise RuntimeError(f"chg_build_csr failed: {lib.chg_last_error().decode()}")
    if not with_reverse:
        i32 = dict(dtype=torch.int32, device=device)
        views["ptr_n"] = views["ptr_x"] = torch.zeros(N + 1, **i32)
        views["ptr_u"] = views["ptr_j"] = torch.zeros(Eu + 1, **i32)
        if Es >= 0:
            views["ptr_js"] = torch.zeros(Es + 1, **i32)
    return views


def _pack_lib():
    global _PACK_LIB
    if _PACK_LIB is None:
        import ctypes

        from chgnet_b200._lib import load_library

        lib = load_library()
        lib.chg_pack_batch_host.restype = ctypes.c_int32
        lib.chg_pack_batch_host.argtypes = [ctypes.c_int32] + [ctypes.c_void_p] * 5
        lib.chg_pack_batch_wire.restype = ctypes.c_int32
        lib.chg_pack_batch_wire.argtypes = [ctypes.c_int32] + [ctypes.c_void_p] * 10
        lib.chg_host_alloc.restype = ctypes.c_int32
        lib.chg_host_alloc.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]
        lib.chg_host_free.restype = ctypes.c_int32
        lib.chg_host_free.argtypes = [ctypes.c_void_p]
        lib.chg_build_csr_scratch_ints.restype = ctypes.c_int64
        lib.chg_build_csr_scratch_ints.argtypes = [ctypes.c_int32] * 4
        lib.chg_build_csr.restype = ctypes.c_int32
        lib.chg_build_csr.argtypes = [ctypes.POINTER(_CsrIn), ctypes.POINTER(_CsrOut), ctypes.c_void_p, ctypes.c_void_p]
        _PACK_LIB = lib
    return _PACK_LIB


_INT_FIELDS = ("atomic_number", "directed2undirected", "undirected2directed")
_FLT_FIELDS = ("atom_frac_coord", "neighbor_image", "lattice")


def _graphs_are_packable(graphs, ag_l, bg_l) -> bool:
    """Host tensors with the hot path's dtypes (int32 / fp32), contiguous: what the C packer memcpy's.
    Checked once per graph object (the verdict is cached on it)."""
    for g, ag, bg in zip(graphs, ag_l, bg_l):
        ok = getattr(g, "_packable", None)
        if ok is None:
            try:
                ok = (all(getattr(g, f).dtype == torch.int32 and getattr(g, f).is_contiguous() and getattr(g, f).device.type == "cpu"
                          for f in _INT_FIELDS)
                      and all(getattr(g, f).dtype == torch.float32 and getattr(g, f).is_contiguous() and getattr(g, f).device.type == "cpu"
                              for f in _FLT_FIELDS)
                      and ag.dtype == torch.int32 and ag.is_contiguous() and bg.dtype == torch.int32 and bg.is_contiguous()
                      and g.lattice.numel() == 9 and ag is g.atom_graph and bg is g.bond_graph or
                      False)
                if not ok and (ag is not g.atom_graph or bg is not g.bond_graph):
                    # reshaped empty graphs (isolated atoms): nothing to copy from them, dtype still matters
                    ok = (all(getattr(g, f).dtype == torch.int32 and getattr(g, f).is_contiguous() for f in _INT_FIELDS)
                          and all(getattr(g, f).dtype == torch.float32 and getattr(g, f).is_contiguous() for f in _FLT_FIELDS)
                          and g.lattice.numel() == 9 and ag.numel() * (ag is not g.atom_graph) == 0
                          and bg.numel() * (bg is not g.bond_graph) == 0 and ag.dtype == torch.int32 and bg.dtype == torch.int32
                          and ag.is_contiguous() and bg.is_contiguous())
                try:
                    g._packable = bool(ok)
                except AttributeError:
                    pass
            except AttributeError:
                ok = False
        if not ok:
            return False
    return True


def build_batch(graphs: Sequence, device: torch.device | str, *, with_reverse: bool = True,
                compact_bonds: bool = True, native_pack: bool = True, native_csr: bool = True,
                wire: bool | None = None, packed: tuple | None = None) -> DeviceBatch:
    """list[CrystalGraph] -> DeviceBatch.  ``native_pack`` / ``native_csr`` = False select the older torch
    implementations of the host packing / the segment structures (kept as the checkers of the C paths).
    ``wire`` (default: on; env CHGNET_B200_WIRE=0 turns it off) ships the compact wire format of csrc/batch_wire.cu
    (derivable bond-graph columns and fp32 images are re-created on the device, packing overlaps the copies) and falls
    back to the full format for graphs that do not satisfy its (verified) assumptions.
    ``packed = (counts int64 [B][4], ptrs uint64 [B][8])`` (``graphs`` = None) batches graphs that exist only as host
    arrays - the rows ``CrystalGraph.pack_info`` produces; used for the handles of ``chg_graph_build_many``."""
    if wire is None:
        wire = os.environ.get("CHGNET_B200_WIRE", "1") != "0"
    device = torch.device(device)
    B = len(graphs) if packed is None else int(packed[0].shape[0])
    # fast path: chgnet_b200.CrystalGraph caches its sizes and raw data pointers (graph.pack_info), so the per-graph
    # Python work is one attribute call; any other graph-like object goes through the generic reads below
    fast = None
    if packed is not None:
        fast = (np.ascontiguousarray(packed[0], dtype=np.int64), np.ascontiguousarray(packed[1], dtype=np.uint64))
    elif native_pack and B > 0:
        infos = [g.pack_info() if hasattr(g, "pack_info") else False for g in graphs]
        if all(i is not False for i in infos):
            fast = (np.stack([i[0] for i in infos]), np.stack([i[1] for i in infos]))
    ag_l, bg_l = [], []
    if fast is not None:
        cnt = fast[0]
        n_at, n_ed, n_eu, n_an = (cnt[:, k].tolist() for k in range(4))
        bad = np.nonzero((cnt[:, 1] != 2 * cnt[:, 2]))[0]
        if len(bad) or (packed is None and any(g.directed2undirected.shape[0] != e for g, e in zip(graphs, n_ed))):
            raise ValueError("CrystalGraph invariant violated: n_directed != 2 * n_undirected")
    else:
        # per-graph Python work is kept to attribute reads (no reshape / detach calls per tensor)
        for g in graphs:
            ag, bg = g.atom_graph, g.bond_graph
            if ag.dim() != 2:  # structure with every atom isolated (model.py:841-843)
                ag = ag.reshape(0, 2)
            if bg.dim() != 2:
                bg = bg.reshape(0, 5)
            ag_l.append(ag)
            bg_l.append(bg)
        n_at = [g.atomic_number.shape[0] for g in graphs]
        n_ed = [a.shape[0] for a in ag_l]
        n_eu = [g.undirected2directed.shape[0] for g in graphs]
        n_an = [b.shape[0] for b in bg_l]
        for g, e_d, e_u in zip(graphs, n_ed, n_eu):
            if e_d != 2 * e_u or g.directed2undirected.shape[0] != e_d:
                raise ValueError("CrystalGraph invariant violated: n_directed != 2 * n_undirected")
    N, Ed, Eu, A = sum(n_at), sum(n_ed), sum(n_eu), sum(n_an)

    src_dev = graphs[0].atomic_number.device if (B and packed is None) else torch.device("cpu")

    def pack_legacy():
        src_dev = graphs[0].atomic_number.device if B else torch.device("cpu")
        host_src = src_dev.type == "cpu"
        pin = host_src and device.type == "cuda"

        # (name, per-graph tensors, inner width, total rows)
        int_fields = [
            ("z", [g.atomic_number for g in graphs], 1, N),
            ("ag", ag_l, 2, Ed),
            ("d2u", [g.directed2undirected for g in graphs], 1, Ed),
            ("u2d", [g.undirected2directed for g in graphs], 1, Eu),
            ("bg", bg_l, 5, A),
        ]
        flt_fields = [
            ("frac", [g.atom_frac_coord for g in graphs], 3, N),
            ("image", [g.neighbor_image for g in graphs], 3, Ed),
            ("lattice", [g.lattice for g in graphs], 3, 3 * B),
        ]
        counts_host = np.array([n_at, n_ed, n_eu, n_an], dtype=np.int32).reshape(-1)  # [4*B]
        n_int = sum(f[2] * f[3] for f in int_fields) + counts_host.size
        n_flt = sum(f[2] * f[3] for f in flt_fields)

        def as_np(t: Tensor) -> np.ndarray:
            return t.detach().numpy() if t.requires_grad else t.numpy()

        def stage(fields, total, dtype, tail=None):
            """Concatenate all parts into one staging buffer and ship it with one copy.  Host
            graphs are packed with numpy (single-threaded memcpy: torch CPU ops fork an OpenMP
            team per call, which costs milliseconds on a many-core host) into a cached pinned
            buffer; device-resident graphs are concatenated on the device."""
            views, off = {}, 0
            if host_src:
                buf = _staging_buffer(total, dtype, pin)
                host = buf.numpy()
                for name, parts, width, rows in fields:
                    size = width * rows
                    if size:
                        dst = host[off : off + size]
                        np.concatenate([as_np(p) for p in parts], axis=0,
                                       out=dst if width == 1 else dst.reshape(rows, width), casting="unsafe")
                    views[name] = (off, size)
                    off += size
                if tail is not None:
                    host[off : off + tail.size] = tail
                    views["_tail"] = (off, tail.size)
                # a CPU "device" would alias the reused staging buffer: copy instead
                dbuf = buf.clone() if device.type == "cpu" else buf.to(device, non_blocking=True)
                _mark_staging_in_flight(dtype, pin, device)
                return dbuf, views, host
            buf = torch.empty(total, dtype=dtype, device=src_dev)
            for name, parts, width, rows in fields:
                size = width * rows
                if size:
                    torch.cat([p.detach().reshape(-1).to(dtype) for p in parts], out=buf[off : off + size])
                views[name] = (off, size)
                off += size
            if tail is not None:
                buf[off : off + tail.size] = torch.from_numpy(tail).to(src_dev)
                views["_tail"] = (off, tail.size)
            return buf.to(device), views, None

        ibuf, iv, ihost = stage(int_fields, n_int, torch.int32, counts_host)
        fbuf, fv, _ = stage(flt_fields, n_flt, torch.float32)
        h2d = (n_int + n_flt) * 4 if src_dev != device else 0

        # sortedness (decides whether the device has to reorder): one vectorised pass over the
        # concatenated raw indices; a decrease is only legal at a graph boundary
        def sorted_within_graphs(raw: np.ndarray | Tensor, sizes: list[int]) -> bool:
            if len(raw) < 2:
                return True
            if isinstance(raw, Tensor):
                return all(_is_sorted(t) for t in torch.split(raw, sizes))
            dec = raw[1:] < raw[:-1]
            ends = np.cumsum(sizes)[:-1] - 1
            ends = ends[(ends >= 0) & (ends < len(dec))]
            dec[ends] = False
            return not bool(dec.any())

        if ihost is not None:
            o, sz = iv["ag"]
            edges_sorted = sorted_within_graphs(ihost[o : o + sz : 2].copy(), n_ed)
            o, sz = iv["bg"]
            angles_sorted = sorted_within_graphs(ihost[o + 1 : o + sz : 5].copy(), n_an)
        else:
            edges_sorted = all(_is_sorted(a[:, 0]) for a in ag_l)
            angles_sorted = all(_is_sorted(b[:, 1]) for b in bg_l)

        def iview(name):
            o, s = iv[name]
            return ibuf[o : o + s]

        def fview(name):
            o, s = fv[name]
            return fbuf[o : o + s]

        cnt = iview("_tail").view(4, B).long()  # device copies of the per-graph counts
        # exclusive prefix sums = per-graph offsets
        offs = torch.cumsum(cnt, dim=1) - cnt
        atom_off, ed_off, eu_off = offs[0], offs[1], offs[2]

        def rep(vals: Tensor, which: int, total: int) -> Tensor:
            return torch.repeat_interleave(vals, cnt[which], output_size=total).to(torch.int32)

        z = iview("z")
        if N:
            zsrc = ihost[iv["z"][0] : iv["z"][0] + N] if ihost is not None else z
            zmin, zmax = int(zsrc.min()), int(zsrc.max())
            if zmin < 1 or zmax > MAX_Z:
                raise IndexError(f"index out of range in self: atomic numbers span [{zmin}, {zmax}], outside [1, {MAX_Z}]")
        owner = rep(torch.arange(B, device=device), 0, N)
        ag = iview("ag").view(Ed, 2)
        e_atom_off = rep(atom_off, 1, Ed)
        center = (ag[:, 0] + e_atom_off).contiguous()
        nbr = (ag[:, 1] + e_atom_off).contiguous()
        d2u = iview("d2u") + rep(eu_off, 1, Ed)
        u2d = iview("u2d") + rep(ed_off, 2, Eu)
        image = fview("image").view(Ed, 3)
        bg = iview("bg").view(A, 5)
        a_atom_off, a_eu_off, a_ed_off = rep(atom_off, 3, A), rep(eu_off, 3, A), rep(ed_off, 3, A)
        ang_atom = bg[:, 0] + a_atom_off
        ang_i = bg[:, 1] + a_eu_off
        ang_di = bg[:, 2] + a_ed_off
        ang_j = bg[:, 3] + a_eu_off
        ang_dj = bg[:, 4] + a_ed_off
        return (z, owner, center, nbr, d2u, u2d, image, fview("frac").view(N, 3), fview("lattice").view(B, 9),
                ang_atom, ang_i, ang_di, ang_j, ang_dj, edges_sorted, angles_sorted, h2d)

    def pack_native():
        """ONE pass over host memory in the C library (chg_pack_batch_host): concatenation, index offsets,
        owners and the sortedness flags, written straight into the cached (pinned) staging buffers."""
        import ctypes

        lib = _pack_lib()
        n_int, n_flt = 2 * N + 3 * Ed + Eu + 5 * A, 3 * N + 3 * Ed + 9 * B
        pin = device.type == "cuda"
        ibuf_h, fbuf_h = _staging_buffer(max(n_int, 1), torch.int32, pin), _staging_buffer(max(n_flt, 1), torch.float32, pin)
        if fast is not None:
            counts, ptrs = np.ascontiguousarray(fast[0]), np.ascontiguousarray(fast[1])
        else:
            counts = np.ascontiguousarray(np.array([n_at, n_ed, n_eu, n_an], dtype=np.int64).T)
            ptrs = np.fromiter((t.data_ptr() for g, ag, bg in zip(graphs, ag_l, bg_l)
                                for t in (g.atomic_number, g.atom_frac_coord, ag, g.neighbor_image, g.directed2undirected,
                                          g.undirected2directed, bg, g.lattice)), dtype=np.uint64, count=8 * B)
        flags = (ctypes.c_int32 * 4)()
        rc = lib.chg_pack_batch_host(B, counts.ctypes.data, ptrs.ctypes.data, ibuf_h.data_ptr(), fbuf_h.data_ptr(), flags)
        if rc != 0:
            raise RuntimeError(f"chg_pack_batch_host failed: {lib.chg_last_error().decode()}")
        if flags[2] >= 0:  # nn.Embedding(94, .) of the reference raises the same (tests/test_encoders.py:25-28)
            raise IndexError(f"index out of range in self: atomic number {int(ibuf_h[flags[2]])} of atom {int(flags[2])} "
                             f"is outside [1, {MAX_Z}]")
        on_cpu = device.type == "cpu"  # a CPU "device" would alias the reused staging buffer: copy instead
        ibuf = ibuf_h[:n_int].clone() if on_cpu else ibuf_h[:n_int].to(device, non_blocking=True)
        _mark_staging_in_flight(torch.int32, pin, device)
        fbuf = fbuf_h[:n_flt].clone() if on_cpu else fbuf_h[:n_flt].to(device, non_blocking=True)
        _mark_staging_in_flight(torch.float32, pin, device)
        sizes = (N, N, Ed, Ed, Ed, Eu, A, A, A, A, A)
        z, owner, center, nbr, d2u, u2d, ang_atom, ang_i, ang_di, ang_j, ang_dj = torch.split(ibuf, sizes)
        frac, image, lattice = torch.split(fbuf, (3 * N, 3 * Ed, 9 * B))
        nonlocal n_short_host
        n_short_host = int(flags[3])
        return (z, owner, center, nbr, d2u, u2d, image.view(Ed, 3), frac.view(N, 3), lattice.view(B, 9), ang_atom, ang_i,
                ang_di, ang_j, ang_dj, bool(flags[0]), bool(flags[1]), (n_int + n_flt) * 4 if device.type != "cpu" else 0)

    def pack_wire():
        """chg_pack_batch_wire: compact wire format, two-phase pack + copy overlap, expansion on the device.  Returns
        None when some graph does not satisfy the format's assumptions (the caller then ships the full format)."""
        import ctypes

        lib = _pack_lib()
        n_int_w, n_flt_w = 2 * N + 3 * Ed + Eu + 2 * A, 3 * N + 9 * B
        cuda = device.type == "cuda"
        wc = cuda and os.environ.get("CHGNET_B200_STAGING", _DEFAULT_STAGING) == "wc"
        if wc:  # write-combined pinned staging: raw pointers, never read on the CPU
            ip, fp, gp = (_wc_staging("int", 4 * max(n_int_w, 1)), _wc_staging("flt", 4 * max(n_flt_w, 1)),
                          _wc_staging("img", max(3 * Ed, 1)))
        else:
            ibuf_h = _staging_buffer(max(n_int_w, 1), torch.int32, cuda)
            fbuf_h = _staging_buffer(max(n_flt_w, 1), torch.float32, cuda)
            img_h = _staging_buffer(max(3 * Ed, 1), torch.int8, cuda)
            ip, fp, gp = ibuf_h.data_ptr(), fbuf_h.data_ptr(), img_h.data_ptr()
        if fast is not None:
            counts, ptrs = np.ascontiguousarray(fast[0]), np.ascontiguousarray(fast[1])
        else:
            counts = np.ascontiguousarray(np.array([n_at, n_ed, n_eu, n_an], dtype=np.int64).T)
            ptrs = np.fromiter((t.data_ptr() for g, ag, bg in zip(graphs, ag_l, bg_l)
                                for t in (g.atomic_number, g.atom_frac_coord, ag, g.neighbor_image, g.directed2undirected,
                                          g.undirected2directed, bg, g.lattice)), dtype=np.uint64, count=8 * B)
        flags = (ctypes.c_int32 * 5)()
        if cuda:
            ibuf = torch.empty(n_int_w + 3 * A, dtype=torch.int32, device=device)
            fbuf = torch.empty(n_flt_w + 3 * Ed, dtype=torch.float32, device=device)
            img_d = torch.empty(max(3 * Ed, 1), dtype=torch.int8, device=device)
            with torch.cuda.device(device):
                rc = lib.chg_pack_batch_wire(B, counts.ctypes.data, ptrs.ctypes.data, ip, fp, gp, ibuf.data_ptr(),
                                             fbuf.data_ptr(), img_d.data_ptr(), flags,
                                             torch.cuda.current_stream(device).cuda_stream)
            if wc:
                _mark_wc_in_flight(device)
            else:
                for dt in (torch.int32, torch.float32, torch.int8):
                    _mark_staging_in_flight(dt, True, device)
        else:
            rc = lib.chg_pack_batch_wire(B, counts.ctypes.data, ptrs.ctypes.data, ip, fp, gp, None, None, None, flags, None)
        if rc != 0:
            raise RuntimeError(f"chg_pack_batch_wire failed: {lib.chg_last_error().decode()}")
        if flags[2] >= 0:  # nn.Embedding(94, .) of the reference raises the same (tests/test_encoders.py:25-28)
            bad = ctypes.c_int32.from_address(ip + 4 * int(flags[2])).value
            raise IndexError(f"index out of range in self: atomic number {bad} of atom {int(flags[2])} "
                             f"is outside [1, {MAX_Z}]")
        if flags[4] != 0:
            return None
        if cuda:
            z, owner, center, nbr, d2u, u2d, ang_di, ang_dj, ang_atom, ang_i, ang_j = torch.split(
                ibuf, (N, N, Ed, Ed, Ed, Eu, A, A, A, A, A))
            frac, lattice, image = torch.split(fbuf, (3 * N, 9 * B, 3 * Ed))
        else:  # CPU "device" (tests): expand on the host what the two device kernels re-create
            z, owner, center, nbr, d2u, u2d, ang_di, ang_dj = (t.clone() for t in torch.split(
                ibuf_h[:n_int_w], (N, N, Ed, Ed, Ed, Eu, A, A)))
            frac, lattice = (t.clone() for t in torch.split(fbuf_h[:n_flt_w], (3 * N, 9 * B)))
            image = img_h[: 3 * Ed].to(torch.float32)
            ang_atom, ang_i, ang_j = center[ang_di.long()], d2u[ang_di.long()], d2u[ang_dj.long()]
        nonlocal n_short_host
        n_short_host = int(flags[3])
        return (z, owner, center, nbr, d2u, u2d, image.view(Ed, 3), frac.view(N, 3), lattice.view(B, 9), ang_atom, ang_i,
                ang_di, ang_j, ang_dj, bool(flags[0]), bool(flags[1]), (n_int_w + n_flt_w) * 4 + 3 * Ed if cuda else 0)

    n_short_host = -1  # number of bond-graph bonds, counted by the C packer (no device sync needed later)
    use_native = fast is not None or (native_pack and B > 0 and src_dev.type == "cpu" and _graphs_are_packable(graphs, ag_l, bg_l))
    packed = pack_wire() if (use_native and wire) else None
    if packed is None:
        packed = pack_native() if use_native else pack_legacy()
    (z, owner, center, nbr, d2u, u2d, image, frac_t, lattice, ang_atom, ang_i, ang_di, ang_j, ang_dj, edges_sorted,
     angles_sorted, h2d) = packed

    if not edges_sorted and Ed:
        # stable sort by center; remap everything that stores a directed index
        _, perm = torch.sort(center, stable=True)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(Ed, device=device)
        center, nbr, d2u, image = center[perm], nbr[perm], d2u[perm], image[perm]
        u2d = inv[u2d.long()].to(torch.int32)
        if A:
            ang_di = inv[ang_di.long()].to(torch.int32)
            ang_dj = inv[ang_dj.long()].to(torch.int32)
    if not angles_sorted and A:
        _, perm = torch.sort(ang_i, stable=True)
        ang_atom, ang_i, ang_j, ang_di, ang_dj = (t[perm] for t in (ang_atom, ang_i, ang_j, ang_di, ang_dj))

    i32 = dict(dtype=torch.int32, device=device)
    if device.type == "cuda" and use_native and native_csr and n_short_host >= 0:
        # ---- device CSR build: one C call, no sorts, no host sync (csrc/batch_csr.cu) ----
        center, nbr, d2u = center.contiguous(), nbr.contiguous(), d2u.contiguous()
        ang_atom, ang_i, ang_j = ang_atom.contiguous(), ang_i.contiguous(), ang_j.contiguous()
        do_compact = bool(A and compact_bonds)
        v = _csr_device(device, N, Ed, Eu, A, n_short_host if do_compact else -1, with_reverse, center, nbr, d2u,
                        ang_atom, ang_i, ang_j)
        ptr_c, ptr_i = v["ptr_c"], v["ptr_i"]
        perm_n, ptr_n, perm_u, ptr_u = v["perm_n"], v["ptr_n"], v["perm_u"], v["ptr_u"]
        perm_j, ptr_j, perm_x, ptr_x = v["perm_j"], v["ptr_j"], v["perm_x"], v["ptr_x"]
        if do_compact:
            short_ids, ang_is, ang_js, ptr_is = v["short_ids"], v["ang_is"], v["ang_js"], v["ptr_is"]
            perm_js, ptr_js = perm_j, v["ptr_js"]  # slot numbers are monotone in the bond index: same permutation
            Es = n_short_host
        else:
            short_ids = torch.arange(Eu, **i32)
            ang_is, ang_js, ptr_is, perm_js, ptr_js, Es = ang_i, ang_j, ptr_i, perm_j, ptr_j, Eu
    else:
        ptr_c = _csr_ptr(center, N)
        ptr_i = _csr_ptr(ang_i, Eu)
        if with_reverse:
            perm_n, ptr_n = _group(nbr, N)
            perm_u, _ = _group(d2u, Eu)
            ptr_u = torch.arange(Eu + 1, **i32) * 2
            perm_j, ptr_j = _group(ang_j, Eu)
            perm_x, ptr_x = _group(ang_atom, N)
        else:
            empty = torch.zeros(0, **i32)
            perm_n = perm_u = perm_j = perm_x = empty
            ptr_n = ptr_x = torch.zeros(N + 1, **i32)
            ptr_u = ptr_j = torch.zeros(Eu + 1, **i32)

        # ---- compact index space of the bond-graph bonds (one host sync: the slot count) ----
        if A and compact_bonds:
            mask = torch.zeros(Eu, dtype=torch.int32, device=device)
            mask[ang_i.long()] = 1
            mask[ang_j.long()] = 1
            slot = torch.cumsum(mask, 0, dtype=torch.int32) - 1
            short_ids = torch.nonzero(mask).view(-1).to(torch.int32)  # syncs; ascending
            ang_is, ang_js = slot[ang_i.long()], slot[ang_j.long()]
        else:  # identity "compaction": every bond has a slot
            short_ids = torch.arange(Eu, **i32)
            ang_is, ang_js = ang_i, ang_j
        Es = int(short_ids.shape[0])
        ptr_is = _csr_ptr(ang_is, Es)
        if with_reverse:
            perm_js, ptr_js = _group(ang_js, Es)
        else:
            perm_js, ptr_js = torch.zeros(0, **i32), torch.zeros(Es + 1, **i32)

    L = lattice.view(B, 3, 3)
    volume = (L[:, 0] * torch.linalg.cross(L[:, 1], L[:, 2])).sum(dim=1)  # model.py:834-836

    c = lambda t: t.contiguous()  # noqa: E731
    return DeviceBatch(
        n_graphs=B, n_atoms=N, n_edges=Ed, n_bonds=Eu, n_angles=A, atoms_per_graph=n_at,
        z=c(z), frac=c(frac_t), owner=c(owner), lattice=c(lattice), volume=volume,
        center=c(center), nbr=c(nbr), image=c(image), d2u=c(d2u), u2d=c(u2d),
        ptr_c=ptr_c, perm_n=perm_n, ptr_n=ptr_n, perm_u=perm_u, ptr_u=ptr_u,
        ang_atom=c(ang_atom), ang_i=c(ang_i), ang_j=c(ang_j), ang_di=c(ang_di), ang_dj=c(ang_dj),
        ptr_i=ptr_i, perm_j=perm_j, ptr_j=ptr_j, perm_x=perm_x, ptr_x=ptr_x,
        n_short=Es, short_ids=c(short_ids), ang_is=c(ang_is), ang_js=c(ang_js), ptr_is=ptr_is,
        perm_js=perm_js, ptr_js=ptr_js, h2d_bytes=h2d,
    )


def partition_graphs(costs: Sequence[float], n_ranks: int) -> list[list[int]]:
    """Greedy longest-processing-time partition of graph indices over ranks
    (SURVEY.md §8e): whole graphs per rank, no device-path collective."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * n_ranks
    parts: list[list[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: loads[k])
        parts[r].append(i)
        loads[r] += costs[i]
    for p in parts:
        p.sort()
    return parts


def graph_cost(g) -> float:
    """Work estimate of one graph: message rows (edges + 2.5 x angles)."""
    return float(g.atom_graph.reshape(-1, 2).shape[0]) + 2.5 * float(g.bond_graph.reshape(-1, 5).shape[0])
