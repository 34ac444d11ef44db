"""Packed on-disk graph store.

The reference writes ONE ``torch.save`` pickle per graph (``CrystalGraph.save``,
chgnet/graph/crystalgraph.py:138-155; ``examples/make_graphs.py``) and the dataset re-opens and
un-pickles a file per sample (``GraphData.__getitem__``, chgnet/data/dataset.py:386-434).  Here a
whole dataset is ONE file: a small JSON header, then every field of every graph concatenated into one
contiguous block per field (int32 / fp32, the dtypes of the hot path) with per-graph row offsets.
``GraphStore`` memory-maps the file; ``store[i]`` / ``store.batch(indices)`` hand out ``CrystalGraph``
objects whose tensors are zero-copy views of the map, which ``build_batch`` concatenates straight into
its pinned staging buffer — no unpickling, no per-graph file open.

Layout (little endian): ``b"CHGPACK1"`` | u64 header bytes | header JSON | padding to 64 B | blocks.
"""
from __future__ import annotations

import json
import os
import warnings
from collections.abc import Sequence

import numpy as np
import torch

from chgnet_b200.graph import CrystalGraph

MAGIC = b"CHGPACK1"
# field -> (dtype, inner width, which per-graph count gives its rows)
_FIELDS = {
    "atomic_number": ("<i4", 1, "n_atoms"),
    "atom_frac_coord": ("<f4", 3, "n_atoms"),
    "atom_graph": ("<i4", 2, "n_edges"),
    "neighbor_image": ("<f4", 3, "n_edges"),
    "directed2undirected": ("<i4", 1, "n_edges"),
    "undirected2directed": ("<i4", 1, "n_bonds"),
    "bond_graph": ("<i4", 5, "n_angles"),
    "lattice": ("<f4", 3, "three"),
}


def save_graphs(path: str, graphs: Sequence[CrystalGraph]) -> str:
    """Write ``graphs`` as one packed file; returns ``path``."""
    counts = {
        "n_atoms": [int(g.atomic_number.shape[0]) for g in graphs],
        "n_edges": [int(g.directed2undirected.shape[0]) for g in graphs],
        "n_bonds": [int(g.undirected2directed.shape[0]) for g in graphs],
        "n_angles": [int(g.bond_graph.reshape(-1, 5).shape[0]) for g in graphs],
        "three": [3] * len(graphs),
    }
    header = {
        "n_graphs": len(graphs), "counts": {k: v for k, v in counts.items() if k != "three"},
        "atom_graph_cutoff": [float(g.atom_graph_cutoff) for g in graphs],
        "bond_graph_cutoff": [float(g.bond_graph_cutoff) for g in graphs],
        "graph_id": [g.graph_id for g in graphs], "mp_id": [getattr(g, "mp_id", None) for g in graphs],
        "composition": [getattr(g, "composition", None) for g in graphs], "blocks": {},
    }
    off = 0
    for name, (dt, width, cnt) in _FIELDS.items():
        nbytes = sum(counts[cnt]) * width * 4
        header["blocks"][name] = [off, nbytes]
        off += (nbytes + 63) // 64 * 64
    hjson = json.dumps(header).encode()
    data_start = (len(MAGIC) + 8 + len(hjson) + 63) // 64 * 64
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(hjson)).tobytes())
        f.write(hjson)
        f.write(b"\0" * (data_start - f.tell()))
        for name, (dt, width, _) in _FIELDS.items():
            start = f.tell()
            for g in graphs:
                t = getattr(g, name).detach().cpu().reshape(-1, width)
                f.write(np.ascontiguousarray(t.numpy(), dtype=dt).tobytes())
            f.write(b"\0" * ((-(f.tell() - start)) % 64))
    os.replace(tmp, path)
    return path


class GraphStore:
    """Memory-mapped reader of a ``save_graphs`` file: ``len``, ``store[i]``, ``store.batch(ids)``."""

    def __init__(self, path: str) -> None:
        self.path = path
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{path} is not a chgnet_b200 packed graph file")
            hlen = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            self.header = json.loads(f.read(hlen).decode())
        data_start = (16 + hlen + 63) // 64 * 64
        self._map = np.memmap(path, dtype=np.uint8, mode="r")
        c = dict(self.header["counts"])
        c["three"] = [3] * self.header["n_graphs"]
        self._arr, self._rows = {}, {}
        for name, (dt, width, cnt) in _FIELDS.items():
            off, nbytes = self.header["blocks"][name]
            flat = self._map[data_start + off : data_start + off + nbytes].view(dt)
            self._arr[name] = flat.reshape(-1, width) if width > 1 else flat
            self._rows[name] = np.concatenate([[0], np.cumsum(c[cnt])]).astype(np.int64)

    def __len__(self) -> int:
        return int(self.header["n_graphs"])

    def __getitem__(self, i: int) -> CrystalGraph:
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        h = self.header
        t = {}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)  # "array is not writable": intended (read-only map)
            for name in _FIELDS:
                lo, hi = self._rows[name][i], self._rows[name][i + 1]
                # zero-copy view of the (read-only) map; the hot path never writes to graph tensors
                t[name] = torch.from_numpy(self._arr[name][lo:hi])
        return CrystalGraph(atom_graph_cutoff=h["atom_graph_cutoff"][i], bond_graph_cutoff=h["bond_graph_cutoff"][i],
                            graph_id=h["graph_id"][i], mp_id=h["mp_id"][i], composition=h["composition"][i], **t)

    def batch(self, indices: Sequence[int]) -> list[CrystalGraph]:
        return [self[int(i)] for i in indices]


def load_graphs(path: str) -> list[CrystalGraph]:
    store = GraphStore(path)
    return store.batch(range(len(store)))
