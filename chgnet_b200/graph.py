"""``CrystalGraph`` — the input type of the hot path.

Host-side mirror of the reference container (reference
chgnet/graph/crystalgraph.py:15-118).  Field names, dtypes (fp32 / int32,
crystalgraph.py:12, converter.py:143-159), the ``len(d2u) == 2 * len(u2d)``
invariant (crystalgraph.py:96-100) and ``.to(device)`` returning a NEW graph
(crystalgraph.py:102-118) are kept so that graphs built by the reference's
converter can be handed to :class:`chgnet_b200.CHGNet` unchanged (the model only
reads the attributes below — any object that has them is accepted).
"""
from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

TORCH_DTYPE = torch.float32

_TENSOR_FIELDS = (
    "atomic_number",
    "atom_frac_coord",
    "atom_graph",
    "neighbor_image",
    "directed2undirected",
    "undirected2directed",
    "bond_graph",
    "lattice",
)


class CrystalGraph:
    """Crystal graph: atoms, directed bonds (atom graph) and angles (bond graph).

    atomic_number        int32 [n]
    atom_frac_coord      fp32  [n, 3]
    atom_graph           int32 [e_d, 2]   (center, neighbor)
    neighbor_image       fp32  [e_d, 3]
    directed2undirected  int32 [e_d]
    undirected2directed  int32 [e_u]      e_d == 2 * e_u
    bond_graph           int32 [a, 5]     (atom, undirected i, directed i,
                                           undirected j, directed j)
    lattice              fp32  [3, 3]
    """

    def __init__(
        self,
        atomic_number: Tensor,
        atom_frac_coord: Tensor,
        atom_graph: Tensor,
        atom_graph_cutoff: float,
        neighbor_image: Tensor,
        directed2undirected: Tensor,
        undirected2directed: Tensor,
        bond_graph: Tensor,
        bond_graph_cutoff: float,
        lattice: Tensor,
        graph_id: str | None = None,
        mp_id: str | None = None,
        composition: str | None = None,
    ) -> None:
        self.atomic_number = atomic_number
        self.atom_frac_coord = atom_frac_coord
        self.atom_graph = atom_graph
        self.atom_graph_cutoff = atom_graph_cutoff
        self.neighbor_image = neighbor_image
        self.directed2undirected = directed2undirected
        self.undirected2directed = undirected2directed
        self.bond_graph = bond_graph
        self.bond_graph_cutoff = bond_graph_cutoff
        self.lattice = lattice
        self.graph_id = graph_id
        self.mp_id = mp_id
        self.composition = composition
        if len(directed2undirected) != 2 * len(undirected2directed):
            raise ValueError(
                f"{graph_id} number of directed indices ({len(directed2undirected)}) !="
                f" 2 * number of undirected indices ({2 * len(undirected2directed)})!"
            )

    def __setattr__(self, name: str, value: Any) -> None:
        if name in _TENSOR_FIELDS:
            object.__setattr__(self, "_pack", None)  # cached raw-pointer view of the tensors (pack_info) is stale
        object.__setattr__(self, name, value)

    def pack_info(self):
        """``(counts, ptrs)`` for the C batch packer (``chg_pack_batch_host``): int64 [4] = atoms, directed edges,
        bonds, angles and uint64 [8] = the data pointers of the eight tensor fields - or ``False`` when the tensors are
        not host-resident, contiguous int32 / fp32 (the packer then goes through the generic torch path).  Cached; the
        cache is dropped whenever a tensor field is reassigned."""
        cached = self.__dict__.get("_pack")
        if cached is not None:
            return cached
        import numpy as np

        ints = (self.atomic_number, self.atom_graph, self.directed2undirected, self.undirected2directed, self.bond_graph)
        flts = (self.atom_frac_coord, self.neighbor_image, self.lattice)
        ok = (all(t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu" for t in ints)
              and all(t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu" for t in flts)
              and self.lattice.numel() == 9
              and (self.atom_graph.dim() == 2 or self.atom_graph.numel() == 0)
              and (self.bond_graph.dim() == 2 or self.bond_graph.numel() == 0))
        if not ok:
            info = False
        else:
            n_ed = self.atom_graph.shape[0] if self.atom_graph.dim() == 2 else 0
            n_an = self.bond_graph.shape[0] if self.bond_graph.dim() == 2 else 0
            counts = np.array([self.atomic_number.shape[0], n_ed, self.undirected2directed.shape[0], n_an], dtype=np.int64)
            ptrs = np.array([t.data_ptr() for t in (self.atomic_number, self.atom_frac_coord, self.atom_graph, self.neighbor_image,
                                                    self.directed2undirected, self.undirected2directed, self.bond_graph,
                                                    self.lattice)], dtype=np.uint64)
            info = (counts, ptrs)
        object.__setattr__(self, "_pack", info)
        return info

    def to(self, device: str | torch.device = "cpu") -> CrystalGraph:
        """Return a copy of the graph with every tensor on ``device``."""
        kw = self.to_dict()
        for name in _TENSOR_FIELDS:
            kw[name] = kw[name].to(device)
        return CrystalGraph(**kw)

    def to_dict(self) -> dict[str, Any]:
        return {
            "atomic_number": self.atomic_number,
            "atom_frac_coord": self.atom_frac_coord,
            "atom_graph": self.atom_graph,
            "atom_graph_cutoff": self.atom_graph_cutoff,
            "neighbor_image": self.neighbor_image,
            "directed2undirected": self.directed2undirected,
            "undirected2directed": self.undirected2directed,
            "bond_graph": self.bond_graph,
            "bond_graph_cutoff": self.bond_graph_cutoff,
            "lattice": self.lattice,
            "graph_id": self.graph_id,
            "mp_id": self.mp_id,
            "composition": self.composition,
        }

    @classmethod
    def from_dict(cls, dic: dict[str, Any]) -> CrystalGraph:
        return cls(**dic)

    def save(self, fname: str) -> str:
        torch.save(self.to_dict(), f=fname)
        return fname

    @classmethod
    def from_file(cls, file_name: str) -> CrystalGraph:
        return cls.from_dict(torch.load(file_name, weights_only=False))

    def __repr__(self) -> str:
        return (
            f"CrystalGraph({self.composition}, num_atoms={len(self.atomic_number)}, "
            f"num_bonds={len(self.atom_graph)}, num_angles={len(self.bond_graph)})"
        )

    @property
    def num_isolated_atoms(self) -> int:
        return len(self.atomic_number) - len(torch.unique(self.atom_graph[:, 0]))


def is_graph_like(obj: Any) -> bool:
    """True for this class and for the reference's own ``CrystalGraph``."""
    return all(hasattr(obj, f) for f in _TENSOR_FIELDS)
