"""Record every kernel call of an engine run executed with the torch specifications on
the CPU, so that the same calls can be replayed through the CUDA library and compared
output by output (tests/test_kernels_gpu.py)."""
from __future__ import annotations

import torch

from oracle.kernel_specs import SpecKernels

# positional indices of the output arguments of each kernel method
OUT_ARGS = {
    "embed_atoms": [2],
    "edge_geometry": [6, 7, 8],
    "bond_basis_embed": [8, 9, 10],
    "bond_basis_bwd": [11],
    "angle_basis_embed": [5],
    "angle_basis_bwd": [6],
    "linear": [4],
    "gather_rows": [2],
    "scatter_rows": [2],
    "atom_conv_fwd": [9, 10],
    "atom_conv_bwd": [10, 11],
    "segment_sum": [4],
    "bond_conv_fwd": [10, 11, 12],
    "bond_conv_bwd": [8, 9, 10],
    "angle_update_fwd": [8, 9],
    "angle_update_bwd": [3],
    "readout": [10, 11, 12, 13, 14],
    "magmom": [3],
    "force_virial": [10, 11],
}


class RecordingKernels(SpecKernels):
    def __init__(self) -> None:
        self.calls: list[tuple[str, list, dict[int, torch.Tensor]]] = []

    def __getattribute__(self, name):
        attr = super().__getattribute__(name)
        if name in OUT_ARGS and callable(attr):
            def wrapped(*args):
                snap = [a.detach().clone().contiguous() if isinstance(a, torch.Tensor) else a for a in args]
                attr(*args)
                outs = {i: args[i].detach().clone().contiguous() for i in OUT_ARGS[name] if args[i] is not None}
                self.calls.append((name, snap, outs))
            return wrapped
        return attr
