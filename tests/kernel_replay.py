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
    "bond_basis_embed": [8, 9, 10, 11],
    "bond_basis_bwd": [11, 12],
    "angle_basis_embed": [5, 6],
    "angle_basis_bwd": [6, 7],
    "linear": [4],
    "gather_rows": [2],
    "scatter_rows": [2],
    "atom_conv_fwd": [9, 10, 11],
    "atom_conv_bwd": [10, 11, 12, 13],
    "segment_sum": [4],
    "atom_conv_fused": [10, 11],
    "bond_conv_fused": [11, 12, 13],
    "bond_conv_fwd": [10, 11, 12],
    "bond_conv_bwd": [8, 9, 10, 11, 12],
    "angle_update_fwd": [8, 9],
    "angle_update_bwd": [3, 4],
    "readout": [10, 11, 12, 13, 14],
    "magmom": [3],
    "force_virial": [10, 11],
    # training (trailing optional outputs above are training-only too)
    "wgrad": [2, 3],
    "colsum": [1],
    "readout_bwd": [7, 8, 9, 10, 11],
    "magmom_bwd": [4, 5],
    # second order
    "edge_tangent": [8, 9],
    "bond_basis_tangent": [9, 10, 11, 12],
    "bond_basis_bwd2": [12],
    "angle_basis_tangent": [6, 7],
    "angle_basis_bwd2": [7],
    "atom_conv_tan": [11, 12, 13],
    "atom_conv_bwd2": [13, 14, 15, 16],
    "bond_conv_tan": [12, 13, 14],
    "bond_conv_bwd2": [13, 14, 15, 16, 17],
    "angle_update_tan": [9, 10],
    "angle_update_bwd2": [5, 6],
    "readout_bwd2": [8, 9, 10, 11, 12, 13, 14, 15, 16],
}
SECOND_ORDER_KERNELS = {"edge_tangent", "bond_basis_tangent", "bond_basis_bwd2", "angle_basis_tangent", "angle_basis_bwd2",
                        "atom_conv_tan", "atom_conv_bwd2", "bond_conv_tan", "bond_conv_bwd2", "angle_update_tan",
                        "angle_update_bwd2", "readout_bwd2"}
# the unfused message kernels only run in training mode (they also save `pre`); inference uses the fused pair
TRAIN_KERNELS = {"wgrad", "colsum", "readout_bwd", "magmom_bwd", "atom_conv_fwd", "bond_conv_fwd"} | SECOND_ORDER_KERNELS
INFER_KERNELS = set(OUT_ARGS) - TRAIN_KERNELS


class RecordingKernels(SpecKernels):
    def __init__(self) -> None:
        self.calls: list[tuple[str, list, dict[int, torch.Tensor]]] = []

    def __getattribute__(self, name):
        attr = super().__getattribute__(name)
        if name in OUT_ARGS and callable(attr):
            def wrapped(*args):
                snap = [a.detach().clone().contiguous() if isinstance(a, torch.Tensor) else a for a in args]
                attr(*args)
                outs = {i: args[i].detach().clone().contiguous() for i in OUT_ARGS[name]
                        if i < len(args) and args[i] is not None}
                self.calls.append((name, snap, outs))
            return wrapped
        return attr
