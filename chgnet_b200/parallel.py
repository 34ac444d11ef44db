"""Multi-GPU: shard independent crystal graphs over ranks (one process per GPU).

Graphs never share edges (index offsets only, reference chgnet/model/model.py:856-877), so
inference shards by whole graphs with NO device-path collective: each rank evaluates its
share and the per-graph results are gathered on the host in the original order.  The
only collective of the hot path is the gradient all-reduce of a training step.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch.distributed as dist

from chgnet_b200.batch import graph_cost, partition_graphs


def shard_indices(graphs: Sequence, world_size: int) -> list[list[int]]:
    """Balanced assignment of graph indices to ranks (LPT on edges + 2.5 x angles)."""
    return partition_graphs([graph_cost(g) for g in graphs], world_size)


def predict_sharded(graphs: Sequence, predict_fn: Callable[[list], list[dict]], group=None) -> list[dict]:
    """Every rank calls this with the SAME list of graphs; rank r evaluates its share with
    ``predict_fn`` (e.g. ``lambda gs: model.predict_graph(gs, task="efs", batch_size=len(gs))``)
    and every rank returns the full list of per-graph result dicts in input order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(predict_fn(list(graphs)))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = shard_indices(graphs, world)
    mine = parts[rank]
    local = list(predict_fn([graphs[i] for i in mine])) if mine else []
    if len(local) != len(mine):
        raise RuntimeError("predict_fn must return one result per graph")
    gathered: list = [None] * world
    dist.all_gather_object(gathered, list(zip(mine, local)), group=group)
    out: list = [None] * len(graphs)
    for chunk in gathered:
        for idx, res in chunk:
            out[idx] = res
    return out
