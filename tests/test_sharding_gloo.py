"""CPU, world_size 2 over gloo: the N>1 host logic (graph partition + ordered gather).
The per-graph evaluation is the oracle here — the device path has no collective."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chgnet_b200 import graphgen
from chgnet_b200.batch import graph_cost, partition_graphs


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 10, size=257).tolist()
    for world in (1, 2, 4, 8):
        parts = partition_graphs(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(257))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs) + 1e-9
        assert all(p == sorted(p) for p in parts)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from chgnet_b200.parallel import predict_sharded
        from oracle import chgnet_oracle as orc

        w = orc.load_weights_npz(os.path.join(os.path.dirname(__file__), "golden", "chgnet_0.3.0_weights.npz"))
        graphs = graphgen.random_graphs(5, 6, 10, 9600)
        seen = []

        def fn(gs):
            seen.extend(g.graph_id for g in gs)
            return orc.predict_graph(w, gs, "ef", batch_size=max(1, len(gs))) if gs else []

        res = predict_sharded(graphs, fn)
        if rank == 0:
            q.put((seen, [r["e"] for r in res], [r["f"] for r in res]))
        else:
            q.put((seen, None, None))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_prediction_matches_single_process(weights030):
    from oracle import chgnet_oracle as orc

    graphs = graphgen.random_graphs(5, 6, 10, 9600)
    ref = orc.predict_graph(weights030, graphs, "ef", batch_size=5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=300) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    seen_all = sorted(s for g in got for s in g[0])
    assert seen_all == sorted(g.graph_id for g in graphs)  # every graph evaluated exactly once
    full = [g for g in got if g[1] is not None][0]
    for e, f, r in zip(full[1], full[2], ref):
        assert np.allclose(e, r["e"], atol=1e-5) and np.allclose(f, r["f"], atol=1e-4)
    costs = [graph_cost(g) for g in graphs]
    assert len(partition_graphs(costs, 2)[0]) in (2, 3)
