"""Structures -> graphs on the host: serial converter loop vs GraphConverter.convert_many (thread pool over the native
builder), then predict_structure(list) end to end.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chgnet_b200.model import CHGNet, GraphConverter

rng = np.random.default_rng(0)
structs = []
for i in range(256):
    n = int(rng.integers(20, 41)); a = (n / 0.1) ** (1 / 3)
    frac = rng.random((n, 3))
    structs.append((rng.integers(1, 90, n), frac, np.eye(3) * a + rng.normal(0, 0.05, (3, 3))))
gc = GraphConverter(on_isolated_atoms="ignore")
gc.convert_many(structs[:16])
t0 = time.perf_counter(); a = [gc(s) for s in structs]; t1 = time.perf_counter()
b = gc.convert_many(structs); t2 = time.perf_counter()
print("256 structures (20-40 atoms): serial loop %.1f ms, convert_many %.1f ms (%d host threads)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, os.cpu_count()))
assert all(torch.equal(x.atom_graph, y.atom_graph) and torch.equal(x.bond_graph, y.bond_graph) for x, y in zip(a, b))
if torch.cuda.is_available():
    m = CHGNet.from_file("tests/golden/chgnet_0.3.0_weights.npz", version="0.3.0").to("cuda")
    m.graph_converter.on_isolated_atoms = "ignore"
    for mode in ("native", "python"):
        os.environ["CHGNET_B200_GRAPH"] = mode
        for _ in range(2):
            m.predict_structure(structs, task="efs", batch_size=256)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            out = m.predict_structure(structs, task="efs", batch_size=256)
        torch.cuda.synchronize(); print("predict_structure(list of 256, batch_size=256) [%s graph path]: %.1f ms per call" % (mode, (time.perf_counter() - t0) / 5 * 1e3))
    t0 = time.perf_counter()
    for _ in range(5):
        b = m.structures_to_batch(structs)
    torch.cuda.synchronize(); print("structures_to_batch alone: %.1f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
    graphs = gc.convert_many(structs)
    m.predict_graph(graphs, task="efs", batch_size=256)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        m.predict_graph(graphs, task="efs", batch_size=256)
    torch.cuda.synchronize(); print("predict_graph(the same 256 as CrystalGraphs): %.1f ms per call" % ((time.perf_counter() - t0) / 5 * 1e3))
