"""Breakdown of build_batch on the GPU box: C packer, H2D, device CSR, the rest (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chgnet_b200 import graphgen
from chgnet_b200 import batch as B

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
if wl == "c3":
    graphs = graphgen.random_graphs(256, 20, 40, 2000)
else:
    z, frac, lat = graphgen.limno2_structure((10, 5, 25), 0.02, 4000)
    graphs = [graphgen.make_crystal_graph(z, frac, lat)]
dev = torch.device("cuda")
for _ in range(3):
    B.build_batch(graphs, dev)
torch.cuda.synchronize()
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for _ in range(10):
    b = B.build_batch(graphs, dev)
    torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
pr.disable()
print(f"{wl}: build_batch {dt*1e3:.3f} ms per call (synchronised)")
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
