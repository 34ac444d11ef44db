"""Where does the host-visible time of build_batch go?  (torch.profiler, run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from chgnet_b200 import graphgen
from chgnet_b200.batch import build_batch

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
if wl == "c2":
    graphs = graphgen.random_graphs(64, 40, 60, 1000)
elif wl == "c3":
    graphs = graphgen.random_graphs(256, 20, 40, 2000)
else:
    z, frac, lat = graphgen.limno2_structure((10, 5, 25), 0.02, 4000)
    graphs = [graphgen.make_crystal_graph(z, frac, lat)]
dev = torch.device("cuda")
for _ in range(3):
    build_batch(graphs, dev); torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); build_batch(graphs, dev); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(wl, "build_batch ms:", [round(t, 2) for t in ts])
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    build_batch(graphs, dev); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=50))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=50))
