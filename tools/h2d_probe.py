"""Where the 2.8 ms device-side wait of build_batch goes: raw pinned H2D bandwidth, and the GPU time span of one
build_batch (CUDA events), for c3 / c4.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chgnet_b200 import graphgen
from chgnet_b200 import batch as B

dev = torch.device("cuda")
if os.environ.get("PROBE_BIND") == "1":
    import bench
    print("cpu binding:", bench.bind_to_gpu_numa_node(0)[1])
for mb in (1, 8, 35, 128):
    h = torch.empty(mb * 2**20 // 4, dtype=torch.int32).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D pinned {mb} MiB: {ms:.3f} ms  ({mb * 2**20 / ms / 1e6:.1f} GB/s)")
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"D2H pinned {mb} MiB: {ms:.3f} ms  ({mb * 2**20 / ms / 1e6:.1f} GB/s)")

for wl in ("c3", "c4"):
    if wl == "c3":
        graphs = graphgen.random_graphs(256, 20, 40, 2000)
    else:
        z, frac, lat = graphgen.limno2_structure((10, 5, 25), 0.02, 4000)
        graphs = [graphgen.make_crystal_graph(z, frac, lat)]
    for _ in range(3):
        B.build_batch(graphs, dev)
    torch.cuda.synchronize()
    span, wall, host = [], [], []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        b = B.build_batch(graphs, dev)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        span.append(e0.elapsed_time(e1)); wall.append((t2 - t0) * 1e3); host.append((t1 - t0) * 1e3)
    span.sort(); wall.sort(); host.sort()
    print(f"{wl}: build_batch wall {wall[5]:.3f} ms, host part {host[5]:.3f} ms, first-to-last GPU event span {span[5]:.3f} ms, h2d bytes {b.h2d_bytes if hasattr(b, 'h2d_bytes') else '?'}")

# write-combined vs default pinned memory (cudaHostAlloc through cuda-python), 19 MiB and 35 MiB
try:
    from cuda import cudart
    for mb in (19, 35):
        nbytes = mb * 2**20
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        for label, flag in (("default", cudart.cudaHostAllocDefault), ("write-combined", cudart.cudaHostAllocWriteCombined)):
            err, hp = cudart.cudaHostAlloc(nbytes, flag)
            assert int(err) == 0, err
            for _ in range(3):
                cudart.cudaMemcpyAsync(d.data_ptr(), hp, nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, 0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                cudart.cudaMemcpyAsync(d.data_ptr(), hp, nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, 0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"H2D cudaHostAlloc({label}) {mb} MiB: {ms:.3f} ms ({nbytes / ms / 1e6:.1f} GB/s)")
            cudart.cudaFreeHost(hp)
except Exception as exc:  # noqa: BLE001
    print("cuda-python probe unavailable:", repr(exc))
