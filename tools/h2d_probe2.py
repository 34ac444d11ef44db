"""H2D of FRESHLY CPU-WRITTEN pinned memory (what a staging buffer is): cudaHostAlloc default vs write-combined vs a
torch pinned tensor; single-threaded and 8-thread writers.  Run under gpurun."""
import ctypes, os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda import cudart

dev = torch.device("cuda")
torch.zeros(1, device=dev)
nbytes = 19 * 2**20
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def write(ptr, val, threads):
    if threads == 1:
        ctypes.memset(ptr, val, nbytes)
        return
    chunk = nbytes // threads
    ts = [threading.Thread(target=ctypes.memset, args=(ptr + k * chunk, val, chunk)) for k in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]


def run(label, ptr, threads, written=True):
    tot = []
    for i in range(8):
        if written:
            write(ptr, i + 1, threads)
        e0.record()
        cudart.cudaMemcpyAsync(d.data_ptr(), ptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, 0)
        e1.record()
        torch.cuda.synchronize()
        tot.append(e0.elapsed_time(e1))
    tot.sort()
    print(f"{label:42s} writers={threads} written={written}: median {tot[4]:.3f} ms ({nbytes / tot[4] / 1e6:.1f} GB/s)")


for label, flag in (("cudaHostAlloc(default)", cudart.cudaHostAllocDefault), ("cudaHostAlloc(write-combined)", cudart.cudaHostAllocWriteCombined)):
    err, hp = cudart.cudaHostAlloc(nbytes, flag)
    assert int(err) == 0
    run(label, hp, 1, written=False)
    run(label, hp, 1)
    run(label, hp, 8)
    cudart.cudaFreeHost(hp)
t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
run("torch pin_memory()", t.data_ptr(), 1, written=False)
run("torch pin_memory()", t.data_ptr(), 1)
run("torch pin_memory()", t.data_ptr(), 8)
t2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
run("torch.empty(pin_memory=True)", t2.data_ptr(), 1, written=False)
run("torch.empty(pin_memory=True)", t2.data_ptr(), 1)
