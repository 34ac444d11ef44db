"""A/B microbenchmark + accuracy of chg_linear: tcgen05 3xTF32 (default) vs fp64 truth.
Run once per implementation: CHG_LINEAR_IMPL=ffma python tools/linear_ab.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chgnet_b200._lib import CudaKernels

K = CudaKernels()
torch.manual_seed(0)
flush = torch.empty(64 << 20, device="cuda")
for impl_id, impl in ((3, "ws"), (1, "tc"), (0, "ffma")):
  K.set_option("linear_impl", impl_id)
  for (m, k, n, gather) in [(300, 64, 128, False), (422077, 64, 128, False), (52000, 64, 256, True), (10000, 64, 256, False),
                            (422077, 128, 64, False), (52000, 256, 64, True), (10000, 256, 64, False), (422077, 64, 64, False)]:
      src_rows = m * 8 if gather else m
      x = torch.randn(src_rows, k, device="cuda")
      wt = torch.randn(k, n, device="cuda") / k ** 0.5
      bias = torch.randn(n, device="cuda")
      rows = torch.sort(torch.randperm(src_rows, device="cuda")[:m]).values.int() if gather else None
      res = torch.randn(src_rows if gather else m, n, device="cuda")
      y = res.clone() if gather else torch.empty(m, n, device="cuda")
      K.linear(x, wt, bias, res, y, rows, rows)
      torch.cuda.synchronize()
      xs = x[rows.long()] if gather else x
      want = xs.double() @ wt.double() + bias.double() + (res[rows.long()] if gather else res).double()
      got = (y[rows.long()] if gather else y).double()
      err = (got - want).abs().max().item()
      ts = []
      for _ in range(5):
          flush.zero_()
          s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          s.record(); K.linear(x, wt, bias, res, y, rows, rows); e.record(); e.synchronize()
          ts.append(s.elapsed_time(e))
      us = min(ts) * 1e3
      gb = (m * k + 2 * m * n) * 4 / 1e9
      print(f"{impl:5s} m={m:7d} k={k:3d} n={n:3d} gather={int(gather)}  max_err={err:.2e}  {us:8.1f} us  "
            f"{2*m*k*n/us/1e6:7.1f} TFLOP/s  {gb/us*1e6:7.0f} GB/s")
