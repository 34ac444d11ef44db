import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chgnet_b200._lib import CudaKernels
K = CudaKernels()
m, k, n = 422077, 64, 128
x = torch.randn(m, k, device="cuda"); wt = torch.randn(k, n, device="cuda") / 8; bias = torch.randn(n, device="cuda")
y = torch.empty(m, n, device="cuda")
for _ in range(3):
    K.linear(x, wt, bias, None, y, None, None)
torch.cuda.synchronize()
