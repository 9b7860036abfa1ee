"""One fp16-split launch of each dense transform at the 1M-face 512->512 shape (for an `ncu --set full` capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dual_dmp_b200 import functional as F_
dev = "cuda:0"
n, cin, cout = 1003520, 512, 512
X = torch.randn(n, cin, device=dev); W = torch.randn(cout, cin, device=dev) / cin ** 0.5
dH = torch.randn(n, cout, device=dev)
sc = torch.rand(cin, device=dev) + 0.5; sh = torch.randn(cin, device=dev)
bx = (torch.nn.functional.leaky_relu(X[:65536] * sc + sh, 0.01).abs().amax(0) * 1.5).contiguous()
bd = dH[:65536].abs().amax(0).mul(1.5).contiguous()
for _ in range(2):
    F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2, amax=bx)
    F_.gemm_dx(dH, W, backend=2, amax=bd)
    F_.gemm_dw(dH, X, cin, scale=sc, shift=sh, backend=2)
    F_.gemm_dw(dH[:, :128].contiguous(), X[:, :128].contiguous(), 128, scale=sc[:128].contiguous(), shift=sh[:128].contiguous(),
               backend=2, amax_dh=bd, amax_x=bx)
torch.cuda.synchronize()
