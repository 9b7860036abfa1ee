"""Micro-benchmark of the FFMA kernel at the narrow layers of the 1M-face networks (widths < 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dual_dmp_b200 import functional as F_

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1003520
for cin, cout in [(16, 32), (32, 64), (64, 32), (32, 16)]:
    X = torch.randn(n, cin, device=dev); W = torch.randn(cout, cin, device=dev) / cin ** 0.5
    dH = torch.randn(n, cout, device=dev)
    sc = torch.rand(cin, device=dev) + 0.5; sh = torch.randn(cin, device=dev)
    for name, fn in (("xw", lambda: F_.gemm_xw(X, W, scale=sc, shift=sh, backend=1)),
                     ("dx", lambda: F_.gemm_dx(dH, W, backend=1)),
                     ("dw", lambda: F_.gemm_dw(dH, X, cin, scale=sc, shift=sh, backend=1))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = n * (cin + cout) * 4 / 1e9
        print(f"{name} {cin:3d}->{cout:3d}: {ms:7.3f} ms  {2.0 * n * cin * cout / ms / 1e9:6.1f} TFLOP/s  {gb / ms * 1e3:6.0f} GB/s", flush=True)
