"""Micro-benchmark of the narrow layers of the 1M-face networks (a width < 64): FFMA kernel vs the default dispatch
(tensor-core kernels where the reduction width / masked tile rows allow it)."""
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
    for name, mk in (("xw", lambda b: (lambda: F_.gemm_xw(X, W, scale=sc, shift=sh, backend=b))),
                     ("dx", lambda b: (lambda: F_.gemm_dx(dH, W, backend=b))),
                     ("dw", lambda b: (lambda: F_.gemm_dw(dH, X, cin, scale=sc, shift=sh, backend=b)))):
        line = f"{name} {cin:3d}->{cout:3d}:"
        for label, b in (("ffma", 1), ("auto", 0)):
            fn = mk(b)
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            gb = n * (cin + cout) * 4 / 1e9
            line += f"  {label} {ms:7.3f} ms {2.0 * n * cin * cout / ms / 1e9:6.1f} TFLOP/s {gb / ms * 1e3:6.0f} GB/s |"
        print(line, flush=True)
