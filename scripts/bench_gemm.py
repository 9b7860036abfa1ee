"""Micro-benchmark of the dense feature transform at the 1M-face shapes: FFMA vs tcgen05 3xTF32."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dual_dmp_b200 import functional as F_

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1003520
shapes = [(64, 128), (128, 256), (256, 256), (256, 512), (512, 512), (512, 256), (256, 128), (128, 64)]
res = {}
for cin, cout in shapes:
    X = torch.randn(n, cin, device=dev); W = torch.randn(cout, cin, device=dev) / cin ** 0.5
    dH = torch.randn(n, cout, device=dev)
    sc = torch.rand(cin, device=dev) + 0.5; sh = torch.randn(cin, device=dev)
    bx = (torch.nn.functional.leaky_relu(X[:65536] * sc + sh, 0.01).abs().amax(0) * 1.5).contiguous()
    bd = dH[:65536].abs().amax(0).mul(1.5).contiguous()
    F16 = os.environ.get("BENCH_F16", "1") == "1"
    for name, fn in (("xw", lambda b: F_.gemm_xw(X, W, scale=sc, shift=sh, backend=b, amax=bx if (F16 and b == 2) else None)),
                     ("dx", lambda b: F_.gemm_dx(dH, W, backend=b, amax=bd if (F16 and b == 2) else None)),
                     ("dw", lambda b: F_.gemm_dw(dH, X, cin, scale=sc, shift=sh, backend=b, amax_dh=bd if (F16 and b == 2) else None,
                                                 amax_x=bx if (F16 and b == 2) else None))):
        for b in (1, 2):
            try:
                fn(b); torch.cuda.synchronize()
            except RuntimeError as e:
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): fn(b)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            tf = 2.0 * n * cin * cout / ms / 1e9
            res[f"{name} {cin}->{cout} backend={b}"] = (round(ms, 3), round(tf, 1))
            print(f"{name} {cin:4d}->{cout:4d} backend={b}: {ms:8.3f} ms  {tf:7.1f} TFLOP/s(fp32-equivalent)", flush=True)
    del X, dH
json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)
