"""Micro-benchmark of the BatchNorm-backward + backward-aggregation step of one layer on the 1M-face graphs:
unfused (ddmp_bn_bwd_reduce, finalize, ddmp_bn_bwd_apply, ddmp_spmm_gcn) against the tile-fused path
(reduce, finalize, ddmp_spmm_bn_bwd_tile).  Bytes: reads of gX, Y and the write of dH = 3 tensor passes (fused floor)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dual_dmp_b200 import functional as F_, synth
from dual_dmp_b200.graph import GcnGraph
from dual_dmp_b200.util.mesh import Mesh
dev = "cuda:0"
vs, faces = synth.icosphere(int(os.environ.get("N", "224")))
m = Mesh(vs=vs * 100, faces=faces)
F, V = len(faces), len(vs)
fg = GcnGraph(torch.from_numpy(m.f_edges), F, dev, coords=torch.from_numpy(m.fc))
e = torch.from_numpy(m.edges.T.astype("int64")); vg = GcnGraph(torch.cat([e, e[[1, 0]]], dim=1), V, dev, coords=torch.from_numpy(m.vs))
tot = {"unfused": 0.0, "fused": 0.0}
for name, g in (("vertex", vg), ("face", fg)):
    for C in (32, 64, 128, 256, 512):
        Y = torch.randn(g.n, C, device=dev) * 1.5 + 0.3
        gX = torch.randn(g.n, C, device=dev)
        st = F_.bn_stats_finalize(F_.spmm_gcn(g, Y, stats=True)[1], g.n, torch.ones(C, device=dev), torch.zeros(C, device=dev))
        dH = torch.empty_like(Y); dY = torch.empty_like(Y)
        def unfused():
            F_.bn_lrelu_backward(gX, Y, st, dY_out=dY)
            F_.spmm_gcn(g, dY, transposed=True, out=dH, amax=C >= 64)
        def fused():
            F_.bn_bwd_spmm_tile(g, gX, Y, st, dH_out=dH, amax=C >= 64)
        for label, fn in (("unfused", unfused), ("fused", fused)):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            w = 2 if C == 256 else 1
            tot[label] += ms * w
            print(f"{name:6s} C={C:3d} {label:8s} {ms:7.3f} ms   3-pass floor {3 * g.n * C * 4 / ms / 1e6:7.0f} GB/s", flush=True)
        del Y, gX, dH, dY
print("per-step totals (ms):", tot)
