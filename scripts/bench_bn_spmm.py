"""A/B: BatchNorm-apply + backward SpMM as two kernels vs the fused ddmp_spmm_bn_bwd (1M-face graphs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dual_dmp_b200 import functional as F_, synth
from dual_dmp_b200.graph import GcnGraph
from dual_dmp_b200.util.mesh import Mesh
dev = "cuda:0"
vs, faces = synth.icosphere(224)
m = Mesh(vs=vs * 100, faces=faces)
F, V = len(faces), len(vs)
fg = GcnGraph(torch.from_numpy(m.f_edges), F, dev, coords=torch.from_numpy(m.fc))
e = torch.from_numpy(m.edges.T.astype("int64")); vg = GcnGraph(torch.cat([e, e[[1, 0]]], dim=1), V, dev, coords=torch.from_numpy(m.vs))
def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
for name, g in (("vertex", vg), ("face", fg)):
    for C in (64, 128, 256, 512):
        Y = torch.randn(g.n, C, device=dev); gX = torch.randn(g.n, C, device=dev)
        st = torch.stack([torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev)])
        dY = torch.empty_like(Y); dH = torch.empty_like(Y)
        def two():
            F_.bn_lrelu_backward(gX, Y, st, dY_out=dY); F_.spmm_gcn(g, dY, transposed=True, out=dH)
        def fused():
            F_.bn_bwd_spmm_fused(g, gX, Y, st, dH_out=dH)
        print(f"{name:6s} C={C:3d} two-kernel {timeit(two):7.3f} ms   fused {timeit(fused):7.3f} ms", flush=True)
        del Y, gX, dY, dH
