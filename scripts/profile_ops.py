"""Runs each dominant kernel of the step once or twice at the 1M-face shapes (for `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dual_dmp_b200 import functional as F_, synth
from dual_dmp_b200.graph import GcnGraph
from dual_dmp_b200.util.mesh import Mesh

dev = "cuda:0"
n_freq = int(sys.argv[1]) if len(sys.argv) > 1 else 224
vs, faces = synth.icosphere(n_freq)
m = Mesh(vs=vs * 100, faces=faces)
F = len(faces); V = len(vs)
fg = GcnGraph(torch.from_numpy(m.f_edges), F, dev, coords=torch.from_numpy(m.fc))
e = torch.from_numpy(m.edges.T.astype("int64")); ei = torch.cat([e, e[[1, 0]]], dim=1)
vg = GcnGraph(ei, V, dev, coords=torch.from_numpy(m.vs))
torch.manual_seed(0)
for C in (512, 64):
    H = torch.randn(F, C, device=dev); b = torch.randn(C, device=dev)
    for _ in range(2):
        Y, p = F_.spmm_gcn(fg, H, bias=b, stats=True)          # forward flavour
        F_.spmm_gcn(fg, H)                                     # backward flavour
    Hv = torch.randn(V, C, device=dev)
    F_.spmm_gcn(vg, Hv, bias=b, stats=True)
    del H, Hv, Y
X = torch.randn(F, 512, device=dev); W = torch.randn(512, 512, device=dev) / 22
sc = torch.rand(512, device=dev) + 0.5; sh = torch.randn(512, device=dev)
dH = torch.randn(F, 512, device=dev)
for _ in range(2):
    F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2)
    F_.gemm_dx(dH, W, backend=2)
    F_.gemm_dw(dH, X, 512, scale=sc, shift=sh, backend=2)
st = torch.stack([sh, sc, sc, sh])
F_.bn_lrelu_backward(dH, X, st)
X2 = torch.randn(F, 128, device=dev); W2 = torch.randn(256, 128, device=dev) / 11
F_.gemm_xw(X2, W2, backend=2)
torch.cuda.synchronize()
print("done")
