"""Runs each dominant kernel of the round-2 step once or twice at the 1M-face shapes (for `ncu --set full`):
aggregation (tile-staged C<=128, gather C>=256; vertex and face graph), BatchNorm backward reduce / apply, the fp16-split
tcgen05 GEMMs (NT pair kernel: X.W^T, dH.W; TN pair kernel: dH^T.X), the fused loss kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dual_dmp_b200 import functional as F_, synth
from dual_dmp_b200.graph import GcnGraph
from dual_dmp_b200.util import loss as L
from dual_dmp_b200.util.mesh import Mesh

dev = "cuda:0"
n_freq = int(sys.argv[1]) if len(sys.argv) > 1 else 224
case = synth.make_case(n_freq)
m = Mesh(vs=case.noise_vs, faces=case.faces)
F = len(case.faces); V = len(case.noise_vs)
fg = GcnGraph(torch.from_numpy(m.f_edges), F, dev, coords=torch.from_numpy(m.fc))
e = torch.from_numpy(m.edges.T.astype("int64")); ei = torch.cat([e, e[[1, 0]]], dim=1)
vg = GcnGraph(ei, V, dev, coords=torch.from_numpy(m.vs))
torch.manual_seed(0)
for C in (32, 128, 512):
    b = torch.randn(C, device=dev)
    for g, rows in ((fg, F), (vg, V)):
        H = torch.randn(rows, C, device=dev)
        for _ in range(2):
            F_.spmm_gcn(g, H, bias=b, stats=True)              # forward flavour
            F_.spmm_gcn(g, H, amax=C >= 64)                    # backward flavour
        del H
X = torch.randn(F, 512, device=dev); W = torch.randn(512, 512, device=dev) / 22
sc = torch.rand(512, device=dev) + 0.5; sh = torch.randn(512, device=dev)
dH = torch.randn(F, 512, device=dev)
bx = (torch.nn.functional.leaky_relu(X[:65536] * sc + sh, 0.01).abs().amax(0) * 1.5).contiguous()
bd = dH[:65536].abs().amax(0).mul(1.5).contiguous()
for _ in range(2):
    F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2, amax=bx)
    F_.gemm_dx(dH, W, backend=2, amax=bd)
    F_.gemm_dw(dH, X, 512, scale=sc, shift=sh, backend=2, amax_dh=bd, amax_x=bx)
st = torch.stack([sh, sc, sc, sh])
for _ in range(2):
    F_.bn_lrelu_backward(dH, X, st)
del X, dH
pos = torch.from_numpy(case.smooth_vs).float().to(dev).requires_grad_(True)
nrm = torch.nn.functional.normalize(torch.from_numpy(m.fn).float() + 0.1 * torch.randn(F, 3), dim=1).to(dev).requires_grad_(True)
for _ in range(2):
    tot, _ = L.dual_loss(pos, nrm, m, m.vs, m.fn, (3.0, 4.0, 4.0, 4.0, 1.0), 1, 1.0)
torch.cuda.synchronize()
print("done")
