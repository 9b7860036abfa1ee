"""Micro-benchmark of the GCN aggregation kernel on the 1M-face graphs (both flavours, every width)."""
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
tot_b = tot_t = 0.0
for name, g in (("vertex", vg), ("face", fg)):
    for C in (32, 64, 128, 256, 512):
        H = torch.randn(g.n, C, device=dev); b = torch.randn(C, device=dev)
        byt = 4 * ((g.n + 1) + 2 * g.nnz + 2 * g.n * C)
        for flav, fn in (("fwd(stats+bias)", lambda: F_.spmm_gcn(g, H, bias=b, stats=True)), ("bwd(plain)", lambda: F_.spmm_gcn(g, H))):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            w = 2 if C == 256 else 1          # a step has twice as many 256-wide launches
            tot_b += byt * w; tot_t += ms * w
            print(f"{name:6s} C={C:3d} {flav:16s} {ms:7.3f} ms {byt / ms / 1e6:7.0f} GB/s", flush=True)
        del H
print(f"step-weighted: {tot_b / tot_t / 1e6:.0f} GB/s")
