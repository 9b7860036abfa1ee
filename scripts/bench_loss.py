"""Micro-benchmark of the fused loss kernel (ddmp_dual_loss) on the 1M-face benchmark mesh: event time with a cold L2
(a 256 MB buffer is rewritten before every launch, as inside the step) and the phase trace of the last launch
(ddmp_dual_loss_trace: %globaltimer stamps taken by block 0 at the phase boundaries).
usage: bench_loss.py [n_freq=224] [bnfloop=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dual_dmp_b200 import synth
from dual_dmp_b200._lib import lib
from dual_dmp_b200.util import loss as L
from dual_dmp_b200.util.mesh import Mesh

dev = "cuda:0"
n_freq = int(sys.argv[1]) if len(sys.argv) > 1 else 224
loop = int(sys.argv[2]) if len(sys.argv) > 2 else 1
case = synth.make_case(n_freq)
m = Mesh(vs=case.noise_vs, faces=case.faces)
F, V = len(case.faces), len(case.noise_vs)
torch.manual_seed(0)
pos = torch.from_numpy(case.smooth_vs).float().to(dev).requires_grad_(True)
nrm = torch.nn.functional.normalize(torch.from_numpy(m.fn).float() + 0.1 * torch.randn(F, 3), dim=1).to(dev).requires_grad_(True)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
names = ["P1 vertices (pos_rec, Laplacian residual)", "P1 faces (geometry, norm_rec, pos_norm)", "publish x4 + barrier",
         "totals x4", "P2 vertices (dpos: Laplacian' + corner gather)", "P2 faces (centroid distances)", "publish + barrier",
         "total + filter iteration(s)", "publish + filter backward: messages", "barrier", "gather (reverse slots)",
         "final total"]
ms = []
tr = np.zeros(16, dtype=np.uint64)
acc = np.zeros(12)
for it in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tot, parts = L.dual_loss(pos, nrm, m, m.vs, m.fn, (3.0, 4.0, 4.0, 4.0, 1.0), loop, 1.0)
    e1.record()
    torch.cuda.synchronize()
    if it >= 2:
        ms.append(e0.elapsed_time(e1))
        lib.call("ddmp_dual_loss_trace", tr.ctypes.data)
        acc += np.diff(tr[:13].astype(np.int64)) / 1e3
byt = 349225272 if (n_freq == 224 and loop == 1) else None
kern_us = float(acc.sum() / len(ms))
print(f"faces {F} loop {loop}: kernel {kern_us:.1f} us (sum of the phase trace; the event time {np.mean(ms) * 1e3:.0f} us includes the "
      f"host side of the autograd call)" + (f", {byt / kern_us / 1e3:.0f} GB/s algorithmic" if byt else ""), "total", float(tot.detach()))
for nme, v in zip(names, acc / len(ms)):
    print(f"  {nme:48s} {v:8.1f} us")
