"""debug: every tensor-core GEMM call of one NormalNet/PosNet step at n=64, compared with float64 on the SAME inputs"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import small_case
from dual_dmp_b200 import functional as F_
from dual_dmp_b200.util import loss as L
from dual_dmp_b200.util.datamaker import dataset_from_meshes
from dual_dmp_b200.util.networks import NormalNet, PosNet
from oracle.networks_ref import NormalNetRef, PosNetRef

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
DEV = "cuda:0"
K = (3.0, 4.0, 4.0, 4.0, 1.0)
n_mesh, s_mesh, _ = small_case("ico", n)
ds = dataset_from_meshes(n_mesh, s_mesh)
torch.manual_seed(1)
pa, na = PosNetRef(), NormalNetRef()
pd, nd = PosNet(DEV).to(DEV), NormalNet(DEV).to(DEV)
pd.load_state_dict(pa.state_dict()); nd.load_state_dict(na.state_dict())
pd.train(); nd.train()
orig_dx, orig_dw = F_.gemm_dx, F_.gemm_dw
tag = [""]

def stats(t):
    a = t.abs().flatten()
    rows = t.abs().amax(dim=1)
    return "max %.2e med %.2e rowmax[min %.2e med %.2e]" % (float(a.max()), float(a.median()), float(rows.min()), float(rows.median()))

def dx(dH, W, out=None, backend=None, amax=None):
    res = orig_dx(dH, W, out=out, backend=backend, amax=amax)
    ref = dH.double() @ W.double()
    ff = orig_dx(dH, W, backend=1)
    e_tc = float((res.double() - ref).abs().max() / ref.abs().max())
    e_ff = float((ff.double() - ref).abs().max() / ref.abs().max())
    # per-row relative error (rows with small magnitude)
    rr = ((res.double() - ref).abs().amax(dim=1) / (ref.abs().amax(dim=1) + 1e-300))
    print(f"{tag[0]} dx {tuple(W.shape)} amax={'y' if amax is not None else 'n'} bound={float(amax.max()) if amax is not None else 0:.2e} "
          f"err tc {e_tc:.1e} ffma {e_ff:.1e} worst-row-rel {float(rr.max()):.1e} med-row-rel {float(rr.median()):.1e} | dH {stats(dH)}")
    return res

def dw(dH, X, Cin, row_map=None, scale=None, shift=None, backend=None, amax_dh=None, amax_x=None):
    res = orig_dw(dH, X, Cin, row_map=row_map, scale=scale, shift=shift, backend=backend, amax_dh=amax_dh, amax_x=amax_x)
    Xa = X.double() if row_map is None else X.double()[row_map.long()]
    if scale is not None:
        Xa = torch.nn.functional.leaky_relu(Xa * scale.double() + shift.double(), 0.01)
    ref = dH.double().t() @ Xa[:, :Cin]
    ff = orig_dw(dH, X, Cin, row_map=row_map, scale=scale, shift=shift, backend=1)
    e_tc = float((res.double() - ref).abs().max() / ref.abs().max())
    e_ff = float((ff.double() - ref).abs().max() / ref.abs().max())
    print(f"{tag[0]} dw {tuple(res.shape)} amax={'y' if amax_dh is not None else 'n'} err tc {e_tc:.1e} ffma {e_ff:.1e} "
          f"cancel {float((dH.double().abs().t() @ Xa[:, :Cin].abs()).max() / ref.abs().max()):.1e}")
    return res

F_.gemm_dx, F_.gemm_dw = dx, dw
pos = pd(ds); nrm = nd(ds)
l4, _ = L.fn_bnf_loss(pos, nrm, n_mesh, loop=1)
parts = [L.pos_rec_loss(pos, n_mesh.vs), L.mesh_laplacian_loss(pos, n_mesh), L.norm_rec_loss(nrm, n_mesh.fn), l4,
         L.pos_norm_loss(pos, nrm, n_mesh)]
tot = sum(k * l for k, l in zip(K, parts))
gp, gn = torch.autograd.grad(tot, [pos, nrm], retain_graph=True)
tag[0] = "NRM"
nrm.backward(gn, retain_graph=True)
tag[0] = "POS"
pos.backward(gp)
