"""debug: per-parameter gradient error of the product vs the float64 oracle on the product's active sets"""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import dataset64, oracle_like, product_masks, rel_err, small_case
from dual_dmp_b200 import functional as F_
from dual_dmp_b200.util import loss as L
from dual_dmp_b200.util.datamaker import dataset_from_meshes
from dual_dmp_b200.util.networks import NormalNet, PosNet
from oracle import step_ref
from oracle.networks_ref import NormalNetRef, PosNetRef

n = int(sys.argv[1]); backend = int(sys.argv[2]); which = sys.argv[3] if len(sys.argv) > 3 else "step"
F_.GEMM_BACKEND = backend
DEV = "cuda:0"
K = (3.0, 4.0, 4.0, 4.0, 1.0)
n_mesh, s_mesh, _ = small_case("ico", n)
ds = dataset_from_meshes(n_mesh, s_mesh)
torch.manual_seed(1)
pa, na = PosNetRef(), NormalNetRef()
pd, nd = PosNet(DEV).to(DEV), NormalNet(DEV).to(DEV)
pd.load_state_dict(pa.state_dict()); nd.load_state_dict(na.state_dict())
pd.train(); nd.train(); pd.taps, nd.taps = [], []
if which == "step":
    pos = pd(ds); nrm = nd(ds)
    l4, _ = L.fn_bnf_loss(pos, nrm, n_mesh, loop=1)
    parts = [L.pos_rec_loss(pos, n_mesh.vs), L.mesh_laplacian_loss(pos, n_mesh), L.norm_rec_loss(nrm, n_mesh.fn), l4,
             L.pos_norm_loss(pos, nrm, n_mesh)]
    sum(k * l for k, l in zip(K, parts)).backward()
else:   # random output gradient, networks only
    pos = pd(ds); nrm = nd(ds)
    gp = torch.randn(pos.shape, generator=torch.Generator().manual_seed(5))
    gn = torch.randn(nrm.shape, generator=torch.Generator().manual_seed(6))
    pos.backward(gp.to(DEV)); nrm.backward(gn.to(DEV))
masks_p, masks_n = product_masks(pd), product_masks(nd)
ds64 = dataset64(ds)
pb, nb = oracle_like(pa, masks_p, double=True), oracle_like(na, masks_n, double=True)
pb.train(); nb.train()
if which == "step":
    tot_b, _, pos_b, nrm_b = step_ref.losses(pb, nb, ds64, n_mesh, K, 1, epoch=101)
    tot_b.backward()
else:
    pos_b = pb(ds64); nrm_b = nb(ds64)
    pos_b.backward(gp.double()); nrm_b.backward(gn.double())
print("n", n, "backend", backend, which, "out err", rel_err(pos, pos_b), rel_err(nrm, nrm_b))
for tag, net_d, net_b in (("pos", pd, pb), ("nrm", nd, nb)):
    for (name, a), (_, b) in zip(net_d.named_parameters(), net_b.named_parameters()):
        if name.startswith("conv") and name.endswith(".bias"):
            continue
        e = rel_err(a.grad, b.grad)
        if e > 3e-5:
            print(f"  {tag}.{name:20s} {e:.2e}  max|g| {float(b.grad.abs().max()):.3e}")
