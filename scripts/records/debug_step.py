import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import small_case, rel_err, oracle_like, product_masks
from tests.test_gpu_nets import _pair
from dual_dmp_b200.util.datamaker import dataset_from_meshes
from dual_dmp_b200.util import loss as L
from dual_dmp_b200 import functional as F_
from oracle import step_ref
if len(sys.argv) > 1: F_.GEMM_BACKEND = int(sys.argv[1])
n_mesh, s_mesh, _ = small_case("ico", 12)
ds = dataset_from_meshes(n_mesh, s_mesh)
pr, nr, pd, nd = _pair(1)
k = (3.0, 4.0, 4.0, 4.0, 1.0); loop = 1
pr.train(); nr.train()
tp, tn = [], []
# plain oracle with taps
pos_r = pr(ds, tp)
from oracle import loss_ref as R
nrm_r = nr(ds, tn)
l = [R.pos_rec_loss(pos_r, n_mesh.vs), R.mesh_laplacian_loss(pos_r, n_mesh), R.norm_rec_loss(nrm_r, n_mesh.fn)]
l4, _ = R.fn_bnf_loss(pos_r, nrm_r, n_mesh, loop=loop); l.append(l4); l.append(R.pos_norm_loss(pos_r, nrm_r, n_mesh))
sum(a * b for a, b in zip(k, l)).backward()
pd.train(); nd.train(); pd.taps, nd.taps = [], []
pos = pd(ds); nrm = nd(ds)
ld = [L.pos_rec_loss(pos, n_mesh.vs), L.mesh_laplacian_loss(pos, n_mesh), L.norm_rec_loss(nrm, n_mesh.fn)]
l4d, _ = L.fn_bnf_loss(pos, nrm, n_mesh, loop=loop); ld.append(l4d); ld.append(L.pos_norm_loss(pos, nrm, n_mesh))
sum(a * b for a, b in zip(k, ld)).backward()
print("losses", [x.item() for x in ld], [x.item() for x in l])
for name, net_d, net_r, taps in (("pos", pd, pr, tp), ("norm", nd, nr, tn)):
    masks = product_masks(net_d)
    for li, (m, (y_r, x_r)) in enumerate(zip(masks, taps)):
        mism = int((m != (x_r > 0)).sum())
        if mism: print(name, "layer", li + 1, "mask mismatches", mism)
pm, nm = oracle_like(pr, product_masks(pd)), oracle_like(nr, product_masks(nd))
pm.train(); nm.train()
tot, _, pos_m, nrm_m = step_ref.losses(pm, nm, ds, n_mesh, k, loop, epoch=101)
tot.backward()
print("out diffs: pos dev-plain", rel_err(pos, pos_r), "nrm dev-plain", rel_err(nrm, nrm_r), "nrm masked-plain", rel_err(nrm_m, nrm_r))
# where do L1 sign kinks differ?
fn_t = torch.from_numpy(n_mesh.fn)
sd = torch.sign(nrm.detach().cpu().double() - fn_t); sr = torch.sign(nrm_r.detach().double() - fn_t)
print("norm_rec sign mismatches dev-vs-plain", int((sd != sr).sum()), "min |nrm-fn|", float((nrm_r.detach().double() - fn_t).abs().min()))
for (name, a), (_, b), (_, c) in zip(nd.named_parameters(), nr.named_parameters(), nm.named_parameters()):
    if name.endswith("bias") and name.startswith("conv"): continue
    print(f"  {name:20s} dev-vs-plain {rel_err(a.grad, b.grad):.2e}  dev-vs-masked {rel_err(a.grad, c.grad):.2e}  masked-vs-plain {rel_err(c.grad, b.grad):.2e}")
