import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import oracle_like, product_masks, rel_err
from tests.test_gpu_nets import _pair
from dual_dmp_b200.util import loss as L
from dual_dmp_b200.util.datamaker import dataset_from_meshes
from dual_dmp_b200.util.mesh import Mesh
from oracle import step_ref
for name in ("tetra", "strip2"):
    g = dict(np.load(f"tests/golden/{name}.npz"))
    rng = np.random.RandomState(1)
    n_mesh = Mesh(vs=g["vs"] + 0.05 * rng.randn(*g["vs"].shape), faces=g["faces"])
    s_mesh = Mesh(vs=g["vs"], faces=g["faces"])
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, nr, pd, nd = _pair(7)
    pr.train(); nr.train(); pd.train(); nd.train()
    pd.taps, nd.taps = [], []
    pos = pd(ds); nrm = nd(ds)
    ls = [L.pos_rec_loss(pos, n_mesh.vs), L.mesh_laplacian_loss(pos, n_mesh), L.norm_rec_loss(nrm, n_mesh.fn)]
    l4, nf = L.fn_bnf_loss(pos, nrm, n_mesh, loop=3); ls += [l4, L.pos_norm_loss(pos, nrm, n_mesh)]
    tot, parts, pos_r, nrm_r = step_ref.losses(pr, nr, ds, n_mesh, (3.0, 4.0, 4.0, 4.0, 1.0), 3, epoch=101)
    print(name, "dev", [float(x) for x in ls]); print(name, "ref", [float(x) for x in parts])
    print("pos err", rel_err(pos, pos_r), "nrm err", rel_err(nrm, nrm_r), "f2f", n_mesh.f2f.tolist())
    from oracle import loss_ref as R
    # losses on identical inputs
    pc, nc = pos.detach().cpu(), nrm.detach().cpu()
    print(name, "ref-on-dev-outputs", float(R.mesh_laplacian_loss(pc, n_mesh)), [float(x) for x in R.fn_bnf_loss(pc, nc, n_mesh, loop=3)[:1]], float(R.pos_norm_loss(pc, nc, n_mesh)))
