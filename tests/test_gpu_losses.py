"""Loss / geometry kernels (through the drop-in util.loss / util.models API) against the REAL reference's golden
vectors (tests/golden, produced by oracle/make_golden.py) and against the oracle on larger meshes."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests.helpers import rel_err, report, small_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["ico3", "ico6", "open4", "tetra", "strip2"]


def _mesh(g):
    return SimpleNamespace(vs=g["vs"], faces=g["faces"], edges=g["edges"], f2f=g["f2f_raw"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("loop", [1, 3, 5])
def test_losses_against_reference_golden(golden_dir, name, loop):
    from dual_dmp_b200.util import loss as L
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    mesh = _mesh(g)
    p = torch.from_numpy(g["pos"]).to(DEV).requires_grad_(True)
    n = torch.from_numpy(g["nrm"]).to(DEV).requires_grad_(True)
    l1 = L.pos_rec_loss(p, g["tgt_vs"])
    l2 = L.mesh_laplacian_loss(p, mesh)
    l3 = L.norm_rec_loss(n, g["fn"])
    l4, new_fn = L.fn_bnf_loss(p, n, mesh, loop=loop)
    l5 = L.pos_norm_loss(p, n, mesh)
    assert [str(x.dtype) for x in (l1, l2, l3, l4, l5)] == g[f"loss_dtypes_loop{loop}"].tolist()
    got = np.array([x.item() for x in (l1, l2, l3, l4, l5)])
    ref = g[f"loss_loop{loop}"]
    assert np.allclose(got, ref, rtol=5e-6, atol=1e-7), (got, ref)
    (3.0 * l1 + 4.0 * l2 + 4.0 * l3 + 4.0 * l4 + 1.0 * l5).backward()
    e_fn = rel_err(new_fn, torch.from_numpy(g[f"bnf_fn_loop{loop}"]))
    e_p = rel_err(p.grad, torch.from_numpy(g[f"gpos_loop{loop}"]))
    e_n = rel_err(n.grad, torch.from_numpy(g[f"gnrm_loop{loop}"]))
    report(f"loss golden {name} loop={loop}", (float(np.abs(got - ref).max()), e_fn, e_p, e_n))
    assert e_fn < 1e-5 and e_p < 5e-5 and e_n < 5e-5, (e_fn, e_p, e_n)


@pytest.mark.parametrize("name", CASES)
def test_single_loss_gradients_golden(golden_dir, name):
    from dual_dmp_b200.util import loss as L
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    mesh = _mesh(g)
    fns = {"pos_rec": lambda p, n: L.pos_rec_loss(p, g["tgt_vs"]),
           "lap": lambda p, n: L.mesh_laplacian_loss(p, mesh),
           "norm_rec": lambda p, n: L.norm_rec_loss(n, g["fn"]),
           "pos_norm": lambda p, n: L.pos_norm_loss(p, n, mesh)}
    for key, f in fns.items():
        p = torch.from_numpy(g["pos"]).to(DEV).requires_grad_(True)
        n = torch.from_numpy(g["nrm"]).to(DEV).requires_grad_(True)
        f(p, n).backward()
        for t, suffix in ((p, "pos"), (n, "nrm")):
            ref = g[f"g_{key}_{suffix}"]
            if ref.size == 0:
                assert t.grad is None
                continue
            e = rel_err(t.grad, torch.from_numpy(ref))
            assert e < 2e-5, (key, suffix, e)


@pytest.mark.parametrize("name", CASES)
def test_mad_and_models_golden(golden_dir, name):
    from dual_dmp_b200.util import loss as L
    from dual_dmp_b200.util import models as M
    from dual_dmp_b200.util.mesh import Mesh
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    n1, n2 = torch.from_numpy(g["nrm"]).to(DEV), torch.from_numpy(g["fn"]).float().to(DEV)
    assert abs(L.mad(n1, n2) - float(g["mad"])) < 2e-3            # float32 normals on the device
    assert abs(L.mad(g["nrm"], g["fn"]) - float(g["mad"])) < 1e-9  # numpy inputs: host expression
    pos = torch.from_numpy(g["pos"]).to(DEV)
    assert rel_err(M.compute_fn(pos, g["faces"]), torch.from_numpy(g["compute_fn"])) < 1e-5
    assert rel_err(M.compute_vn(pos, n2, g["faces"]), torch.from_numpy(g["compute_vn"])) < 1e-5
    if "vertex_updating" in g:
        mesh = Mesh(vs=g["vs"], faces=g["faces"])
        out = M.vertex_updating(pos, n1, mesh, loop=3)
        assert rel_err(out, torch.from_numpy(g["vertex_updating"])) < 1e-5


def test_compute_fn_backward_matches_autograd():
    from dual_dmp_b200.util import models as M
    from oracle import models_ref
    n_mesh, _, _ = small_case("open", 6)
    pos = torch.from_numpy(n_mesh.vs).float()
    gfn = torch.randn(len(n_mesh.faces), 3)
    pr = pos.clone().requires_grad_(True)
    models_ref.compute_fn(pr, n_mesh.faces).backward(gfn)
    pd = pos.to(DEV).requires_grad_(True)
    M.compute_fn(pd, n_mesh.faces).backward(gfn.to(DEV))
    assert rel_err(pd.grad, pr.grad) < 2e-5


@pytest.mark.parametrize("kind,n,loop", [("ico", 24, 1), ("ico", 24, 5), ("open", 20, 5), ("ico", 24, 0)])
def test_losses_against_oracle_medium(kind, n, loop):
    from dual_dmp_b200.util import loss as L
    from oracle import loss_ref as R
    n_mesh, s_mesh, _ = small_case(kind, n)
    torch.manual_seed(1)
    V, F = len(n_mesh.vs), len(n_mesh.faces)
    pos0 = torch.from_numpy(s_mesh.vs).float() + 0.05 * torch.randn(V, 3)
    nrm0 = torch.from_numpy(n_mesh.fn).float() + 0.3 * torch.randn(F, 3)
    nrm0 = nrm0 / nrm0.norm(dim=1, keepdim=True)

    def run(mod, dev):
        p = pos0.clone().to(dev).requires_grad_(True)
        q = nrm0.clone().to(dev).requires_grad_(True)
        ls = [mod.pos_rec_loss(p, n_mesh.vs), mod.mesh_laplacian_loss(p, n_mesh), mod.norm_rec_loss(q, n_mesh.fn)]
        l4, nf = mod.fn_bnf_loss(p, q, n_mesh, loop=loop)
        ls += [l4, mod.pos_norm_loss(p, q, n_mesh)]
        (3.0 * ls[0] + 4.0 * ls[1] + 4.0 * ls[2] + 4.0 * ls[3] + 1.0 * ls[4]).backward()
        return [x.item() for x in ls], p.grad, q.grad, nf

    lr, gpr, gnr, nfr = run(R, "cpu")
    ld, gpd, gnd, nfd = run(L, DEV)
    errs = (max(abs(a - b) / (abs(b) + 1e-12) for a, b in zip(ld, lr)), rel_err(gpd, gpr), rel_err(gnd, gnr),
            rel_err(nfd, nfr))
    report(f"loss oracle {kind}{n} loop={loop}", errs)
    assert errs[0] < 1e-5 and errs[1] < 5e-5 and errs[2] < 5e-5 and errs[3] < 1e-5, errs
    # determinism: bitwise equal on a second run
    ld2, gpd2, gnd2, _ = run(L, DEV)
    assert ld == ld2 and torch.equal(gpd, gpd2) and torch.equal(gnd, gnd2)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("loop", [1, 3, 5])
def test_fused_dual_loss_against_reference_golden(golden_dir, name, loop):
    """ddmp_dual_loss (one cooperative kernel for the five losses, the weighted sum and the backward of it) against the
    REAL reference's golden vectors: loss values, d/dpos and d/dnorm of 3*l1 + 4*l2 + 4*l3 + 4*l4 + 1*l5"""
    from dual_dmp_b200.util import loss as L
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    mesh = _mesh(g)
    p = torch.from_numpy(g["pos"]).to(DEV).requires_grad_(True)
    n = torch.from_numpy(g["nrm"]).to(DEV).requires_grad_(True)
    total, parts = L.dual_loss(p, n, mesh, g["tgt_vs"], g["fn"], (3.0, 4.0, 4.0, 4.0, 1.0), loop, 1.0)
    assert total.dtype == torch.float64 and parts.dtype == torch.float64
    ref = g[f"loss_loop{loop}"]
    got = parts.cpu().numpy()
    assert np.allclose(got, ref, rtol=5e-6, atol=1e-7), (got, ref)
    ref_total = 3.0 * ref[0] + 4.0 * ref[1] + 4.0 * ref[2] + 4.0 * ref[3] + 1.0 * ref[4]
    assert abs(total.item() - ref_total) <= 5e-6 * abs(ref_total)
    total.backward()
    e_p = rel_err(p.grad, torch.from_numpy(g[f"gpos_loop{loop}"]))
    e_n = rel_err(n.grad, torch.from_numpy(g[f"gnrm_loop{loop}"]))
    report(f"fused dual loss golden {name} loop={loop}", (float(np.abs(got - ref).max()), e_p, e_n))
    assert e_p < 5e-5 and e_n < 5e-5, (e_p, e_n)


@pytest.mark.parametrize("kind,n", [("ico", 12), ("open", 10), ("ico", 40)])
@pytest.mark.parametrize("loop,scale,k", [(0, 1.0, (3.0, 4.0, 4.0, 4.0, 1.0)), (1, 1.0, (3.0, 4.0, 4.0, 4.0, 1.0)),
                                          (1, 0.0, (3.0, 4.0, 4.0, 4.0, 1.0)), (5, 1.0, (3.0, 0.0, 3.0, 4.0, 2.0))])
def test_fused_dual_loss_equals_separate_losses(kind, n, loop, scale, k):
    """same arithmetic as the stand-alone kernels: values to rounding of the final sums, gradients to 1e-6; the
    epoch <= 100 switch (bnf term times 0.0) zeroes the filter's gradient but not the others; upstream gradient != 1"""
    from dual_dmp_b200.util import loss as L
    n_mesh, s_mesh, _ = small_case(kind, n)
    torch.manual_seed(n + loop)
    V, F = len(n_mesh.vs), len(n_mesh.faces)
    pos0 = torch.from_numpy(s_mesh.vs).float() + 0.05 * torch.randn(V, 3)
    nrm0 = torch.nn.functional.normalize(torch.from_numpy(n_mesh.fn).float() + 0.3 * torch.randn(F, 3), dim=1)
    res = []
    for fused in (True, False):
        p = pos0.clone().to(DEV).requires_grad_(True)
        q = nrm0.clone().to(DEV).requires_grad_(True)
        if fused:
            total, parts = L.dual_loss(p, q, n_mesh, n_mesh.vs, n_mesh.fn, k, loop, scale)
        else:
            l4, _ = L.fn_bnf_loss(p, q, n_mesh, loop=loop)
            ls = [L.pos_rec_loss(p, n_mesh.vs), L.mesh_laplacian_loss(p, n_mesh), L.norm_rec_loss(q, n_mesh.fn),
                  l4 * scale, L.pos_norm_loss(p, q, n_mesh)]
            total = k[0] * ls[0] + k[1] * ls[1] + k[2] * ls[2] + k[3] * ls[3] + k[4] * ls[4]
            parts = torch.stack([x.double() for x in ls])
        (total * 1.7).backward()
        res.append((total.detach(), parts.detach(), p.grad, q.grad))
    (ta, pa, gpa, gqa), (tb, pb, gpb, gqb) = res
    assert ta.dtype == tb.dtype == torch.float64
    assert abs(float(ta) - float(tb)) <= 1e-6 * abs(float(tb))
    assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-9), (pa, pb)
    e_p, e_q = rel_err(gpa, gpb), rel_err(gqa, gqb)
    report(f"fused vs separate losses {kind}{n} loop={loop} scale={scale}", (e_p, e_q))
    assert e_p < 2e-6 and e_q < 2e-6, (e_p, e_q)
    # deterministic
    p = pos0.clone().to(DEV).requires_grad_(True)
    q = nrm0.clone().to(DEV).requires_grad_(True)
    t2, _ = L.dual_loss(p, q, n_mesh, n_mesh.vs, n_mesh.fn, k, loop, scale)
    (t2 * 1.7).backward()
    assert torch.equal(t2, ta) and torch.equal(p.grad, gpa) and torch.equal(q.grad, gqa)
