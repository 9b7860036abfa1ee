"""Whole-network and whole-step parity: identical weights, noise input and mesh in the oracle and in the product.
Per-layer activations and per-step gradients within 1e-4 relative (BASELINE.json north_star); GCN biases feed
straight into BatchNorm, so their true gradient is 0 and both sides hold rounding noise — they are compared with an
absolute floor (SURVEY.md §7 hard part 4)."""
import numpy as np
import pytest
import torch

from tests.helpers import oracle_like, product_masks, rel_err, report, small_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(seed=0):
    from dual_dmp_b200.util.networks import NormalNet, PosNet
    from oracle.networks_ref import NormalNetRef, PosNetRef
    torch.manual_seed(seed)
    pr, nr = PosNetRef(), NormalNetRef()
    # non-trivial BatchNorm affine parameters and conv biases, so every parameter gradient is exercised
    with torch.no_grad():
        for net in (pr, nr):
            for i in range(1, 13):
                getattr(net, f"bn{i}").weight.uniform_(0.5, 1.5)
                getattr(net, f"bn{i}").bias.normal_(0, 0.2)
                getattr(net, f"conv{i}").bias.normal_(0, 0.1)
    pd, nd = PosNet(DEV).to(DEV), NormalNet(DEV).to(DEV)
    pd.load_state_dict(pr.state_dict())
    nd.load_state_dict(nr.state_dict())
    return pr, nr, pd, nd


def _compare_grads(net_d, net_x, tag, tol=1e-4):
    """parameter gradients of the product against an oracle network (fp32 oracle, or the float64 oracle evaluated on
    the product's own LeakyReLU active sets — see tests/helpers.MaskedLeaky)"""
    worst = 0.0
    for (name, pd), (_, px) in zip(net_d.named_parameters(), net_x.named_parameters()):
        assert pd.grad is not None, name
        if name.startswith("conv") and name.endswith(".bias"):
            # true gradient of a bias that feeds BatchNorm is 0; both sides hold rounding noise
            floor = 1e-5 * max(1.0, float(px.grad.abs().max()))
            assert float(pd.grad.abs().max()) < 1e-4 + floor, (name, float(pd.grad.abs().max()))
            continue
        e = rel_err(pd.grad, px.grad)
        worst = max(worst, e)
        assert e < tol, (tag, name, e)
    return worst


def _count_flips(net_d, taps_r):
    """LeakyReLU sign disagreements between the product and the fp32 oracle (pre-activations within rounding of 0)"""
    masks = product_masks(net_d)
    flips = 0
    for m, (y_r, x_r) in zip(masks, taps_r):
        flips += int((m != (x_r > 0)).sum())
    return flips, masks


@pytest.mark.parametrize("kind,n", [("ico", 8), ("ico", 20), ("open", 12)])
@pytest.mark.parametrize("reorder", [True, False])
def test_per_layer_activations_and_grads(kind, n, reorder):
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    n_mesh, s_mesh, _ = small_case(kind, n)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, nr, pd, nd = _pair()
    for net_r, net_d, gshape in ((pr, pd, len(n_mesh.vs)), (nr, nd, len(n_mesh.faces))):
        net_r.train(); net_d.train()
        net_d.reorder = reorder
        taps_r, taps_d = [], []
        net_d.taps = taps_d
        out_r = net_r(ds, taps_r)
        out_d = net_d(ds)
        g = torch.randn(gshape, 3, generator=torch.Generator().manual_seed(5))
        out_r.backward(g)
        out_d.backward(g.to(DEV))
        # ---- per-layer activations against the fp32 oracle ----
        worst_act = 0.0
        p = torch.from_numpy(net_d.last_graph.perm_host)          # product row i is node perm[i]
        assert (net_d.last_graph.perm is not None) == reorder
        for l, ((y_r, x_r), (y_d, st)) in enumerate(zip(taps_r, taps_d)):
            x_d = torch.nn.functional.leaky_relu(y_d * st[2] + st[3], 0.01)
            e_y, e_x = rel_err(y_d.cpu(), y_r[p]), rel_err(x_d.cpu(), x_r[p])
            worst_act = max(worst_act, e_y, e_x)
            assert e_y < 1e-4 and e_x < 1e-4, (l, e_y, e_x)
        e_out = rel_err(out_d, out_r)
        assert e_out < 1e-4
        # ---- gradients: fp32 oracle evaluated on the product's LeakyReLU active sets (well-posed even when a
        #      pre-activation within rounding of zero got a different sign); the plain oracle too when none flipped
        flips, masks = _count_flips(net_d, taps_r)
        net_m = oracle_like(net_r, masks)
        net_m(ds).backward(g)
        e_gm = _compare_grads(net_d, net_m, f"{kind}{n} vs oracle on product active sets")
        e_g32 = _compare_grads(net_d, net_r, f"{kind}{n} vs oracle") if flips == 0 else float("nan")
        report(f"net {type(net_d).__name__} {kind}{n} reorder={reorder}", (worst_act, e_out, e_gm, e_g32, flips))
        # BatchNorm running statistics follow the reference semantics
        for i in (1, 6, 12):
            assert rel_err(getattr(net_d, f"bn{i}").running_mean, getattr(net_r, f"bn{i}").running_mean) < 1e-4
            assert rel_err(getattr(net_d, f"bn{i}").running_var, getattr(net_r, f"bn{i}").running_var) < 1e-4
            assert int(getattr(net_d, f"bn{i}").num_batches_tracked) == 1


@pytest.mark.parametrize("cfg", [dict(k=(3.0, 4.0, 4.0, 4.0, 1.0), loop=1), dict(k=(3.0, 0.0, 3.0, 4.0, 2.0), loop=5)])
def test_full_step_losses_and_gradients(cfg):
    """reference main.py:88-110 with the default and the CAD loss weights; three optimiser steps"""
    from dual_dmp_b200.util import loss as L
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    from oracle import step_ref
    n_mesh, s_mesh, _ = small_case("ico", 12)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, nr, pd, nd = _pair(1)
    k = cfg["k"]
    opt_r = (torch.optim.Adam(pr.parameters(), lr=0.01), torch.optim.Adam(nr.parameters(), lr=0.01))
    opt_d = (torch.optim.Adam(pd.parameters(), lr=0.01), torch.optim.Adam(nd.parameters(), lr=0.01))
    for it in range(3):
        for o in opt_r + opt_d:
            o.zero_grad()
        taps_p, taps_n = [], []
        pr.train(); nr.train()
        tot_r, parts_r, _, _ = step_ref.losses(pr, nr, ds, n_mesh, k, cfg["loop"], epoch=101)
        tot_r.backward()
        pd.train(); nd.train()
        pd.taps, nd.taps = [], []
        pos = pd(ds)
        l1 = L.pos_rec_loss(pos, n_mesh.vs)
        l2 = L.mesh_laplacian_loss(pos, n_mesh)
        nrm = nd(ds)
        l3 = L.norm_rec_loss(nrm, n_mesh.fn)
        l4, _ = L.fn_bnf_loss(pos, nrm, n_mesh, loop=cfg["loop"])
        l5 = L.pos_norm_loss(pos, nrm, n_mesh)
        tot = k[0] * l1 + k[1] * l2 + k[2] * l3 + k[3] * l4 + k[4] * l5
        tot.backward()
        parts_d = [x.item() for x in (l1, l2, l3, l4, l5)]
        e_l = max(abs(a - b.item()) / (abs(b.item()) + 1e-9) for a, b in zip(parts_d, parts_r))
        if it == 0:
            assert e_l < 1e-4, (parts_d, [p.item() for p in parts_r])
            # gradients of the whole step: fp32 oracle evaluated on the product's LeakyReLU active sets
            pm, nm = oracle_like(pr, product_masks(pd)), oracle_like(nr, product_masks(nd))
            pm.train(); nm.train()
            totm, _, _, _ = step_ref.losses(pm, nm, ds, n_mesh, k, cfg["loop"], epoch=101)
            totm.backward()
            e_gp = _compare_grads(pd, pm, "posnet step")
            e_gn = _compare_grads(nd, nm, "normnet step")
            report(f"step k={k} loop={cfg['loop']}", (e_l, e_gp, e_gn))
        else:
            # Adam turns rounding noise of near-zero gradients into +-lr steps (sign-like update), so later
            # iterations are only required to track the oracle loosely
            assert e_l < 5e-2, (it, parts_d, [p.item() for p in parts_r])
        for net in (nr, nd):
            torch.nn.utils.clip_grad_norm_(net.parameters(), 0.8)
        for o in opt_r + opt_d:
            o.step()


def test_permutation_equivariance_and_determinism():
    """relabelling nodes (the Morton reorder) must not change the result; two runs are bitwise identical"""
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    n_mesh, s_mesh, _ = small_case("ico", 16)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    _, _, pd, nd = _pair(2)
    for net in (pd, nd):
        net.train()
        net.reorder = True
        a = net(ds).detach().clone()
        b = net(ds).detach().clone()
        assert torch.equal(a, b)
        net.reorder = False
        c = net(ds).detach()
        assert rel_err(a, c) < 1e-4


def test_eval_mode_uses_running_statistics():
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    n_mesh, s_mesh, _ = small_case("ico", 6)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, _, pd, _ = _pair(3)
    pr.train(); pd.train()
    for _ in range(2):
        pr(ds); pd(ds)
    pr.eval(); pd.eval()
    with torch.no_grad():
        assert rel_err(pd(ds), pr(ds)) < 1e-4


def test_dual_step_cuda_graph_matches_eager_and_oracle():
    """dual_dmp_b200.step.DualStep (reference main.py:88-110 as one object): the CUDA-graph replayed iteration gives
    the same losses as the eager drop-in loop, crosses the epoch-100 BNF switch, and tracks the CPU oracle"""
    import copy
    from dual_dmp_b200.step import DualStep
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    from oracle import step_ref
    n_mesh, s_mesh, _ = small_case("ico", 10)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, nr, pd, nd = _pair(4)
    pd2, nd2 = copy.deepcopy(pd), copy.deepcopy(nd)
    graph_step = DualStep(pd, nd, ds, n_mesh, bnfloop=2, capture=True)
    eager_step = DualStep(pd2, nd2, ds, n_mesh, bnfloop=2, capture=False)
    opt_r = (torch.optim.Adam(pr.parameters(), lr=0.01), torch.optim.Adam(nr.parameters(), lr=0.01))
    epochs = [97, 98, 99, 100, 101, 102, 103]           # 3 eager warm-ups, capture (bnf off), switch, capture (bnf on)
    lg, le, lo = [], [], []
    for ep in epochs:
        lg.append(graph_step.step(ep).clone())
        le.append(eager_step.step(ep).clone())
        lo.append(step_ref.train_step(pr, nr, opt_r[0], opt_r[1], ds, n_mesh, bnfloop=2, epoch=ep)[0])
    torch.cuda.synchronize()
    assert len(graph_step._graphs) == 2
    lg, le, lo = [float(x) for x in lg], [float(x) for x in le], [float(x) for x in lo]
    report("dual_step losses graph/eager/oracle", (lg, le, lo))
    assert abs(lg[0] - le[0]) <= 1e-6 * abs(le[0]), (lg, le)      # same kernels, same order
    for a, b in zip(lg, le):                            # capturable vs plain Adam differ in the last bits, and Adam
        assert abs(a - b) <= 1e-2 * abs(b), (lg, le)    # turns them into +-lr steps on noise-level gradients
    assert abs(lg[0] - lo[0]) <= 1e-4 * abs(lo[0])
    for a, b in zip(lg, lo):                            # later iterations: Adam amplifies rounding, track loosely
        assert abs(a - b) <= 5e-2 * abs(b), (lg, lo)
    assert graph_step.pos.shape == (len(n_mesh.vs), 3) and torch.isfinite(graph_step.pos).all()


def test_dual_step_streamed_inputs():
    """DualStep.prefetch: per-step uploads from pinned host memory on the copy stream.  Streaming the SAME inputs
    reproduces the resident run bit for bit; streaming different inputs reaches the captured graph (the loss changes
    to what a resident run on those inputs gives)"""
    import copy
    from dual_dmp_b200.step import DualStep
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    n_mesh, s_mesh, _ = small_case("ico", 8)
    ds = dataset_from_meshes(n_mesh, s_mesh)
    _, _, pd, nd = _pair(5)
    pd2, nd2 = copy.deepcopy(pd), copy.deepcopy(nd)
    pd3, nd3 = copy.deepcopy(pd), copy.deepcopy(nd)
    resident = DualStep(pd, nd, ds, n_mesh)
    streamed = DualStep(pd2, nd2, ds, n_mesh)
    host = copy.copy(ds).pin_memory()
    vs_h, fn_h = torch.from_numpy(n_mesh.vs).pin_memory(), torch.from_numpy(n_mesh.fn).pin_memory()
    a, b = [], []
    for ep in range(101, 108):
        a.append(resident.step(ep).clone())
        nbytes = streamed.prefetch(host, vs_h, fn_h)
        b.append(streamed.step(ep).clone())
    torch.cuda.synchronize()
    assert nbytes == sum(t.numel() * t.element_size() for t in (host.z1, host.z2, host.x_pos, vs_h, fn_h))
    assert [float(x) for x in a] == [float(x) for x in b]
    # different inputs, once the graph is live: compare with a resident stepper built on those inputs
    ds2 = copy.copy(ds)
    ds2.z1 = ds.z1.detach() * 1.25
    ds2.z2 = ds.z2.detach().clone()
    ds2.z2[:, 3:] *= 0.75                          # centroids (which order the face graph) untouched
    ds2.x = ds2.z1
    other = DualStep(pd3, nd3, ds2, n_mesh)
    fresh = DualStep(copy.deepcopy(pd3), copy.deepcopy(nd3), ds, n_mesh)
    host2 = copy.copy(ds2).pin_memory()
    lo, lf = [], []
    for ep in range(101, 107):
        lo.append(other.step(ep).clone())
        fresh.prefetch(host2, vs_h, fn_h)          # built on ds, fed ds2 from the host every step
        lf.append(fresh.step(ep).clone())
    torch.cuda.synchronize()
    assert [float(x) for x in lo] == [float(x) for x in lf]
    with pytest.raises(ValueError):
        fresh.prefetch(host2, vs_h[:-1], fn_h)


@pytest.mark.parametrize("name", ["tetra", "strip2", "ico3", "open4"])
def test_tiny_and_boundary_meshes_match_oracle(golden_dir, name):
    """edge cases of the reference's own fixtures: 4-vertex / 2-face meshes (fewer rows than one tile), an open mesh
    whose f2f has -1 entries and whose face graph has degree-2 rows; forward, all five losses and the parameter
    gradients of the whole step against the oracle"""
    import os
    import numpy as np
    from dual_dmp_b200.util import loss as L
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    from dual_dmp_b200.util.mesh import Mesh
    from oracle import step_ref
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    rng = np.random.RandomState(1)
    n_mesh = Mesh(vs=g["vs"] + 0.05 * rng.randn(*g["vs"].shape), faces=g["faces"])
    s_mesh = Mesh(vs=g["vs"], faces=g["faces"])
    ds = dataset_from_meshes(n_mesh, s_mesh)
    pr, nr, pd, nd = _pair(7)
    pr.train(); nr.train(); pd.train(); nd.train()
    pd.taps, nd.taps = [], []
    pos = pd(ds)
    nrm = nd(ds)
    ls = [L.pos_rec_loss(pos, n_mesh.vs), L.mesh_laplacian_loss(pos, n_mesh), L.norm_rec_loss(nrm, n_mesh.fn)]
    l4, _ = L.fn_bnf_loss(pos, nrm, n_mesh, loop=3)
    ls += [l4, L.pos_norm_loss(pos, nrm, n_mesh)]
    (3.0 * ls[0] + 4.0 * ls[1] + 4.0 * ls[2] + 4.0 * ls[3] + 1.0 * ls[4]).backward()
    pm, nm = oracle_like(pr, product_masks(pd)), oracle_like(nr, product_masks(nd))
    pm.train(); nm.train()
    tot, parts, pos_r, nrm_r = step_ref.losses(pm, nm, ds, n_mesh, (3.0, 4.0, 4.0, 4.0, 1.0), 3, epoch=101)
    tot.backward()
    e_pos, e_nrm = rel_err(pos, pos_r), rel_err(nrm, nrm_r)
    # absolute floor: on the 2- and 4-face meshes the filtered normals coincide and the BNF loss is ~1e-7
    e_l = max(abs(float(a) - float(b)) / (abs(float(b)) + 1e-2) for a, b in zip(ls, parts))
    report(f"tiny mesh {name}", (e_pos, e_nrm, e_l))
    # with a handful of rows the batch statistics are ill-conditioned (variance of 2-4 samples), so 5e-3 here
    assert e_pos < 5e-3 and e_nrm < 5e-3 and e_l < 5e-3, (e_pos, e_nrm, e_l)
    assert torch.isfinite(pos).all() and torch.isfinite(nrm).all()
    for net in (pd, nd):
        for p_ in net.parameters():
            assert p_.grad is not None and torch.isfinite(p_.grad).all()


def test_empty_inputs_are_rejected_or_noops():
    """n = 0 rows: kernels are no-ops; shape mismatches raise instead of reading out of bounds"""
    from dual_dmp_b200 import functional as F_
    from dual_dmp_b200.graph import GcnGraph
    g = GcnGraph(torch.zeros(2, 0, dtype=torch.long), 5, DEV, reorder=False)
    H = torch.randn(5, 32, device=DEV)
    assert torch.equal(F_.spmm_gcn(g, H), H)                       # isolated nodes: A_hat = I
    W = torch.randn(64, 32, device=DEV)
    assert F_.gemm_xw(torch.empty(0, 32, device=DEV), W).shape == (0, 64)
    net_in = torch.randn(7, 16, device=DEV)
    from dual_dmp_b200.util.networks import PosNet
    from types import SimpleNamespace
    ds = SimpleNamespace(z1=net_in, x_pos=torch.zeros(5, 3, device=DEV), edge_index=torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises((RuntimeError, ValueError)):
        PosNet(DEV).to(DEV)(ds)                                     # 7 feature rows vs 5 positions
