"""Whole-step parity against the UNMODIFIED oracle at a size where the graph no longer fits one CTA wave: icosphere
n = 64 (81,920 faces / 40,962 vertices), identical weights, noise input and mesh (BASELINE.json north_star).

Three CPU evaluations of the same step stand beside the product (P):
  A  fp32 oracle, unmodified                      -- "the reference's PyG path"
  C  float64 oracle, unmodified                   -- the same algorithm in (practically) exact arithmetic
  B  float64 oracle on P's LeakyReLU active sets  -- tests/helpers.MaskedLeaky

* activations / outputs / loss terms:  P vs A  <= 1e-4   (unmodified oracle, the north_star bar)
* gradients, well-posed form, in two halves so that one kink cannot hide the rest:
    - loss gradients: the product's d total/d pos, d total/d norm against the float64 oracle LOSSES evaluated at the
      product's own outputs (<= 1e-4, measured 4e-5 / 2e-6; a handful of rows whose |(p-c).n| or |n-fn| term sits within rounding of its kink
      may take the other sign and are counted, not compared);
    - network gradients: P vs B with the SAME upstream gradients, <= 1e-4 on both GEMM paths (tcgen05 fp16 split and
      FFMA) -- every kernel of the backward pass at this size;
* gradients vs the unmodified oracle:  |P - A| is bounded by the reference's OWN rounding sensitivity |A - C|.
  A pre-activation within rounding of zero takes a different LeakyReLU slope in any two fp32 evaluations; the
  number of such rows grows with N while each one's weight falls as 1/N, so the effect decays only like
  1/sqrt(N): measured |A - C| is 1e-3 .. 6e-3 at 8K AND at 82K faces (scripts/e2e_mad.py header).
  A 1e-4 bar against A is therefore not attainable by ANY second fp32 implementation (A itself misses it against
  C); what is attainable, and asserted, is that P is as close to A as A is to exact arithmetic (3 x |A - C| + 5e-3;
  |A - C| itself moves between 6e-4 and 2.5e-3 with the oracle's thread count).
"""
import copy

import pytest
import torch

from tests.helpers import dataset64, oracle_like, product_masks, rel_err, report, small_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K = (3.0, 4.0, 4.0, 4.0, 1.0)


def _grad_errs(net_a, net_b):
    worst, name_w = 0.0, None
    for (name, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        if name.startswith("conv") and name.endswith(".bias"):
            continue                       # exact gradient 0 (feeds BatchNorm): rounding noise on both sides
        e = rel_err(pa.grad, pb.grad)
        if e > worst:
            worst, name_w = e, name
    return worst, name_w


def _product_step(pa, na, ds, n_mesh, backend):
    """one product step (forward, five losses, backward) with the given GEMM backend (0 = auto: tcgen05 where the width
    allows; 1 = FFMA everywhere); also returns the loss gradients d total / d pos, d total / d norm it back-propagated"""
    from dual_dmp_b200 import functional as F_
    from dual_dmp_b200.util import loss as L
    from dual_dmp_b200.util.networks import NormalNet, PosNet
    pd, nd = PosNet(DEV).to(DEV), NormalNet(DEV).to(DEV)
    pd.load_state_dict(pa.state_dict())
    nd.load_state_dict(na.state_dict())
    old = F_.GEMM_BACKEND
    F_.GEMM_BACKEND = backend
    try:
        pd.train(); nd.train()
        pd.taps, nd.taps = [], []
        pos = pd(ds)
        nrm = nd(ds)
        l4, _ = L.fn_bnf_loss(pos, nrm, n_mesh, loop=1)
        parts = [L.pos_rec_loss(pos, n_mesh.vs), L.mesh_laplacian_loss(pos, n_mesh), L.norm_rec_loss(nrm, n_mesh.fn), l4,
                 L.pos_norm_loss(pos, nrm, n_mesh)]
        total = sum(k * l for k, l in zip(K, parts))
        gp, gn = torch.autograd.grad(total, [pos, nrm], retain_graph=True)
        torch.autograd.backward([pos, nrm], [gp, gn])
        torch.cuda.synchronize()
    finally:
        F_.GEMM_BACKEND = old
    return pd, nd, pos.detach(), nrm.detach(), [float(x.detach()) for x in parts], gp, gn


def _masked_f64_grads(pa, na, pd, nd, ds64, gp, gn):
    """B: float64 oracle networks on the product's LeakyReLU active sets, driven by the SAME upstream gradients the
    product back-propagated (the losses' own kinks -- |.| of the L1 / pos_norm terms -- are checked separately)"""
    pb, nb = oracle_like(pa, product_masks(pd), double=True), oracle_like(na, product_masks(nd), double=True)
    pb.train(); nb.train()
    pb(ds64).backward(gp.double().cpu())
    nb(ds64).backward(gn.double().cpu())
    return _grad_errs(pd, pb), _grad_errs(nd, nb)


def _loss_gradient_parity(pos, nrm, gp, gn, n_mesh):
    """the product's loss gradients against the float64 oracle losses evaluated AT the product's outputs; a face whose
    |(p - c).n| or |n - fn| term sits within rounding of its kink may take the other sign: such rows are counted"""
    from oracle import loss_ref as LR
    p = pos.double().cpu().requires_grad_(True)
    q = nrm.double().cpu().requires_grad_(True)
    l4, _ = LR.fn_bnf_loss(p, q, n_mesh, loop=1)
    parts = [LR.pos_rec_loss(p, n_mesh.vs), LR.mesh_laplacian_loss(p, n_mesh), LR.norm_rec_loss(q, n_mesh.fn), l4,
             LR.pos_norm_loss(p, q, n_mesh)]
    sum(k * l for k, l in zip(K, parts)).backward()
    out = []
    for g, ref in ((gp, p.grad), (gn, q.grad)):
        err = (g.double().cpu() - ref).abs().amax(dim=1) / ref.abs().max()
        bad = int((err > 1e-4).sum())
        out.append((bad, float(err[err <= 1e-4].max()) if bad < len(err) else float("nan"), float(err.max())))
    return out


def test_step_vs_unmodified_oracle_81920_faces():
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    from oracle import loss_ref as LR
    from oracle import step_ref
    from oracle.networks_ref import NormalNetRef, PosNetRef
    n_mesh, s_mesh, _ = small_case("ico", 64)
    assert len(n_mesh.faces) == 81920
    ds = dataset_from_meshes(n_mesh, s_mesh)
    torch.manual_seed(1)
    pa, na = PosNetRef(), NormalNetRef()
    pc, nc = copy.deepcopy(pa).double(), copy.deepcopy(na).double()

    # ---- P: product, tensor-core path (default) and FFMA path ----------------------------------------------------
    pd, nd, pos, nrm, parts_p, gp, gn = _product_step(pa, na, ds, n_mesh, backend=0)
    pf, nf, pos_f, nrm_f, _, gpf, gnf = _product_step(pa, na, ds, n_mesh, backend=1)

    # ---- A: unmodified fp32 oracle ------------------------------------------------------------------------------
    pa.train(); na.train()
    taps_pa, taps_na = [], []
    pos_a = pa(ds, taps_pa)
    nrm_a = na(ds, taps_na)
    l4a, _ = LR.fn_bnf_loss(pos_a, nrm_a, n_mesh, loop=1)
    parts_a = [LR.pos_rec_loss(pos_a, n_mesh.vs), LR.mesh_laplacian_loss(pos_a, n_mesh),
               LR.norm_rec_loss(nrm_a, n_mesh.fn), l4a, LR.pos_norm_loss(pos_a, nrm_a, n_mesh)]
    sum(k * l for k, l in zip(K, parts_a)).backward()

    # activations of every layer, outputs, loss terms: unmodified oracle, 1e-4
    worst_act = 0.0
    for net_d, taps_r in ((pd, taps_pa), (nd, taps_na)):
        p = torch.from_numpy(net_d.last_graph.perm_host)
        for l, ((y_r, x_r), (y_d, st)) in enumerate(zip(taps_r, net_d.taps)):
            x_d = torch.nn.functional.leaky_relu(y_d * st[2] + st[3], 0.01)
            e_y, e_x = rel_err(y_d.cpu(), y_r[p]), rel_err(x_d.cpu(), x_r[p])
            worst_act = max(worst_act, e_y, e_x)
            assert e_y < 1e-4 and e_x < 1e-4, (type(net_d).__name__, l, e_y, e_x)
    e_pos, e_nrm = rel_err(pos, pos_a), rel_err(nrm, nrm_a)
    assert e_pos < 1e-4 and e_nrm < 1e-4, (e_pos, e_nrm)
    assert rel_err(pos_f, pos_a) < 1e-4 and rel_err(nrm_f, nrm_a) < 1e-4
    e_loss = max(abs(a - float(b.detach())) / abs(float(b.detach())) for a, b in zip(parts_p, parts_a))
    assert e_loss < 1e-4, e_loss
    flips = 0
    for masks, taps in ((product_masks(pd), taps_pa), (product_masks(nd), taps_na)):
        flips += sum(int((m != (x_r > 0)).sum()) for m, (y_r, x_r) in zip(masks, taps))
    del taps_pa, taps_na

    # ---- loss gradients at the product's outputs vs float64 losses ------------------------------------------------
    (bad_p, e_gp, worst_gp), (bad_n, e_gn, worst_gn) = _loss_gradient_parity(pos, nrm, gp, gn, n_mesh)

    # ---- B: float64 oracle networks on the product's active sets, same upstream gradients -----------------------------
    ds64 = dataset64(ds)
    (e_pb, name_pb), (e_nb, name_nb) = _masked_f64_grads(pa, na, pd, nd, ds64, gp, gn)          # tensor-core path
    (e_pbf, name_pbf), (e_nbf, name_nbf) = _masked_f64_grads(pa, na, pf, nf, ds64, gpf, gnf)    # FFMA path

    # ---- C: unmodified float64 oracle: the reference's own rounding sensitivity |A - C| -------------------------
    pc.train(); nc.train()
    tot_c, _, _, _ = step_ref.losses(pc, nc, ds64, n_mesh, K, 1, epoch=101)
    tot_c.backward()
    e_pa, name_pa = _grad_errs(pd, pa)          # product vs unmodified fp32 oracle
    e_na, name_na = _grad_errs(nd, na)
    s_p, _ = _grad_errs(pa, pc)                 # fp32 oracle vs its own float64 evaluation
    s_n, _ = _grad_errs(na, nc)
    report("step n=64 (81,920 faces): act / pos / nrm / loss vs unmodified fp32 oracle", (worst_act, e_pos, e_nrm, e_loss))
    report("step n=64: LeakyReLU sign flips product vs fp32 oracle", flips)
    report("step n=64: loss gradients vs float64 losses at the product's outputs: rows off a kink (pos, norm) / err",
           (bad_p, e_gp, worst_gp, bad_n, e_gn, worst_gn))
    report("step n=64: network grads vs float64 oracle on product active sets, tensor-core path (posnet, normnet)",
           (e_pb, name_pb, e_nb, name_nb))
    report("step n=64: network grads vs float64 oracle on product active sets, FFMA path (posnet, normnet)",
           (e_pbf, name_pbf, e_nbf, name_nbf))
    report("step n=64: grads vs UNMODIFIED fp32 oracle (posnet, normnet)", (e_pa, name_pa, e_na, name_na))
    report("step n=64: UNMODIFIED fp32 oracle vs UNMODIFIED float64 oracle (posnet, normnet)", (s_p, s_n))
    assert bad_p <= 8 and bad_n <= 8 and e_gp < 1e-4 and e_gn < 1e-4, (bad_p, e_gp, bad_n, e_gn)
    assert e_pbf < 1e-4 and e_nbf < 1e-4, (e_pbf, name_pbf, e_nbf, name_nbf)
    assert e_pb < 1e-4 and e_nb < 1e-4, (e_pb, name_pb, e_nb, name_nb)
    # |A - C| itself moves between 6e-4 and 2.5e-3 with the host's thread count (reduction order decides which rows
    # flip), so the bound is 3 x that sensitivity plus a 5e-3 floor; measured |P - A|: 2.2e-3 / 1.1e-3
    assert e_pa < 3.0 * s_p + 5e-3, (e_pa, s_p, name_pa)
    assert e_na < 3.0 * s_n + 5e-3, (e_na, s_n, name_na)
