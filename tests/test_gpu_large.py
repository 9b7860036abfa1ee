"""Size-independent properties at the benchmark size (1M faces), where the oracle cannot run: determinism,
permutation equivariance of the Morton reorder, the A_hat symmetry checksum of the SpMM, tcgen05 vs FFMA GEMM."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_err, report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def big():
    from dual_dmp_b200 import synth
    from dual_dmp_b200.util.datamaker import dataset_from_meshes
    from dual_dmp_b200.util.mesh import Mesh
    case = synth.make_case(224)
    n_mesh, s_mesh = Mesh(vs=case.noise_vs, faces=case.faces), Mesh(vs=case.smooth_vs, faces=case.faces)
    return n_mesh, s_mesh, dataset_from_meshes(n_mesh, s_mesh)


def test_graph_build_invariants_1m(big):
    from dual_dmp_b200.graph import GcnGraph
    n_mesh, _, ds = big
    F, V = len(n_mesh.faces), len(n_mesh.vs)
    assert (F, V) == (1003520, 501762) and len(n_mesh.edges) == 1505280
    deg = n_mesh.v_dims.numpy()
    assert (deg == 5).sum() == 12 and (deg == 6).sum() == V - 12 and (n_mesh.f2f >= 0).all()
    g = GcnGraph(ds.face_index, F, DEV, coords=ds.z2.detach()[:, :3])
    assert g.symmetric and g.nnz == 4 * F and not g.identity
    rowptr = g.rowptr.cpu().numpy()
    assert (np.diff(rowptr) == 4).all()
    assert torch.allclose(g.w, torch.full_like(g.w, 0.25))
    # locality of the Morton order: median |row - col| is tiny compared with F
    col = g.col.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(F), 4)
    assert np.median(np.abs(rows - col)) < 2000


def test_spmm_checksum_and_determinism_1m(big):
    """A_hat symmetric  =>  1^T (A_hat H) = (A_hat 1)^T H ;  two runs are bitwise identical"""
    from dual_dmp_b200 import functional as F_
    from dual_dmp_b200.graph import GcnGraph
    n_mesh, _, ds = big
    V = len(n_mesh.vs)
    g = GcnGraph(ds.edge_index, V, DEV, coords=ds.x_pos)
    torch.manual_seed(0)
    for C in (64, 512):
        H = torch.randn(V, C, device=DEV)
        Y, partials = F_.spmm_gcn(g, H, stats=True)
        Y2, partials2 = F_.spmm_gcn(g, H, stats=True)
        assert torch.equal(Y, Y2) and torch.equal(partials, partials2)
        rowsum = F_.spmm_gcn(g, torch.ones(V, 4, device=DEV))[:, 0].double()
        lhs = Y.double().sum(dim=0)
        rhs = (rowsum[:, None] * H.double()).sum(dim=0)
        e = rel_err(lhs, rhs)
        report(f"spmm checksum 1M C={C}", e)
        assert e < 1e-5
        assert rel_err(partials.double().sum(dim=0)[0], lhs) < 1e-6


def test_tc_and_ffma_gemm_agree_1m():
    from dual_dmp_b200 import functional as F_
    torch.manual_seed(1)
    n = 300001
    X = torch.randn(n, 256, device=DEV)
    W = torch.randn(512, 256, device=DEV) / 16
    dH = torch.randn(n, 512, device=DEV)
    sc, sh = torch.rand(256, device=DEV) + 0.5, torch.randn(256, device=DEV)
    assert rel_err(F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2), F_.gemm_xw(X, W, scale=sc, shift=sh, backend=1)) < 1e-5
    assert rel_err(F_.gemm_dx(dH, W, backend=2), F_.gemm_dx(dH, W, backend=1)) < 1e-5
    a, b = F_.gemm_dw(dH, X, 256, scale=sc, shift=sh, backend=2), F_.gemm_dw(dH, X, 256, scale=sc, shift=sh, backend=1)
    e = rel_err(a, b)
    report("gemm_dw tc vs ffma n=300001", e)
    assert e < 5e-5


def test_f16_split_gemms_full_size_vs_float64():
    """the fp16-split tensor-core GEMMs (CTA-pair kernels, the default of the training step) at the full benchmark
    shape, 1,003,520 rows x 512 -> 512, with the operand bounds the step supplies: sampled rows of X.W^T and dH.W and
    the whole of dH^T.X against float64, linearity in the weights, bitwise determinism"""
    from dual_dmp_b200 import functional as F_
    torch.manual_seed(2)
    n, C = 1003520, 512
    X = torch.randn(n, C, device=DEV)
    W = torch.randn(C, C, device=DEV) / C ** 0.5
    W2 = torch.randn(C, C, device=DEV) / C ** 0.5
    dH = torch.randn(n, C, device=DEV) * 1e-6 * torch.exp(torch.randn(n, 1, device=DEV))      # gradient-like scales
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    mean, var = X.mean(0), X.var(0, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    sc, sh = gamma * rstd, beta - mean * gamma * rstd
    bound = gamma.abs() * (n - 1) ** 0.5 + beta.abs()                       # what ddmp_bn_stats_finalize emits
    nblk = F_.num_row_blocks(n, C)                                          # what ddmp_spmm_gcn(amax_blocks) emits
    rpb = -(-n // nblk)
    blockmax = torch.nn.functional.pad(dH.abs().amax(1), (0, nblk * rpb - n)).view(nblk, rpb).amax(1).contiguous()
    rows = torch.randint(0, n, (257,), device=DEV)
    act = torch.nn.functional.leaky_relu(X[rows].double() * sc.double() + sh.double(), 0.01)
    H = F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2, amax=bound)
    e_xw = rel_err(H[rows], act @ W.double().t())
    gX = F_.gemm_dx(dH, W, backend=2, amax=blockmax)
    e_dx = rel_err(gX[rows], dH[rows].double() @ W.double())
    dW = F_.gemm_dw(dH, X, C, scale=sc, shift=sh, backend=2, amax_dh=blockmax, amax_x=bound)
    ref = torch.zeros(C, C, dtype=torch.float64, device=DEV)
    for lo in range(0, n, 131072):                                          # float64 reference in slabs (memory)
        a = torch.nn.functional.leaky_relu(X[lo:lo + 131072].double() * sc.double() + sh.double(), 0.01)
        ref += dH[lo:lo + 131072].double().t() @ a
    e_dw = rel_err(dW, ref)
    e_lin = rel_err(F_.gemm_xw(X, W + W2, scale=sc, shift=sh, backend=2, amax=bound)[rows],
                    H[rows].double() + F_.gemm_xw(X, W2, scale=sc, shift=sh, backend=2, amax=bound)[rows].double())
    report("f16-split GEMMs 1M x 512x512 (xw, dx, dw, linearity)", (e_xw, e_dx, e_dw, e_lin))
    assert e_xw < 5e-6 and e_dx < 5e-6 and e_dw < 2e-5 and e_lin < 5e-6, (e_xw, e_dx, e_dw, e_lin)
    assert torch.equal(H, F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2, amax=bound))
    assert torch.equal(dW, F_.gemm_dw(dH, X, C, scale=sc, shift=sh, backend=2, amax_dh=blockmax, amax_x=bound))


@pytest.mark.parametrize("n", [8_000_000, 16_020_500])
def test_f16_split_gemms_at_partitioned_mesh_row_counts(n):
    """the operand bound of activated inputs, |gamma|*sqrt(n-1)+|beta|, grows with the row count (2,800 sigma at 8M rows,
    4,000 sigma at the 16M-face config of BASELINE.json): every doubling of n costs half a bit of the `lo` piece.  At the
    row counts of the partitioned configs (256 -> 256 transform, 8 / 16 GB per tensor) sampled rows of X.W^T and dH.W and
    the whole dH^T.X still meet the 1M-row bars against float64"""
    from dual_dmp_b200 import functional as F_
    free, _ = torch.cuda.mem_get_info()
    if free < 5 * n * 256 * 4:
        pytest.skip("not enough device memory for this row count")
    torch.manual_seed(3)
    C = 256
    X = torch.randn(n, C, device=DEV)
    W = torch.randn(C, C, device=DEV) / C ** 0.5
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    sc, sh = gamma, beta                                                      # X is already ~N(0,1) per channel
    bound = gamma.abs() * (n - 1) ** 0.5 + beta.abs()                         # what ddmp_bn_stats_finalize emits
    rows = torch.randint(0, n, (513,), device=DEV)
    act = torch.nn.functional.leaky_relu(X[rows].double() * sc.double() + sh.double(), 0.01)
    H = F_.gemm_xw(X, W, scale=sc, shift=sh, backend=2, amax=bound)
    e_xw = rel_err(H[rows], act @ W.double().t())
    del H
    dH = torch.randn(n, C, device=DEV) * 1e-6 * torch.exp(torch.randn(n, 1, device=DEV))
    amax = dH.abs().amax().reshape(1).contiguous()
    gX = F_.gemm_dx(dH, W, backend=2, amax=amax)
    e_dx = rel_err(gX[rows], dH[rows].double() @ W.double())
    del gX
    dW = F_.gemm_dw(dH, X, C, scale=sc, shift=sh, backend=2, amax_dh=amax, amax_x=bound)
    ref = torch.zeros(C, C, dtype=torch.float64, device=DEV)
    for lo in range(0, n, 262144):
        a = torch.nn.functional.leaky_relu(X[lo:lo + 262144].double() * sc.double() + sh.double(), 0.01)
        ref += dH[lo:lo + 262144].double().t() @ a
    e_dw = rel_err(dW, ref)
    report(f"f16-split GEMMs {n} x 256x256, bound {float(bound.max()):.0f} (xw, dx, dw)", (e_xw, e_dx, e_dw))
    assert e_xw < 5e-6 and e_dx < 5e-6 and e_dw < 2e-5, (e_xw, e_dx, e_dw)


def test_network_equivariance_and_determinism_1m(big):
    """the Morton relabelling must not change the result (this is what legitimises the reorder); reruns are bitwise
    identical; outputs are finite and NormalNet rows are unit vectors"""
    from dual_dmp_b200.util.networks import NormalNet, PosNet
    n_mesh, _, ds = big
    torch.manual_seed(0)
    for cls in (PosNet, NormalNet):
        net = cls(DEV).to(DEV)
        net.train()
        with torch.no_grad():
            pass
        a = net(ds).detach().clone()
        b = net(ds).detach().clone()
        assert torch.equal(a, b) and torch.isfinite(a).all()
        net.reorder = False
        c = net(ds).detach()
        e = rel_err(a, c)
        report(f"equivariance 1M {cls.__name__}", e)
        assert e < 1e-4
        if cls is NormalNet:
            assert (a.norm(dim=1) - 1).abs().max() < 1e-5
        del net, a, b, c
        torch.cuda.empty_cache()
