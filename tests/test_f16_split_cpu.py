"""The numerical contract of the fp16-split tensor-core GEMMs (csrc/gemm_tc.cu), restated on the CPU:
the power-of-two operand scale derived from a bound, the hi/lo split, the three-term product sum, and the BatchNorm
bound the training step supplies.  The CUDA kernels are checked against float64 in tests/test_gpu_ops.py and
tests/test_gpu_large.py; this file pins the arithmetic they implement."""
import numpy as np
import torch


def scale_from_bits(bound: np.ndarray):
    """mirror of tc::scale_from_bits: E = biased exponent of the bound, clamped to [16, 252]; s = 2^(141-E)"""
    bits = np.abs(bound).astype(np.float32).view(np.uint32)
    E = np.clip((bits >> 23) & 0xFF, 16, 252).astype(np.int64)
    return np.ldexp(1.0, 141 - E), np.ldexp(1.0, E - 141)


def split16(x: torch.Tensor):
    hi = x.to(torch.float16)
    lo = (x - hi.to(torch.float32)).to(torch.float16)
    return hi, lo


def test_scale_keeps_the_operand_inside_fp16():
    rng = np.random.RandomState(0)
    bound = np.exp(rng.uniform(np.log(1e-30), np.log(1e30), 20000)).astype(np.float32)
    s, inv = scale_from_bits(bound)
    assert np.all(s * inv == 1.0)
    assert np.all(np.log2(s) == np.round(np.log2(s)))                # powers of two: scaling is exact
    scaled = bound.astype(np.float64) * s
    assert np.all(scaled < 2.0 ** 15) and np.all(scaled >= 2.0 ** 14)   # top of the fp16 range, never above it
    assert np.all(scaled < 65504.0)
    # degenerate bounds: zero and denormal operands get the largest finite scale, inf / NaN the smallest
    s0, _ = scale_from_bits(np.array([0.0, 1e-42], dtype=np.float32))
    assert np.all(np.isfinite(s0)) and np.all(s0 == 2.0 ** 125)
    s1, _ = scale_from_bits(np.array([np.inf, np.nan], dtype=np.float32))
    assert np.all(s1 == 2.0 ** -111)


def test_split_carries_fp32_precision_inside_the_range():
    torch.manual_seed(0)
    x = torch.randn(200000) * torch.exp(torch.randn(200000) * 3)     # 6 decades of magnitudes
    s, _ = scale_from_bits(np.array([float(x.abs().max())], dtype=np.float32))
    xs = x * float(s[0])
    hi, lo = split16(xs)
    rec = hi.double() + lo.double()
    err = (rec - xs.double()).abs()
    big = xs.abs() >= 2.0 ** -2                                      # lo is still a normal fp16 number
    assert float((err[big] / xs.double().abs()[big]).max()) <= 2.0 ** -22
    assert float(err[~big].max()) <= 2.0 ** -25                      # below: half an fp16 subnormal ulp, absolute
    assert torch.isfinite(hi.float()).all()


def test_three_term_product_sum_matches_float64():
    torch.manual_seed(1)
    n, K, N = 512, 512, 64
    X = torch.randn(n, K) * (torch.rand(1, K) * 3 + 0.1)
    W = torch.randn(N, K) / K ** 0.5
    ref = X.double() @ W.double().t()
    sx, ix = scale_from_bits(np.array([float(X.abs().max()) * 1000.0], dtype=np.float32))   # 1000x pessimistic bound
    sw, iw = scale_from_bits(W.abs().amax(1).numpy())                                       # per weight row
    xh, xl = split16(X * float(sx[0]))
    wh, wl = split16(W * torch.from_numpy(sw).float()[:, None])
    acc = xl.double() @ wh.double().t() + xh.double() @ wl.double().t() + xh.double() @ wh.double().t()
    out = acc * float(ix[0]) * torch.from_numpy(iw)[None, :]
    e16 = float((out - ref).norm() / ref.norm())
    e32 = float(((X @ W.t()).double() - ref).norm() / ref.norm())
    assert e16 < 2e-7 and e16 < e32, (e16, e32)                      # better than an fp32 matmul of the same data


def test_batchnorm_bound_is_rigorous_and_attained():
    """|gamma * (x - mean) / sqrt(var + eps) + beta| <= |gamma| sqrt(n - 1) + |beta| for every sample of a batch"""
    torch.manual_seed(2)
    n, C = 4097, 16
    x = torch.randn(n, C, dtype=torch.float64) * 5 + 3
    x[7, 0] = 1e9                                                     # one huge outlier: the extremal case
    x[:, 1] = 0.0
    x[11, 1] = 1.0                                                    # one-hot column attains sqrt(n-1) as eps -> 0
    gamma, beta = torch.randn(C, dtype=torch.float64), torch.randn(C, dtype=torch.float64)
    mean, var = x.mean(0), x.var(0, unbiased=False)
    z = gamma * (x - mean) / torch.sqrt(var + 1e-5) + beta
    bound = gamma.abs() * (n - 1) ** 0.5 + beta.abs()
    assert bool((z.abs() <= bound * (1 + 1e-12)).all())
    assert float(((x[:, 0] - mean[0]) / var[0].sqrt()).abs().max()) > 0.999 * (n - 1) ** 0.5
    # LeakyReLU only shrinks magnitudes, so the bound also covers the activated operand
    assert bool((torch.nn.functional.leaky_relu(z, 0.01).abs() <= bound * (1 + 1e-12)).all())
