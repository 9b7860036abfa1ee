"""CPU restatement of the BatchNorm statistics contract (csrc/spmm_tile.cu, csrc/spmm.cu epilogues + csrc/bn.cu finalize):
per thread Welford in float32 -> per block Chan merge in float32 -> (sum_b, M2_b) -> float64 combine
Q = sum_b (M2_b + sum_b^2 / n_b), var = Q/n - (S/n)^2.  Pins the arithmetic and shows why it replaced the float32
sum / sum-of-squares partials: at |mean|/sigma = 100 the old form loses 3-4 digits of the variance, the new one none."""
import numpy as np
import pytest


def _welford_f32(x):
    mean, m2, cnt = np.float32(0), np.float32(0), np.float32(0)
    for v in x.astype(np.float32):
        cnt += np.float32(1)
        d = v - mean
        mean = np.float32(mean + d * (np.float32(1) / cnt))
        m2 = np.float32(m2 + d * (v - mean))
    return cnt, mean, m2


def _block_moments(y, groups=8):
    """one row block: rows g, g+groups, ... per thread, merged in group order"""
    n_a, mean, m2 = np.float32(0), np.float32(0), np.float32(0)
    for g in range(min(groups, len(y))):
        n_b, mb, qb = _welford_f32(y[g::groups])
        n = n_a + n_b
        d = mb - mean
        mean = np.float32(mean + d * (n_b / n))
        m2 = np.float32(m2 + qb + d * d * (n_a * n_b / n))
        n_a = n
    return np.float32(mean * n_a), m2


def _finalize(sums, m2s, n, rpb):
    nb = np.full(len(sums), float(rpb))
    nb[-1] = n - (len(sums) - 1) * rpb
    S = sums.astype(np.float64).sum()
    Q = (m2s.astype(np.float64) + sums.astype(np.float64) ** 2 / nb).sum()
    return S / n, Q / n - (S / n) ** 2


@pytest.mark.parametrize("ratio", [0.0, 100.0, 1000.0])
def test_moment_partials_keep_the_variance(ratio):
    rng = np.random.RandomState(0)
    n, rpb = 5000, 128
    y = (rng.randn(n) * 1.7 + ratio * 1.7).astype(np.float32)
    blocks = [y[i:i + rpb] for i in range(0, n, rpb)]
    pm = np.array([_block_moments(b) for b in blocks])
    mean, var = _finalize(pm[:, 0], pm[:, 1], n, rpb)
    ref_mean, ref_var = y.astype(np.float64).mean(), y.astype(np.float64).var()
    assert abs(mean - ref_mean) <= 2e-7 * max(1.0, abs(ref_mean))
    assert abs(var - ref_var) <= (2e-6 + 4e-7 * ratio) * ref_var
    # the float32 sum / sum-of-squares partials this replaced
    s_old = np.array([b.sum(dtype=np.float32) for b in blocks])
    q_old = np.array([(b * b).sum(dtype=np.float32) for b in blocks])
    var_old = q_old.astype(np.float64).sum() / n - (s_old.astype(np.float64).sum() / n) ** 2
    if ratio >= 100.0:
        assert abs(var_old - ref_var) > 20 * abs(var - ref_var)


def test_single_block_and_ragged_last_block():
    rng = np.random.RandomState(1)
    for n, rpb in ((5, 128), (129, 128), (256, 256)):
        y = (rng.randn(n) + 3.0).astype(np.float32)
        blocks = [y[i:i + rpb] for i in range(0, n, rpb)]
        pm = np.array([_block_moments(b, groups=32) for b in blocks])
        mean, var = _finalize(pm[:, 0], pm[:, 1], n, rpb)
        assert abs(mean - y.astype(np.float64).mean()) < 1e-6 and abs(var - y.astype(np.float64).var()) < 1e-5
