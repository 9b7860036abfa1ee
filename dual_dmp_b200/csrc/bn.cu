// BatchNorm1d (training mode) statistics / backward and column sums.
// Replaces nn.BatchNorm1d + nn.LeakyReLU at reference util/networks.py:31-44,51-62 (two full passes forward and
// two backward per layer in the library) by: statistics taken in the producing kernel's epilogue (spmm.cu),
// a tiny fixed-order float64 finalize, and normalise+LeakyReLU applied lazily where the activation is consumed.
// All per-row-block partials are [nblk][sets][C] float32, combined in block order (deterministic).
#include "common.cuh"

namespace ddmp {

// One CTA = kFinCh channels x kFinLanes block-lanes (1024 threads): lane ty accumulates blocks ty, ty+kFinLanes, ...
// in float64 and the lanes are combined in lane order, so the summation order is fixed.  8 channels per CTA (one
// 32-byte sector per block row) keeps the loads sector-exact while giving C/8 CTAs: with 32 channels per CTA the
// 512-wide finalize ran on 16 CTAs and took 59 us per launch (ncu launch list, profiles/).
constexpr int kFinCh = 8;
constexpr int kFinLanes = 128;
constexpr int kFinThreads = kFinCh * kFinLanes;
// MOMENTS (SETS == 2 only): the partials are per-block (sum_b, M2_b about the block mean) as the aggregation kernels
// emit them; each block contributes (sum_b, M2_b + sum_b^2 / n_b) -- its sum of squares, rebuilt in float64 -- with
// n_b = rows_per_block except for the last block.  Summed in float64, var = Q/n - (S/n)^2 then cancels in float64
// instead of float32: relative error ~1e-7 * |mean|/sigma instead of ~1e-7 * (mean/sigma)^2.
template <int SETS, bool MOMENTS = false>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partials, int64_t nblk, int C, int c,
                                                double (&out)[SETS], int64_t n_rows = 0, int rows_per_block = 0) {
    __shared__ double red[kFinLanes][SETS][kFinCh + 1];
    const int tx = threadIdx.x % kFinCh, ty = threadIdx.x / kFinCh;
    double acc[SETS];
#pragma unroll
    for (int s = 0; s < SETS; ++s) acc[s] = 0.0;
    if (c < C) {
        int64_t b = ty;
        constexpr int UN = 8;           // block rows in flight per lane: the loop is a chain of L2 round trips
        for (; b + (UN - 1) * kFinLanes < nblk; b += UN * kFinLanes) {
            float v[UN][SETS];
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int s = 0; s < SETS; ++s) v[u][s] = __ldg(partials + ((b + u * kFinLanes) * SETS + s) * C + c);
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (MOMENTS) {          // b + (UN-1)*kFinLanes < nblk: none of these is the last block
                    acc[0] += (double)v[u][0];
                    acc[SETS - 1] += (double)v[u][SETS - 1] + (double)v[u][0] * (double)v[u][0] / (double)rows_per_block;
                } else {
#pragma unroll
                    for (int s = 0; s < SETS; ++s) acc[s] += (double)v[u][s];
                }
            }
        }
        for (; b < nblk; b += kFinLanes) {
            if (MOMENTS) {
                const double nb = (b == nblk - 1) ? (double)(n_rows - (nblk - 1) * (int64_t)rows_per_block)
                                                  : (double)rows_per_block;
                const double sb = (double)__ldg(partials + (b * SETS) * C + c);
                acc[0] += sb;
                acc[SETS - 1] += (double)__ldg(partials + (b * SETS + SETS - 1) * C + c) + sb * sb / nb;
            } else {
#pragma unroll
                for (int s = 0; s < SETS; ++s) acc[s] += (double)__ldg(partials + (b * SETS + s) * C + c);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < SETS; ++s) red[ty][s][tx] = acc[s];
    __syncthreads();
    // fold the lanes in lane order, in two levels (16 groups of 8 lanes, then the 16 group sums): 8 + 16 dependent adds
    // instead of 128.  The result is valid in the threads with ty == 0 (threadIdx.x < kFinCh), the only ones that use it.
    constexpr int GRP = 8;
    if (ty < kFinLanes / GRP) {
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
            double t = 0.0;
#pragma unroll
            for (int y = 0; y < GRP; ++y) t += red[ty * GRP + y][s][tx];
            acc[s] = t;
        }
    }
    __syncthreads();
    if (ty < kFinLanes / GRP) {
#pragma unroll
        for (int s = 0; s < SETS; ++s) red[ty][s][tx] = acc[s];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SETS; ++s) {
        double t = 0.0;
        if (ty == 0) {
#pragma unroll
            for (int y = 0; y < kFinLanes / GRP; ++y) t += red[y][s][tx];
        }
        out[s] = t;
    }
}

// per-channel batch statistics from S = sum y and Q = sum y^2 (float64)
__device__ __forceinline__ void bn_stats_write(double S, double Q, int64_t n, int c, const float* gamma, const float* beta,
                                               float eps, float momentum, float* running_mean, float* running_var,
                                               float* mean, float* rstd, float* scale, float* shift, float* bound) {
    const double m = S / (double)n;
    double var = Q / (double)n - m * m;
    if (var < 0.0) var = 0.0;
    const float r = (float)(1.0 / sqrt(var + (double)eps));
    const float mf = (float)m;
    const float sc = gamma[c] * r;
    mean[c] = mf;
    rstd[c] = r;
    scale[c] = sc;
    shift[c] = beta[c] - mf * sc;
    // no sample of a batch of n lies more than sqrt(n-1) (biased) standard deviations from the batch mean, so
    // |scale*y + shift| <= |gamma|*sqrt(n-1) + |beta| for every row: the operand bound of the fp16-split GEMMs
    if (bound) bound[c] = fabsf(gamma[c]) * sqrtf((float)(n > 1 ? n - 1 : 1)) + fabsf(beta[c]);
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mf;
    if (running_var) {
        const double unbiased = (n > 1) ? var * ((double)n / (double)(n - 1)) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// partials: [nblk][2][C] per-block (sum, M2 about the block mean) from the aggregation kernels' epilogue
__global__ void __launch_bounds__(kFinThreads)
bn_stats_finalize_kernel(const float* __restrict__ partials, int64_t nblk, int64_t n, int C, int rows_per_block,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                         float* running_mean, float* running_var, float* mean, float* rstd, float* scale,
                         float* shift, float* bound) {
    const int c = blockIdx.x * kFinCh + (threadIdx.x % kFinCh);
    double s[2];
    reduce_partials<2, true>(partials, nblk, C, c, s, n, rows_per_block);
    if (threadIdx.x < kFinCh && c < C)
        bn_stats_write(s[0], s[1], n, c, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale,
                       shift, bound);
}

// the same reduction without the finalize: this rank's (sum y, sum y^2) in float64 (partitioned mode: all-reduced next)
__global__ void __launch_bounds__(kFinThreads)
bn_stats_rank_sums_kernel(const float* __restrict__ partials, int64_t nblk, int64_t n, int C, int rows_per_block,
                          double* __restrict__ sums) {
    const int c = blockIdx.x * kFinCh + (threadIdx.x % kFinCh);
    double s[2];
    reduce_partials<2, true>(partials, nblk, C, c, s, n, rows_per_block);
    if (threadIdx.x < kFinCh && c < C) { sums[c] = s[0]; sums[C + c] = s[1]; }
}

// eval mode: the same [mean, rstd, scale, shift] table from the running statistics (reference: nn.BatchNorm1d.eval())
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int C,
                                     float* mean, float* rstd, float* scale, float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float r = 1.0f / sqrtf(running_var[c] + eps);
    const float sc = gamma[c] * r;
    mean[c] = running_mean[c];
    rstd[c] = r;
    scale[c] = sc;
    shift[c] = beta[c] - running_mean[c] * sc;
}

__global__ void bn_stats_from_sums_kernel(const double* __restrict__ sums, int64_t n, int C,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                          float momentum, float* running_mean, float* running_var, float* mean,
                                          float* rstd, float* scale, float* shift, float* bound) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C)
        bn_stats_write(sums[c], sums[C + c], n, c, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd,
                       scale, shift, bound);
}

__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int64_t nblk, int64_t n, int C, float* dgamma,
                       float* dbeta, float* c1, float* c2) {
    const int c = blockIdx.x * kFinCh + (threadIdx.x % kFinCh);
    double s[2];
    reduce_partials<2>(partials, nblk, C, c, s);
    if (threadIdx.x < kFinCh && c < C) {
        dbeta[c] = (float)s[0];
        dgamma[c] = (float)s[1];
        c1[c] = (float)(s[0] / (double)n);
        c2[c] = (float)(s[1] / (double)n);
    }
}

template <int SETS>
__global__ void __launch_bounds__(kFinThreads)
colsum_finalize_kernel(const float* __restrict__ partials, int64_t nblk, int C, float* out) {
    const int c = blockIdx.x * kFinCh + (threadIdx.x % kFinCh);
    double s[SETS];
    reduce_partials<SETS>(partials, nblk, C, c, s);
    if (threadIdx.x < kFinCh && c < C) {
#pragma unroll
        for (int k = 0; k < SETS; ++k) out[k * C + c] = (float)s[k];
    }
}

// Row-block elementwise kernels.  CTA = 256 threads walks rows [b*rpb, (b+1)*rpb); a thread owns one float4 of
// channels (cv) and every RL-th row, RL = 256/(C/4) when C/4 <= 256.  MODE 0: BN-backward reduce (2 sets),
// MODE 1: BN-backward apply (+ 1 set: column sum of dY), MODE 2: plain column sum (1 set).
// 5 CTAs per SM (<= 48 registers; the unconstrained build takes 56 / 60 and fits 4): the kernels are pure streams, more
// rows in flight is all that counts -- 15.7 -> 14.9 ms per step at 1M faces (scripts/records/gpu_r2_m.sh; 6 CTAs = 40 registers
// spills in the loop).  One CTA also fits beside a resident tcgen05 GEMM CTA of the other network's stream.
#ifndef DDMP_ROWBLOCK_MINB
#define DDMP_ROWBLOCK_MINB 5
#endif
template <int MODE>
__global__ void __launch_bounds__(256, DDMP_ROWBLOCK_MINB)
rowblock_kernel(const float* __restrict__ gX, const float* __restrict__ Y, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                float slope, const float* __restrict__ c1, const float* __restrict__ c2, float* __restrict__ dY,
                float* __restrict__ partials, int64_t n, int C, int rows_per_block) {
    constexpr int SETS = (MODE == 0) ? 2 : 1;
    extern __shared__ float red[];   // [RL][SETS][C]
    const int CV = C / 4;
    const int lanes_c = CV < 256 ? CV : 256;       // threads across channels
    const int RL = 256 / lanes_c;                  // row lanes
    const int tc = threadIdx.x % lanes_c, tr = threadIdx.x / lanes_c;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t row_end = (row0 + rows_per_block < n) ? (row0 + rows_per_block) : n;

    for (int cv = tc; cv < CV; cv += lanes_c) {
        const int c = cv * 4;
        float4 mu = make_float4(0, 0, 0, 0), rs = mu, sc = mu, sh = mu, k1 = mu, k2 = mu;
        if (MODE != 2) {
            mu = ldg4(mean + c); rs = ldg4(rstd + c); sc = ldg4(scale + c); sh = ldg4(shift + c);
        }
        if (MODE == 1) { k1 = ldg4(c1 + c); k2 = ldg4(c2 + c); }
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
        if (tr < RL) {
            for (int64_t r = row0 + tr; r < row_end; r += RL) {
                const float4 g = ldg4(gX + r * C + c);
                if (MODE == 2) {
                    a0.x += g.x; a0.y += g.y; a0.z += g.z; a0.w += g.w;
                    continue;
                }
                const float4 y = ldg4(Y + r * C + c);
                float4 gz, xh;
                gz.x = (fmaf(y.x, sc.x, sh.x) > 0.f) ? g.x : g.x * slope;
                gz.y = (fmaf(y.y, sc.y, sh.y) > 0.f) ? g.y : g.y * slope;
                gz.z = (fmaf(y.z, sc.z, sh.z) > 0.f) ? g.z : g.z * slope;
                gz.w = (fmaf(y.w, sc.w, sh.w) > 0.f) ? g.w : g.w * slope;
                xh.x = (y.x - mu.x) * rs.x; xh.y = (y.y - mu.y) * rs.y;
                xh.z = (y.z - mu.z) * rs.z; xh.w = (y.w - mu.w) * rs.w;
                if (MODE == 0) {
                    a0.x += gz.x; a0.y += gz.y; a0.z += gz.z; a0.w += gz.w;
                    a1.x = fmaf(gz.x, xh.x, a1.x); a1.y = fmaf(gz.y, xh.y, a1.y);
                    a1.z = fmaf(gz.z, xh.z, a1.z); a1.w = fmaf(gz.w, xh.w, a1.w);
                } else {
                    float4 d;
                    d.x = sc.x * (gz.x - k1.x - xh.x * k2.x);
                    d.y = sc.y * (gz.y - k1.y - xh.y * k2.y);
                    d.z = sc.z * (gz.z - k1.z - xh.z * k2.z);
                    d.w = sc.w * (gz.w - k1.w - xh.w * k2.w);
                    st4(dY + r * C + c, d);
                    a0.x += d.x; a0.y += d.y; a0.z += d.z; a0.w += d.w;
                }
            }
        }
        if (partials && tr < RL) {
            st4(red + (tr * SETS + 0) * C + c, a0);
            if (SETS == 2) st4(red + (tr * SETS + 1) * C + c, a1);
        }
    }
    if (!partials) return;
    __syncthreads();
    float* outp = partials + (int64_t)blockIdx.x * SETS * C;
    for (int i = threadIdx.x; i < SETS * C; i += 256) {
        float t = 0.f;
        for (int r = 0; r < RL; ++r) t += red[r * SETS * C + i];
        outp[i] = t;
    }
}

// dst[i,:] = src[idx[i],:]  (halo packing for the partitioned mode; one float4 per thread)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst,
                                   int64_t m, int C4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m * C4) return;
    const int64_t r = i / C4;
    const int c = (int)(i % C4);
    const int64_t s = __ldg(idx + r);
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + s * C4 + c);
}

template <int MODE>
static int launch_rowblock(const float* gX, const float* Y, const float* mean, const float* rstd, const float* scale,
                           const float* shift, float slope, const float* c1, const float* c2, float* dY,
                           float* partials, int64_t n, int C, cudaStream_t st, const char* what) {
    DDMP_REQUIRE(C % 4 == 0 && C >= 4, "%s: C must be a multiple of 4 (got %d)", what, C);
    constexpr int SETS = (MODE == 0) ? 2 : 1;
    const int CV = C / 4;
    const int lanes_c = CV < 256 ? CV : 256;
    const int RL = 256 / lanes_c;
    const size_t smem = (size_t)RL * SETS * C * sizeof(float);
    DDMP_REQUIRE(smem <= 48 * 1024, "%s: C=%d too wide", what, C);
    const int rpb = ddmp_elem_rows_per_block(C);
    rowblock_kernel<MODE><<<(unsigned)ceil_div(n, rpb), 256, smem, st>>>(gX, Y, mean, rstd, scale, shift, slope, c1,
                                                                         c2, dY, partials, n, C, rpb);
    return check_launch(what);
}

}  // namespace ddmp

extern "C" {

int ddmp_bn_stats_finalize(const float* partials, int64_t nblk, int64_t n, int32_t C, const float* gamma,
                           const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                           float* mean, float* rstd, float* scale, float* shift, float* bound, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(partials && gamma && beta && mean && rstd && scale && shift, "bn_stats_finalize: null pointer");
    DDMP_REQUIRE(n > 0 && C > 0 && nblk > 0, "bn_stats_finalize: bad shape");
    const int rpb = ddmp_rows_per_block(C);
    DDMP_REQUIRE(nblk == ceil_div(n, rpb), "bn_stats_finalize: %lld partial blocks do not cover %lld rows in blocks of %d",
                 (long long)nblk, (long long)n, rpb);
    bn_stats_finalize_kernel<<<(unsigned)ceil_div(C, kFinCh), kFinThreads, 0, as_stream(stream)>>>(
        partials, nblk, n, C, rpb, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift, bound);
    return check_launch("bn_stats_finalize");
}

int ddmp_bn_stats_rank_sums(const float* partials, int64_t nblk, int64_t n, int32_t C, double* sums, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(partials && sums && n > 0 && C > 0 && nblk > 0, "bn_stats_rank_sums: bad arguments");
    const int rpb = ddmp_rows_per_block(C);
    DDMP_REQUIRE(nblk == ceil_div(n, rpb), "bn_stats_rank_sums: partial blocks do not cover the rows");
    bn_stats_rank_sums_kernel<<<(unsigned)ceil_div(C, kFinCh), kFinThreads, 0, as_stream(stream)>>>(partials, nblk, n, C,
                                                                                                  rpb, sums);
    return check_launch("bn_stats_rank_sums");
}

int ddmp_bn_stats_finalize_sums(const double* sums, int64_t n, int32_t C, const float* gamma, const float* beta,
                                float eps, float momentum, float* running_mean, float* running_var, float* mean,
                                float* rstd, float* scale, float* shift, float* bound, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(sums && gamma && beta && mean && rstd && scale && shift && n > 0 && C > 0,
                 "bn_stats_finalize_sums: bad arguments");
    bn_stats_from_sums_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, as_stream(stream)>>>(
        sums, n, C, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift, bound);
    return check_launch("bn_stats_finalize_sums");
}

int ddmp_bn_eval_stats(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                       float eps, int32_t C, float* mean, float* rstd, float* scale, float* shift, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(running_mean && running_var && gamma && beta && mean && rstd && scale && shift && C > 0,
                 "bn_eval_stats: bad arguments");
    bn_eval_stats_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, as_stream(stream)>>>(running_mean, running_var, gamma, beta,
                                                                                  eps, C, mean, rstd, scale, shift);
    return check_launch("bn_eval_stats");
}

int ddmp_bn_bwd_reduce(const float* gX, const float* Y, const float* mean, const float* rstd, const float* scale,
                       const float* shift, float slope, float* partials, int64_t n, int32_t C, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(gX && Y && mean && rstd && scale && shift && partials, "bn_bwd_reduce: null pointer");
    if (n == 0) return DDMP_OK;
    return launch_rowblock<0>(gX, Y, mean, rstd, scale, shift, slope, nullptr, nullptr, nullptr, partials, n, C,
                              as_stream(stream), "bn_bwd_reduce");
}

int ddmp_bn_bwd_finalize(const float* partials, int64_t nblk, int64_t n, int32_t C, float* dgamma, float* dbeta,
                         float* c1, float* c2, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(partials && dgamma && dbeta && c1 && c2, "bn_bwd_finalize: null pointer");
    DDMP_REQUIRE(n > 0 && C > 0 && nblk > 0, "bn_bwd_finalize: bad shape");
    bn_bwd_finalize_kernel<<<(unsigned)ceil_div(C, kFinCh), kFinThreads, 0, as_stream(stream)>>>(partials, nblk, n, C, dgamma,
                                                                                    dbeta, c1, c2);
    return check_launch("bn_bwd_finalize");
}

int ddmp_bn_bwd_apply(const float* gX, const float* Y, const float* mean, const float* rstd, const float* scale,
                      const float* shift, float slope, const float* c1, const float* c2, float* dY,
                      float* colsum_partials, int64_t n, int32_t C, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(gX && Y && mean && rstd && scale && shift && c1 && c2 && dY, "bn_bwd_apply: null pointer");
    if (n == 0) return DDMP_OK;
    return launch_rowblock<1>(gX, Y, mean, rstd, scale, shift, slope, c1, c2, dY, colsum_partials, n, C,
                              as_stream(stream), "bn_bwd_apply");
}

int ddmp_gather_rows(const float* src, const int32_t* idx, float* dst, int64_t m, int32_t C, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(src && idx && dst && m >= 0 && C > 0 && C % 4 == 0, "gather_rows: bad arguments (C must be a multiple of 4)");
    if (m == 0) return DDMP_OK;
    const int64_t total = m * (C / 4);
    gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(src, idx, dst, m, C / 4);
    return check_launch("gather_rows");
}

int ddmp_colsum_partials(const float* X, float* partials, int64_t n, int32_t C, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(X && partials, "colsum_partials: null pointer");
    if (n == 0) return DDMP_OK;
    return launch_rowblock<2>(X, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr,
                              partials, n, C, as_stream(stream), "colsum_partials");
}

int ddmp_colsum_finalize(const float* partials, int64_t nblk, int32_t sets, int32_t C, float* out, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(partials && out && nblk > 0 && C > 0, "colsum_finalize: bad arguments");
    const unsigned grid = (unsigned)ceil_div(C, kFinCh);
    cudaStream_t st = as_stream(stream);
    if (sets == 1) colsum_finalize_kernel<1><<<grid, kFinThreads, 0, st>>>(partials, nblk, C, out);
    else if (sets == 2) colsum_finalize_kernel<2><<<grid, kFinThreads, 0, st>>>(partials, nblk, C, out);
    else { set_error("colsum_finalize: sets must be 1 or 2"); return DDMP_ERR_INVALID; }
    return check_launch("colsum_finalize");
}

}  // extern "C"
