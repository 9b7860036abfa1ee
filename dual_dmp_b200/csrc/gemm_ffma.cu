// Dense feature transform, FP32 FFMA path (DDMP_GEMM_FFMA).
//   xw : H[n,Cout]    = act(X)[n,Cin] * W[Cout,Cin]^T      (GCNConv.lin, reference util/networks.py:51-62)
//   dx : gX[n,Cin]    = dH[n,Cout] * W[Cout,Cin]
//   dw : dW[Cout,Cin] = dH[n,Cout]^T * act(X)[n,Cin]        (split-K over rows, fixed-order second stage)
// One register-tiled SIMT kernel serves all three through operand layout flags.  It is the path for channel
// widths below 64 (memory-bound there) and the correctness baseline for the tcgen05 path (gemm_tc.cu).
// "act" = the previous layer's BatchNorm + LeakyReLU applied while the operand is loaded, so activations are
// never materialised (SURVEY.md §7 hard part 2).
#include "common.cuh"

namespace ddmp {

struct GemmArgs {
    const float* A;      // A_KC: A[m*lda + k] (row m optionally through a_map) ; else A[k*lda + m]
    const float* B;      // B_KC: B[n*ldb + k] ; else B[k*ldb + n] (row k optionally through b_map)
    float* C;            // C[m*ldc + n] (+ z*M*N for split-K partials)
    const int* a_map;    // gather map over m (A_KC only)
    const int* b_map;    // gather map over k (!B_KC only)
    const float* scale;  // act: over k for the A operand (xw) or over n for the B operand (dw)
    const float* shift;
    float slope;
    int64_t M, N, K;
    int64_t lda, ldb, ldc;
    int64_t kchunk;      // K range per blockIdx.y (split-K); == K when not split
    int tiles_n;
};

template <int BM, int BN, int BK, int TM, int TN, bool A_KC, bool B_KC, bool ACT_A, bool ACT_B, bool VEC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const GemmArgs g) {
    constexpr int T = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    constexpr int MR = TM / 4, NR = TN / 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int t = threadIdx.x;
    const int tn = blockIdx.x % g.tiles_n;
    const int64_t tm = blockIdx.x / g.tiles_n;
    const int64_t m0 = tm * BM;
    const int n0 = tn * BN;
    const int64_t kbeg = (int64_t)blockIdx.y * g.kchunk;
    const int64_t kend = (kbeg + g.kchunk < g.K) ? (kbeg + g.kchunk) : g.K;
    const int tx = t % (BN / TN), ty = t / (BN / TN);

    // per-thread staging registers for one k-tile
    constexpr int A_V = (BM * BK / 4 + T - 1) / T;   // float4 slots per thread
    constexpr int B_V = (BN * BK / 4 + T - 1) / T;
    float4 ra[A_V], rb[B_V];

    auto act = [&](float x, int64_t ch) -> float {
        const float z = fmaf(x, __ldg(g.scale + ch), __ldg(g.shift + ch));
        return z > 0.f ? z : z * g.slope;
    };

    auto load_tiles = [&](int64_t k0) {
        // ---- A tile ----
#pragma unroll
        for (int i = 0; i < A_V; ++i) {
            const int idx = t + i * T;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < BM * BK / 4) {
                if (A_KC) {
                    const int row = idx / (BK / 4), kv = idx % (BK / 4);
                    const int64_t m = m0 + row, k = k0 + kv * 4;
                    if (m < g.M && k < kend) {
                        const int64_t src = g.a_map ? (int64_t)__ldg(g.a_map + m) : m;
                        const float* p = g.A + src * g.lda + k;
                        if (VEC && k + 3 < kend) {
                            v = ldg4(p);
                        } else {
                            v.x = __ldg(p);
                            if (k + 1 < kend) v.y = __ldg(p + 1);
                            if (k + 2 < kend) v.z = __ldg(p + 2);
                            if (k + 3 < kend) v.w = __ldg(p + 3);
                        }
                        if (ACT_A && g.scale) {
                            v.x = act(v.x, k);
                            if (k + 1 < kend) v.y = act(v.y, k + 1);
                            if (k + 2 < kend) v.z = act(v.z, k + 2);
                            if (k + 3 < kend) v.w = act(v.w, k + 3);
                        }
                    }
                } else {
                    const int kr = idx / (BM / 4), mv = idx % (BM / 4);
                    const int64_t k = k0 + kr, m = m0 + mv * 4;
                    if (k < kend && m < g.M) {
                        const float* p = g.A + k * g.lda + m;
                        if (VEC && m + 3 < g.M) {
                            v = ldg4(p);
                        } else {
                            v.x = __ldg(p);
                            if (m + 1 < g.M) v.y = __ldg(p + 1);
                            if (m + 2 < g.M) v.z = __ldg(p + 2);
                            if (m + 3 < g.M) v.w = __ldg(p + 3);
                        }
                    }
                }
            }
            ra[i] = v;
        }
        // ---- B tile ----
#pragma unroll
        for (int i = 0; i < B_V; ++i) {
            const int idx = t + i * T;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < BN * BK / 4) {
                if (B_KC) {
                    const int col = idx / (BK / 4), kv = idx % (BK / 4);
                    const int64_t n = n0 + col, k = k0 + kv * 4;
                    if (n < g.N && k < kend) {
                        const float* p = g.B + n * g.ldb + k;
                        if (VEC && k + 3 < kend) {
                            v = ldg4(p);
                        } else {
                            v.x = __ldg(p);
                            if (k + 1 < kend) v.y = __ldg(p + 1);
                            if (k + 2 < kend) v.z = __ldg(p + 2);
                            if (k + 3 < kend) v.w = __ldg(p + 3);
                        }
                    }
                } else {
                    const int kr = idx / (BN / 4), nv = idx % (BN / 4);
                    const int64_t k = k0 + kr, n = n0 + nv * 4;
                    if (k < kend && n < g.N) {
                        const int64_t src = g.b_map ? (int64_t)__ldg(g.b_map + k) : k;
                        const float* p = g.B + src * g.ldb + n;
                        if (VEC && n + 3 < g.N) {
                            v = ldg4(p);
                        } else {
                            v.x = __ldg(p);
                            if (n + 1 < g.N) v.y = __ldg(p + 1);
                            if (n + 2 < g.N) v.z = __ldg(p + 2);
                            if (n + 3 < g.N) v.w = __ldg(p + 3);
                        }
                        if (ACT_B && g.scale) {
                            v.x = act(v.x, n);
                            if (n + 1 < g.N) v.y = act(v.y, n + 1);
                            if (n + 2 < g.N) v.z = act(v.z, n + 2);
                            if (n + 3 < g.N) v.w = act(v.w, n + 3);
                        }
                    }
                }
            }
            rb[i] = v;
        }
    };

    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_V; ++i) {
            const int idx = t + i * T;
            if (idx < BM * BK / 4) {
                if (A_KC) {
                    const int row = idx / (BK / 4), kv = idx % (BK / 4);
                    As[buf][kv * 4 + 0][row] = ra[i].x;
                    As[buf][kv * 4 + 1][row] = ra[i].y;
                    As[buf][kv * 4 + 2][row] = ra[i].z;
                    As[buf][kv * 4 + 3][row] = ra[i].w;
                } else {
                    const int kr = idx / (BM / 4), mv = idx % (BM / 4);
                    *reinterpret_cast<float4*>(&As[buf][kr][mv * 4]) = ra[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < B_V; ++i) {
            const int idx = t + i * T;
            if (idx < BN * BK / 4) {
                if (B_KC) {
                    const int col = idx / (BK / 4), kv = idx % (BK / 4);
                    Bs[buf][kv * 4 + 0][col] = rb[i].x;
                    Bs[buf][kv * 4 + 1][col] = rb[i].y;
                    Bs[buf][kv * 4 + 2][col] = rb[i].z;
                    Bs[buf][kv * 4 + 3][col] = rb[i].w;
                } else {
                    const int kr = idx / (BN / 4), nv = idx % (BN / 4);
                    *reinterpret_cast<float4*>(&Bs[buf][kr][nv * 4]) = rb[i];
                }
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int64_t nk = (kend > kbeg) ? ((kend - kbeg + BK - 1) / BK) : 0;
    if (nk > 0) {
        load_tiles(kbeg);
        store_tiles(0);
    }
    __syncthreads();
    for (int64_t kt = 0; kt < nk; ++kt) {
        const int buf = (int)(kt & 1);
        if (kt + 1 < nk) load_tiles(kbeg + (kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < MR; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][i * (BM / MR) + ty * 4]);
                a[i * 4 + 0] = v.x; a[i * 4 + 1] = v.y; a[i * 4 + 2] = v.z; a[i * 4 + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][j * (BN / NR) + tx * 4]);
                b[j * 4 + 0] = v.x; b[j * 4 + 1] = v.y; b[j * 4 + 2] = v.z; b[j * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    float* Cz = g.C + (int64_t)blockIdx.y * g.M * g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + (i / 4) * (BM / MR) + ty * 4 + (i % 4);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const int64_t n = n0 + j * (BN / NR) + tx * 4;
            float* p = Cz + m * g.ldc + n;
            if (VEC && n + 3 < g.N) {
                st4(p, make_float4(acc[i][j * 4 + 0], acc[i][j * 4 + 1], acc[i][j * 4 + 2], acc[i][j * 4 + 3]));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (n + e < g.N) p[e] = acc[i][j * 4 + e];
            }
        }
    }
}

// second stage of split-K: out[i] = sum_z partials[z][i], z ascending (deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out, int64_t count,
                                     int splits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partials[(int64_t)z * count + i];
    out[i] = s;
}

// The same sum for SMALL outputs (narrow layers: 64 ... 2,048 numbers, ~600 partials): the kernel above would run 1-8 CTAs
// whose threads each walk the ~600 partials in one dependent chain (~70 us -- most of what ddmp_gemm_dw took on the narrow
// layers).  Here a CTA owns 32 outputs, 8 z-lanes per output add every 8th partial in ascending order (independent
// loads), and the 8 lane sums are folded in lane order: fixed order, deterministic.
__global__ void __launch_bounds__(256)
splitk_reduce_small_kernel(const float* __restrict__ partials, float* __restrict__ out, int64_t count, int splits) {
    __shared__ float red[8][33];
    const int o = threadIdx.x & 31, zl = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + o;
    float s = 0.f;
    if (i < count) {
#pragma unroll 4
        for (int z = zl; z < splits; z += 8) s += __ldg(partials + (int64_t)z * count + i);
    }
    red[zl][o] = s;
    __syncthreads();
    if (zl == 0 && i < count) {
        float v = red[0][o];
#pragma unroll
        for (int z = 1; z < 8; ++z) v += red[z][o];
        out[i] = v;
    }
}

static int launch_splitk_reduce(const float* partials, float* out, int64_t count, int splits, cudaStream_t st) {
    if (count <= 16384 && splits >= 32)
        splitk_reduce_small_kernel<<<(unsigned)ceil_div(count, 32), 256, 0, st>>>(partials, out, count, splits);
    else
        splitk_reduce_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(partials, out, count, splits);
    return check_launch("splitk_reduce");
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int BM, int BN, int TM, int TN, bool A_KC, bool B_KC, bool ACT_A, bool ACT_B>
static int launch_cfg(const GemmArgs& g, int splits, bool vec, cudaStream_t st) {
    constexpr int BK = 16;
    constexpr int T = (BM / TM) * (BN / TN);
    const int64_t tiles_m = ceil_div(g.M, BM);
    const int64_t tiles = tiles_m * g.tiles_n;
    DDMP_REQUIRE(tiles < (1ll << 31), "gemm: too many tiles");
    dim3 grid((unsigned)tiles, (unsigned)splits, 1);
    if (vec) sgemm_kernel<BM, BN, BK, TM, TN, A_KC, B_KC, ACT_A, ACT_B, true><<<grid, T, 0, st>>>(g);
    else sgemm_kernel<BM, BN, BK, TM, TN, A_KC, B_KC, ACT_A, ACT_B, false><<<grid, T, 0, st>>>(g);
    return check_launch("sgemm");
}

// tile choice by output width N (M is the huge dimension for xw/dx)
template <bool A_KC, bool B_KC, bool ACT_A, bool ACT_B>
static int launch_by_shape(GemmArgs g, int splits, bool vec, cudaStream_t st) {
    if (g.N > 64 && g.M > 64) {
        g.tiles_n = (int)ceil_div(g.N, 128);
        return launch_cfg<128, 128, 8, 8, A_KC, B_KC, ACT_A, ACT_B>(g, splits, vec, st);
    } else if (g.N > 32) {
        g.tiles_n = (int)ceil_div(g.N, 64);
        if (g.M > 64) return launch_cfg<128, 64, 8, 4, A_KC, B_KC, ACT_A, ACT_B>(g, splits, vec, st);
        return launch_cfg<64, 64, 4, 4, A_KC, B_KC, ACT_A, ACT_B>(g, splits, vec, st);
    } else {
        g.tiles_n = (int)ceil_div(g.N, 32);
        if (g.M > 32) return launch_cfg<128, 32, 4, 4, A_KC, B_KC, ACT_A, ACT_B>(g, splits, vec, st);
        return launch_cfg<32, 32, 4, 4, A_KC, B_KC, ACT_A, ACT_B>(g, splits, vec, st);
    }
}

int ffma_gemm_xw(const float* X, const int32_t* row_map, const float* scale, const float* shift, float slope,
                 const float* W, float* H, int64_t n, int32_t Cin, int32_t Cout, cudaStream_t st) {
    GemmArgs g{};
    g.A = X; g.B = W; g.C = H; g.a_map = row_map; g.b_map = nullptr;
    g.scale = scale; g.shift = shift; g.slope = slope;
    g.M = n; g.N = Cout; g.K = Cin; g.lda = Cin; g.ldb = Cin; g.ldc = Cout; g.kchunk = Cin;
    const bool vec = (Cin % 4 == 0) && (Cout % 4 == 0) && aligned16(X) && aligned16(W) && aligned16(H);
    return launch_by_shape<true, true, true, false>(g, 1, vec, st);
}

int ffma_gemm_dx(const float* dH, const float* W, float* gX, int64_t n, int32_t Cin, int32_t Cout, cudaStream_t st) {
    GemmArgs g{};
    g.A = dH; g.B = W; g.C = gX; g.a_map = nullptr; g.b_map = nullptr;
    g.scale = nullptr; g.shift = nullptr; g.slope = 0.f;
    g.M = n; g.N = Cin; g.K = Cout; g.lda = Cout; g.ldb = Cin; g.ldc = Cin; g.kchunk = Cout;
    const bool vec = (Cin % 4 == 0) && (Cout % 4 == 0) && aligned16(dH) && aligned16(W) && aligned16(gX);
    return launch_by_shape<true, false, false, false>(g, 1, vec, st);
}

// ---- dW of the narrow layers (both widths <= 64) --------------------------------------------------------------------
// dW[M = Cout, N = Cin] = sum over rows of dH[r, :]^T * act(X)[r, :] is a reduction over ~1M rows into at most 64 x 64
// numbers.  The general kernel above gives every output a thread of ONE tile: at 16 x 32 outputs that is a 64-thread CTA
// (4.4 TFLOP/s, 0.23 ms for 0.19 GB of operands at 1M rows), at 32 x 64 half of a 64 x 64 tile idles.  Here a CTA of 256
// threads always works: G = (M/4)(N/4) threads hold the output in 4 x 4 register tiles and the remaining factor
// KG = 256 / G splits the ROWS of a stage between KG such groups (split-K inside the CTA, groups folded in group order
// through shared memory at the end).  Stages of RS rows are staged through shared memory (coalesced float4 loads, the
// BatchNorm + LeakyReLU of the X operand applied once per element on the way, double-buffered through registers), so the
// inner loop is 2 LDS.128 + 16 FFMA per row and thread.  One partial per CTA, summed in CTA order by splitk_reduce_kernel.
// rows per stage: the largest power of two whose two stage buffers fit DDMP_DWN_STAGE_BYTES (12 KB per stage: 64 rows at
// 32 x 16, 32 at 64 x 32).  24 KB stages need 80-89 registers (3 CTAs per SM) or spill at 64 and measured the same
// (scripts/bench_gemm_small.py: 0.073-0.186 vs 0.077-0.176 ms); what is left at 16 x 32 outputs is the one-stage-deep
// pipeline (latency per stage), at 64 x 32 the FFMA issue rate (~40 % of peak).
#ifndef DDMP_DWN_STAGE_BYTES
#define DDMP_DWN_STAGE_BYTES 12288
#endif
constexpr int dw_narrow_stage_rows(int M, int N, int KG) {
    int rs = 32;
    while (2 * rs * (M + N) * 4 <= DDMP_DWN_STAGE_BYTES) rs *= 2;
    return rs > KG ? rs : KG;
}
template <int M, int N>
__global__ void __launch_bounds__(256, 4)     // 4 CTAs per SM = the 592 splits of dw_plan in one wave
dw_narrow_kernel(const float* __restrict__ dH, const float* __restrict__ X, const int* __restrict__ x_map,
                 const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                 float* __restrict__ partials, int64_t n, int64_t rows_per_cta) {
    constexpr int G = (M / 4) * (N / 4);
    constexpr int KG = 256 / G;
    constexpr int RS = dw_narrow_stage_rows(M, N, KG);    // rows per stage
    constexpr int A_F4 = RS * M / 4, B_F4 = RS * N / 4;   // float4 per stage
    constexpr int A_V = (A_F4 + 255) / 256, B_V = (B_F4 + 255) / 256;
    static_assert(256 % G == 0 && RS % KG == 0, "thread layout");
    static_assert(256 % (M / 4) == 0 && 256 % (N / 4) == 0, "a thread stages a fixed channel quad");
    constexpr int STAGE = RS * (M + N);
    constexpr int RED = 256 * 16;
    __shared__ __align__(16) float sm[(2 * STAGE > RED) ? 2 * STAGE : RED];
    const int t = threadIdx.x;
    const int kg = t / G, gi = t % G;
    const int tm = gi / (N / 4), tn = gi % (N / 4);
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = (r_begin + rows_per_cta < n) ? (r_begin + rows_per_cta) : n;

    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_act = scale != nullptr;
    if (has_act) { sc = ldg4(scale + (t % (N / 4)) * 4); sh = ldg4(shift + (t % (N / 4)) * 4); }
    float4 ra[A_V], rb[B_V];
    auto load_stage = [&](int64_t r0) {
#pragma unroll
        for (int i = 0; i < A_V; ++i) {
            const int idx = t + i * 256;
            const int64_t r = r0 + idx / (M / 4);
            ra[i] = (idx < A_F4 && r < r_end) ? ldg4(dH + r * M + (idx % (M / 4)) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < B_V; ++i) {
            const int idx = t + i * 256;
            const int64_t r = r0 + idx / (N / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < B_F4 && r < r_end) {
                const int64_t src = x_map ? (int64_t)__ldg(x_map + r) : r;
                v = ldg4(X + src * N + (idx % (N / 4)) * 4);
                if (has_act) {
                    float z;
                    z = fmaf(v.x, sc.x, sh.x); v.x = z > 0.f ? z : z * slope;
                    z = fmaf(v.y, sc.y, sh.y); v.y = z > 0.f ? z : z * slope;
                    z = fmaf(v.z, sc.z, sh.z); v.z = z > 0.f ? z : z * slope;
                    z = fmaf(v.w, sc.w, sh.w); v.w = z > 0.f ? z : z * slope;
                }
            }
            rb[i] = v;
        }
    };
    auto store_stage = [&](int buf) {
        float* As = sm + buf * STAGE;
        float* Bs = As + RS * M;
#pragma unroll
        for (int i = 0; i < A_V; ++i) {
            const int idx = t + i * 256;
            if (idx < A_F4) st4(As + idx * 4, ra[i]);
        }
#pragma unroll
        for (int i = 0; i < B_V; ++i) {
            const int idx = t + i * 256;
            if (idx < B_F4) st4(Bs + idx * 4, rb[i]);
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int64_t ns = (r_end > r_begin) ? ((r_end - r_begin + RS - 1) / RS) : 0;
    if (ns > 0) {
        load_stage(r_begin);
        store_stage(0);
    }
    __syncthreads();
    for (int64_t s = 0; s < ns; ++s) {
        const int buf = (int)(s & 1);
        if (s + 1 < ns) load_stage(r_begin + (s + 1) * RS);
        const float* As = sm + buf * STAGE;
        const float* Bs = As + RS * M;
#pragma unroll 4
        for (int k = kg; k < RS; k += KG) {
            const float4 a = *reinterpret_cast<const float4*>(As + k * M + tm * 4);
            const float4 b = *reinterpret_cast<const float4*>(Bs + k * N + tn * 4);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
            acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
            acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
            acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
            acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
        }
        if (s + 1 < ns) store_stage(buf ^ 1);
        __syncthreads();
    }
    // fold the KG row groups in group order: red[kg][m][n]
    float* red = sm;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        st4(red + (kg * M + tm * 4 + i) * N + tn * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    __syncthreads();
    float* outp = partials + (int64_t)blockIdx.x * M * N;
    for (int i = t; i < M * N; i += 256) {
        float v = 0.f;
#pragma unroll 4
        for (int z = 0; z < KG; ++z) v += red[z * M * N + i];
        outp[i] = v;
    }
}

template <int M, int N>
static int launch_dw_narrow(const float* dH, const float* X, const int32_t* row_map, const float* scale,
                            const float* shift, float slope, float* partials, int64_t n, int splits, cudaStream_t st) {
    constexpr int G = (M / 4) * (N / 4);
    constexpr int KG = 256 / G;
    constexpr int RS = dw_narrow_stage_rows(M, N, KG);
    int64_t rows = ceil_div(n, splits);
    rows = ceil_div(rows, RS) * RS;
    dw_narrow_kernel<M, N><<<(unsigned)splits, 256, 0, st>>>(dH, X, row_map, scale, shift, slope, partials, n, rows);
    return check_launch("dw_narrow");
}

// shapes of the Dual-DMP networks with both widths below 64 (Cout x Cin); 0 = not covered
static int dw_narrow_shape(int32_t Cin, int32_t Cout) {
    if (Cout == 32 && Cin == 16) return 1;
    if (Cout == 64 && Cin == 32) return 2;
    if (Cout == 32 && Cin == 64) return 3;
    if (Cout == 16 && Cin == 32) return 4;
    if (Cout == 4 && Cin == 16) return 5;
    return 0;
}

static void dw_plan(int64_t n, int32_t Cin, int32_t Cout, int* splits, int64_t* kchunk) {
    int bm, bn;
    if (Cin > 64 && Cout > 64) { bm = 128; bn = 128; }
    else if (Cin > 32) { bn = 64; bm = (Cout > 64) ? 128 : 64; }
    else { bn = 32; bm = (Cout > 32) ? 128 : 32; }
    const int64_t tiles = ceil_div(Cout, bm) * ceil_div(Cin, bn);
    int64_t s = ceil_div(4ll * kNumSMs, tiles);           // ~4 CTAs per SM in flight
    const int64_t max_by_rows = ceil_div(n, 256);          // at least 256 rows per split
    if (s > max_by_rows) s = max_by_rows;
    if (s < 1) s = 1;
    int64_t kc = ceil_div(n, s);
    kc = ceil_div(kc, 16) * 16;
    s = ceil_div(n, kc);
    if (s < 1) s = 1;
    *splits = (int)s;
    *kchunk = kc;
}

int64_t ffma_gemm_dw_workspace_bytes(int64_t n, int32_t Cin, int32_t Cout) {
    int s; int64_t kc;
    dw_plan(n, Cin, Cout, &s, &kc);
    return (int64_t)s * Cin * Cout * (int64_t)sizeof(float);
}

int ffma_gemm_dw(const float* dH, const float* X, const int32_t* row_map, const float* scale, const float* shift,
                 float slope, float* dW, void* workspace, int64_t workspace_bytes, int64_t n, int32_t Cin,
                 int32_t Cout, cudaStream_t st) {
    int splits; int64_t kc;
    dw_plan(n, Cin, Cout, &splits, &kc);
    DDMP_REQUIRE(workspace && workspace_bytes >= (int64_t)splits * Cin * Cout * (int64_t)sizeof(float),
                 "gemm_dw: workspace too small (%lld bytes)", (long long)workspace_bytes);
    const int narrow = dw_narrow_shape(Cin, Cout);
    if (narrow && splits > 1 && aligned16(dH) && aligned16(X) && aligned16(workspace)) {
        float* P = reinterpret_cast<float*>(workspace);
        int rc = DDMP_OK;
        switch (narrow) {
            case 1: rc = launch_dw_narrow<32, 16>(dH, X, row_map, scale, shift, slope, P, n, splits, st); break;
            case 2: rc = launch_dw_narrow<64, 32>(dH, X, row_map, scale, shift, slope, P, n, splits, st); break;
            case 3: rc = launch_dw_narrow<32, 64>(dH, X, row_map, scale, shift, slope, P, n, splits, st); break;
            case 4: rc = launch_dw_narrow<16, 32>(dH, X, row_map, scale, shift, slope, P, n, splits, st); break;
            default: rc = launch_dw_narrow<4, 16>(dH, X, row_map, scale, shift, slope, P, n, splits, st); break;
        }
        if (rc != DDMP_OK) return rc;
        return launch_splitk_reduce(P, dW, (int64_t)Cin * Cout, splits, st);
    }
    GemmArgs g{};
    g.A = dH; g.B = X; g.C = (splits == 1) ? dW : reinterpret_cast<float*>(workspace);
    g.a_map = nullptr; g.b_map = row_map;
    g.scale = scale; g.shift = shift; g.slope = slope;
    g.M = Cout; g.N = Cin; g.K = n; g.lda = Cout; g.ldb = Cin; g.ldc = Cin; g.kchunk = kc;
    const bool vec = (Cin % 4 == 0) && (Cout % 4 == 0) && aligned16(dH) && aligned16(X) && aligned16(g.C);
    int rc = launch_by_shape<false, false, false, true>(g, splits, vec, st);
    if (rc != DDMP_OK) return rc;
    if (splits > 1) {
        return launch_splitk_reduce(reinterpret_cast<const float*>(workspace), dW, (int64_t)Cin * Cout, splits, st);
    }
    return DDMP_OK;
}

}  // namespace ddmp
