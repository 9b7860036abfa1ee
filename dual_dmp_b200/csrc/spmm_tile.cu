// GCN aggregation, tile-staged variant: the row block's OWN rows of H arrive in shared memory through ONE TMA tensor
// copy (cp.async.bulk.tensor.2d, SASS UTMALDG) and its (rowptr, col, w) stream through coalesced loads into shared
// memory, so the gathers of a Morton-ordered mesh graph -- 85-92 % of whose neighbour references fall inside the
// block (measured on the 1M-face icosphere) -- are shared-memory reads; only the references that
// leave the block go to L1/L2.  Replaces torch_geometric GCNConv.propagate (index_select -> mul -> scatter_add,
// reference util/networks.py:51-62,112-123) like spmm.cu; same epilogues (bias, BatchNorm partial sums, max|Y|).
//
// Mapping: CTA = (row block of R rows) x (channel slice of CS <= 128 channels), 256 threads.  A group of G = CS/4
// lanes owns one output row at a time (one float4 per lane); the 256/G groups walk the block's rows.  No global
// dependent-load chain is left in the loop: rowptr / col / w come from shared memory, in-block rows from the tile.
//   C = 32, 64  : CS = C,   R = 256  (tile 32 / 64 KB)
//   C >= 128    : CS = 128, R = 128  (tile 64 KB), C/128 CTAs per row block (neighbouring block indices, so the
//                 slices of one row are fetched from DRAM at about the same time)
// Deterministic: no atomics; per-CTA BatchNorm partials are combined in group order.
// Roofline: HBM; algorithmic bytes 4*[(n+1) + 2*nnz + 2*n*C].
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ddmp {

namespace tile {

constexpr int kThreads = 256;
constexpr int kMetaPerRow = 9;          // (col, w) entries per row kept in shared memory (mesh graphs: <= 8 incl. loop)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tile of a row-major [rows, C] fp32 tensor: coordinates (c0 = first channel, c1 = first row)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

struct Args {
    const int* rowptr;
    const int* col;
    const float* w;
    const float* H;
    const float* bias;
    float* Y;
    float* partials;      // [nblk][2][C] or null
    float* amax;          // [nblk * slices] or null
    int64_t n;
    int C;
};

template <int CS, int R, bool STATS, bool BIAS>
__global__ void __launch_bounds__(kThreads)
spmm_tile_kernel(const __grid_constant__ CUtensorMap hmap, const Args a) {
    constexpr int G = CS / 4;
    constexpr int GROUPS = kThreads / G;
    constexpr int kMetaCap = R * kMetaPerRow;
    static_assert(GROUPS * 2 * CS * 4 <= R * CS * 4, "the statistics scratch aliases the tile");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);    // TMA destination: 128 B aligned
    float* tile = reinterpret_cast<float*>(base);                                       // [R][CS]
    float* red = tile;                                                                  // [GROUPS][2*CS], after the loop
    int* scol = reinterpret_cast<int*>(base + (size_t)R * CS * 4);                      // [kMetaCap]
    float* sw = reinterpret_cast<float*>(scol + kMetaCap);                              // [kMetaCap]
    int* srow = reinterpret_cast<int*>(sw + kMetaCap);                                  // [R + 1]
    uint64_t* barp = reinterpret_cast<uint64_t*>(srow + R + 2 + ((R & 1) ? 1 : 0));     // 8 B aligned
    float* wmax = reinterpret_cast<float*>(barp + 1);                                   // [kThreads / 32]
    uint64_t& bar = *barp;

    const int slices = a.C / CS;
    const int64_t blk = blockIdx.x / slices;
    const int ch0 = (blockIdx.x % slices) * CS;
    const int64_t row0 = blk * R;
    const int rows = (int)((row0 + R <= a.n) ? R : (a.n - row0));
    const int lane = threadIdx.x & 31;
    const int lg = threadIdx.x % G;
    const int gid = threadIdx.x / G;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (uint32_t)(R * CS * 4));            // rows past n are zero-filled and still counted
        tma_load_2d(tile, &hmap, ch0, (int)row0, &bar);
    }
    // (rowptr, col, w) of the block -> shared memory, while the tile is in flight
    for (int i = threadIdx.x; i <= rows; i += kThreads) srow[i] = __ldg(a.rowptr + row0 + i);
    __syncthreads();
    const int kb = srow[0];
    const int nk = srow[rows] - kb;
    const bool meta_in_smem = nk <= kMetaCap;
    if (meta_in_smem) {
        for (int i = threadIdx.x; i < nk; i += kThreads) {
            scol[i] = __ldg(a.col + kb + i);
            sw[i] = __ldg(a.w + kb + i);
        }
    }
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BIAS) bsum = ldg4(a.bias + ch0 + lg * 4);
    float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), qsum = make_float4(0.f, 0.f, 0.f, 0.f);     // running mean, M2
    float cnt = 0.f;
    float amx = 0.f;
    __syncthreads();
    mbar_wait(&bar, 0);

    const float* Hs = a.H + ch0 + lg * 4;
    for (int r = gid; r < rows; r += GROUPS) {
        const int s = srow[r] - kb, e = srow[r + 1] - kb;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = s;
        // four references at a time: the (rare) out-of-block gathers of a batch are in flight together
        for (; k + 4 <= e; k += 4) {
            int c[4];
            float wv[4];
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                c[u] = meta_in_smem ? scol[k + u] : __ldg(a.col + kb + k + u);
                wv[u] = meta_in_smem ? sw[k + u] : __ldg(a.w + kb + k + u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t lc = (int64_t)c[u] - row0;
                if ((uint64_t)lc < (uint64_t)rows) x[u] = *reinterpret_cast<const float4*>(tile + (int)lc * CS + lg * 4);
                else x[u] = ldg4(Hs + (int64_t)c[u] * a.C);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x = fmaf(wv[u], x[u].x, acc.x); acc.y = fmaf(wv[u], x[u].y, acc.y);
                acc.z = fmaf(wv[u], x[u].z, acc.z); acc.w = fmaf(wv[u], x[u].w, acc.w);
            }
        }
        for (; k < e; ++k) {
            const int c1 = meta_in_smem ? scol[k] : __ldg(a.col + kb + k);
            const float w1 = meta_in_smem ? sw[k] : __ldg(a.w + kb + k);
            const int64_t lc = (int64_t)c1 - row0;
            float4 x;
            if ((uint64_t)lc < (uint64_t)rows) x = *reinterpret_cast<const float4*>(tile + (int)lc * CS + lg * 4);
            else x = ldg4(Hs + (int64_t)c1 * a.C);
            acc.x = fmaf(w1, x.x, acc.x); acc.y = fmaf(w1, x.y, acc.y);
            acc.z = fmaf(w1, x.z, acc.z); acc.w = fmaf(w1, x.w, acc.w);
        }
        if (BIAS) { acc.x += bsum.x; acc.y += bsum.y; acc.z += bsum.z; acc.w += bsum.w; }
        st4(a.Y + (row0 + r) * a.C + ch0 + lg * 4, acc);
        amx = fmaxf(fmaxf(amx, fmaxf(fabsf(acc.x), fabsf(acc.y))), fmaxf(fabsf(acc.z), fabsf(acc.w)));
        if (STATS) {                                // Welford: running mean and M2 of this thread's rows (see bn.cu)
            cnt += 1.f;
            const float inv = 1.f / cnt;
            float d;
            d = acc.x - ssum.x; ssum.x = fmaf(d, inv, ssum.x); qsum.x = fmaf(d, acc.x - ssum.x, qsum.x);
            d = acc.y - ssum.y; ssum.y = fmaf(d, inv, ssum.y); qsum.y = fmaf(d, acc.y - ssum.y, qsum.y);
            d = acc.z - ssum.z; ssum.z = fmaf(d, inv, ssum.z); qsum.z = fmaf(d, acc.z - ssum.z, qsum.z);
            d = acc.w - ssum.w; ssum.w = fmaf(d, inv, ssum.w); qsum.w = fmaf(d, acc.w - ssum.w, qsum.w);
        }
    }

    if (STATS) {
        // per-group (mean, M2) -> shared memory (over the tile, which is dead now); then one thread per channel merges the
        // groups in group order (Chan et al.) and writes the block's (sum, M2 about the block mean)
        __syncthreads();
        st4(red + gid * 2 * CS + lg * 4, ssum);
        st4(red + gid * 2 * CS + CS + lg * 4, qsum);
        __syncthreads();
        float* outp = a.partials + blk * 2 * a.C;
        for (int ch = threadIdx.x; ch < CS; ch += kThreads) {
            float n_a = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll 4
            for (int g = 0; g < GROUPS; ++g) {
                const float n_b = (float)((rows - g + GROUPS - 1) / GROUPS);      // rows g, g+GROUPS, ... of this block
                if (g < rows) {
                    const float mb = red[g * 2 * CS + ch], qb = red[g * 2 * CS + CS + ch];
                    const float n = n_a + n_b;
                    const float d = mb - mean;
                    mean = fmaf(d, n_b / n, mean);
                    m2 += qb + d * d * (n_a * n_b / n);
                    n_a = n;
                }
            }
            outp[ch0 + ch] = mean * n_a;
            outp[a.C + ch0 + ch] = m2;
        }
    }
    if (a.amax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
        if (lane == 0) wmax[threadIdx.x >> 5] = amx;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = wmax[0];
#pragma unroll
            for (int i = 1; i < kThreads / 32; ++i) m = fmaxf(m, wmax[i]);
            a.amax[blockIdx.x] = m;
        }
    }
}

// Backward aggregation fused with the BatchNorm / LeakyReLU backward "apply" (replaces ddmp_bn_bwd_apply + ddmp_spmm_gcn
// on the backward pass; reference: autograd of nn.BatchNorm1d + nn.LeakyReLU + GCNConv.propagate,
// util/networks.py:51-62):
//     dY = scale * (gZ - c1 - xhat * c2),  gZ = gX * lrelu'(scale*Y + shift),  xhat = (Y - mean) * rstd
//     dH[i,:] = sum_k w[k] * dY[col[k],:]          colsum[c] = sum_i dY[i,c]   (conv-bias gradient)
// The block's own rows of gX and Y arrive by two TMA tensor copies; dY of those rows is formed ONCE in shared memory
// (in place over the gX tile) and the in-block gathers read it from there; a reference that leaves the block recomputes
// dY from the two global rows.  dY is never written to HBM: per layer the pass reads gX and Y and writes dH (3 tensor
// passes + halo) instead of 5 (apply: read gX, Y, write dY; aggregate: read dY, write dH).
struct BwdArgs {
    const int* rowptr;
    const int* col;
    const float* w;
    const float* gX;
    const float* Y;
    const float* mean;
    const float* rstd;
    const float* scale;
    const float* shift;
    const float* c1;
    const float* c2;
    float slope;
    float* dH;
    float* colsum;        // [nblk][C] or null
    float* amax;          // [nblk * slices] or null
    int64_t n;
    int C;
};

template <int CS, int R>
__global__ void __launch_bounds__(kThreads)
spmm_bn_bwd_tile_kernel(const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap ymap,
                        const BwdArgs a) {
    constexpr int G = CS / 4;
    constexpr int GROUPS = kThreads / G;
    constexpr int kMetaCap = R * kMetaPerRow;
    static_assert(GROUPS * CS * 4 <= R * CS * 4, "the column-sum scratch aliases the Y tile");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float* tg = reinterpret_cast<float*>(base);                                         // gX tile, then dY  [R][CS]
    float* ty = tg + R * CS;                                                            // Y tile, then scratch
    int* scol = reinterpret_cast<int*>(ty + R * CS);
    float* sw = reinterpret_cast<float*>(scol + kMetaCap);
    int* srow = reinterpret_cast<int*>(sw + kMetaCap);
    uint64_t* barp = reinterpret_cast<uint64_t*>(srow + R + 2 + ((R & 1) ? 1 : 0));
    float* wmax = reinterpret_cast<float*>(barp + 1);
    uint64_t& bar = *barp;

    const int slices = a.C / CS;
    const int64_t blk = blockIdx.x / slices;
    const int ch0 = (blockIdx.x % slices) * CS;
    const int64_t row0 = blk * R;
    const int rows = (int)((row0 + R <= a.n) ? R : (a.n - row0));
    const int lane = threadIdx.x & 31;
    const int lg = threadIdx.x % G;
    const int gid = threadIdx.x / G;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (uint32_t)(2 * R * CS * 4));
        tma_load_2d(tg, &gmap, ch0, (int)row0, &bar);
        tma_load_2d(ty, &ymap, ch0, (int)row0, &bar);
    }
    for (int i = threadIdx.x; i <= rows; i += kThreads) srow[i] = __ldg(a.rowptr + row0 + i);
    // per-lane channel constants (this lane's four channels of the slice)
    const int c = ch0 + lg * 4;
    const float4 mu = ldg4(a.mean + c), rs = ldg4(a.rstd + c), sc = ldg4(a.scale + c), sh = ldg4(a.shift + c);
    const float4 k1 = ldg4(a.c1 + c), k2 = ldg4(a.c2 + c);
    const float slope = a.slope;
    auto dy4 = [&](const float4& g, const float4& y) {          // same arithmetic as bn.cu rowblock_kernel<1>
        float4 d;
        const float gzx = (fmaf(y.x, sc.x, sh.x) > 0.f) ? g.x : g.x * slope;
        const float gzy = (fmaf(y.y, sc.y, sh.y) > 0.f) ? g.y : g.y * slope;
        const float gzz = (fmaf(y.z, sc.z, sh.z) > 0.f) ? g.z : g.z * slope;
        const float gzw = (fmaf(y.w, sc.w, sh.w) > 0.f) ? g.w : g.w * slope;
        d.x = sc.x * (gzx - k1.x - ((y.x - mu.x) * rs.x) * k2.x);
        d.y = sc.y * (gzy - k1.y - ((y.y - mu.y) * rs.y) * k2.y);
        d.z = sc.z * (gzz - k1.z - ((y.z - mu.z) * rs.z) * k2.z);
        d.w = sc.w * (gzw - k1.w - ((y.w - mu.w) * rs.w) * k2.w);
        return d;
    };
    __syncthreads();
    const int kb = srow[0];
    const int nk = srow[rows] - kb;
    const bool meta_in_smem = nk <= kMetaCap;
    if (meta_in_smem) {
        for (int i = threadIdx.x; i < nk; i += kThreads) {
            scol[i] = __ldg(a.col + kb + i);
            sw[i] = __ldg(a.w + kb + i);
        }
    }
    mbar_wait(&bar, 0);
    // dY of the block's own rows, in place over the gX tile; column sums of dY (rows of this block only)
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = gid; r < rows; r += GROUPS) {
        float4* pg = reinterpret_cast<float4*>(tg + r * CS + lg * 4);
        const float4 d = dy4(*pg, *reinterpret_cast<const float4*>(ty + r * CS + lg * 4));
        *pg = d;
        csum.x += d.x; csum.y += d.y; csum.z += d.z; csum.w += d.w;
    }
    __syncthreads();
    if (a.colsum) {
        float* red = ty;                                        // the Y tile is dead now
        st4(red + gid * CS + lg * 4, csum);
        __syncthreads();
        for (int i = threadIdx.x; i < CS; i += kThreads) {
            float t = 0.f;
#pragma unroll 8
            for (int g = 0; g < GROUPS; ++g) t += red[g * CS + i];
            a.colsum[blk * a.C + ch0 + i] = t;
        }
    }

    float amx = 0.f;
    const float* Gs = a.gX + c;
    const float* Ys = a.Y + c;
    for (int r = gid; r < rows; r += GROUPS) {
        const int s = srow[r] - kb, e = srow[r + 1] - kb;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = s;
        for (; k + 4 <= e; k += 4) {
            int cc[4];
            float wv[4];
            float4 x[4], y[4];
            bool in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                cc[u] = meta_in_smem ? scol[k + u] : __ldg(a.col + kb + k + u);
                wv[u] = meta_in_smem ? sw[k + u] : __ldg(a.w + kb + k + u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t lc = (int64_t)cc[u] - row0;
                in[u] = (uint64_t)lc < (uint64_t)rows;
                if (in[u]) x[u] = *reinterpret_cast<const float4*>(tg + (int)lc * CS + lg * 4);
                else { x[u] = ldg4(Gs + (int64_t)cc[u] * a.C); y[u] = ldg4(Ys + (int64_t)cc[u] * a.C); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 d = in[u] ? x[u] : dy4(x[u], y[u]);
                acc.x = fmaf(wv[u], d.x, acc.x); acc.y = fmaf(wv[u], d.y, acc.y);
                acc.z = fmaf(wv[u], d.z, acc.z); acc.w = fmaf(wv[u], d.w, acc.w);
            }
        }
        for (; k < e; ++k) {
            const int c1_ = meta_in_smem ? scol[k] : __ldg(a.col + kb + k);
            const float w1 = meta_in_smem ? sw[k] : __ldg(a.w + kb + k);
            const int64_t lc = (int64_t)c1_ - row0;
            float4 d;
            if ((uint64_t)lc < (uint64_t)rows) d = *reinterpret_cast<const float4*>(tg + (int)lc * CS + lg * 4);
            else d = dy4(ldg4(Gs + (int64_t)c1_ * a.C), ldg4(Ys + (int64_t)c1_ * a.C));
            acc.x = fmaf(w1, d.x, acc.x); acc.y = fmaf(w1, d.y, acc.y);
            acc.z = fmaf(w1, d.z, acc.z); acc.w = fmaf(w1, d.w, acc.w);
        }
        st4(a.dH + (row0 + r) * a.C + c, acc);
        amx = fmaxf(fmaxf(amx, fmaxf(fabsf(acc.x), fabsf(acc.y))), fmaxf(fabsf(acc.z), fabsf(acc.w)));
    }
    if (a.amax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
        if (lane == 0) wmax[threadIdx.x >> 5] = amx;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = wmax[0];
#pragma unroll
            for (int i = 1; i < kThreads / 32; ++i) m = fmaxf(m, wmax[i]);
            a.amax[blockIdx.x] = m;
        }
    }
}

template <int CS, int R>
constexpr size_t smem_bytes() {
    return 128 + (size_t)R * CS * 4 + (size_t)R * kMetaPerRow * 8 + (size_t)(R + 4) * 4 + 8 + (kThreads / 32) * 4 + 16;
}

static int make_map(CUtensorMap* map, const float* H, int64_t n, int C, int CS, int R) {
    return make_tensor_map_2d(map, H, n, C, R, CS, false, "spmm_tile");
}

template <int CS, int R>
static int launch(const Args& a, cudaStream_t st) {
    CUtensorMap map;
    int rc = make_map(&map, a.H, a.n, a.C, CS, R);
    if (rc != DDMP_OK) return rc;
    const int64_t nblk = ceil_div(a.n, R);
    const unsigned grid = (unsigned)(nblk * (a.C / CS));
    const bool stats = a.partials != nullptr, bias = a.bias != nullptr;
    const size_t smem = smem_bytes<CS, R>();
    static PerDeviceOnce configured;
    if (configured.need()) {
        const int mx = (int)smem;
        DDMP_CUDA(cudaFuncSetAttribute(spmm_tile_kernel<CS, R, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        DDMP_CUDA(cudaFuncSetAttribute(spmm_tile_kernel<CS, R, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        DDMP_CUDA(cudaFuncSetAttribute(spmm_tile_kernel<CS, R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        DDMP_CUDA(cudaFuncSetAttribute(spmm_tile_kernel<CS, R, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        configured.mark();
    }
    if (stats) {
        if (bias) spmm_tile_kernel<CS, R, true, true><<<grid, kThreads, smem, st>>>(map, a);
        else spmm_tile_kernel<CS, R, true, false><<<grid, kThreads, smem, st>>>(map, a);
    } else {
        if (bias) spmm_tile_kernel<CS, R, false, true><<<grid, kThreads, smem, st>>>(map, a);
        else spmm_tile_kernel<CS, R, false, false><<<grid, kThreads, smem, st>>>(map, a);
    }
    return check_launch("spmm_tile");
}

template <int CS, int R>
constexpr size_t bwd_smem_bytes() {
    return 128 + (size_t)2 * R * CS * 4 + (size_t)R * kMetaPerRow * 8 + (size_t)(R + 4) * 4 + 8 + (kThreads / 32) * 4 + 16;
}

template <int CS, int R>
static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
    CUtensorMap gmap, ymap;
    int rc = make_map(&gmap, a.gX, a.n, a.C, CS, R);
    if (rc != DDMP_OK) return rc;
    rc = make_map(&ymap, a.Y, a.n, a.C, CS, R);
    if (rc != DDMP_OK) return rc;
    const int64_t nblk = ceil_div(a.n, R);
    const unsigned grid = (unsigned)(nblk * (a.C / CS));
    const size_t smem = bwd_smem_bytes<CS, R>();
    static PerDeviceOnce configured;
    if (configured.need()) {
        DDMP_CUDA(cudaFuncSetAttribute(spmm_bn_bwd_tile_kernel<CS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.mark();
    }
    spmm_bn_bwd_tile_kernel<CS, R><<<grid, kThreads, smem, st>>>(gmap, ymap, a);
    return check_launch("spmm_bn_bwd_tile");
}

}  // namespace tile

// rows per CTA of the fused backward kernel (two tiles per CTA, so half the forward kernel's block)
int spmm_bn_bwd_tile_rows(int C) { return C == 32 ? 256 : (C == 64 ? 128 : 64); }

int spmm_bn_bwd_tile_launch(const int* rowptr, const int* col, const float* w, const float* gX, const float* Y,
                            const float* mean, const float* rstd, const float* scale, const float* shift,
                            const float* c1, const float* c2, float slope, float* dH, float* colsum, float* amax,
                            int64_t n, int C, cudaStream_t st) {
    tile::BwdArgs a{rowptr, col, w, gX, Y, mean, rstd, scale, shift, c1, c2, slope, dH, colsum, amax, n, C};
    if (C == 32) return tile::launch_bwd<32, 256>(a, st);
    if (C == 64) return tile::launch_bwd<64, 128>(a, st);
    return tile::launch_bwd<128, 64>(a, st);
}

// Kernel choice (ddmp_spmm_use_tile_kernel(v) / environment DDMP_SPMM_TILE; v = mode | flags << 4):
//   mode 0 = gather kernel for every width
//        1 = tile-staged kernel for C <= 128, gather kernel above (default: measured fastest on B200, 1M-face graphs,
//            profiles/spmm_tile_ab_r2.txt, profiles/spmm_pipe_ab_r2.txt)
//        2 = tile-staged kernel for every supported width
//   flags bit 1 = streaming (evict-first) stores of Y in the gather kernel (default on: Y is never re-read from L2 by this
//   kernel, and keeping it out leaves the cache to the gathered rows: +1 % on the wide layers).  -1: not decided yet.
static std::atomic<int> g_tile_mode{-1};
constexpr int kDefaultSetting = 1 | ((2 | 4 | 8) << 4);  // tile kernel for C <= 128, streaming stores, TMEM Welford state at C = 256 / 512

static int tile_setting() {
    int v = g_tile_mode.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("DDMP_SPMM_TILE");
        v = e ? atoi(e) : kDefaultSetting;
        if (v < 0 || (v & 15) > 2) v = kDefaultSetting;
        g_tile_mode.store(v, std::memory_order_relaxed);
    }
    return v;
}
static int tile_mode() { return tile_setting() & 15; }
int spmm_flags() { return tile_setting() >> 4; }

int spmm_tile_set(int v) {
    const int prev = tile_setting();
    if (v < 0 || (v & 15) > 2) v = kDefaultSetting;
    g_tile_mode.store(v, std::memory_order_relaxed);
    return prev;
}

bool spmm_tile_supported(int64_t n, int C) {
    const int mode = tile_mode();
    if (mode == 0 || n >= (1ll << 31)) return false;
    if (C == 32 || C == 64 || C == 128) return true;
    return mode == 2 && C > 128 && C <= 512 && C % 128 == 0;
}

int spmm_tile_launch(const int* rowptr, const int* col, const float* w, const float* H, const float* bias, float* Y,
                     float* partials, float* amax, int64_t n, int C, cudaStream_t st) {
    tile::Args a{rowptr, col, w, H, bias, Y, partials, amax, n, C};
    const int rpb = ddmp_rows_per_block(C);               // the BatchNorm partials are laid out per row block
    if (C == 32 && rpb == 256) return tile::launch<32, 256>(a, st);
    if (C == 64 && rpb == 256) return tile::launch<64, 256>(a, st);
    if (C == 64 && rpb == 128) return tile::launch<64, 128>(a, st);
    if (C >= 128 && rpb == 128) return tile::launch<128, 128>(a, st);
    set_error("spmm_tile: no instantiation for C=%d rows_per_block=%d", C, rpb);
    return DDMP_ERR_UNSUPPORTED;
}

}  // namespace ddmp
