// BatchNorm in the partitioned mode: statistics reduction + all-reduce + finalize as ONE kernel over NVLink peer memory.
//
// The partitioned step needs 96 all-reduces of 2*C numbers per iteration (BatchNorm forward statistics and the two
// BatchNorm-backward sums of every layer; reference: nn.BatchNorm1d at util/networks.py:31-42,51-62 sees the whole
// batch).  As NCCL calls they cost ~80 us each on 8 GPUs (latency, not bandwidth: 8 KB messages) between a reduction
// kernel and a finalize kernel.  Here every rank owns a small exchange buffer that all peers map (CUDA IPC); the kernel
//   1. reduces this rank's row-block partials in float64, (channels / 8) x (up to 16 row-block chunks) CTAs, fixed order;
//   2. the last CTA to finish combines the chunks in order and stores the rank's 2*C sums into slot [parity][rank] of
//      EVERY peer's buffer (P2P stores through NVSwitch),
//   3. publishes a sequence number to every peer (release, system scope), waits until the sequence numbers of all peers
//      have arrived in its own buffer (acquire), and
//   4. sums the slots in RANK order -- every rank forms bit-identical totals -- and writes the BatchNorm table.
// One-shot, latency-bound by a single NVLink round trip; slots are double-buffered by the parity of the sequence number
// (a rank can be at most one exchange ahead of its slowest peer, because it cannot finish exchange s+1 before every peer
// has published s+1, i.e. has finished reading s).  Waits are bounded: a peer that never arrives sets an error word
// instead of hanging the GPU.
#include "common.cuh"

namespace ddmp {
namespace comm {

constexpr int kMaxRanks = 16;
constexpr int kMaxLen = 1024;                 // doubles per message (2 * C, C <= 512)
constexpr int kFinCh = 8;
constexpr int kFinLanes = 128;
constexpr int kFinThreads = kFinCh * kFinLanes;
constexpr unsigned long long kSpinLimit = 1ull << 24;   // ~10 s of polling: ranks are at most one layer apart

constexpr int kMaxSplit = 16;                 // CTAs along the row-block axis of the local reduction

struct Buffer {                               // lives in device memory of its owner, mapped by every peer
    double slots[2][kMaxRanks][kMaxLen];
    double local[kMaxSplit][kMaxLen];         // local: this rank's per-chunk sums, combined by the last CTA
    unsigned long long flags[2][kMaxRanks];   // sequence number last published by each rank, per parity
    unsigned int ticket;                      // local: CTAs of the running kernel that have finished
    unsigned int error;                       // local: set when a wait ran into kSpinLimit
};

struct Peers {
    Buffer* buf[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// float64 sums of [nblk][2][C] float32 partials for channel c, blocks dealt to kFinLanes lanes, lanes combined in order.
// MOMENTS: partials are (sum_b, M2_b): set 1 contributes M2_b + sum_b^2 / n_b (see bn.cu reduce_partials)
template <bool MOMENTS>
__device__ __forceinline__ void reduce2(const float* __restrict__ partials, int64_t nblk, int C, int c, double (&out)[2],
                                        int64_t n_rows, int rpb, int64_t b_begin, int64_t b_end) {
    __shared__ double red[kFinLanes][2][kFinCh + 1];
    const int tx = threadIdx.x % kFinCh, ty = threadIdx.x / kFinCh;
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
        for (int64_t b = b_begin + ty; b < b_end; b += kFinLanes) {
            const double s0 = (double)__ldg(partials + (b * 2) * C + c);
            const double s1 = (double)__ldg(partials + (b * 2 + 1) * C + c);
            if (MOMENTS) {
                const double nb = (b == nblk - 1) ? (double)(n_rows - (nblk - 1) * (int64_t)rpb) : (double)rpb;
                a0 += s0;
                a1 += s1 + s0 * s0 / nb;
            } else {
                a0 += s0;
                a1 += s1;
            }
        }
    }
    red[ty][0][tx] = a0;
    red[ty][1][tx] = a1;
    __syncthreads();
    double t0 = 0.0, t1 = 0.0;
#pragma unroll 8
    for (int y = 0; y < kFinLanes; ++y) { t0 += red[y][0][tx]; t1 += red[y][1][tx]; }
    out[0] = t0;
    out[1] = t1;
}

struct StatsOut {           // MODE 0: BatchNorm forward table
    const float* gamma; const float* beta; float eps; float momentum;
    float* running_mean; float* running_var; float* mean; float* rstd; float* scale; float* shift; float* bound;
};
struct BwdOut {             // MODE 1: BatchNorm backward sums
    float* dgamma; float* dbeta; float* c1; float* c2;
};

template <int MODE>
__global__ void __launch_bounds__(kFinThreads)
bn_allreduce_kernel(const float* __restrict__ partials, int64_t nblk, int64_t n_local, int C, int rpb, Peers peers,
                    int rank, int world, unsigned long long seq, int64_t n_global, StatsOut so, BwdOut bo) {
    Buffer* mine = peers.buf[rank];
    const int parity = (int)(seq & 1ull);
    const int c = blockIdx.x * kFinCh + (threadIdx.x % kFinCh);
    // stage 1: the row blocks are cut into gridDim.y chunks; every CTA reduces (8 channels) x (one chunk) in float64
    const int64_t chunk = (nblk + gridDim.y - 1) / gridDim.y;
    const int64_t b0 = (int64_t)blockIdx.y * chunk;
    const int64_t b1 = (b0 + chunk < nblk) ? (b0 + chunk) : nblk;
    double s[2];
    reduce2<MODE == 0>(partials, nblk, C, c, s, n_local, rpb, b0, b1);
    if (threadIdx.x < kFinCh && c < C) {
        mine->local[blockIdx.y][c] = s[0];
        mine->local[blockIdx.y][C + c] = s[1];
    }
    // last CTA of this rank: combine the chunks in order, publish to every peer, wait for the peers, total, finalize
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&mine->ticket, 1u);
        is_last = (t == gridDim.x * gridDim.y - 1);
        if (is_last) mine->ticket = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < 2 * C; i += kFinThreads) {
        double t = 0.0;
        for (unsigned y = 0; y < gridDim.y; ++y) t += ld_volatile_f64(&mine->local[y][i]);
        for (int p = 0; p < world; ++p) peers.buf[p]->slots[parity][rank][i] = t;      // P2P stores (own buffer included)
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) st_release_sys(&peers.buf[threadIdx.x]->flags[parity][rank], seq);
    if (threadIdx.x < world) {
        unsigned long long spins = 0;
        while (ld_acquire_sys(&mine->flags[parity][threadIdx.x]) < seq) {
            if (++spins > kSpinLimit) { mine->error = 1u + threadIdx.x; break; }
        }
    }
    __syncthreads();
    __threadfence_system();
    for (int ch = threadIdx.x; ch < C; ch += kFinThreads) {
        double S = 0.0, Q = 0.0;
        for (int q = 0; q < world; ++q) {                      // rank order: identical totals on every rank
            S += ld_volatile_f64(&mine->slots[parity][q][ch]);
            Q += ld_volatile_f64(&mine->slots[parity][q][C + ch]);
        }
        if (MODE == 0) {
            const double m = S / (double)n_global;
            double var = Q / (double)n_global - m * m;
            if (var < 0.0) var = 0.0;
            const float r = (float)(1.0 / sqrt(var + (double)so.eps));
            const float mf = (float)m;
            const float sc = so.gamma[ch] * r;
            so.mean[ch] = mf;
            so.rstd[ch] = r;
            so.scale[ch] = sc;
            so.shift[ch] = so.beta[ch] - mf * sc;
            if (so.bound) so.bound[ch] = fabsf(so.gamma[ch]) * sqrtf((float)(n_global > 1 ? n_global - 1 : 1)) + fabsf(so.beta[ch]);
            if (so.running_mean) so.running_mean[ch] = (1.f - so.momentum) * so.running_mean[ch] + so.momentum * mf;
            if (so.running_var) {
                const double unbiased = (n_global > 1) ? var * ((double)n_global / (double)(n_global - 1)) : var;
                so.running_var[ch] = (1.f - so.momentum) * so.running_var[ch] + so.momentum * (float)unbiased;
            }
        } else {
            bo.dbeta[ch] = (float)S;
            bo.dgamma[ch] = (float)Q;
            bo.c1[ch] = (float)(S / (double)n_global);
            bo.c2[ch] = (float)(Q / (double)n_global);
        }
    }
}

// (channels / 8) x (row-block chunks): >= 512 blocks per CTA, at most kMaxSplit chunks
static dim3 grid_for(int C, int64_t nblk) {
    int64_t split = nblk / 512;
    split = split < 1 ? 1 : (split > kMaxSplit ? kMaxSplit : split);
    return dim3((unsigned)ceil_div(C, kFinCh), (unsigned)split);
}

static int fill_peers(Peers& p, const void* const* peer_ptrs, int world) {
    DDMP_REQUIRE(peer_ptrs && world >= 1 && world <= kMaxRanks, "peer all-reduce: world must be 1..%d", kMaxRanks);
    for (int i = 0; i < kMaxRanks; ++i) p.buf[i] = i < world ? (Buffer*)peer_ptrs[i] : nullptr;
    for (int i = 0; i < world; ++i) DDMP_REQUIRE(p.buf[i] != nullptr, "peer all-reduce: null peer buffer");
    return DDMP_OK;
}

}  // namespace comm
}  // namespace ddmp

extern "C" {

int64_t ddmp_comm_buffer_bytes(void) { return (int64_t)sizeof(ddmp::comm::Buffer); }

int ddmp_comm_alloc(void** out) {
    using namespace ddmp;
    DDMP_REQUIRE(out, "comm_alloc: null pointer");
    void* p = nullptr;
    DDMP_CUDA(cudaMalloc(&p, sizeof(comm::Buffer)));
    DDMP_CUDA(cudaMemset(p, 0, sizeof(comm::Buffer)));
    DDMP_CUDA(cudaDeviceSynchronize());
    *out = p;
    return DDMP_OK;
}

int ddmp_comm_free(void* p) {
    using namespace ddmp;
    if (p) DDMP_CUDA(cudaFree(p));
    return DDMP_OK;
}

int ddmp_comm_ipc_handle(void* p, void* handle64) {
    using namespace ddmp;
    DDMP_REQUIRE(p && handle64, "comm_ipc_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DDMP_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p));
    return DDMP_OK;
}

int ddmp_comm_ipc_open(const void* handle64, void** out) {
    using namespace ddmp;
    DDMP_REQUIRE(handle64 && out, "comm_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    DDMP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out = p;
    return DDMP_OK;
}

int ddmp_comm_ipc_close(void* p) {
    using namespace ddmp;
    if (p) DDMP_CUDA(cudaIpcCloseMemHandle(p));
    return DDMP_OK;
}

/* 0 = no wait of a peer kernel ran into its bound since ddmp_comm_alloc (synchronises the device) */
int ddmp_comm_error(const void* p, int32_t* out) {
    using namespace ddmp;
    DDMP_REQUIRE(p && out, "comm_error: null pointer");
    unsigned int e = 0;
    DDMP_CUDA(cudaMemcpy(&e, &reinterpret_cast<const comm::Buffer*>(p)->error, sizeof(e), cudaMemcpyDeviceToHost));
    *out = (int32_t)e;
    return DDMP_OK;
}

int ddmp_bn_stats_finalize_peer(const float* partials, int64_t nblk, int64_t n_local, int32_t C,
                                const void* const* peer_buffers, int32_t rank, int32_t world, int64_t seq,
                                int64_t n_global, const float* gamma, const float* beta, float eps, float momentum,
                                float* running_mean, float* running_var, float* mean, float* rstd, float* scale,
                                float* shift, float* bound, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(partials && gamma && beta && mean && rstd && scale && shift, "bn_stats_finalize_peer: null pointer");
    DDMP_REQUIRE(n_local > 0 && n_global >= n_local && C > 0 && 2 * C <= comm::kMaxLen && nblk > 0 && seq > 0 && rank >= 0 &&
                     rank < world, "bn_stats_finalize_peer: bad arguments");
    const int rpb = ddmp_rows_per_block(C);
    DDMP_REQUIRE(nblk == ceil_div(n_local, rpb), "bn_stats_finalize_peer: partial blocks do not cover the rows");
    comm::Peers peers;
    int rc = comm::fill_peers(peers, peer_buffers, world);
    if (rc != DDMP_OK) return rc;
    comm::StatsOut so{gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift, bound};
    comm::BwdOut bo{};
    comm::bn_allreduce_kernel<0><<<comm::grid_for(C, nblk), comm::kFinThreads, 0, as_stream(stream)>>>(
        partials, nblk, n_local, C, rpb, peers, rank, world, (unsigned long long)seq, n_global, so, bo);
    return check_launch("bn_stats_finalize_peer");
}

int ddmp_bn_bwd_finalize_peer(const float* partials, int64_t nblk, int32_t C, const void* const* peer_buffers,
                              int32_t rank, int32_t world, int64_t seq, int64_t n_global, float* dgamma, float* dbeta,
                              float* c1, float* c2, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(partials && dgamma && dbeta && c1 && c2, "bn_bwd_finalize_peer: null pointer");
    DDMP_REQUIRE(n_global > 0 && C > 0 && 2 * C <= comm::kMaxLen && nblk > 0 && seq > 0 && rank >= 0 && rank < world,
                 "bn_bwd_finalize_peer: bad arguments");
    comm::Peers peers;
    int rc = comm::fill_peers(peers, peer_buffers, world);
    if (rc != DDMP_OK) return rc;
    comm::StatsOut so{};
    comm::BwdOut bo{dgamma, dbeta, c1, c2};
    comm::bn_allreduce_kernel<1><<<comm::grid_for(C, nblk), comm::kFinThreads, 0, as_stream(stream)>>>(
        partials, nblk, 0, C, 1, peers, rank, world, (unsigned long long)seq, n_global, so, bo);
    return check_launch("bn_bwd_finalize_peer");
}

}  // extern "C"
