// Network heads 32 -> 16 -> 3, fused per row (one thread per node).
//   PosNet    (reference util/networks.py:64-67)  : out = x_pos + linear2(lrelu(linear1(x)))
//   NormalNet (reference util/networks.py:125-129): t = tanh(linear2(lrelu(linear1(x)))); out = t/(||t||+1e-12)
// x = lrelu(scale*Y12 + shift) is the lazily applied BatchNorm+LeakyReLU of the last GCN layer.  Results are
// written in the CALLER's node numbering (out[perm[i]]) so the space-filling-curve order never leaks out.
// Weight gradients of linear1/linear2 are produced by the dW GEMM from the (gh, go) rows written here.
#include "common.cuh"

namespace ddmp {

constexpr int HC0 = 32, HC1 = 16, HC2 = 3;

template <int KIND>
__global__ void __launch_bounds__(128)
head_fwd_kernel(const float* __restrict__ Y12, const float* __restrict__ scale, const float* __restrict__ shift,
                float slope, const float* __restrict__ W1, const float* __restrict__ b1,
                const float* __restrict__ W2, const float* __restrict__ b2, const int* __restrict__ perm,
                const float* __restrict__ x_pos, float* __restrict__ out, float* __restrict__ h_save,
                float* __restrict__ t_save, int64_t n) {
    __shared__ float sW1[HC1 * HC0], sb1[HC1], sW2[HC2 * HC1], sb2[HC2], ssc[HC0], ssh[HC0];
    for (int i = threadIdx.x; i < HC1 * HC0; i += blockDim.x) sW1[i] = W1[i];
    for (int i = threadIdx.x; i < HC2 * HC1; i += blockDim.x) sW2[i] = W2[i];
    if (threadIdx.x < HC1) sb1[threadIdx.x] = b1[threadIdx.x];
    if (threadIdx.x < HC2) sb2[threadIdx.x] = b2[threadIdx.x];
    if (threadIdx.x < HC0) { ssc[threadIdx.x] = scale[threadIdx.x]; ssh[threadIdx.x] = shift[threadIdx.x]; }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;

    float x[HC0];
#pragma unroll
    for (int v = 0; v < HC0 / 4; ++v) {
        const float4 y = ldg4(Y12 + r * HC0 + v * 4);
        x[v * 4 + 0] = lrelu(fmaf(y.x, ssc[v * 4 + 0], ssh[v * 4 + 0]), slope);
        x[v * 4 + 1] = lrelu(fmaf(y.y, ssc[v * 4 + 1], ssh[v * 4 + 1]), slope);
        x[v * 4 + 2] = lrelu(fmaf(y.z, ssc[v * 4 + 2], ssh[v * 4 + 2]), slope);
        x[v * 4 + 3] = lrelu(fmaf(y.w, ssc[v * 4 + 3], ssh[v * 4 + 3]), slope);
    }
    float h[HC1];
#pragma unroll
    for (int j = 0; j < HC1; ++j) {
        float a = sb1[j];
#pragma unroll
        for (int k = 0; k < HC0; ++k) a = fmaf(sW1[j * HC0 + k], x[k], a);
        h[j] = lrelu(a, slope);
    }
#pragma unroll
    for (int v = 0; v < HC1 / 4; ++v)
        st4(h_save + r * HC1 + v * 4, make_float4(h[v * 4], h[v * 4 + 1], h[v * 4 + 2], h[v * 4 + 3]));
    float o[HC2];
#pragma unroll
    for (int i = 0; i < HC2; ++i) {
        float a = sb2[i];
#pragma unroll
        for (int j = 0; j < HC1; ++j) a = fmaf(sW2[i * HC1 + j], h[j], a);
        o[i] = a;
    }
    const int64_t p = perm ? (int64_t)perm[r] : r;
    if (KIND == DDMP_HEAD_POS) {
        out[p * 3 + 0] = x_pos[p * 3 + 0] + o[0];
        out[p * 3 + 1] = x_pos[p * 3 + 1] + o[1];
        out[p * 3 + 2] = x_pos[p * 3 + 2] + o[2];
    } else {
        const float t0 = tanhf(o[0]), t1 = tanhf(o[1]), t2 = tanhf(o[2]);
        const float nrm = sqrtf(t0 * t0 + t1 * t1 + t2 * t2);
        const float inv = 1.0f / (nrm + 1.0e-12f);
        out[p * 3 + 0] = t0 * inv;
        out[p * 3 + 1] = t1 * inv;
        out[p * 3 + 2] = t2 * inv;
        st4(t_save + r * 4, make_float4(t0, t1, t2, nrm));
    }
}

template <int KIND>
__global__ void __launch_bounds__(128)
head_bwd_kernel(const float* __restrict__ g_out, const int* __restrict__ perm, const float* __restrict__ W1,
                const float* __restrict__ W2, const float* __restrict__ h_save, const float* __restrict__ t_save,
                float slope, float* __restrict__ go, float* __restrict__ gh, float* __restrict__ gX12, int64_t n) {
    __shared__ float sW1[HC1 * HC0], sW2[HC2 * HC1];
    for (int i = threadIdx.x; i < HC1 * HC0; i += blockDim.x) sW1[i] = W1[i];
    for (int i = threadIdx.x; i < HC2 * HC1; i += blockDim.x) sW2[i] = W2[i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int64_t p = perm ? (int64_t)perm[r] : r;
    float g0 = g_out[p * 3 + 0], g1 = g_out[p * 3 + 1], g2 = g_out[p * 3 + 2];
    if (KIND == DDMP_HEAD_NORM) {
        const float4 t = ldg4(t_save + r * 4);      // (t0, t1, t2, ||t||)
        const float s = 1.0f / (t.w + 1.0e-12f);
        const float dot = g0 * t.x + g1 * t.y + g2 * t.z;
        const float k = (t.w > 0.f) ? (s * s * dot / t.w) : 0.f;
        const float gt0 = s * g0 - k * t.x, gt1 = s * g1 - k * t.y, gt2 = s * g2 - k * t.z;
        g0 = gt0 * (1.f - t.x * t.x);
        g1 = gt1 * (1.f - t.y * t.y);
        g2 = gt2 * (1.f - t.z * t.z);
    }
    st4(go + r * 4, make_float4(g0, g1, g2, 0.f));
    float ghv[HC1];
#pragma unroll
    for (int v = 0; v < HC1 / 4; ++v) {
        const float4 h = ldg4(h_save + r * HC1 + v * 4);
        const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = v * 4 + e;
            const float a = sW2[0 * HC1 + j] * g0 + sW2[1 * HC1 + j] * g1 + sW2[2 * HC1 + j] * g2;
            ghv[j] = (hh[e] > 0.f) ? a : a * slope;
        }
        st4(gh + r * HC1 + v * 4, make_float4(ghv[v * 4], ghv[v * 4 + 1], ghv[v * 4 + 2], ghv[v * 4 + 3]));
    }
#pragma unroll
    for (int v = 0; v < HC0 / 4; ++v) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < HC1; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) a[e] = fmaf(sW1[j * HC0 + v * 4 + e], ghv[j], a[e]);
        }
        st4(gX12 + r * HC0 + v * 4, make_float4(a[0], a[1], a[2], a[3]));
    }
}

}  // namespace ddmp

extern "C" {

int ddmp_head_fwd(int kind, const float* Y12, const float* scale, const float* shift, float slope, const float* W1,
                  const float* b1, const float* W2, const float* b2, const int32_t* perm, const float* x_pos,
                  float* out, float* h_save, float* t_save, int64_t n, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(Y12 && scale && shift && W1 && b1 && W2 && b2 && out && h_save, "head_fwd: null pointer");
    DDMP_REQUIRE(kind == DDMP_HEAD_POS || kind == DDMP_HEAD_NORM, "head_fwd: bad kind %d", kind);
    DDMP_REQUIRE(kind == DDMP_HEAD_NORM ? (t_save != nullptr) : (x_pos != nullptr),
                 "head_fwd: POS needs x_pos, NORM needs t_save");
    if (n == 0) return DDMP_OK;
    const unsigned grid = (unsigned)ceil_div(n, 128);
    cudaStream_t st = as_stream(stream);
    if (kind == DDMP_HEAD_POS)
        head_fwd_kernel<DDMP_HEAD_POS><<<grid, 128, 0, st>>>(Y12, scale, shift, slope, W1, b1, W2, b2, perm, x_pos,
                                                             out, h_save, t_save, n);
    else
        head_fwd_kernel<DDMP_HEAD_NORM><<<grid, 128, 0, st>>>(Y12, scale, shift, slope, W1, b1, W2, b2, perm, x_pos,
                                                              out, h_save, t_save, n);
    return check_launch("head_fwd");
}

int ddmp_head_bwd(int kind, const float* g_out, const int32_t* perm, const float* W1, const float* W2,
                  const float* h_save, const float* t_save, float slope, float* go, float* gh, float* gX12,
                  int64_t n, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(g_out && W1 && W2 && h_save && go && gh && gX12, "head_bwd: null pointer");
    DDMP_REQUIRE(kind == DDMP_HEAD_POS || (kind == DDMP_HEAD_NORM && t_save), "head_bwd: bad kind / t_save");
    if (n == 0) return DDMP_OK;
    const unsigned grid = (unsigned)ceil_div(n, 128);
    cudaStream_t st = as_stream(stream);
    if (kind == DDMP_HEAD_POS)
        head_bwd_kernel<DDMP_HEAD_POS><<<grid, 128, 0, st>>>(g_out, perm, W1, W2, h_save, t_save, slope, go, gh,
                                                             gX12, n);
    else
        head_bwd_kernel<DDMP_HEAD_NORM><<<grid, 128, 0, st>>>(g_out, perm, W1, W2, h_save, t_save, slope, go, gh,
                                                              gX12, n);
    return check_launch("head_bwd");
}

}  // extern "C"
