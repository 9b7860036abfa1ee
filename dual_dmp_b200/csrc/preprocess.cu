// Mesh preprocessing on the device, float64: the conventions of the reference's offline tools
// (preprocess/preprocess.py:22-28,68-72 and preprocess/noisemaker.py:25-42, which need pymeshlab) restated as kernels
// over the index tables the loss kernels already use (vertex adjacency CSR, corner CSR):
//   * uniform Laplacian smoothing sweep  x_i <- (x_i + sum_{j in N(i)} x_j) / (deg_i + 1)   (x30: the *_smooth mesh)
//   * float64 face normals / areas and vertex normals (normalised sum of incident face normals, util/mesh.py:87-107)
//   * Gaussian noise along the vertex normal  vs += vn * noise   (noise drawn on the host with np.random.seed(314))
//   * mean edge length (rescale so it is 1) and bounding box (unit-box normalisation + centring)
// All HBM-bound streaming kernels over V / F / E elements; deterministic (fixed-order float64 reductions).
#include "common.cuh"

namespace ddmp {
namespace prep {

constexpr int kT = 256;
constexpr int kMaxBlocks = 1024;

static inline unsigned grid_for(int64_t count) {
    int64_t g = ceil_div(count, kT);
    if (g > kMaxBlocks) g = kMaxBlocks;
    return (unsigned)(g < 1 ? 1 : g);
}

#define PREP_STRIDE(i, count) \
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (count); i += (int64_t)gridDim.x * blockDim.x)

__global__ void smooth_sweep_kernel(const double* __restrict__ x, const int* __restrict__ rowptr,
                                    const int* __restrict__ col, double* __restrict__ y, int64_t V) {
    PREP_STRIDE(i, V) {
        const int s = rowptr[i], e = rowptr[i + 1];
        double ax = x[3 * i], ay = x[3 * i + 1], az = x[3 * i + 2];
        for (int k = s; k < e; ++k) {
            const int64_t j = col[k];
            ax += x[3 * j]; ay += x[3 * j + 1]; az += x[3 * j + 2];
        }
        const double inv = 1.0 / ((double)(e - s) + 1.0);
        y[3 * i] = ax * inv; y[3 * i + 1] = ay * inv; y[3 * i + 2] = az * inv;
    }
}

// util/mesh.py:87-92: n = cross(v1-v0, v2-v0); fa = 0.5*sqrt(sum n^2); fn = n / (|n| + 1e-24)
__global__ void face_geometry_kernel(const double* __restrict__ vs, const int* __restrict__ faces,
                                     double* __restrict__ fn, double* __restrict__ fa, double* __restrict__ fc,
                                     int64_t F) {
    PREP_STRIDE(f, F) {
        double p[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int64_t v = faces[3 * f + k];
            p[k][0] = vs[3 * v]; p[k][1] = vs[3 * v + 1]; p[k][2] = vs[3 * v + 2];
        }
        const double ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1], az = p[1][2] - p[0][2];
        const double bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1], bz = p[2][2] - p[0][2];
        const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        const double s2 = cx * cx + cy * cy + cz * cz;
        const double nr = sqrt(s2) + 1e-24;
        if (fn) { fn[3 * f] = cx / nr; fn[3 * f + 1] = cy / nr; fn[3 * f + 2] = cz / nr; }
        if (fa) fa[f] = 0.5 * sqrt(s2);
        if (fc) {
#pragma unroll
            for (int a = 0; a < 3; ++a) fc[3 * f + a] = (p[0][a] + p[1][a] + p[2][a]) / 3.0;
        }
    }
}

// util/mesh.py:94-107: vn = normalise(sum of incident face normals); all-zero rows stay zero (sklearn normalize)
__global__ void vertex_normals_kernel(const double* __restrict__ fn, const int* __restrict__ corner_ptr,
                                      const int* __restrict__ corner_slot, double* __restrict__ vn, int64_t V) {
    PREP_STRIDE(v, V) {
        double x = 0.0, y = 0.0, z = 0.0;
        for (int k = corner_ptr[v]; k < corner_ptr[v + 1]; ++k) {
            const int64_t f = corner_slot[k] / 3;
            x += fn[3 * f]; y += fn[3 * f + 1]; z += fn[3 * f + 2];
        }
        double nr = sqrt(x * x + y * y + z * z);
        if (nr == 0.0) nr = 1.0;
        vn[3 * v] = x / nr; vn[3 * v + 1] = y / nr; vn[3 * v + 2] = z / nr;
    }
}

// out = (a + b * t[v]) * s   per vertex row  (noise along the normal: a = vs, b = vn, t = noise, s = 1;
// rescale / recentre: b = null, shift = -centre)
__global__ void affine_rows_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                   const double* __restrict__ t, const double* __restrict__ shift, double scale,
                                   double* __restrict__ out, int64_t V) {
    PREP_STRIDE(e, 3 * V) {
        double x = a[e];
        if (b) x += b[e] * t[e / 3];
        if (shift) x += shift[e % 3];
        out[e] = x * scale;
    }
}

struct Scratch {
    unsigned int ticket;
    unsigned int pad[63];
    double partials[7][kMaxBlocks];
};

// sum of edge lengths, and the bounding box (min / max per axis), one launch each, last block combines in block order
__global__ void __launch_bounds__(kT)
edge_length_kernel(const double* __restrict__ vs, const int* __restrict__ edges, double* out, Scratch* sc, int64_t E) {
    __shared__ double sm[kT / 32];
    double acc = 0.0;
    PREP_STRIDE(e, E) {
        const int64_t a = edges[2 * e], b = edges[2 * e + 1];
        const double dx = vs[3 * a] - vs[3 * b], dy = vs[3 * a + 1] - vs[3 * b + 1], dz = vs[3 * a + 2] - vs[3 * b + 2];
        acc += sqrt(dx * dx + dy * dy + dz * dz);
    }
    const double b = block_sum<kT>(acc, sm);
    if (threadIdx.x == 0) sc->partials[0][blockIdx.x] = b;
    if (publish_and_am_last(&sc->ticket, gridDim.x) && threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned i = 0; i < gridDim.x; ++i) t += __ldcg(&sc->partials[0][i]);
        *out = t;
    }
}

__global__ void __launch_bounds__(kT)
bbox_kernel(const double* __restrict__ vs, double* out /* [6]: min xyz, max xyz */, Scratch* sc, int64_t V) {
    __shared__ double sm[6][kT / 32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    PREP_STRIDE(v, V) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double x = vs[3 * v + a];
            lo[a] = fmin(lo[a], x);
            hi[a] = fmax(hi[a], x);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { sm[a][threadIdx.x >> 5] = lo[a]; sm[3 + a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double r = sm[threadIdx.x][0];
        for (int i = 1; i < kT / 32; ++i) r = threadIdx.x < 3 ? fmin(r, sm[threadIdx.x][i]) : fmax(r, sm[threadIdx.x][i]);
        sc->partials[1 + threadIdx.x][blockIdx.x] = r;
    }
    if (publish_and_am_last(&sc->ticket, gridDim.x) && threadIdx.x == 0) {
        for (int q = 0; q < 6; ++q) {
            double r = __ldcg(&sc->partials[1 + q][0]);
            for (unsigned i = 1; i < gridDim.x; ++i) {
                const double x = __ldcg(&sc->partials[1 + q][i]);
                r = q < 3 ? fmin(r, x) : fmax(r, x);
            }
            out[q] = r;
        }
    }
}

}  // namespace prep
}  // namespace ddmp

extern "C" {

int64_t ddmp_prep_scratch_bytes(void) { return (int64_t)sizeof(ddmp::prep::Scratch); }

int ddmp_prep_smooth_sweep(const double* pos_in, const int32_t* rowptr, const int32_t* col, double* pos_out, int64_t V,
                           void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(pos_in && rowptr && col && pos_out && V > 0 && pos_in != pos_out, "prep_smooth_sweep: bad arguments");
    prep::smooth_sweep_kernel<<<prep::grid_for(V), prep::kT, 0, as_stream(stream)>>>(pos_in, rowptr, col, pos_out, V);
    return check_launch("prep_smooth_sweep");
}

int ddmp_prep_face_geometry(const double* vs, const int32_t* faces, double* fn, double* fa, double* fc, int64_t F,
                            void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(vs && faces && F > 0 && (fn || fa || fc), "prep_face_geometry: bad arguments");
    prep::face_geometry_kernel<<<prep::grid_for(F), prep::kT, 0, as_stream(stream)>>>(vs, faces, fn, fa, fc, F);
    return check_launch("prep_face_geometry");
}

int ddmp_prep_vertex_normals(const double* fn, const int32_t* corner_ptr, const int32_t* corner_slot, double* vn,
                             int64_t V, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(fn && corner_ptr && corner_slot && vn && V > 0, "prep_vertex_normals: bad arguments");
    prep::vertex_normals_kernel<<<prep::grid_for(V), prep::kT, 0, as_stream(stream)>>>(fn, corner_ptr, corner_slot, vn, V);
    return check_launch("prep_vertex_normals");
}

int ddmp_prep_affine_rows(const double* a, const double* b, const double* t, const double* shift, double scale,
                          double* out, int64_t V, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(a && out && V > 0 && (!b || t), "prep_affine_rows: bad arguments");
    prep::affine_rows_kernel<<<prep::grid_for(3 * V), prep::kT, 0, as_stream(stream)>>>(a, b, t, shift, scale, out, V);
    return check_launch("prep_affine_rows");
}

int ddmp_prep_edge_length_sum(const double* vs, const int32_t* edges, double* out, void* scratch, int64_t E,
                              void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(vs && edges && out && scratch && E > 0, "prep_edge_length_sum: bad arguments");
    prep::edge_length_kernel<<<prep::grid_for(E), prep::kT, 0, as_stream(stream)>>>(vs, edges, out, (prep::Scratch*)scratch, E);
    return check_launch("prep_edge_length_sum");
}

int ddmp_prep_bbox(const double* vs, double* out6, void* scratch, int64_t V, void* stream) {
    using namespace ddmp;
    DDMP_REQUIRE(vs && out6 && scratch && V > 0, "prep_bbox: bad arguments");
    prep::bbox_kernel<<<prep::grid_for(V), prep::kT, 0, as_stream(stream)>>>(vs, out6, (prep::Scratch*)scratch, V);
    return check_launch("prep_bbox");
}

}  // extern "C"
