// cov_hull.cu — GPU stage of Katz hidden-point removal: vertex set of conv(flipped U {0})
// (reference src/tools.py:56-64 hands this to scipy/Qhull on the CPU).
//
// Instead of building hull facets, every point decides for itself whether it is an extreme point
// (hull_core.h): a 2-D least-tilt LP over its neighbours on the direction sphere, with a coverage bound
// that makes the neighbourhood search rigorous and fp64 certificates for both outcomes.  The flipped
// cloud is a thin shell (|f| in [199,200] * max|p| for R_param = 2), so a point's fate is decided by the
// few dozen..thousand points within a small cap around its direction; a counting sort by direction voxel
// makes those contiguous.  Kernels:
//   hull_prep      voxel key + histogram + max |f|                       (12 B/point read)
//   hull_scan      exclusive scan of the G^3 histogram (one block, coalesced chunks with a running carry)
//   hull_occupied  compact list of non-empty voxels (for the far phase)
//   hull_scatter   counting-sort scatter into (x, y, z, original index) records
//   hull_classify  one thread per point: hull_classify_point -> vertex mask, counters
//   hull_origin    GJK distance from the origin to conv(F) in one block (is the origin a vertex?)
#include "cov_common.cuh"
#include "hull_core.h"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kHullMaxG = 128;

int hull_grid_size(int64_t n) {
    int G = (int)lround(sqrt((double)n / (12.0 * 3.141592653589793)));
    return G < 1 ? 1 : (G > kHullMaxG ? kHullMaxG : G);
}

struct HullWs {  // carve-up of the caller's workspace
    int* key;             // n
    int* cell_count;      // G^3 + 1 (becomes cell_start after the scan)
    int* cursor;          // G^3
    int* occ;             // G^3
    float4* sorted;       // n
    unsigned long long* rho_max_bits;  // 1
    int* n_occ;           // 1
    int* n_valid;         // 1
};

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

size_t hull_carve(void* base, int64_t n, int G, HullWs* w) {
    const size_t ncell = (size_t)G * G * G;
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off += align16(bytes); return p; };
    HullWs t;
    t.sorted = (float4*)take((size_t)n * sizeof(float4));
    t.key = (int*)take((size_t)n * sizeof(int));
    t.cell_count = (int*)take((ncell + 1) * sizeof(int));
    t.cursor = (int*)take(ncell * sizeof(int));
    t.occ = (int*)take(ncell * sizeof(int));
    t.rho_max_bits = (unsigned long long*)take(16);
    t.n_occ = (int*)take(16);
    t.n_valid = (int*)take(16);
    if (w) *w = t;
    return off;
}

__global__ void __launch_bounds__(256)
hull_prep_kernel(const float* __restrict__ f, int64_t n, int G, int* __restrict__ key, int* __restrict__ cell_count,
                 unsigned long long* __restrict__ rho_max_bits) {
    double mx = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double x = f[i * 3], y = f[i * 3 + 1], z = f[i * 3 + 2];
        const double rho = sqrt(x * x + y * y + z * z);
        int k = -1;
        if (rho > 0.0 && rho < 1e300) {  // a point at the origin (or a NaN) duplicates the extra hull point: never a vertex
            k = (hull_cell_coord(x / rho, G) * G + hull_cell_coord(y / rho, G)) * G + hull_cell_coord(z / rho, G);
            atomicAdd(cell_count + k, 1);
            mx = fmax(mx, rho);
        }
        key[i] = k;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(rho_max_bits, (unsigned long long)__double_as_longlong(mx));
}

// exclusive scan in place over ncell+1 entries (entry ncell receives the total).  One block walks the histogram in
// chunks of 4096 consecutive cells (4 per thread, so a warp touches 512 contiguous bytes) and carries the running total.
__global__ void __launch_bounds__(1024) hull_scan_kernel(int* __restrict__ cnt, int64_t ncell, int* __restrict__ n_valid) {
    __shared__ int wtot[32];
    __shared__ int chunk_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int carry = 0;
    for (int64_t base = 0; base < ncell; base += 4096) {
        const int64_t i = base + (int64_t)t * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i + k < ncell) ? cnt[i + k] : 0;
        const int s = v[0] + v[1] + v[2] + v[3];
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = wtot[lane];
            int sc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, sc, o);
                if (lane >= o) sc += u;
            }
            wtot[lane] = sc - w;  // exclusive over warps
            if (lane == 31) chunk_total = sc;
        }
        __syncthreads();
        int run = carry + wtot[warp] + incl - s;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < ncell) cnt[i + k] = run;
            run += v[k];
        }
        carry += chunk_total;
        __syncthreads();  // wtot / chunk_total are rewritten by the next chunk
    }
    if (t == 0) {
        cnt[ncell] = carry;
        *n_valid = carry;
    }
}

__global__ void __launch_bounds__(256) hull_occupied_kernel(const int* __restrict__ cell_start, int64_t ncell,
                                                             int* __restrict__ occ, int* __restrict__ n_occ) {
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < ncell && cell_start[c + 1] > cell_start[c]) occ[atomicAdd(n_occ, 1)] = (int)c;
}

__global__ void __launch_bounds__(256)
hull_scatter_kernel(const float* __restrict__ f, int64_t n, const int* __restrict__ key, const int* __restrict__ cell_start,
                    int* __restrict__ cursor, float4* __restrict__ sorted, uint8_t* __restrict__ mask) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int k = key[i];
        mask[i] = 0;
        if (k < 0) continue;
        const int pos = cell_start[k] + atomicAdd(cursor + k, 1);
        sorted[pos] = make_float4(f[i * 3], f[i * 3 + 1], f[i * 3 + 2], __int_as_float((int)i));
    }
}

__global__ void __launch_bounds__(128)
hull_classify_kernel(int G, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                     const int* __restrict__ occ, const int* __restrict__ n_occ,
                     const unsigned long long* __restrict__ rho_max_bits, const int* __restrict__ n_valid,
                     uint8_t* __restrict__ mask, int* __restrict__ info) {
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= *n_valid) return;
    HullGrid g;
    g.G = G;
    g.h = 2.0 / G;
    g.cell_start = cell_start;
    g.sorted = sorted;
    g.occ = occ;
    g.n_occ = *n_occ;
    g.rho_max = __longlong_as_double((long long)*rho_max_bits);
    int cert[3];
    const int rc = hull_classify_point(g, k, cert);
    const bool vertex = (rc == HULL_EXTREME || rc == HULL_EXTREME_UNCERT);  // OVERFLOW: the LP ran away, no feasible region in reach
    mask[__float_as_int(sorted[k].w)] = vertex ? 1 : 0;
    if (rc != HULL_EXTREME && rc != HULL_INSIDE) atomicAdd(info + 1, 1);  // decided by the LP but not certified in fp64
    if (vertex) atomicAdd(info + 2, 1);
}

// One block: GJK on conv(F).  info[0] = 1 when the origin is outside conv(F) (i.e. a vertex of conv(F U {0})).
__global__ void __launch_bounds__(1024)
hull_origin_kernel(const float* __restrict__ f, int64_t n, int* __restrict__ info) {
    __shared__ HullSimplex S;
    __shared__ double rv[32];
    __shared__ long long ri[32];
    __shared__ int state;  // 0 = iterate, 1 = done
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        int64_t j0 = 0;  // first finite non-zero point
        for (; j0 < n; ++j0) {
            const double q = (double)f[j0 * 3] * f[j0 * 3] + (double)f[j0 * 3 + 1] * f[j0 * 3 + 1] + (double)f[j0 * 3 + 2] * f[j0 * 3 + 2];
            if (q > 0.0 && q < 1e300) break;
        }
        state = 0;
        if (j0 >= n) { info[0] = 1; state = 1; }
        else {
            S.n = 1;
            for (int c = 0; c < 3; ++c) S.v[0][c] = S.x[c] = f[j0 * 3 + c];
        }
    }
    __syncthreads();
    for (int it = 0; it < 64 && state == 0; ++it) {
        const double x0 = S.x[0], x1 = S.x[1], x2 = S.x[2];
        double best = 1e300;
        long long bj = -1;
        for (int64_t j = t; j < n; j += 1024) {
            const double d = x0 * f[j * 3] + x1 * f[j * 3 + 1] + x2 * f[j * 3 + 2];
            if (d < best) { best = d; bj = j; }  // NaNs never compare smaller
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const long long oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob < best || (ob == best && oj >= 0 && (bj < 0 || oj < bj))) { best = ob; bj = oj; }
        }
        if (lane == 0) { rv[warp] = best; ri[warp] = bj; }
        __syncthreads();
        if (t == 0) {
            for (int w = 1; w < 32; ++w)
                if (rv[w] < best || (rv[w] == best && ri[w] >= 0 && (bj < 0 || ri[w] < bj))) { best = rv[w]; bj = ri[w]; }
            const double xx = hull_dot3(S.x, S.x);
            info[3] = it + 1;
            if (bj < 0 || best >= xx * (1.0 - 1e-10)) {  // no point lies further towards the origin: x is the closest point
                info[0] = 1;
                if (!(best > 1e-9 * xx)) atomicAdd(info + 1, 1);  // separating margin not certified
                state = 1;
            } else {
                for (int c = 0; c < 3; ++c) S.v[S.n][c] = f[bj * 3 + c];
                S.n++;
                if (hull_simplex_update(S)) {
                    info[0] = 0;
                    if (!hull_certify_origin_inside(S)) atomicAdd(info + 1, 1);
                    state = 1;
                }
            }
        }
        __syncthreads();
    }
    if (t == 0 && state == 0) {  // did not converge in 64 iterations: decide by distance, flag as uncertified
        info[0] = hull_dot3(S.x, S.x) > 0.0 ? 1 : 0;
        atomicAdd(info + 1, 1);
    }
}

}  // namespace

extern "C" size_t cov_hpr_hull_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return hull_carve(nullptr, n, hull_grid_size(n), nullptr);
}

extern "C" int cov_hpr_hull(const float* flipped, int64_t n, uint8_t* vertex_mask, int32_t* info, void* ws,
                            size_t ws_bytes, void* stream) {
    if (n <= 0 || !flipped || !vertex_mask || !info || !ws) {
        cov_set_error("cov_hpr_hull: null pointer or empty cloud");
        return COV_ERR_ARG;
    }
    if (n >= ((int64_t)1 << 31) - 1) {
        cov_set_error("cov_hpr_hull: n >= 2^31 not supported");
        return COV_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)ws) & 15) {
        cov_set_error("cov_hpr_hull: workspace must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    const int G = hull_grid_size(n);
    HullWs w;
    const size_t need = hull_carve(ws, n, G, &w);
    if (ws_bytes < need) {
        cov_set_error("cov_hpr_hull: workspace %zu < %zu bytes", ws_bytes, need);
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ncell = (int64_t)G * G * G;
    // one memset clears histogram, cursors, occupied list and the three scalars (contiguous in the carve-up)
    cudaMemsetAsync(w.cell_count, 0, (size_t)((char*)w.n_valid + 16 - (char*)w.cell_count), s);
    cudaMemsetAsync(info, 0, 4 * sizeof(int32_t), s);
    int64_t nb = (n + 255) / 256;
    const int64_t cap = (int64_t)cov_sm_count_cached() * 16;
    if (nb > cap) nb = cap;
    hull_prep_kernel<<<(unsigned)nb, 256, 0, s>>>(flipped, n, G, w.key, w.cell_count, w.rho_max_bits);
    hull_scan_kernel<<<1, 1024, 0, s>>>(w.cell_count, ncell, w.n_valid);
    hull_occupied_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(w.cell_count, ncell, w.occ, w.n_occ);
    hull_scatter_kernel<<<(unsigned)nb, 256, 0, s>>>(flipped, n, w.key, w.cell_count, w.cursor, w.sorted, vertex_mask);
    hull_classify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(G, w.cell_count, w.sorted, w.occ, w.n_occ,
                                                                    w.rho_max_bits, w.n_valid, vertex_mask, info);
    hull_origin_kernel<<<1, 1024, 0, s>>>(flipped, n, info);
    return cov_check_launch("cov_hpr_hull");
}
