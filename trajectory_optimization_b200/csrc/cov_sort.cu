// cov_sort.cu — spatial (Morton) ordering of a point cloud, done once per cloud.
//
// The reference builds one ModelTraj per cloud and then iterates the optimiser on it
// (src/trajectory_optimization.py:83-127, src/model.py:164), so the cloud is constant over hundreds of objective
// evaluations.  Ordering it along a Z-curve makes every run of consecutive points spatially compact, which is what
// the tile-level pruning of cov_traj.cu feeds on.  Pipeline (all on the caller's stream, workspace from the caller):
//   1. bounding box of the cloud           (block reduce + integer atomics on order-preserving keys)
//   2. 30-bit Morton key per point         (cubic cells: 1024 along the longest extent)
//   3. stable LSD radix sort of (key, index) pairs — covradix::sort_pairs (cov_radix.cuh: own kernels, 8 bits per
//      pass, MATCH.ANY ranking, tiles reordered in shared memory; 4 passes for the 30 key bits)
//   4. gather xyz_sorted[j] = xyz[perm[j]]
#include <algorithm>

#include "cov_common.cuh"
#include "cov_radix.cuh"
#include "../../include/coverage_b200.h"

namespace {

__device__ __forceinline__ unsigned sort_f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sort_ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

// box[0..2] = min xyz, box[3..5] = max xyz as order-preserving uints
__global__ void sort_box_init_kernel(unsigned* box) {
    if (threadIdx.x < 3) box[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) box[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) sort_box_kernel(const float* __restrict__ xyz, int64_t n, unsigned* __restrict__ box) {
    const float inf = __uint_as_float(0x7f800000u);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = __ldg(xyz + j * 3 + k);
            if (v == v && fabsf(v) != inf) {  // NaN / inf points do not shape the grid (they land in cell 0 or 1023)
                lo[k] = fminf(lo[k], v);
                hi[k] = fmaxf(hi[k], v);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned l = __reduce_min_sync(0xffffffffu, sort_f2ord(lo[k]));
        const unsigned h = __reduce_max_sync(0xffffffffu, sort_f2ord(hi[k]));
        if ((threadIdx.x & 31) == 0) {
            atomicMin(box + k, l);
            atomicMax(box + 3 + k, h);
        }
    }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256) sort_key_kernel(const float* __restrict__ xyz, int64_t n, const unsigned* __restrict__ box,
                                                       unsigned* __restrict__ keys, int32_t* __restrict__ idx) {
    const float lx = sort_ord2f(box[0]), ly = sort_ord2f(box[1]), lz = sort_ord2f(box[2]);
    const float ext = fmaxf(fmaxf(sort_ord2f(box[3]) - lx, sort_ord2f(box[4]) - ly), sort_ord2f(box[5]) - lz);
    const float scale = (ext > 0.f && ext < __uint_as_float(0x7f800000u)) ? 1023.999f / ext : 0.f;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
        // fminf/fmaxf drop NaN operands, so a NaN coordinate maps to cell 0
        const unsigned ix = (unsigned)fminf(fmaxf((x - lx) * scale, 0.f), 1023.f);
        const unsigned iy = (unsigned)fminf(fmaxf((y - ly) * scale, 0.f), 1023.f);
        const unsigned iz = (unsigned)fminf(fmaxf((z - lz) * scale, 0.f), 1023.f);
        keys[j] = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
        idx[j] = (int32_t)j;
    }
}

__global__ void __launch_bounds__(256) sort_gather_kernel(const float* __restrict__ xyz, int64_t n, const int32_t* __restrict__ perm,
                                                          float* __restrict__ out) {
    // 3 consecutive threads move one point, so the writes are fully coalesced; four independent (index, coordinate) load
    // chains per thread keep enough of the scattered reads in flight
    const int64_t total = 3 * n, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
        int64_t src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            const int64_t j = i / 3;
            src[u] = i < total ? (int64_t)__ldg(perm + j) * 3 + (i - 3 * j) : -1;
        }
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = src[u] >= 0 ? __ldg(xyz + src[u]) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (src[u] >= 0) out[i0 + u * stride] = v[u];
    }
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

// workspace: [box 256 B][keys A n u32][keys B n u32][idx B n i32][digit counts]
extern "C" size_t cov_spatial_sort_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return 256 + 3 * align256((size_t)n * 4) + align256(covradix::temp_bytes()) + 256;
}

extern "C" int cov_spatial_sort(const float* xyz, int64_t n, float* xyz_sorted, int32_t* perm, void* ws, size_t ws_bytes,
                                void* stream) {
    if (!xyz || !xyz_sorted || !perm || !ws || n <= 0) {
        cov_set_error("cov_spatial_sort: null pointer or empty cloud (n=%lld)", (long long)n);
        return COV_ERR_ARG;
    }
    if (n >= ((int64_t)1 << 31)) {
        cov_set_error("cov_spatial_sort: %lld points exceed the int32 index range", (long long)n);
        return COV_ERR_UNSUPPORTED;
    }
    if (ws_bytes < cov_spatial_sort_workspace_bytes(n)) {
        cov_set_error("cov_spatial_sort: workspace %zu < %zu bytes", ws_bytes, cov_spatial_sort_workspace_bytes(n));
        return COV_ERR_WORKSPACE;
    }
    if ((((uintptr_t)ws) & 255) || (((uintptr_t)perm) & 15)) {
        cov_set_error("cov_spatial_sort: workspace must be 256-byte aligned, perm 16-byte aligned (TMA bulk copies)");
        return COV_ERR_ALIGN;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    unsigned* box = reinterpret_cast<unsigned*>(base);
    const size_t col = align256((size_t)n * 4);
    unsigned* keys_a = reinterpret_cast<unsigned*>(base + 256);
    unsigned* keys_b = reinterpret_cast<unsigned*>(base + 256 + col);
    int32_t* idx_b = reinterpret_cast<int32_t*>(base + 256 + 2 * col);
    void* temp = base + 256 + 3 * col;

    const int sms = cov_sm_count_cached();
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    sort_box_init_kernel<<<1, 32, 0, s>>>(box);
    sort_box_kernel<<<grid, 256, 0, s>>>(xyz, n, box);
    sort_key_kernel<<<grid, 256, 0, s>>>(xyz, n, box, keys_a, perm);
    // 30 key bits = 4 passes (8, 8, 8, 6): the pairs end where they started, in (keys_a, perm)
    if (covradix::sort_pairs(keys_a, perm, keys_b, idx_b, n, 0, 30, temp, sms, s) != 0)
        cudaMemcpyAsync(perm, idx_b, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    const int ggrid = (int)std::min<int64_t>((3 * n + 255) / 256, (int64_t)cov_sm_count_cached() * 32);
    sort_gather_kernel<<<ggrid, 256, 0, s>>>(xyz, n, perm, xyz_sorted);
    return cov_check_launch("cov_spatial_sort");
}

// ================================= voxel-grid downsample =================================
// The reference runs a pcl/VoxelGrid nodelet in front of the optimiser (launch/voxels_filtering.launch:8-21: leaf_size,
// pass-through on z, filter_limit_negative False).  PCL is a third-party dependency that is not in the tree; this is a
// restatement of pcl::VoxelGrid<PointXYZ>::applyFilter (PCL 1.8-1.10): finite points inside the pass-through limits
// (inclusive) are binned into leaf-sized voxels, index = (floor(x/leaf) - floor(min/leaf)) + ... x fastest; one output
// point per occupied voxel = the fp32 centroid of its points, in ascending voxel index.  PCL sorts the point indices
// with std::sort (unstable), so its own centroids are only reproducible up to fp32 summation order; here the sort is
// stable (original order within a voxel) and the sum sequential: bitwise deterministic.
namespace {

struct VoxHeader {          // first 256 bytes of the workspace
    unsigned box[6];        // ord-mapped min xyz, max xyz of the kept points
    int min_b[3], div_b[3];
    int overflow;           // the grid has more than 2^31-1 cells (PCL: "Leaf size is too small")
    int n_voxels;
};

__device__ __forceinline__ bool vox_keep(float x, float y, float z, int axis, float lo, float hi) {
    const float inf = __uint_as_float(0x7f800000u);
    if (!(fabsf(x) < inf && fabsf(y) < inf && fabsf(z) < inf)) return false;
    if (axis < 0) return true;
    const float v = axis == 0 ? x : (axis == 1 ? y : z);
    return !(v > hi || v < lo);
}

__global__ void vox_init_kernel(VoxHeader* h) {
    if (threadIdx.x < 3) h->box[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) h->box[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        h->overflow = 0;
        h->n_voxels = 0;
    }
}

__global__ void __launch_bounds__(256) vox_box_kernel(const float* __restrict__ xyz, int64_t n, int axis, float lo, float hi,
                                                      VoxHeader* __restrict__ h) {
    const float inf = __uint_as_float(0x7f800000u);
    float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const float p[3] = {__ldg(xyz + j * 3), __ldg(xyz + j * 3 + 1), __ldg(xyz + j * 3 + 2)};
        if (vox_keep(p[0], p[1], p[2], axis, lo, hi)) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                mn[k] = fminf(mn[k], p[k]);
                mx[k] = fmaxf(mx[k], p[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned l = __reduce_min_sync(0xffffffffu, sort_f2ord(mn[k]));
        const unsigned u = __reduce_max_sync(0xffffffffu, sort_f2ord(mx[k]));
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&h->box[k], l);
            atomicMax(&h->box[3 + k], u);
        }
    }
}

__global__ void vox_dims_kernel(VoxHeader* h, float leaf) {
    if (threadIdx.x != 0) return;
    const float inv = __fdiv_rn(1.f, leaf);
    long long cells = 1;
    for (int k = 0; k < 3; ++k) {
        const float mn = sort_ord2f(h->box[k]), mx = sort_ord2f(h->box[3 + k]);
        if (!(mn <= mx)) {  // no point kept
            h->min_b[k] = 0;
            h->div_b[k] = 1;
            continue;
        }
        const float fmn = floorf(__fmul_rn(mn, inv)), fmx = floorf(__fmul_rn(mx, inv));
        if (!(fabsf(fmn) < 2.0e9f && fabsf(fmx) < 2.0e9f)) {
            h->overflow = 1;
            h->min_b[k] = 0;
            h->div_b[k] = 1;
            continue;
        }
        h->min_b[k] = (int)fmn;
        h->div_b[k] = (int)fmx - (int)fmn + 1;
        cells *= (long long)h->div_b[k];
        if (cells > 2147483647LL) h->overflow = 1;
    }
}

__global__ void __launch_bounds__(256) vox_key_kernel(const float* __restrict__ xyz, int64_t n, float leaf, int axis, float lo,
                                                      float hi, const VoxHeader* __restrict__ h, unsigned* __restrict__ keys,
                                                      int32_t* __restrict__ idx) {
    const float inv = __fdiv_rn(1.f, leaf);
    const float b0 = (float)h->min_b[0], b1 = (float)h->min_b[1], b2 = (float)h->min_b[2];
    const int d0 = h->div_b[0], d01 = h->div_b[0] * h->div_b[1];
    const bool bad = h->overflow != 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
        unsigned key = 0xffffffffu;  // dropped points sort to the end
        if (!bad && vox_keep(x, y, z, axis, lo, hi)) {
            const int i0 = (int)__fsub_rn(floorf(__fmul_rn(x, inv)), b0);
            const int i1 = (int)__fsub_rn(floorf(__fmul_rn(y, inv)), b1);
            const int i2 = (int)__fsub_rn(floorf(__fmul_rn(z, inv)), b2);
            key = (unsigned)(i0 + i1 * d0 + i2 * d01);
        }
        keys[j] = key;
        idx[j] = (int32_t)j;
    }
}

__global__ void __launch_bounds__(256) vox_head_kernel(const unsigned* __restrict__ keys, int64_t n, int* __restrict__ head) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned k = keys[i];
        head[i] = (k != 0xffffffffu && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) vox_centroid_kernel(const float* __restrict__ xyz, const unsigned* __restrict__ keys,
                                                           const int32_t* __restrict__ idx, const int* __restrict__ head,
                                                           const int* __restrict__ pos, int64_t n, float* __restrict__ out,
                                                           int64_t* __restrict__ count, VoxHeader* __restrict__ h) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (i == n - 1) {
            const int total = pos[i] + head[i];
            *count = total;
            h->n_voxels = total;
        }
        if (!head[i]) continue;
        const unsigned k = keys[i];
        float sx = 0.f, sy = 0.f, sz = 0.f;
        int cnt = 0;
        for (int64_t j = i; j < n && keys[j] == k; ++j) {  // ascending original index: the sort is stable
            const int64_t p = idx[j];
            sx = __fadd_rn(sx, xyz[p * 3]);
            sy = __fadd_rn(sy, xyz[p * 3 + 1]);
            sz = __fadd_rn(sz, xyz[p * 3 + 2]);
            ++cnt;
        }
        const float c = (float)cnt;
        const int64_t o = pos[i];
        out[o * 3] = __fdiv_rn(sx, c);
        out[o * 3 + 1] = __fdiv_rn(sy, c);
        out[o * 3 + 2] = __fdiv_rn(sz, c);
    }
}

size_t vox_sort_temp_bytes(int64_t) { return std::max(covradix::temp_bytes(), covradix::scan_temp_bytes()); }

}  // namespace

// workspace: [header 256 B][keys A][keys B][idx A][idx B][head][pos][digit counts / scan partials]
extern "C" size_t cov_voxel_grid_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return 256 + 6 * align256((size_t)n * 4) + align256(vox_sort_temp_bytes(n)) + 256;
}

extern "C" int cov_voxel_grid(const float* xyz, int64_t n, float leaf, int filter_axis, float limit_min, float limit_max,
                              float* xyz_out, int64_t* count, int32_t* info, void* ws, size_t ws_bytes, void* stream) {
    if (!xyz || !xyz_out || !count || !info || !ws || n <= 0 || !(leaf > 0.f) || filter_axis < -1 || filter_axis > 2) {
        cov_set_error("cov_voxel_grid: bad argument (n=%lld, leaf=%g, filter_axis=%d)", (long long)n, (double)leaf, filter_axis);
        return COV_ERR_ARG;
    }
    if (n >= ((int64_t)1 << 31)) {
        cov_set_error("cov_voxel_grid: %lld points exceed the int32 index range", (long long)n);
        return COV_ERR_UNSUPPORTED;
    }
    if (ws_bytes < cov_voxel_grid_workspace_bytes(n) || (((uintptr_t)ws) & 255)) {
        cov_set_error("cov_voxel_grid: workspace too small or not 256-byte aligned");
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    VoxHeader* h = reinterpret_cast<VoxHeader*>(base);
    const size_t col = align256((size_t)n * 4);
    unsigned* keys_a = reinterpret_cast<unsigned*>(base + 256);
    unsigned* keys_b = reinterpret_cast<unsigned*>(base + 256 + col);
    int32_t* idx_a = reinterpret_cast<int32_t*>(base + 256 + 2 * col);
    int32_t* idx_b = reinterpret_cast<int32_t*>(base + 256 + 3 * col);
    int* head = reinterpret_cast<int*>(base + 256 + 4 * col);
    int* pos = reinterpret_cast<int*>(base + 256 + 5 * col);
    void* temp = base + 256 + 6 * col;
    const int sms = cov_sm_count_cached();
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    vox_init_kernel<<<1, 32, 0, s>>>(h);
    vox_box_kernel<<<grid, 256, 0, s>>>(xyz, n, filter_axis, limit_min, limit_max, h);
    vox_dims_kernel<<<1, 32, 0, s>>>(h, leaf);
    vox_key_kernel<<<grid, 256, 0, s>>>(xyz, n, leaf, filter_axis, limit_min, limit_max, h, keys_a, idx_a);
    // stable sort by the full 32-bit voxel index (dropped points carry 0xffffffff and end up last): 4 passes, so the
    // sorted pairs are back in (keys_a, idx_a)
    const int where = covradix::sort_pairs(keys_a, idx_a, keys_b, idx_b, n, 0, 32, temp, sms, s);
    const unsigned* keys_sorted = where ? keys_b : keys_a;
    const int32_t* idx_sorted = where ? idx_b : idx_a;
    vox_head_kernel<<<grid, 256, 0, s>>>(keys_sorted, n, head);
    covradix::exclusive_sum(head, pos, n, temp, sms, s);
    vox_centroid_kernel<<<grid, 256, 0, s>>>(xyz, keys_sorted, idx_sorted, head, pos, n, xyz_out, count, h);
    cudaMemcpyAsync(info, &h->min_b[0], 8 * sizeof(int), cudaMemcpyDeviceToDevice, s);  // min_b, div_b, overflow, n_voxels
    return cov_check_launch("cov_voxel_grid");
}

// ================================= the pair sort on its own (tests, tools) =================================
// workspace: [keys B n u32][vals B n i32][digit counts]
extern "C" size_t cov_sort_pairs_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return 2 * align256((size_t)n * 4) + align256(covradix::temp_bytes()) + 256;
}

extern "C" int cov_sort_pairs(uint32_t* keys, int32_t* vals, int64_t n, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
                              void* stream) {
    if (!keys || !vals || !ws || n <= 0 || begin_bit < 0 || end_bit > 32 || begin_bit >= end_bit) {
        cov_set_error("cov_sort_pairs: bad argument (n=%lld, bits [%d, %d))", (long long)n, begin_bit, end_bit);
        return COV_ERR_ARG;
    }
    if (n >= ((int64_t)1 << 31)) {
        cov_set_error("cov_sort_pairs: %lld pairs exceed the 32-bit position range", (long long)n);
        return COV_ERR_UNSUPPORTED;
    }
    if (ws_bytes < cov_sort_pairs_workspace_bytes(n) || (((uintptr_t)ws) & 255)) {
        cov_set_error("cov_sort_pairs: workspace too small or not 256-byte aligned");
        return COV_ERR_WORKSPACE;
    }
    if ((((uintptr_t)keys) | ((uintptr_t)vals)) & 15) {
        cov_set_error("cov_sort_pairs: keys and vals must be 16-byte aligned (TMA bulk copies)");
        return COV_ERR_ALIGN;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    const size_t col = align256((size_t)n * 4);
    unsigned* keys_b = reinterpret_cast<unsigned*>(base);
    int32_t* vals_b = reinterpret_cast<int32_t*>(base + col);
    void* temp = base + 2 * col;
    if (covradix::sort_pairs(keys, vals, keys_b, vals_b, n, begin_bit, end_bit, temp, cov_sm_count_cached(), s) != 0) {
        cudaMemcpyAsync(keys, keys_b, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(vals, vals_b, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
    }
    return cov_check_launch("cov_sort_pairs");
}
