// cov_tools.cu — binary frustum cull with ordered stream compaction, and the fp32 spherical flip.
//
// Cull (reference src/tools.py:176-187, src/model.py:34-39): h = K p evaluated as the fma chain
// k0*x, +k1*y, +k2*z; u = h0/h2, v = h1/h2 with IEEE division; strict comparisons.  Index set and
// order are exactly those of torch's boolean-mask gather.  Three small kernels: flags + per-block
// counts, scan of the block counts, ordered scatter.  24 B/point.
//
// Flip (reference src/tools.py:38-53): n = sqrtf(fma(z,z,fma(y,y,x*x))) (= torch CPU linalg.norm),
// R = max n * 10^param, f = (2*((R-n)*p))/n + p with every operation rounded to fp32 in that order.
#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kCullBlock = 256;

__device__ __forceinline__ void cull_tests(const float* __restrict__ xyz, int64_t j, const float* k, float wlim,
                                           float hlim, float min_d, float max_d, bool& dm, bool& fm) {
    const float x = xyz[j * 3], y = xyz[j * 3 + 1], z = xyz[j * 3 + 2];
    dm = (z > min_d) && (z < max_d);
    const float h0 = __fmaf_rn(k[2], z, __fmaf_rn(k[1], y, __fmul_rn(k[0], x)));
    const float h1 = __fmaf_rn(k[5], z, __fmaf_rn(k[4], y, __fmul_rn(k[3], x)));
    const float h2 = __fmaf_rn(k[8], z, __fmaf_rn(k[7], y, __fmul_rn(k[6], x)));
    const float u = __fdiv_rn(h0, h2), v = __fdiv_rn(h1, h2);
    fm = (h2 > 0.f) && (u > 1.f) && (u < wlim) && (v > 1.f) && (v < hlim);
}

__global__ void __launch_bounds__(kCullBlock)
cull_flags_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ K9, float wlim, float hlim,
                  float min_d, float max_d, uint8_t* __restrict__ dmask, uint8_t* __restrict__ fmask,
                  int* __restrict__ block_counts) {
    __shared__ float k[9];
    if (threadIdx.x < 9) k[threadIdx.x] = K9[threadIdx.x];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * kCullBlock + threadIdx.x;
    bool dm = false, fm = false;
    if (j < n) {
        cull_tests(xyz, j, k, wlim, hlim, min_d, max_d, dm, fm);
        dmask[j] = dm;
        fmask[j] = fm;
    }
    const int c = __syncthreads_count(dm && fm);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of block_counts in place (one block); total -> count_out.  The array is walked in chunks of 4096
// consecutive entries (4 per thread: coalesced) with a running carry.
__global__ void __launch_bounds__(1024) cull_scan_kernel(int* __restrict__ block_counts, int64_t nblocks,
                                                          int64_t* __restrict__ count_out) {
    __shared__ int wtot[32];
    __shared__ int chunk_total;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long carry = 0;
    for (int64_t base = 0; base < nblocks; base += 4096) {
        const int64_t i = base + (int64_t)t * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i + k < nblocks) ? block_counts[i + k] : 0;
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
        long long run = carry + wtot[warp] + incl - s;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < nblocks) block_counts[i + k] = (int)run;  // < 2^31 points kept per call (idx is int32)
            run += v[k];
        }
        carry += chunk_total;
        __syncthreads();  // wtot / chunk_total are rewritten by the next chunk
    }
    if (t == 0) *count_out = carry;
}

__global__ void __launch_bounds__(kCullBlock)
cull_scatter_kernel(const uint8_t* __restrict__ dmask, const uint8_t* __restrict__ fmask, int64_t n,
                    const int* __restrict__ block_offsets, int32_t* __restrict__ idx) {
    __shared__ int warp_off[kCullBlock / 32];
    const int64_t j = (int64_t)blockIdx.x * kCullBlock + threadIdx.x;
    const bool keep = (j < n) && dmask[j] && fmask[j];
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < kCullBlock / 32; ++i) {
            const int c = warp_off[i];
            warp_off[i] = run;
            run += c;
        }
    }
    __syncthreads();
    if (keep) idx[block_offsets[blockIdx.x] + warp_off[warp] + __popc(bal & ((1u << lane) - 1u))] = (int32_t)j;
}

__device__ __forceinline__ float flip_norm(float x, float y, float z) {
    return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

__global__ void __launch_bounds__(256) flip_maxnorm_kernel(const float* __restrict__ xyz, int64_t n,
                                                            unsigned* __restrict__ max_bits) {
    float mx = 0.f;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < n; j += (int64_t)gridDim.x * 256)
        mx = fmaxf(mx, flip_norm(xyz[j * 3], xyz[j * 3 + 1], xyz[j * 3 + 2]));
    const unsigned w = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, w);
}

__global__ void __launch_bounds__(256) flip_apply_kernel(const float* __restrict__ xyz, int64_t n, float scale,
                                                          float* __restrict__ radius_io, float* __restrict__ out) {
    const float radius = __fmul_rn(radius_io[1], scale);  // radius_io[1] = max norm, [0] = radius (written below)
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < n; j += (int64_t)gridDim.x * 256) {
        const float p[3] = {xyz[j * 3], xyz[j * 3 + 1], xyz[j * 3 + 2]};
        const float nr = flip_norm(p[0], p[1], p[2]);
        const float rn = __fsub_rn(radius, nr);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[j * 3 + c] = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, __fmul_rn(rn, p[c])), nr), p[c]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) radius_io[0] = radius;
}

// ---- PointCloud2 payload <-> xyz (reference src/pointcloud_utils.py:58-80,180-198,290-338) ----
// One record of `point_step` bytes per point; x, y, z are FLOAT32 (PointField datatype 7) or FLOAT64 (8) fields at
// byte offsets that need not be aligned.  Points whose x, y, z are all finite are kept (np.isfinite), in order.
__device__ __forceinline__ float pc2_field(const uint8_t* __restrict__ rec, int off, int datatype, bool& finite) {
    if (datatype == 8) {
        unsigned long long b = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) b |= (unsigned long long)rec[off + i] << (8 * i);
        finite = finite && ((b >> 52) & 0x7ffull) != 0x7ffull;  // judged on the double, before the conversion to fp32
        return (float)__longlong_as_double((long long)b);
    }
    unsigned b = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) b |= (unsigned)rec[off + i] << (8 * i);
    finite = finite && ((b >> 23) & 0xffu) != 0xffu;
    return __uint_as_float(b);
}
// x, y, z of record j; returns whether all three are finite (np.isfinite: neither NaN nor +-inf)
__device__ __forceinline__ bool pc2_load(const uint8_t* __restrict__ data, int64_t j, int step, int ox, int oy, int oz,
                                         int datatype, bool aligned4, float& x, float& y, float& z) {
    const uint8_t* rec = data + j * step;
    bool finite = true;
    if (aligned4 && datatype == 7) {  // the common layout: 4-byte aligned float fields
        x = *reinterpret_cast<const float*>(rec + ox);
        y = *reinterpret_cast<const float*>(rec + oy);
        z = *reinterpret_cast<const float*>(rec + oz);
        const float inf = __uint_as_float(0x7f800000u);
        finite = fabsf(x) < inf && fabsf(y) < inf && fabsf(z) < inf;
    } else {
        x = pc2_field(rec, ox, datatype, finite);
        y = pc2_field(rec, oy, datatype, finite);
        z = pc2_field(rec, oz, datatype, finite);
    }
    return finite;
}
__device__ __forceinline__ bool pc2_finite(float x, float y, float z) {
    const float inf = __uint_as_float(0x7f800000u);
    return fabsf(x) < inf && fabsf(y) < inf && fabsf(z) < inf;  // false for NaN and +-inf
}

__global__ void __launch_bounds__(kCullBlock)
pc2_flags_kernel(const uint8_t* __restrict__ data, int64_t n, int step, int ox, int oy, int oz, int datatype,
                 bool aligned4, int remove_nans, int* __restrict__ block_counts) {
    const int64_t j = (int64_t)blockIdx.x * kCullBlock + threadIdx.x;
    bool keep = false;
    if (j < n) {
        float x, y, z;
        const bool fin = pc2_load(data, j, step, ox, oy, oz, datatype, aligned4, x, y, z);
        keep = !remove_nans || fin;
    }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kCullBlock)
pc2_scatter_kernel(const uint8_t* __restrict__ data, int64_t n, int step, int ox, int oy, int oz, int datatype,
                   bool aligned4, int remove_nans, const int* __restrict__ block_offsets, float* __restrict__ xyz) {
    __shared__ int warp_off[kCullBlock / 32];
    const int64_t j = (int64_t)blockIdx.x * kCullBlock + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    bool keep = false;
    if (j < n) {
        const bool fin = pc2_load(data, j, step, ox, oy, oz, datatype, aligned4, x, y, z);
        keep = !remove_nans || fin;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < kCullBlock / 32; ++i) {
            const int c = warp_off[i];
            warp_off[i] = run;
            run += c;
        }
    }
    __syncthreads();
    if (keep) {
        const int64_t o = (int64_t)block_offsets[blockIdx.x] + warp_off[warp] + __popc(bal & ((1u << lane) - 1u));
        xyz[o * 3] = x;
        xyz[o * 3 + 1] = y;
        xyz[o * 3 + 2] = z;
    }
}

// xyz (+ optional 4th float per point) -> little-endian FLOAT32 records of 12 or 16 bytes; dense[0] &= all finite
__global__ void __launch_bounds__(256)
pc2_pack_kernel(const float* __restrict__ xyz, const float* __restrict__ extra, int64_t n, float* __restrict__ out,
                int* __restrict__ dense) {
    bool ok = true;
    const int per = extra ? 4 : 3;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const float x = xyz[j * 3], y = xyz[j * 3 + 1], z = xyz[j * 3 + 2];
        out[j * per] = x;
        out[j * per + 1] = y;
        out[j * per + 2] = z;
        ok = ok && pc2_finite(x, y, z);
        if (extra) {
            const float e = extra[j];
            out[j * per + 3] = e;
            ok = ok && pc2_finite(e, 0.f, 0.f);
        }
    }
    if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(dense, 0);
}

}  // namespace

extern "C" size_t cov_cull_workspace_bytes(int64_t n) {
    const int64_t nb = (n + kCullBlock - 1) / kCullBlock + 1;
    return (size_t)nb * sizeof(int);
}

extern "C" int cov_frustum_cull(const float* xyz, int64_t n, const float* K, float img_width, float img_height,
                                float min_dist, float max_dist, uint8_t* dmask, uint8_t* fmask, int32_t* idx,
                                int64_t* count, void* ws, size_t ws_bytes, void* stream) {
    if (n < 0 || !K || !count || (n > 0 && (!xyz || !dmask || !fmask || !idx || !ws))) {
        cov_set_error("cov_frustum_cull: null pointer or negative n");
        return COV_ERR_ARG;
    }
    if (n >= (int64_t)1 << 31) {
        cov_set_error("cov_frustum_cull: n >= 2^31 not supported (int32 indices)");
        return COV_ERR_UNSUPPORTED;
    }
    if (ws_bytes < cov_cull_workspace_bytes(n)) {
        cov_set_error("cov_frustum_cull: workspace too small");
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(count, 0, sizeof(int64_t), s);
        return cov_check_launch("cov_frustum_cull");
    }
    const int64_t nb = (n + kCullBlock - 1) / kCullBlock;
    int* counts = (int*)ws;
    // the comparisons `u < img_width - 1` are against the fp32 value of the Python float (exact for integers)
    cull_flags_kernel<<<(unsigned)nb, kCullBlock, 0, s>>>(xyz, n, K, (float)((double)img_width - 1.0),
                                                         (float)((double)img_height - 1.0), min_dist, max_dist, dmask,
                                                         fmask, counts);
    cull_scan_kernel<<<1, 1024, 0, s>>>(counts, nb, count);
    cull_scatter_kernel<<<(unsigned)nb, kCullBlock, 0, s>>>(dmask, fmask, n, counts, idx);
    return cov_check_launch("cov_frustum_cull");
}

extern "C" int cov_hpr_flip(const float* xyz, int64_t n, float scale, float* flipped, float* radius, void* stream) {
    if (n <= 0 || !xyz || !flipped || !radius) {
        cov_set_error("cov_hpr_flip: null pointer or empty cloud");
        return COV_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // radius[0] = R (out), radius[1] = max norm scratch
    cudaMemsetAsync(radius, 0, 2 * sizeof(float), s);
    int64_t nb = (n + 255) / 256;
    const int64_t cap = (int64_t)cov_sm_count_cached() * 8;
    if (nb > cap) nb = cap;
    flip_maxnorm_kernel<<<(unsigned)nb, 256, 0, s>>>(xyz, n, reinterpret_cast<unsigned*>(radius + 1));
    flip_apply_kernel<<<(unsigned)nb, 256, 0, s>>>(xyz, n, scale, radius, flipped);
    return cov_check_launch("cov_hpr_flip");
}

extern "C" size_t cov_pc2_workspace_bytes(int64_t n) { return cov_cull_workspace_bytes(n); }

extern "C" int cov_pc2_to_xyz(const uint8_t* data, int64_t n, int point_step, int off_x, int off_y, int off_z,
                              int datatype, int remove_nans, float* xyz, int64_t* count, void* ws, size_t ws_bytes,
                              void* stream) {
    const int fsz = datatype == 8 ? 8 : 4;
    if (n < 0 || !count || (datatype != 7 && datatype != 8) || point_step <= 0 || off_x < 0 || off_y < 0 || off_z < 0 ||
        off_x + fsz > point_step || off_y + fsz > point_step || off_z + fsz > point_step ||
        (n > 0 && (!data || !xyz || !ws))) {
        cov_set_error("cov_pc2_to_xyz: bad argument (n=%lld, point_step=%d, offsets %d/%d/%d, datatype %d: only "
                      "FLOAT32=7 and FLOAT64=8 fields are supported)", (long long)n, point_step, off_x, off_y, off_z, datatype);
        return COV_ERR_ARG;
    }
    if (n >= (int64_t)1 << 31) {
        cov_set_error("cov_pc2_to_xyz: n >= 2^31 not supported");
        return COV_ERR_UNSUPPORTED;
    }
    if (ws_bytes < cov_pc2_workspace_bytes(n)) {
        cov_set_error("cov_pc2_to_xyz: workspace too small");
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(count, 0, sizeof(int64_t), s);
        return cov_check_launch("cov_pc2_to_xyz");
    }
    const bool aligned4 = ((((uintptr_t)data) | (unsigned)point_step | (unsigned)off_x | (unsigned)off_y | (unsigned)off_z) & 3) == 0;
    const int64_t nb = (n + kCullBlock - 1) / kCullBlock;
    int* counts = (int*)ws;
    pc2_flags_kernel<<<(unsigned)nb, kCullBlock, 0, s>>>(data, n, point_step, off_x, off_y, off_z, datatype, aligned4,
                                                        remove_nans, counts);
    cull_scan_kernel<<<1, 1024, 0, s>>>(counts, nb, count);
    pc2_scatter_kernel<<<(unsigned)nb, kCullBlock, 0, s>>>(data, n, point_step, off_x, off_y, off_z, datatype, aligned4,
                                                          remove_nans, counts, xyz);
    return cov_check_launch("cov_pc2_to_xyz");
}

extern "C" int cov_xyz_to_pc2(const float* xyz, const float* extra, int64_t n, uint8_t* data, int* is_dense, void* stream) {
    if (n < 0 || !is_dense || (n > 0 && (!xyz || !data)) || (((uintptr_t)data) & 3)) {
        cov_set_error("cov_xyz_to_pc2: bad argument (null pointer, negative n or payload not 4-byte aligned)");
        return COV_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(is_dense, 1, sizeof(int), s);  // nonzero = dense; the kernel clears it on the first non-finite value
    if (n > 0) {
        int64_t nb = (n + 255) / 256;
        const int64_t cap = (int64_t)cov_sm_count_cached() * 16;
        if (nb > cap) nb = cap;
        pc2_pack_kernel<<<(unsigned)nb, 256, 0, s>>>(xyz, extra, n, reinterpret_cast<float*>(data), is_dense);
    }
    return cov_check_launch("cov_xyz_to_pc2");
}
