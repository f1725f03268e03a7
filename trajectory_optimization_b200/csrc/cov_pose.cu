// cov_pose.cu — ModelPose fused forward+backward (reference src/model.py:98-127).
//
// One pass over the cloud: each thread streams groups of four points (three LDG.128, 48 B),
// evaluates m and dm/dy in registers, writes the four observations with one STG.128 and keeps
// seven running sums (sum m, F = sum dm/dy, T = sum dm/dy x y).  Warp shuffles + one shared-memory
// step reduce them per block into a row of fp32 partials; a second tiny kernel adds the rows in
// fp64 in a fixed order (deterministic, no atomics).  HBM-bound: 16 B/point.
#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

struct PoseSums {
    float m, fx, fy, fz, tx, ty, tz;
};

__device__ __forceinline__ void pose_point(float x, float y, float z, float wgt, const float4* row, const CovConst& C,
                                           float& obs, PoseSums& S) {
    CovEval ev;
    float m = cov_vis<true>(x, y, z, row[0], row[1], row[2], row[3], C, &ev);
    float gx, gy, gz;
    cov_vis_grad(m, ev, row[0], row[1], row[2], C, gx, gy, gz);
    m = m > 0.f ? m : 0.f;  // also drops the NaN of the measure-zero h2 == -eps case
    m *= wgt; gx *= wgt; gy *= wgt; gz *= wgt;
    const float yx = x - row[5].x, yy = y - row[5].y, yz = z - row[5].z;  // lever arm about the camera centre
    obs = m;
    S.m += m;
    S.fx += gx; S.fy += gy; S.fz += gz;
    S.tx += gy * yz - gz * yy;   // (g x y)_x
    S.ty += gz * yx - gx * yz;
    S.tz += gx * yy - gy * yx;
}

template <bool HAS_W, bool HAS_OBS>
__global__ void __launch_bounds__(COV_THREADS)
cov_pose_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ weight,
                const float* __restrict__ trans, const float* __restrict__ quat, const float* __restrict__ K9,
                CovConst C, float* __restrict__ obs, float* __restrict__ partials) {
    __shared__ float4 row[COV_ROW_F4];
    __shared__ float red[COV_THREADS / 32][8];
    if (threadIdx.x == 0) cov_pose_row(trans, quat, K9, C, row);
    __syncthreads();

    PoseSums S = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int64_t ngroups = n >> 2;  // full groups of four points
    const int64_t stride = (int64_t)gridDim.x * COV_THREADS;
    const float4* __restrict__ xyz4 = reinterpret_cast<const float4*>(xyz);
    for (int64_t g = (int64_t)blockIdx.x * COV_THREADS + threadIdx.x; g < ngroups; g += stride) {
        const float4 a = __ldg(xyz4 + g * 3), b = __ldg(xyz4 + g * 3 + 1), c = __ldg(xyz4 + g * 3 + 2);
        float4 wv = make_float4(1.f, 1.f, 1.f, 1.f);
        if (HAS_W) wv = __ldg(reinterpret_cast<const float4*>(weight) + g);
        float4 o;
        pose_point(a.x, a.y, a.z, wv.x, row, C, o.x, S);
        pose_point(a.w, b.x, b.y, wv.y, row, C, o.y, S);
        pose_point(b.z, b.w, c.x, wv.z, row, C, o.z, S);
        pose_point(c.y, c.z, c.w, wv.w, row, C, o.w, S);
        if (HAS_OBS) reinterpret_cast<float4*>(obs)[g] = o;
    }
    // ragged tail (n % 4 points), handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t j = (ngroups << 2) + threadIdx.x;
        float o;
        pose_point(xyz[j * 3], xyz[j * 3 + 1], xyz[j * 3 + 2], HAS_W ? weight[j] : 1.f, row, C, o, S);
        if (HAS_OBS) obs[j] = o;
    }
    float v[7] = {S.m, S.fx, S.fy, S.fz, S.tx, S.ty, S.tz};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float r = cov_warp_sum(v[k]);
        if (lane == 0) red[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float r = 0.f;
        if (threadIdx.x < 7)
            for (int wp = 0; wp < COV_THREADS / 32; ++wp) r += red[wp][threadIdx.x];
        partials[blockIdx.x * 8 + threadIdx.x] = r;
    }
}

// acc[k] = sum over blocks of partials[b][k], fp64, fixed order.
__global__ void cov_pose_reduce_kernel(const float* __restrict__ partials, int nblocks, double* __restrict__ acc) {
    const int k = threadIdx.x & 7, part = threadIdx.x >> 3;  // 8 columns x 32 row-strides
    double s = 0.0;
    for (int b = part; b < nblocks; b += 32) s += (double)partials[b * 8 + k];
    __shared__ double sh[32][8];
    sh[part][k] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
        for (int p = 0; p < 32; ++p) t += sh[p][threadIdx.x];
        acc[threadIdx.x] = t;
    }
}

// d(sum)/dt = -F ;  d(sum)/dq~ = 2 ((0,T) (x) q) / |q~|   (tangent to the unit sphere, so the
// (I - q q^T) projection of F.normalize's backward is the identity on it).
__global__ void cov_pose_epilogue_kernel(const double* __restrict__ acc, const float* __restrict__ quat,
                                         float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double w = quat[0], x = quat[1], y = quat[2], z = quat[3];
    double n = sqrt(w * w + x * x + y * y + z * z);
    n = n > 1e-12 ? n : 1e-12;
    w /= n; x /= n; y /= n; z /= n;
    const double tx = acc[4], ty = acc[5], tz = acc[6];
    out[0] = (float)acc[0];
    out[1] = (float)(-acc[1]);
    out[2] = (float)(-acc[2]);
    out[3] = (float)(-acc[3]);
    // (0,T) (x) (w,x,y,z)
    out[4] = (float)(2.0 * (-tx * x - ty * y - tz * z) / n);
    out[5] = (float)(2.0 * (tx * w + ty * z - tz * y) / n);
    out[6] = (float)(2.0 * (ty * w - tx * z + tz * x) / n);
    out[7] = (float)(2.0 * (tz * w + tx * y - ty * x) / n);
}

int pose_grid(int64_t n) {
    const int64_t groups = (n >> 2) + 1;
    int64_t blocks = (groups + COV_THREADS - 1) / COV_THREADS;
    const int64_t cap = (int64_t)cov_sm_count_cached() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks > COV_MAX_GRID) blocks = COV_MAX_GRID;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" size_t cov_pose_workspace_bytes(int64_t n) {
    (void)n;
    return (size_t)COV_MAX_GRID * 8 * sizeof(float);
}

extern "C" int cov_pose_fused(const float* xyz, int64_t n, const float* weight, const float* trans, const float* quat,
                              const float* K, const cov_camera* cam, float* obs, double* acc, void* ws, size_t ws_bytes,
                              void* stream) {
    if (n < 0 || (n > 0 && !xyz) || !trans || !quat || !K || !cam || !acc || !ws) {
        cov_set_error("cov_pose_fused: null pointer or negative n");
        return COV_ERR_ARG;
    }
    if (ws_bytes < cov_pose_workspace_bytes(n)) {
        cov_set_error("cov_pose_fused: workspace %zu < %zu bytes", ws_bytes, cov_pose_workspace_bytes(n));
        return COV_ERR_WORKSPACE;
    }
    if (((uintptr_t)xyz | (uintptr_t)obs | (uintptr_t)weight | (uintptr_t)ws) & 15) {
        cov_set_error("cov_pose_fused: xyz/obs/weight/workspace must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    const int grid = pose_grid(n);
    float* partials = (float*)ws;
    if (weight) {
        if (obs) cov_pose_kernel<true, true><<<grid, COV_THREADS, 0, s>>>(xyz, n, weight, trans, quat, K, C, obs, partials);
        else cov_pose_kernel<true, false><<<grid, COV_THREADS, 0, s>>>(xyz, n, weight, trans, quat, K, C, obs, partials);
    } else {
        if (obs) cov_pose_kernel<false, true><<<grid, COV_THREADS, 0, s>>>(xyz, n, weight, trans, quat, K, C, obs, partials);
        else cov_pose_kernel<false, false><<<grid, COV_THREADS, 0, s>>>(xyz, n, weight, trans, quat, K, C, obs, partials);
    }
    cov_pose_reduce_kernel<<<1, 256, 0, s>>>(partials, grid, acc);
    return cov_check_launch("cov_pose_fused");
}

extern "C" int cov_pose_epilogue(const double* acc, const float* trans, const float* quat, float* out, void* stream) {
    (void)trans;
    if (!acc || !quat || !out) {
        cov_set_error("cov_pose_epilogue: null pointer");
        return COV_ERR_ARG;
    }
    cov_pose_epilogue_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, quat, out);
    return cov_check_launch("cov_pose_epilogue");
}
