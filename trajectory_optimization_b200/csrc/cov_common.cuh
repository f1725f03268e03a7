// cov_common.cuh — device math shared by every coverage kernel (sm_100a).
//
// One (point, pose) visibility evaluation, written with explicit round-to-nearest
// intrinsics so that every kernel that instantiates it (min/max pass, fused pass, the
// gradient walk, the sweep) produces BIT-IDENTICAL m for the same inputs: the tie tests
// of the min/max-normalisation backward (m - a == b, m == a) rely on that.
//
// Math (reference src/model.py:13-57), restated in the world frame:
//   y = x - t,  h = P y with P = K R^T,  e = y - R mu1  (so |e|^2 = |c - mu1|^2, c = R^T y)
//   zi = 1/(h2+eps), u = h0 zi, v = h1 zi, s = sigmoid(h2)
//   m  = s * 2^-(kd |e|^2 + kf ((u/W-.5)^2 + (v/H-.5)^2)),  kd = log2e/(2 sigma^2), kf = log2e/2
// The three Gaussians share one ex2, and the sigmoid and 1/(h2+eps) share one rcp:
//   r = 1/((1+2^(-h2 log2e)) (h2+eps));  zi = r (1+..);  s = r (h2+eps)       -> 3 MUFU per eval.
// Gradient in the world frame (verified against fp64 autograd in tests/test_oracle_golden.py):
//   dm/dy = m [ -e/sigma^2 + A P0 + B P1 + (1-s - A u - B v) P2 ],  A = -(u/W-.5)/W zi, B likewise
//   dm/dtheta (world-frame rotation of the camera about its centre) = dm/dy x y.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define COV_LOG2E_F 1.4426950408889634f
#define COV_LN2_F 0.6931471805599453f
#define COV_THREADS 256
#define COV_MAX_GRID 1024

// Per-launch constants, computed on the host from cov_camera.
struct CovConst {
    float eps;      // added to the depth in the two projections
    float kd;       // 0.5*log2(e)/sigma^2
    float kf;       // 0.5*log2(e)
    float inv_w;    // 1/img_width
    float inv_h;    // 1/img_height
    float inv_s2;   // 1/sigma^2
    float mu;       // (min_dist+max_dist)/2
    float hi;       // fp32(1 - eps): upper clip of the normalised observation
};

// One pose row = 5 float4 in shared memory (broadcast LDS.128):
//   v0 = (tx, ty, tz, rmx)  v1 = (rmy, rmz, p00, p01)  v2 = (p02, p10, p11, p12)
//   v3 = (p20, p21, p22, a) v4 = (hb, b, 1/b, unused)      a = min_j m, b = max_j m - a, hb = b/2
#define COV_ROW_F4 5

__device__ __forceinline__ float cov_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cov_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cov_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Build one pose row from (t, q~, K): R(q) camera->world with q = q~/max(|q~|,1e-12)
// (F.normalize, src/model.py:53), P = K R^T, rm = R (mu,mu,mu).  Computed in fp64, stored fp32.
__device__ inline void cov_pose_row(const float* __restrict__ t3, const float* __restrict__ q4,
                                    const float* __restrict__ K9, float mu, float4* __restrict__ row) {
    double w = q4[0], x = q4[1], y = q4[2], z = q4[3];
    double n = sqrt(w * w + x * x + y * y + z * z);
    n = n > 1e-12 ? n : 1e-12;
    w /= n; x /= n; y /= n; z /= n;
    double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
                      {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
                      {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    float P[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            P[i * 3 + j] = (float)((double)K9[i * 3 + 0] * R[j][0] + (double)K9[i * 3 + 1] * R[j][1] +
                                   (double)K9[i * 3 + 2] * R[j][2]);
    float rm[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rm[i] = (float)((double)mu * (R[i][0] + R[i][1] + R[i][2]));
    row[0] = make_float4(t3[0], t3[1], t3[2], rm[0]);
    row[1] = make_float4(rm[1], rm[2], P[0], P[1]);
    row[2] = make_float4(P[2], P[3], P[4], P[5]);
    row[3] = make_float4(P[6], P[7], P[8], 0.f);
    row[4] = make_float4(0.f, 0.f, 0.f, 0.f);
}

struct CovEval {  // intermediates the gradient needs
    float yx, yy, yz, ex, ey, ez, zi, s, e2z, u, v, du, dv;
};

// m for one (point, pose).  WANT = true also fills the intermediates.
template <bool WANT>
__device__ __forceinline__ float cov_vis(float x, float y, float z, const float4& v0, const float4& v1,
                                         const float4& v2, const float4& v3, const CovConst& C, CovEval* ev) {
    const float yx = __fsub_rn(x, v0.x), yy = __fsub_rn(y, v0.y), yz = __fsub_rn(z, v0.z);
    const float ex = __fsub_rn(yx, v0.w), ey = __fsub_rn(yy, v1.x), ez = __fsub_rn(yz, v1.y);
    const float q2 = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
    const float h0 = __fmaf_rn(v2.x, yz, __fmaf_rn(v1.w, yy, __fmul_rn(v1.z, yx)));
    const float h1 = __fmaf_rn(v2.w, yz, __fmaf_rn(v2.z, yy, __fmul_rn(v2.y, yx)));
    const float h2 = __fmaf_rn(v3.z, yz, __fmaf_rn(v3.y, yy, __fmul_rn(v3.x, yx)));
    const float den = __fadd_rn(h2, C.eps);
    const float e2z = cov_ex2(fminf(__fmul_rn(h2, -COV_LOG2E_F), 80.f));  // exp(-h2), capped
    const float opz = __fadd_rn(1.f, e2z);
    const float r = cov_rcp(__fmul_rn(opz, den));
    const float zi = __fmul_rn(r, opz);   // 1/(h2+eps)
    const float s = __fmul_rn(r, den);    // sigmoid(h2)
    const float u = __fmul_rn(h0, zi), v = __fmul_rn(h1, zi);
    const float du = __fmaf_rn(u, C.inv_w, -0.5f), dv = __fmaf_rn(v, C.inv_h, -0.5f);
    const float tq = __fmaf_rn(dv, dv, __fmul_rn(du, du));
    const float Q = __fmaf_rn(q2, C.kd, __fmul_rn(tq, C.kf));
    const float E = cov_ex2(-Q);
    if (WANT) {
        ev->yx = yx; ev->yy = yy; ev->yz = yz; ev->ex = ex; ev->ey = ey; ev->ez = ez;
        ev->zi = zi; ev->s = s; ev->e2z = e2z; ev->u = u; ev->v = v; ev->du = du; ev->dv = dv;
    }
    return __fmul_rn(E, s);
}

// dm/dy (world frame) from the intermediates; g = 0 when m is not a positive finite number.
__device__ __forceinline__ void cov_vis_grad(float m, const CovEval& ev, const float4& v1, const float4& v2,
                                             const float4& v3, const CovConst& C, float& gx, float& gy, float& gz) {
    const float oms = ev.e2z * ev.s;                  // 1 - sigmoid(h2)
    const float A = -(ev.du * C.inv_w) * ev.zi;       // coefficient of P row 0
    const float B = -(ev.dv * C.inv_h) * ev.zi;       // coefficient of P row 1
    const float Cz = oms - A * ev.u - B * ev.v;       // coefficient of P row 2
    const float ax = -ev.ex * C.inv_s2 + A * v1.z + B * v2.y + Cz * v3.x;
    const float ay = -ev.ey * C.inv_s2 + A * v1.w + B * v2.z + Cz * v3.y;
    const float az = -ev.ez * C.inv_s2 + A * v2.x + B * v2.w + Cz * v3.z;
    const bool ok = m > 0.f;
    gx = ok ? m * ax : 0.f;
    gy = ok ? m * ay : 0.f;
    gz = ok ? m * az : 0.f;
}

__device__ __forceinline__ float cov_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double cov_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- host-side helpers (cov_api.cu) ----
void cov_set_error(const char* fmt, ...);
int cov_check_launch(const char* what);
CovConst cov_make_const(const struct cov_camera* cam);
int cov_sm_count_cached();
