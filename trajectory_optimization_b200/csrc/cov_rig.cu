// cov_rig.cu — multi-camera front end: body waypoints (x, y, z, yaw) + fixed camera extrinsics -> per-camera
// poses (t, q) for the coverage kernels, and the chain rule back to the 4 body parameters.
//
// The reference optimises camera poses directly (src/model.py:66-90); its cameras come from the tf extrinsics of
// a 5-6 camera rig (src/pc_processor.py:33-39,161-165).  BASELINE's north star parametrises a waypoint as
// (X, Y, Z, yaw) plus fixed extrinsics; this is that map, as two O(W) kernels instead of ~300 small torch launches:
//   q_wc = q_z(yaw) (x) q_bc,   t_wc = xyz + R_z(yaw) t_bc
//   d/dxyz = sum_c g_t,   d/dyaw = sum_c [ g_t . (dR_z/dyaw t_bc) + g_q . (dq_z/dyaw (x) q_bc) ]
#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

// Hamilton product a (x) b, (w, x, y, z)
__device__ __forceinline__ void qmul(const float a[4], const float b[4], float o[4]) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

__global__ void cov_rig_poses_kernel(const float* __restrict__ body, int B, const float* __restrict__ rig, int C,
                                     float* __restrict__ poses, float* __restrict__ quats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    const float x = body[4 * b], y = body[4 * b + 1], z = body[4 * b + 2], yaw = body[4 * b + 3];
    float sh, ch, s, co;
    sincosf(0.5f * yaw, &sh, &ch);
    sincosf(yaw, &s, &co);
    const float* r = rig + 7 * c;
    const float qz[4] = {ch, 0.f, 0.f, sh}, qb[4] = {r[0], r[1], r[2], r[3]};
    float q[4];
    qmul(qz, qb, q);
    quats[4 * i] = q[0]; quats[4 * i + 1] = q[1]; quats[4 * i + 2] = q[2]; quats[4 * i + 3] = q[3];
    poses[3 * i] = x + co * r[4] - s * r[5];
    poses[3 * i + 1] = y + s * r[4] + co * r[5];
    poses[3 * i + 2] = z + r[6];
}

__global__ void cov_rig_poses_backward_kernel(const float* __restrict__ body, int B, const float* __restrict__ rig, int C,
                                              const float* __restrict__ g_poses, const float* __restrict__ g_quats,
                                              float scale, float* __restrict__ g_body) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float yaw = body[4 * b + 3];
    float sh, ch, s, co;
    sincosf(0.5f * yaw, &sh, &ch);
    sincosf(yaw, &s, &co);
    const float dqz[4] = {-0.5f * sh, 0.f, 0.f, 0.5f * ch};
    float gx = 0.f, gy = 0.f, gz = 0.f, gyaw = 0.f;
    for (int c = 0; c < C; ++c) {
        const int i = b * C + c;
        const float* r = rig + 7 * c;
        if (g_poses) {
            const float a0 = g_poses[3 * i], a1 = g_poses[3 * i + 1], a2 = g_poses[3 * i + 2];
            gx += a0; gy += a1; gz += a2;
            gyaw += a0 * (-s * r[4] - co * r[5]) + a1 * (co * r[4] - s * r[5]);
        }
        if (g_quats) {
            const float qb[4] = {r[0], r[1], r[2], r[3]};
            float dq[4];
            qmul(dqz, qb, dq);
            gyaw += g_quats[4 * i] * dq[0] + g_quats[4 * i + 1] * dq[1] + g_quats[4 * i + 2] * dq[2] + g_quats[4 * i + 3] * dq[3];
        }
    }
    g_body[4 * b] = scale * gx; g_body[4 * b + 1] = scale * gy; g_body[4 * b + 2] = scale * gz; g_body[4 * b + 3] = scale * gyaw;
}

}  // namespace

extern "C" int cov_rig_poses(const float* body, int n_body, const float* rig, int n_cams, float* poses, float* quats,
                             void* stream) {
    if (!body || !rig || !poses || !quats || n_body <= 0 || n_cams <= 0) {
        cov_set_error("cov_rig_poses: null pointer or empty input (n_body=%d, n_cams=%d)", n_body, n_cams);
        return COV_ERR_ARG;
    }
    const int total = n_body * n_cams;
    cov_rig_poses_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(body, n_body, rig, n_cams, poses, quats);
    return cov_check_launch("cov_rig_poses");
}

extern "C" int cov_rig_poses_backward(const float* body, int n_body, const float* rig, int n_cams, const float* g_poses,
                                      const float* g_quats, float scale, float* g_body, void* stream) {
    if (!body || !rig || !g_body || n_body <= 0 || n_cams <= 0) {
        cov_set_error("cov_rig_poses_backward: null pointer or empty input (n_body=%d, n_cams=%d)", n_body, n_cams);
        return COV_ERR_ARG;
    }
    cov_rig_poses_backward_kernel<<<(n_body + 127) / 128, 128, 0, (cudaStream_t)stream>>>(body, n_body, rig, n_cams, g_poses,
                                                                                        g_quats, scale, g_body);
    return cov_check_launch("cov_rig_poses_backward");
}
