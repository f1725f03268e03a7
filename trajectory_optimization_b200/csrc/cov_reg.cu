// cov_reg.cu — the O(W) regularisers of ModelTraj.criterion and their gradients in one launch
// (reference src/model.py:244-260 with length_calc :135-139 and mean_angle_calc :142-155):
//   l2     = |poses[0] - poses0[0]|
//   smooth = w_s / (mean_i arccos( (ab.ac) / (|ab||ac| + eps) ) + eps),  ab = P[i-1]-P[i], ac = P[i+1]-P[i], 0 < i < W-1
//   length = w_l * | sum_i |P[i+1]-P[i]|  -  sum_i |P0[i+1]-P0[i]| |
// Once the coverage term is fused, these ~45 small torch launches (and as many again in backward) dominate the step on
// small clouds.  One block, fp64 arithmetic; the gradients follow torch's conventions where the
// functions are not differentiable: d|v|/dv = 0 at v = 0 and d|x|/dx = 0 at x = 0.
#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kRegThreads = 256;

__device__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kRegThreads / 32; ++i) t += sh[i];
    return t;
}

struct Angle {  // angle at waypoint i and its derivatives w.r.t. P[i-1], P[i+1] (the one w.r.t. P[i] is minus their sum)
    double ang, da[3], dc[3];
};

__device__ Angle angle_at(const float* __restrict__ P, int i, double eps) {
    Angle r;
    double ab[3], ac[3];
    for (int k = 0; k < 3; ++k) {
        ab[k] = (double)P[3 * (i - 1) + k] - (double)P[3 * i + k];
        ac[k] = (double)P[3 * (i + 1) + k] - (double)P[3 * i + k];
    }
    const double na = sqrt(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2]);
    const double nc = sqrt(ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2]);
    const double dot = ab[0] * ac[0] + ab[1] * ac[1] + ab[2] * ac[2];
    const double D = na * nc + eps;
    const double c = dot / D;
    r.ang = acos(c);
    const double k = -1.0 / sqrt(1.0 - c * c);  // d arccos / dc (inf at |c| = 1, as in torch)
    for (int q = 0; q < 3; ++q) {
        // dc/dab = ac/D - c * nc * (ab/na) / D   (d|ab|/dab = ab/na, 0 at ab = 0), symmetrically for ac
        const double ua = na > 0.0 ? ab[q] / na : 0.0, uc = nc > 0.0 ? ac[q] / nc : 0.0;
        r.da[q] = k * (ac[q] / D - c * nc * ua / D);
        r.dc[q] = k * (ab[q] / D - c * na * uc / D);
    }
    return r;
}

// out: [0] l2, [1] smooth, [2] length, then three (W,3) gradient arrays in that order
__global__ void __launch_bounds__(kRegThreads)
cov_traj_regularizers_kernel(const float* __restrict__ P, const float* __restrict__ P0, int W, double sw, double lw,
                             double eps, float* __restrict__ out) {
    __shared__ double sh[kRegThreads / 32];
    const int tid = threadIdx.x;
    double len = 0.0, len0 = 0.0, ang = 0.0;
    for (int i = tid; i + 1 < W; i += kRegThreads) {
        double s = 0.0, s0 = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double d = (double)P[3 * (i + 1) + k] - (double)P[3 * i + k];
            const double d0 = (double)P0[3 * (i + 1) + k] - (double)P0[3 * i + k];
            s += d * d;
            s0 += d0 * d0;
        }
        len += sqrt(s);
        len0 += sqrt(s0);
    }
    for (int i = 1 + tid; i + 1 < W; i += kRegThreads) ang += angle_at(P, i, eps).ang;
    len = block_sum(len, sh);
    len0 = block_sum(len0, sh);
    ang = block_sum(ang, sh);
    const double mean_angle = ang / (double)(W - 2);
    double l2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double d = (double)P[k] - (double)P0[k];
        l2 += d * d;
    }
    l2 = sqrt(l2);
    const double dlen = len - len0;
    if (tid == 0) {
        out[0] = (float)l2;
        out[1] = (float)(sw / (mean_angle + eps));
        out[2] = (float)(lw * fabs(dlen));
    }
    const double ksm = -sw / ((mean_angle + eps) * (mean_angle + eps)) / (double)(W - 2);
    const double klen = lw * (dlen > 0.0 ? 1.0 : (dlen < 0.0 ? -1.0 : 0.0));
    float* g_l2 = out + 3;
    float* g_sm = g_l2 + 3 * W;
    float* g_len = g_sm + 3 * W;
    for (int j = tid; j < W; j += kRegThreads) {
        double gs[3] = {0.0, 0.0, 0.0}, gl[3] = {0.0, 0.0, 0.0};
        if (j >= 2) {                       // P[j] is the "c" end of the angle at j-1
            const Angle a = angle_at(P, j - 1, eps);
            for (int k = 0; k < 3; ++k) gs[k] += a.dc[k];
        }
        if (j >= 1 && j + 1 < W) {          // the apex of the angle at j
            const Angle a = angle_at(P, j, eps);
            for (int k = 0; k < 3; ++k) gs[k] -= a.da[k] + a.dc[k];
        }
        if (j + 2 < W) {                    // the "b" end of the angle at j+1
            const Angle a = angle_at(P, j + 1, eps);
            for (int k = 0; k < 3; ++k) gs[k] += a.da[k];
        }
        if (j >= 1) {                       // segment (j-1, j)
            double d[3], s = 0.0;
            for (int k = 0; k < 3; ++k) { d[k] = (double)P[3 * j + k] - (double)P[3 * (j - 1) + k]; s += d[k] * d[k]; }
            s = sqrt(s);
            if (s > 0.0) for (int k = 0; k < 3; ++k) gl[k] += d[k] / s;
        }
        if (j + 1 < W) {                    // segment (j, j+1)
            double d[3], s = 0.0;
            for (int k = 0; k < 3; ++k) { d[k] = (double)P[3 * (j + 1) + k] - (double)P[3 * j + k]; s += d[k] * d[k]; }
            s = sqrt(s);
            if (s > 0.0) for (int k = 0; k < 3; ++k) gl[k] -= d[k] / s;
        }
        for (int k = 0; k < 3; ++k) {
            g_sm[3 * j + k] = (float)(ksm * gs[k]);
            g_len[3 * j + k] = (float)(klen * gl[k]);
            g_l2[3 * j + k] = (j == 0 && l2 > 0.0) ? (float)(((double)P[k] - (double)P0[k]) / l2) : 0.f;
        }
    }
}

}  // namespace

extern "C" int cov_traj_regularizers(const float* poses, const float* poses0, int W, float smoothness_weight,
                                     float traj_length_weight, float eps, float* out, void* stream) {
    if (!poses || !poses0 || !out || W < 3) {
        cov_set_error("cov_traj_regularizers: null pointer or fewer than 3 waypoints (W=%d)", W);
        return COV_ERR_ARG;
    }
    cov_traj_regularizers_kernel<<<1, kRegThreads, 0, (cudaStream_t)stream>>>(poses, poses0, W, (double)smoothness_weight,
                                                                             (double)traj_length_weight, (double)eps, out);
    return cov_check_launch("cov_traj_regularizers");
}
