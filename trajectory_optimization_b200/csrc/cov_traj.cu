// cov_traj.cu — ModelTraj visibility term, fused forward+backward (reference src/model.py:200-246).
//
// Data layout: the cloud stays in HBM as row-major (N,3) fp32 and is streamed once per pass;
// the W pose rows (t, R mu1, P = K R^T, normalisers) live in shared memory as 5 float4 each and are
// read with broadcast LDS.128; per-point state (log-odds sum) lives in registers.
//
// Pass A (cov_traj_minmax_kernel): per pose min_j m and max_j m.  Thread-local fmin/fmax over the
//   thread's points, one integer REDUX per warp (m >= 0, so the float order is the uint order), one
//   shared-memory atomic per warp and pose, one global atomic per block and pose.  Deterministic.
// Pass B (cov_traj_fused_kernel), per tile of 256*PPT points:
//   phase 1  every (point, pose): m, gate (m - a >= b/2  <=>  p >= 0.5, exact), warp ballot of the
//            gate stored as a pose-major bit matrix in shared memory; gated lanes add their log-odds
//            to the point's running sum in pose order (same order as the reference loop).
//            rewards_j = sigmoid(L_j) is written, G_j = r_j (1 - r_j) kept in shared memory.
//   phase 2  the bit matrix is walked pose-major: a lane owns (pose, row segment), pops its set bits,
//            re-evaluates m and dm/dy for that pair and accumulates the 8 weighted sums in registers —
//            no atomics, fixed order.  Segments of a pose are combined with xor-shuffles and added to
//            the block's per-pose accumulators in shared memory by one owner lane.
//   Block accumulators go to a [block][W][8] fp32 slab; a small kernel adds the slabs in fp64.
//   The arg-max / arg-min tie sets of the normalisation backward are rare (one point per pose unless
//   the minimum underflowed to 0, in which case their gradient is exactly negligible and skipped) and
//   go straight to the fp64 accumulator with atomics.
#include <cstdlib>

#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kWarps = COV_THREADS / 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ constexpr int tile_points(int ppt) { return COV_THREADS * ppt; }
__host__ __device__ constexpr int bit_words(int ppt) { return kWarps * ppt; }         // ballot words per pose
__host__ __device__ constexpr int bit_stride(int ppt) { return bit_words(ppt) + 4; }  // rows stay 16-byte aligned

size_t minmax_smem_bytes(int W) { return (size_t)W * (COV_ROW_F4 * sizeof(float4) + 3 * sizeof(unsigned)); }
size_t fused_smem_bytes(int W, int ppt) {
    return (size_t)W * COV_ROW_F4 * sizeof(float4) + (size_t)W * bit_stride(ppt) * sizeof(unsigned) +
           (size_t)tile_points(ppt) * 4 * sizeof(float) + (size_t)W * 8 * sizeof(float);
}

// Unweighted dm/dy and dm/dy x y of one (point, pose) into a tie-set accumulator (7 doubles).
__device__ __noinline__ void tie_accumulate(float x, float y, float z, const float4* row, CovConst C, double* acc,
                                            int slot) {
    double* dst = acc + slot;
    CovEval ev;
    const float m = cov_vis<true>(x, y, z, row[0], row[1], row[2], row[3], C, &ev);
    float gx, gy, gz;
    cov_vis_grad(m, ev, row[0], row[1], row[2], C, gx, gy, gz);
    const float yx = x - row[5].x, yy = y - row[5].y, yz = z - row[5].z;
    atomicAdd(dst + 0, (double)gx);
    atomicAdd(dst + 1, (double)gy);
    atomicAdd(dst + 2, (double)gz);
    atomicAdd(dst + 3, (double)(gy * yz - gz * yy));
    atomicAdd(dst + 4, (double)(gz * yx - gx * yz));
    atomicAdd(dst + 5, (double)(gx * yy - gy * yx));
    atomicAdd(dst + 6, 1.0);
}

// PRUNE: once this block has seen m == 0 for a pose (so the global minimum is 0), a warp skips the evaluation of
// its points for that pose when none of them can reach the block's running maximum:
//   m <= 2^-(kd q2) (1+1.3e-5)  and  q2 > qcap = (-log2(max) + 1e-4)/kd  =>  m < max.
// Skipped pairs can change neither the minimum (already 0, and m >= 0) nor the maximum: results are identical.
template <int PPT, int MINB, int U, bool PRUNE>
__global__ void __launch_bounds__(COV_THREADS, MINB)
cov_traj_minmax_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                       const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                       unsigned* __restrict__ gmin, unsigned* __restrict__ gmax, unsigned long long* __restrict__ stats) {
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    unsigned* smin = reinterpret_cast<unsigned*>(ptab + (size_t)W * COV_ROW_F4);
    unsigned* smax = smin + W;
    float* sqcap = reinterpret_cast<float*>(smax + W);
    const int tid = threadIdx.x, lane = tid & 31;
    const float inv_kd = 1.f / C.kd;
    unsigned n_iter = 0, n_full = 0;
    for (int w = tid; w < W; w += COV_THREADS) {
        cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, ptab + (size_t)w * COV_ROW_F4);
        smin[w] = 0x7f800000u;
        smax[w] = 0u;
        sqcap[w] = __uint_as_float(0x7f800000u);  // +inf: evaluate everything until a zero minimum is known
    }
    __syncthreads();
    constexpr int T = tile_points(PPT);
    const int64_t ntiles = (n + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float px[PPT], py[PPT], pz[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            int64_t j = tile * T + s * COV_THREADS + tid;
            j = j < n ? j : n - 1;  // a duplicate cannot change a min or a max
            px[s] = __ldg(xyz + j * 3);
            py[s] = __ldg(xyz + j * 3 + 1);
            pz[s] = __ldg(xyz + j * 3 + 2);
        }
        for (int w0 = 0; w0 < W; w0 += 32) {
            // lane i keeps the warp-wide min/max of pose w0+i in registers; one shared atomic per 32 poses
            unsigned keep_mn = 0x7f800000u, keep_mx = 0u;
            const int wn = (W - w0 < 32) ? (W - w0) : 32;
            for (int i = 0; i < wn; i += U) {
                float mn[U], mx[U];
                int wi[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {  // U poses in flight: U*PPT independent evaluation chains
                    wi[u] = (i + u < wn) ? (i + u) : (wn - 1);  // odd remainder: re-evaluate the last pose (harmless)
                    const float4* row = ptab + (size_t)(w0 + wi[u]) * COV_ROW_F4;
                    const float4 v3 = row[3];
                    mn[u] = __uint_as_float(0x7f800000u);  // neutral elements when the pose is skipped
                    mx[u] = 0.f;
                    if (PRUNE) {
                        float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
                        for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
                        ++n_iter;
                        if (!__any_sync(kFull, qmin <= sqcap[w0 + wi[u]])) continue;
                        ++n_full;
                    }
                    const float4 v0 = row[0], v1 = row[1], v2 = row[2];
                    float m[PPT];
#pragma unroll
                    for (int s = 0; s < PPT; ++s) m[s] = cov_vis<false>(px[s], py[s], pz[s], v0, v1, v2, v3, C, nullptr);
                    mn[u] = m[0];
                    mx[u] = m[0];
#pragma unroll
                    for (int s = 1; s + 1 < PPT; s += 2) {
                        mn[u] = fminf(mn[u], fminf(m[s], m[s + 1]));
                        mx[u] = fmaxf(mx[u], fmaxf(m[s], m[s + 1]));
                    }
                    if ((PPT & 1) == 0) {
                        mn[u] = fminf(mn[u], m[PPT - 1]);
                        mx[u] = fmaxf(mx[u], m[PPT - 1]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned umn = __reduce_min_sync(kFull, __float_as_uint(mn[u]));
                    const unsigned umx = __reduce_max_sync(kFull, __float_as_uint(mx[u]));
                    if (lane == wi[u]) {
                        keep_mn = umn;
                        keep_mx = umx;
                    }
                }
            }
            if (lane < wn) {
                atomicMin(smin + w0 + lane, keep_mn);
                atomicMax(smax + w0 + lane, keep_mx);
                if (PRUNE) {
                    // stale values are only ever larger (the maximum grows), i.e. conservative
                    const unsigned bmn = smin[w0 + lane], bmx = smax[w0 + lane];
                    if (bmn == 0u && bmx != 0u)
                        sqcap[w0 + lane] = (1e-4f - __log2f(__uint_as_float(bmx))) * inv_kd * 1.000001f;
                }
            }
        }
    }
    __syncthreads();
    for (int w = tid; w < W; w += COV_THREADS) {
        atomicMin(gmin + w, smin[w]);
        atomicMax(gmax + w, smax[w]);
    }
    if (PRUNE && lane == 0) {
        atomicAdd(stats + 2, (unsigned long long)n_iter);
        atomicAdd(stats + 3, (unsigned long long)n_full);
    }
}

__global__ void cov_minmax_init_kernel(unsigned* gmin, unsigned* gmax, int W) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < W) {
        gmin[w] = 0x7f800000u;
        gmax[w] = 0u;
    }
}

// Phase-1 body of the fused pass for U consecutive poses starting at w (U*PPT independent chains).
template <int PPT, int U, bool AMIN, bool THR5>
__device__ __forceinline__ void fused_pose_iter(int w, const float4* __restrict__ ptab, unsigned* __restrict__ bits_warp,
                                                int RS, const float (&px)[PPT], const float (&py)[PPT],
                                                const float (&pz)[PPT], float (&L)[PPT], const CovConst& C,
                                                double* __restrict__ acc, int lane) {
    float m[U][PPT];
    float mmax[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float4* row = ptab + (size_t)(w + u) * COV_ROW_F4;
        const float4 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3];
#pragma unroll
        for (int s = 0; s < PPT; ++s) m[u][s] = cov_vis<false>(px[s], py[s], pz[s], v0, v1, v2, v3, C, nullptr);
        mmax[u] = m[u][0];
#pragma unroll
        for (int s = 1; s + 1 < PPT; s += 2) mmax[u] = fmaxf(mmax[u], fmaxf(m[u][s], m[u][s + 1]));
        if ((PPT & 1) == 0) mmax[u] = fmaxf(mmax[u], m[u][PPT - 1]);
        // >= 0  <=>  some point of this lane may pass the gate (conservative threshold; lives in v5.w when pruning)
        mmax[u] -= THR5 ? row[5].w : v3.w;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float4* row = ptab + (size_t)(w + u) * COV_ROW_F4;
        unsigned bal[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) bal[s] = 0u;
        if (__any_sync(kFull, mmax[u] >= 0.f)) {  // warp-uniform; a few % of (warp, pose) iterations
            const float4 v4 = row[4];
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                const float d = __fsub_rn(m[u][s], v4.w);
                const bool act = d >= v4.x;  // exactly p >= 0.5
                bal[s] = __ballot_sync(kFull, act);
                if (act) {
                    const float p = __fmul_rn(d, v4.z);
                    const float qc = fminf(p, C.hi);
                    L[s] += COV_LN2_F * cov_lg2(qc * cov_rcp(1.f - qc));
                    if (d == v4.y) tie_accumulate(px[s], py[s], pz[s], row, C, acc, (w + u) * COV_ACC_STRIDE + 8);
                }
            }
        }
        if (lane == 0) {
            unsigned* brow = bits_warp + (size_t)(w + u) * RS;
            if (PPT == 4) *reinterpret_cast<uint4*>(brow) = make_uint4(bal[0], bal[1 % PPT], bal[2 % PPT], bal[3 % PPT]);
            else if (PPT == 2) *reinterpret_cast<uint2*>(brow) = make_uint2(bal[0], bal[1 % PPT]);
            else brow[0] = bal[0];
        }
        if (AMIN) {  // only compact clouds whose minimum did not underflow to 0 (block-uniform choice of the loop)
            const float a = row[4].w;
            if (a > 0.f) {
#pragma unroll
                for (int s = 0; s < PPT; ++s)
                    if (m[u][s] == a) tie_accumulate(px[s], py[s], pz[s], row, C, acc, (w + u) * COV_ACC_STRIDE + 15);
            }
        }
    }
}

// Pruned phase-1 body: v3.w holds qthr = (-log2(thr) + 1e-4)/kd; a pair with q2 > qthr has
// m <= 2^-(kd q2)(1+1.3e-5) < thr, so it is neither gated nor the arg-max.  The warp runs the exact body only when
// one of its 32*PPT points passes; the bit matrix was zeroed at the start of the tile.
template <int PPT>
__device__ __forceinline__ void fused_pose_iter_pruned(int w, const float4* __restrict__ ptab,
                                                       unsigned* __restrict__ bits_warp, int RS, const float (&px)[PPT],
                                                       const float (&py)[PPT], const float (&pz)[PPT], float (&L)[PPT],
                                                       const CovConst& C, double* __restrict__ acc, int lane,
                                                       unsigned& n_full) {
    const float4 v3 = ptab[(size_t)w * COV_ROW_F4 + 3];
    float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
    for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
    if (__any_sync(kFull, qmin <= v3.w)) {
        ++n_full;
        fused_pose_iter<PPT, 1, false, true>(w, ptab, bits_warp, RS, px, py, pz, L, C, acc, lane);
    }
}

template <int PPT, bool HAS_UP, int U, bool PRUNE>
__global__ void __launch_bounds__(COV_THREADS, 2)
cov_traj_fused_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                      const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                      const float* __restrict__ minmax, const float* __restrict__ upstream,
                      float* __restrict__ rewards, float* __restrict__ partials, double* __restrict__ sumr_partials,
                      double* __restrict__ acc, int seg_log2, unsigned long long* __restrict__ stats) {
    constexpr int T = tile_points(PPT);
    constexpr int NW = bit_words(PPT);
    constexpr int RS = bit_stride(PPT);
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    unsigned* bits = reinterpret_cast<unsigned*>(ptab + (size_t)W * COV_ROW_F4);
    float* xs = reinterpret_cast<float*>(bits + (size_t)W * RS);
    float* ys = xs + T;
    float* zs = ys + T;
    float* Gs = zs + T;
    float* accs = Gs + T;
    __shared__ double red[kWarps];
    __shared__ int amin_pos;  // some pose has min_j m > 0: its arg-min points carry gradient

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) amin_pos = 0;
    __syncthreads();
    for (int w = tid; w < W; w += COV_THREADS) {
        float4* row = ptab + (size_t)w * COV_ROW_F4;
        cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, row);
        const float a = minmax[w];
        const float b = __fsub_rn(minmax[W + w], a);
        const float hb = 0.5f * b;
        const float thr = (a + hb) * (1.f - 9.5367431640625e-07f);  // conservative gate threshold (exact test in the rare path)
        row[5].w = thr;
        // pruning: q2 above this cannot reach thr (thr <= 0 or NaN: never prune)
        const float qthr = (thr > 0.f) ? (float)((1e-4 - log2((double)thr)) / (double)C.kd * 1.000001) : __uint_as_float(0x7f800000u);
        row[3].w = PRUNE ? qthr : thr;
        row[4] = make_float4(hb, b, __frcp_rn(b), a);
        if (a > 0.f) amin_pos = 1;  // benign race: every writer stores 1
    }
    for (int i = tid; i < W * 8; i += COV_THREADS) accs[i] = 0.f;
    __syncthreads();

    double sum_r = 0.0;
    unsigned n_iter = 0, n_full = 0;
    const int64_t ntiles = (n + T - 1) / T;
    const int nseg = 1 << seg_log2;       // lanes that share one pose row in phase 2
    const int wps = NW >> seg_log2;       // ballot words per lane
    const int ntask = W << seg_log2;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ------------------------------ phase 1: every (point, pose) ------------------------------
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const int64_t j = tile * T + s * COV_THREADS + tid;
            valid[s] = j < n;
            // a point past the end sits 3e18 m away: m = 0 exactly, never gated, never a tie
            px[s] = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;
            py[s] = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
            pz[s] = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
            xs[s * COV_THREADS + tid] = px[s];
            ys[s * COV_THREADS + tid] = py[s];
            zs[s * COV_THREADS + tid] = pz[s];
            L[s] = 0.f;
        }
        const bool check_amin = amin_pos != 0;
        {
            unsigned* bits_warp = bits + warp * PPT;
            int w = 0;
            if (check_amin) {  // arg-min points carry gradient: every pair must be looked at
                for (; w < W; ++w) fused_pose_iter<PPT, 1, true, PRUNE>(w, ptab, bits_warp, RS, px, py, pz, L, C, acc, lane);
            } else if (PRUNE) {
                for (int wz = lane; wz < W; wz += 32) {  // this warp's columns of the gate bit matrix start at zero
                    unsigned* brow = bits_warp + (size_t)wz * RS;
#pragma unroll
                    for (int s = 0; s < PPT; ++s) brow[s] = 0u;
                }
                __syncwarp();
                for (; w < W; ++w) fused_pose_iter_pruned<PPT>(w, ptab, bits_warp, RS, px, py, pz, L, C, acc, lane, n_full);
                n_iter += W;
            } else {
                for (; w + U <= W; w += U) fused_pose_iter<PPT, U, false, false>(w, ptab, bits_warp, RS, px, py, pz, L, C, acc, lane);
                for (; w < W; ++w) fused_pose_iter<PPT, 1, false, false>(w, ptab, bits_warp, RS, px, py, pz, L, C, acc, lane);
            }
        }
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float r = 1.f / (1.f + expf(-L[s]));
            float g = r * (1.f - r);
            if (valid[s]) {
                const int64_t j = tile * T + s * COV_THREADS + tid;
                rewards[j] = r;
                sum_r += (double)r;
                if (HAS_UP) g *= upstream[j];
            }
            Gs[s * COV_THREADS + tid] = g;
        }
        __syncthreads();
        // ------------------------------ phase 2: gated pairs, pose-major ------------------------------
        for (int base = 0; base < ntask; base += COV_THREADS) {
            const int task = base + tid;
            const bool live = task < ntask;
            const int w = live ? (task >> seg_log2) : 0;
            const int seg = task & (nseg - 1);
            const float4* row = ptab + (size_t)w * COV_ROW_F4;
            const float4 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3], v4 = row[4], v5 = row[5];
            const unsigned* brow = bits + (size_t)w * RS;
            int k = seg * wps;
            const int kend = live ? k + wps : k;
            unsigned word = live ? brow[k] : 0u;
            float f0 = 0.f, f1 = 0.f, f2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, se = 0.f, sep = 0.f;
            while (true) {
                while (word == 0u && k + 1 < kend) word = brow[++k];
                if (!__any_sync(kFull, word != 0u)) break;
                if (word != 0u) {
                    const int bit = __ffs(word) - 1;
                    word &= word - 1;
                    const int local = (k % PPT) * COV_THREADS + (k / PPT) * 32 + bit;
                    CovEval ev;
                    const float x = xs[local], y = ys[local], z = zs[local];
                    const float m = cov_vis<true>(x, y, z, v0, v1, v2, v3, C, &ev);
                    const float d = __fsub_rn(m, v4.w);
                    const float p = __fdiv_rn(d, v4.y);
                    if (p <= C.hi) {  // clamp backward gate (inclusive); p >= 0.5 holds for every set bit
                        float gx, gy, gz;
                        cov_vis_grad(m, ev, v0, v1, v2, C, gx, gy, gz);
                        const float yx = x - v5.x, yy = y - v5.y, yz = z - v5.z;
                        const float e = Gs[local] / (p * (1.f - p));
                        const float om = e * v4.z;
                        f0 += om * gx; f1 += om * gy; f2 += om * gz;
                        t0 += om * (gy * yz - gz * yy);
                        t1 += om * (gz * yx - gx * yz);
                        t2 += om * (gx * yy - gy * yx);
                        se += e;
                        sep += e * p;
                    }
                }
            }
            for (int o = 1; o < nseg; o <<= 1) {
                f0 += __shfl_xor_sync(kFull, f0, o); f1 += __shfl_xor_sync(kFull, f1, o);
                f2 += __shfl_xor_sync(kFull, f2, o); t0 += __shfl_xor_sync(kFull, t0, o);
                t1 += __shfl_xor_sync(kFull, t1, o); t2 += __shfl_xor_sync(kFull, t2, o);
                se += __shfl_xor_sync(kFull, se, o); sep += __shfl_xor_sync(kFull, sep, o);
            }
            if (live && seg == 0) {
                float* a8 = accs + (size_t)w * 8;
                a8[0] += f0; a8[1] += f1; a8[2] += f2; a8[3] += t0;
                a8[4] += t1; a8[5] += t2; a8[6] += se; a8[7] += sep;
            }
        }
        __syncthreads();
    }
    float* slab = partials + (size_t)blockIdx.x * W * 8;
    for (int i = tid; i < W * 8; i += COV_THREADS) slab[i] = accs[i];
    const double ws = cov_warp_sum(sum_r);
    if (lane == 0) red[warp] = ws;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kWarps; ++i) t += red[i];
        sumr_partials[blockIdx.x] = t;
    }
    if (PRUNE && lane == 0) {
        atomicAdd(stats + 0, (unsigned long long)n_iter);
        atomicAdd(stats + 1, (unsigned long long)n_full);
    }
}

// acc[w][0..7] = sum over blocks of the fp32 slabs (fp64, fixed order); acc[W*STRIDE] = sum of rewards.
__global__ void cov_traj_reduce_kernel(const float* __restrict__ partials, const double* __restrict__ sumr_partials,
                                       int nblocks, int W, double* __restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W * 8) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += (double)partials[(size_t)b * W * 8 + i];
        acc[(size_t)(i >> 3) * COV_ACC_STRIDE + (i & 7)] = s;
    }
    if (i == 0) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += sumr_partials[b];
        acc[(size_t)W * COV_ACC_STRIDE] = s;
    }
}

// fold the min/max-path terms in and map (F, T) to (d/dt, d/dq~).
__global__ void cov_traj_epilogue_kernel(const double* __restrict__ acc, const float* __restrict__ minmax,
                                         const float* __restrict__ quats, int W, double n_total, int upstream_mode,
                                         float* __restrict__ out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w == 0) out[0] = (float)(acc[(size_t)W * COV_ACC_STRIDE] / n_total);
    if (w >= W) return;
    const double* A = acc + (size_t)w * COV_ACC_STRIDE;
    const float a = minmax[w];
    const double b = (double)__fsub_rn(minmax[W + w], a);
    const double dLdb = -A[7] / b;
    const double dLda = -A[6] / b - dLdb;
    double F[3] = {A[0], A[1], A[2]}, Tq[3] = {A[3], A[4], A[5]};
    if (A[14] > 0.0) {
        const double c = dLdb / A[14];
        for (int k = 0; k < 3; ++k) { F[k] += c * A[8 + k]; Tq[k] += c * A[11 + k]; }
    }
    if (A[21] > 0.0) {
        const double c = dLda / A[21];
        for (int k = 0; k < 3; ++k) { F[k] += c * A[15 + k]; Tq[k] += c * A[18 + k]; }
    }
    const double c0 = upstream_mode ? 1.0 : 1.0 / n_total;
    double qw = quats[4 * w], qx = quats[4 * w + 1], qy = quats[4 * w + 2], qz = quats[4 * w + 3];
    double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qn = qn > 1e-12 ? qn : 1e-12;
    qw /= qn; qx /= qn; qy /= qn; qz /= qn;
    float* gp = out + 1 + 3 * w;
    float* gq = out + 1 + 3 * W + 4 * w;
    for (int k = 0; k < 3; ++k) gp[k] = (float)(-c0 * F[k]);
    const double s = 2.0 * c0 / qn;
    gq[0] = (float)(s * (-Tq[0] * qx - Tq[1] * qy - Tq[2] * qz));
    gq[1] = (float)(s * (Tq[0] * qw + Tq[1] * qz - Tq[2] * qy));
    gq[2] = (float)(s * (Tq[1] * qw - Tq[0] * qz + Tq[2] * qx));
    gq[3] = (float)(s * (Tq[2] * qw + Tq[0] * qy - Tq[1] * qx));
}

constexpr size_t kSmemCap = 227 * 1024 - 256;

// development-only kernel-variant switch (COV_DEV_* environment variables); 0 = shipped configuration
int dev_variant(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}  // opt-in shared memory per block on sm_100, minus static use

int pick_ppt(int64_t n, int W, bool fused) {
    const int sms = cov_sm_count_cached();
    const int cand[3] = {4, 2, 1};
    for (int i = 0; i < 3; ++i) {
        const int ppt = cand[i];
        const size_t sm = fused ? fused_smem_bytes(W, ppt) : minmax_smem_bytes(W);
        if (sm > kSmemCap) continue;
        const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
        if (ntiles >= 2 * (int64_t)sms || ppt == 1) return ppt;
    }
    return 0;
}

template <typename Kern>
int grid_for(Kern kern, size_t smem, int64_t ntiles) {
    int per_sm = 0;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, COV_THREADS, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    int64_t g = (int64_t)per_sm * cov_sm_count_cached();
    if (g > ntiles) g = ntiles;
    if (g > COV_MAX_GRID) g = COV_MAX_GRID;
    return g < 1 ? 1 : (int)g;
}

int check_traj_args(const char* who, const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                    const float* K, const cov_camera* cam) {
    if (!xyz || n <= 0 || !poses || !quats || W <= 0 || !K || !cam) {
        cov_set_error("%s: null pointer, empty cloud or no poses (n=%lld, W=%d)", who, (long long)n, W);
        return COV_ERR_ARG;
    }
    if (W > cov_traj_max_poses()) {
        cov_set_error("%s: %d poses exceed the shared-memory pose table (max %d)", who, W, cov_traj_max_poses());
        return COV_ERR_UNSUPPORTED;
    }
    return COV_OK;
}

}  // namespace

extern "C" int cov_traj_max_poses(void) {
    int w = 1;
    while (fused_smem_bytes(w + 1, 1) <= kSmemCap) ++w;
    return w;
}

extern "C" size_t cov_traj_workspace_bytes(int64_t n, int n_poses) {
    (void)n;
    if (n_poses < 1) n_poses = 1;
    return (size_t)COV_MAX_GRID * ((size_t)n_poses * 8 * sizeof(float) + sizeof(double));
}

extern "C" int cov_traj_minmax(const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                               const float* K, const cov_camera* cam, float* minmax, void* stream) {
    int rc = check_traj_args("cov_traj_minmax", xyz, n, poses, quats, W, K, cam);
    if (rc) return rc;
    if (!minmax) {
        cov_set_error("cov_traj_minmax: null minmax");
        return COV_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    unsigned* gmin = reinterpret_cast<unsigned*>(minmax);
    unsigned* gmax = gmin + W;
    cov_minmax_init_kernel<<<(W + 255) / 256, 256, 0, s>>>(gmin, gmax, W);
    int ppt = pick_ppt(n, W, false);
    if (ppt == 4 && (n + tile_points(8) - 1) / tile_points(8) >= 4 * (int64_t)cov_sm_count_cached()) ppt = 8;
    const size_t smem = minmax_smem_bytes(W);
    const int mm_variant = dev_variant("COV_DEV_MM", 0);
    const int eff_ppt = (ppt == 8) ? ((mm_variant == 0 || mm_variant == 4) ? 8 : (mm_variant == 5 ? 2 : 4)) : (ppt >= 4 ? 4 : ppt);
    const int64_t ntiles = (n + tile_points(eff_ppt) - 1) / tile_points(eff_ppt);
#define LAUNCH_MM(P, B, U)                                                                                        \
    {                                                                                                             \
        if (prune) {                                                                                              \
            const int grid = grid_for(cov_traj_minmax_kernel<P, B, 1, true>, smem, ntiles);                       \
            cov_traj_minmax_kernel<P, B, 1, true><<<grid, COV_THREADS, smem, s>>>(xyz, n, poses, quats, W, K, C,  \
                                                                                  gmin, gmax, stats);             \
        } else {                                                                                                  \
            const int grid = grid_for(cov_traj_minmax_kernel<P, B, U, false>, smem, ntiles);                      \
            cov_traj_minmax_kernel<P, B, U, false><<<grid, COV_THREADS, smem, s>>>(xyz, n, poses, quats, W, K, C, \
                                                                                   gmin, gmax, stats);            \
        }                                                                                                         \
    }
    const bool prune = cov_pruning_enabled() != 0;
    unsigned long long* stats = cov_stats_device_ptr();
    const int variant = dev_variant("COV_DEV_MM", 0);
    if (ppt == 8 && variant == 0) LAUNCH_MM(8, 2, 2)
    else if (ppt == 8 && variant == 1) LAUNCH_MM(4, 3, 1)
    else if (ppt == 8 && variant == 2) LAUNCH_MM(4, 2, 2)
    else if (ppt == 8 && variant == 3) LAUNCH_MM(4, 3, 2)
    else if (ppt == 8 && variant == 4) LAUNCH_MM(8, 2, 1)
    else if (ppt == 8 && variant == 5) LAUNCH_MM(2, 3, 4)
    else if (ppt >= 4) LAUNCH_MM(4, 2, 1) else if (ppt == 2) LAUNCH_MM(2, 2, 1) else LAUNCH_MM(1, 2, 1)
#undef LAUNCH_MM
    return cov_check_launch("cov_traj_minmax");
}

extern "C" int cov_traj_fused(const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                              const float* K, const cov_camera* cam, const float* minmax, const float* upstream,
                              float* rewards, double* acc, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_traj_args("cov_traj_fused", xyz, n, poses, quats, W, K, cam);
    if (rc) return rc;
    if (!minmax || !rewards || !acc || !ws) {
        cov_set_error("cov_traj_fused: null minmax/rewards/acc/workspace");
        return COV_ERR_ARG;
    }
    if (ws_bytes < cov_traj_workspace_bytes(n, W)) {
        cov_set_error("cov_traj_fused: workspace %zu < %zu bytes", ws_bytes, cov_traj_workspace_bytes(n, W));
        return COV_ERR_WORKSPACE;
    }
    if (((uintptr_t)ws) & 15) {
        cov_set_error("cov_traj_fused: workspace must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    int ppt = pick_ppt(n, W, true);
    const int fvariant = dev_variant("COV_DEV_F", 0);
    if (ppt == 4 && (fvariant == 2 || fvariant == 3)) ppt = 2;
    if (ppt == 0) {
        cov_set_error("cov_traj_fused: %d poses do not fit in shared memory", W);
        return COV_ERR_UNSUPPORTED;
    }
    const size_t smem = fused_smem_bytes(W, ppt);
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
    // phase-2 parallelism: split each pose row over 2^seg_log2 lanes until there are >= 2 tasks per thread
    int seg_log2 = 0;
    while ((W << seg_log2) < 2 * COV_THREADS && (2 << seg_log2) <= bit_words(ppt) && seg_log2 < 5) ++seg_log2;
    double* sumr = reinterpret_cast<double*>(ws);
    float* partials = reinterpret_cast<float*>(sumr + COV_MAX_GRID);
    cudaMemsetAsync(acc, 0, ((size_t)W * COV_ACC_STRIDE + 1) * sizeof(double), s);
    int grid = 1;
#define LAUNCH_F(P, UP, U)                                                                                        \
    {                                                                                                             \
        if (prune) {                                                                                              \
            grid = grid_for(cov_traj_fused_kernel<P, UP, 1, true>, smem, ntiles);                                 \
            cov_traj_fused_kernel<P, UP, 1, true><<<grid, COV_THREADS, smem, s>>>(                                \
                xyz, n, poses, quats, W, K, C, minmax, upstream, rewards, partials, sumr, acc, seg_log2, stats);  \
        } else {                                                                                                  \
            grid = grid_for(cov_traj_fused_kernel<P, UP, U, false>, smem, ntiles);                                \
            cov_traj_fused_kernel<P, UP, U, false><<<grid, COV_THREADS, smem, s>>>(                               \
                xyz, n, poses, quats, W, K, C, minmax, upstream, rewards, partials, sumr, acc, seg_log2, stats);  \
        }                                                                                                         \
    }
    const bool prune = cov_pruning_enabled() != 0;
    unsigned long long* stats = cov_stats_device_ptr();
    if (upstream) {
        if (ppt == 4) LAUNCH_F(4, true, 1) else if (ppt == 2) LAUNCH_F(2, true, 1) else LAUNCH_F(1, true, 1)
    } else {
        if (ppt == 4 && fvariant == 1) LAUNCH_F(4, false, 2)
        else if (ppt == 2 && fvariant == 2) LAUNCH_F(2, false, 2)
        else if (ppt == 2 && fvariant == 3) LAUNCH_F(2, false, 4)
        else if (ppt == 4) LAUNCH_F(4, false, 1) else if (ppt == 2) LAUNCH_F(2, false, 1) else LAUNCH_F(1, false, 1)
    }
#undef LAUNCH_F
    cov_traj_reduce_kernel<<<(W * 8 + 255) / 256, 256, 0, s>>>(partials, sumr, grid, W, acc);
    return cov_check_launch("cov_traj_fused");
}

extern "C" int cov_traj_epilogue(const double* acc, const float* minmax, const float* quats, int W, int64_t n_total,
                                 int upstream_mode, float* out, void* stream) {
    if (!acc || !minmax || !quats || !out || W <= 0 || n_total <= 0) {
        cov_set_error("cov_traj_epilogue: bad argument");
        return COV_ERR_ARG;
    }
    cov_traj_epilogue_kernel<<<(W + 127) / 128, 128, 0, (cudaStream_t)stream>>>(acc, minmax, quats, W, (double)n_total,
                                                                               upstream_mode, out);
    return cov_check_launch("cov_traj_epilogue");
}
