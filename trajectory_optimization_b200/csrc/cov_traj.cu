// cov_traj.cu — ModelTraj visibility term, fused forward+backward (reference src/model.py:200-246).
//
// Data layout: the cloud stays in HBM as row-major (N,3) fp32; the W pose rows (constants of cov_common.cuh) live
// in shared memory as 6 float4 each and are read with broadcast LDS.128; per-point state lives in registers.
//
// Two passes per objective evaluation (the min/max normalisation of src/model.py:226-227 needs the extrema first):
//   pass A  cov_traj_minmax   per pose min_j m and max_j m
//   pass B  cov_traj_fused    rewards_j = sigmoid(sum_w logit(clip(p_jw))) + the gradient accumulators
//
// Each pass exists in two forms that give bit-identical normalisers and rewards:
//
// DENSE (cov_traj_opts.dense, and every cloud below 65 536 points): every (point, pose) pair is evaluated, on packed
//   fp32 pairs — one point against TWO poses per instruction stream (cov_vis2p, FFMA2); bound by the FP32 and MUFU pipes.
//   pass A: thread-local fmin/fmax over the thread's points, one integer REDUX per warp (m >= 0, so the float order
//           is the uint order), one shared atomic per warp and 32 poses, one global atomic per block and pose.
//   pass B, per tile of 256*PPT points:
//     phase 1  every pair: m, one conservative vote per pose pair; when it passes: the exact gate (m - a >= b/2  <=>
//              p >= 0.5), warp ballots of the gate into a pose-major bit matrix in shared memory (non-zero words only),
//              gated lanes add their log-odds to the point's running sum in pose order.
//     phase 2  the bit matrix is walked pose-major: a lane owns (pose, row segment), pops its set bits, re-evaluates
//              m and dm/dx for that pair and accumulates the 8 weighted sums in registers — fixed order within a block.
//
// PRUNED (default): cull -> list -> evaluate, on boxes of 128 consecutive points (cov_tile_boxes; tight when the
//   cloud is Morton-ordered by cov_spatial_sort, valid for any order).  A pair whose distance Gaussian alone bounds m
//   below what can matter is never evaluated:  m <= 2^-(kd q2)(1+1.3e-5)  and  q2 >= box bound > qcap  =>  m < bound.
//     pass A: bound = a lower bound of max_j m from a strided sample of the cloud (seed launch), valid once the
//             minimum is known to be exactly 0 (skipped points have m >= 0);
//     pass B: bound = the conservative gate threshold (a + b/2)(1 - 2^-20): a pair below it is neither gated nor
//             the arg-max, adds logit(1/2) = 0 to the log-odds sum and nothing to the gradient.
//   cov_traj_prepare_kernel (A) / cov_traj_table_kernel (B)   pose table; A also evaluates a strided sample of the
//                        cloud (lower bounds of the maxima, which minima are exactly 0); B zeroes the accumulators
//   cov_cull_kernel      one warp per group of 8 tiles: boxes against every pose -> per-tile pose bit mask, and the
//                        ascending work list of the tiles with a non-empty mask, positions from a decoupled
//                        look-back scan over the blocks' counts (deterministic; no second kernel); in pass B the
//                        same blocks pre-fill the rewards with 1/2 (what every unlisted point gets, exactly)
//   *_tiles_kernel       persistent blocks walk the work list; tile points, boxes and mask arrive together through a
//                        double-buffered TMA bulk copy (one mbarrier per buffer); a warp evaluates a listed pose only
//                        if its own 128-point box and then one of its points pass the same test.  Pass B queues its
//                        gated (point, pose) pairs per warp and differentiates them 32 at a time with every lane busy.
//   Tiles that are not listed are never read: their rewards are the pre-filled 1/2 (exact), their sum is 0.5 * count.
//
// Block accumulators are added to the caller's fp64 accumulator rows with fp64 atomics (the order of those additions is
// the only thing that differs between runs: ~1e-16 relative).  The arg-max / arg-min tie sets of the normalisation
// backward are rare (one point per pose unless the minimum underflowed to 0, in which case their gradient is exactly
// negligible and skipped) and go straight to the fp64 accumulator as well.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kWarps = COV_THREADS / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaskWords = 64;   // pose bit mask of a tile: covers cov_traj_max_poses() <= 2048
constexpr int kBoxPts = COV_BOX_POINTS;  // points per precomputed bounding box

__host__ __device__ constexpr int tile_points(int ppt) { return COV_THREADS * ppt; }
__host__ __device__ constexpr int bit_words(int ppt) { return kWarps * ppt; }         // ballot words per pose
__host__ __device__ constexpr int bit_stride(int ppt) { return bit_words(ppt) + 4; }  // dense rows: padded, 16-byte aligned
__host__ __device__ constexpr int tile_boxes(int ppt) { return tile_points(ppt) / kBoxPts; }
__host__ __device__ inline int mask_stride_words(int W) { return (((W + 31) >> 5) + 3) & ~3; }  // 16-byte multiple

size_t minmax_smem_bytes(int W) {  // pose pairs (COV_PAIR_F4 float4 each) + block minima and maxima
    return (size_t)((W + 1) / 2) * (COV_PAIR_F4 * sizeof(float4) + 4 * sizeof(unsigned));
}
// fused pass: pose table, then (128-byte aligned) the tile stage(s), G_j, gate bits, accumulators, gated-pose list
__host__ __device__ inline int fused_stage_offset_floats(int W) {  // W rounded up to whole pose pairs (dense kernel)
    return (int)((((size_t)((W + 1) & ~1) * COV_ROW_F4 * 16 + 127) & ~(size_t)127) / 4);
}
size_t fused_smem_bytes(int W, int ppt) {  // dense pass B: pose table | tile points | G_j | gate bits | block accumulators
    return (size_t)fused_stage_offset_floats(W) * 4 + (size_t)tile_points(ppt) * 12 + (size_t)tile_points(ppt) * 4 +
           (size_t)W * bit_stride(ppt) * sizeof(unsigned) + (size_t)W * 8 * sizeof(float);
}

// Unweighted dm/dy and dm/dy x y of one (point, pose) into a tie-set accumulator (7 doubles: F, T, count).  Rare (one
// point per pose and step), so these are real calls; they take the pose INDEX (or its row in the global table) and find
// the constants themselves — passing loaded rows would keep them live across the callers' hot loops.
// The dense pass B keeps its table in pose PAIRS at the start of dynamic shared memory (COV_PAIR_F4, cov_common.cuh).
__device__ __noinline__ void tie_accumulate_pair(float x, float y, float z, int w, CovConst C, double* acc, int slot) {
    extern __shared__ float4 smem4[];
    const float4* pair = smem4 + (size_t)(w >> 1) * COV_PAIR_F4;
    const int h = w & 1;
    const float4 v0 = cov_pair_row(pair, h, 0), v1 = cov_pair_row(pair, h, 1), v2 = cov_pair_row(pair, h, 2),
                 v3 = cov_pair_row(pair, h, 3), v5 = pair[10 + h];
    double* dst = acc + slot;
    CovEval ev;
    const float m = cov_vis<true>(x, y, z, v0, v1, v2, v3, C, &ev);
    float gx, gy, gz;
    cov_vis_grad(m, ev, v0, v1, v2, C, gx, gy, gz);
    const float yx = x - v5.x, yy = y - v5.y, yz = z - v5.z;
    atomicAdd(dst + 0, (double)gx);
    atomicAdd(dst + 1, (double)gy);
    atomicAdd(dst + 2, (double)gz);
    atomicAdd(dst + 3, (double)(gy * yz - gz * yy));
    atomicAdd(dst + 4, (double)(gz * yx - gx * yz));
    atomicAdd(dst + 5, (double)(gx * yy - gy * yx));
    atomicAdd(dst + 6, 1.0);
}

// Same for the pruned pass B, whose pose table stays in global memory (read through L1, a handful of rows per tile).
__device__ __noinline__ void tie_accumulate_g(float x, float y, float z, const float4* __restrict__ row, CovConst C,
                                              double* dst) {
    const float4 v0 = __ldg(row), v1 = __ldg(row + 1), v2 = __ldg(row + 2), v3 = __ldg(row + 3), v5 = __ldg(row + 5);
    CovEval ev;
    const float m = cov_vis<true>(x, y, z, v0, v1, v2, v3, C, &ev);
    float gx, gy, gz;
    cov_vis_grad(m, ev, v0, v1, v2, C, gx, gy, gz);
    const float yx = x - v5.x, yy = y - v5.y, yz = z - v5.z;
    atomicAdd(dst + 0, (double)gx);
    atomicAdd(dst + 1, (double)gy);
    atomicAdd(dst + 2, (double)gz);
    atomicAdd(dst + 3, (double)(gy * yz - gz * yy));
    atomicAdd(dst + 4, (double)(gz * yx - gx * yz));
    atomicAdd(dst + 5, (double)(gx * yy - gy * yx));
    atomicAdd(dst + 6, 1.0);
}

// ---- pruning helpers ------------------------------------------------------------------------------------------
// Order-preserving float <-> uint map (so one integer REDUX gives a float min or max of either sign).
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

// Lower bound of cov_q2(x, y, z, v3) over every point of the box [lo, hi].  Rounding is monotone, so with the same
// operation sequence as cov_q2 (fsub, then fmul/fma/fma) the bound holds for the COMPUTED q2 of each point, exactly:
// |fl(x - td)| >= max(fl(lo - td), fl(td - hi), 0) for lo <= x <= hi.  An empty box (lo = +inf) gives +inf.
__device__ __forceinline__ float box_q2lb(const float4& lo, const float4& hi, const float4& v3) {
    const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, v3.x), __fsub_rn(v3.x, hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, v3.y), __fsub_rn(v3.y, hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, v3.z), __fsub_rn(v3.z, hi.z)), 0.f);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
__device__ __forceinline__ void box_union(float4& lo, float4& hi, const float4& lo2, const float4& hi2) {
    lo.x = fminf(lo.x, lo2.x); lo.y = fminf(lo.y, lo2.y); lo.z = fminf(lo.z, lo2.z);
    hi.x = fmaxf(hi.x, hi2.x); hi.y = fmaxf(hi.y, hi2.y); hi.z = fmaxf(hi.z, hi2.z);
}
// Pass-A cap on q2 from the extrema seen so far (as uints): +inf (evaluate everything) unless the minimum is known
// to be exactly 0.  NaN propagates and every test against it evaluates.
__device__ __forceinline__ float minmax_qcap(unsigned mn, unsigned mx, float inv_kd) {
    float cap = __uint_as_float(0x7f800000u);
    if (mn == 0u && mx != 0u) cap = (1e-4f - __log2f(__uint_as_float(mx))) * inv_kd * 1.000001f;
    return cap;
}

// ---- TMA bulk copies (global -> shared) completing on an mbarrier ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_copy(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// =============================================== pass A, dense ===============================================
// Every (point, pose) pair.  The evaluation runs on packed fp32 pairs (FFMA2): a thread evaluates each of its points
// against TWO poses at a time, the pose rows interleaved pairwise in shared memory (COV_PAIR_F4) so that one LDS.128
// brings two packed constants; the points are kept as broadcast pairs in registers.
template <int PPT, int MINB>
__global__ void __launch_bounds__(COV_THREADS, MINB)
cov_traj_minmax_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                       const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                       unsigned* __restrict__ gmin, unsigned* __restrict__ gmax) {
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    const int NP = (W + 1) >> 1;  // pose pairs; an odd W repeats its last pose in the spare slot
    unsigned* smin = reinterpret_cast<unsigned*>(ptab + (size_t)NP * COV_PAIR_F4);
    unsigned* smax = smin + 2 * NP;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int w = tid; w < 2 * NP; w += COV_THREADS) {
        const int ws = w < W ? w : W - 1;
        float4 row[COV_ROW_F4];
        cov_pose_row(poses + 3 * ws, quats + 4 * ws, K9, C, row);
        cov_pair_store(ptab + (size_t)(w >> 1) * COV_PAIR_F4, w & 1, row);
        smin[w] = 0x7f800000u;
        smax[w] = 0u;
    }
    __syncthreads();
    const CovConst2 C2 = cov_make_const2(C);
    constexpr int T = tile_points(PPT);
    const int64_t ntiles = (n + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        f2_t X[PPT], Y[PPT], Z[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            int64_t j = tile * T + s * COV_THREADS + tid;
            j = j < n ? j : n - 1;  // a duplicate cannot change a min or a max
            const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
            X[s] = f2_pack(x, x);
            Y[s] = f2_pack(y, y);
            Z[s] = f2_pack(z, z);
        }
        for (int p0 = 0; p0 < NP; p0 += 16) {
            // lane 2i (2i+1) keeps the warp-wide min/max of pose A (B) of pair p0+i; one shared atomic per 32 poses
            unsigned keep_mn = 0x7f800000u, keep_mx = 0u;
            const int pn = (NP - p0 < 16) ? (NP - p0) : 16;
            for (int i = 0; i < pn; ++i) {
                const float4* q = ptab + (size_t)(p0 + i) * COV_PAIR_F4;
                const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5], q6 = q[6], q7 = q[7];
                float mnA, mxA, mnB, mxB;
                f2_unpack(cov_vis2p(X[0], Y[0], Z[0], q0, q1, q2, q3, q4, q5, q6, q7, C2), mnA, mnB);
                mxA = mnA;
                mxB = mnB;
#pragma unroll
                for (int s = 1; s < PPT; ++s) {
                    float a, b;
                    f2_unpack(cov_vis2p(X[s], Y[s], Z[s], q0, q1, q2, q3, q4, q5, q6, q7, C2), a, b);
                    mnA = fminf(mnA, a); mxA = fmaxf(mxA, a);
                    mnB = fminf(mnB, b); mxB = fmaxf(mxB, b);
                }
                const unsigned umnA = __reduce_min_sync(kFull, __float_as_uint(mnA));
                const unsigned umxA = __reduce_max_sync(kFull, __float_as_uint(mxA));
                const unsigned umnB = __reduce_min_sync(kFull, __float_as_uint(mnB));
                const unsigned umxB = __reduce_max_sync(kFull, __float_as_uint(mxB));
                if (lane == 2 * i) {
                    keep_mn = umnA;
                    keep_mx = umxA;
                } else if (lane == 2 * i + 1) {
                    keep_mn = umnB;
                    keep_mx = umxB;
                }
            }
            if (lane < 2 * pn) {
                atomicMin(smin + 2 * p0 + lane, keep_mn);
                atomicMax(smax + 2 * p0 + lane, keep_mx);
            }
        }
    }
    __syncthreads();
    for (int w = tid; w < W; w += COV_THREADS) {
        atomicMin(gmin + w, smin[w]);
        atomicMax(gmax + w, smax[w]);
    }
}

__global__ void cov_minmax_init_kernel(unsigned* gmin, unsigned* gmax, int W) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < W) {
        gmin[w] = 0x7f800000u;
        gmax[w] = 0u;
    }
}

// ============================================ pruned passes: set-up ============================================
// Control words of one pruned call, at the head of the workspace (zeroed once per call: by a memset node in pass A,
// by cov_traj_table_kernel in pass B):
//   ctrl[0] work-list length   ctrl[1] some pose has min_j m > 0   ctrl[2] block ticket of the evaluation kernel
//   ctrl[4..5] (tile, pose) pairs the cull listed (u64)
//   enc[0..W)  ~bits(min_j m)   enc[W..2W) bits(max_j m)   pass A's extrema in an encoding whose neutral element is 0:
//              m >= +0, so the unsigned order of the bits is the float order; both are kept with atomicMax
//   desc[c]    look-back descriptor of cull chunk c (status in the top two bits)
constexpr int kCtrlInts = 16;
constexpr unsigned long long kDescAgg = 1ull << 62, kDescPrefix = 2ull << 62, kDescValue = (1ull << 62) - 1;
constexpr int kChunkTiles = 64;  // tiles per cull block iteration (8 warps x 8 tiles) = one look-back descriptor

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Pass A, first launch: pose table + seed.  Grid = (sample blocks) x (chunks of 32 poses).  The first 32 threads of a
// block build the chunk's pose rows in shared memory (fp64 inside; blocks of the first grid column also store them to
// the global table the later kernels copy), then every thread evaluates ONE point of a strided sample of the cloud
// against the 32 poses (dense), so that the cull starts from a good lower bound of each maximum and knows which
// minima are exactly 0.  Per pose one REDUX pair per warp, the 8 warps meet in shared memory, 64 global atomics per
// block into the zero-initialised encoded extrema.
__global__ void __launch_bounds__(COV_THREADS)
cov_traj_prepare_kernel(const float* __restrict__ xyz, int64_t nsamples, int64_t stride, const float* __restrict__ poses,
                        const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                        float4* __restrict__ table, unsigned* __restrict__ enc) {
    __shared__ float4 rows[32 * COV_ROW_F4];
    __shared__ unsigned smn[32], smx[32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int w0 = blockIdx.y * 32;
    const int wn = (W - w0 < 32) ? (W - w0) : 32;
    if (tid < wn) {
        cov_pose_row(poses + 3 * (w0 + tid), quats + 4 * (w0 + tid), K9, C, rows + tid * COV_ROW_F4);
        if (blockIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < COV_ROW_F4; ++i) table[(size_t)(w0 + tid) * COV_ROW_F4 + i] = rows[tid * COV_ROW_F4 + i];
        }
        smn[tid] = 0x7f800000u;
        smx[tid] = 0u;
    }
    __syncthreads();
    int64_t j = (int64_t)blockIdx.x * COV_THREADS + tid;
    j = (j < nsamples ? j : nsamples - 1) * stride;  // a duplicate cannot change a min or a max
    const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
    unsigned keep_mn = 0x7f800000u, keep_mx = 0u;
    for (int i = 0; i < wn; ++i) {
        const float4* r = rows + i * COV_ROW_F4;
        const float m = cov_vis<false>(x, y, z, r[0], r[1], r[2], r[3], C, nullptr);
        const unsigned umn = __reduce_min_sync(kFull, __float_as_uint(m));
        const unsigned umx = __reduce_max_sync(kFull, __float_as_uint(m));
        if (lane == i) {
            keep_mn = umn;
            keep_mx = umx;
        }
    }
    if (lane < wn) {
        atomicMin(smn + lane, keep_mn);
        atomicMax(smx + lane, keep_mx);
    }
    __syncthreads();
    if (tid < wn) {
        atomicMax(enc + w0 + tid, ~smn[tid]);
        atomicMax(enc + W + w0 + tid, smx[tid]);
    }
}

// Pass B, first launch: block 0 zeroes the control words, the look-back descriptors and the caller's accumulator rows
// (acc[W * STRIDE] starts at `sum_base` = 0.5 * n: what the unlisted points add to the reward sum) and sets ctrl[1] when
// some pose has min_j m > 0; every block builds 64 rows of the pose table with the normalisation constants:
// v3.w = qthr, v4 = (b/2, b, 1/b, a), v5.w = thr.
__global__ void __launch_bounds__(64)
cov_traj_table_kernel(const float* __restrict__ poses, const float* __restrict__ quats, int W,
                      const float* __restrict__ K9, CovConst C, const float* __restrict__ minmax,
                      float4* __restrict__ table, int* __restrict__ ctrl, int ctrl_words, double* __restrict__ acc,
                      double sum_base) {
    const int tid = threadIdx.x;
    if (blockIdx.x == 0) {
        int amin = 0;
        for (int w = tid; w < W; w += blockDim.x) amin |= minmax[w] > 0.f;
        amin = __syncthreads_or(amin);
        for (int i = tid; i < ctrl_words; i += blockDim.x) ctrl[i] = (i == 1) ? amin : 0;
        if (acc) {
            for (int i = tid; i < W * COV_ACC_STRIDE; i += blockDim.x) acc[i] = 0.0;
            if (tid == 0) acc[(size_t)W * COV_ACC_STRIDE] = sum_base;
        }
    }
    const int w = blockIdx.x * blockDim.x + tid;
    if (w >= W) return;
    float4 row[COV_ROW_F4];
    cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, row);
    const float a = minmax[w];
    const float b = __fsub_rn(minmax[W + w], a);
    const float hb = 0.5f * b;
    const float thr = (a + hb) * (1.f - 9.5367431640625e-07f);  // conservative gate threshold
    row[5].w = thr;
    // q2 above qthr cannot reach thr (thr <= 0 or NaN, or a > 0 — arg-min points carry gradient: never prune)
    row[3].w = (thr > 0.f && !(a > 0.f)) ? (float)((1e-4 - log2((double)thr)) / (double)C.kd * 1.000001)
                                         : __uint_as_float(0x7f800000u);
    row[4] = make_float4(hb, b, __frcp_rn(b), a);
#pragma unroll
    for (int i = 0; i < COV_ROW_F4; ++i) table[(size_t)w * COV_ROW_F4 + i] = row[i];
}

// Bounding boxes of runs of kBoxPts consecutive points: boxes[2b] = (lo, 0), boxes[2b+1] = (hi, 0); boxes past the
// end of the cloud are empty (+inf, -inf).  One warp per box.
__global__ void __launch_bounds__(256) cov_tile_boxes_kernel(const float* __restrict__ xyz, int64_t n, int64_t nboxes,
                                                             float4* __restrict__ boxes) {
    const float inf = __uint_as_float(0x7f800000u);
    const int lane = threadIdx.x & 31;
    for (int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < nboxes; b += (int64_t)gridDim.x * 8) {
        float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
#pragma unroll
        for (int s = 0; s < kBoxPts / 32; ++s) {
            const int64_t j = b * kBoxPts + s * 32 + lane;
            if (j < n) {
                const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
                lx = fminf(lx, x); ly = fminf(ly, y); lz = fminf(lz, z);
                hx = fmaxf(hx, x); hy = fmaxf(hy, y); hz = fmaxf(hz, z);
            }
        }
        const unsigned ulx = __reduce_min_sync(kFull, f2ord(lx)), uly = __reduce_min_sync(kFull, f2ord(ly));
        const unsigned ulz = __reduce_min_sync(kFull, f2ord(lz)), uhx = __reduce_max_sync(kFull, f2ord(hx));
        const unsigned uhy = __reduce_max_sync(kFull, f2ord(hy)), uhz = __reduce_max_sync(kFull, f2ord(hz));
        if (lane == 0) {
            boxes[2 * b] = make_float4(ord2f(ulx), ord2f(uly), ord2f(ulz), 0.f);
            boxes[2 * b + 1] = make_float4(ord2f(uhx), ord2f(uhy), ord2f(uhz), 0.f);
        }
    }
}

// Cull + work list.  mask[tile][c] bit i <=> pose 32c+i can matter for some point of the tile.  One warp per GROUP of 8
// consecutive tiles: every pose is tested against the group's box first (one lane per pose), and only the few that
// pass are tested against the 8 tile boxes (one lane per tile).  Pass A passes `enc` (after the prepare launch) and the
// cap is derived here; pass B reads qthr from the table.
// A block iteration covers a CHUNK of 64 consecutive tiles; the block first culls ALL its chunks (warps independent, no
// barrier), then appends each chunk's listed tiles to the work list in ascending order: the chunk's offset is the
// exclusive prefix sum of the chunks' counts, obtained with a decoupled look-back over per-chunk descriptors (publish
// the own aggregates, then sum predecessors until one carries an inclusive prefix).  The list is a pure function of the
// masks (deterministic).  Grid <= resident capacity, and every block publishes its aggregates before it waits for
// anybody's, so every predecessor a block waits for has been (or is being) published.  `fill_dst` (pass B): the same
// blocks write 1/2 to every reward.
constexpr int kCullMaxIter = 32;  // chunks per block at most (the host sizes the grid accordingly)

__global__ void __launch_bounds__(256)
cov_cull_kernel(const float4* __restrict__ boxes, int boxes_per_tile, int64_t ntiles, const float4* __restrict__ table,
                int W, const unsigned* __restrict__ enc, float inv_kd, unsigned* __restrict__ amask_g, int mask_stride,
                int* __restrict__ worklist, int* __restrict__ ctrl, unsigned long long* __restrict__ desc,
                float* __restrict__ fill_dst, int64_t fill_n) {
    extern __shared__ float4 v3s[];
    __shared__ unsigned char listed_s[kCullMaxIter][8];  // per own chunk and warp: which of the warp's 8 tiles are listed
    __shared__ int chunk_base[kCullMaxIter];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = tid; w < W; w += blockDim.x) {
        float4 v3 = table[(size_t)w * COV_ROW_F4 + 3];
        if (enc) v3.w = minmax_qcap(~enc[w], enc[W + w], inv_kd);
        v3s[w] = v3;
    }
    if (fill_dst) {  // fire-and-forget stores: they drain while the blocks cull
        const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
        for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + tid) * 4; i < fill_n; i += stride) {
            if (i + 4 <= fill_n && ((reinterpret_cast<uintptr_t>(fill_dst + i) & 15) == 0)) {
                __stcs(reinterpret_cast<float4*>(fill_dst + i), make_float4(0.5f, 0.5f, 0.5f, 0.5f));
            } else {
                for (int64_t k = i; k < fill_n && k < i + 4; ++k) fill_dst[k] = 0.5f;
            }
        }
    }
    __syncthreads();
    const float inf = __uint_as_float(0x7f800000u);
    const int nwords = (W + 31) >> 5;
    const int64_t nchunks = (ntiles + kChunkTiles - 1) / kChunkTiles;
    const int sub = lane >> 3, tl = lane & 7;  // box loads: lane = (pass-local box slot, tile of the group)
    unsigned long long npairs = 0;
    int it = 0;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
        // lane -> tile tl of this warp's group; the 4 lanes with the same tl share the tile's boxes_per_tile boxes (<= 16)
        const int64_t tile_l = (chunk * 8 + warp) * 8 + tl;
        float4 lo = make_float4(inf, inf, inf, 0.f), hi = make_float4(-inf, -inf, -inf, 0.f);
        if (tile_l < ntiles) {
            for (int b = sub; b < boxes_per_tile; b += 4) {
                const float4 l2 = boxes[(tile_l * boxes_per_tile + b) * 2], h2 = boxes[(tile_l * boxes_per_tile + b) * 2 + 1];
                box_union(lo, hi, l2, h2);
            }
        }
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {  // join the 4 lanes of a tile: afterwards every lane holds ITS tile's box
            lo.x = fminf(lo.x, __shfl_xor_sync(kFull, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(kFull, lo.y, o));
            lo.z = fminf(lo.z, __shfl_xor_sync(kFull, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(kFull, hi.x, o));
            hi.y = fmaxf(hi.y, __shfl_xor_sync(kFull, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(kFull, hi.z, o));
        }
        float4 glo = lo, ghi = hi;  // the group's box: join the 8 tiles
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            glo.x = fminf(glo.x, __shfl_xor_sync(kFull, glo.x, o)); glo.y = fminf(glo.y, __shfl_xor_sync(kFull, glo.y, o));
            glo.z = fminf(glo.z, __shfl_xor_sync(kFull, glo.z, o)); ghi.x = fmaxf(ghi.x, __shfl_xor_sync(kFull, ghi.x, o));
            ghi.y = fmaxf(ghi.y, __shfl_xor_sync(kFull, ghi.y, o)); ghi.z = fmaxf(ghi.z, __shfl_xor_sync(kFull, ghi.z, o));
        }
        unsigned any = 0u;  // lanes 0..7: number of poses listed for tile tl
        for (int c = 0; c < mask_stride; ++c) {
            const int w = c * 32 + lane;
            bool cand = false;
            if (c < nwords && w < W) {
                const float4 v3 = v3s[w];
                cand = !(box_q2lb(glo, ghi, v3) > v3.w);  // NaN cap: evaluate
            }
            unsigned gword = __ballot_sync(kFull, cand);
            unsigned mine = 0u;  // lanes 0..7: word c of tile tl's mask
            while (gword) {
                const int b = __ffs(gword) - 1;
                gword &= gword - 1;
                const float4 v3 = v3s[c * 32 + b];
                const bool hit = !(box_q2lb(lo, hi, v3) > v3.w);  // an empty box (+inf, -inf) gives +inf: never hit
                const unsigned tb = __ballot_sync(kFull, hit) & 0xffu;  // lanes 0..7 speak for the 8 tiles
                if ((tb >> tl) & 1u) mine |= 1u << b;
            }
            if (lane < 8 && tile_l < ntiles) {
                amask_g[tile_l * mask_stride + c] = mine;
                any += __popc(mine);
            }
        }
        const bool listed = lane < 8 && tile_l < ntiles && any != 0u;
        if (listed) npairs += any;
        const unsigned listed8 = __ballot_sync(kFull, listed);
        if (lane == 0) listed_s[it][warp] = (unsigned char)listed8;
    }
    const int n_it = it;
    __syncthreads();
    if (warp == 0) {
        // own chunk `k` is chunk blockIdx.x + k * gridDim.x; lane l < 8 holds warp l's listed bits
        for (int k = lane; k < n_it; k += 32) {  // publish every own aggregate first
            int total = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) total += __popc((unsigned)listed_s[k][q]);
            const int64_t chunk = blockIdx.x + (int64_t)k * gridDim.x;
            if (chunk > 0) st_release_u64(desc + chunk, kDescAgg | (unsigned long long)total);
        }
        __syncwarp();
        for (int k = 0; k < n_it; ++k) {
            const int64_t chunk = blockIdx.x + (int64_t)k * gridDim.x;
            unsigned long long total = 0ull;
#pragma unroll
            for (int q = 0; q < 8; ++q) total += (unsigned long long)__popc((unsigned)listed_s[k][q]);
            unsigned long long excl = 0ull;
            if (chunk > 0) {
                int64_t look = chunk - 1;
                while (true) {  // windows of 32 predecessors, nearest first
                    const int64_t idx = look - lane;
                    unsigned long long d = kDescPrefix;  // before chunk 0: an inclusive prefix of 0
                    if (idx >= 0) d = ld_acquire_u64(desc + idx);
                    const unsigned empty = __ballot_sync(kFull, (d >> 62) == 0ull);
                    const unsigned pref = __ballot_sync(kFull, (d >> 62) == 2ull);
                    const int first_pref = pref ? (__ffs(pref) - 1) : 32;
                    const unsigned need = first_pref >= 31 ? kFull : ((2u << first_pref) - 1u);
                    if (empty & need) continue;  // a predecessor in the window has not published yet: look again
                    unsigned long long v = (lane <= first_pref) ? (d & kDescValue) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
                    excl += v;
                    if (first_pref < 32) break;
                    look -= 32;
                }
            }
            if (lane == 0) {
                st_release_u64(desc + chunk, kDescPrefix | (excl + total));
                if (chunk == nchunks - 1) ctrl[0] = (int)(excl + total);
                chunk_base[k] = (int)excl;
            }
        }
    }
    __syncthreads();
    for (int k = 0; k < n_it; ++k) {
        const int64_t chunk = blockIdx.x + (int64_t)k * gridDim.x;
        const int64_t tile_l = (chunk * 8 + warp) * 8 + tl;
        const unsigned listed8 = listed_s[k][warp];
        int before = 0;  // listed tiles of the chunk in the warps ahead of this one
        for (int q = 0; q < warp; ++q) before += __popc((unsigned)listed_s[k][q]);
        if (lane < 8 && ((listed8 >> lane) & 1u))
            worklist[chunk_base[k] + before + __popc(listed8 & ((1u << lane) - 1u))] = (int)tile_l;
    }
    npairs = __reduce_add_sync(kFull, (unsigned)npairs);
    if (lane == 0 && npairs) atomicAdd(reinterpret_cast<unsigned long long*>(ctrl + 4), npairs);
}

// =============================================== pass A, pruned ===============================================
// Persistent WARPS over the work list, like pass B below: the unit of work is one box of 128 consecutive points of a
// listed tile, drawn by a warp from a global ticket counter; lane 0 stages the next item (points, box, the tile's pose
// mask) by TMA into the warp's other private stage while the warp works on the current one.  A listed pose is evaluated
// when the cap ball (from the seed's extrema; any later value is tighter and valid) meets the warp's box and one of its
// points.  Extrema meet per block in shared memory, then in the zero-initialised encoded arrays; the last block decodes
// them into the caller's floats.
constexpr int kItemPts = 128;                 // points per work item = one warp's registers = one precomputed box
constexpr int kItemsPerTile = 8;              // a listed tile (1024 points) is eight items
constexpr int kItemPpt = kItemPts / 32;
constexpr int kItemStageWordsA = kItemPts * 3 + 8 + kMaskWords;               // points | box | pose mask
constexpr int kItemStageWords = kItemStageWordsA + kItemPts;                  // ... | permutation (pass B)

__global__ void __launch_bounds__(COV_THREADS, 3)
cov_traj_minmax_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float4* __restrict__ table, int W, CovConst C,
                             unsigned* __restrict__ enc, float* __restrict__ minmax_out, const float4* __restrict__ boxes,
                             const unsigned* __restrict__ amask_g, int mask_stride, const int* __restrict__ worklist,
                             int* __restrict__ ctrl, int64_t ntiles, unsigned long long* __restrict__ stats,
                             float* __restrict__ fill_dst, int64_t fill_n) {
    constexpr int PPT = kItemPpt;
    extern __shared__ float4 smem4[];                      // block minima | maxima (uint) | per-pose caps
    unsigned* smin = reinterpret_cast<unsigned*>(smem4);
    unsigned* smax = smin + W;
    float* sqcap = reinterpret_cast<float*>(smax + W);
    __shared__ __align__(128) float stages[kWarps][2][kItemStageWordsA];
    __shared__ __align__(8) unsigned long long mbar[kWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv_kd = 1.f / C.kd;
    unsigned long long* bar = mbar[warp];
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    for (int w = tid; w < W; w += COV_THREADS) {
        smin[w] = 0x7f800000u;
        smax[w] = 0u;
        sqcap[w] = minmax_qcap(~enc[w], enc[W + w], inv_kd);  // from the seed (any later value is tighter and valid)
    }
    __syncthreads();
    const int n_items = ctrl[0] * kItemsPerTile;
    const int nwords = (W + 31) >> 5;
    int* ticket = ctrl + 3;
    // Pre-fill of pass B's rewards with 1/2 under this kernel's idle memory bandwidth (cov_traj_opts.prefill_dev): the
    // warp that takes item i also writes slice i of the buffer (fire-and-forget streaming stores), in 16-byte units.
    const int64_t fill_q = (fill_n + 3) / 4;                                          // float4 slots (tail handled below)
    const int64_t fill_per_item = fill_dst ? (fill_q + (n_items > 0 ? n_items : 1) - 1) / (n_items > 0 ? n_items : 1) : 0;
    auto fill_slice = [&](int64_t q0, int64_t q1) {
        if (q1 > fill_q) q1 = fill_q;
        for (int64_t q = q0 + lane; q < q1; q += 32) {
            if (q * 4 + 4 <= fill_n) __stcs(reinterpret_cast<float4*>(fill_dst) + q, make_float4(0.5f, 0.5f, 0.5f, 0.5f));
            else for (int64_t k = q * 4; k < fill_n; ++k) fill_dst[k] = 0.5f;
        }
    };
    if (fill_dst && n_items == 0) {  // nothing listed: the warps of the grid share the whole buffer
        const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
        const int64_t per = (fill_q + nw - 1) / nw;
        fill_slice(gw * per, (gw + 1) * per);
    }
    auto issue = [&](int it, int b) {
        if (it >= n_items) return;
        const int64_t tile = worklist[it >> 3];
        const int64_t j0 = tile * (kItemsPerTile * kItemPts) + (int64_t)(it & 7) * kItemPts;
        float* st = stages[warp][b];
        const bool whole = j0 + kItemPts <= n;
        mbar_expect_tx(&bar[b], (whole ? kItemPts * 12u : 0u) + 32u + (unsigned)mask_stride * 4u);
        if (whole) tma_copy(st, xyz + j0 * 3, kItemPts * 12u, &bar[b]);
        tma_copy(st + kItemPts * 3, boxes + (j0 / kBoxPts) * 2, 32u, &bar[b]);
        tma_copy(st + kItemPts * 3 + 8, amask_g + tile * mask_stride, (unsigned)mask_stride * 4u, &bar[b]);
    };
    unsigned n_box = 0, n_pre = 0, n_full = 0;
    unsigned uses0 = 0, uses1 = 0;
    int cur = 0;
    if (lane == 0) {
        cur = atomicAdd(ticket, 1);
        issue(cur, 0);
    }
    cur = __shfl_sync(kFull, cur, 0);
    int buf = 0;
    while (cur < n_items) {
        int nxt = 0;
        if (lane == 0) {  // the other stage was last read before the __syncwarp that ended the previous item
            nxt = atomicAdd(ticket, 1);
            issue(nxt, buf ^ 1);
        }
        nxt = __shfl_sync(kFull, nxt, 0);
        if (fill_dst) fill_slice((int64_t)cur * fill_per_item, (int64_t)(cur + 1) * fill_per_item);
        if (buf == 0) mbar_wait(&bar[0], uses0++ & 1u);
        else mbar_wait(&bar[1], uses1++ & 1u);
        const float* st = stages[warp][buf];
        const int64_t tile = worklist[cur >> 3];
        const int64_t j0 = tile * (kItemsPerTile * kItemPts) + (int64_t)(cur & 7) * kItemPts;
        float px[PPT], py[PPT], pz[PPT];
        if (j0 + kItemPts <= n) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                px[s] = st[(s * 32 + lane) * 3];
                py[s] = st[(s * 32 + lane) * 3 + 1];
                pz[s] = st[(s * 32 + lane) * 3 + 2];
            }
        } else {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                int64_t j = j0 + s * 32 + lane;
                j = j < n ? j : n - 1;  // a duplicate cannot change a min or a max (j0 < n: the item holds a real point)
                px[s] = __ldg(xyz + j * 3);
                py[s] = __ldg(xyz + j * 3 + 1);
                pz[s] = __ldg(xyz + j * 3 + 2);
            }
        }
        const float4* tb = reinterpret_cast<const float4*>(st + kItemPts * 3);
        const float4 wlo = tb[0], whi = tb[1];
        const unsigned* am = reinterpret_cast<const unsigned*>(st + kItemPts * 3 + 8);
        if (j0 < n) {
            for (int c = 0; c < nwords; ++c) {
                unsigned word = am[c];
                while (word) {
                    const int w = c * 32 + __ffs(word) - 1;
                    word &= word - 1;
                    const float4* row = table + (size_t)w * COV_ROW_F4;
                    const float4 v3 = __ldg(row + 3);
                    const float cap = sqcap[w];
                    ++n_box;
                    if (box_q2lb(wlo, whi, v3) > cap) continue;
                    ++n_pre;
                    float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
                    for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
                    if (!__any_sync(kFull, !(qmin > cap))) continue;
                    ++n_full;
                    const float4 v0 = __ldg(row), v1 = __ldg(row + 1), v2 = __ldg(row + 2);
                    float m[PPT];
#pragma unroll
                    for (int s = 0; s < PPT; ++s) m[s] = cov_vis<false>(px[s], py[s], pz[s], v0, v1, v2, v3, C, nullptr);
                    float mn = m[0], mx = m[0];
#pragma unroll
                    for (int s = 1; s < PPT; ++s) {
                        mn = fminf(mn, m[s]);
                        mx = fmaxf(mx, m[s]);
                    }
                    const unsigned umn = __reduce_min_sync(kFull, __float_as_uint(mn));
                    const unsigned umx = __reduce_max_sync(kFull, __float_as_uint(mx));
                    if (lane == 0) {
                        atomicMin(smin + w, umn);
                        atomicMax(smax + w, umx);
                    }
                }
            }
        }
        __syncwarp();  // every lane is done with this stage before lane 0 refills it
        cur = nxt;
        buf ^= 1;
    }
    __syncthreads();
    for (int w = tid; w < W; w += COV_THREADS) {
        if (smax[w] != 0u || smin[w] != 0x7f800000u) {
            atomicMax(enc + w, ~smin[w]);
            atomicMax(enc + W + w, smax[w]);
        }
    }
    if (stats && lane == 0) {
        if (tid == 0 && blockIdx.x == 0) atomicAdd(stats + 2, (unsigned long long)ntiles * kWarps * W);
        atomicAdd(stats + 3, (unsigned long long)n_full);
        atomicAdd(stats + 5, (unsigned long long)n_pre);
        atomicAdd(stats + 7, (unsigned long long)n_box);
    }
    // the last block to arrive decodes the extrema into the caller's floats (minima, then maxima)
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ctrl + 2, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        unsigned* out = reinterpret_cast<unsigned*>(minmax_out);
        for (int w = tid; w < W; w += COV_THREADS) {
            out[w] = ~__ldcg(enc + w);
            out[W + w] = __ldcg(enc + W + w);
        }
    }
}

// =================================================== pass B ===================================================
// Gate bit matrix of the dense kernel.  Row w holds kWarps groups of PPT ballot words (group g = the words of warp g;
// word k of a row covers points [32k, 32k+32) of the tile, since a warp owns 32*PPT consecutive points).  Rows are
// padded by 4 words (conflict-free when lanes walk different rows at the same word).
template <int PPT>
__device__ __forceinline__ unsigned* bit_row_group(unsigned* bits, int w, int group) {
    return bits + (size_t)w * bit_stride(PPT) + group * PPT;
}

__device__ __forceinline__ float4 lds_f4(unsigned saddr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// Phase 1 of the dense pass B for one pose PAIR: every point of the thread against both poses on packed fp32 pairs
// (cov_vis2p), then ONE conservative warp vote for the pair and — only when some lane may pass for either pose — the
// exact gate, the ballots and the log-odds per pose.  Non-zero ballot words go to the gate bit matrix (which is all zero
// otherwise: zeroed at kernel start, and the gradient walk clears what it consumed).  The spare slot of an odd W repeats
// the last pose and is skipped after the vote.  Returns whether this warp gated a pair (warp-uniform).
template <int PPT, bool AMIN>
__device__ __forceinline__ bool dense_pair_iter(int pair, int W, unsigned ptab_s, unsigned* __restrict__ bits, int warp,
                                                const f2_t (&X)[PPT], const f2_t (&Y)[PPT], const f2_t (&Z)[PPT],
                                                float (&L)[PPT], const CovConst& C, const CovConst2& C2,
                                                double* __restrict__ acc, int lane) {
    const unsigned base = ptab_s + (unsigned)pair * (COV_PAIR_F4 * 16u);
    const float4 q0 = lds_f4(base), q1 = lds_f4(base + 16u), q2 = lds_f4(base + 32u), q3 = lds_f4(base + 48u),
                 q4 = lds_f4(base + 64u), q5 = lds_f4(base + 80u), q6 = lds_f4(base + 96u), q7 = lds_f4(base + 112u);
    float m[2][PPT];
#pragma unroll
    for (int s = 0; s < PPT; ++s) f2_unpack(cov_vis2p(X[s], Y[s], Z[s], q0, q1, q2, q3, q4, q5, q6, q7, C2), m[0][s], m[1][s]);
    float mxa = m[0][0], mxb = m[1][0];
#pragma unroll
    for (int s = 1; s < PPT; ++s) {
        mxa = fmaxf(mxa, m[0][s]);
        mxb = fmaxf(mxb, m[1][s]);
    }
    // conservative gate thresholds (v3.w of the two poses): one vote for the pair; a few % of the iterations pass
    bool hot = __any_sync(kFull, (mxa >= q7.z) | (mxb >= q7.w));
    if (AMIN) hot = true;  // arg-min points carry gradient: every pair is looked at (block-uniform)
    if (!hot) return false;
    unsigned anyb = 0u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int w = 2 * pair + h;
        if (w >= W) break;  // the spare slot of an odd pose count
        const float4 v4 = lds_f4(base + (8u + h) * 16u);
        unsigned bal[PPT];
        unsigned any_h = 0u;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float d = __fsub_rn(m[h][s], v4.w);
            const bool act = d >= v4.x;  // exactly p >= 0.5
            bal[s] = __ballot_sync(kFull, act);
            any_h |= bal[s];
            if (act) {
                const float p = __fmul_rn(d, v4.z);
                const float qc = (p > C.hi) ? C.hi : p;  // upper clip; a NaN p (pose that sees nothing: 0/0) stays NaN, as torch.clip
                L[s] += COV_LN2_F * cov_lg2(qc * cov_rcp(1.f - qc));
                if (d == v4.y) {
                    float x, y, z, t;
                    f2_unpack(X[s], x, t); f2_unpack(Y[s], y, t); f2_unpack(Z[s], z, t);
                    tie_accumulate_pair(x, y, z, w, C, acc, w * COV_ACC_STRIDE + 8);
                }
            }
        }
        anyb |= any_h;
        if (any_h != 0u && lane == 0) {
            unsigned* brow = bit_row_group<PPT>(bits, w, warp);
            if (PPT == 4) *reinterpret_cast<uint4*>(brow) = make_uint4(bal[0], bal[1 % PPT], bal[2 % PPT], bal[3 % PPT]);
            else if (PPT == 2) *reinterpret_cast<uint2*>(brow) = make_uint2(bal[0], bal[1 % PPT]);
            else brow[0] = bal[0];
        }
        if (AMIN) {  // only compact clouds whose minimum did not underflow to 0
            if (v4.w > 0.f) {
#pragma unroll
                for (int s = 0; s < PPT; ++s)
                    if (m[h][s] == v4.w) {
                        float x, y, z, t;
                        f2_unpack(X[s], x, t); f2_unpack(Y[s], y, t); f2_unpack(Z[s], z, t);
                        tie_accumulate_pair(x, y, z, w, C, acc, w * COV_ACC_STRIDE + 15);
                    }
            }
        }
    }
    return anyb != 0u;
}

// Forward of one listed pose for a warp of the pruned kernel: m for the warp's points, log-odds of the gated ones added to
// L; returns whether the warp has a gated pair for the pose.  `row` = the pose's 6 float4 in the global table (L1 hits:
// a tile touches a handful of rows).  The conservative threshold lives in v5.w (v3.w holds qthr); the exact gate test
// runs only when some lane may pass.
template <int PPT, bool AMIN>
__device__ __forceinline__ bool tiles_pose_iter(const float4* __restrict__ row, const float4& v3, const float (&px)[PPT],
                                                const float (&py)[PPT], const float (&pz)[PPT], float (&L)[PPT],
                                                const CovConst& C, double* __restrict__ acc_row) {
    float m[PPT];
    const float4 v0 = __ldg(row), v1 = __ldg(row + 1), v2 = __ldg(row + 2);
    const float thr = __ldg(&row[5].w);
#pragma unroll
    for (int s = 0; s < PPT; ++s) m[s] = cov_vis<false>(px[s], py[s], pz[s], v0, v1, v2, v3, C, nullptr);
    float mmax = m[0];
#pragma unroll
    for (int s = 1; s < PPT; ++s) mmax = fmaxf(mmax, m[s]);
    bool any = false;
    if (__any_sync(kFull, mmax >= thr)) {  // warp-uniform: some point may pass the gate (conservative threshold)
        const float4 v4 = __ldg(row + 4);
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float d = __fsub_rn(m[s], v4.w);
            const bool act = d >= v4.x;  // exactly p >= 0.5
            any |= act;
            if (act) {
                const float p = __fmul_rn(d, v4.z);
                const float qc = (p > C.hi) ? C.hi : p;  // upper clip; a NaN p (pose that sees nothing: 0/0) stays NaN, as torch.clip
                L[s] += COV_LN2_F * cov_lg2(qc * cov_rcp(1.f - qc));
                if (acc_row && d == v4.y) tie_accumulate_g(px[s], py[s], pz[s], row, C, acc_row + 8);
            }
        }
        any = __any_sync(kFull, any);
    }
    if (AMIN) {  // only compact clouds whose minimum did not underflow to 0 (block-uniform choice of the caller)
        const float a = __ldg(&row[4].w);
        if (a > 0.f) {
#pragma unroll
            for (int s = 0; s < PPT; ++s)
                if (m[s] == a) tie_accumulate_g(px[s], py[s], pz[s], row, C, acc_row + 15);
        }
    }
    return any;
}

// Phase 2 of the dense kernel: the gated pairs of all W bit rows, each row split over 2^seg_log2 lanes.  A lane pops
// its set bits in ascending point order, recomputes m (bit-identical) and dm/dx, and accumulates in registers;
// segments are combined with xor-shuffles; one owner lane adds into accs.
template <int PPT>
__device__ __forceinline__ void fused_phase2(const float4* __restrict__ ptab, unsigned* __restrict__ bits,
                                             const float* __restrict__ pt, const float* __restrict__ Gs,
                                             float* __restrict__ accs, int W, int seg_log2, const CovConst& C, int tid) {
    constexpr int NW = bit_words(PPT);
    constexpr int RS = bit_stride(PPT);
    const int nseg = 1 << seg_log2;       // lanes that share one pose row
    const int wps = NW >> seg_log2;       // ballot words per lane
    const int ntask = W << seg_log2;
    for (int base = 0; base < ntask; base += COV_THREADS) {
        const int task = base + tid;
        const bool live = task < ntask;
        const int w = live ? (task >> seg_log2) : 0;
        const int seg = task & (nseg - 1);
        const float4* pair = ptab + (size_t)(w >> 1) * COV_PAIR_F4;
        const int hslot = w & 1;
        const float4 v0 = cov_pair_row(pair, hslot, 0), v1 = cov_pair_row(pair, hslot, 1), v2 = cov_pair_row(pair, hslot, 2),
                     v3 = cov_pair_row(pair, hslot, 3), v4 = pair[8 + hslot], v5 = pair[10 + hslot];
        unsigned* brow = bits + (size_t)w * RS;  // every word of a row belongs to exactly one lane: it is cleared once read
        int k = seg * wps;
        const int kend = live ? k + wps : k;
        unsigned word = live ? brow[k] : 0u;
        if (word != 0u) brow[k] = 0u;
        float f0 = 0.f, f1 = 0.f, f2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, se = 0.f, sep = 0.f;
        while (true) {
            while (word == 0u && k + 1 < kend) {
                word = brow[++k];
                if (word != 0u) brow[k] = 0u;
            }
            if (!__any_sync(kFull, word != 0u)) break;
            if (word != 0u) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                const int local = k * 32 + bit;
                CovEval ev;
                const float x = pt[local * 3], y = pt[local * 3 + 1], z = pt[local * 3 + 2];
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
}

// A block's accumulators -> the caller's fp64 rows (fp64 atomics; the reward sum likewise).
template <typename T>
__device__ __forceinline__ void fused_block_flush(const T* accs, int W, double* __restrict__ acc, double sum_r,
                                                  double* red, int tid) {
    for (int i = tid; i < W * 8; i += COV_THREADS) {
        const T v = accs[i];
        if (v != (T)0) atomicAdd(acc + (size_t)(i >> 3) * COV_ACC_STRIDE + (i & 7), (double)v);
    }
    const double ws = cov_warp_sum(sum_r);
    if ((tid & 31) == 0) red[tid >> 5] = ws;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kWarps; ++i) t += red[i];
        if (t != 0.0) atomicAdd(acc + (size_t)W * COV_ACC_STRIDE, t);
    }
}

// ---- dense ----
template <int PPT, bool HAS_UP, int U>
__global__ void __launch_bounds__(COV_THREADS, 2)
cov_traj_fused_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                      const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                      const float* __restrict__ minmax, const float* __restrict__ upstream,
                      const int32_t* __restrict__ out_index, float* __restrict__ rewards, double* __restrict__ acc,
                      int seg_log2) {
    constexpr int T = tile_points(PPT);
    constexpr int RS = bit_stride(PPT);
    // shared memory: pose table | the tile's points (xyz interleaved) | G_j | gate bits | block accumulators
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    float* pt = reinterpret_cast<float*>(smem4) + fused_stage_offset_floats(W);
    float* Gs = pt + T * 3;
    unsigned* bits = reinterpret_cast<unsigned*>(Gs + T);
    float* accs = reinterpret_cast<float*>(bits + (size_t)W * RS);
    __shared__ double red[kWarps];
    __shared__ int amin_pos;  // some pose has min_j m > 0: its arg-min points carry gradient

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) amin_pos = 0;
    __syncthreads();
    const int NP = (W + 1) >> 1;  // pose pairs; an odd W repeats its last pose in the spare slot (never used)
    for (int w = tid; w < 2 * NP; w += COV_THREADS) {
        const int ws = w < W ? w : W - 1;
        float4 row[COV_ROW_F4];
        cov_pose_row(poses + 3 * ws, quats + 4 * ws, K9, C, row);
        const float a = minmax[ws];
        const float b = __fsub_rn(minmax[W + ws], a);
        const float hb = 0.5f * b;
        row[3].w = (a + hb) * (1.f - 9.5367431640625e-07f);  // conservative gate threshold (exact test in the rare path)
        row[4] = make_float4(hb, b, __frcp_rn(b), a);
        cov_pair_store(ptab + (size_t)(w >> 1) * COV_PAIR_F4, w & 1, row);
        if (a > 0.f) amin_pos = 1;  // benign race: every writer stores 1
    }
    for (int i = tid; i < W * 8; i += COV_THREADS) accs[i] = 0.f;
    for (int i = tid; i < W * RS; i += COV_THREADS) bits[i] = 0u;  // phase 1 stores non-zero ballots only; phase 2 clears what it reads
    __syncthreads();

    double sum_r = 0.0;
    const int64_t ntiles = (n + T - 1) / T;
    const bool check_amin = amin_pos != 0;
    const unsigned ptab_s = smem_u32(ptab);
    const CovConst2 C2 = cov_make_const2(C);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        f2_t X[PPT], Y[PPT], Z[PPT];   // the thread's points as broadcast pairs
        float L[PPT];
        bool valid[PPT];
        const int lbase = warp * (32 * PPT) + lane;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const int64_t j = tile * T + lbase + s * 32;
            valid[s] = j < n;
            // a point past the end sits 3e18 m away: m = 0 exactly, never gated, never a tie
            const float x = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;
            const float y = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
            const float z = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
            pt[(lbase + s * 32) * 3] = x;
            pt[(lbase + s * 32) * 3 + 1] = y;
            pt[(lbase + s * 32) * 3 + 2] = z;
            X[s] = f2_pack(x, x);
            Y[s] = f2_pack(y, y);
            Z[s] = f2_pack(z, z);
            L[s] = 0.f;
        }
        bool any_gate = false;  // warp-uniform
        if (check_amin) {  // arg-min points carry gradient: every pair must be looked at
            for (int pr = 0; pr < NP; ++pr) any_gate |= dense_pair_iter<PPT, true>(pr, W, ptab_s, bits, warp, X, Y, Z, L, C, C2, acc, lane);
        } else {
            for (int pr = 0; pr < NP; ++pr) any_gate |= dense_pair_iter<PPT, false>(pr, W, ptab_s, bits, warp, X, Y, Z, L, C, C2, acc, lane);
        }
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float r = 1.f / (1.f + expf(-L[s]));
            float g = r * (1.f - r);
            if (valid[s]) {
                const int64_t j = tile * T + lbase + s * 32;
                const int64_t jo = out_index ? (int64_t)out_index[j] : j;
                rewards[jo] = r;
                sum_r += (double)r;
                if (HAS_UP) g *= upstream[jo];
            }
            Gs[lbase + s * 32] = g;
        }
        // one barrier per tile; the gradient walk (and its closing barrier) only when somebody gated a pair
        if (__syncthreads_or(any_gate ? 1 : 0)) {
            fused_phase2<PPT>(ptab, bits, pt, Gs, accs, W, seg_log2, C, tid);
            __syncthreads();
        }
    }
    fused_block_flush(accs, W, acc, sum_r, red, tid);
}

// ---- pruned: persistent WARPS over the work list of tiles with a non-empty pose mask ----
// The unit of work is one box of 128 consecutive points (one eighth of a listed tile), taken by ONE WARP from a global
// ticket counter: warps never wait for each other (no block barrier, no static assignment), so neither the uneven cost
// of neighbouring boxes nor a heavy tile at the end of the list leaves lanes idle.  A warp keeps two private stages in
// shared memory; its lane 0 draws the next ticket and issues the TMA bulk copies for it (points, the box, the tile's pose
// mask, the slice of the permutation — one mbarrier per stage) before the warp starts on the current item.
//   forward   walk the tile's pose mask; a listed pose is evaluated when its qthr-ball meets the warp's box; gated lanes
//             add their log-odds; the warp notes (one bit per pose) whether it had a gated pair.  Then r_j,
//             G_j = r_j (1 - r_j) per point; rewards stored.
//   backward  the warp walks its noted poses: m again for its points (bit-identical), the gate, dm/dx, the 8 weighted
//             sums in registers; a fixed 9-shuffle tree reduces them over the warp and one lane per component adds the
//             result to the caller's fp64 accumulator row (fire-and-forget RED.F64).
// On the bench cloud ~58 % of the point slots of an evaluated (warp, pose) are gated (profiles/r02_*), so the
// re-evaluation keeps most lanes busy.  The pose table stays in global memory (a box touches ~5 rows: L1 hits).
// `rewards` was pre-filled with 1/2, only other values are stored; the warps add sum_j (r_j - 1/2) of the listed
// tiles to acc[W * STRIDE], which starts at 0.5 * n.
constexpr int kTilesMinBlocks = 3;            // resident blocks per SM the register allocation aims at

template <bool HAS_UP>
__global__ void __launch_bounds__(COV_THREADS, kTilesMinBlocks)
cov_traj_fused_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float4* __restrict__ table, int W, CovConst C,
                            const float* __restrict__ upstream, const int32_t* __restrict__ out_index,
                            float* __restrict__ rewards, double* __restrict__ acc,
                            const float4* __restrict__ boxes, const unsigned* __restrict__ amask_g, int mask_stride,
                            const int* __restrict__ worklist, int* __restrict__ ctrl, int64_t ntiles,
                            unsigned long long* __restrict__ stats) {
    constexpr int PPT = kItemPpt;
    __shared__ __align__(128) float stages[kWarps][2][kItemStageWords];
    __shared__ unsigned wgate[kWarps][kMaskWords];   // per warp: poses with a gated pair among its points (this item)
    __shared__ __align__(8) unsigned long long mbar[kWarps][2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long* bar = mbar[warp];
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const bool check_amin = ctrl[1] != 0;  // some pose has min_j m > 0
    const int n_items = ctrl[0] * kItemsPerTile;
    const int nwords = (W + 31) >> 5;
    unsigned* wg = wgate[warp];
    int* ticket = ctrl + 3;

    // lane 0: stage item `it` (tile = worklist[it / 8], box k = it % 8) unless it is past the end or ragged
    auto issue = [&](int it, int b) {
        if (it >= n_items) return;
        const int64_t tile = worklist[it >> 3];
        const int64_t j0 = tile * (kItemsPerTile * kItemPts) + (int64_t)(it & 7) * kItemPts;
        float* st = stages[warp][b];
        const bool whole = j0 + kItemPts <= n;
        const bool with_perm = whole && out_index != nullptr;
        mbar_expect_tx(&bar[b], (whole ? kItemPts * 12u : 0u) + 32u + (unsigned)mask_stride * 4u + (with_perm ? kItemPts * 4u : 0u));
        if (whole) tma_copy(st, xyz + j0 * 3, kItemPts * 12u, &bar[b]);
        tma_copy(st + kItemPts * 3, boxes + (j0 / kBoxPts) * 2, 32u, &bar[b]);
        tma_copy(st + kItemPts * 3 + 8, amask_g + tile * mask_stride, (unsigned)mask_stride * 4u, &bar[b]);
        if (with_perm) tma_copy(st + kItemPts * 3 + 8 + kMaskWords, out_index + j0, kItemPts * 4u, &bar[b]);
    };

    double sum_r = 0.0;
    unsigned n_box = 0, n_full = 0, n_bwd = 0;
    unsigned uses0 = 0, uses1 = 0;
    int cur = 0;
    if (lane == 0) {
        cur = atomicAdd(ticket, 1);
        issue(cur, 0);
    }
    cur = __shfl_sync(kFull, cur, 0);
    int buf = 0;
    while (cur < n_items) {
        int nxt = 0;
        if (lane == 0) {  // the other stage was last read before the __syncwarp that ended the previous item
            nxt = atomicAdd(ticket, 1);
            issue(nxt, buf ^ 1);
        }
        nxt = __shfl_sync(kFull, nxt, 0);
        if (buf == 0) mbar_wait(&bar[0], uses0++ & 1u);
        else mbar_wait(&bar[1], uses1++ & 1u);
        const float* st = stages[warp][buf];
        const int64_t tile = worklist[cur >> 3];
        const int64_t j0 = tile * (kItemsPerTile * kItemPts) + (int64_t)(cur & 7) * kItemPts;
        const bool whole = j0 + kItemPts <= n;
        // ------------------------------ forward ------------------------------
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
        if (whole) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                valid[s] = true;
                px[s] = st[(s * 32 + lane) * 3];
                py[s] = st[(s * 32 + lane) * 3 + 1];
                pz[s] = st[(s * 32 + lane) * 3 + 2];
                L[s] = 0.f;
            }
        } else {  // the ragged end of the cloud: loaded by hand
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                const int64_t j = j0 + s * 32 + lane;
                valid[s] = j < n;
                // a point past the end sits 3e18 m away: m = 0 exactly, never gated, never a tie
                px[s] = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;
                py[s] = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
                pz[s] = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
                L[s] = 0.f;
            }
        }
        const float4* tb = reinterpret_cast<const float4*>(st + kItemPts * 3);
        const float4 wlo = tb[0], whi = tb[1];
        const unsigned* am = reinterpret_cast<const unsigned*>(st + kItemPts * 3 + 8);
        const int32_t* sperm = reinterpret_cast<const int32_t*>(st + kItemPts * 3 + 8 + kMaskWords);
        bool warp_gated = false;
        for (int c = 0; c < nwords; ++c) {
            unsigned word = am[c];
            unsigned gbits = 0u;
            while (word) {
                const int b = __ffs(word) - 1;
                const int w = c * 32 + b;
                word &= word - 1;
                const float4* row = table + (size_t)w * COV_ROW_F4;
                const float4 v3 = __ldg(row + 3);
                ++n_box;
                // the warp's own box against the pose's qthr-ball; no per-point pre-filter here: it rejected only 11 % of
                // what the box test let through and cost a fifth of the evaluation it saved
                if (box_q2lb(wlo, whi, v3) > v3.w) continue;
                ++n_full;
                double* acc_row = acc + (size_t)w * COV_ACC_STRIDE;
                const bool g = check_amin ? tiles_pose_iter<PPT, true>(row, v3, px, py, pz, L, C, acc_row)
                                          : tiles_pose_iter<PPT, false>(row, v3, px, py, pz, L, C, acc_row);
                if (g) gbits |= 1u << b;
            }
            if (lane == 0) wg[c] = gbits;
            warp_gated |= gbits != 0u;
        }
        float G[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            float r = 0.5f;
            G[s] = 0.25f;  // exactly what the formulas below give for L = 0
            if (L[s] != 0.f) {
                r = 1.f / (1.f + expf(-L[s]));
                G[s] = r * (1.f - r);
            }
            if (valid[s]) {
                const int64_t j = j0 + s * 32 + lane;
                sum_r += (double)(r - 0.5f);
                const bool store = r != 0.5f;
                int64_t jo = j;
                if (out_index && (store || HAS_UP))
                    jo = whole ? (int64_t)sperm[s * 32 + lane] : (int64_t)out_index[j];  // staged with the item
                if (store) rewards[jo] = r;
                if (HAS_UP) G[s] *= upstream[jo];
            }
        }
        // ------------------------------ backward ------------------------------
        if (warp_gated) {
            __syncwarp();  // lane 0's notes are visible to the warp
            for (int c = 0; c < nwords; ++c) {
                unsigned word = wg[c];
                while (word) {
                    const int w = c * 32 + __ffs(word) - 1;
                    word &= word - 1;
                    ++n_bwd;
                    const float4* row = table + (size_t)w * COV_ROW_F4;
                    const float4 v0 = __ldg(row), v1 = __ldg(row + 1), v2 = __ldg(row + 2), v3 = __ldg(row + 3),
                                 v4 = __ldg(row + 4), v5 = __ldg(row + 5);
                    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int s = 0; s < PPT; ++s) {
                        CovEval ev;
                        const float m = cov_vis<true>(px[s], py[s], pz[s], v0, v1, v2, v3, C, &ev);
                        const float d = __fsub_rn(m, v4.w);
                        if (d >= v4.x) {  // exactly p >= 0.5, as in the forward
                            const float p = __fdiv_rn(d, v4.y);
                            if (p <= C.hi) {  // clamp backward gate (inclusive)
                                float gx, gy, gz;
                                cov_vis_grad(m, ev, v0, v1, v2, C, gx, gy, gz);
                                const float yx = px[s] - v5.x, yy = py[s] - v5.y, yz = pz[s] - v5.z;
                                const float e = G[s] * cov_rcp(p * (1.f - p));
                                const float om = e * v4.z;
                                v[0] += om * gx; v[1] += om * gy; v[2] += om * gz;
                                v[3] += om * (gy * yz - gz * yy);
                                v[4] += om * (gz * yx - gx * yz);
                                v[5] += om * (gx * yy - gy * yx);
                                v[6] += e;
                                v[7] += e * p;
                            }
                        }
                    }
                    // 8 sums over 32 lanes in 9 shuffles: halve the number of values a lane carries at offsets 16, 8, 4
                    // (a lane keeps the half its lane bit selects and hands the other half over), then two plain steps;
                    // lanes with (lane & 3) == 0 end with component (lane bits 4,3,2)
                    {
                        const bool up = (lane & 16) != 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float keep = up ? v[q + 4] : v[q], send = up ? v[q] : v[q + 4];
                            v[q] = keep + __shfl_xor_sync(kFull, send, 16);
                        }
                    }
                    {
                        const bool up = (lane & 8) != 0;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float keep = up ? v[q + 2] : v[q], send = up ? v[q] : v[q + 2];
                            v[q] = keep + __shfl_xor_sync(kFull, send, 8);
                        }
                    }
                    {
                        const bool up = (lane & 4) != 0;
                        const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
                        v[0] = keep + __shfl_xor_sync(kFull, send, 4);
                    }
                    v[0] += __shfl_xor_sync(kFull, v[0], 2);
                    v[0] += __shfl_xor_sync(kFull, v[0], 1);
                    if ((lane & 3) == 0 && v[0] != 0.f) {  // one fire-and-forget fp64 RED per component
                        const int comp = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        atomicAdd(acc + (size_t)w * COV_ACC_STRIDE + comp, (double)v[0]);
                    }
                }
            }
        }
        __syncwarp();  // every lane is done with this stage (and with wg) before lane 0 refills it
        cur = nxt;
        buf ^= 1;
    }
    {
        const double ws = cov_warp_sum(sum_r);
        if (lane == 0 && ws != 0.0) atomicAdd(acc + (size_t)W * COV_ACC_STRIDE, ws);
    }
    if (stats && lane == 0) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(stats + 0, (unsigned long long)ntiles * kWarps * W);
        atomicAdd(stats + 1, (unsigned long long)n_full);
        atomicAdd(stats + 4, (unsigned long long)n_bwd);
        atomicAdd(stats + 6, (unsigned long long)n_box);
    }
}

// ---- candidate sweep, pruned (BASELINE config 5): forward only, per-trajectory sum_j r_j ----
// Persistent warps as in the two passes above (128-point items from a ticket counter, private double-buffered TMA
// stages, pose table read through L1).  Poses are trajectory-major (pose = traj * per_traj + i), so walking a tile's pose
// mask in ascending order visits the trajectories one after the other: a warp keeps the log-odds sums of its points for
// the current trajectory, and when the trajectory changes it adds sum_j (sigmoid(L_j) - 1/2) to that trajectory's total
// (fp64 shared atomic; the block adds its totals to the output at the end).  Trajectories no pose of which is listed for
// an item contribute exactly 1/2 per point: 0.5 * n is added once.
__global__ void __launch_bounds__(COV_THREADS, 3)
cov_sweep_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float4* __restrict__ table, int W, int per_traj,
                       int n_traj, CovConst C, const float4* __restrict__ boxes, const unsigned* __restrict__ amask_g,
                       int mask_stride, const int* __restrict__ worklist, int* __restrict__ ctrl,
                       double* __restrict__ sum_out) {
    constexpr int PPT = kItemPpt;
    extern __shared__ float4 smem4[];
    double* ssum = reinterpret_cast<double*>(smem4);   // per-trajectory totals of this block
    __shared__ __align__(128) float stages[kWarps][2][kItemStageWordsA];
    __shared__ __align__(8) unsigned long long mbar[kWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* bar = mbar[warp];
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    for (int t = tid; t < n_traj; t += COV_THREADS) ssum[t] = 0.0;
    __syncthreads();
    const int n_items = ctrl[0] * kItemsPerTile;
    const int nwords = (W + 31) >> 5;
    int* ticket = ctrl + 3;
    auto issue = [&](int it, int b) {
        if (it >= n_items) return;
        const int64_t tile = worklist[it >> 3];
        const int64_t j0 = tile * (kItemsPerTile * kItemPts) + (int64_t)(it & 7) * kItemPts;
        float* st = stages[warp][b];
        const bool whole = j0 + kItemPts <= n;
        mbar_expect_tx(&bar[b], (whole ? kItemPts * 12u : 0u) + 32u + (unsigned)mask_stride * 4u);
        if (whole) tma_copy(st, xyz + j0 * 3, kItemPts * 12u, &bar[b]);
        tma_copy(st + kItemPts * 3, boxes + (j0 / kBoxPts) * 2, 32u, &bar[b]);
        tma_copy(st + kItemPts * 3 + 8, amask_g + tile * mask_stride, (unsigned)mask_stride * 4u, &bar[b]);
    };
    unsigned uses0 = 0, uses1 = 0;
    int cur_it = 0;
    if (lane == 0) {
        cur_it = atomicAdd(ticket, 1);
        issue(cur_it, 0);
    }
    cur_it = __shfl_sync(kFull, cur_it, 0);
    int buf = 0;
    while (cur_it < n_items) {
        int nxt = 0;
        if (lane == 0) {
            nxt = atomicAdd(ticket, 1);
            issue(nxt, buf ^ 1);
        }
        nxt = __shfl_sync(kFull, nxt, 0);
        if (buf == 0) mbar_wait(&bar[0], uses0++ & 1u);
        else mbar_wait(&bar[1], uses1++ & 1u);
        const float* st = stages[warp][buf];
        const int64_t tile = worklist[cur_it >> 3];
        const int64_t j0 = tile * (kItemsPerTile * kItemPts) + (int64_t)(cur_it & 7) * kItemPts;
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
        if (j0 + kItemPts <= n) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                valid[s] = true;
                px[s] = st[(s * 32 + lane) * 3];
                py[s] = st[(s * 32 + lane) * 3 + 1];
                pz[s] = st[(s * 32 + lane) * 3 + 2];
            }
        } else {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                const int64_t j = j0 + s * 32 + lane;
                valid[s] = j < n;
                px[s] = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;  // past the end: m = 0 exactly, never gated
                py[s] = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
                pz[s] = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
            }
        }
        const float4* tb = reinterpret_cast<const float4*>(st + kItemPts * 3);
        const float4 wlo = tb[0], whi = tb[1];
        const unsigned* am = reinterpret_cast<const unsigned*>(st + kItemPts * 3 + 8);
        int cur = -1;      // trajectory whose log-odds sums are in L
        bool touched = false;
        auto flush = [&]() {
            if (cur >= 0 && touched) {
                float a = 0.f;
#pragma unroll
                for (int s = 0; s < PPT; ++s)
                    if (valid[s] && L[s] != 0.f) a += 1.f / (1.f + expf(-L[s])) - 0.5f;
                a = cov_warp_sum(a);
                if (lane == 0 && a != 0.f) atomicAdd(ssum + cur, (double)a);
            }
        };
        for (int c = 0; c < nwords; ++c) {
            unsigned word = am[c];
            while (word) {
                const int w = c * 32 + __ffs(word) - 1;
                word &= word - 1;
                const float4* row = table + (size_t)w * COV_ROW_F4;
                const float4 v3 = __ldg(row + 3);
                if (box_q2lb(wlo, whi, v3) > v3.w) continue;
                const int t = w / per_traj;
                if (t != cur) {
                    flush();
                    cur = t;
                    touched = false;
#pragma unroll
                    for (int s = 0; s < PPT; ++s) L[s] = 0.f;
                }
                touched |= tiles_pose_iter<PPT, false>(row, v3, px, py, pz, L, C, nullptr);
            }
        }
        flush();
        __syncwarp();  // every lane is done with this stage before lane 0 refills it
        cur_it = nxt;
        buf ^= 1;
    }
    __syncthreads();
    for (int t = tid; t < n_traj; t += COV_THREADS) {
        double v = ssum[t];
        if (blockIdx.x == 0) v += 0.5 * (double)n;
        if (v != 0.0) atomicAdd(sum_out + t, v);
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

// ==================================================== host ====================================================
constexpr size_t kSmemCap = (227 - 4) * 1024;  // opt-in shared memory per block on sm_100, minus static use
constexpr int64_t kSeedSamples = 16384;          // pruned pass A: size of the strided sample that seeds the bounds
constexpr int64_t kDenseBelow = 4 * kSeedSamples;  // clouds this small go straight to the dense kernels

// Resident blocks per SM of a kernel at a given dynamic shared-memory size, cached per (kernel, device, size): the
// opt-in shared-memory attribute is raised once per kernel and device to the cap (not per call: two host threads with
// different pose counts must not race between "set" and "launch"), the occupancy query runs once per size.
struct OccKey {
    const void* fn;
    int dev;
    size_t smem;
    bool operator<(const OccKey& o) const {
        return fn != o.fn ? fn < o.fn : dev != o.dev ? dev < o.dev : smem < o.smem;
    }
};
std::mutex g_occ_mutex;
std::map<OccKey, int> g_occ;

template <typename Kern>
int blocks_per_sm(Kern kern, int threads, size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    const OccKey key{reinterpret_cast<const void*>(kern), dev, smem};
    std::lock_guard<std::mutex> lock(g_occ_mutex);
    auto it = g_occ.find(key);
    if (it != g_occ.end()) return it->second;
    const OccKey attr_key{reinterpret_cast<const void*>(kern), dev, (size_t)-1};
    if (g_occ.find(attr_key) == g_occ.end()) {
        cudaFuncAttributes fa;
        size_t cap = kSmemCap;
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess && fa.sharedSizeBytes + cap > 227 * 1024)
            cap = 227 * 1024 - fa.sharedSizeBytes;  // opt-in limit covers static + dynamic
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap) != cudaSuccess)
            cudaGetLastError();
        g_occ[attr_key] = 1;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    g_occ[key] = per_sm;
    return per_sm;
}

template <typename Kern>
int grid_for(Kern kern, size_t smem, int64_t ntiles) {
    int64_t g = (int64_t)blocks_per_sm(kern, COV_THREADS, smem) * cov_sm_count_cached();
    if (g > ntiles) g = ntiles;
    if (g > COV_MAX_GRID) g = COV_MAX_GRID;
    return g < 1 ? 1 : (int)g;
}

// points per thread of the DENSE kernels: the largest tile that still gives every SM two tiles and fits shared memory
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

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
int64_t boxes_padded(int64_t n) { return ((n + 2047) / 2048) * (2048 / kBoxPts); }

// Workspace of both passes (every region 256-byte aligned):
//   [control: ctrl ints | encoded extrema 2W | look-back descriptors][pose table][work list][tile masks][boxes]
struct TrajWorkspace {
    int* ctrl;
    unsigned* enc;
    unsigned long long* desc;
    size_t ctrl_bytes;  // the span a pruned call zeroes first
    float4* table;
    int* worklist;
    unsigned* amask;
    float4* boxes;
    size_t bytes;
};
TrajWorkspace carve_workspace(void* ws, int64_t n, int W) {
    const int64_t nt = (n + 255) / 256;  // tiles at the smallest tile size
    const int64_t nchunks = (nt + kChunkTiles - 1) / kChunkTiles;
    char* p = reinterpret_cast<char*>(ws);
    size_t off = 0;
    TrajWorkspace t;
    auto take = [&](size_t bytes) { char* q = p + off; off += align256(bytes); return q; };
    const size_t enc_off = kCtrlInts * sizeof(int);
    const size_t desc_off = (enc_off + 2 * (size_t)W * sizeof(unsigned) + 7) & ~(size_t)7;
    t.ctrl_bytes = desc_off + (size_t)nchunks * sizeof(unsigned long long);
    char* c = take(t.ctrl_bytes);
    t.ctrl = reinterpret_cast<int*>(c);
    t.enc = reinterpret_cast<unsigned*>(c + enc_off);
    t.desc = reinterpret_cast<unsigned long long*>(c + desc_off);
    t.table = reinterpret_cast<float4*>(take((size_t)W * COV_ROW_F4 * sizeof(float4)));
    t.worklist = reinterpret_cast<int*>(take((size_t)nt * sizeof(int)));
    t.amask = reinterpret_cast<unsigned*>(take((size_t)nt * mask_stride_words(W) * sizeof(unsigned)));
    t.boxes = reinterpret_cast<float4*>(take((size_t)boxes_padded(n) * 2 * sizeof(float4)));
    t.bytes = off;
    return t;
}

int check_traj_args(const char* who, const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                    const float* K, const cov_camera* cam, const void* ws, size_t ws_bytes, bool pruned) {
    if (!xyz || n <= 0 || !poses || !quats || W <= 0 || !K || !cam) {
        cov_set_error("%s: null pointer, empty cloud or no poses (n=%lld, W=%d)", who, (long long)n, W);
        return COV_ERR_ARG;
    }
    // the dense kernels keep the pose table in shared memory; the pruned ones only a bit per pose in the tile masks
    const int max_w = pruned ? cov_traj_max_poses_pruned() : cov_traj_max_poses();
    if (W > max_w) {
        cov_set_error("%s: %d poses exceed what one call takes on this path (max %d)", who, W, max_w);
        return COV_ERR_UNSUPPORTED;
    }
    if (n >= ((int64_t)1 << 31) * 256) {
        cov_set_error("%s: %lld points exceed the int32 tile index range", who, (long long)n);
        return COV_ERR_UNSUPPORTED;
    }
    if (!ws) {
        cov_set_error("%s: null workspace", who);
        return COV_ERR_ARG;
    }
    if (ws_bytes < cov_traj_workspace_bytes(n, W)) {
        cov_set_error("%s: workspace %zu < %zu bytes", who, ws_bytes, cov_traj_workspace_bytes(n, W));
        return COV_ERR_WORKSPACE;
    }
    if ((((uintptr_t)ws) & 255) || (((uintptr_t)xyz) & 15)) {
        cov_set_error("%s: workspace must be 256-byte aligned and the cloud 16-byte aligned", who);
        return COV_ERR_ALIGN;
    }
    return COV_OK;
}

// boxes for this call: the caller's (cov_tile_boxes, once per cloud) or built into the workspace now
const float4* boxes_for_call(const float* xyz, int64_t n, const float* boxes_dev, const TrajWorkspace& t, cudaStream_t s) {
    if (boxes_dev) return reinterpret_cast<const float4*>(boxes_dev);
    const int64_t nb = boxes_padded(n);
    const int grid = (int)std::min<int64_t>((nb + 7) / 8, (int64_t)cov_sm_count_cached() * 16);
    cov_tile_boxes_kernel<<<grid, 256, 0, s>>>(xyz, n, nb, t.boxes);
    return t.boxes;
}

// cull + work list in one launch; every block must be resident (the blocks wait on one another's descriptors)
int launch_cull(const float4* boxes, int ppt, int64_t ntiles, const TrajWorkspace& t, int W, const unsigned* enc,
                float inv_kd, float* fill_dst, int64_t fill_n, cudaStream_t s) {
    const size_t smem = (size_t)W * sizeof(float4);
    const int64_t nchunks = (ntiles + kChunkTiles - 1) / kChunkTiles;
    const int64_t resident = (int64_t)blocks_per_sm(cov_cull_kernel, 256, smem) * cov_sm_count_cached();
    int64_t want = nchunks;
    if (fill_dst) want = std::max<int64_t>(want, (int64_t)cov_sm_count_cached() * 4);  // enough stores in flight for the fill
    int64_t g = std::max<int64_t>(1, std::min<int64_t>(want, resident));
    if ((nchunks + g - 1) / g > kCullMaxIter) {
        cov_set_error("cov_cull: %lld tiles exceed what one launch can list (%lld blocks x %d chunks)", (long long)ntiles,
                      (long long)g, kCullMaxIter);
        return COV_ERR_UNSUPPORTED;
    }
    const int grid = (int)g;
    cov_cull_kernel<<<grid, 256, smem, s>>>(boxes, tile_boxes(ppt), ntiles, t.table, W, enc, inv_kd, t.amask,
                                            mask_stride_words(W), t.worklist, t.ctrl, t.desc, fill_dst, fill_n);
    return COV_OK;
}

}  // namespace

extern "C" int cov_traj_max_poses(void) {
    static int cached = 0;
    if (cached) return cached;
    int w = 1;
    while (w < 32 * kMaskWords && fused_smem_bytes(w + 1, 1) <= kSmemCap && minmax_smem_bytes(w + 1) <= kSmemCap &&
           (size_t)(w + 1) * sizeof(float4) <= 48 * 1024)
        ++w;
    cached = w;
    return w;
}

extern "C" int cov_traj_max_poses_pruned(void) { return 32 * kMaskWords; }

extern "C" size_t cov_traj_workspace_bytes(int64_t n, int n_poses) {
    if (n < 1) n = 1;
    if (n_poses < 1) n_poses = 1;
    return carve_workspace(nullptr, n, n_poses).bytes;
}

extern "C" int64_t cov_tile_boxes_count(int64_t n) { return n > 0 ? boxes_padded(n) : 0; }

extern "C" int cov_traj_prefill_applies(int64_t n, const cov_traj_opts* opts) {
    return (!(opts && opts->dense) && n >= kDenseBelow) ? 1 : 0;
}

extern "C" int cov_tile_boxes(const float* xyz, int64_t n, float* boxes, void* stream) {
    if (!xyz || !boxes || n <= 0) {
        cov_set_error("cov_tile_boxes: null pointer or empty cloud (n=%lld)", (long long)n);
        return COV_ERR_ARG;
    }
    if (((uintptr_t)boxes) & 15) {
        cov_set_error("cov_tile_boxes: boxes must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    const int64_t nb = boxes_padded(n);
    const int grid = (int)std::min<int64_t>((nb + 7) / 8, (int64_t)cov_sm_count_cached() * 16);
    cov_tile_boxes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, n, nb, reinterpret_cast<float4*>(boxes));
    return cov_check_launch("cov_tile_boxes");
}

extern "C" int cov_traj_minmax(const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                               const float* K, const cov_camera* cam, const float* boxes_dev, float* minmax,
                               const cov_traj_opts* opts, void* ws, size_t ws_bytes, void* stream) {
    const bool prune = !(opts && opts->dense) && n >= kDenseBelow;
    int rc = check_traj_args("cov_traj_minmax", xyz, n, poses, quats, W, K, cam, ws, ws_bytes, prune);
    if (rc) return rc;
    if (!minmax) {
        cov_set_error("cov_traj_minmax: null minmax");
        return COV_ERR_ARG;
    }
    if (boxes_dev && (((uintptr_t)boxes_dev) & 15)) {
        cov_set_error("cov_traj_minmax: boxes must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    int ppt = prune ? 4 : pick_ppt(n, W, false);
    if (ppt == 4 && (n + tile_points(8) - 1) / tile_points(8) >= 4 * (int64_t)cov_sm_count_cached()) ppt = 8;
    if (ppt == 0) {
        cov_set_error("cov_traj_minmax: %d poses do not fit in shared memory", W);
        return COV_ERR_UNSUPPORTED;
    }
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
    if (!prune) {
        // (prefill_dev is not honoured here: the dense pass B writes every reward itself)
        unsigned* gmin = reinterpret_cast<unsigned*>(minmax);
        unsigned* gmax = gmin + W;
        const size_t smem = minmax_smem_bytes(W);
        cov_minmax_init_kernel<<<(W + 255) / 256, 256, 0, s>>>(gmin, gmax, W);
#define LAUNCH_DENSE(P, B)                                                                                       \
    {                                                                                                            \
        const int grid = grid_for(cov_traj_minmax_kernel<P, B>, smem, ntiles);                                   \
        cov_traj_minmax_kernel<P, B><<<grid, COV_THREADS, smem, s>>>(xyz, n, poses, quats, W, K, C, gmin, gmax); \
    }
        if (ppt == 8) LAUNCH_DENSE(8, 2)
        else if (ppt == 4) LAUNCH_DENSE(4, 2)
        else if (ppt == 2) LAUNCH_DENSE(2, 2)
        else LAUNCH_DENSE(1, 2)
#undef LAUNCH_DENSE
        return cov_check_launch("cov_traj_minmax");
    }
    // pruned: pose table + seed on a strided sample, cull + work list, evaluation of the listed tiles: 3 launches
    const TrajWorkspace t = carve_workspace(ws, n, W);
    const float4* boxes = boxes_for_call(xyz, n, boxes_dev, t, s);
    constexpr int ppt_t = 4;  // tiles of 1024 points, eight 128-point items each (as pass B)
    const int64_t ntiles_t = (n + tile_points(ppt_t) - 1) / tile_points(ppt_t);
    cudaMemsetAsync(t.ctrl, 0, t.ctrl_bytes, s);
    {
        const int64_t stride = (n + kSeedSamples - 1) / kSeedSamples;
        const int64_t nsamples = (n + stride - 1) / stride;
        const dim3 sgrid((unsigned)((nsamples + COV_THREADS - 1) / COV_THREADS), (unsigned)((W + 31) / 32));
        cov_traj_prepare_kernel<<<sgrid, COV_THREADS, 0, s>>>(xyz, nsamples, stride, poses, quats, W, K, C, t.table, t.enc);
    }
    if ((rc = launch_cull(boxes, ppt_t, ntiles_t, t, W, t.enc, 1.f / C.kd, nullptr, 0, s)) != COV_OK) return rc;
    {
        const size_t smem_t = (size_t)W * 12;
        int64_t grid = (int64_t)blocks_per_sm(cov_traj_minmax_tiles_kernel, COV_THREADS, smem_t) * cov_sm_count_cached();
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, ntiles_t));
        cov_traj_minmax_tiles_kernel<<<(unsigned)grid, COV_THREADS, smem_t, s>>>(
            xyz, n, t.table, W, C, t.enc, minmax, boxes, t.amask, mask_stride_words(W), t.worklist, t.ctrl, ntiles_t,
            opts ? opts->stats_dev : nullptr, opts ? opts->prefill_dev : nullptr, n);
    }
    return cov_check_launch("cov_traj_minmax");
}

extern "C" int cov_traj_fused(const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                              const float* K, const cov_camera* cam, const float* boxes_dev, const float* minmax,
                              const float* upstream, const int32_t* reward_index, float* rewards, double* acc,
                              const cov_traj_opts* opts, void* ws, size_t ws_bytes, void* stream) {
    const bool prune = !(opts && opts->dense) && n >= kDenseBelow;
    int rc = check_traj_args("cov_traj_fused", xyz, n, poses, quats, W, K, cam, ws, ws_bytes, prune);
    if (rc) return rc;
    if (!minmax || !rewards || !acc) {
        cov_set_error("cov_traj_fused: null minmax/rewards/acc");
        return COV_ERR_ARG;
    }
    if ((boxes_dev && (((uintptr_t)boxes_dev) & 15)) || (reward_index && (((uintptr_t)reward_index) & 15))) {
        cov_set_error("cov_traj_fused: boxes and reward_index must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    const TrajWorkspace t = carve_workspace(ws, n, W);
    if (!prune) {  // every pair evaluated: accumulators zeroed, one kernel
        const int ppt_d = pick_ppt(n, W, true);
        if (ppt_d == 0) {
            cov_set_error("cov_traj_fused: %d poses do not fit in shared memory", W);
            return COV_ERR_UNSUPPORTED;
        }
        cudaMemsetAsync(acc, 0, ((size_t)W * COV_ACC_STRIDE + 1) * sizeof(double), s);
        const size_t smem = fused_smem_bytes(W, ppt_d);
        const int64_t ntiles = (n + tile_points(ppt_d) - 1) / tile_points(ppt_d);
        // phase-2 parallelism: split each pose row over 2^seg_log2 lanes until there are >= 2 tasks per thread
        int seg_log2 = 0;
        while ((W << seg_log2) < 2 * COV_THREADS && (2 << seg_log2) <= bit_words(ppt_d) && seg_log2 < 5) ++seg_log2;
#define LAUNCH_F(P, UP)                                                                                           \
    {                                                                                                             \
        const int grid = grid_for(cov_traj_fused_kernel<P, UP, 1>, smem, ntiles);                                 \
        cov_traj_fused_kernel<P, UP, 1><<<grid, COV_THREADS, smem, s>>>(xyz, n, poses, quats, W, K, C, minmax,    \
                                                                        upstream, reward_index, rewards, acc, seg_log2); \
    }
        if (upstream) {
            if (ppt_d == 4) LAUNCH_F(4, true) else if (ppt_d == 2) LAUNCH_F(2, true) else LAUNCH_F(1, true)
        } else {
            if (ppt_d == 4) LAUNCH_F(4, false) else if (ppt_d == 2) LAUNCH_F(2, false) else LAUNCH_F(1, false)
        }
#undef LAUNCH_F
        return cov_check_launch("cov_traj_fused");
    }
    // pruned: table (+ zeroing), cull + work list (+ rewards pre-fill), evaluation of the listed tiles: 3 launches
    constexpr int ppt = kItemPpt * kItemsPerTile * 32 / COV_THREADS;  // tiles of 1024 points, eight 128-point items each
    static_assert(ppt == 4 && tile_points(ppt) == kItemsPerTile * kItemPts, "tile = 8 items");
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
    const float4* boxes = boxes_for_call(xyz, n, boxes_dev, t, s);
    cov_traj_table_kernel<<<(W + 63) / 64, 64, 0, s>>>(poses, quats, W, K, C, minmax, t.table, t.ctrl,
                                                       (int)(t.ctrl_bytes / sizeof(int)), acc, 0.5 * (double)n);
    const bool prefilled = opts && opts->rewards_prefilled;
    if ((rc = launch_cull(boxes, ppt, ntiles, t, W, nullptr, 0.f, prefilled ? nullptr : rewards, n, s)) != COV_OK) return rc;
    unsigned long long* stats = opts ? opts->stats_dev : nullptr;
    {
        // persistent warps: as many blocks as are resident; every warp draws 128-point items from a ticket counter
        int64_t grid = 0;
        if (upstream) grid = (int64_t)blocks_per_sm(cov_traj_fused_tiles_kernel<true>, COV_THREADS, 0) * cov_sm_count_cached();
        else grid = (int64_t)blocks_per_sm(cov_traj_fused_tiles_kernel<false>, COV_THREADS, 0) * cov_sm_count_cached();
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, ntiles));
        if (upstream)
            cov_traj_fused_tiles_kernel<true><<<(unsigned)grid, COV_THREADS, 0, s>>>(
                xyz, n, t.table, W, C, upstream, reward_index, rewards, acc, boxes, t.amask, mask_stride_words(W),
                t.worklist, t.ctrl, ntiles, stats);
        else
            cov_traj_fused_tiles_kernel<false><<<(unsigned)grid, COV_THREADS, 0, s>>>(
                xyz, n, t.table, W, C, upstream, reward_index, rewards, acc, boxes, t.amask, mask_stride_words(W),
                t.worklist, t.ctrl, ntiles, stats);
    }
    return cov_check_launch("cov_traj_fused");
}

extern "C" int cov_sweep_rewards(const float* xyz, int64_t n, const float* poses, const float* quats, int n_traj,
                                 int per_traj, const float* K, const cov_camera* cam, const float* boxes_dev,
                                 const float* minmax, double* sum_rewards, const cov_traj_opts* opts, void* ws,
                                 size_t ws_bytes, void* stream) {
    if (!xyz || n <= 0 || !poses || !quats || n_traj <= 0 || per_traj <= 0 || !K || !cam || !minmax || !sum_rewards) {
        cov_set_error("cov_sweep_rewards: bad argument");
        return COV_ERR_ARG;
    }
    if (opts && opts->dense)
        return cov_sweep_rewards_dense(xyz, n, poses, quats, n_traj, per_traj, K, cam, minmax, sum_rewards, stream);
    // trajectories per launch: as many as the per-tile pose mask holds (32 * kMaskWords poses); the pose table stays in
    // global memory, the block keeps one fp64 total per trajectory in shared memory
    if (per_traj > 32 * kMaskWords) {
        cov_set_error("cov_sweep_rewards: %d poses per trajectory exceed the pose mask (%d)", per_traj, 32 * kMaskWords);
        return COV_ERR_UNSUPPORTED;
    }
    int chunk = (32 * kMaskWords) / per_traj;
    if (chunk > n_traj) chunk = n_traj;
    const int Wc = chunk * per_traj;
    if (!ws || ws_bytes < cov_sweep_workspace_bytes(n, n_traj, per_traj) || (((uintptr_t)ws) & 255) || (((uintptr_t)xyz) & 15)) {
        cov_set_error("cov_sweep_rewards: workspace missing, misaligned or smaller than cov_sweep_workspace_bytes(n, %d, %d) = %zu",
                      n_traj, per_traj, cov_sweep_workspace_bytes(n, n_traj, per_traj));
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    const TrajWorkspace t = carve_workspace(ws, n, Wc);
    const float4* boxes = boxes_for_call(xyz, n, boxes_dev, t, s);
    const int W = n_traj * per_traj;
    constexpr int ppt = 4;  // tiles of 1024 points, eight 128-point items each
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
    float* mm = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + t.bytes);  // the chunk's minima and maxima, contiguous
    for (int t0 = 0; t0 < n_traj; t0 += chunk) {
        const int nt = (n_traj - t0 < chunk) ? n_traj - t0 : chunk;
        const int w0 = t0 * per_traj, Wn = nt * per_traj;
        cudaMemcpyAsync(mm, minmax + w0, (size_t)Wn * sizeof(float), cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(mm + Wn, minmax + W + w0, (size_t)Wn * sizeof(float), cudaMemcpyDeviceToDevice, s);
        cov_traj_table_kernel<<<(Wn + 63) / 64, 64, 0, s>>>(poses + 3 * (size_t)w0, quats + 4 * (size_t)w0, Wn, K, C, mm, t.table,
                                                            t.ctrl, (int)(t.ctrl_bytes / sizeof(int)), nullptr, 0.0);
        if (launch_cull(boxes, ppt, ntiles, t, Wn, nullptr, 0.f, nullptr, 0, s) != COV_OK) return COV_ERR_UNSUPPORTED;
        const size_t smem = (size_t)nt * sizeof(double);
        int64_t grid = (int64_t)blocks_per_sm(cov_sweep_tiles_kernel, COV_THREADS, smem) * cov_sm_count_cached();
        grid = std::max<int64_t>(1, std::min<int64_t>(grid, ntiles));
        cov_sweep_tiles_kernel<<<(unsigned)grid, COV_THREADS, smem, s>>>(xyz, n, t.table, Wn, per_traj, nt, C, boxes, t.amask,
                                                                        mask_stride_words(Wn), t.worklist, t.ctrl,
                                                                        sum_rewards + t0);
    }
    return cov_check_launch("cov_sweep_rewards");
}

extern "C" size_t cov_sweep_workspace_bytes(int64_t n, int n_traj, int per_traj) {
    if (n < 1) n = 1;
    int W = n_traj * per_traj;
    if (W > 32 * kMaskWords) W = 32 * kMaskWords;
    if (W < 1) W = 1;
    return cov_traj_workspace_bytes(n, W) + align256(2 * (size_t)W * sizeof(float));
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
