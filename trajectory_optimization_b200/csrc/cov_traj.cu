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
// DENSE (cov_set_pruning(0)): every (point, pose) pair is evaluated.  FP32-issue bound.
//   pass A: thread-local fmin/fmax over the thread's points, one integer REDUX per warp (m >= 0, so the float order
//           is the uint order), one shared atomic per warp and 32 poses, one global atomic per block and pose.
//   pass B, per tile of 256*PPT points:
//     phase 1  every pair: m, gate (m - a >= b/2  <=>  p >= 0.5, exact), warp ballot of the gate into a pose-major
//              bit matrix in shared memory; gated lanes add their log-odds to the point's running sum in pose order.
//     phase 2  the bit matrix is walked pose-major: a lane owns (pose, row segment), pops its set bits, re-evaluates
//              m and dm/dx for that pair and accumulates the 8 weighted sums in registers — no atomics, fixed order.
//
// PRUNED (default): cull -> compact -> evaluate, on boxes of 128 consecutive points (cov_tile_boxes; tight when the
//   cloud is Morton-ordered by cov_spatial_sort, valid for any order).  A pair whose distance Gaussian alone bounds m
//   below what can matter is never evaluated:  m <= 2^-(kd q2)(1+1.3e-5)  and  q2 >= box bound > qcap  =>  m < bound.
//     pass A: bound = a lower bound of max_j m from a strided sample of the cloud (seed launch), valid once the
//             minimum is known to be exactly 0 (skipped points have m >= 0);
//     pass B: bound = the conservative gate threshold (a + b/2)(1 - 2^-20): a pair below it is neither gated nor
//             the arg-max, adds logit(1/2) = 0 to the log-odds sum and nothing to the gradient.
//   cov_cull_kernel      one warp per tile: union of the tile's boxes against every pose -> per-tile pose bit mask
//   cov_worklist_kernel  list of the tiles with a non-empty mask, heaviest first (single block scan: deterministic)
//   *_tiles_kernel       persistent blocks walk the work list; tile points, boxes and mask arrive together through a
//                        double-buffered TMA bulk copy (one mbarrier per buffer); a warp evaluates a listed pose only
//                        if its own 128-point box and then one of its points pass the same test.
//   Tiles that are not listed are never read: their rewards are the pre-filled 1/2 (exact), their sum is 0.5 * count.
//
// Block accumulators go to a [block][W][8] fp32 slab; a small kernel adds the slabs in fp64 in a fixed order.
// The arg-max / arg-min tie sets of the normalisation backward are rare (one point per pose unless the minimum
// underflowed to 0, in which case their gradient is exactly negligible and skipped) and go straight to the fp64
// accumulator with atomics.
#include <algorithm>
#include <cstdlib>

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
// one staged tile of the pruned kernels: points | boxes | pose mask (all 16-byte multiples, buffer 128-byte multiple)
__host__ __device__ constexpr int stage_floats(int ppt) { return tile_points(ppt) * 3 + tile_boxes(ppt) * 8 + kMaskWords; }

size_t minmax_smem_bytes(int W) { return (size_t)W * (COV_ROW_F4 * sizeof(float4) + 2 * sizeof(unsigned)); }
__host__ __device__ inline int minmax_tiles_stage_offset_floats(int W) {
    return (int)((((size_t)W * (COV_ROW_F4 * 16 + 12) + 127) & ~(size_t)127) / 4);
}
size_t minmax_tiles_smem_bytes(int W, int ppt) {
    return (size_t)minmax_tiles_stage_offset_floats(W) * 4 + 2 * (size_t)stage_floats(ppt) * 4;
}
// fused pass: pose table, then (128-byte aligned) the tile stage(s), G_j, gate bits, accumulators, gated-pose list
__host__ __device__ inline int fused_stage_offset_floats(int W) {
    return (int)((((size_t)W * COV_ROW_F4 * 16 + 127) & ~(size_t)127) / 4);
}
size_t fused_smem_bytes(int W, int ppt, bool prune) {
    if (prune)  // pose table | 2 tile stages | block accumulators
        return (size_t)fused_stage_offset_floats(W) * 4 + 2 * (size_t)(stage_floats(ppt) + tile_points(ppt)) * 4 +
               (size_t)W * 8 * sizeof(float);
    // pose table | tile points | G_j | gate bits | block accumulators
    return (size_t)fused_stage_offset_floats(W) * 4 + (size_t)tile_points(ppt) * 12 + (size_t)tile_points(ppt) * 4 +
           (size_t)W * bit_stride(ppt) * sizeof(unsigned) + (size_t)W * 8 * sizeof(float);
}

// Unweighted dm/dy and dm/dy x y of one (point, pose) into a tie-set accumulator (7 doubles).  Rare (one point per
// pose and step), so it is a real call; it takes the pose INDEX and finds the row in the pose table at the start of
// dynamic shared memory (where every kernel of this file keeps it) — passing a row pointer would make the callers
// compute a generic shared-memory address on every iteration of their hot loops.
__device__ __noinline__ void tie_accumulate(float x, float y, float z, int w, CovConst C, double* acc, int slot) {
    extern __shared__ float4 smem4[];
    const float4* row = smem4 + (size_t)w * COV_ROW_F4;
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

// Pose-table row loads by 32-bit shared address (the address is formed once per kernel; with generic pointers the
// compiler re-derives the shared window base inside the hot loops).  The table is constant after the prologue barrier.
__device__ __forceinline__ float4 lds_row(unsigned table_saddr, int w, int i) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"(table_saddr + (unsigned)(w * COV_ROW_F4 + i) * 16u));
    return v;
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

// One staged tile: points (unless the tile is the ragged last one), its boxes and its pose mask — and, for pass B on an
// ordered cloud, the tile's slice of the permutation — on one mbarrier.
template <int PPT>
__device__ __forceinline__ void stage_issue(float* stage, unsigned long long* bar, const float* __restrict__ xyz,
                                            const float4* __restrict__ boxes, const unsigned* __restrict__ amask_g,
                                            int mask_stride, int64_t tile, int64_t nfull,
                                            const int32_t* __restrict__ perm = nullptr) {
    constexpr int T = tile_points(PPT);
    constexpr int NB = tile_boxes(PPT);
    const bool whole = tile < nfull;
    mbar_expect_tx(bar, (whole ? T * 12u : 0u) + NB * 32u + (unsigned)mask_stride * 4u + ((whole && perm) ? T * 4u : 0u));
    if (whole) tma_copy(stage, xyz + tile * (T * 3), T * 12u, bar);
    tma_copy(stage + T * 3, boxes + tile * (NB * 2), NB * 32u, bar);
    tma_copy(stage + T * 3 + NB * 8, amask_g + tile * mask_stride, (unsigned)mask_stride * 4u, bar);
    if (whole && perm) tma_copy(stage + stage_floats(PPT), perm + tile * T, T * 4u, bar);
}

// =============================================== pass A, dense ===============================================
// Also the SEED launch of the pruned pass (point_stride > 1: every point_stride-th point, n = number of samples).
template <int PPT, int MINB, int U>
__global__ void __launch_bounds__(COV_THREADS, MINB)
cov_traj_minmax_kernel(const float* __restrict__ xyz, int64_t n, int64_t point_stride, const float* __restrict__ poses,
                       const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                       unsigned* __restrict__ gmin, unsigned* __restrict__ gmax) {
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    unsigned* smin = reinterpret_cast<unsigned*>(ptab + (size_t)W * COV_ROW_F4);
    unsigned* smax = smin + W;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int w = tid; w < W; w += COV_THREADS) {
        cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, ptab + (size_t)w * COV_ROW_F4);
        smin[w] = 0x7f800000u;
        smax[w] = 0u;
    }
    __syncthreads();
    constexpr int T = tile_points(PPT);
    const int64_t ntiles = (n + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float px[PPT], py[PPT], pz[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            int64_t j = tile * T + s * COV_THREADS + tid;
            j = (j < n ? j : n - 1) * point_stride;  // a duplicate cannot change a min or a max
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
                    const float4 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3];
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
            }
        }
    }
    __syncthreads();
    for (int w = tid; w < W; w += COV_THREADS) {
        atomicMin(gmin + w, smin[w]);
        atomicMax(gmax + w, smax[w]);
    }
}

__global__ void __launch_bounds__(256) cov_fill_kernel(float* __restrict__ dst, int64_t n, float v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(dst + i) & 15) == 0)) {
            *reinterpret_cast<float4*>(dst + i) = make_float4(v, v, v, v);
        } else {
            for (int64_t k = i; k < n && k < i + 4; ++k) dst[k] = v;
        }
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
// Pose table in global memory, built once per call (the pruned kernels copy it to shared memory).  With `minmax`
// (pass B) the rows carry the normalisation constants: v3.w = qthr, v4 = (b/2, b, 1/b, a), v5.w = thr.
// flags[1] is set when some pose has min_j m > 0.
__global__ void cov_pose_table_kernel(const float* __restrict__ poses, const float* __restrict__ quats, int W,
                                      const float* __restrict__ K9, CovConst C, const float* __restrict__ minmax,
                                      float4* __restrict__ table, int* __restrict__ flags) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    float4 row[COV_ROW_F4];
    cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, row);
    if (minmax) {
        const float a = minmax[w];
        const float b = __fsub_rn(minmax[W + w], a);
        const float hb = 0.5f * b;
        const float thr = (a + hb) * (1.f - 9.5367431640625e-07f);  // conservative gate threshold
        row[5].w = thr;
        // q2 above qthr cannot reach thr (thr <= 0 or NaN, or a > 0 — arg-min points carry gradient: never prune)
        row[3].w = (thr > 0.f && !(a > 0.f)) ? (float)((1e-4 - log2((double)thr)) / (double)C.kd * 1.000001)
                                             : __uint_as_float(0x7f800000u);
        row[4] = make_float4(hb, b, __frcp_rn(b), a);
        if (a > 0.f) atomicOr(flags + 1, 1);
    }
#pragma unroll
    for (int i = 0; i < COV_ROW_F4; ++i) table[(size_t)w * COV_ROW_F4 + i] = row[i];
}

// Seed of the pruned pass A: every `stride`-th point against all poses (dense), so that the cull starts from a good
// lower bound of each maximum and knows which minima are exactly 0.  Grid = (sample blocks) x (chunks of 32 poses):
// a thread owns one sample point, a block 32 pose rows of the table; per pose one REDUX pair per warp, the 8 warps
// meet in shared memory, 32 global atomics per block.
__global__ void __launch_bounds__(COV_THREADS)
cov_traj_seed_kernel(const float* __restrict__ xyz, int64_t nsamples, int64_t stride, const float4* __restrict__ table,
                     int W, CovConst C, unsigned* __restrict__ gmin, unsigned* __restrict__ gmax) {
    __shared__ float4 rows[32 * 4];
    __shared__ unsigned smn[32], smx[32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int w0 = blockIdx.y * 32;
    const int wn = (W - w0 < 32) ? (W - w0) : 32;
    if (tid < wn * 4) rows[tid] = table[(size_t)(w0 + (tid >> 2)) * COV_ROW_F4 + (tid & 3)];
    if (tid < 32) {
        smn[tid] = 0x7f800000u;
        smx[tid] = 0u;
    }
    __syncthreads();
    int64_t j = (int64_t)blockIdx.x * COV_THREADS + tid;
    j = (j < nsamples ? j : nsamples - 1) * stride;  // a duplicate cannot change a min or a max
    const float x = __ldg(xyz + j * 3), y = __ldg(xyz + j * 3 + 1), z = __ldg(xyz + j * 3 + 2);
    unsigned keep_mn = 0x7f800000u, keep_mx = 0u;
    for (int i = 0; i < wn; ++i) {
        const float m = cov_vis<false>(x, y, z, rows[4 * i], rows[4 * i + 1], rows[4 * i + 2], rows[4 * i + 3], C, nullptr);
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
        atomicMin(gmin + w0 + tid, smn[tid]);
        atomicMax(gmax + w0 + tid, smx[tid]);
    }
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

// Cull: mask[tile][c] bit i <=> pose 32c+i can matter for some point of the tile; flags[tile] = the mask is not
// empty.  One warp per GROUP of 8 consecutive tiles: every pose is tested against the group's box first (one lane
// per pose), and only the few that pass are tested against the 8 tile boxes (one lane per tile).  Pass A passes
// gmin/gmax (after the seed launch) and the cap is derived here; pass B reads qthr from the table.
__global__ void __launch_bounds__(256)
cov_cull_kernel(const float4* __restrict__ boxes, int boxes_per_tile, int64_t ntiles, const float4* __restrict__ table,
                int W, const unsigned* __restrict__ gmin, const unsigned* __restrict__ gmax, float inv_kd,
                unsigned* __restrict__ amask_g, int mask_stride, unsigned char* __restrict__ flags,
                unsigned long long* __restrict__ listed_pairs) {
    extern __shared__ float4 v3s[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int w = tid; w < W; w += blockDim.x) {
        float4 v3 = table[(size_t)w * COV_ROW_F4 + 3];
        if (gmin) v3.w = minmax_qcap(gmin[w], gmax[w], inv_kd);
        v3s[w] = v3;
    }
    __syncthreads();
    const float inf = __uint_as_float(0x7f800000u);
    const int nwords = (W + 31) >> 5;
    const int64_t ngroups = (ntiles + 7) / 8;
    const int sub = lane >> 3, tl = lane & 7;  // box loads: lane = (pass-local box slot, tile of the group)
    unsigned long long npairs = 0;
    for (int64_t grp = (int64_t)blockIdx.x * 8 + (tid >> 5); grp < ngroups; grp += (int64_t)gridDim.x * 8) {
        // lane -> tile tl; the 4 lanes with the same tl share the tile's boxes_per_tile boxes (<= 16)
        const int64_t tile_l = grp * 8 + tl;
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
        if (lane < 8 && tile_l < ntiles) {
            npairs += any;
            // cost class of the tile for the work list: 0 = nothing listed, 1..4 = 1-2, 3-6, 7-14, >= 15 poses
            flags[tile_l] = (unsigned char)(any == 0u ? 0 : any <= 2u ? 1 : any <= 6u ? 2 : any <= 14u ? 3 : 4);
        }
    }
    npairs = __reduce_add_sync(kFull, (unsigned)npairs);
    if (lane == 0 && npairs) atomicAdd(listed_pairs, npairs);
}

// Work list of the tiles with a non-empty mask: heaviest cost class first, ascending tile index within a class, so
// that blocks taking entries i, i + grid, i + 2 grid, ... all get the same mix of heavy and light tiles (the order is
// a pure function of the masks: deterministic, no atomics; single block).
// ints[0] = count, ints[2] = 1 when the masks list more than `dense_above` (tile, pose) pairs: the cloud has no
// spatial coherence to exploit, the dense kernel does the call instead (it checks ints[2]) and the list is left empty.
__global__ void __launch_bounds__(1024) cov_worklist_kernel(const unsigned char* __restrict__ flags, int64_t ntiles,
                                                            int* __restrict__ worklist, int* __restrict__ ints,
                                                            unsigned long long dense_above) {
    __shared__ int warp_tot[32][4];
    __shared__ int class_start[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t chunk = (((ntiles + 1023) / 1024) + 15) & ~(int64_t)15;  // flags per thread, 16 per vector load
    const int64_t lo = (int64_t)tid * chunk < ntiles ? (int64_t)tid * chunk : ntiles;
    const int64_t hi = lo + chunk < ntiles ? lo + chunk : ntiles;
    int c[4] = {0, 0, 0, 0};  // tiles of class 4, 3, 2, 1 in this thread's chunk
    {
        int64_t t = lo;
        for (; t + 16 <= hi; t += 16) {
            const uint4 v = *reinterpret_cast<const uint4*>(flags + t);
            const unsigned words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (words[q] == 0u) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned f = (words[q] >> (8 * k)) & 0xffu;
                    c[0] += f == 4u; c[1] += f == 3u; c[2] += f == 2u; c[3] += f == 1u;
                }
            }
        }
        for (; t < hi; ++t) {
            const unsigned f = flags[t];
            c[0] += f == 4u; c[1] += f == 3u; c[2] += f == 2u; c[3] += f == 1u;
        }
    }
    int incl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        incl[k] = c[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl[k], o);
            if (lane >= o) incl[k] += v;
        }
        if (lane == 31) warp_tot[warp][k] = incl[k];
    }
    __syncthreads();
    if (warp == 0) {
        int tot[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = warp_tot[lane][k];
            int sc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(kFull, sc, o);
                if (lane >= o) sc += u;
            }
            warp_tot[lane][k] = sc - v;  // exclusive over warps
            tot[k] = __shfl_sync(kFull, sc, 31);
        }
        if (lane == 0) {
            const bool dense = *reinterpret_cast<const unsigned long long*>(ints + 4) > dense_above;
            class_start[0] = 0;
            class_start[1] = tot[0];
            class_start[2] = tot[0] + tot[1];
            class_start[3] = tot[0] + tot[1] + tot[2];
            ints[0] = dense ? 0 : tot[0] + tot[1] + tot[2] + tot[3];
            ints[2] = dense ? 1 : 0;
        }
    }
    __syncthreads();
    int pos[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pos[k] = class_start[k] + warp_tot[warp][k] + incl[k] - c[k];
    if (c[0] + c[1] + c[2] + c[3] == 0) return;
    int64_t t = lo;
    for (; t + 16 <= hi; t += 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(flags + t);
        const unsigned words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (words[q] == 0u) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned f = (words[q] >> (8 * k)) & 0xffu;
                const int tile = (int)(t + q * 4 + k);
                if (f == 4u) worklist[pos[0]++] = tile;
                else if (f == 3u) worklist[pos[1]++] = tile;
                else if (f == 2u) worklist[pos[2]++] = tile;
                else if (f == 1u) worklist[pos[3]++] = tile;
            }
        }
    }
    for (; t < hi; ++t) {
        const unsigned f = flags[t];
        if (f == 4u) worklist[pos[0]++] = (int)t;
        else if (f == 3u) worklist[pos[1]++] = (int)t;
        else if (f == 2u) worklist[pos[2]++] = (int)t;
        else if (f == 1u) worklist[pos[3]++] = (int)t;
    }
}

// =============================================== pass A, pruned ===============================================
template <int PPT, int MINB>
__global__ void __launch_bounds__(COV_THREADS, MINB)
cov_traj_minmax_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float4* __restrict__ table, int W, CovConst C,
                             unsigned* __restrict__ gmin, unsigned* __restrict__ gmax, const float4* __restrict__ boxes,
                             const unsigned* __restrict__ amask_g, int mask_stride, const int* __restrict__ worklist,
                             const int* __restrict__ count_ptr, int64_t ntiles, unsigned long long* __restrict__ stats) {
    constexpr int T = tile_points(PPT);
    constexpr int NB = tile_boxes(PPT);
    constexpr int SF = stage_floats(PPT);
    constexpr int WB = (32 * PPT >= kBoxPts) ? (32 * PPT / kBoxPts) : 1;  // boxes covering one warp's points
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    unsigned* smin = reinterpret_cast<unsigned*>(ptab + (size_t)W * COV_ROW_F4);
    unsigned* smax = smin + W;
    float* sqcap = reinterpret_cast<float*>(smax + W);
    float* stage = reinterpret_cast<float*>(smem4) + minmax_tiles_stage_offset_floats(W);  // [2][SF]
    __shared__ __align__(8) unsigned long long mbar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv_kd = 1.f / C.kd;
    unsigned n_box = 0, n_pre = 0, n_full = 0;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < W * COV_ROW_F4; i += COV_THREADS) ptab[i] = table[i];
    for (int w = tid; w < W; w += COV_THREADS) {
        smin[w] = 0x7f800000u;
        smax[w] = 0u;
        sqcap[w] = minmax_qcap(gmin[w], gmax[w], inv_kd);  // from the seed launch (any later value is tighter and valid)
    }
    __syncthreads();
    const int count = *count_ptr;
    const int64_t nfull = n / T;
    const int nwords = (W + 31) >> 5;
    unsigned uses0 = 0, uses1 = 0;
    if (tid == 0 && (int)blockIdx.x < count)
        stage_issue<PPT>(stage, &mbar[0], xyz, boxes, amask_g, mask_stride, (int64_t)worklist[blockIdx.x], nfull);
    int buf = 0;
    for (int i = blockIdx.x; i < count; i += gridDim.x, buf ^= 1) {
        const int64_t tile = worklist[i];
        if (tid == 0 && i + (int)gridDim.x < count)  // the other stage was last read before the previous barrier
            stage_issue<PPT>(stage + (buf ^ 1) * SF, &mbar[buf ^ 1], xyz, boxes, amask_g, mask_stride,
                             (int64_t)worklist[i + gridDim.x], nfull);
        if (buf == 0) mbar_wait(&mbar[0], uses0++ & 1u);
        else mbar_wait(&mbar[1], uses1++ & 1u);
        const float* st = stage + buf * SF;
        float px[PPT], py[PPT], pz[PPT];
        if (tile < nfull) {
            const float* src = st + (warp * (32 * PPT) + lane) * 3;
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                px[s] = src[s * 96];
                py[s] = src[s * 96 + 1];
                pz[s] = src[s * 96 + 2];
            }
        } else {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                int64_t j = tile * T + warp * (32 * PPT) + s * 32 + lane;
                j = j < n ? j : n - 1;  // a duplicate cannot change a min or a max
                px[s] = __ldg(xyz + j * 3);
                py[s] = __ldg(xyz + j * 3 + 1);
                pz[s] = __ldg(xyz + j * 3 + 2);
            }
        }
        // box of this warp's points = union of the precomputed boxes that cover them
        const float4* tb = reinterpret_cast<const float4*>(st + T * 3);
        const int b0 = (warp * 32 * PPT) / kBoxPts;
        float4 wlo = tb[2 * b0], whi = tb[2 * b0 + 1];
#pragma unroll
        for (int k = 1; k < WB; ++k) box_union(wlo, whi, tb[2 * (b0 + k)], tb[2 * (b0 + k) + 1]);
        const unsigned* am = reinterpret_cast<const unsigned*>(st + T * 3 + NB * 8);
        for (int c = 0; c < nwords; ++c) {
            unsigned word = am[c];
            while (word) {
                const int w = c * 32 + __ffs(word) - 1;
                word &= word - 1;
                const float4* row = ptab + (size_t)w * COV_ROW_F4;
                const float4 v3 = row[3];
                const float cap = sqcap[w];
                ++n_box;
                if (box_q2lb(wlo, whi, v3) > cap) continue;
                ++n_pre;
                float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
                for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
                if (!__any_sync(kFull, !(qmin > cap))) continue;
                ++n_full;
                const float4 v0 = row[0], v1 = row[1], v2 = row[2];
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
        __syncthreads();  // every warp is done with this stage before it is refilled
    }
    for (int w = tid; w < W; w += COV_THREADS) {
        if (smax[w] != 0u || smin[w] != 0x7f800000u) {
            atomicMin(gmin + w, smin[w]);
            atomicMax(gmax + w, smax[w]);
        }
    }
    if (lane == 0) {
        if (tid == 0 && blockIdx.x == 0) atomicAdd(stats + 2, (unsigned long long)ntiles * kWarps * W);
        atomicAdd(stats + 3, (unsigned long long)n_full);
        atomicAdd(stats + 5, (unsigned long long)n_pre);
        atomicAdd(stats + 7, (unsigned long long)n_box);
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

// Phase-1 body for U consecutive poses starting at w (U*PPT independent chains).  Dense kernel: the ballots go to
// the gate bit matrix.  TILES = pruned kernel: the conservative threshold lives in v5.w (v3.w holds qthr), nothing
// is stored, and the return value says whether this warp has a gated pair for the pose (U = 1).
template <int PPT, int U, bool AMIN, bool TILES>
__device__ __forceinline__ bool fused_pose_iter(int w, unsigned ptab, unsigned* __restrict__ bits,
                                                int warp, const float (&px)[PPT], const float (&py)[PPT],
                                                const float (&pz)[PPT], float (&L)[PPT], const CovConst& C,
                                                double* __restrict__ acc, int lane) {
    float m[U][PPT];
    float mmax[U];
    unsigned anyb = 0u;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float4 v0 = lds_row(ptab, w + u, 0), v1 = lds_row(ptab, w + u, 1), v2 = lds_row(ptab, w + u, 2),
                     v3 = lds_row(ptab, w + u, 3);
#pragma unroll
        for (int s = 0; s < PPT; ++s) m[u][s] = cov_vis<false>(px[s], py[s], pz[s], v0, v1, v2, v3, C, nullptr);
        mmax[u] = m[u][0];
#pragma unroll
        for (int s = 1; s + 1 < PPT; s += 2) mmax[u] = fmaxf(mmax[u], fmaxf(m[u][s], m[u][s + 1]));
        if ((PPT & 1) == 0) mmax[u] = fmaxf(mmax[u], m[u][PPT - 1]);
        // >= 0  <=>  some point of this lane may pass the gate (conservative threshold)
        mmax[u] -= TILES ? lds_row(ptab, w + u, 5).w : v3.w;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        unsigned bal[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) bal[s] = 0u;
        if (__any_sync(kFull, mmax[u] >= 0.f)) {  // warp-uniform; a few % of (warp, pose) iterations
            const float4 v4 = lds_row(ptab, w + u, 4);
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                const float d = __fsub_rn(m[u][s], v4.w);
                const bool act = d >= v4.x;  // exactly p >= 0.5
                bal[s] = __ballot_sync(kFull, act);
                if (act) {
                    const float p = __fmul_rn(d, v4.z);
                    const float qc = (p > C.hi) ? C.hi : p;  // upper clip; a NaN p (pose that sees nothing: 0/0) stays NaN, as torch.clip
                    L[s] += COV_LN2_F * cov_lg2(qc * cov_rcp(1.f - qc));
                    if (acc && d == v4.y) tie_accumulate(px[s], py[s], pz[s], w + u, C, acc, (w + u) * COV_ACC_STRIDE + 8);
                }
            }
        }
        if (TILES) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) anyb |= bal[s];
        } else if (lane == 0) {
            unsigned* brow = bit_row_group<PPT>(bits, w + u, warp);
            if (PPT == 4) *reinterpret_cast<uint4*>(brow) = make_uint4(bal[0], bal[1 % PPT], bal[2 % PPT], bal[3 % PPT]);
            else if (PPT == 2) *reinterpret_cast<uint2*>(brow) = make_uint2(bal[0], bal[1 % PPT]);
            else brow[0] = bal[0];
        }
        if (AMIN) {  // only compact clouds whose minimum did not underflow to 0 (block-uniform choice of the loop)
            const float a = lds_row(ptab, w + u, 4).w;
            if (a > 0.f) {
#pragma unroll
                for (int s = 0; s < PPT; ++s)
                    if (m[u][s] == a) tie_accumulate(px[s], py[s], pz[s], w + u, C, acc, (w + u) * COV_ACC_STRIDE + 15);
            }
        }
    }
    return anyb != 0u;
}

// Phase 2 of the dense kernel: the gated pairs of all W bit rows, each row split over 2^seg_log2 lanes.  A lane pops
// its set bits in ascending point order, recomputes m (bit-identical) and dm/dx, and accumulates in registers;
// segments are combined with xor-shuffles; one owner lane adds into accs.
template <int PPT>
__device__ __forceinline__ void fused_phase2(const float4* __restrict__ ptab, const unsigned* __restrict__ bits,
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

__device__ __forceinline__ void fused_block_epilogue(const float* accs, int W, float* __restrict__ partials,
                                                     double* __restrict__ sumr_partials, double sum_r, double* red,
                                                     int tid) {
    float* slab = partials + (size_t)blockIdx.x * W * 8;
    for (int i = tid; i < W * 8; i += COV_THREADS) slab[i] = accs[i];
    const double ws = cov_warp_sum(sum_r);
    if ((tid & 31) == 0) red[tid >> 5] = ws;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kWarps; ++i) t += red[i];
        sumr_partials[blockIdx.x] = t;
    }
}

// ---- dense ----
template <int PPT, bool HAS_UP, int U>
__global__ void __launch_bounds__(COV_THREADS, 2)
cov_traj_fused_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                      const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                      const float* __restrict__ minmax, const float* __restrict__ upstream,
                      const int32_t* __restrict__ out_index, float* __restrict__ rewards,
                      float* __restrict__ partials, double* __restrict__ sumr_partials, double* __restrict__ acc,
                      int seg_log2, const int* __restrict__ run_flag) {
    if (run_flag && *run_flag == 0) return;  // stand-by launch behind the pruned kernel (see cov_worklist_kernel)
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
    for (int w = tid; w < W; w += COV_THREADS) {
        float4* row = ptab + (size_t)w * COV_ROW_F4;
        cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, row);
        const float a = minmax[w];
        const float b = __fsub_rn(minmax[W + w], a);
        const float hb = 0.5f * b;
        row[3].w = (a + hb) * (1.f - 9.5367431640625e-07f);  // conservative gate threshold (exact test in the rare path)
        row[4] = make_float4(hb, b, __frcp_rn(b), a);
        if (a > 0.f) amin_pos = 1;  // benign race: every writer stores 1
    }
    for (int i = tid; i < W * 8; i += COV_THREADS) accs[i] = 0.f;
    __syncthreads();

    double sum_r = 0.0;
    const int64_t ntiles = (n + T - 1) / T;
    const bool check_amin = amin_pos != 0;
    const unsigned ptab_s = smem_u32(ptab);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
        const int lbase = warp * (32 * PPT) + lane;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const int64_t j = tile * T + lbase + s * 32;
            valid[s] = j < n;
            // a point past the end sits 3e18 m away: m = 0 exactly, never gated, never a tie
            px[s] = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;
            py[s] = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
            pz[s] = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
            pt[(lbase + s * 32) * 3] = px[s];
            pt[(lbase + s * 32) * 3 + 1] = py[s];
            pt[(lbase + s * 32) * 3 + 2] = pz[s];
            L[s] = 0.f;
        }
        {
            int w = 0;
            if (check_amin) {  // arg-min points carry gradient: every pair must be looked at
                for (; w < W; ++w) fused_pose_iter<PPT, 1, true, false>(w, ptab_s, bits, warp, px, py, pz, L, C, acc, lane);
            } else {
                for (; w + U <= W; w += U) fused_pose_iter<PPT, U, false, false>(w, ptab_s, bits, warp, px, py, pz, L, C, acc, lane);
                for (; w < W; ++w) fused_pose_iter<PPT, 1, false, false>(w, ptab_s, bits, warp, px, py, pz, L, C, acc, lane);
            }
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
        __syncthreads();
        fused_phase2<PPT>(ptab, bits, pt, Gs, accs, W, seg_log2, C, tid);
        __syncthreads();
    }
    fused_block_epilogue(accs, W, partials, sumr_partials, sum_r, red, tid);
}

// ---- pruned: persistent blocks over the work list of tiles with a non-empty pose mask ----
// Each warp owns 32*PPT consecutive points of the tile in registers, for both phases:
//   phase 1  walk the tile's pose mask; a listed pose is evaluated when its qthr-ball meets the warp's own box and
//            one of its points passes the same test; gated lanes add their log-odds; the warp notes (one bit per
//            pose) whether it had a gated pair.  Then r_j, G_j = r_j (1 - r_j) per point, rewards stored.
//   phase 2  (only when some warp of the block noted a gate)  the warp walks the mask again; for a noted pose it
//            re-evaluates m for its points (bit-identical), re-derives the gate, and accumulates the 8 weighted
//            sums of dm/dx in registers; a fixed xor-shuffle tree reduces them over the warp.  The 8 warps' partial
//            sums meet in a slot table indexed by the pose's rank in the mask and are added to the block
//            accumulators in warp order by one thread per (pose, component): fixed order, bitwise reproducible.
// sumr_partials receive sum_j (r_j - 1/2) over the listed tiles (the rest is 0.5 * n, added by the reduce kernel);
// `rewards` was pre-filled with 1/2, only other values are stored.
constexpr int kSlotPoses = 32;  // poses per round of the slot table

template <int PPT, bool HAS_UP>
__global__ void __launch_bounds__(COV_THREADS, 2)
cov_traj_fused_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float4* __restrict__ table, int W, CovConst C,
                            const float* __restrict__ upstream, const int32_t* __restrict__ out_index,
                            float* __restrict__ rewards, float* __restrict__ partials,
                            double* __restrict__ sumr_partials, double* __restrict__ acc,
                            const float4* __restrict__ boxes, const unsigned* __restrict__ amask_g, int mask_stride,
                            const int* __restrict__ worklist, const int* __restrict__ count_ptr, int64_t ntiles,
                            unsigned long long* __restrict__ stats) {
    constexpr int T = tile_points(PPT);
    constexpr int NB = tile_boxes(PPT);
    constexpr int SF = stage_floats(PPT) + T;  // + the tile's slice of the permutation
    constexpr int WB = (32 * PPT >= kBoxPts) ? (32 * PPT / kBoxPts) : 1;
    // shared memory: pose table | 2 tile stages (points, boxes, pose mask) | block accumulators
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    float* stage = reinterpret_cast<float*>(smem4) + fused_stage_offset_floats(W);
    float* accs = stage + 2 * SF;
    __shared__ double red[kWarps];
    __shared__ float slots[kSlotPoses][kWarps][8];   // per-(pose rank, warp) partial sums of one round
    __shared__ int slot_pose[kSlotPoses];
    __shared__ unsigned wgate[kWarps][kMaskWords];   // per warp: poses with a gated pair among its points (this tile)
    __shared__ __align__(8) unsigned long long mbar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < W * COV_ROW_F4; i += COV_THREADS) ptab[i] = table[i];
    for (int i = tid; i < W * 8; i += COV_THREADS) accs[i] = 0.f;
    const bool check_amin = count_ptr[1] != 0;  // flags[1]: some pose has min_j m > 0
    __syncthreads();

    double sum_r = 0.0;
    unsigned n_box = 0, n_pre = 0, n_full = 0;
    const int count = count_ptr[0];
    const int64_t nfull = n / T;
    const int nwords = (W + 31) >> 5;
    unsigned* wg = wgate[warp];
    const unsigned ptab_s = smem_u32(ptab);
    unsigned uses0 = 0, uses1 = 0;
    if (tid == 0 && (int)blockIdx.x < count)
        stage_issue<PPT>(stage, &mbar[0], xyz, boxes, amask_g, mask_stride, (int64_t)worklist[blockIdx.x], nfull, out_index);
    int buf = 0;
    for (int i = blockIdx.x; i < count; i += gridDim.x, buf ^= 1) {
        const int64_t tile = worklist[i];
        if (tid == 0 && i + (int)gridDim.x < count)  // the other stage was last read before the previous barrier
            stage_issue<PPT>(stage + (buf ^ 1) * SF, &mbar[buf ^ 1], xyz, boxes, amask_g, mask_stride,
                             (int64_t)worklist[i + gridDim.x], nfull, out_index);
        if (buf == 0) mbar_wait(&mbar[0], uses0++ & 1u);
        else mbar_wait(&mbar[1], uses1++ & 1u);
        const float* st = stage + buf * SF;
        // ------------------------------ phase 1 ------------------------------
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
        const int lbase = warp * (32 * PPT) + lane;
        if (tile < nfull) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                valid[s] = true;
                px[s] = st[(lbase + s * 32) * 3];
                py[s] = st[(lbase + s * 32) * 3 + 1];
                pz[s] = st[(lbase + s * 32) * 3 + 2];
                L[s] = 0.f;
            }
        } else {  // the ragged last tile: loaded by hand
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                const int64_t j = tile * T + lbase + s * 32;
                valid[s] = j < n;
                // a point past the end sits 3e18 m away: m = 0 exactly, never gated, never a tie
                px[s] = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;
                py[s] = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
                pz[s] = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
                L[s] = 0.f;
            }
        }
        const float4* tb = reinterpret_cast<const float4*>(st + T * 3);
        const int b0 = (warp * 32 * PPT) / kBoxPts;
        float4 wlo = tb[2 * b0], whi = tb[2 * b0 + 1];
#pragma unroll
        for (int k = 1; k < WB; ++k) box_union(wlo, whi, tb[2 * (b0 + k)], tb[2 * (b0 + k) + 1]);
        const unsigned* am = reinterpret_cast<const unsigned*>(st + T * 3 + NB * 8);
        const int32_t* sperm = reinterpret_cast<const int32_t*>(st + stage_floats(PPT));
        bool warp_gated = false;
        for (int c = 0; c < nwords; ++c) {
            unsigned word = am[c];
            unsigned gbits = 0u;
            while (word) {
                const int b = __ffs(word) - 1;
                const int w = c * 32 + b;
                word &= word - 1;
                const float4 v3 = ptab[(size_t)w * COV_ROW_F4 + 3];
                ++n_box;
                if (box_q2lb(wlo, whi, v3) > v3.w) continue;
                ++n_pre;
                float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
                for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
                if (!__any_sync(kFull, !(qmin > v3.w))) continue;
                ++n_full;
                const bool g = check_amin
                                   ? fused_pose_iter<PPT, 1, true, true>(w, ptab_s, nullptr, warp, px, py, pz, L, C, acc, lane)
                                   : fused_pose_iter<PPT, 1, false, true>(w, ptab_s, nullptr, warp, px, py, pz, L, C, acc, lane);
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
                const int64_t j = tile * T + lbase + s * 32;
                sum_r += (double)(r - 0.5f);
                const bool store = r != 0.5f;
                int64_t jo = j;
                if (out_index && (store || HAS_UP))
                    jo = tile < nfull ? (int64_t)sperm[lbase + s * 32] : (int64_t)out_index[j];  // staged with the tile
                if (store) rewards[jo] = r;
                if (HAS_UP) G[s] *= upstream[jo];
            }
        }
        // one barrier per tile: protects the stage that is refilled next and tells whether anybody has a gated pair
        if (!__syncthreads_or(warp_gated ? 1 : 0)) continue;
        // ------------------------------ phase 2 ------------------------------
        int listed = 0;
        for (int c = lane; c < nwords; c += 32) listed += __popc(am[c]);
        listed = __reduce_add_sync(kFull, listed);
        // rounds of kSlotPoses listed poses (one round unless the cloud is unordered); slot = rank within the round
        for (int round0 = 0; round0 < listed; round0 += kSlotPoses) {
            int rank = 0;
            for (int c = 0; c < nwords && rank < round0 + kSlotPoses; ++c) {
                unsigned word = am[c];
                const int cnt = __popc(word);
                if (rank + cnt <= round0) {  // the whole word belongs to an earlier round
                    rank += cnt;
                    continue;
                }
                const unsigned mine = wg[c];
                while (word) {
                    const int b = __ffs(word) - 1;
                    const int w = c * 32 + b;
                    word &= word - 1;
                    const int slot = rank - round0;
                    ++rank;
                    if (slot < 0) continue;
                    if (slot >= kSlotPoses) break;
                    float f0 = 0.f, f1 = 0.f, f2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, se = 0.f, sep = 0.f;
                    if ((mine >> b) & 1u) {
                        const float4* row = ptab + (size_t)w * COV_ROW_F4;
                        const float4 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3], v4 = row[4], v5 = row[5];
#pragma unroll
                        for (int s = 0; s < PPT; ++s) {
                            CovEval ev;
                            const float m = cov_vis<true>(px[s], py[s], pz[s], v0, v1, v2, v3, C, &ev);
                            const float d = __fsub_rn(m, v4.w);
                            const bool act = d >= v4.x;  // exactly p >= 0.5, as in phase 1
                            if (act) {
                                const float p = __fdiv_rn(d, v4.y);
                                if (p <= C.hi) {  // clamp backward gate (inclusive)
                                    float gx, gy, gz;
                                    cov_vis_grad(m, ev, v0, v1, v2, C, gx, gy, gz);
                                    const float yx = px[s] - v5.x, yy = py[s] - v5.y, yz = pz[s] - v5.z;
                                    const float e = G[s] * cov_rcp(p * (1.f - p));
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
                        // 8 sums over 32 lanes in 9 shuffles: halve the number of values a lane carries at offsets 16,
                        // 8, 4 (a lane keeps the half its lane bit selects and hands the other half over), then two
                        // plain steps; lane l ends with component (l>>2)&7 ... in bit order (16,8,4) -> (4,2,1)
                        float v[8] = {f0, f1, f2, t0, t1, t2, se, sep};
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
                        f0 = v[0];
                    }
                    if ((lane & 3) == 0) {
                        const int comp = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        slots[slot][warp][comp] = f0;  // 0 when this warp has no gated pair for the pose
                        if (tid == 0) slot_pose[slot] = w;
                    }
                }
            }
            __syncthreads();  // this round's partial sums are in the slot table (and nobody reads the stage any more)
            {
                const int nslots = (listed - round0) < kSlotPoses ? (listed - round0) : kSlotPoses;
                const int k = tid >> 3, comp = tid & 7;
                if (k < nslots) {
                    float sacc = 0.f;
#pragma unroll
                    for (int g = 0; g < kWarps; ++g) sacc += slots[k][g][comp];
                    accs[(size_t)slot_pose[k] * 8 + comp] += sacc;
                }
            }
            if (round0 + kSlotPoses < listed) __syncthreads();  // the slot table is rewritten by the next round
        }
    }
    __syncthreads();
    fused_block_epilogue(accs, W, partials, sumr_partials, sum_r, red, tid);
    if (lane == 0) {
        if (tid == 0 && blockIdx.x == 0) atomicAdd(stats + 0, (unsigned long long)ntiles * kWarps * W);
        atomicAdd(stats + 1, (unsigned long long)n_full);
        atomicAdd(stats + 4, (unsigned long long)n_pre);
        atomicAdd(stats + 6, (unsigned long long)n_box);
    }
}

// ---- candidate sweep, pruned (BASELINE config 5): forward only, per-trajectory sum_j r_j ----
// Poses are trajectory-major (pose = traj * per_traj + i), so walking a tile's pose mask in ascending order visits the
// trajectories one after the other: a warp keeps the log-odds sums of its points for the current trajectory, and when
// the trajectory changes it adds sum_j (sigmoid(L_j) - 1/2) to that trajectory's total (fp64 shared atomic).
// Trajectories no pose of which is listed for a tile contribute exactly 1/2 per point: 0.5 * n is added once.
template <int PPT>
__global__ void __launch_bounds__(COV_THREADS, 2)
cov_sweep_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float4* __restrict__ table, int W, int per_traj,
                       int n_traj, CovConst C, const float4* __restrict__ boxes, const unsigned* __restrict__ amask_g,
                       int mask_stride, const int* __restrict__ worklist, const int* __restrict__ count_ptr,
                       double* __restrict__ sum_out) {
    constexpr int T = tile_points(PPT);
    constexpr int NB = tile_boxes(PPT);
    constexpr int SF = stage_floats(PPT);
    constexpr int WB = (32 * PPT >= kBoxPts) ? (32 * PPT / kBoxPts) : 1;
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    float* stage = reinterpret_cast<float*>(smem4) + fused_stage_offset_floats(W);
    double* ssum = reinterpret_cast<double*>(stage + 2 * SF);
    __shared__ __align__(8) unsigned long long mbar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < W * COV_ROW_F4; i += COV_THREADS) ptab[i] = table[i];
    for (int t = tid; t < n_traj; t += COV_THREADS) ssum[t] = 0.0;
    __syncthreads();
    const int count = count_ptr[0];
    const int64_t nfull = n / T;
    const int nwords = (W + 31) >> 5;
    const unsigned ptab_s = smem_u32(ptab);
    unsigned uses0 = 0, uses1 = 0;
    if (tid == 0 && (int)blockIdx.x < count)
        stage_issue<PPT>(stage, &mbar[0], xyz, boxes, amask_g, mask_stride, (int64_t)worklist[blockIdx.x], nfull);
    int buf = 0;
    for (int i = blockIdx.x; i < count; i += gridDim.x, buf ^= 1) {
        const int64_t tile = worklist[i];
        if (tid == 0 && i + (int)gridDim.x < count)
            stage_issue<PPT>(stage + (buf ^ 1) * SF, &mbar[buf ^ 1], xyz, boxes, amask_g, mask_stride,
                             (int64_t)worklist[i + gridDim.x], nfull);
        if (buf == 0) mbar_wait(&mbar[0], uses0++ & 1u);
        else mbar_wait(&mbar[1], uses1++ & 1u);
        const float* st = stage + buf * SF;
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
        const int lbase = warp * (32 * PPT) + lane;
        if (tile < nfull) {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                valid[s] = true;
                px[s] = st[(lbase + s * 32) * 3];
                py[s] = st[(lbase + s * 32) * 3 + 1];
                pz[s] = st[(lbase + s * 32) * 3 + 2];
            }
        } else {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                const int64_t j = tile * T + lbase + s * 32;
                valid[s] = j < n;
                px[s] = valid[s] ? __ldg(xyz + j * 3) : 3.0e18f;  // past the end: m = 0 exactly, never gated
                py[s] = valid[s] ? __ldg(xyz + j * 3 + 1) : 3.0e18f;
                pz[s] = valid[s] ? __ldg(xyz + j * 3 + 2) : 3.0e18f;
            }
        }
        const float4* tb = reinterpret_cast<const float4*>(st + T * 3);
        const int b0 = (warp * 32 * PPT) / kBoxPts;
        float4 wlo = tb[2 * b0], whi = tb[2 * b0 + 1];
#pragma unroll
        for (int k = 1; k < WB; ++k) box_union(wlo, whi, tb[2 * (b0 + k)], tb[2 * (b0 + k) + 1]);
        const unsigned* am = reinterpret_cast<const unsigned*>(st + T * 3 + NB * 8);
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
                const float4 v3 = ptab[(size_t)w * COV_ROW_F4 + 3];
                if (box_q2lb(wlo, whi, v3) > v3.w) continue;
                float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
                for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
                if (!__any_sync(kFull, !(qmin > v3.w))) continue;
                const int t = w / per_traj;
                if (t != cur) {
                    flush();
                    cur = t;
                    touched = false;
#pragma unroll
                    for (int s = 0; s < PPT; ++s) L[s] = 0.f;
                }
                touched |= fused_pose_iter<PPT, 1, false, true>(w, ptab_s, nullptr, warp, px, py, pz, L, C, nullptr, lane);
            }
        }
        flush();
        __syncthreads();  // every warp is done with this stage before it is refilled
    }
    __syncthreads();
    for (int t = tid; t < n_traj; t += COV_THREADS) {
        double v = ssum[t];
        if (blockIdx.x == 0) v += 0.5 * (double)n;
        if (v != 0.0) atomicAdd(sum_out + t, v);
    }
}

// acc[w][0..7] = sum over blocks of the fp32 slabs (fp64, fixed order); acc[W*STRIDE] = base + sum of the blocks'
// reward sums (base = 0.5 * n for the pruned kernel, whose blocks sum r - 1/2 over the listed tiles only).
__global__ void cov_traj_reduce_kernel(const float* __restrict__ partials, const double* __restrict__ sumr_partials,
                                       int nblocks, int nblocks_dense, int W, double base,
                                       const int* __restrict__ dense_flag, double* __restrict__ acc) {
    if (dense_flag && *dense_flag) {  // the dense kernel did the call: its blocks wrote the slabs and summed r itself
        base = 0.0;
        nblocks = nblocks_dense;
    }
    // 8 lanes per output: lane q adds blocks q, q+8, ... in order, then a fixed xor tree joins the 8 partial sums
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid >> 3, q = gid & 7;
    double s = 0.0;
    if (i < W * 8)
        for (int b = q; b < nblocks; b += 8) s += (double)partials[(size_t)b * W * 8 + i];
    else if (i == W * 8)
        for (int b = q; b < nblocks; b += 8) s += sumr_partials[b];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (q == 0) {
        if (i < W * 8) acc[(size_t)(i >> 3) * COV_ACC_STRIDE + (i & 7)] = s;
        else if (i == W * 8) acc[(size_t)W * COV_ACC_STRIDE] = base + s;
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
constexpr size_t kSmemCap = (227 - 11) * 1024;  // opt-in shared memory per block on sm_100, minus static use (slot table)
constexpr int64_t kSeedSamples = 16384;          // pruned pass A: size of the strided sample that seeds the bounds
constexpr int64_t kDenseBelow = 4 * kSeedSamples;  // clouds this small go straight to the dense pass A

int pick_ppt(int64_t n, int W, bool fused, bool prune) {
    const int sms = cov_sm_count_cached();
    const int cand[3] = {4, 2, 1};
    for (int i = 0; i < 3; ++i) {
        const int ppt = cand[i];
        const size_t sm = fused ? fused_smem_bytes(W, ppt, prune)
                                : (prune ? minmax_tiles_smem_bytes(W, ppt) : minmax_smem_bytes(W));
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

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
int64_t boxes_padded(int64_t n) { return ((n + 2047) / 2048) * (2048 / kBoxPts); }

// Workspace of both passes (every region 256-byte aligned):
//   [reward-sum partials][accumulator slabs][pose table][ints: count, amin flag][work list][tile flags][tile masks][boxes]
struct TrajWorkspace {
    double* sumr;
    float* partials;
    float4* table;
    int* ints;
    int* worklist;
    unsigned char* flags;
    unsigned* amask;
    float4* boxes;
    size_t bytes;
};
TrajWorkspace carve_workspace(void* ws, int64_t n, int W) {
    const int64_t nt = (n + 255) / 256;  // tiles at the smallest tile size
    char* p = reinterpret_cast<char*>(ws);
    size_t off = 0;
    TrajWorkspace t;
    auto take = [&](size_t bytes) { char* q = p + off; off += align256(bytes); return q; };
    t.sumr = reinterpret_cast<double*>(take(COV_MAX_GRID * sizeof(double)));
    t.partials = reinterpret_cast<float*>(take((size_t)COV_MAX_GRID * W * 8 * sizeof(float)));
    t.table = reinterpret_cast<float4*>(take((size_t)W * COV_ROW_F4 * sizeof(float4)));
    t.ints = reinterpret_cast<int*>(take(256));
    t.worklist = reinterpret_cast<int*>(take((size_t)nt * sizeof(int)));
    t.flags = reinterpret_cast<unsigned char*>(take((size_t)nt));
    t.amask = reinterpret_cast<unsigned*>(take((size_t)nt * mask_stride_words(W) * sizeof(unsigned)));
    t.boxes = reinterpret_cast<float4*>(take((size_t)boxes_padded(n) * 2 * sizeof(float4)));
    t.bytes = off;
    return t;
}

int check_traj_args(const char* who, const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                    const float* K, const cov_camera* cam, const void* ws, size_t ws_bytes) {
    if (!xyz || n <= 0 || !poses || !quats || W <= 0 || !K || !cam) {
        cov_set_error("%s: null pointer, empty cloud or no poses (n=%lld, W=%d)", who, (long long)n, W);
        return COV_ERR_ARG;
    }
    if (W > cov_traj_max_poses()) {
        cov_set_error("%s: %d poses exceed the shared-memory pose table (max %d)", who, W, cov_traj_max_poses());
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

// dense_frac: hand the call to the dense kernel when more than this fraction of all (tile, pose) pairs is listed
// (2.0 = never)
void launch_cull(const float4* boxes, int ppt, int64_t ntiles, const TrajWorkspace& t, int W, const unsigned* gmin,
                 const unsigned* gmax, float inv_kd, double dense_frac, cudaStream_t s) {
    const int grid = (int)std::min<int64_t>((ntiles + 63) / 64, (int64_t)cov_sm_count_cached() * 8);
    cov_cull_kernel<<<grid, 256, (size_t)W * sizeof(float4), s>>>(boxes, tile_boxes(ppt), ntiles, t.table, W, gmin, gmax,
                                                                 inv_kd, t.amask, mask_stride_words(W), t.flags,
                                                                 reinterpret_cast<unsigned long long*>(t.ints + 4));
    cov_worklist_kernel<<<1, 1024, 0, s>>>(t.flags, ntiles, t.worklist, t.ints,
                                           (unsigned long long)(dense_frac * (double)ntiles * (double)W));
}

}  // namespace

extern "C" int cov_traj_max_poses(void) {
    int w = 1;
    while (w < 32 * kMaskWords && fused_smem_bytes(w + 1, 1, false) <= kSmemCap && fused_smem_bytes(w + 1, 1, true) <= kSmemCap &&
           minmax_tiles_smem_bytes(w + 1, 1) <= kSmemCap && (size_t)(w + 1) * sizeof(float4) <= 48 * 1024)
        ++w;
    return w;
}

extern "C" size_t cov_traj_workspace_bytes(int64_t n, int n_poses) {
    if (n < 1) n = 1;
    if (n_poses < 1) n_poses = 1;
    return carve_workspace(nullptr, n, n_poses).bytes;
}

extern "C" int64_t cov_tile_boxes_count(int64_t n) { return n > 0 ? boxes_padded(n) : 0; }

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
                               const float* K, const cov_camera* cam, const float* boxes_dev, float* minmax, void* ws,
                               size_t ws_bytes, void* stream) {
    int rc = check_traj_args("cov_traj_minmax", xyz, n, poses, quats, W, K, cam, ws, ws_bytes);
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
    unsigned* gmin = reinterpret_cast<unsigned*>(minmax);
    unsigned* gmax = gmin + W;
    cov_minmax_init_kernel<<<(W + 255) / 256, 256, 0, s>>>(gmin, gmax, W);
    const bool prune = cov_pruning_enabled() != 0 && n >= kDenseBelow;
    int ppt = pick_ppt(n, W, false, prune);
    if (ppt == 4 && (n + tile_points(8) - 1) / tile_points(8) >= 4 * (int64_t)cov_sm_count_cached() &&
        (!prune || minmax_tiles_smem_bytes(W, 8) <= kSmemCap))
        ppt = 8;
    if (ppt == 0) {
        cov_set_error("cov_traj_minmax: %d poses do not fit in shared memory", W);
        return COV_ERR_UNSUPPORTED;
    }
    const size_t smem = minmax_smem_bytes(W);
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
#define LAUNCH_DENSE(P, B, U, NPTS, STRIDE)                                                                           \
    {                                                                                                                 \
        const int64_t nt_ = ((NPTS) + tile_points(P) - 1) / tile_points(P);                                           \
        const int grid = grid_for(cov_traj_minmax_kernel<P, B, U>, smem, nt_);                                        \
        cov_traj_minmax_kernel<P, B, U><<<grid, COV_THREADS, smem, s>>>(xyz, NPTS, STRIDE, poses, quats, W, K, C, gmin, \
                                                                        gmax);                                        \
    }
    if (!prune) {
        if (ppt == 8) LAUNCH_DENSE(8, 2, 2, n, 1)
        else if (ppt == 4) LAUNCH_DENSE(4, 2, 1, n, 1)
        else if (ppt == 2) LAUNCH_DENSE(2, 2, 1, n, 1)
        else LAUNCH_DENSE(1, 2, 1, n, 1)
        return cov_check_launch("cov_traj_minmax");
    }
#undef LAUNCH_DENSE
    // pruned: seed the bounds on a strided sample, cull tiles against them, evaluate the listed tiles
    const TrajWorkspace t = carve_workspace(ws, n, W);
    const float4* boxes = boxes_for_call(xyz, n, boxes_dev, t, s);
    cudaMemsetAsync(t.ints, 0, 256, s);
    cov_pose_table_kernel<<<(W + 127) / 128, 128, 0, s>>>(poses, quats, W, K, C, nullptr, t.table, t.ints);
    {
        const int64_t stride = (n + kSeedSamples - 1) / kSeedSamples;
        const int64_t nsamples = (n + stride - 1) / stride;
        const dim3 sgrid((unsigned)((nsamples + COV_THREADS - 1) / COV_THREADS), (unsigned)((W + 31) / 32));
        cov_traj_seed_kernel<<<sgrid, COV_THREADS, 0, s>>>(xyz, nsamples, stride, t.table, W, C, gmin, gmax);
    }
    launch_cull(boxes, ppt, ntiles, t, W, gmin, gmax, 1.f / C.kd, 2.0, s);
    unsigned long long* stats = cov_stats_device_ptr();
#define LAUNCH_TILES(P)                                                                                             \
    {                                                                                                               \
        const size_t smem_t = minmax_tiles_smem_bytes(W, P);                                                        \
        const int grid = grid_for(cov_traj_minmax_tiles_kernel<P, 2>, smem_t, ntiles);                              \
        cov_traj_minmax_tiles_kernel<P, 2><<<grid, COV_THREADS, smem_t, s>>>(                                       \
            xyz, n, t.table, W, C, gmin, gmax, boxes, t.amask, mask_stride_words(W), t.worklist, t.ints, ntiles, stats); \
    }
    if (ppt == 8) LAUNCH_TILES(8) else if (ppt == 4) LAUNCH_TILES(4) else if (ppt == 2) LAUNCH_TILES(2) else LAUNCH_TILES(1)
#undef LAUNCH_TILES
    return cov_check_launch("cov_traj_minmax");
}

extern "C" int cov_traj_fused(const float* xyz, int64_t n, const float* poses, const float* quats, int W,
                              const float* K, const cov_camera* cam, const float* boxes_dev, const float* minmax,
                              const float* upstream, const int32_t* reward_index, float* rewards, double* acc, void* ws,
                              size_t ws_bytes, void* stream) {
    int rc = check_traj_args("cov_traj_fused", xyz, n, poses, quats, W, K, cam, ws, ws_bytes);
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
    const bool prune = cov_pruning_enabled() != 0;
    const TrajWorkspace t = carve_workspace(ws, n, W);
    cudaMemsetAsync(acc, 0, ((size_t)W * COV_ACC_STRIDE + 1) * sizeof(double), s);

    // the dense kernel: the whole call when pruning is off, a stand-by launch behind the pruned kernel otherwise
    // (it runs only if the cull found nothing to prune: run_flag = ints[2])
    const int ppt_d = pick_ppt(n, W, true, false);
    if (ppt_d == 0) {
        cov_set_error("cov_traj_fused: %d poses do not fit in shared memory", W);
        return COV_ERR_UNSUPPORTED;
    }
    auto launch_dense = [&](const int* run_flag) -> int {
        const size_t smem = fused_smem_bytes(W, ppt_d, false);
        const int64_t ntiles = (n + tile_points(ppt_d) - 1) / tile_points(ppt_d);
        // phase-2 parallelism: split each pose row over 2^seg_log2 lanes until there are >= 2 tasks per thread
        int seg_log2 = 0;
        while ((W << seg_log2) < 2 * COV_THREADS && (2 << seg_log2) <= bit_words(ppt_d) && seg_log2 < 5) ++seg_log2;
        int grid = 1;
#define LAUNCH_F(P, UP)                                                                                           \
    {                                                                                                             \
        grid = grid_for(cov_traj_fused_kernel<P, UP, 2>, smem, ntiles);                                           \
        cov_traj_fused_kernel<P, UP, 2><<<grid, COV_THREADS, smem, s>>>(xyz, n, poses, quats, W, K, C, minmax,    \
                                                                        upstream, reward_index, rewards, t.partials, \
                                                                        t.sumr, acc, seg_log2, run_flag);         \
    }
        if (upstream) {
            if (ppt_d == 4) LAUNCH_F(4, true) else if (ppt_d == 2) LAUNCH_F(2, true) else LAUNCH_F(1, true)
        } else {
            if (ppt_d == 4) LAUNCH_F(4, false) else if (ppt_d == 2) LAUNCH_F(2, false) else LAUNCH_F(1, false)
        }
#undef LAUNCH_F
        return grid;
    };
    if (!prune) {
        const int grid = launch_dense(nullptr);
        cov_traj_reduce_kernel<<<((W * 8 + 1) * 8 + 255) / 256, 256, 0, s>>>(t.partials, t.sumr, grid, grid, W, 0.0, nullptr, acc);
        return cov_check_launch("cov_traj_fused");
    }
    // pruned: rewards start at 1/2; cull tiles against the gate thresholds; evaluate the listed tiles
    const int ppt = pick_ppt(n, W, true, true);
    if (ppt == 0) {
        cov_set_error("cov_traj_fused: %d poses do not fit in shared memory", W);
        return COV_ERR_UNSUPPORTED;
    }
    const size_t smem = fused_smem_bytes(W, ppt, true);
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
    const float4* boxes = boxes_for_call(xyz, n, boxes_dev, t, s);
    cudaMemsetAsync(t.ints, 0, 256, s);
    cov_pose_table_kernel<<<(W + 127) / 128, 128, 0, s>>>(poses, quats, W, K, C, minmax, t.table, t.ints);
    {
        const int fgrid = (int)std::min<int64_t>((n + 1023) / 1024, (int64_t)cov_sm_count_cached() * 16);
        cov_fill_kernel<<<fgrid, 256, 0, s>>>(rewards, n, 0.5f);
    }
    launch_cull(boxes, ppt, ntiles, t, W, nullptr, nullptr, 0.f, 0.25, s);
    unsigned long long* stats = cov_stats_device_ptr();
    int grid = 1;
#define LAUNCH_FT(P, UP)                                                                                          \
    {                                                                                                             \
        grid = grid_for(cov_traj_fused_tiles_kernel<P, UP>, smem, ntiles);                                        \
        cov_traj_fused_tiles_kernel<P, UP><<<grid, COV_THREADS, smem, s>>>(                                       \
            xyz, n, t.table, W, C, upstream, reward_index, rewards, t.partials, t.sumr, acc, boxes, t.amask,      \
            mask_stride_words(W), t.worklist, t.ints, ntiles, stats);                                             \
    }
    if (upstream) {
        if (ppt == 4) LAUNCH_FT(4, true) else if (ppt == 2) LAUNCH_FT(2, true) else LAUNCH_FT(1, true)
    } else {
        if (ppt == 4) LAUNCH_FT(4, false) else if (ppt == 2) LAUNCH_FT(2, false) else LAUNCH_FT(1, false)
    }
#undef LAUNCH_FT
    const int grid_dense = launch_dense(t.ints + 2);
    cov_traj_reduce_kernel<<<((W * 8 + 1) * 8 + 255) / 256, 256, 0, s>>>(t.partials, t.sumr, grid, grid_dense, W, 0.5 * (double)n,
                                                               t.ints + 2, acc);
    return cov_check_launch("cov_traj_fused");
}

extern "C" int cov_sweep_rewards(const float* xyz, int64_t n, const float* poses, const float* quats, int n_traj,
                                 int per_traj, const float* K, const cov_camera* cam, const float* boxes_dev,
                                 const float* minmax, double* sum_rewards, void* ws, size_t ws_bytes, void* stream) {
    if (!xyz || n <= 0 || !poses || !quats || n_traj <= 0 || per_traj <= 0 || !K || !cam || !minmax || !sum_rewards) {
        cov_set_error("cov_sweep_rewards: bad argument");
        return COV_ERR_ARG;
    }
    if (!cov_pruning_enabled())
        return cov_sweep_rewards_dense(xyz, n, poses, quats, n_traj, per_traj, K, cam, minmax, sum_rewards, stream);
    // trajectories per launch: pose table + two stages + per-trajectory sums within ~100 KB (two blocks per SM)
    constexpr int PPT = 4;
    auto smem_for = [&](int nt) {
        return (size_t)fused_stage_offset_floats(nt * per_traj) * 4 + 2 * (size_t)stage_floats(PPT) * 4 + (size_t)nt * sizeof(double);
    };
    if (smem_for(1) > kSmemCap || per_traj > 32 * kMaskWords) {
        cov_set_error("cov_sweep_rewards: %d poses per trajectory do not fit in shared memory", per_traj);
        return COV_ERR_UNSUPPORTED;
    }
    int chunk = 1;
    while (chunk < n_traj && smem_for(chunk + 1) <= 100 * 1024 && (chunk + 1) * per_traj <= 32 * kMaskWords) ++chunk;
    const int Wc = chunk * per_traj;
    if (!ws || ws_bytes < cov_traj_workspace_bytes(n, Wc) || (((uintptr_t)ws) & 255) || (((uintptr_t)xyz) & 15)) {
        cov_set_error("cov_sweep_rewards: workspace missing, misaligned or smaller than cov_traj_workspace_bytes(n, %d) = %zu",
                      Wc, cov_traj_workspace_bytes(n, Wc));
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const CovConst C = cov_make_const(cam);
    const TrajWorkspace t = carve_workspace(ws, n, Wc);
    const float4* boxes = boxes_for_call(xyz, n, boxes_dev, t, s);
    const int W = n_traj * per_traj;
    const int64_t ntiles = (n + tile_points(PPT) - 1) / tile_points(PPT);
    float* mm = reinterpret_cast<float*>(t.partials);  // the chunk's minima and maxima, contiguous (2 * Wc floats)
    for (int t0 = 0; t0 < n_traj; t0 += chunk) {
        const int nt = (n_traj - t0 < chunk) ? n_traj - t0 : chunk;
        const int w0 = t0 * per_traj, Wn = nt * per_traj;
        cudaMemcpyAsync(mm, minmax + w0, (size_t)Wn * sizeof(float), cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(mm + Wn, minmax + W + w0, (size_t)Wn * sizeof(float), cudaMemcpyDeviceToDevice, s);
        cudaMemsetAsync(t.ints, 0, 256, s);
        cov_pose_table_kernel<<<(Wn + 127) / 128, 128, 0, s>>>(poses + 3 * (size_t)w0, quats + 4 * (size_t)w0, Wn, K, C, mm,
                                                              t.table, t.ints);
        launch_cull(boxes, PPT, ntiles, t, Wn, nullptr, nullptr, 0.f, 2.0, s);
        const size_t smem = smem_for(nt);
        const int grid = grid_for(cov_sweep_tiles_kernel<PPT>, smem, ntiles);
        cov_sweep_tiles_kernel<PPT><<<grid, COV_THREADS, smem, s>>>(xyz, n, t.table, Wn, per_traj, nt, C, boxes, t.amask,
                                                                   mask_stride_words(Wn), t.worklist, t.ints,
                                                                   sum_rewards + t0);
    }
    return cov_check_launch("cov_sweep_rewards");
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
