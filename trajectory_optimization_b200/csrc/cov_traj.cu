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
#include <algorithm>
#include <cstdlib>

#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kWarps = COV_THREADS / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaskWords = 64;  // active-pose bit mask: covers cov_traj_max_poses() <= 2048

__host__ __device__ constexpr int tile_points(int ppt) { return COV_THREADS * ppt; }
__host__ __device__ constexpr int bit_words(int ppt) { return kWarps * ppt; }         // ballot words per pose
__host__ __device__ constexpr int bit_stride(int ppt) { return bit_words(ppt) + 4; }  // rows stay 16-byte aligned

size_t minmax_smem_bytes(int W) { return (size_t)W * (COV_ROW_F4 * sizeof(float4) + 3 * sizeof(unsigned)); }
// pruned pass A: pose table + block min/max/cap, then the 128-byte aligned double buffer of raw tiles [2][T*3]
__host__ __device__ inline int minmax_tiles_raw_offset_floats(int W) {
    return (int)((((size_t)W * (COV_ROW_F4 * 16 + 12) + 127) & ~(size_t)127) / 4);
}
size_t minmax_tiles_smem_bytes(int W, int ppt) {
    return (size_t)minmax_tiles_raw_offset_floats(W) * 4 + 2 * (size_t)tile_points(ppt) * 12;
}
// fused pass: pose table, then (128-byte aligned) the tile buffer(s), G_j, gate bits, accumulators, active-pose list
__host__ __device__ inline int fused_raw_offset_floats(int W) {
    return (int)((((size_t)W * COV_ROW_F4 * 16 + 127) & ~(size_t)127) / 4);
}
size_t fused_smem_bytes(int W, int ppt, bool prune) {
    const size_t rs = prune ? bit_words(ppt) : bit_stride(ppt);
    return (size_t)fused_raw_offset_floats(W) * 4 + (size_t)(prune ? 2 : 1) * tile_points(ppt) * 12 +
           (size_t)tile_points(ppt) * 4 + (size_t)W * rs * sizeof(unsigned) + (size_t)W * 8 * sizeof(float) +
           (prune ? (((size_t)W * sizeof(unsigned short) + 15) & ~(size_t)15) : 0);
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

// ---- tile-level pruning helpers -------------------------------------------------------------------------------
// Order-preserving float <-> uint map (so one integer REDUX gives a float min or max of either sign).
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

// Axis-aligned box of the 32*PPT points a warp holds (identical in every lane).  Points flagged invalid are ignored.
template <int PPT>
__device__ __forceinline__ void warp_box(const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                         const bool (&valid)[PPT], float3& lo, float3& hi) {
    const float inf = __uint_as_float(0x7f800000u);
    float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
        if (valid[s]) {
            lx = fminf(lx, px[s]); ly = fminf(ly, py[s]); lz = fminf(lz, pz[s]);
            hx = fmaxf(hx, px[s]); hy = fmaxf(hy, py[s]); hz = fmaxf(hz, pz[s]);
        }
    }
    lo.x = ord2f(__reduce_min_sync(kFull, f2ord(lx)));
    lo.y = ord2f(__reduce_min_sync(kFull, f2ord(ly)));
    lo.z = ord2f(__reduce_min_sync(kFull, f2ord(lz)));
    hi.x = ord2f(__reduce_max_sync(kFull, f2ord(hx)));
    hi.y = ord2f(__reduce_max_sync(kFull, f2ord(hy)));
    hi.z = ord2f(__reduce_max_sync(kFull, f2ord(hz)));
}

// Lower bound of cov_q2(x, y, z, v3) over every point of the box [lo, hi].  Rounding is monotone, so with the same
// operation sequence as cov_q2 (fsub, then fmul/fma/fma) the bound holds for the COMPUTED q2 of each point, exactly:
// |fl(x - td)| >= max(fl(lo - td), fl(td - hi), 0) for lo <= x <= hi.  An empty box (lo = +inf) gives +inf.
__device__ __forceinline__ float box_q2lb(const float3& lo, const float3& hi, const float4& v3) {
    const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, v3.x), __fsub_rn(v3.x, hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, v3.y), __fsub_rn(v3.y, hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, v3.z), __fsub_rn(v3.z, hi.z)), 0.f);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Union of the kWarps warp boxes a block left in shared memory (wbox[warp] = lo.xyz, hi.xyz).
__device__ __forceinline__ void block_box(const float (*wbox)[8], float3& lo, float3& hi) {
    lo = make_float3(wbox[0][0], wbox[0][1], wbox[0][2]);
    hi = make_float3(wbox[0][4], wbox[0][5], wbox[0][6]);
#pragma unroll
    for (int i = 1; i < kWarps; ++i) {
        lo.x = fminf(lo.x, wbox[i][0]); lo.y = fminf(lo.y, wbox[i][1]); lo.z = fminf(lo.z, wbox[i][2]);
        hi.x = fmaxf(hi.x, wbox[i][4]); hi.y = fmaxf(hi.y, wbox[i][5]); hi.z = fmaxf(hi.z, wbox[i][6]);
    }
}

// Ask the L2 to fetch the points of a tile this block will read next (one thread issues it).
__device__ __forceinline__ void prefetch_tile_l2(const float* __restrict__ xyz, int64_t n, int64_t tile, int tile_pts) {
    const int64_t first = tile * tile_pts;
    if (first >= n) return;
    int64_t pts = n - first;
    if (pts > tile_pts) pts = tile_pts;
    const unsigned bytes = (unsigned)(pts * 12) & ~15u;
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xyz + first * 3), "r"(bytes) : "memory");
}

// ---- double-buffered TMA stream of whole tiles (pruned kernels) ------------------------------------------------
// One elected thread arms an mbarrier with the byte count and issues one cp.async.bulk (global -> shared) per
// tile; everybody waits on the barrier's phase parity.  Tile k+1 is in flight while tile k is processed.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
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

// Pass A with tile-level pruning (the product default).  Works on any point order and pays off when consecutive
// points are spatially close (cov_spatial_sort): per tile the block builds the bounding boxes of its warps' points,
// tests every pose against the block box (thread-parallel, 1 pose per thread and round) and keeps a bit mask of
// the poses that can still matter; each warp then walks that mask, re-tests against its own box, runs the per-lane
// distance pre-filter and only then the full evaluation.
// A pose can be skipped for a set of points when (i) its global minimum is already known to be exactly 0 (skipped
// points have m >= 0, so they cannot lower it) and (ii) every skipped point has m < the largest m seen so far:
//   m <= 2^-(kd q2) (1 + 1.3e-5),  q2 >= box bound > qcap = (1e-4 - log2(max_seen)) / kd   =>   m < max_seen.
// "Seen so far" is global: blocks publish their per-pose min/max to gmin/gmax after every tile and read them
// (relaxed loads, L2) before the next one; stale values are only looser.  Tiles are visited in a golden-ratio
// stride so the first rounds sample the whole cloud and the bounds tighten early.  min/max stay exact.
template <int PPT, int MINB>
__global__ void __launch_bounds__(COV_THREADS, MINB)
cov_traj_minmax_tiles_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                             const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                             unsigned* __restrict__ gmin, unsigned* __restrict__ gmax, unsigned long long tile_stride,
                             unsigned long long* __restrict__ stats) {
    constexpr int T = tile_points(PPT);
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    unsigned* smin = reinterpret_cast<unsigned*>(ptab + (size_t)W * COV_ROW_F4);
    unsigned* smax = smin + W;
    float* sqcap = reinterpret_cast<float*>(smax + W);
    float* raw = reinterpret_cast<float*>(smem4) + minmax_tiles_raw_offset_floats(W);  // [2][T*3], 128-byte aligned
    __shared__ unsigned amask[kMaskWords];
    __shared__ float wbox[kWarps][8];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv_kd = 1.f / C.kd;
    const float inf = __uint_as_float(0x7f800000u);
    unsigned n_box = 0, n_pre = 0, n_full = 0;
    unsigned long long n_iter = 0;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    for (int w = tid; w < W; w += COV_THREADS) {
        cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, ptab + (size_t)w * COV_ROW_F4);
        smin[w] = 0x7f800000u;
        smax[w] = 0u;
    }
    __syncthreads();
    const unsigned long long ntiles = (unsigned long long)((n + T - 1) / T);
    const int nwords = (W + 31) >> 5;
    const int64_t nfull = n / T;  // tiles [0, nfull) are complete and arrive by TMA; a ragged last tile is loaded by hand
    unsigned uses0 = 0, uses1 = 0;  // completed TMA phases per buffer (uniform over the block)
    if (tid == 0 && blockIdx.x < ntiles) {
        const int64_t t0 = (int64_t)(((unsigned long long)blockIdx.x * tile_stride) % ntiles);
        if (t0 < nfull) tma_load_1d(raw, xyz + t0 * (T * 3), T * 12, &mbar[0]);
    }
    int buf = 0;
    for (unsigned long long it = blockIdx.x; it < ntiles; it += gridDim.x, buf ^= 1) {
        const int64_t tile = (int64_t)((it * tile_stride) % ntiles);
        if (tid == 0 && it + gridDim.x < ntiles) {  // the other buffer was last read two barriers ago
            const int64_t tn = (int64_t)(((it + gridDim.x) * tile_stride) % ntiles);
            if (tn < nfull) tma_load_1d(raw + (buf ^ 1) * (T * 3), xyz + tn * (T * 3), T * 12, &mbar[buf ^ 1]);
        }
        float px[PPT], py[PPT], pz[PPT];
        bool valid[PPT];
        if (tile < nfull) {
            if (buf == 0) mbar_wait(&mbar[0], uses0++ & 1u);
            else mbar_wait(&mbar[1], uses1++ & 1u);
            const float* src = raw + buf * (T * 3) + (warp * (32 * PPT) + lane) * 3;
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                valid[s] = true;
                px[s] = src[s * 96];
                py[s] = src[s * 96 + 1];
                pz[s] = src[s * 96 + 2];
            }
        } else {
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                int64_t j = tile * T + warp * (32 * PPT) + s * 32 + lane;
                j = j < n ? j : n - 1;  // a duplicate cannot change a min or a max
                valid[s] = true;
                px[s] = __ldg(xyz + j * 3);
                py[s] = __ldg(xyz + j * 3 + 1);
                pz[s] = __ldg(xyz + j * 3 + 2);
            }
        }
        float3 wlo, whi;
        warp_box<PPT>(px, py, pz, valid, wlo, whi);
        if (lane == 0) {
            wbox[warp][0] = wlo.x; wbox[warp][1] = wlo.y; wbox[warp][2] = wlo.z;
            wbox[warp][4] = whi.x; wbox[warp][5] = whi.y; wbox[warp][6] = whi.z;
        }
        __syncthreads();
        {
            float3 blo, bhi;
            block_box(wbox, blo, bhi);
            for (int c = warp; c < nwords; c += kWarps) {
                const int w = c * 32 + lane;
                bool active = false;
                if (w < W) {
                    const unsigned mn = min(ld_relaxed(gmin + w), smin[w]);
                    const unsigned mx = max(ld_relaxed(gmax + w), smax[w]);
                    float cap = inf;
                    if (mn == 0u && mx != 0u) cap = (1e-4f - __log2f(__uint_as_float(mx))) * inv_kd * 1.000001f;
                    sqcap[w] = cap;
                    active = !(box_q2lb(blo, bhi, ptab[(size_t)w * COV_ROW_F4 + 3]) > cap);  // NaN cap: evaluate
                }
                const unsigned bal = __ballot_sync(kFull, active);
                if (lane == 0) amask[c] = bal;
            }
        }
        __syncthreads();
        n_iter += (unsigned long long)W;
        for (int c = 0; c < nwords; ++c) {
            unsigned word = amask[c];
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
        __syncthreads();
        for (int w = tid; w < W; w += COV_THREADS) {  // publish what this tile may have changed
            if ((amask[w >> 5] >> (w & 31)) & 1u) {
                atomicMin(gmin + w, smin[w]);
                atomicMax(gmax + w, smax[w]);
            }
        }
    }
    if (lane == 0) {
        atomicAdd(stats + 2, n_iter);
        atomicAdd(stats + 3, (unsigned long long)n_full);
        atomicAdd(stats + 5, (unsigned long long)n_pre);
        atomicAdd(stats + 7, (unsigned long long)n_box);
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

// Gate bit matrix addressing.  Row w holds kWarps groups of PPT ballot words (group g = the words of warp g).
// Dense kernel: rows are padded by 4 words (conflict-free when lanes walk different rows at the same word).
// Pruned kernel: rows are unpadded (shared memory goes to the second tile buffer) and group g sits at slot
// g ^ (w & 7) instead, which spreads the same access pattern over 8 bank groups.
template <int PPT, bool SWZ>
__device__ __forceinline__ unsigned* bit_row_group(unsigned* bits, int w, int group) {
    constexpr int RS = SWZ ? bit_words(PPT) : bit_stride(PPT);
    return bits + (size_t)w * RS + (SWZ ? ((group ^ (w & 7)) * PPT) : group * PPT);
}
template <int PPT, bool SWZ>
__device__ __forceinline__ unsigned bit_word(const unsigned* bits, int w, int k) {  // word k of row w (k = group*PPT + s)
    constexpr int RS = SWZ ? bit_words(PPT) : bit_stride(PPT);
    const int g = k / PPT, sidx = k - g * PPT;
    return bits[(size_t)w * RS + (SWZ ? ((g ^ (w & 7)) * PPT) : g * PPT) + sidx];
}

// Phase-1 body of the fused pass for U consecutive poses starting at w (U*PPT independent chains).
template <int PPT, int U, bool AMIN, bool THR5>
__device__ __forceinline__ void fused_pose_iter(int w, const float4* __restrict__ ptab, unsigned* __restrict__ bits,
                                                int warp, const float (&px)[PPT], const float (&py)[PPT],
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
            unsigned* brow = bit_row_group<PPT, THR5>(bits, w + u, warp);
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

// Pass B.  PRUNE (the product default) adds tile-level pruning: v3.w of a pose row holds
// qthr = (1e-4 - log2(thr))/kd with thr the conservative gate threshold; a pair with q2 > qthr has
// m <= 2^-(kd q2)(1+1.3e-5) < thr, so it is neither gated nor the arg-max, contributes logit(1/2) = 0 to the
// log-odds sum and nothing to the gradient.  Per tile the block keeps the ascending list of poses whose qthr-ball
// meets the tile's bounding box; a warp runs the exact body for a listed pose only when the ball also meets the
// warp's own box and one of its 32*PPT points passes the per-point test.  Poses with min_j m > 0 are always
// listed (their arg-min points carry gradient).  Results are bit-identical to the dense kernel.
// Point layout of a tile: local = warp*32*PPT + s*32 + lane (a warp owns 32*PPT consecutive points), so ballot
// word k of a pose row covers points [32k, 32k+32).
template <int PPT, bool HAS_UP, int U, bool PRUNE>
__global__ void __launch_bounds__(COV_THREADS, 2)
cov_traj_fused_kernel(const float* __restrict__ xyz, int64_t n, const float* __restrict__ poses,
                      const float* __restrict__ quats, int W, const float* __restrict__ K9, CovConst C,
                      const float* __restrict__ minmax, const float* __restrict__ upstream,
                      const int32_t* __restrict__ out_index, float* __restrict__ rewards,
                      float* __restrict__ partials, double* __restrict__ sumr_partials, double* __restrict__ acc,
                      int seg_log2_dense, unsigned long long* __restrict__ stats) {
    constexpr int T = tile_points(PPT);
    constexpr int NW = bit_words(PPT);
    constexpr int RS = PRUNE ? NW : bit_stride(PPT);
    // shared memory: pose table | tile buffer 0 (T points, xyz interleaved) | [PRUNE: tile buffer 1] | G_j | gate bits |
    // block accumulators | [PRUNE: active-pose list]          (tile buffers first: they need 128-byte alignment)
    extern __shared__ float4 smem4[];
    float4* ptab = smem4;
    float* raw = reinterpret_cast<float*>(smem4) + fused_raw_offset_floats(W);
    float* Gs = raw + (PRUNE ? 2 : 1) * (T * 3);
    unsigned* bits = reinterpret_cast<unsigned*>(Gs + T);
    float* accs = reinterpret_cast<float*>(bits + (size_t)W * RS);
    unsigned short* alist = reinterpret_cast<unsigned short*>(accs + (size_t)W * 8);
    __shared__ double red[kWarps];
    __shared__ int amin_pos;  // some pose has min_j m > 0: its arg-min points carry gradient
    __shared__ unsigned amask[kMaskWords];
    __shared__ float wbox[kWarps][8];
    __shared__ int n_active;
    __shared__ __align__(8) unsigned long long mbar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        amin_pos = 0;
        if (PRUNE) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            mbar_fence_init();
        }
    }
    __syncthreads();
    for (int w = tid; w < W; w += COV_THREADS) {
        float4* row = ptab + (size_t)w * COV_ROW_F4;
        cov_pose_row(poses + 3 * w, quats + 4 * w, K9, C, row);
        const float a = minmax[w];
        const float b = __fsub_rn(minmax[W + w], a);
        const float hb = 0.5f * b;
        const float thr = (a + hb) * (1.f - 9.5367431640625e-07f);  // conservative gate threshold (exact test in the rare path)
        row[5].w = thr;
        // pruning: q2 above this cannot reach thr (thr <= 0 or NaN, or a > 0: never prune)
        const float qthr = (thr > 0.f && !(a > 0.f)) ? (float)((1e-4 - log2((double)thr)) / (double)C.kd * 1.000001)
                                                      : __uint_as_float(0x7f800000u);
        row[3].w = PRUNE ? qthr : thr;
        row[4] = make_float4(hb, b, __frcp_rn(b), a);
        if (a > 0.f) amin_pos = 1;  // benign race: every writer stores 1
    }
    for (int i = tid; i < W * 8; i += COV_THREADS) accs[i] = 0.f;
    __syncthreads();

    double sum_r = 0.0;
    unsigned long long n_iter = 0;
    unsigned n_box = 0, n_pre = 0, n_full = 0;
    const int64_t ntiles = (n + T - 1) / T;
    const int64_t nfull = n / T;  // complete tiles arrive by TMA (PRUNE); a ragged last tile is loaded by hand
    const int nwords = (W + 31) >> 5;
    const bool check_amin = amin_pos != 0;
    unsigned uses0 = 0, uses1 = 0;
    int buf = 0;
    if (PRUNE && tid == 0 && (int64_t)blockIdx.x < nfull) tma_load_1d(raw, xyz + (int64_t)blockIdx.x * (T * 3), T * 12, &mbar[0]);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ------------------------------ phase 1: every (point, pose) that can matter ------------------------------
        float px[PPT], py[PPT], pz[PPT], L[PPT];
        bool valid[PPT];
        const int lbase = warp * (32 * PPT) + lane;
        float* pt = raw + (PRUNE ? buf : 0) * (T * 3);  // this tile's points in shared memory (phase 2 reads them)
        if (PRUNE) {
            // the other buffer was last read in phase 2 of the previous tile, which ended at a barrier
            if (tid == 0 && tile + gridDim.x < nfull)
                tma_load_1d(raw + (buf ^ 1) * (T * 3), xyz + (tile + gridDim.x) * (T * 3), T * 12, &mbar[buf ^ 1]);
        }
        if (PRUNE && tile < nfull) {
            if (buf == 0) mbar_wait(&mbar[0], uses0++ & 1u);
            else mbar_wait(&mbar[1], uses1++ & 1u);
#pragma unroll
            for (int s = 0; s < PPT; ++s) {
                valid[s] = true;
                px[s] = pt[(lbase + s * 32) * 3];
                py[s] = pt[(lbase + s * 32) * 3 + 1];
                pz[s] = pt[(lbase + s * 32) * 3 + 2];
                L[s] = 0.f;
            }
        } else {
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
        }
        if (PRUNE) {
            float3 wlo, whi;
            warp_box<PPT>(px, py, pz, valid, wlo, whi);
            if (lane == 0) {
                wbox[warp][0] = wlo.x; wbox[warp][1] = wlo.y; wbox[warp][2] = wlo.z;
                wbox[warp][4] = whi.x; wbox[warp][5] = whi.y; wbox[warp][6] = whi.z;
            }
            __syncthreads();
            {
                float3 blo, bhi;
                block_box(wbox, blo, bhi);
                for (int c = warp; c < nwords; c += kWarps) {
                    const int w = c * 32 + lane;
                    bool active = false;
                    if (w < W) {
                        const float4 v3 = ptab[(size_t)w * COV_ROW_F4 + 3];
                        active = !(box_q2lb(blo, bhi, v3) > v3.w);
                    }
                    const unsigned bal = __ballot_sync(kFull, active);
                    if (lane == 0) amask[c] = bal;
                }
            }
            __syncthreads();
            n_iter += (unsigned long long)W;
            {   // nothing listed (every warp reaches the same verdict): rewards are exactly 1/2, no gradient, next tile
                unsigned anyw = 0u;
                for (int c = lane; c < nwords; c += 32) anyw |= amask[c];
                if (!__any_sync(kFull, anyw != 0u)) {
                    int cnt = 0;
#pragma unroll
                    for (int s = 0; s < PPT; ++s) cnt += valid[s] ? 1 : 0;
                    sum_r += 0.5 * (double)cnt;
                    if (!out_index) {  // (with out_index the caller-order rewards were pre-filled with 1/2)
#pragma unroll
                        for (int s = 0; s < PPT; ++s)
                            if (valid[s]) rewards[tile * T + lbase + s * 32] = 0.5f;
                    }
                    buf ^= 1;
                    continue;
                }
            }
            if (warp == 0) {  // ascending list of the tile's active poses for phase 2 (read after the next barrier)
                int base = 0;
                for (int c0 = 0; c0 < nwords; c0 += 32) {
                    const unsigned word = (c0 + lane < nwords) ? amask[c0 + lane] : 0u;
                    const int cnt = __popc(word);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += t;
                    }
                    int pos = base + incl - cnt;
                    unsigned wv = word;
                    while (wv) {
                        alist[pos++] = (unsigned short)((c0 + lane) * 32 + __ffs(wv) - 1);
                        wv &= wv - 1;
                    }
                    base += __shfl_sync(kFull, incl, 31);
                }
                if (lane == 0) n_active = base;
            }
            for (int c = 0; c < nwords; ++c) {
                unsigned word = amask[c];
                while (word) {
                    const int w = c * 32 + __ffs(word) - 1;
                    word &= word - 1;
                    const float4 v3 = ptab[(size_t)w * COV_ROW_F4 + 3];
                    ++n_box;
                    bool run = !(box_q2lb(wlo, whi, v3) > v3.w);
                    if (run) {
                        ++n_pre;
                        float qmin = cov_q2(px[0], py[0], pz[0], v3);
#pragma unroll
                        for (int s = 1; s < PPT; ++s) qmin = fminf(qmin, cov_q2(px[s], py[s], pz[s], v3));
                        run = __any_sync(kFull, !(qmin > v3.w));
                    }
                    if (run) {
                        ++n_full;
                        if (check_amin) fused_pose_iter<PPT, 1, true, true>(w, ptab, bits, warp, px, py, pz, L, C, acc, lane);
                        else fused_pose_iter<PPT, 1, false, true>(w, ptab, bits, warp, px, py, pz, L, C, acc, lane);
                    } else if (lane == 0) {  // this warp's words of a listed row must still be defined
                        unsigned* brow = bit_row_group<PPT, true>(bits, w, warp);
#pragma unroll
                        for (int s = 0; s < PPT; ++s) brow[s] = 0u;
                    }
                }
            }
        } else {
            int w = 0;
            if (check_amin) {  // arg-min points carry gradient: every pair must be looked at
                for (; w < W; ++w) fused_pose_iter<PPT, 1, true, false>(w, ptab, bits, warp, px, py, pz, L, C, acc, lane);
            } else {
                for (; w + U <= W; w += U) fused_pose_iter<PPT, U, false, false>(w, ptab, bits, warp, px, py, pz, L, C, acc, lane);
                for (; w < W; ++w) fused_pose_iter<PPT, 1, false, false>(w, ptab, bits, warp, px, py, pz, L, C, acc, lane);
            }
        }
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const float r = 1.f / (1.f + expf(-L[s]));
            float g = r * (1.f - r);
            if (valid[s]) {
                const int64_t j = tile * T + lbase + s * 32;
                sum_r += (double)r;
                // pruned + out_index: the caller-order rewards were pre-filled with 1/2 (what every ungated point gets,
                // exactly), so only the others are scattered
                const bool store = !(PRUNE && out_index) || r != 0.5f;
                int64_t jo = j;
                if (out_index && (store || HAS_UP)) jo = (int64_t)out_index[j];
                if (store) rewards[jo] = r;
                if (HAS_UP) g *= upstream[jo];
            }
            Gs[lbase + s * 32] = g;
        }
        __syncthreads();
        // ------------------------------ phase 2: gated pairs, pose-major ------------------------------
        const int nposes2 = PRUNE ? n_active : W;
        int seg_log2 = seg_log2_dense;
        if (PRUNE) {  // split each listed pose row over 2^seg_log2 lanes until there are >= 2 tasks per thread
            seg_log2 = 0;
            while ((nposes2 << seg_log2) < 2 * COV_THREADS && (2 << seg_log2) <= NW && seg_log2 < 5) ++seg_log2;
        }
        const int nseg = 1 << seg_log2;       // lanes that share one pose row
        const int wps = NW >> seg_log2;       // ballot words per lane
        const int ntask = nposes2 << seg_log2;
        for (int base = 0; base < ntask; base += COV_THREADS) {
            const int task = base + tid;
            const bool live = task < ntask;
            const int wi = live ? (task >> seg_log2) : 0;
            const int w = PRUNE ? (int)alist[wi] : wi;
            const int seg = task & (nseg - 1);
            const float4* row = ptab + (size_t)w * COV_ROW_F4;
            const float4 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3], v4 = row[4], v5 = row[5];
            int k = seg * wps;
            const int kend = live ? k + wps : k;
            unsigned word = live ? bit_word<PPT, PRUNE>(bits, w, k) : 0u;
            float f0 = 0.f, f1 = 0.f, f2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, se = 0.f, sep = 0.f;
            while (true) {
                while (word == 0u && k + 1 < kend) word = bit_word<PPT, PRUNE>(bits, w, ++k);
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
        __syncthreads();
        buf ^= 1;
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
        atomicAdd(stats + 0, n_iter);
        atomicAdd(stats + 1, (unsigned long long)n_full);
        atomicAdd(stats + 4, (unsigned long long)n_pre);
        atomicAdd(stats + 6, (unsigned long long)n_box);
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

// Tile visiting order of the pruned pass A: it -> (it * stride) mod ntiles with stride ~ ntiles/phi, coprime to
// ntiles, so any run of consecutive `it` is spread evenly over the (spatially sorted) cloud.
unsigned long long golden_stride(unsigned long long ntiles) {
    if (ntiles < 3) return 1;
    auto gcd = [](unsigned long long a, unsigned long long b) {
        while (b) { const unsigned long long t = a % b; a = b; b = t; }
        return a;
    };
    unsigned long long s = (unsigned long long)((double)ntiles * 0.6180339887498949);
    if (s < 1) s = 1;
    while (gcd(s, ntiles) != 1) ++s;
    return s % ntiles ? s % ntiles : 1;
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
    while (fused_smem_bytes(w + 1, 1, false) <= kSmemCap && fused_smem_bytes(w + 1, 1, true) <= kSmemCap) ++w;
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
    const bool prune = cov_pruning_enabled() != 0;
    int ppt = pick_ppt(n, W, false, prune);
    if (ppt == 4 && (n + tile_points(8) - 1) / tile_points(8) >= 4 * (int64_t)cov_sm_count_cached() &&
        (!prune || minmax_tiles_smem_bytes(W, 8) <= kSmemCap))
        ppt = 8;
    const size_t smem = minmax_smem_bytes(W);
    const int mm_variant = dev_variant("COV_DEV_MM", 0);
    const int eff_ppt = (ppt == 8) ? ((mm_variant == 0 || mm_variant == 4) ? 8 : (mm_variant == 5 ? 2 : 4)) : (ppt >= 4 ? 4 : ppt);
    const int64_t ntiles = (n + tile_points(eff_ppt) - 1) / tile_points(eff_ppt);
#define LAUNCH_MM(P, B, U)                                                                                        \
    {                                                                                                             \
        if (prune) {                                                                                              \
            const size_t smem_t = minmax_tiles_smem_bytes(W, P);                                                  \
            const int grid = grid_for(cov_traj_minmax_tiles_kernel<P, B>, smem_t, ntiles);                        \
            cov_traj_minmax_tiles_kernel<P, B><<<grid, COV_THREADS, smem_t, s>>>(                                 \
                xyz, n, poses, quats, W, K, C, gmin, gmax, golden_stride((unsigned long long)ntiles), stats);     \
        } else {                                                                                                  \
            const int grid = grid_for(cov_traj_minmax_kernel<P, B, U, false>, smem, ntiles);                      \
            cov_traj_minmax_kernel<P, B, U, false><<<grid, COV_THREADS, smem, s>>>(xyz, n, poses, quats, W, K, C, \
                                                                                   gmin, gmax, stats);            \
        }                                                                                                         \
    }
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
                              const int32_t* reward_index, float* rewards, double* acc, void* ws, size_t ws_bytes,
                              void* stream) {
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
    const bool prune = cov_pruning_enabled() != 0;
    int ppt = pick_ppt(n, W, true, prune);
    const int fvariant = dev_variant("COV_DEV_F", 0);
    if (ppt == 4 && (fvariant == 2 || fvariant == 3)) ppt = 2;
    if (ppt == 0) {
        cov_set_error("cov_traj_fused: %d poses do not fit in shared memory", W);
        return COV_ERR_UNSUPPORTED;
    }
    const size_t smem = fused_smem_bytes(W, ppt, prune);
    const int64_t ntiles = (n + tile_points(ppt) - 1) / tile_points(ppt);
    // phase-2 parallelism: split each pose row over 2^seg_log2 lanes until there are >= 2 tasks per thread
    int seg_log2 = 0;
    while ((W << seg_log2) < 2 * COV_THREADS && (2 << seg_log2) <= bit_words(ppt) && seg_log2 < 5) ++seg_log2;
    double* sumr = reinterpret_cast<double*>(ws);
    float* partials = reinterpret_cast<float*>(sumr + COV_MAX_GRID);
    cudaMemsetAsync(acc, 0, ((size_t)W * COV_ACC_STRIDE + 1) * sizeof(double), s);
    if (reward_index && prune) {  // the pruned kernel scatters only rewards != 1/2
        const int fgrid = (int)std::min<int64_t>((n + 1023) / 1024, (int64_t)cov_sm_count_cached() * 16);
        cov_fill_kernel<<<fgrid, 256, 0, s>>>(rewards, n, 0.5f);
    }
    int grid = 1;
#define LAUNCH_F(P, UP, U)                                                                                        \
    {                                                                                                             \
        if (prune) {                                                                                              \
            grid = grid_for(cov_traj_fused_kernel<P, UP, 1, true>, smem, ntiles);                                 \
            cov_traj_fused_kernel<P, UP, 1, true><<<grid, COV_THREADS, smem, s>>>(                                \
                xyz, n, poses, quats, W, K, C, minmax, upstream, reward_index, rewards, partials, sumr, acc, seg_log2,  \
                stats);  \
        } else {                                                                                                  \
            grid = grid_for(cov_traj_fused_kernel<P, UP, U, false>, smem, ntiles);                                \
            cov_traj_fused_kernel<P, UP, U, false><<<grid, COV_THREADS, smem, s>>>(                               \
                xyz, n, poses, quats, W, K, C, minmax, upstream, reward_index, rewards, partials, sumr, acc, seg_log2,  \
                stats);  \
        }                                                                                                         \
    }
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
