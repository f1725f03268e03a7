// cov_hull.cu — GPU stage of Katz hidden-point removal: vertex set of conv(flipped U {0})
// (reference src/tools.py:56-64 hands this to scipy/Qhull on the CPU).
//
// Instead of building hull facets, every point decides for itself whether it is an extreme point
// (hull_core.h): a 2-D least-tilt LP over its neighbours on the direction sphere, with a coverage bound
// that makes the neighbourhood search rigorous and fp64 certificates for both outcomes.  The flipped
// cloud is a thin shell (|f| in [199,200] * max|p| for R_param = 2), so a point's fate is decided by the
// few dozen..thousand points within a small cap around its direction; a counting sort by direction voxel
// makes those contiguous.  Kernels:
//   hull_prep      voxel key + histogram + max |f|                       (12 B/point read)
//   hull_scan      exclusive scan of the G^3 histogram (three small launches over 4096-voxel chunks)
//   hull_occupied  compact list of non-empty voxels (for the all-voxel stage)
//   hull_scatter   counting-sort scatter into (x, y, z, original index) records
//   hull_classify  stage 1, one thread per point: radius <= 1 and an evaluation budget, one flat loop per warp
//   hull_local<32>   stage 2, one warp per point handed on: cubes up to radius 6, candidates spread over the lanes
//   hull_local<256>  stage 3, one block per point handed on: radius 7..16
//   hull_far       stage 4, one block per point: every occupied voxel, culled by a bound (wide tilt box, full active set)
//   hull_origin    GJK distance from the origin to conv(F) in one cluster of 8 blocks (is the origin a vertex?)
#include <cooperative_groups.h>

#include <cstdlib>

#include "cov_common.cuh"
#include "hull_core.h"
#include "../../include/coverage_b200.h"

namespace {

constexpr int kHullMaxG = 128;
// one-thread-per-point near phase: radius and evaluation budget after which a point is handed to the warp-per-point
// stage.  Measured on 1 M-point clouds (scripts/hull_knobs.py, profiles/r02ab_hull_knobs.txt): a radius beyond 1 makes the
// warp wait for its slowest lane (radius 6, no budget: 4.9 ms shell, 23 ms half space; radius 1: 3.9 and 7.7 ms).  The
// budget bounds what one lane can cost; the best absolute value follows the density of the direction grid (shell cloud,
// 13 records per voxel: 200-300; half-space cloud, 50 per voxel: 700-1000), hence a budget per record of an average
// occupied voxel (a per-point budget from the point's OWN voxel was worse: 8.3 instead of 7.7 ms on the half-space cloud).
constexpr int kNearRadius = 1;
constexpr int kNearBudget = -20;   // < 0: evaluations per record of an average occupied voxel, clamped to [256, 4096]
constexpr int kWarpRadius = 6;    // warp-per-point stage: largest Chebyshev radius; beyond it a block takes the point
constexpr int kMidRadius = 16;    // block-per-point stage: largest Chebyshev radius before the all-voxel sweep (half-space cloud:
                                  // 747 points need more than radius 6, none more than 16)

int hull_grid_size(int64_t n) {
    int G = (int)lround(sqrt((double)n / (12.0 * 3.141592653589793)));
    return G < 1 ? 1 : (G > kHullMaxG ? kHullMaxG : G);
}

struct HullWs {  // carve-up of the caller's workspace
    int* key;             // n
    int* cell_count;      // G^3 + 1 (becomes cell_start after the scan)
    int* cursor;          // G^3
    int* occ;             // G^3
    float4* sorted;       // n
    unsigned long long* rho_max_bits;  // 1
    int* n_occ;           // 1
    int* n_valid;         // 1
    int* n_far;           // 1: points handed on to the all-voxel sweep (their sorted ids: `far`)
    int* n_mid;           // 1: points the near phase gave up on (their sorted ids reuse `key`)
    int* n_far2;          // 1: points the block stage hands to the all-voxel sweep (their sorted ids reuse `key` again)
    int* far;             // n
    int* block_tot;       // ceil(G^3 / 4096): chunk totals / offsets of the histogram scan
};

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

size_t hull_carve(void* base, int64_t n, int G, HullWs* w) {
    const size_t ncell = (size_t)G * G * G;
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off += align16(bytes); return p; };
    HullWs t;
    t.sorted = (float4*)take((size_t)n * sizeof(float4));
    t.key = (int*)take((size_t)n * sizeof(int));
    t.far = (int*)take((size_t)n * sizeof(int));
    t.cell_count = (int*)take((ncell + 1) * sizeof(int));
    t.cursor = (int*)take(ncell * sizeof(int));
    t.occ = (int*)take(ncell * sizeof(int));
    t.rho_max_bits = (unsigned long long*)take(16);
    t.n_occ = (int*)take(16);
    t.n_valid = (int*)take(16);
    t.n_far = (int*)take(16);
    t.n_mid = (int*)take(16);
    t.n_far2 = (int*)take(16);
    t.block_tot = (int*)take(((ncell + 4095) / 4096 + 1) * sizeof(int));
    if (w) *w = t;
    return off;
}

__global__ void __launch_bounds__(256)
hull_prep_kernel(const float* __restrict__ f, int64_t n, int G, int* __restrict__ key, int* __restrict__ cell_count,
                 unsigned long long* __restrict__ rho_max_bits) {
    double mx = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double x = f[i * 3], y = f[i * 3 + 1], z = f[i * 3 + 2];
        const double rho = sqrt(x * x + y * y + z * z);
        int k = -1;
        if (rho > 0.0 && rho < 1e300) {  // a point at the origin (or a NaN) duplicates the extra hull point: never a vertex
            k = (hull_cell_coord(x / rho, G) * G + hull_cell_coord(y / rho, G)) * G + hull_cell_coord(z / rho, G);
            atomicAdd(cell_count + k, 1);
            mx = fmax(mx, rho);
        }
        key[i] = k;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(rho_max_bits, (unsigned long long)__double_as_longlong(mx));
}

// exclusive scan in place over ncell+1 entries (entry ncell receives the total).  One block walks the histogram in
// chunks of 4096 consecutive cells (4 per thread, so a warp touches 512 contiguous bytes) and carries the running total.
// Exclusive scan of the cell histogram in three small launches (a single block walking 2 M cells took 1 ms):
// per-block totals of 4096-cell chunks, a one-block scan of those totals, then each block scans its chunk from its offset.
constexpr int kScanChunk = 4096;

__device__ __forceinline__ int hull_block_exclusive_scan(int s, int* wtot, int& total) {  // 1024 threads; s = thread's sum
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
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
        if (lane == 31) wtot[32] = sc;
    }
    __syncthreads();
    total = wtot[32];
    return wtot[warp] + incl - s;
}

__global__ void __launch_bounds__(1024) hull_scan_totals_kernel(const int* __restrict__ cnt, int64_t ncell, int* __restrict__ block_tot) {
    __shared__ int wtot[33];
    const int64_t i = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * 4;
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (i + k < ncell) ? cnt[i + k] : 0;
    int total;
    hull_block_exclusive_scan(s, wtot, total);
    if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) hull_scan_offsets_kernel(int* __restrict__ block_tot, int nblocks, int* __restrict__ cnt,
                                                                 int64_t ncell, int* __restrict__ n_valid) {
    __shared__ int wtot[33];
    int carry = 0;
    for (int base = 0; base < nblocks; base += 1024) {  // one round up to 1024 chunks = 4 M cells
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? block_tot[i] : 0;
        int total;
        const int ex = hull_block_exclusive_scan(v, wtot, total);
        if (i < nblocks) block_tot[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cnt[ncell] = carry;
        *n_valid = carry;
    }
}

__global__ void __launch_bounds__(1024) hull_scan_apply_kernel(int* __restrict__ cnt, int64_t ncell, const int* __restrict__ block_off) {
    __shared__ int wtot[33];
    const int64_t i = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i + k < ncell) ? cnt[i + k] : 0;
    int total;
    int run = block_off[blockIdx.x] + hull_block_exclusive_scan(v[0] + v[1] + v[2] + v[3], wtot, total);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i + k < ncell) cnt[i + k] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(256) hull_occupied_kernel(const int* __restrict__ cell_start, int64_t ncell,
                                                             int* __restrict__ occ, int* __restrict__ n_occ) {
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c < ncell && cell_start[c + 1] > cell_start[c]) occ[atomicAdd(n_occ, 1)] = (int)c;
}

__global__ void __launch_bounds__(256)
hull_scatter_kernel(const float* __restrict__ f, int64_t n, const int* __restrict__ key, const int* __restrict__ cell_start,
                    int* __restrict__ cursor, float4* __restrict__ sorted, uint8_t* __restrict__ mask) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int k = key[i];
        mask[i] = 0;
        if (k < 0) continue;
        const int pos = cell_start[k] + atomicAdd(cursor + k, 1);
        sorted[pos] = make_float4(f[i * 3], f[i * 3 + 1], f[i * 3 + 2], __int_as_float((int)i));
    }
}

// ---- cooperative stages for the points the one-thread near phase gives up on ----
// Every participating thread keeps the same LP (the updates are deterministic, so the copies never diverge); per round
// the threads scan disjoint candidates for the MOST VIOLATED half-plane, agree on one, and every thread adds it.  The
// answer (feasible / infeasible) does not depend on the pivot order; certificates as in hull_core.h.
struct HullPick {
    double score, a, b, c;   // most violated half-plane seen so far (score < 0), none: id < 0
    int id;
};

__device__ __forceinline__ void hull_pick_init(HullPick& pk) {
    pk.score = 0.0; pk.a = 0.0; pk.b = 0.0; pk.c = 0.0; pk.id = -1;
}

// half-plane (a, bq, c) of record j against the current LP: kept in pk when it is violated, not active yet, and more
// violated than what pk holds
__device__ __forceinline__ void hull_score(const HullLP& L, double a, double bq, double c, int j, HullPick& pk) {
    if (!hull_violated(L, a, bq, c)) return;
    bool active = false;
    for (int q = 0; q < L.n; ++q) active |= L.id[q] == j;
    if (active) return;  // the residual is evaluation noise of an optimum that sits on its line
    const double score = (a * L.x0 + bq * L.x1 + c) / (fabs(a * L.x0) + fabs(bq * L.x1) + fabs(c));
    if (score < pk.score || (score == pk.score && (pk.id < 0 || j < pk.id))) { pk.score = score; pk.a = a; pk.b = bq; pk.c = c; pk.id = j; }
}

__device__ __forceinline__ void hull_consider(const HullFrame& F, const HullLP& L, const float4& sp, int j, HullPick& pk) {
    double a, bq, c;
    hull_constraint(F, L.tilt, (double)sp.x, (double)sp.y, (double)sp.z, a, bq, c);
    hull_score(L, a, bq, c, j, pk);
}

// the warp's most violated half-plane (ties: smallest id) — the same on every lane afterwards
__device__ __forceinline__ void hull_warp_pick(HullPick& pk) {
    double wbest = pk.score;
    int wid = pk.id;
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, wbest, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wid, o);
        if (oi >= 0 && (wid < 0 || ob < wbest || (ob == wbest && oi < wid))) { wbest = ob; wid = oi; }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, pk.id == wid && pk.score == wbest)) - 1;
    pk.a = __shfl_sync(0xffffffffu, pk.a, src);
    pk.b = __shfl_sync(0xffffffffu, pk.b, src);
    pk.c = __shfl_sync(0xffffffffu, pk.c, src);
    pk.score = wbest;
    pk.id = wid;
}

// the GROUP's most violated half-plane, the same on every thread afterwards: a warp's pick, or the warps' picks joined
// through shared memory (two block barriers)
template <int GROUP>
__device__ __forceinline__ void hull_group_pick(HullPick& pk, double* s_score, double (*s_abc)[3], int* s_id) {
    hull_warp_pick(pk);
    if (GROUP > 32) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
            s_score[warp] = pk.score; s_id[warp] = pk.id;
            s_abc[warp][0] = pk.a; s_abc[warp][1] = pk.b; s_abc[warp][2] = pk.c;
        }
        __syncthreads();
        hull_pick_init(pk);
#pragma unroll
        for (int w = 0; w < GROUP / 32; ++w) {
            const int oi = s_id[w];
            const double ob = s_score[w];
            if (oi >= 0 && (pk.id < 0 || ob < pk.score || (ob == pk.score && oi < pk.id))) {
                pk.score = ob; pk.id = oi; pk.a = s_abc[w][0]; pk.b = s_abc[w][1]; pk.c = s_abc[w][2];
            }
        }
        __syncthreads();  // everybody has read the picks before the next call overwrites them
    }
}

// one thread writes the decision for sorted point `ps` (certifying an INSIDE answer first)
__device__ __forceinline__ void hull_finalize(int rc, const int* cert, const HullFrame& F, const float4& ps,
                                              const float4* __restrict__ sorted, uint8_t* __restrict__ mask, int* __restrict__ info) {
    if (rc == HULL_INSIDE) {
        bool ok = cert[0] >= 0 && cert[1] >= 0 && cert[2] >= 0 && cert[0] != cert[1] && cert[1] != cert[2] && cert[0] != cert[2];
        if (ok) {
            const float4 s1 = sorted[cert[0]], s2 = sorted[cert[1]], s3 = sorted[cert[2]];
            const double a1[3] = {s1.x, s1.y, s1.z}, a2[3] = {s2.x, s2.y, s2.z}, a3[3] = {s3.x, s3.y, s3.z};
            ok = hull_certify_inside(F.p, a1, a2, a3) != 0;
        }
        if (!ok) rc = HULL_INSIDE_UNCERT;
    }
    const bool vertex = (rc == HULL_EXTREME || rc == HULL_EXTREME_UNCERT);
    mask[__float_as_int(ps.w)] = vertex ? 1 : 0;
    if (rc != HULL_EXTREME && rc != HULL_INSIDE) atomicAdd(info + 1, 1);
    if (vertex) atomicAdd(info + 2, 1);
}

// Stage 1, one THREAD per point: hull_classify_attempt's near phase (hull_core.h; same candidates in the same order, so
// every point takes the same decisions as there) unrolled into ONE flat loop in which a lane handles one candidate per
// iteration and the whole warp meets again at the top: [advance the lane's cursor over voxels and rings] -> [evaluate
// the candidate's half-plane] -> [hull_lp_add if violated].  In the nested-loop form the lanes of a warp fell out of
// step at their first LP update and never met again (ncu: 3.0 active threads per instruction in the candidate loop,
// 1.0 in the LP update); here the evaluation runs with every undecided lane and lanes that are violated in the same
// iteration update together.  A point that exhausts its evaluation budget or needs a wider search, a wider tilt box or a
// bigger active set goes on the list of the cooperative stages (hull_local_kernel, hull_far_kernel).
// (register allocation: the compiler's choice is 91 registers = 5 blocks per SM; capping it at 80 or 64 registers measured
// within 2 % of that, 40 registers with spills 20 % slower)
__global__ void __launch_bounds__(128)
hull_classify_kernel(int G, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                     const unsigned long long* __restrict__ rho_max_bits, const int* __restrict__ n_valid,
                     const int* __restrict__ n_occ, uint8_t* __restrict__ mask, int* __restrict__ info,
                     int* __restrict__ mid_list, int* __restrict__ n_mid, int r_near, int budget) {
    const int self = blockIdx.x * 128 + threadIdx.x;
    const bool mine = self < *n_valid;
    const double h = 2.0 / G;
    const double rho_max = __longlong_as_double((long long)*rho_max_bits);
    const float4 ps = mine ? sorted[self] : make_float4(1.f, 0.f, 0.f, 0.f);
    HullFrame F;
    hull_frame_init(F, (double)ps.x, (double)ps.y, (double)ps.z);
    HullLP L;
    hull_lp_init(L, HULL_TILT_NEAR);
    const int cx = hull_cell_coord(F.u[0], G), cy = hull_cell_coord(F.u[1], G), cz = hull_cell_coord(F.u[2], G);
    int r = 0, clean = -1;      // current radius; every voxel within Chebyshev radius `clean` was swept without moving the LP
    int dx = 0, dy = 0, dz = 0; // next voxel of the cube of radius r (counters, no divisions: the cursor runs one lane at a time)
    int k = 0, e = 0;           // remaining records of the current voxel
    int work = 0;
    if (budget < 0)  // relative budget: -budget evaluations per record of an average occupied voxel (the clouds differ in density)
        budget = min(max(-budget * (*n_valid / max(*n_occ, 1)), 256), 4096);
    bool changed = false;       // the LP moved during the current sweep
    bool alive = mine, defer = false;
    int rc = HULL_UNDECIDED;
    while (__any_sync(0xffffffffu, alive)) {
        int cand = -1;
        if (alive) {
            while (k >= e) {  // advance to the next non-empty voxel that still has to be swept, or end the sweep
                if (dx <= r) {
                    const int ix = cx + dx, iy = cy + dy, iz = cz + dz;
                    const int ad = dx < 0 ? -dx : dx, bd = dy < 0 ? -dy : dy, cd = dz < 0 ? -dz : dz;
                    if (++dz > r) {
                        dz = -r;
                        if (++dy > r) { dy = -r; ++dx; }
                    }
                    if (ix < 0 || ix >= G || iy < 0 || iy >= G || iz < 0 || iz >= G) continue;
                    if (max(ad, max(bd, cd)) <= clean) continue;
                    const int cell = (ix * G + iy) * G + iz;
                    const int b = cell_start[cell], e2 = cell_start[cell + 1];
                    if (budget > 0) {
                        work += e2 - b;
                        if (work > budget) { defer = true; alive = false; break; }
                    }
                    k = b;
                    e = e2;
                    continue;
                }
                // the sweep of radius r is complete
                if (changed) {  // the LP moved: everything swept so far must be re-checked
                    changed = false;
                    clean = -1;
                } else {
                    clean = r;
                    const bool whole = (cx - r <= 0 && cx + r >= G - 1 && cy - r <= 0 && cy + r >= G - 1 && cz - r <= 0 && cz + r >= G - 1);
                    if (whole || hull_coverage_ok(L, F.rho, rho_max, r * h)) {
                        if (fabs(L.x0) > L.tilt || fabs(L.x1) > L.tilt) defer = true;  // the rounding margin needs the wide tilt box
                        else rc = HULL_EXTREME;
                        alive = false;
                        break;
                    }
                    if (++r > r_near) { defer = true; alive = false; break; }
                }
                dx = dy = dz = -r;
            }
            if (alive) {
                cand = k++;
                if (cand == self) cand = -1;
            }
        }
        __syncwarp();
        double a = 0.0, bq = 0.0, c = 0.0;
        bool viol = false;
        if (cand >= 0) {
            const float4 sp = sorted[cand];
            hull_constraint(F, L.tilt, (double)sp.x, (double)sp.y, (double)sp.z, a, bq, c);
            viol = hull_violated(L, a, bq, c);
        }
        __syncwarp();
        if (viol) {
            const int ra = hull_lp_add(L, a, bq, c, cand);
            if (ra == HULL_UNDECIDED) changed = true;
            else if (ra == HULL_INSIDE) { rc = HULL_INSIDE; alive = false; }
            else if (ra == HULL_OVERFLOW) { defer = true; alive = false; }
        }
    }
    if (!mine) return;
    if (defer) mid_list[atomicAdd(n_mid, 1)] = self;
    else hull_finalize(rc, L.cert, F, ps, sorted, mask, info);
}

// Middle stages, a GROUP of threads per point (a warp, then a block of 256 for what the warp stage hands on): the near
// phase's neighbourhood search (Chebyshev cubes of growing radius around the point's direction voxel, coverage bound
// after a clean sweep) with the candidates of a cube spread over the group's threads.  The voxels of a (dx, dy) column
// are consecutive in the counting sort, so a column is one contiguous run of records.  Every round scans the WHOLE
// cube of the current radius, so a stage may start at any radius (the block stage starts where the warp stage ends).
// A point that needs a wider tilt box, a bigger active set or a radius beyond r_max goes on the next stage's list.
// (A single warp needs ~1 ms for a silhouette point whose search reaches radius 16; with only hundreds of such points
// the warp stage alone left the GPU idle for 3 of its 3.9 ms on the half-space cloud.)
template <int GROUP>
__global__ void __launch_bounds__(GROUP == 32 ? 128 : GROUP)
hull_local_kernel(int G, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                  const unsigned long long* __restrict__ rho_max_bits, const int* __restrict__ list_in,
                  const int* __restrict__ n_in, uint8_t* __restrict__ mask, int* __restrict__ info, int* __restrict__ list_out,
                  int* __restrict__ n_out, int r_first, int r_max) {
    __shared__ double s_score[8], s_abc[8][3];
    __shared__ int s_id[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gtid = GROUP == 32 ? lane : (int)threadIdx.x;           // thread within the group
    const int first = GROUP == 32 ? (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (int)blockIdx.x;
    const int stride = GROUP == 32 ? (int)((gridDim.x * blockDim.x) >> 5) : (int)gridDim.x;
    const double h = 2.0 / G;
    const double rho_max = __longlong_as_double((long long)*rho_max_bits);
    for (int item = first; item < *n_in; item += stride) {  // GROUP > 32: block-uniform control flow from here on
        const int self = list_in[item];
        const float4 ps = sorted[self];
        HullFrame F;
        hull_frame_init(F, (double)ps.x, (double)ps.y, (double)ps.z);
        const int cx = hull_cell_coord(F.u[0], G), cy = hull_cell_coord(F.u[1], G), cz = hull_cell_coord(F.u[2], G);
        HullLP L;
        hull_lp_init(L, HULL_TILT_NEAR);
        int rc = HULL_UNDECIDED;
        bool defer = false;
        int r = r_first;
        for (int round = 0; round < 4096 && rc == HULL_UNDECIDED && !defer; ++round) {
            HullPick pk;
            hull_pick_init(pk);
            const int side = 2 * r + 1;
            const int izlo = max(cz - r, 0), izhi = min(cz + r, G - 1);
            // a warp takes 32 columns at a time: every lane fetches the bounds of one column (one round trip to the L2 for 32
            // columns instead of one per column: the dependent loads were what a round's time consisted of), then the
            // warp walks the non-empty ones, its lanes striding over a column's records
            for (int col0 = (GROUP == 32 ? 0 : warp * 32); col0 < side * side; col0 += GROUP) {
                const int col = col0 + lane;
                int cb = 0, ce = 0;
                if (col < side * side) {
                    const int ix = cx + col / side - r, iy = cy + col % side - r;
                    if (ix >= 0 && ix < G && iy >= 0 && iy < G) {
                        const int cbase = (ix * G + iy) * G;
                        cb = cell_start[cbase + izlo];
                        ce = cell_start[cbase + izhi + 1];
                    }
                }
                unsigned todo = __ballot_sync(0xffffffffu, ce > cb);
                while (todo) {
                    const int q = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int b = __shfl_sync(0xffffffffu, cb, q), e = __shfl_sync(0xffffffffu, ce, q);
                    for (int j = b + lane; j < e; j += 32)
                        if (j != self) hull_consider(F, L, sorted[j], j, pk);
                }
            }
            hull_group_pick<GROUP>(pk, s_score, s_abc, s_id);
            if (pk.id < 0) {  // clean sweep of the cube of radius r
                const bool whole = (cx - r <= 0 && cx + r >= G - 1 && cy - r <= 0 && cy + r >= G - 1 && cz - r <= 0 && cz + r >= G - 1);
                if (whole || hull_coverage_ok(L, F.rho, rho_max, r * h)) {
                    if (fabs(L.x0) > L.tilt || fabs(L.x1) > L.tilt) defer = true;  // the rounding margin needs the wide tilt box
                    else rc = HULL_EXTREME;
                } else if (++r > r_max) {
                    defer = true;
                }
                continue;
            }
            const int ra = hull_lp_add(L, pk.a, pk.b, pk.c, pk.id);
            if (ra == HULL_INSIDE) rc = HULL_INSIDE;
            else if (ra == HULL_OVERFLOW) defer = true;
        }
        if (rc == HULL_UNDECIDED) defer = true;  // round limit
        if (gtid == 0) {
            if (defer) list_out[atomicAdd(n_out, 1)] = self;
            else hull_finalize(rc, L.cert, F, ps, sorted, mask, info);
        }
    }
}

// The all-voxel sweep for what is left (silhouette points of a cloud that does not surround the camera: large tilt), ONE
// BLOCK PER POINT: per round the threads scan disjoint subsets of the occupied voxels, culled by the bound on n.f over
// a voxel; warps agree by shuffles, the block through shared memory.  (One warp per point left the GPU idle: a single
// point takes a warp milliseconds, and there are only hundreds of such points.)
constexpr int kFarThreads = 256;

__global__ void __launch_bounds__(kFarThreads)
hull_far_kernel(int G, const int* __restrict__ cell_start, const float4* __restrict__ sorted, const int* __restrict__ occ,
                const int* __restrict__ n_occ_p, const unsigned long long* __restrict__ rho_max_bits,
                const int* __restrict__ far_list, const int* __restrict__ n_far, uint8_t* __restrict__ mask,
                int* __restrict__ info) {
    __shared__ double s_score[kFarThreads / 32], s_abc[kFarThreads / 32][3];
    __shared__ int s_id[kFarThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_occ = *n_occ_p;
    const double h = 2.0 / G;
    const double rho_max = __longlong_as_double((long long)*rho_max_bits);
    for (int item = blockIdx.x; item < *n_far; item += gridDim.x) {  // block-uniform control flow from here on
        const int self = far_list[item];
        const float4 ps = sorted[self];
        HullFrame F;
        hull_frame_init(F, (double)ps.x, (double)ps.y, (double)ps.z);
        int rc = HULL_UNDECIDED;
        int cert[3] = {-1, -1, -1};
        for (int attempt = 0; attempt < 2; ++attempt) {
            HullLP L;
            hull_lp_init(L, attempt == 0 ? HULL_TILT_NEAR : HULL_TILT_MAX);
            rc = HULL_UNDECIDED;
            for (int round = 0; round < 4096 && rc == HULL_UNDECIDED; ++round) {
                const double n0 = F.u[0] + L.x0 * F.e1[0] + L.x1 * F.e2[0];
                const double n1 = F.u[1] + L.x0 * F.e1[1] + L.x1 * F.e2[1];
                const double n2 = F.u[2] + L.x0 * F.e1[2] + L.x1 * F.e2[2];
                const double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
                const double np = (n0 * F.p[0] + n1 * F.p[1] + n2 * F.p[2]) * (1.0 - 1e-12);
                const double slack = nn * h * 0.8660254037844387;
                HullPick pk;
                hull_pick_init(pk);
                for (int k = threadIdx.x; k < n_occ; k += kFarThreads) {
                    const int cell = occ[k];
                    const int iz = cell % G, iy = (cell / G) % G, ix = cell / (G * G);
                    const double c0 = (ix + 0.5) * h - 1.0, c1 = (iy + 0.5) * h - 1.0, c2 = (iz + 0.5) * h - 1.0;
                    if (rho_max * (n0 * c0 + n1 * c1 + n2 * c2 + slack) < np) continue;
                    const int b = cell_start[cell], e = cell_start[cell + 1];
                    for (int j = b; j < e; ++j)
                        if (j != self) hull_consider(F, L, sorted[j], j, pk);
                }
                hull_warp_pick(pk);
                if (lane == 0) {
                    s_score[warp] = pk.score; s_id[warp] = pk.id;
                    s_abc[warp][0] = pk.a; s_abc[warp][1] = pk.b; s_abc[warp][2] = pk.c;
                }
                __syncthreads();
                double best = 0.0;
                int bid = -1, bw = 0;
#pragma unroll
                for (int w = 0; w < kFarThreads / 32; ++w) {
                    const int oi = s_id[w];
                    const double ob = s_score[w];
                    if (oi >= 0 && (bid < 0 || ob < best || (ob == best && oi < bid))) { best = ob; bid = oi; bw = w; }
                }
                const double a = s_abc[bw][0], bq = s_abc[bw][1], c = s_abc[bw][2];
                __syncthreads();  // everybody has read the picks before the next round overwrites them
                if (bid < 0) {  // clean sweep over everything the bound could not exclude
                    rc = (fabs(L.x0) > L.tilt || fabs(L.x1) > L.tilt) ? HULL_EXTREME_UNCERT : HULL_EXTREME;
                    break;
                }
                const int r = hull_lp_add(L, a, bq, c, bid);
                if (r == HULL_INSIDE || r == HULL_OVERFLOW) rc = r;
            }
            if (rc == HULL_UNDECIDED) rc = HULL_OVERFLOW;  // round limit
            if (rc == HULL_INSIDE) { cert[0] = L.cert[0]; cert[1] = L.cert[1]; cert[2] = L.cert[2]; }
            if (rc != HULL_EXTREME_UNCERT && rc != HULL_OVERFLOW) break;  // settled within this tilt box
        }
        if (threadIdx.x == 0) hull_finalize(rc, cert, F, ps, sorted, mask, info);
    }
}

// Is the origin a vertex of conv(F U {0})?  GJK distance from the origin to conv(F).  The support search of an
// iteration is a scan of the whole cloud; one block does that at one SM's load bandwidth (0.6-0.9 ms per call at 1 M
// points), so the kernel runs as ONE THREAD-BLOCK CLUSTER of 8 CTAs: each scans an eighth, the 8 candidates meet in
// CTA 0's shared memory (distributed shared memory), every CTA reads them back and updates its own copy of the simplex
// (the update is deterministic, so the copies stay identical and the loop control is uniform over the cluster).
constexpr int kOriginCtas = 8;

__global__ void __cluster_dims__(kOriginCtas, 1, 1) __launch_bounds__(1024)
hull_origin_kernel(const float* __restrict__ f, int64_t n, int* __restrict__ info) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    __shared__ HullSimplex S;
    __shared__ double rv[32];
    __shared__ long long ri[32];
    __shared__ double slot_v[kOriginCtas];       // CTA 0's copies collect the cluster's candidates
    __shared__ long long slot_i[kOriginCtas];
    __shared__ int state;  // 0 = iterate, 1 = done
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        int64_t j0 = 0;  // first finite non-zero point
        for (; j0 < n; ++j0) {
            const double q = (double)f[j0 * 3] * f[j0 * 3] + (double)f[j0 * 3 + 1] * f[j0 * 3 + 1] + (double)f[j0 * 3 + 2] * f[j0 * 3 + 2];
            if (q > 0.0 && q < 1e300) break;
        }
        state = 0;
        if (j0 >= n) {
            if (rank == 0) info[0] = 1;
            state = 1;
        } else {
            S.n = 1;
            for (int c = 0; c < 3; ++c) S.v[0][c] = S.x[c] = f[j0 * 3 + c];
        }
    }
    __syncthreads();
    double* slot_v0 = cluster.map_shared_rank(slot_v, 0);
    long long* slot_i0 = cluster.map_shared_rank(slot_i, 0);
    for (int it = 0; it < 64 && state == 0; ++it) {
        const double x0 = S.x[0], x1 = S.x[1], x2 = S.x[2];
        double best = 1e300;
        long long bj = -1;
        for (int64_t j = (int64_t)rank * 1024 + t; j < n; j += (int64_t)kOriginCtas * 1024) {
            const double d = x0 * f[j * 3] + x1 * f[j * 3 + 1] + x2 * f[j * 3 + 2];
            if (d < best) { best = d; bj = j; }  // NaNs never compare smaller
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const long long oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (oj >= 0 && (bj < 0 || ob < best || (ob == best && oj < bj))) { best = ob; bj = oj; }
        }
        if (lane == 0) { rv[warp] = best; ri[warp] = bj; }
        __syncthreads();
        if (t == 0) {
            for (int w = 1; w < 32; ++w)
                if (ri[w] >= 0 && (bj < 0 || rv[w] < best || (rv[w] == best && ri[w] < bj))) { best = rv[w]; bj = ri[w]; }
            slot_v0[rank] = best;
            slot_i0[rank] = bj;
        }
        cluster.sync();  // the 8 candidates are in CTA 0's shared memory
        if (t == 0) {
            best = 1e300;
            bj = -1;
            for (int c = 0; c < kOriginCtas; ++c) {
                const double ob = slot_v0[c];
                const long long oj = slot_i0[c];
                if (oj >= 0 && (bj < 0 || ob < best || (ob == best && oj < bj))) { best = ob; bj = oj; }
            }
            const double xx = hull_dot3(S.x, S.x);
            if (rank == 0) info[3] = it + 1;
            if (bj < 0 || best >= xx * (1.0 - 1e-10)) {  // no point lies further towards the origin: x is the closest point
                if (rank == 0) {
                    info[0] = 1;
                    if (!(best > 1e-9 * xx)) atomicAdd(info + 1, 1);  // separating margin not certified
                }
                state = 1;
            } else {
                for (int c = 0; c < 3; ++c) S.v[S.n][c] = f[bj * 3 + c];
                S.n++;
                if (hull_simplex_update(S)) {
                    if (rank == 0) {
                        info[0] = 0;
                        if (!hull_certify_origin_inside(S)) atomicAdd(info + 1, 1);
                    }
                    state = 1;
                }
            }
        }
        cluster.sync();  // everybody has read the candidates (the next iteration overwrites them) and `state` is set
    }
    if (t == 0 && rank == 0 && state == 0) {  // did not converge in 64 iterations: decide by distance, flag as uncertified
        info[0] = hull_dot3(S.x, S.x) > 0.0 ? 1 : 0;
        atomicAdd(info + 1, 1);
    }
    cluster.sync();  // no CTA leaves while another may still read CTA 0's shared memory
}

}  // namespace

extern "C" size_t cov_hpr_hull_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return hull_carve(nullptr, n, hull_grid_size(n), nullptr);
}

extern "C" int cov_hpr_hull(const float* flipped, int64_t n, uint8_t* vertex_mask, int32_t* info, void* ws,
                            size_t ws_bytes, void* stream) {
    if (n <= 0 || !flipped || !vertex_mask || !info || !ws) {
        cov_set_error("cov_hpr_hull: null pointer or empty cloud");
        return COV_ERR_ARG;
    }
    if (n >= ((int64_t)1 << 31) - 1) {
        cov_set_error("cov_hpr_hull: n >= 2^31 not supported");
        return COV_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)ws) & 15) {
        cov_set_error("cov_hpr_hull: workspace must be 16-byte aligned");
        return COV_ERR_ALIGN;
    }
    const int G = hull_grid_size(n);
    HullWs w;
    const size_t need = hull_carve(ws, n, G, &w);
    if (ws_bytes < need) {
        cov_set_error("cov_hpr_hull: workspace %zu < %zu bytes", ws_bytes, need);
        return COV_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ncell = (int64_t)G * G * G;
    // one memset clears histogram, cursors, occupied list and the three scalars (contiguous in the carve-up)
    cudaMemsetAsync(w.cell_count, 0, (size_t)((char*)w.n_far2 + 16 - (char*)w.cell_count), s);
    cudaMemsetAsync(info, 0, 4 * sizeof(int32_t), s);
    int64_t nb = (n + 255) / 256;
    const int64_t cap = (int64_t)cov_sm_count_cached() * 16;
    if (nb > cap) nb = cap;
    hull_prep_kernel<<<(unsigned)nb, 256, 0, s>>>(flipped, n, G, w.key, w.cell_count, w.rho_max_bits);
    {
        // `cursor` (zero until the scatter below) lends its head to the chunk totals
        const int nchunk = (int)((ncell + kScanChunk - 1) / kScanChunk);
        hull_scan_totals_kernel<<<nchunk, 1024, 0, s>>>(w.cell_count, ncell, w.block_tot);
        hull_scan_offsets_kernel<<<1, 1024, 0, s>>>(w.block_tot, nchunk, w.cell_count, ncell, w.n_valid);
        hull_scan_apply_kernel<<<nchunk, 1024, 0, s>>>(w.cell_count, ncell, w.block_tot);
    }
    hull_occupied_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(w.cell_count, ncell, w.occ, w.n_occ);
    hull_scatter_kernel<<<(unsigned)nb, 256, 0, s>>>(flipped, n, w.key, w.cell_count, w.cursor, w.sorted, vertex_mask);
    // `key` is free once the points are scattered: it becomes the list of the points the near phase hands on
    int r_near = kNearRadius, budget = kNearBudget, r_mid = kMidRadius;
#ifdef COV_HULL_KNOBS
    if (const char* e = getenv("COV_HULL_R_NEAR")) r_near = atoi(e);
    if (const char* e = getenv("COV_HULL_BUDGET")) budget = atoi(e);
    if (const char* e = getenv("COV_HULL_R_MID")) r_mid = atoi(e);
#endif
    const int sms = cov_sm_count_cached();
    hull_classify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(G, w.cell_count, w.sorted, w.rho_max_bits, w.n_valid,
                                                                    w.n_occ, vertex_mask, info, w.key, w.n_mid, r_near, budget);
    // warp per point up to radius kWarpRadius (list: key -> far), block per point beyond it (far -> key), then the all-voxel
    // sweep for what is still open (key)
    hull_local_kernel<32><<<(unsigned)(sms * 8), 128, 0, s>>>(G, w.cell_count, w.sorted, w.rho_max_bits, w.key, w.n_mid,
                                                             vertex_mask, info, w.far, w.n_far, 1, min(kWarpRadius, r_mid));
    hull_local_kernel<256><<<(unsigned)(sms * 4), 256, 0, s>>>(G, w.cell_count, w.sorted, w.rho_max_bits, w.far, w.n_far,
                                                              vertex_mask, info, w.key, w.n_far2, min(kWarpRadius, r_mid) + 1, r_mid);
    hull_far_kernel<<<(unsigned)(sms * 4), kFarThreads, 0, s>>>(G, w.cell_count, w.sorted, w.occ, w.n_occ, w.rho_max_bits,
                                                               w.key, w.n_far2, vertex_mask, info);
    hull_origin_kernel<<<kOriginCtas, 1024, 0, s>>>(flipped, n, info);
#ifdef COV_HULL_KNOBS
    cudaMemcpyAsync(info + 3, w.n_mid, sizeof(int), cudaMemcpyDeviceToDevice, s);  // probe builds report the list lengths
    cudaMemcpyAsync(info + 2, w.n_far, sizeof(int), cudaMemcpyDeviceToDevice, s);   // handed to the block stage
#endif
    return cov_check_launch("cov_hpr_hull");
}
