// hull_core.h — per-point extremeness test for the Katz hull stage, host + device.
//
// Problem (reference src/tools.py:56-64,79): which of the flipped points f_i are vertices of
// conv({f_j} U {0})?  The reference asks Qhull; on non-degenerate input its vertex set is the exact
// set of extreme points, which is what is computed here, point by point and in parallel:
//
//   f_i is extreme  <=>  exists n with n.(f_i - s) > 0 for every other point s and for s = 0.
//
// n.f_i > 0 (the origin constraint) lets us normalise n = u + alpha e1 + beta e2 with u = f_i/|f_i|
// and (e1,e2) an orthonormal tangent basis, so each other point contributes a half-plane
//   a alpha + b beta + c > 0,  (a,b,c) = (e1.d, e2.d, u.d),  d = f_i - s
// and the question is a 2-D feasibility problem.  We keep the minimum-norm feasible (alpha,beta)
// (least tilt from the radial direction) with a cutting-plane variant of Seidel's incremental LP:
// only constraints that were ever violated enter the active set; adding one costs O(|active|).
//
// Decisions carry certificates:
//   EXTREME      a direction n whose margin against every candidate exceeds the fp64 rounding bound
//                (constraints are shrunk by HULL_EPS), plus a coverage bound showing that points outside
//                the searched neighbourhood cannot reach the supporting plane;
//   NOT EXTREME  three points s1,s2,s3 with f_i in conv(0,s1,s2,s3), checked with four 3x3 determinants
//                whose signs are validated against a forward error bound.
// A decision whose certificate cannot be validated in fp64 is still returned (from the LP) but counted
// as "uncertified" so callers can see that the input was degenerate at fp64 resolution.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define HULL_HD __host__ __device__ __forceinline__
#else
#define HULL_HD inline
#endif

#define HULL_MAX_ACTIVE 64
#define HULL_EPS 4.0e-15      /* half-planes are shrunk by HULL_EPS (1 + 2 tilt) |d|_1: ~4x the fp64 rounding of n.d */
#define HULL_TILT_NEAR 1.0    /* first attempt: |alpha|,|beta| <= 1 (tilt below 55 degrees), tight rounding margin */
#define HULL_TILT_MAX 64.0    /* second attempt (silhouette points): tilt below 89.1 degrees */
#define HULL_BOX 1.0e6        /* bounding box of the LP itself */

enum { HULL_UNDECIDED = 0, HULL_EXTREME = 1, HULL_INSIDE = 2, HULL_INSIDE_UNCERT = 3, HULL_OVERFLOW = 4, HULL_EXTREME_UNCERT = 5,
       HULL_NOCHANGE = 6 };

struct HullFrame {  // local frame of the point under test
    double p[3], rho, u[3], e1[3], e2[3];
};

struct HullLP {
    double tilt;                    // box |alpha|,|beta| <= tilt the rounding margin was computed for
    double x0, x1;                  // current least-tilt feasible (alpha, beta)
    int n;                          // active constraints
    int cert[3];                    // on infeasibility: candidate ids of the blocking constraints
    double a[HULL_MAX_ACTIVE], b[HULL_MAX_ACTIVE], c[HULL_MAX_ACTIVE];
    int id[HULL_MAX_ACTIVE];
};

HULL_HD void hull_frame_init(HullFrame& F, double px, double py, double pz) {
    F.p[0] = px; F.p[1] = py; F.p[2] = pz;
    F.rho = sqrt(px * px + py * py + pz * pz);
    F.u[0] = px / F.rho; F.u[1] = py / F.rho; F.u[2] = pz / F.rho;
    // tangent basis: cross u with the axis it is least aligned with
    const double ax = fabs(F.u[0]), ay = fabs(F.u[1]), az = fabs(F.u[2]);
    double t[3] = {0, 0, 0};
    if (ax <= ay && ax <= az) t[0] = 1; else if (ay <= az) t[1] = 1; else t[2] = 1;
    double e[3] = {F.u[1] * t[2] - F.u[2] * t[1], F.u[2] * t[0] - F.u[0] * t[2], F.u[0] * t[1] - F.u[1] * t[0]};
    const double en = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    F.e1[0] = e[0] / en; F.e1[1] = e[1] / en; F.e1[2] = e[2] / en;
    F.e2[0] = F.u[1] * F.e1[2] - F.u[2] * F.e1[1];
    F.e2[1] = F.u[2] * F.e1[0] - F.u[0] * F.e1[2];
    F.e2[2] = F.u[0] * F.e1[1] - F.u[1] * F.e1[0];
}

HULL_HD void hull_lp_init(HullLP& L, double tilt) {
    L.tilt = tilt;
    L.x0 = 0.0; L.x1 = 0.0; L.n = 0;
    L.cert[0] = L.cert[1] = L.cert[2] = -1;
}

// Half-plane of candidate s in the frame of p, shrunk by the rounding margin.
HULL_HD void hull_constraint(const HullFrame& F, double tilt, double sx, double sy, double sz, double& a, double& b,
                             double& c) {
    const double dx = F.p[0] - sx, dy = F.p[1] - sy, dz = F.p[2] - sz;
    a = F.e1[0] * dx + F.e1[1] * dy + F.e1[2] * dz;
    b = F.e2[0] * dx + F.e2[1] * dy + F.e2[2] * dz;
    c = F.u[0] * dx + F.u[1] * dy + F.u[2] * dz;
    c -= HULL_EPS * (fabs(dx) + fabs(dy) + fabs(dz)) * (1.0 + 2.0 * tilt);
}

// Violated beyond evaluation noise?  The half-planes are shrunk by HULL_EPS >> this tolerance, so a constraint
// accepted here still holds strictly for the true geometry; the tolerance keeps an active (tight) constraint from
// being re-added when the optimum sits on its boundary line.
HULL_HD bool hull_violated(const HullLP& L, double a, double b, double c) {
    const double v = a * L.x0 + b * L.x1 + c;
    return v < -1.0e-15 * (fabs(a * L.x0) + fabs(b * L.x1) + fabs(c));
}

// One constraint g s + h > 0 on the line parameter s: tightens lo (g > 0: s > -h/g), hi (g < 0: s < h/(-g)) or empties the
// interval (g == 0 > h).  Bounds are fractions with positive denominators, compared by cross-multiplication.
#define HULL_LP_BOUND(G_, H_, ID_)                                                                    \
    {                                                                                                  \
        const double g_ = (G_), h_ = (H_);                                                             \
        if (g_ > 0.0) { if (-h_ * lo_d > lo_n * g_) { lo_n = -h_; lo_d = g_; ilo = (ID_); } }          \
        else if (g_ < 0.0) { if (h_ * hi_d < hi_n * -g_) { hi_n = h_; hi_d = -g_; ihi = (ID_); } }     \
        else if (h_ < 0.0) { empty = true; ilo = (ID_); ihi = (ID_); }                                 \
    }

// Add a violated half-plane (a,b,c) [candidate id cid].  Returns HULL_UNDECIDED (new optimum stored),
// HULL_INSIDE (active set infeasible, L.cert filled) or HULL_OVERFLOW (active set full / tilt box hit).
HULL_HD int hull_lp_add(HullLP& L, double a, double b, double c, int cid) {
    for (int k = 0; k < L.n; ++k)
        if (L.id[k] == cid) return HULL_NOCHANGE;  // already active: the residual is evaluation noise of the optimum
    const double nn = a * a + b * b;
    if (!(nn > 0.0)) {  // degenerate direction: constraint is c >= 0 for every (alpha,beta) and it is violated
        L.cert[0] = cid; L.cert[1] = cid; L.cert[2] = cid;
        return HULL_INSIDE;
    }
    if (L.n >= HULL_MAX_ACTIVE) {
        // active set full: keep only the binding half-planes.  They alone determine the current optimum (KKT), so
        // the invariant "x is the least-norm point of the active set" survives; a dropped constraint that the new
        // optimum violates is simply found again by the next sweep, and |x| grows with every addition (no cycling).
        int m = 0;
        for (int k = 0; k < L.n; ++k) {
            const double v = L.a[k] * L.x0 + L.b[k] * L.x1 + L.c[k];
            // binding ones, plus the most recent third (the ones most likely to block the next move)
            if (k >= L.n - HULL_MAX_ACTIVE / 3 || v <= 1.0e-9 * (fabs(L.a[k] * L.x0) + fabs(L.b[k] * L.x1) + fabs(L.c[k]))) {
                L.a[m] = L.a[k]; L.b[m] = L.b[k]; L.c[m] = L.c[k]; L.id[m] = L.id[k]; ++m;
            }
        }
        L.n = m;
        if (L.n >= HULL_MAX_ACTIVE) return HULL_OVERFLOW;
    }
    // the new optimum lies on the line a x + b y + c = 0: q + s t, q = closest point to the origin
    const double inv = 1.0 / nn, rn = sqrt(inv);
    const double q0 = -c * a * inv, q1 = -c * b * inv, t0 = -b * rn, t1 = a * rn;
    // The feasible interval [lo, hi] of the line parameter.  Each constraint bounds it by -h/g; the running bounds are
    // kept as fractions with positive denominators and compared by cross-multiplication, so the loop has no division
    // (an fp64 division is ~30 instructions on the GPU and this loop is most of what the near phase executes).
    double lo_n = -1e300, lo_d = 1.0, hi_n = 1e300, hi_d = 1.0;
    int ilo = -1, ihi = -1;
    bool empty = false;
    // LP bounding box |alpha|, |beta| <= HULL_BOX (ids -2: hitting it is not a proof of infeasibility -> HULL_OVERFLOW),
    // in straight-line code.  It is much wider than the tilt the rounding margin covers; a solution beyond L.tilt is
    // re-done with the wider margin by the caller.
    HULL_LP_BOUND(t0, q0 + HULL_BOX, -2)
    HULL_LP_BOUND(-t0, HULL_BOX - q0, -2)
    HULL_LP_BOUND(t1, q1 + HULL_BOX, -2)
    HULL_LP_BOUND(-t1, HULL_BOX - q1, -2)
    for (int k = 0; k < L.n && !empty; ++k) {
        const double ai = L.a[k], bi = L.b[k];
        HULL_LP_BOUND(ai * t0 + bi * t1, ai * q0 + bi * q1 + L.c[k], L.id[k])
    }
    const double lo = empty ? 1e300 : lo_n / lo_d, hi = empty ? -1e300 : hi_n / hi_d;
    if (lo > hi) {
        if (ilo == -2 || ihi == -2) return HULL_OVERFLOW;
        L.cert[0] = cid; L.cert[1] = ilo; L.cert[2] = ihi;
        return HULL_INSIDE;
    }
    L.a[L.n] = a; L.b[L.n] = b; L.c[L.n] = c; L.id[L.n] = cid; ++L.n;
    const double s = lo > 0.0 ? lo : (hi < 0.0 ? hi : 0.0);
    L.x0 = q0 + s * t0;
    L.x1 = q1 + s * t1;
    return HULL_UNDECIDED;
}

// ---- double-double arithmetic for the determinant signs the fp64 filter cannot settle ----
// (hi, lo) with hi = fl(hi + lo); products by an error-free fma split, sums by the two-sum.  Relative error of a
// dd product or sum <= 2^-104.
struct HullDD { double hi, lo; };
HULL_HD HullDD hull_two_prod(double a, double b) {
    HullDD r;
    r.hi = a * b;
    r.lo = fma(a, b, -r.hi);
    return r;
}
HULL_HD HullDD hull_two_sum(double a, double b) {
    HullDD r;
    r.hi = a + b;
    const double bb = r.hi - a;
    r.lo = (a - (r.hi - bb)) + (b - bb);
    return r;
}
HULL_HD HullDD hull_dd_add(HullDD a, HullDD b) {
    HullDD s = hull_two_sum(a.hi, b.hi);
    s.lo += a.lo + b.lo;
    return hull_two_sum(s.hi, s.lo);
}
HULL_HD HullDD hull_dd_neg(HullDD a) { HullDD r; r.hi = -a.hi; r.lo = -a.lo; return r; }
HULL_HD HullDD hull_dd_mul_d(HullDD a, double b) {
    HullDD p = hull_two_prod(a.hi, b);
    p.lo += a.lo * b;
    return hull_two_sum(p.hi, p.lo);
}
// b_i c_j - b_k c_l in double-double (each product is error-free)
HULL_HD HullDD hull_dd_minor(double bi, double cj, double bk, double cl) {
    return hull_dd_add(hull_two_prod(bi, cj), hull_dd_neg(hull_two_prod(bk, cl)));
}

// sign of det[a;b;c]: +1 / -1 when certain, 0 when not.  Stage 1 is an fp64 evaluation with a forward error bound;
// when that is inconclusive the determinant is recomputed in double-double (error < 2^-96 of the permanent, i.e.
// 13 orders of magnitude finer): what still comes out as 0 is degenerate at ~100 bits — for fp32-valued input that
// means exactly coplanar with the origin (duplicate points, points on a common plane through 0).
HULL_HD int hull_det_sign(const double* a, const double* b, const double* c) {
    const double m0 = b[1] * c[2] - b[2] * c[1], m1 = b[2] * c[0] - b[0] * c[2], m2 = b[0] * c[1] - b[1] * c[0];
    const double det = a[0] * m0 + a[1] * m1 + a[2] * m2;
    const double perm = fabs(a[0]) * (fabs(b[1] * c[2]) + fabs(b[2] * c[1])) +
                        fabs(a[1]) * (fabs(b[2] * c[0]) + fabs(b[0] * c[2])) +
                        fabs(a[2]) * (fabs(b[0] * c[1]) + fabs(b[1] * c[0]));
    const double bound = 1.0e-15 * perm;  // > 8 * 2^-53 * perm
    if (det > bound) return 1;
    if (det < -bound) return -1;
    if (!(perm > 0.0)) return 0;
    const HullDD d = hull_dd_add(hull_dd_add(hull_dd_mul_d(hull_dd_minor(b[1], c[2], b[2], c[1]), a[0]),
                                             hull_dd_mul_d(hull_dd_minor(b[2], c[0], b[0], c[2]), a[1])),
                                 hull_dd_mul_d(hull_dd_minor(b[0], c[1], b[1], c[0]), a[2]));
    const double bound2 = 1.3e-29 * perm;  // 2^-96 * perm, far above the ~12 * 2^-104 * perm the dd evaluation can lose
    return d.hi > bound2 ? 1 : (d.hi < -bound2 ? -1 : 0);
}

// Is p inside conv(0, s1, s2, s3)?  1 = certified inside (closed, non-degenerate), 0 = cannot certify.
HULL_HD int hull_certify_inside(const double* p, const double* s1, const double* s2, const double* s3) {
    const int D = hull_det_sign(s1, s2, s3);
    if (D == 0) return 0;
    const int D1 = hull_det_sign(p, s2, s3), D2 = hull_det_sign(s1, p, s3), D3 = hull_det_sign(s1, s2, p);
    const double q1[3] = {s1[0] - p[0], s1[1] - p[1], s1[2] - p[2]};
    const double q2[3] = {s2[0] - p[0], s2[1] - p[1], s2[2] - p[2]};
    const double q3[3] = {s3[0] - p[0], s3[1] - p[1], s3[2] - p[2]};
    const int D0 = hull_det_sign(q1, q2, q3);  // sign of D - D1 - D2 - D3 (differences of fp32 values: see bound)
    return (D1 == D && D2 == D && D3 == D && D0 == D) ? 1 : 0;
}

// After a clean sweep over every point j with |u_j - u_i| <= Dcov: can a point outside that neighbourhood
// still reach the supporting plane of n = u + x0 e1 + x1 e2 ?  (rho_max = largest norm in the cloud.)
// For a point at angle gamma from u:  n.f_j <= rho_max (cos gamma + |tau| sin gamma),  n.f_i = rho_i.
HULL_HD bool hull_coverage_ok(const HullLP& L, double rho_i, double rho_max, double Dcov) {
    if (Dcov >= 2.0) return true;  // the whole sphere was searched
    const double tau = sqrt(L.x0 * L.x0 + L.x1 * L.x1);
    const double cD = 1.0 - 0.5 * Dcov * Dcov;
    const double sD = sqrt(fmax(0.0, 1.0 - cD * cD));
    if (!(sD >= tau * fabs(cD)) && cD > 0.0) return false;  // maximum of the bound lies outside the searched cap
    if (cD <= 0.0) return false;                             // beyond a hemisphere: let the caller search everything
    return rho_max * (cD + tau * sD) * (1.0 + 1e-12) < rho_i;
}

// ------------------------------------------------------------------------------------------
// Direction grid: unit directions u = f/|f| binned into G^3 voxels of [-1,1]^3 (only the voxels
// cut by the unit sphere are occupied).  Points are counting-sorted by voxel; `sorted` holds
// (x, y, z, original index as int bits) so a sweep over a voxel is a contiguous read.
// ------------------------------------------------------------------------------------------
#ifndef __CUDACC__
struct float4 { float x, y, z, w; };
#endif

struct HullGrid {
    int G;                       // voxels per axis
    double h;                    // voxel edge = 2/G
    const int* cell_start;       // G^3 + 1 exclusive offsets into `sorted`
    const float4* sorted;        // n_sorted entries
    const int* occ;              // occupied voxel ids
    int n_occ;
    double rho_max;              // largest |f| in the cloud
};

HULL_HD int hull_cell_coord(double u, int G) {
    int c = (int)floor((u + 1.0) * 0.5 * G);
    return c < 0 ? 0 : (c >= G ? G - 1 : c);
}

HULL_HD int hull_float_as_int(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    union { float f; int i; } v; v.f = f; return v.i;
#endif
}

// Sweep the points of one voxel against the current LP.  Returns a decision code or HULL_UNDECIDED;
// sets `changed` when the LP moved.
HULL_HD int hull_sweep_cell(const HullGrid& g, int cell, int self, const HullFrame& F, HullLP& L, bool& changed) {
    const int b = g.cell_start[cell], e = g.cell_start[cell + 1];
    for (int k = b; k < e; ++k) {
        if (k == self) continue;
        const float4 s = g.sorted[k];
        double a, bb, c;
        hull_constraint(F, L.tilt, (double)s.x, (double)s.y, (double)s.z, a, bb, c);
        if (hull_violated(L, a, bb, c)) {
            const int rc = hull_lp_add(L, a, bb, c, k);
            if (rc == HULL_NOCHANGE) continue;
            if (rc != HULL_UNDECIDED) return rc;
            changed = true;
        }
    }
    return HULL_UNDECIDED;
}

#define HULL_R_NEAR 6

// Classify sorted point `self`.  cert_out receives the three sorted-array ids of an INSIDE certificate.
// far_ok = false: stop after the near phase and return HULL_UNDECIDED when the point needs the all-voxel sweep (the
// CUDA path hands those points to a kernel that runs that sweep with a whole warp per point).
// r_near / budget (far_ok = false only): give up the near phase beyond Chebyshev radius r_near or after `budget` constraint
// evaluations (0 = no limit) — the point is deferred, not decided, so the answer does not depend on either.
HULL_HD int hull_classify_attempt(const HullGrid& g, int self, double tilt, int* cert_out, bool far_ok = true,
                                  int r_near = HULL_R_NEAR, int budget = 0) {
    const float4 ps = g.sorted[self];
    HullFrame F;
    hull_frame_init(F, (double)ps.x, (double)ps.y, (double)ps.z);
    HullLP L;
    hull_lp_init(L, tilt);
    const int G = g.G;
    const int cx = hull_cell_coord(F.u[0], G), cy = hull_cell_coord(F.u[1], G), cz = hull_cell_coord(F.u[2], G);
    int clean = -1;  // every voxel within Chebyshev radius `clean` has been swept without moving the LP
    int rc = HULL_UNDECIDED;
    int work = 0;
    if (far_ok) { r_near = HULL_R_NEAR; budget = 0; }
    for (int r = 0; r <= r_near && rc == HULL_UNDECIDED;) {
        bool changed = false;
        for (int dx = -r; dx <= r && rc == HULL_UNDECIDED; ++dx) {
            const int ix = cx + dx;
            if (ix < 0 || ix >= G) continue;
            for (int dy = -r; dy <= r && rc == HULL_UNDECIDED; ++dy) {
                const int iy = cy + dy;
                if (iy < 0 || iy >= G) continue;
                for (int dz = -r; dz <= r; ++dz) {
                    const int iz = cz + dz;
                    if (iz < 0 || iz >= G) continue;
                    const int ad = (dx < 0 ? -dx : dx), bd = (dy < 0 ? -dy : dy), cd = (dz < 0 ? -dz : dz);
                    const int cheb = ad > bd ? (ad > cd ? ad : cd) : (bd > cd ? bd : cd);
                    if (cheb <= clean) continue;
                    const int cell = (ix * G + iy) * G + iz;
                    if (budget > 0) {
                        work += g.cell_start[cell + 1] - g.cell_start[cell];
                        if (work > budget) return HULL_UNDECIDED;
                    }
                    rc = hull_sweep_cell(g, cell, self, F, L, changed);
                    if (rc != HULL_UNDECIDED) break;
                }
            }
        }
        if (rc != HULL_UNDECIDED) break;
        if (changed) { clean = -1; continue; }  // the LP moved: everything swept so far must be re-checked
        clean = r;
        const bool whole = (cx - r <= 0 && cx + r >= G - 1 && cy - r <= 0 && cy + r >= G - 1 && cz - r <= 0 && cz + r >= G - 1);
        if (whole || hull_coverage_ok(L, F.rho, g.rho_max, r * g.h)) { rc = HULL_EXTREME; break; }
        ++r;
    }
    if (rc == HULL_UNDECIDED && !far_ok) return HULL_UNDECIDED;
    // far phase (large tilt, e.g. silhouette points of a half-space cloud): every occupied voxel, culled by a
    // bound on n.f over the voxel: rho_max (n.c + |n| h sqrt(3)/2), c = voxel centre.
    while (rc == HULL_UNDECIDED) {
        bool changed = false;
        const double n0 = F.u[0] + L.x0 * F.e1[0] + L.x1 * F.e2[0];
        const double n1 = F.u[1] + L.x0 * F.e1[1] + L.x1 * F.e2[1];
        const double n2 = F.u[2] + L.x0 * F.e1[2] + L.x1 * F.e2[2];
        const double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
        const double np = (n0 * F.p[0] + n1 * F.p[1] + n2 * F.p[2]) * (1.0 - 1e-12);
        const double slack = nn * g.h * 0.8660254037844387;
        for (int k = 0; k < g.n_occ && rc == HULL_UNDECIDED; ++k) {
            const int cell = g.occ[k];
            const int iz = cell % G, iy = (cell / G) % G, ix = cell / (G * G);
            const double c0 = (ix + 0.5) * g.h - 1.0, c1 = (iy + 0.5) * g.h - 1.0, c2 = (iz + 0.5) * g.h - 1.0;
            if (g.rho_max * (n0 * c0 + n1 * c1 + n2 * c2 + slack) < np) continue;
            rc = hull_sweep_cell(g, cell, self, F, L, changed);
            if (changed) break;  // the direction moved: restart the culled sweep with the new n
        }
        if (rc == HULL_UNDECIDED && !changed) rc = HULL_EXTREME;
    }
    if (rc == HULL_EXTREME && (fabs(L.x0) > L.tilt || fabs(L.x1) > L.tilt)) rc = HULL_EXTREME_UNCERT;  // margin not valid
    if (rc == HULL_INSIDE) {
        cert_out[0] = L.cert[0]; cert_out[1] = L.cert[1]; cert_out[2] = L.cert[2];
        const float4 s1 = g.sorted[L.cert[0]], s2 = g.sorted[L.cert[1]], s3 = g.sorted[L.cert[2]];
        const double a1[3] = {s1.x, s1.y, s1.z}, a2[3] = {s2.x, s2.y, s2.z}, a3[3] = {s3.x, s3.y, s3.z};
        if (L.cert[0] == L.cert[1] || L.cert[1] == L.cert[2] || L.cert[0] == L.cert[2] ||
            !hull_certify_inside(F.p, a1, a2, a3))
            rc = HULL_INSIDE_UNCERT;
    }
    return rc;
}

// Two attempts: a tight tilt box (and rounding margin) that settles everything but silhouette points, then the wide one.
HULL_HD int hull_classify_point(const HullGrid& g, int self, int* cert_out) {
    int rc = hull_classify_attempt(g, self, HULL_TILT_NEAR, cert_out);
    if (rc == HULL_EXTREME_UNCERT || rc == HULL_OVERFLOW) rc = hull_classify_attempt(g, self, HULL_TILT_MAX, cert_out);
    return rc;
}

// ------------------------------------------------------------------------------------------
// Is the origin a vertex of conv(F U {0}), i.e. is 0 outside conv(F)?  GJK on the point set:
// the simplex update below (closest point of conv(S) to the origin, |S| <= 4) is shared by the
// host harness and the device kernel; the support search over all points is the caller's.
// ------------------------------------------------------------------------------------------
struct HullSimplex {
    double v[4][3];
    int n;
    double x[3];  // closest point of conv(v[0..n)) to the origin
};

HULL_HD double hull_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Recompute x and shrink the simplex to the face that supports it.  Returns 1 when the origin is inside a
// tetrahedron of the simplex (x = 0).  Brute force over the 15 faces: tiny and branch-simple.
HULL_HD int hull_simplex_update(HullSimplex& S) {
    double best = 1e300;
    int best_mask = 0;
    double best_x[3] = {0, 0, 0};
    const int n = S.n;
    for (int mask = 1; mask < (1 << n); ++mask) {
        int idx[4], m = 0;
        for (int i = 0; i < n; ++i) if (mask & (1 << i)) idx[m++] = i;
        double lam[4] = {0, 0, 0, 0};
        bool ok = true;
        if (m == 1) lam[0] = 1.0;
        else {
            // minimise |sum lam_k v_k|^2 s.t. sum lam = 1: with e_k = v_k - v_0, solve (E^T E) mu = -E^T v_0
            const double* v0 = S.v[idx[0]];
            double e[3][3], Gm[3][3], rhs[3];
            for (int k = 1; k < m; ++k) for (int c = 0; c < 3; ++c) e[k - 1][c] = S.v[idx[k]][c] - v0[c];
            const int d = m - 1;
            for (int r = 0; r < d; ++r) { rhs[r] = -hull_dot3(e[r], v0); for (int c = 0; c < d; ++c) Gm[r][c] = hull_dot3(e[r], e[c]); }
            double mu[3] = {0, 0, 0};
            if (d == 1) { if (!(Gm[0][0] > 0)) ok = false; else mu[0] = rhs[0] / Gm[0][0]; }
            else if (d == 2) {
                const double det = Gm[0][0] * Gm[1][1] - Gm[0][1] * Gm[1][0];
                if (!(fabs(det) > 0)) ok = false;
                else { mu[0] = (rhs[0] * Gm[1][1] - rhs[1] * Gm[0][1]) / det; mu[1] = (Gm[0][0] * rhs[1] - Gm[1][0] * rhs[0]) / det; }
            } else {
                const double c00 = Gm[1][1] * Gm[2][2] - Gm[1][2] * Gm[2][1], c01 = Gm[1][2] * Gm[2][0] - Gm[1][0] * Gm[2][2],
                             c02 = Gm[1][0] * Gm[2][1] - Gm[1][1] * Gm[2][0];
                const double det = Gm[0][0] * c00 + Gm[0][1] * c01 + Gm[0][2] * c02;
                if (!(fabs(det) > 0)) ok = false;
                else {
                    mu[0] = (rhs[0] * c00 + rhs[1] * (Gm[0][2] * Gm[2][1] - Gm[0][1] * Gm[2][2]) + rhs[2] * (Gm[0][1] * Gm[1][2] - Gm[0][2] * Gm[1][1])) / det;
                    mu[1] = (rhs[0] * c01 + rhs[1] * (Gm[0][0] * Gm[2][2] - Gm[0][2] * Gm[2][0]) + rhs[2] * (Gm[0][2] * Gm[1][0] - Gm[0][0] * Gm[1][2])) / det;
                    mu[2] = (rhs[0] * c02 + rhs[1] * (Gm[0][1] * Gm[2][0] - Gm[0][0] * Gm[2][1]) + rhs[2] * (Gm[0][0] * Gm[1][1] - Gm[0][1] * Gm[1][0])) / det;
                }
            }
            if (ok) {
                double s = 0;
                for (int k = 1; k < m; ++k) { lam[k] = mu[k - 1]; s += mu[k - 1]; }
                lam[0] = 1.0 - s;
                for (int k = 0; k < m; ++k) if (!(lam[k] >= 0.0)) ok = false;
            }
        }
        if (!ok) continue;
        double x[3] = {0, 0, 0};
        for (int k = 0; k < m; ++k) for (int c = 0; c < 3; ++c) x[c] += lam[k] * S.v[idx[k]][c];
        const double d2 = hull_dot3(x, x);
        if (d2 < best) { best = d2; best_mask = mask; best_x[0] = x[0]; best_x[1] = x[1]; best_x[2] = x[2]; }
    }
    // keep only the supporting face
    HullSimplex T;
    T.n = 0;
    for (int i = 0; i < n; ++i)
        if (best_mask & (1 << i)) { for (int c = 0; c < 3; ++c) T.v[T.n][c] = S.v[i][c]; ++T.n; }
    for (int c = 0; c < 3; ++c) T.x[c] = best_x[c];
    S = T;
    return (S.n == 4) ? 1 : 0;  // a full tetrahedron supports x only when x = 0 lies inside it
}

// 0 strictly inside the tetrahedron (v0..v3)?  1 = certified.
HULL_HD int hull_certify_origin_inside(const HullSimplex& S) {
    if (S.n != 4) return 0;
    const int d0 = hull_det_sign(S.v[1], S.v[2], S.v[3]);
    const int d1 = -hull_det_sign(S.v[0], S.v[2], S.v[3]);
    const int d2 = hull_det_sign(S.v[0], S.v[1], S.v[3]);
    const int d3 = -hull_det_sign(S.v[0], S.v[1], S.v[2]);
    return (d0 != 0 && d0 == d1 && d1 == d2 && d2 == d3) ? 1 : 0;
}
