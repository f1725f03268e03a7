// CPU harness for trajectory_optimization_b200/csrc/hull_core.h (the per-point extremeness test the CUDA
// hull kernel runs): builds the direction grid sequentially, classifies every point, runs GJK for the
// origin, and writes the vertex mask.  Used by tests/test_hull_core_cpu.py to check the algorithm against
// the Qhull fixtures without a GPU.   usage: hull_host <points.f32> <n> <mask_out.u8>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../trajectory_optimization_b200/csrc/hull_core.h"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int n = atoi(argv[2]);
    std::vector<float> pts((size_t)n * 3);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(pts.data(), 4, pts.size(), f) != pts.size()) return 3;
    fclose(f);
    int G = (int)lround(sqrt((double)n / (12.0 * 3.141592653589793)));
    if (argc > 4) G = atoi(argv[4]);
    G = std::max(1, std::min(G, 128));
    const size_t ncell = (size_t)G * G * G;
    std::vector<int> key(n), start(ncell + 1, 0), cursor(ncell, 0), occ;
    double rho_max = 0;
    for (int i = 0; i < n; ++i) {
        const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        const double rho = sqrt(x * x + y * y + z * z);
        rho_max = std::max(rho_max, rho);
        key[i] = (hull_cell_coord(x / rho, G) * G + hull_cell_coord(y / rho, G)) * G + hull_cell_coord(z / rho, G);
        start[key[i] + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) { if (start[c + 1]) occ.push_back((int)c); start[c + 1] += start[c]; }
    std::vector<float4> sorted(n);
    for (int i = 0; i < n; ++i) {
        const int pos = start[key[i]] + cursor[key[i]]++;
        union { float f; int k; } v; v.k = i;
        sorted[pos] = float4{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], v.f};
    }
    HullGrid g{G, 2.0 / G, start.data(), sorted.data(), occ.data(), (int)occ.size(), rho_max};
    std::vector<unsigned char> mask(n, 0);
    long counts[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n; ++k) {
        int cert[3];
        const int rc = hull_classify_point(g, k, cert);
        counts[rc]++;
        mask[hull_float_as_int(sorted[k].w)] = (rc == HULL_EXTREME || rc == HULL_EXTREME_UNCERT) ? 1 : 0;  // an LP that ran away (OVERFLOW) had no feasible region in reach
    }
    // origin: GJK with a sequential support search
    HullSimplex S;
    S.n = 1;
    for (int c = 0; c < 3; ++c) S.v[0][c] = S.x[c] = pts[c];
    int origin_vertex = -1, cert_ok = 0;
    for (int it = 0; it < 64 && origin_vertex < 0; ++it) {
        double best = 1e300; int bj = -1;
        for (int j = 0; j < n; ++j) {
            const double q[3] = {pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]};
            const double d = hull_dot3(S.x, q);
            if (d < best) { best = d; bj = j; }
        }
        const double xx = hull_dot3(S.x, S.x);
        if (best >= xx * (1.0 - 1e-10)) { origin_vertex = 1; cert_ok = best > 1e-9 * xx; break; }
        for (int c = 0; c < 3; ++c) S.v[S.n][c] = pts[3 * bj + c];
        S.n++;
        if (hull_simplex_update(S)) { origin_vertex = 0; cert_ok = hull_certify_origin_inside(S); }
    }
    printf("G=%d n_occ=%zu extreme=%ld inside=%ld inside_uncert=%ld overflow=%ld extreme_uncert=%ld origin_vertex=%d origin_cert=%d\n", G,
           occ.size(), counts[HULL_EXTREME], counts[HULL_INSIDE], counts[HULL_INSIDE_UNCERT], counts[HULL_OVERFLOW], counts[HULL_EXTREME_UNCERT],
           origin_vertex, cert_ok);
    f = fopen(argv[3], "wb");
    fwrite(mask.data(), 1, n, f);
    fclose(f);
    return 0;
}
