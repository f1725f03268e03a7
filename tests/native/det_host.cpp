// Host harness for hull_core.h's determinant sign (tests/test_hull_core_cpu.py): reads triples of 3-vectors (27... 9
// doubles per case) from stdin and prints the certified sign per case.
#include <cstdio>

#include "../../trajectory_optimization_b200/csrc/hull_core.h"

int main() {
    double v[9];
    while (std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) == 9)
        std::printf("%d\n", hull_det_sign(v, v + 3, v + 6));
    return 0;
}
