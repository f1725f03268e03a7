// cov_hull.cu — placeholder until the GPU hull lands; fails loudly (no CPU fallback).
#include "cov_common.cuh"
#include "../../include/coverage_b200.h"
extern "C" size_t cov_hpr_hull_workspace_bytes(int64_t n) { return (size_t)(n > 0 ? n : 1) * 64; }
extern "C" int cov_hpr_hull(const float*, int64_t, uint8_t*, int32_t*, void*, size_t, void*) {
    cov_set_error("cov_hpr_hull: not implemented yet");
    return COV_ERR_UNSUPPORTED;
}
