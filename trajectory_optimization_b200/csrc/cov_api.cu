// cov_api.cu — host-side plumbing of the C ABI: error string, launch checks, constants, probes.
#include <cstdarg>
#include <cstdio>
#include <cmath>

#include "cov_common.cuh"
#include "../../include/coverage_b200.h"

static thread_local char g_err[512] = "";

void cov_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cov_check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cov_set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return COV_ERR_CUDA;
    }
    return COV_OK;
}

CovConst cov_make_const(const cov_camera* cam) {
    CovConst C;
    const double mu = ((double)cam->min_dist + (double)cam->max_dist) / 2.0;
    const double sigma = ((double)cam->max_dist - (double)cam->min_dist) / 2.0;
    const double log2e = 1.4426950408889634;
    const double rkf = std::sqrt(0.5 * log2e);
    C.eps = cam->eps;
    C.kd = (float)(0.5 * log2e / (sigma * sigma));
    C.zk = (float)(-log2e);
    C.zc = (float)(log2e * (double)cam->eps);
    C.c0 = (float)(-0.5 * rkf);
    C.cw = (float)(rkf / (double)cam->img_width);
    C.ch = (float)(rkf / (double)cam->img_height);
    C.inv_s2 = (float)(1.0 / (sigma * sigma));
    C.mu = (float)mu;
    C.hi = (float)(1.0 - (double)cam->eps);
    return C;
}

int cov_sm_count_cached() {
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != dev_cached) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms = v;
        dev_cached = dev;
    }
    return sms;
}

extern "C" int cov_version(void) { return 200; }
extern "C" const char* cov_last_error(void) { return g_err; }
extern "C" int cov_device_sm_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 0;
    }
    return cov_sm_count_cached();
}

// ---- roofline probes: FP32 FMA and MUFU.EX2 issue rates ----
namespace {
constexpr int kProbeIlp = 8;
__global__ void __launch_bounds__(256) probe_fma_kernel(int iters, float* sink) {
    float a[kProbeIlp];
#pragma unroll
    for (int i = 0; i < kProbeIlp; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    const float b = 0.999f, c = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kProbeIlp; ++i) a[i] = __fmaf_rn(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kProbeIlp; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}
// packed fp32 pairs: one FFMA2 per instruction = two FMAs per lane (sm_100 fma.rn.f32x2)
__global__ void __launch_bounds__(256) probe_fma2_kernel(int iters, float* sink) {
    unsigned long long a[kProbeIlp];
#pragma unroll
    for (int i = 0; i < kProbeIlp; ++i) {
        const float lo = 1.0f + 1e-3f * (threadIdx.x + i), hi = 1.0f + 2e-3f * (threadIdx.x + i);
        asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
    }
    unsigned long long b, c;
    {
        const float bf = 0.999f, cf = 1e-4f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
        asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(cf));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kProbeIlp; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kProbeIlp; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
        s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) probe_ex2_kernel(int iters, float* sink) {
    float a[kProbeIlp];
#pragma unroll
    for (int i = 0; i < kProbeIlp; ++i) a[i] = -1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kProbeIlp; ++i) a[i] = cov_ex2(a[i]) - 1.0f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kProbeIlp; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}
}  // namespace

extern "C" int64_t cov_probe_fma(int iters, float* sink, void* stream) {
    const int grid = cov_sm_count_cached() * 8;
    probe_fma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (cov_check_launch("cov_probe_fma")) return -1;
    return (int64_t)grid * 256 * kProbeIlp * iters;
}
extern "C" int64_t cov_probe_fma2(int iters, float* sink, void* stream) {
    const int grid = cov_sm_count_cached() * 8;
    probe_fma2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (cov_check_launch("cov_probe_fma2")) return -1;
    return (int64_t)grid * 256 * kProbeIlp * iters * 2;
}
extern "C" int64_t cov_probe_ex2(int iters, float* sink, void* stream) {
    const int grid = cov_sm_count_cached() * 8;
    probe_ex2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (cov_check_launch("cov_probe_ex2")) return -1;
    return (int64_t)grid * 256 * kProbeIlp * iters;
}
