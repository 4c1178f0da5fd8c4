// Shared device/host helpers for libpamr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pamr_b200.h"

namespace pamr {

// Dilation list passed to kernels by value.
struct Dilations {
    int nd;
    int d[PAMR_MAX_DILATIONS];
};

// 3x3 neighbourhood without the centre, row-major (reference pamr.py:25-34).
__device__ __forceinline__ int tap_dy(int j) { return (j < 3) ? -1 : (j < 5 ? 0 : 1); }
__device__ __forceinline__ int tap_dx(int j) {
    // j: 0 1 2 | 3 4 | 5 6 7  ->  -1 0 1 | -1 1 | -1 0 1
    return (j == 0 || j == 3 || j == 5) ? -1 : ((j == 1 || j == 6) ? 0 : 1);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// Monotone float <-> unsigned map so that atomicMax(unsigned) implements a float max for any sign.
__host__ __device__ __forceinline__ unsigned ordered_from_float(float v) {
#ifdef __CUDA_ARCH__
    unsigned u = __float_as_uint(v);
#else
    union { float f; unsigned u; } c; c.f = v; unsigned u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float float_from_ordered(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}

// ---- host-side error plumbing (defined in pamr_capi.cu) ----
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define PAMR_CUDA_TRY(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::pamr::set_error(PAMR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define PAMR_REQUIRE(cond, ...)                                                                  \
    do {                                                                                         \
        if (!(cond)) return ::pamr::set_error(PAMR_ERR_INVALID_ARGUMENT, __VA_ARGS__);           \
    } while (0)

// Kernel launchers implemented in the .cu files (all enqueue on `s`, return a PAMR_* code).
int launch_resize_bilinear(const float* src, float* dst, int n_planes, int h, int w, int H, int W, cudaStream_t s);
int launch_affinity(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil, cudaStream_t s);
int launch_propagate(const float* aff, const float* m_in, float* m_out, float* m_tmp, int B, int C, int H, int W,
                     const Dilations& dil, int iters, unsigned* cls_max, int dev, cudaStream_t s);
int launch_clean(const float* m, const float* labels, float* cleaned, unsigned* cls_max, int B, int C, int h, int w,
                 int H, int W, cudaStream_t s);
int launch_pseudo_labels(const float* m, const float* labels, const unsigned* cls_max, uint8_t* label,
                         float* pseudo_gt, int* class_count, int B, int C, int h, int w, int H, int W, float bg_cut,
                         float fg_cut, float low_cut, bool max_is_gated, cudaStream_t s);

}  // namespace pamr
