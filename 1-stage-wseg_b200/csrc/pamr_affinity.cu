// Local-affinity kernel: replaces reference models/mods/pamr.py:132-136
//   x_std = LocalStDev(x)            (:77-103; 9*nd samples incl. the centre, unbiased std)
//   a     = -|x - x_nbr| / (1e-8 + 0.1*x_std)   (:134, LocalAffinityAbs :105-109)
//   a     = mean over the K image channels      (:135)
//   w     = softmax over the 8*nd neighbours    (:136)
// One thread per pixel; the per-channel 9*nd samples are read once into registers (template ND)
// and serve both the std and the 8*nd differences.  Output is tap-major [B,P,H,W] so that every
// store of a warp is one coalesced 128-byte line.
//
// Numerics (parity bar 1e-5 on the refined masks): the std is accumulated in double (two-pass),
// as torch's CPU kernel carries Welford in double; the division is IEEE (no fast-math); expf is
// the accurate libdevice one.
#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int AFF_BX = 32;
constexpr int AFF_BY = 4;

// ND > 0: number of dilations known at compile time (arrays stay in registers).
// ND == 0: runtime nd <= PAMR_MAX_DILATIONS (arrays in local memory; generic fallback).
template <int ND>
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
affinity_kernel(const float* __restrict__ img, float* __restrict__ aff, int K, int H, int W, Dilations dil) {
    constexpr int MAXND = (ND > 0) ? ND : PAMR_MAX_DILATIONS;
    const int nd = (ND > 0) ? ND : dil.nd;
    const int x = blockIdx.x * AFF_BX + threadIdx.x;
    const int y = blockIdx.y * AFF_BY + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;
    const int P = 8 * nd;

    float abar[8 * MAXND];
#pragma unroll
    for (int p = 0; p < 8 * MAXND; ++p) abar[p] = 0.f;

    for (int k = 0; k < K; ++k) {
        const float* __restrict__ pl = img + ((size_t)b * K + k) * HW;
        float smp[9 * MAXND];
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < MAXND; ++i) {
            if (i < nd) {
                const int d = dil.d[i];
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const int yy = clampi(y + (j / 3 - 1) * d, 0, H - 1);
                    const int xx = clampi(x + (j % 3 - 1) * d, 0, W - 1);
                    const float v = __ldg(pl + (size_t)yy * W + xx);
                    smp[9 * i + j] = v;
                    sum += (double)v;
                }
            }
        }
        const double mean = sum / (double)(9 * nd);
        double m2 = 0.0;
#pragma unroll
        for (int i = 0; i < MAXND; ++i) {
            if (i < nd) {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const double dv = (double)smp[9 * i + j] - mean;
                    m2 = fma(dv, dv, m2);
                }
            }
        }
        const float sd = (float)sqrt(m2 / (double)(9 * nd - 1));
        const float den = __fadd_rn(1e-8f, __fmul_rn(0.1f, sd));
        const float c = smp[4];
#pragma unroll
        for (int i = 0; i < MAXND; ++i) {
            if (i < nd) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int j9 = (j < 4) ? j : j + 1;  // skip the centre sample
                    const float a = __fdiv_rn(-fabsf(__fsub_rn(c, smp[9 * i + j9])), den);
                    abar[8 * i + j] = (k == 0) ? a : __fadd_rn(abar[8 * i + j], a);
                }
            }
        }
    }

    const float kf = (float)K;
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < 8 * MAXND; ++p) {
        if (p < P) {
            abar[p] = __fdiv_rn(abar[p], kf);
            mx = fmaxf(mx, abar[p]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 8 * MAXND; ++p) {
        if (p < P) {
            abar[p] = expf(abar[p] - mx);
            s += abar[p];
        }
    }
    float* __restrict__ out = aff + (size_t)b * P * HW + (size_t)y * W + x;
#pragma unroll
    for (int p = 0; p < 8 * MAXND; ++p) {
        if (p < P) out[(size_t)p * HW] = __fdiv_rn(abar[p], s);
    }
}

}  // namespace

int launch_affinity(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil, cudaStream_t s) {
    dim3 block(AFF_BX, AFF_BY);
    dim3 grid((W + AFF_BX - 1) / AFF_BX, (H + AFF_BY - 1) / AFF_BY, B);
    if (grid.y > 65535 || grid.z > 65535)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity: H/4 and B must be <= 65535");
    if (dil.nd == 6)
        affinity_kernel<6><<<grid, block, 0, s>>>(img, aff, K, H, W, dil);
    else
        affinity_kernel<0><<<grid, block, 0, s>>>(img, aff, K, H, W, dil);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr
