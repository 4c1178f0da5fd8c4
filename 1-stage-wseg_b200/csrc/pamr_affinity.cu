// Local-affinity kernel: replaces reference models/mods/pamr.py:132-136
//   x_std = LocalStDev(x)            (:77-103; 9*nd samples incl. the centre, unbiased std)
//   a     = -|x - x_nbr| / (1e-8 + 0.1*x_std)   (:134, LocalAffinityAbs :105-109)
//   a     = mean over the K image channels      (:135)
//   w     = softmax over the 8*nd neighbours    (:136)
// One thread per pixel; the per-channel 9*nd samples are read once into registers (template ND)
// and serve both the std and the 8*nd differences.  Output is either the reference's tap-major
// [B,P,H,W] (every store of a warp is one coalesced 128-byte line) or, for the tuned propagation
// kernel, the tile-major layout of pamr_common.cuh (also one 128-byte line per store).
//
// Numerics (parity bar 1e-5 on the refined masks): the std is accumulated in double (two-pass),
// as torch's CPU kernel carries Welford in double; the 8*nd*K divisions by the per-channel
// denominator use a correctly rounded reciprocal plus one FMA residual correction (Markstein),
// which returns the IEEE quotient for these operand ranges at a third of the cost of the generic
// division routine; expf is the accurate libdevice one (no fast-math anywhere).
#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int AFF_BX = 32;
constexpr int AFF_BY = 4;

// x / y for y in the normal range, given r = rn(1/y): q = rn(x*r) corrected by the exact residual.
__device__ __forceinline__ float div_markstein(float x, float y, float r) {
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, y, x);
    return __fmaf_rn(e, r, q);
}

// ND > 0: number of dilations known at compile time (arrays stay in registers).
// ND == 0: runtime nd <= PAMR_MAX_DILATIONS (arrays in local memory; generic fallback).
// TILED: write the tile-major layout (requires ND == 6); the grid then covers whole tiles and
// threads outside the image store zeros.
template <int ND, bool TILED>
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
affinity_kernel(const float* __restrict__ img, float* __restrict__ aff, int K, int H, int W, Dilations dil,
                AffTiling tiling) {
    constexpr int MAXND = (ND > 0) ? ND : PAMR_MAX_DILATIONS;
    const int nd = (ND > 0) ? ND : dil.nd;
    const int x = blockIdx.x * AFF_BX + threadIdx.x;
    const int y = blockIdx.y * AFF_BY + threadIdx.y;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int P = 8 * nd;
    if (x >= W || y >= H) {
        if (TILED) {
            // inside a partial tile but outside the image: the propagation kernel expects zeros
            if (x < tiling.tiles_x * 32 && y < tiling.tiles_y * 4 * tiling.R) {
                float* __restrict__ out = aff + aff_tiled_index(tiling, b, 0, y, x);
                const size_t sstride = (size_t)tiling.R * 32;
#pragma unroll
                for (int s = 0; s < 48; ++s) out[s * sstride] = 0.f;
            }
        }
        return;
    }

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
        const float den = __fadd_rn(1e-8f, __fmul_rn(0.1f, sd));  // in [1e-8, ~0.06]: normal range
        const float rden = __frcp_rn(den);
        const float c = smp[4];
#pragma unroll
        for (int i = 0; i < MAXND; ++i) {
            if (i < nd) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int j9 = (j < 4) ? j : j + 1;  // skip the centre sample
                    const float a = div_markstein(-fabsf(__fsub_rn(c, smp[9 * i + j9])), den, rden);
                    abar[8 * i + j] = (k == 0) ? a : __fadd_rn(abar[8 * i + j], a);
                }
            }
        }
    }

    const float kf = (float)K;
    const float rk = __frcp_rn(kf);
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < 8 * MAXND; ++p) {
        if (p < P) {
            abar[p] = div_markstein(abar[p], kf, rk);
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
    const float rs = __frcp_rn(s);  // s in [1, P]
    if (TILED) {
        float* __restrict__ out = aff + aff_tiled_index(tiling, b, 0, y, x);
        const size_t sstride = (size_t)tiling.R * 32;
#pragma unroll
        for (int p = 0; p < 8 * MAXND; ++p) out[tap_seq(p) * sstride] = div_markstein(abar[p], s, rs);
    } else {
        float* __restrict__ out = aff + (size_t)b * P * HW + (size_t)y * W + x;
#pragma unroll
        for (int p = 0; p < 8 * MAXND; ++p) {
            if (p < P) out[(size_t)p * HW] = div_markstein(abar[p], s, rs);
        }
    }
}

// standard [B,48,H,W] -> tile-major (for callers of the public propagate API)
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
aff_relayout_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, AffTiling tiling) {
    const int x = blockIdx.x * AFF_BX + threadIdx.x;
    const int y = blockIdx.y * AFF_BY + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= tiling.tiles_x * 32 || y >= tiling.tiles_y * 4 * tiling.R) return;
    const bool in = (x < W) && (y < H);
    const size_t HW = (size_t)H * W;
    const float* __restrict__ ip = src + (size_t)b * 48 * HW + (size_t)y * W + x;
    float* __restrict__ out = dst + aff_tiled_index(tiling, b, 0, y, x);
    const size_t sstride = (size_t)tiling.R * 32;
#pragma unroll 8
    for (int p = 0; p < 48; ++p) out[tap_seq(p) * sstride] = in ? __ldg(ip + (size_t)p * HW) : 0.f;
}

}  // namespace

int launch_affinity(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil,
                    const AffTiling& tiling, cudaStream_t s) {
    dim3 block(AFF_BX, AFF_BY);
    const bool tiled = tiling.R > 0;
    const int gw = tiled ? tiling.tiles_x * 32 : W, gh = tiled ? tiling.tiles_y * 4 * tiling.R : H;
    dim3 grid((gw + AFF_BX - 1) / AFF_BX, (gh + AFF_BY - 1) / AFF_BY, B);
    if (grid.y > 65535 || grid.z > 65535)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity: H/4 and B must be <= 65535");
    if (tiled) {
        if (dil.nd != 6) return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity: tiled layout needs 6 dilations");
        affinity_kernel<6, true><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
    } else if (dil.nd == 6) {
        affinity_kernel<6, false><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
    } else {
        affinity_kernel<0, false><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
    }
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_aff_relayout(const float* aff_std, float* aff_tiled, int B, int H, int W, const AffTiling& tiling,
                        cudaStream_t s) {
    dim3 block(AFF_BX, AFF_BY);
    dim3 grid(tiling.tiles_x, (tiling.tiles_y * 4 * tiling.R + AFF_BY - 1) / AFF_BY, B);
    if (grid.y > 65535 || grid.z > 65535)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity relayout: H/4 and B must be <= 65535");
    aff_relayout_kernel<<<grid, block, 0, s>>>(aff_std, aff_tiled, H, W, tiling);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr
