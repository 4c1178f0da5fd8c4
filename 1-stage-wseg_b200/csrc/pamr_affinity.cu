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
#include <atomic>

#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int AFF_BX = 32;
constexpr int AFF_BY = 4;

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
                AffTiling full = tiling;  // address it as a tile pixel (it is neither of the strips)
                full.Wt = tiling.tiles_x * 32; full.Ht = tiling.tiles_y * 4 * tiling.R;
                const AffPixel px = aff_pixel(full, b, y, x);
#pragma unroll
                for (int s = 0; s < 48; ++s) aff[px.at(s)] = 0.f;
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
        const AffPixel px = aff_pixel(tiling, b, y, x);
#pragma unroll
        for (int p = 0; p < 8 * MAXND; ++p) aff[px.at(tap_seq(p))] = div_markstein(abar[p], s, rs);
    } else {
        float* __restrict__ out = aff + (size_t)b * P * HW + (size_t)y * W + x;
#pragma unroll
        for (int p = 0; p < 8 * MAXND; ++p) {
            if (p < P) out[(size_t)p * HW] = div_markstein(abar[p], s, rs);
        }
    }
}

// ---- shared-memory staged variant for the standard dilation set [1,2,4,8,12,24] ----
// A CTA owns a 32 x 8 pixel tile; the K image planes of the tile plus its 24-pixel halo are staged
// in shared memory once, with replicate padding applied while staging (clamped source
// coordinates), so every neighbour is a shared-memory load at an immediate offset: no 64-bit
// address arithmetic and no clamps in the inner loops (128 registers, 2 CTAs = 16 warps
// per SM).
constexpr int SA_BX = 32, SA_BY = 8, SA_HALO = 24;
constexpr int SA_W = SA_BX + 2 * SA_HALO;  // 80
constexpr int SA_H = SA_BY + 2 * SA_HALO;  // 56
constexpr int SA_MAXK = 8;
__host__ __device__ constexpr int sa_dil(int id) { return id == 0 ? 1 : id == 1 ? 2 : id == 2 ? 4 : id == 3 ? 8 : id == 4 ? 12 : 24; }

template <bool TILED>
__global__ void __launch_bounds__(SA_BX * SA_BY, 2)  // 128 registers: 3 CTAs/SM (85) spilled the 48 accumulators, 6 % slower
affinity_smem_kernel(const float* __restrict__ img, float* __restrict__ aff, int K, int H, int W, AffTiling tiling) {
    extern __shared__ float sa_tile[];  // [K][SA_H][SA_W]
    const int x0 = blockIdx.x * SA_BX, y0 = blockIdx.y * SA_BY, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    // staging: thread (tx,ty) covers window columns tx, tx+32, tx+64 and rows ty, ty+8, ... of every
    // plane; all 21 loads of a plane are in flight before the first shared-memory store
    for (int k = 0; k < K; ++k) {
        const float* __restrict__ pl = img + ((size_t)b * K + k) * HW;
        float v[3][SA_H / SA_BY];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int wx = threadIdx.x + 32 * cc;
            const int gx = clampi(x0 - SA_HALO + wx, 0, W - 1);
#pragma unroll
            for (int rr = 0; rr < SA_H / SA_BY; ++rr) {
                const int gy = clampi(y0 - SA_HALO + (int)threadIdx.y + SA_BY * rr, 0, H - 1);
                v[cc][rr] = (wx < SA_W) ? __ldg(pl + (size_t)gy * W + gx) : 0.f;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int wx = threadIdx.x + 32 * cc;
#pragma unroll
            for (int rr = 0; rr < SA_H / SA_BY; ++rr)
                if (wx < SA_W) sa_tile[(k * SA_H + threadIdx.y + SA_BY * rr) * SA_W + wx] = v[cc][rr];
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) {
        if (TILED && x < tiling.tiles_x * 32 && y < tiling.tiles_y * 4 * tiling.R) {
            AffTiling full = tiling;  // a partial tile's pixel outside the image: zeros
            full.Wt = tiling.tiles_x * 32; full.Ht = tiling.tiles_y * 4 * tiling.R;
            const AffPixel px = aff_pixel(full, b, y, x);
#pragma unroll
            for (int s = 0; s < 48; ++s) aff[px.at(s)] = 0.f;
        }
        return;
    }
    float abar[48];
    for (int k = 0; k < K; ++k) {
        const float* __restrict__ c0 = sa_tile + (k * SA_H + threadIdx.y + SA_HALO) * SA_W + threadIdx.x + SA_HALO;
        // the 54 samples are re-read from shared memory in every pass (immediate offsets) rather
        // than held in registers: 48 accumulators + 54 samples would spill at 128 registers
#define SA_SMP(i, j) c0[((j) / 3 - 1) * sa_dil(i) * SA_W + ((j) % 3 - 1) * sa_dil(i)]
        // Unbiased std of the 54 samples in fp32, conditioned so that it tracks torch's double
        // Welford to ~1e-7 relative: samples are first shifted by the centre value (all rounding
        // errors then scale with the local contrast, not with the absolute intensity), and both
        // sums are formed as 6 per-dilation partial sums combined at the end (pairwise-style), so
        // the accumulation error stays at a few ulp instead of ~54 ulp.  (The FP64 pipe, used by an
        // earlier version for 810 conversions/adds/FMAs per pixel, was this kernel's bottleneck.)
        const float cc = c0[0];
        float su[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) t += SA_SMP(i, j) - cc;
            su[i] = t;
        }
        const float mean_u = (((su[0] + su[1]) + (su[2] + su[3])) + (su[4] + su[5])) * (1.0f / 54.0f);
        asm volatile("" ::: "memory");  // keep the compiler from caching all 54 samples in registers
        float sq[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const float dv = (SA_SMP(i, j) - cc) - mean_u;
                t = fmaf(dv, dv, t);
            }
            sq[i] = t;
        }
        const float m2 = ((sq[0] + sq[1]) + (sq[2] + sq[3])) + (sq[4] + sq[5]);
        const float sd = sqrtf(m2 / 53.0f);
        const float den = __fadd_rn(1e-8f, __fmul_rn(0.1f, sd));
        const float rden = __frcp_rn(den);
        asm volatile("" ::: "memory");
        const float c = c0[0];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int j9 = (j < 4) ? j : j + 1;  // skip the centre sample
                const float a = div_markstein(-fabsf(__fsub_rn(c, SA_SMP(i, j9))), den, rden);
                abar[8 * i + j] = (k == 0) ? a : __fadd_rn(abar[8 * i + j], a);
            }
        }
#undef SA_SMP
    }
    const float kf = (float)K, rk = __frcp_rn(kf);
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < 48; ++p) {
        abar[p] = div_markstein(abar[p], kf, rk);
        mx = fmaxf(mx, abar[p]);
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 48; ++p) {
        abar[p] = expf(abar[p] - mx);
        s += abar[p];
    }
    const float rs = __frcp_rn(s);
    if (TILED) {
        const AffPixel px = aff_pixel(tiling, b, y, x);
#pragma unroll
        for (int p = 0; p < 48; ++p) aff[px.at(tap_seq(p))] = div_markstein(abar[p], s, rs);
    } else {
        float* __restrict__ out = aff + (size_t)b * 48 * HW + (size_t)y * W + x;
#pragma unroll
        for (int p = 0; p < 48; ++p) out[(size_t)p * HW] = div_markstein(abar[p], s, rs);
    }
}

bool standard_dilations(const Dilations& dil) {
    if (dil.nd != 6) return false;
    for (int i = 0; i < 6; ++i)
        if (dil.d[i] != sa_dil(i)) return false;
    return true;
}

// standard [B,48,H,W] -> tile-major (for callers of the public propagate API)
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
aff_relayout_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, AffTiling tiling) {
    const int x = blockIdx.x * AFF_BX + threadIdx.x;
    const int y = blockIdx.y * AFF_BY + threadIdx.y;
    const int b = blockIdx.z;
    const bool in = (x < W) && (y < H);
    if (!in && (x >= tiling.tiles_x * 32 || y >= tiling.tiles_y * 4 * tiling.R)) return;
    const size_t HW = (size_t)H * W;
    const float* __restrict__ ip = src + (size_t)b * 48 * HW + (size_t)y * W + x;
    AffTiling full = tiling;
    if (!in) { full.Wt = tiling.tiles_x * 32; full.Ht = tiling.tiles_y * 4 * tiling.R; }  // partial-tile padding: zeros
    const AffPixel px = aff_pixel(full, b, y, x);
#pragma unroll 8
    for (int p = 0; p < 48; ++p) dst[px.at(tap_seq(p))] = in ? __ldg(ip + (size_t)p * HW) : 0.f;
}

}  // namespace

int launch_affinity(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil,
                    const AffTiling& tiling, cudaStream_t s) {
    const bool tiled = tiling.R > 0;
    const int gw = tiled ? max(tiling.tiles_x * 32, W) : W, gh = tiled ? max(tiling.tiles_y * 4 * tiling.R, H) : H;
    if (standard_dilations(dil) && K <= SA_MAXK) {
        dim3 sblock(SA_BX, SA_BY);
        dim3 sgrid((gw + SA_BX - 1) / SA_BX, (gh + SA_BY - 1) / SA_BY, B);
        if (sgrid.y > 65535 || sgrid.z > 65535)
            return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity: H/8 and B must be <= 65535");
        const size_t smem = sizeof(float) * (size_t)K * SA_H * SA_W;
        static std::atomic<int> attr_set[64];
        int dev = 0;
        PAMR_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
            PAMR_CUDA_TRY(cudaFuncSetAttribute(affinity_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(sizeof(float) * SA_MAXK * SA_H * SA_W)));
            PAMR_CUDA_TRY(cudaFuncSetAttribute(affinity_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(sizeof(float) * SA_MAXK * SA_H * SA_W)));
            if (dev < 64) attr_set[dev].store(1, std::memory_order_release);
        }
        if (tiled) affinity_smem_kernel<true><<<sgrid, sblock, smem, s>>>(img, aff, K, H, W, tiling);
        else affinity_smem_kernel<false><<<sgrid, sblock, smem, s>>>(img, aff, K, H, W, tiling);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        return PAMR_OK;
    }
    dim3 block(AFF_BX, AFF_BY);
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
    const int gw = max(tiling.tiles_x * 32, W), gh = max(tiling.tiles_y * 4 * tiling.R, H);
    dim3 grid((gw + AFF_BX - 1) / AFF_BX, (gh + AFF_BY - 1) / AFF_BY, B);
    if (grid.y > 65535 || grid.z > 65535)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity relayout: H/4 and B must be <= 65535");
    aff_relayout_kernel<<<grid, block, 0, s>>>(aff_std, aff_tiled, H, W, tiling);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr
