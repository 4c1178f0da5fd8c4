// Per-pixel arithmetic of the local affinity, shared by every kernel that computes it (the TMA / Tensor-Memory tile
// kernel, the shared-memory kernel of the standard layout, the generic global-memory kernel, the resident small-map
// kernel).  Replaces reference models/mods/pamr.py:132-136:
//   x_std = LocalStDev(x)                          (:77-103; 9*nd samples incl. the centre, unbiased std)
//   a     = -|x - x_nbr| / (1e-8 + 0.1*x_std)      (:134, LocalAffinityAbs :105-109)
//   a     = mean over the K image channels         (:135)
//   w     = softmax over the 8*nd neighbours       (:136)
//
// The kernels are instruction-issue bound, so the arithmetic is arranged for few instructions at (measured) no loss of
// accuracy against the reference -- tools/affinity_numerics.py emulates every variant in fp32 on nine input families
// (max-abs error of the 48 weights against the double-Welford / IEEE-division CPU restatement of the reference):
//   fully IEEE fp32 (two-pass std, division, expf, division)             2.4e-7 .. 5.4e-7
//   this file (two-pass std, reciprocal + FMA, ex2.approx, reciprocal)   3.0e-7 .. 5.7e-7
//   one-pass std (sum u, sum u^2) + the same                             up to 2.2e-6   (rejected)
//  * std: samples are shifted by the centre value first (u = v - c: every rounding error then scales with the local
//    contrast, not with the absolute intensity), two passes (mean of u, then sum of (u - mean)^2), per-dilation
//    partial sums.  The one-pass form loses a factor ~9 to cancellation when the centre pixel is an outlier.
//  * the 8*nd*K divisions by den_k = 1e-8 + 0.1 std_k become ONE FMA each: A[p] += |u_kp| * nr_k with
//    nr_k = -(log2(e) / K) * rcp(den_k), which also folds the mean over the channels (:135) and the conversion to
//    base 2 for the exponential;
//  * softmax: w = ex2.approx(A - max A) * rcp(sum); ex2.approx is good to 2 ulp, and the flush of results below
//    1.2e-38 to zero is invisible at the 1e-5 parity bar.
#pragma once
#include "pamr_common.cuh"

namespace pamr {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Unbiased standard deviation of the 9*ND samples around a pixel, one image channel.
// u[8*i + j] = sample(dilation i, neighbour j) - centre for the 8 neighbours j (the centre itself contributes u = 0,
// once per dilation).  NDT > 0: number of dilations at compile time (everything stays in registers).
template <int NDT>
__device__ __forceinline__ float std_from_shifted(const float (&u)[8 * NDT], int nd) {
    float part[NDT];
#pragma unroll
    for (int i = 0; i < NDT; ++i) {
        float t = 0.f;
        if (i < nd) {
#pragma unroll
            for (int j = 0; j < 8; ++j) t += u[8 * i + j];
        }
        part[i] = t;
    }
    float s1;
    if (NDT == 6) {
        s1 = ((part[0] + part[1]) + (part[2] + part[3])) + (part[4] + part[5]);
    } else {
        s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NDT; ++i) s1 += part[i];
    }
    const float n = (float)(9 * nd);
    const float mean_u = (nd == 6) ? s1 * (1.0f / 54.0f) : s1 / n;
#pragma unroll
    for (int i = 0; i < NDT; ++i) {
        float t = 0.f;
        if (i < nd) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dv = u[8 * i + j] - mean_u;
                t = fmaf(dv, dv, t);
            }
            t = fmaf(mean_u, mean_u, t);  // the dilation's centre sample
        }
        part[i] = t;
    }
    float m2;
    if (NDT == 6) {
        m2 = ((part[0] + part[1]) + (part[2] + part[3])) + (part[4] + part[5]);
    } else {
        m2 = 0.f;
#pragma unroll
        for (int i = 0; i < NDT; ++i) m2 += part[i];
    }
    return sqrtf((nd == 6) ? m2 * (1.0f / 53.0f) : m2 / (n - 1.0f));
}

// nr_k of the header comment from the channel's std
__device__ __forceinline__ float neg_scaled_rcp_den(float sd, float log2e_over_k) {
    const float den = __fadd_rn(1e-8f, __fmul_rn(0.1f, sd));  // in [1e-8, ...): normal range
    return -(__frcp_rn(den) * log2e_over_k);
}

// Softmax of the 8*ND base-2 logits A (in place: A becomes the weights), pamr.py:136.  `scale` multiplies every
// weight (1, or 0 for a pixel whose weights must be zero).  Maximum and sum are formed as four interleaved partial
// chains: the kernels run few warps per scheduler, so a 48-long dependent chain would sit on the issue slot.
template <int NDT>
__device__ __forceinline__ void softmax_base2(float (&A)[8 * NDT], int nd, float scale = 1.0f) {
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int p = 0; p < 8 * NDT; ++p)
        if (p < 8 * nd) m4[p & 3] = fmaxf(m4[p & 3], A[p]);
    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < 8 * NDT; ++p) {
        if (p < 8 * nd) {
            A[p] = ex2_approx(A[p] - mx);
            s[p & 3] += A[p];
        }
    }
    const float rs = __frcp_rn((s[0] + s[1]) + (s[2] + s[3])) * scale;  // the sum is in [1, 8 nd]
#pragma unroll
    for (int p = 0; p < 8 * NDT; ++p)
        if (p < 8 * nd) A[p] *= rs;
}

// The whole pixel.  fetch(k, i, j9) returns the sample of image channel k at dilation index i, position j9 of the
// 3x3 neighbourhood (j9 = 4: the centre; replicate padding is the fetcher's business).  w[8*i + j] receives the
// weight of reference tap p = 8*i + j (pamr.py:25-34).  sd_out: nullptr or K floats receiving the channels' std.
// The channel loops are kept rolled (#pragma unroll 1): the unrolled body of one channel is ~400 instructions.
constexpr int AFF_MAXK = 8;
template <int NDT, class Fetch>
__device__ __forceinline__ void affinity_pixel(const Fetch& fetch, int K, int nd, float (&w)[8 * NDT], float scale = 1.0f) {
    float nr[AFF_MAXK];  // indexed only with compile-time constants below (registers)
#pragma unroll
    for (int k = 0; k < AFF_MAXK; ++k) nr[k] = 0.f;
    const float l2k = 1.4426950408889634f / (float)K;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        float u[8 * NDT];
        const float c = fetch(k, 0, 4);
#pragma unroll
        for (int i = 0; i < NDT; ++i) {
            if (i < nd) {
#pragma unroll
                for (int j = 0; j < 8; ++j) u[8 * i + j] = __fsub_rn(fetch(k, i, j < 4 ? j : j + 1), c);
            }
        }
        const float v = neg_scaled_rcp_den(std_from_shifted<NDT>(u, nd), l2k);
#pragma unroll
        for (int q = 0; q < AFF_MAXK; ++q)
            if (q == k) nr[q] = v;
    }
#pragma unroll
    for (int p = 0; p < 8 * NDT; ++p) w[p] = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        float v = nr[0];
#pragma unroll
        for (int q = 1; q < AFF_MAXK; ++q)
            if (q == k) v = nr[q];
        const float c = fetch(k, 0, 4);
#pragma unroll
        for (int i = 0; i < NDT; ++i) {
            if (i < nd) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    w[8 * i + j] = fmaf(fabsf(__fsub_rn(fetch(k, i, j < 4 ? j : j + 1), c)), v, w[8 * i + j]);
            }
        }
    }
    softmax_base2<NDT>(w, nd, scale);
}

}  // namespace pamr
