// Propagation kernels: replace the loop at reference models/mods/pamr.py:138-140
//   for _ in range(num_iter):  m = aff_m(mask);  mask = (m * x).sum(2)
// i.e. M'[b,c,y,x] = sum_p w[b,p,y,x] * M[b,c,clamp(y+dy_p),clamp(x+dx_p)].
// The reference materialises the [B,C,48,H,W] unfolded tensor every iteration; here every
// iteration is one stencil pass (read affinity once, mask once, write mask once).
//
// This file holds the GENERIC kernel (any dilation list, any C, any H/W): one thread per pixel,
// neighbours fetched through L1 with clamped coordinates.  The tuned sm_100a kernel for the
// standard dilation set lives in pamr_propagate_sm100.cu and is selected in launch_propagate().
#include "pamr_common.cuh"

namespace pamr {

int launch_propagate_tuned(const float* aff, const float* m_in, float* m_out, int B, int C, int H, int W,
                           const Dilations& dil, unsigned* cls_max, int dev, cudaStream_t s, bool* handled);

namespace {

constexpr int GEN_BX = 32;
constexpr int GEN_BY = 8;
constexpr int GEN_CC = 7;  // classes accumulated per pass over the taps

__global__ void __launch_bounds__(GEN_BX * GEN_BY)
propagate_generic_kernel(const float* __restrict__ aff, const float* __restrict__ m_in, float* __restrict__ m_out,
                         int C, int H, int W, Dilations dil, unsigned* __restrict__ cls_max) {
    const int x = blockIdx.x * GEN_BX + threadIdx.x;
    const int y = blockIdx.y * GEN_BY + threadIdx.y;
    const int b = blockIdx.z;
    const bool valid = (x < W) && (y < H);
    const int xc = min(x, W - 1), yc = min(y, H - 1);
    const size_t HW = (size_t)H * W;
    const int P = 8 * dil.nd;
    const float* __restrict__ wp = aff + (size_t)b * P * HW + (size_t)yc * W + xc;

    for (int c0 = 0; c0 < C; c0 += GEN_CC) {
        float acc[GEN_CC];
#pragma unroll
        for (int cc = 0; cc < GEN_CC; ++cc) acc[cc] = 0.f;
        const float* __restrict__ mb = m_in + ((size_t)b * C + c0) * HW;
        for (int i = 0; i < dil.nd; ++i) {
            const int d = dil.d[i];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int yy = clampi(yc + tap_dy(j) * d, 0, H - 1);
                const int xx = clampi(xc + tap_dx(j) * d, 0, W - 1);
                const float wv = __ldg(wp + (size_t)(8 * i + j) * HW);
                const float* __restrict__ q = mb + (size_t)yy * W + xx;
#pragma unroll
                for (int cc = 0; cc < GEN_CC; ++cc)
                    if (c0 + cc < C) acc[cc] = fmaf(wv, __ldg(q + (size_t)cc * HW), acc[cc]);
            }
        }
#pragma unroll
        for (int cc = 0; cc < GEN_CC; ++cc) {
            if (c0 + cc < C) {
                if (valid) m_out[((size_t)b * C + c0 + cc) * HW + (size_t)y * W + x] = acc[cc];
                if (cls_max != nullptr) {
                    unsigned u = valid ? ordered_from_float(acc[cc]) : 0u;
                    u = __reduce_max_sync(0xffffffffu, u);
                    if (threadIdx.x == 0 && u != 0u) atomicMax(cls_max + (size_t)b * C + c0 + cc, u);
                }
            }
        }
    }
}

__global__ void class_max_kernel(const float* __restrict__ m, unsigned* __restrict__ cls_max, size_t HW) {
    // one block per (b,c) plane slice; used only when iters == 0 and a max is requested
    const size_t plane = blockIdx.y;
    const float* __restrict__ p = m + plane * HW;
    unsigned u = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x)
        u = max(u, ordered_from_float(p[i]));
    u = __reduce_max_sync(0xffffffffu, u);
    if ((threadIdx.x & 31) == 0 && u != 0u) atomicMax(cls_max + plane, u);
}

}  // namespace

int launch_propagate(const float* aff, const float* m_in, float* m_out, float* m_tmp, int B, int C, int H, int W,
                     const Dilations& dil, int iters, unsigned* cls_max, int dev, cudaStream_t s) {
    const size_t N = (size_t)B * C * H * W;
    if (cls_max != nullptr) PAMR_CUDA_TRY(cudaMemsetAsync(cls_max, 0, sizeof(unsigned) * (size_t)B * C, s));
    if (iters <= 0) {
        PAMR_CUDA_TRY(cudaMemcpyAsync(m_out, m_in, N * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (cls_max != nullptr) {
            dim3 grid((unsigned)min((size_t)64, ((size_t)H * W + 255) / 256), B * C);
            class_max_kernel<<<grid, 256, 0, s>>>(m_in, cls_max, (size_t)H * W);
            count_launch();
            PAMR_CUDA_TRY(cudaGetLastError());
        }
        return PAMR_OK;
    }
    if (iters > 1 && m_tmp == nullptr)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "propagate: m_tmp is required when iters > 1");

    const float* src = m_in;
    for (int it = 0; it < iters; ++it) {
        // ping-pong so that the last iteration lands in m_out and m_in is never written
        float* dst = ((iters - 1 - it) & 1) ? m_tmp : m_out;
        unsigned* mx = (it == iters - 1) ? cls_max : nullptr;
        bool handled = false;
        int rc = launch_propagate_tuned(aff, src, dst, B, C, H, W, dil, mx, dev, s, &handled);
        if (rc != PAMR_OK) return rc;
        if (!handled) {
            dim3 block(GEN_BX, GEN_BY);
            dim3 grid((W + GEN_BX - 1) / GEN_BX, (H + GEN_BY - 1) / GEN_BY, B);
            if (grid.y > 65535 || grid.z > 65535)
                return set_error(PAMR_ERR_INVALID_ARGUMENT, "propagate: H/8 and B must be <= 65535");
            propagate_generic_kernel<<<grid, block, 0, s>>>(aff, src, dst, C, H, W, dil, mx);
            count_launch();
            PAMR_CUDA_TRY(cudaGetLastError());
        }
        src = dst;
    }
    return PAMR_OK;
}

}  // namespace pamr
