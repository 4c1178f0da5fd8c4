// Class-balanced cross-entropy on the pseudo-labels and its gradient.
// Replaces (paths relative to the reference repo root)  models/SoftMaxAE.py:52-88  balanced_mask_loss_ce:
//   z     = bilinear(mask logits -> label resolution, align_corners=True)          (:58)  never materialised
//   label = argmax pseudo_gt / 255 where no class is set                           (:61-66) = the uint8 label map
//   n[b,c], cw[b,c] = (tot - n)/(1 + tot), bw[b] = (sum gt_labels + 1 == #present)  (:71-74, :82-84)
//   loss[b] = bw[b] * mean_px( cw[b,label] * (logsumexp_c z - z[label]) )           (:77, :86)
// and d(sum_b g[b] loss[b]) / d logits through the transpose of the interpolation.
//
// Forward: one thread per label pixel, online log-sum-exp over the C interpolated logits, per-batch
// sum in double (warp shuffle -> block -> one atomicAdd(double) per block); it also leaves lse(px) and
// coef(px) = cw[label]/(H*W) (0 where ignored) in the workspace.
// Backward: gather, deterministic, no atomics: one thread per logit element (b,c,i,j) visits the label
// pixels whose bilinear footprint touches (i,j) (about (2*scale)^2 of them), recomputes z_c there and
// adds coef * (exp(z_c - lse) - [c == label]) * wy * wx.
#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int LS_BX = 32, LS_BY = 8;
constexpr int CE_BATCH = 7;  // class planes loaded per batch in the forward kernel (21 = 3 x 7)

struct CeWorkspace {
    double* acc;   // [B]    sum_px cw*ce
    float* cwbw;   // [B, C+1]  cw[b,0..C-1], bw[b]
    float* lse;    // [B,H,W]
    float* coef;   // [B,H,W]
};
__host__ __device__ inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
inline CeWorkspace carve(void* ws, int B, int C, int H, int W) {
    char* p = (char*)ws;
    CeWorkspace r;
    r.acc = (double*)p; p += up256(sizeof(double) * B);
    r.cwbw = (float*)p; p += up256(sizeof(float) * (size_t)B * (C + 1));
    r.lse = (float*)p; p += up256(sizeof(float) * (size_t)B * H * W);
    r.coef = (float*)p;
    return r;
}

// pseudo_gt float one-hot-or-empty -> uint8 labels (first maximum, like torch.argmax; 255 where the
// pixel's sum is < 1) and per-class pixel counts.  grid (tiles_x, tiles_y, B)
__global__ void __launch_bounds__(LS_BX * LS_BY)
labels_from_onehot_kernel(const float* __restrict__ pg, uint8_t* __restrict__ label, int* __restrict__ count, int C,
                          int H, int W) {
    const int x = blockIdx.x * LS_BX + threadIdx.x, y = blockIdx.y * LS_BY + threadIdx.y, b = blockIdx.z;
    const bool valid = x < W && y < H;
    const size_t HW = (size_t)H * W, i = (size_t)y * W + x;
    int lab = 255;
    if (valid) {
        float best = 0.f, sum = 0.f;
        int arg = 0;
        for (int c0 = 0; c0 < C; c0 += CE_BATCH) {
            float vb[CE_BATCH];
#pragma unroll
            for (int j = 0; j < CE_BATCH; ++j) vb[j] = __ldg(pg + ((size_t)b * C + min(c0 + j, C - 1)) * HW + i);
#pragma unroll
            for (int j = 0; j < CE_BATCH; ++j) {
                const int c = c0 + j;
                if (c < C) {
                    sum = __fadd_rn(sum, vb[j]);
                    if (c == 0 || vb[j] > best) { best = vb[j]; arg = c; }
                }
            }
        }
        lab = (sum < 1.f) ? 255 : arg;
        label[(size_t)b * HW + i] = (uint8_t)lab;
    }
    if (count != nullptr) {
        for (int c = 0; c < C; ++c) {  // warp-aggregated
            const unsigned m = __ballot_sync(0xffffffffu, lab == c);
            if (m != 0u && (threadIdx.x & 31) == 0) atomicAdd(count + (size_t)b * C + c, __popc(m));
        }
    }
}

// cw / bw from the counts (float arithmetic as in the reference) and zero the accumulators.  one block
__global__ void ce_stats_kernel(const int* __restrict__ count, const float* __restrict__ gt_labels,
                                float* __restrict__ cwbw, double* __restrict__ acc, int B, int C) {
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float tot = 0.f;
        int present = 0;
        for (int c = 0; c < C; ++c) {
            const int n = count[(size_t)b * C + c];
            tot = __fadd_rn(tot, (float)n);
            present += n > 0;
        }
        for (int c = 0; c < C; ++c)
            cwbw[(size_t)b * (C + 1) + c] = __fdiv_rn(__fsub_rn(tot, (float)count[(size_t)b * C + c]), __fadd_rn(1.f, tot));
        float gsum = 1.f;  // + BG
        for (int c = 0; c < C - 1; ++c) gsum = __fadd_rn(gsum, gt_labels[(size_t)b * (C - 1) + c]);
        cwbw[(size_t)b * (C + 1) + C] = (gsum == (float)present) ? 1.f : 0.f;
        acc[b] = 0.0;
    }
}

template <bool kResize>
__device__ __forceinline__ float logit_at(const float* __restrict__ pl, int w, size_t i, const Lerp& ly, const Lerp& lx) {
    return kResize ? bilerp(pl, w, ly, lx) : __ldg(pl + i);
}

// grid (tiles_x, tiles_y, B)
template <bool kResize>
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_forward_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const float* __restrict__ cwbw,
                  double* __restrict__ acc, float* __restrict__ lse_out, float* __restrict__ coef_out, int C, int h, int w,
                  int H, int W, float sh, float sw) {
    const int x = blockIdx.x * LS_BX + threadIdx.x, y = blockIdx.y * LS_BY + threadIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W, hw = (size_t)h * w, i = (size_t)y * W + x;
    double term = 0.0;
    if (x < W && y < H) {
        const int lab = label[(size_t)b * HW + i];
        float lse = 0.f, coef = 0.f;
        if (lab < C) {
            const Lerp ly = make_lerp(y, sh, h), lx = make_lerp(x, sw, w);
            const float* __restrict__ base = logits + (size_t)b * C * hw;
            float m = -INFINITY, s = 0.f, zl = 0.f;  // online log-sum-exp
            // classes in batches of CE_BATCH: the loads of a batch are issued before the dependent exp chain
            for (int c0 = 0; c0 < C; c0 += CE_BATCH) {
                float vb[CE_BATCH];
#pragma unroll
                for (int j = 0; j < CE_BATCH; ++j)
                    vb[j] = logit_at<kResize>(base + (size_t)min(c0 + j, C - 1) * hw, w, i, ly, lx);
#pragma unroll
                for (int j = 0; j < CE_BATCH; ++j) {
                    const int c = c0 + j;
                    if (c < C) {
                        const float v = vb[j];
                        if (c == lab) zl = v;
                        if (v > m) {
                            s = s * expf(m - v) + 1.f;
                            m = v;
                        } else {
                            s += expf(v - m);
                        }
                    }
                }
            }
            lse = m + logf(s);
            const float cw = cwbw[(size_t)b * (C + 1) + lab];
            term = (double)cw * (double)(lse - zl);
            coef = cw / (float)HW;
        }
        lse_out[(size_t)b * HW + i] = lse;
        coef_out[(size_t)b * HW + i] = coef;
    }
    // block sum in double -> one atomic per block
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    __shared__ double red[LS_BY];
    if (threadIdx.x == 0) red[threadIdx.y] = term;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double t = 0.0;
        for (int k = 0; k < LS_BY; ++k) t += red[k];
        if (t != 0.0) atomicAdd(acc + b, t);
    }
}

__global__ void ce_finalize_kernel(const double* __restrict__ acc, const float* __restrict__ cwbw, float* __restrict__ loss,
                                   int B, int C, double inv_hw) {
    for (int b = threadIdx.x; b < B; b += blockDim.x)
        loss[b] = cwbw[(size_t)b * (C + 1) + C] * (float)(acc[b] * inv_hw);
}

// Destination index range [lo, hi] whose interpolation can touch source index i (conservative; the
// exact weights decide).  scale = (in-1)/(out-1).
__device__ __forceinline__ void footprint(int i, float scale, int out_size, int& lo, int& hi) {
    if (scale <= 0.f) { lo = 0; hi = out_size - 1; return; }
    lo = max(0, (int)floorf((float)(i - 1) / scale) - 1);
    hi = min(out_size - 1, (int)ceilf((float)(i + 1) / scale) + 1);
}
__device__ __forceinline__ float weight_to(const Lerp& l, int i) {
    return (l.i0 == i ? l.l0 : 0.f) + (l.i1 == i ? l.l1 : 0.f);
}

// grid (ceil(w/32), ceil(h/8), B*C): one thread per logit element.  With lse(px) stored by the forward
// pass, class c needs only its own interpolated logit: softmax_c = exp(z_c - lse).
template <bool kResize>
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_backward_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const float* __restrict__ cwbw,
                   const float* __restrict__ lse_in, const float* __restrict__ coef_in, const float* __restrict__ gout,
                   float* __restrict__ grad, int C, int h, int w, int H, int W, float sh, float sw) {
    const int j = blockIdx.x * LS_BX + threadIdx.x, i = blockIdx.y * LS_BY + threadIdx.y;
    const int plane = blockIdx.z, b = plane / C, c = plane - b * C;
    if (j >= w || i >= h) return;
    const size_t HW = (size_t)H * W, hw = (size_t)h * w;
    const float g = gout[b] * cwbw[(size_t)b * (C + 1) + C];
    float acc = 0.f;
    if (g != 0.f) {
        const float* __restrict__ pl = logits + (size_t)plane * hw;
        int ylo = i, yhi = i, xlo = j, xhi = j;
        if (kResize) {
            footprint(i, sh, H, ylo, yhi);
            footprint(j, sw, W, xlo, xhi);
        }
        for (int y = ylo; y <= yhi; ++y) {
            const Lerp ly = make_lerp(y, sh, h);
            const float wy = kResize ? weight_to(ly, i) : 1.f;
            if (wy == 0.f) continue;
            for (int x = xlo; x <= xhi; ++x) {
                const Lerp lx = make_lerp(x, sw, w);
                const float wx = kResize ? weight_to(lx, j) : 1.f;
                const size_t pi = (size_t)y * W + x, p = (size_t)b * HW + pi;
                const float coef = __ldg(coef_in + p);
                if (wx == 0.f || coef == 0.f) continue;
                const float z = logit_at<kResize>(pl, w, pi, ly, lx);
                acc = fmaf(coef * wy * wx, expf(z - __ldg(lse_in + p)) - (label[p] == c ? 1.f : 0.f), acc);
            }
        }
    }
    grad[(size_t)plane * hw + (size_t)i * w + j] = g * acc;
}

}  // namespace

size_t mask_ce_workspace_bytes(int B, int C, int H, int W) {
    return up256(sizeof(double) * B) + up256(sizeof(float) * (size_t)B * (C + 1)) + 2 * up256(sizeof(float) * (size_t)B * H * W);
}

int launch_labels_from_onehot(const float* pseudo_gt, uint8_t* label, int* class_count, int B, int C, int H, int W,
                              cudaStream_t s) {
    if (class_count != nullptr) PAMR_CUDA_TRY(cudaMemsetAsync(class_count, 0, sizeof(int) * (size_t)B * C, s));
    dim3 grid((W + LS_BX - 1) / LS_BX, (H + LS_BY - 1) / LS_BY, B), block(LS_BX, LS_BY);
    if (grid.y > 65535 || grid.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "labels_from_onehot: H/8 and B must be <= 65535");
    labels_from_onehot_kernel<<<grid, block, 0, s>>>(pseudo_gt, label, class_count, C, H, W);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_mask_ce_forward(const float* logits, const uint8_t* label, const int* class_count, const float* gt_labels,
                           float* loss, void* ws, size_t ws_bytes, int B, int C, int h, int w, int H, int W,
                           cudaStream_t s) {
    if (ws == nullptr || ws_bytes < mask_ce_workspace_bytes(B, C, H, W))
        return set_error(PAMR_ERR_WORKSPACE, "mask_ce: workspace of %zu bytes given, %zu needed", ws_bytes,
                         mask_ce_workspace_bytes(B, C, H, W));
    if (((uintptr_t)ws & 255) != 0) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce: workspace must be 256-byte aligned");
    const CeWorkspace k = carve(ws, B, C, H, W);
    ce_stats_kernel<<<1, 128, 0, s>>>(class_count, gt_labels, k.cwbw, k.acc, B, C);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    dim3 grid((W + LS_BX - 1) / LS_BX, (H + LS_BY - 1) / LS_BY, B), block(LS_BX, LS_BY);
    if (grid.y > 65535 || grid.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce: H/8 and B must be <= 65535");
    const float sh = scale_of(h, H), sw = scale_of(w, W);
    if (h != H || w != W)
        ce_forward_kernel<true><<<grid, block, 0, s>>>(logits, label, k.cwbw, k.acc, k.lse, k.coef, C, h, w, H, W, sh, sw);
    else
        ce_forward_kernel<false><<<grid, block, 0, s>>>(logits, label, k.cwbw, k.acc, k.lse, k.coef, C, h, w, H, W, sh, sw);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    ce_finalize_kernel<<<1, 128, 0, s>>>(k.acc, k.cwbw, loss, B, C, 1.0 / ((double)H * (double)W));
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_mask_ce_backward(const float* logits, const uint8_t* label, const float* grad_loss, float* grad_logits,
                            const void* ws, size_t ws_bytes, int B, int C, int h, int w, int H, int W, cudaStream_t s) {
    if (ws == nullptr || ws_bytes < mask_ce_workspace_bytes(B, C, H, W))
        return set_error(PAMR_ERR_WORKSPACE, "mask_ce backward: workspace of %zu bytes given, %zu needed", ws_bytes,
                         mask_ce_workspace_bytes(B, C, H, W));
    const CeWorkspace k = carve(const_cast<void*>(ws), B, C, H, W);
    dim3 grid((w + LS_BX - 1) / LS_BX, (h + LS_BY - 1) / LS_BY, B * C), block(LS_BX, LS_BY);
    if (grid.y > 65535 || grid.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce backward: h/8 and B*C must be <= 65535");
    const float sh = scale_of(h, H), sw = scale_of(w, W);
    if (h != H || w != W)
        ce_backward_kernel<true><<<grid, block, 0, s>>>(logits, label, k.cwbw, k.lse, k.coef, grad_loss, grad_logits, C, h, w, H, W, sh, sw);
    else
        ce_backward_kernel<false><<<grid, block, 0, s>>>(logits, label, k.cwbw, k.lse, k.coef, grad_loss, grad_logits, C, h, w, H, W, sh, sw);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr
