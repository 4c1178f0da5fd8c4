// Epilogue kernels: bilinear resize, label gate + class max, thresholds / ambiguity / argmax.
// Replace (reference paths relative to the reference repo root):
//   F.interpolate(..., bilinear, align_corners=True)   models/mods/pamr.py:125, models/SoftMaxAE.py:177, :266
//   masks[:,1:] *= labels[:,:,None,None]               models/SoftMaxAE.py:267   (_rescale_and_clean)
//   pseudo_gtmask                                      models/SoftMaxAE.py:29-50
//   argmax + ignore 255                                models/SoftMaxAE.py:61-67
// The resize is never materialised for the label path: both kernels evaluate the same
// bilinear expression on the fly from the low-resolution source, so the values compared with
// the thresholds are bit-identical to the ones whose maximum defined the thresholds.
#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int EP_BX = 32;
constexpr int EP_BY = 8;

__global__ void __launch_bounds__(EP_BX * EP_BY)
resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int H, int W, float sh, float sw) {
    const int x = blockIdx.x * EP_BX + threadIdx.x;
    const int y = blockIdx.y * EP_BY + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t n = blockIdx.z;
    const Lerp ly = make_lerp(y, sh, h), lx = make_lerp(x, sw, w);
    dst[(n * H + y) * W + x] = bilerp(src + n * h * w, w, ly, lx);
}

// v = gate * bilinear(m); optional store; block max -> atomicMax(cls_max[b,c]).
// grid: (tiles_x, tiles_y, B*C)
template <bool kResize>
__global__ void __launch_bounds__(EP_BX * EP_BY)
clean_kernel(const float* __restrict__ m, const float* __restrict__ labels, float* __restrict__ cleaned,
             unsigned* __restrict__ cls_max, int C, int h, int w, int H, int W, float sh, float sw) {
    const int x = blockIdx.x * EP_BX + threadIdx.x;
    const int y = blockIdx.y * EP_BY + threadIdx.y;
    const size_t plane = blockIdx.z;
    const int b = (int)(plane / C), c = (int)(plane % C);
    const bool valid = (x < W) && (y < H);
    float v = 0.f;
    if (valid) {
        if (kResize) {
            const Lerp ly = make_lerp(y, sh, h), lx = make_lerp(x, sw, w);
            v = bilerp(m + plane * h * w, w, ly, lx);
        } else {
            v = __ldg(m + (plane * H + y) * W + x);
        }
        if (labels != nullptr && c > 0) v = __fmul_rn(v, __ldg(labels + (size_t)b * (C - 1) + (c - 1)));
        if (cleaned != nullptr) cleaned[(plane * H + y) * W + x] = v;
    }
    if (cls_max != nullptr) {
        __shared__ unsigned red[EP_BY];
        unsigned u = valid ? ordered_from_float(v) : 0u;
        u = __reduce_max_sync(0xffffffffu, u);
        if (threadIdx.x == 0) red[threadIdx.y] = u;
        __syncthreads();
        if (threadIdx.y == 0 && threadIdx.x < EP_BY) {
            u = red[threadIdx.x];
            u = __reduce_max_sync((1u << EP_BY) - 1u, u);
            if (threadIdx.x == 0 && u != 0u) atomicMax(cls_max + plane, u);
        }
    }
}

constexpr int PL_THREADS = 256;

// One thread per output pixel, loops over the C classes (coalesced plane reads).
// grid: (ceil(H*W/256), B)
// PL_BATCH class planes are loaded before the first compare: 7 (21 classes = 3 batches) when the values are
// interpolated (4 loads each), 21 when they are read directly -- a thread then has all of its 84 bytes in flight
template <bool kResize, int PL_BATCH>
__global__ void __launch_bounds__(PL_THREADS)
pseudo_labels_kernel(const float* __restrict__ m, const float* __restrict__ labels,
                     const unsigned* __restrict__ cls_max, uint8_t* __restrict__ label,
                     float* __restrict__ pseudo_gt, int* __restrict__ class_count, int C, int h, int w, int H,
                     int W, float sh, float sw, float bg_cut, float fg_cut, float low_cut, bool max_is_gated) {
    extern __shared__ float sm[];
    float* thr = sm;        // [C]
    float* gate = sm + C;   // [C]
    int* cnt = reinterpret_cast<int*>(sm + 2 * C);  // [C]
    const int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += PL_THREADS) {
        const float g = (labels != nullptr && c > 0) ? __ldg(labels + (size_t)b * (C - 1) + (c - 1)) : 1.f;
        // the max may come un-gated from the fused propagation step: x -> fl(g*x) is monotone for
        // g >= 0, so max(g*v) == g*max(v) exactly and the gate can be applied to the max here
        float mx = float_from_ordered(__ldg(cls_max + (size_t)b * C + c));
        if (!max_is_gated) mx = __fmul_rn(mx, g);
        const float t = __fmul_rn(mx, c == 0 ? bg_cut : fg_cut);  // single rounding (mask_max *= 0.7)
        thr[c] = fmaxf(t, low_cut);
        gate[c] = g;
        cnt[c] = 0;
    }
    __syncthreads();
    const size_t HW = (size_t)H * W;
    const size_t i = (size_t)blockIdx.x * PL_THREADS + threadIdx.x;
    int first = -1, n = 0;
    if (i < HW) {
        const int y = (int)(i / W), x = (int)(i % W);
        Lerp ly, lx;
        if (kResize) { ly = make_lerp(y, sh, h); lx = make_lerp(x, sw, w); }
        // classes in batches of PL_BATCH: all loads of a batch are issued before the first compare
        // (one thread otherwise has a single 4-byte load in flight: latency-, not bandwidth-bound)
        for (int c0 = 0; c0 < C; c0 += PL_BATCH) {
            float v[PL_BATCH];
#pragma unroll
            for (int j = 0; j < PL_BATCH; ++j) {
                const int c = min(c0 + j, C - 1);
                const size_t plane = (size_t)b * C + c;
                v[j] = kResize ? bilerp(m + plane * h * w, w, ly, lx) : __ldg(m + plane * HW + i);
            }
#pragma unroll
            for (int j = 0; j < PL_BATCH; ++j) {
                const int c = c0 + j;
                if (c < C) {
                    const float g = (labels != nullptr && c > 0) ? __fmul_rn(v[j], gate[c]) : v[j];
                    if (g > thr[c]) {
                        if (first < 0) first = c;
                        ++n;
                    }
                }
            }
        }
        const bool one = (n == 1);
        if (label != nullptr) label[(size_t)b * HW + i] = one ? (uint8_t)first : (uint8_t)255;
        if (pseudo_gt != nullptr)
            for (int c = 0; c < C; ++c) pseudo_gt[((size_t)b * C + c) * HW + i] = (one && c == first) ? 1.f : 0.f;
        if (!one) first = -1;
    }
    if (class_count != nullptr) {
        if (first >= 0) atomicAdd(&cnt[first], 1);
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += PL_THREADS)
            if (cnt[c] != 0) atomicAdd(class_count + (size_t)b * C + c, cnt[c]);
    }
}

// ---- the data set's denorm (datasets/pascal_voc.py:85-101, called at train.py:120) folded
// into the image down-sampling of run_pamr (SoftMaxAE.py:177): raw = norm * std[k] + mean[k] (two roundings,
// as t.mul_(s).add_(m) does), evaluated on the four source pixels, then the bilinear expression.
struct ChannelAffine {
    float scale[8], shift[8];
};
template <bool kResize>
__global__ void __launch_bounds__(EP_BX * EP_BY)
denorm_resize_kernel(const float* __restrict__ src, float* __restrict__ dst, const ChannelAffine aff, int K, int h, int w,
                     int H, int W, float sh, float sw) {
    const int x = blockIdx.x * EP_BX + threadIdx.x;
    const int y = blockIdx.y * EP_BY + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t n = blockIdx.z;
    const float sc = aff.scale[n % K], sf = aff.shift[n % K];
    const float* __restrict__ pl = src + n * h * w;
    float v;
    if (kResize) {
        const Lerp ly = make_lerp(y, sh, h), lx = make_lerp(x, sw, w);
        const float p00 = __fadd_rn(__fmul_rn(__ldg(pl + (size_t)ly.i0 * w + lx.i0), sc), sf);
        const float p01 = __fadd_rn(__fmul_rn(__ldg(pl + (size_t)ly.i0 * w + lx.i1), sc), sf);
        const float p10 = __fadd_rn(__fmul_rn(__ldg(pl + (size_t)ly.i1 * w + lx.i0), sc), sf);
        const float p11 = __fadd_rn(__fmul_rn(__ldg(pl + (size_t)ly.i1 * w + lx.i1), sc), sf);
        const float t0 = __fadd_rn(__fmul_rn(lx.l0, p00), __fmul_rn(lx.l1, p01));
        const float t1 = __fadd_rn(__fmul_rn(lx.l0, p10), __fmul_rn(lx.l1, p11));
        v = __fadd_rn(__fmul_rn(ly.l0, t0), __fmul_rn(ly.l1, t1));
    } else {
        v = __fadd_rn(__fmul_rn(__ldg(pl + (size_t)y * w + x), sc), sf);
    }
    dst[(n * H + y) * W + x] = v;
}

// ---- multi-scale merge + prediction (utils/inference_tools.py:134-161, :85-88) ----
// torch's align_corners=False source index: src = max(0, (in/out)*(dst + 0.5) - 0.5)
__device__ __forceinline__ Lerp make_lerp_half_pixel(int dst, int in_size, int out_size) {
    Lerp r;
    const float scale = __fdiv_rn((float)in_size, (float)out_size);
    const float f = fmaxf(__fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f), 0.f);
    r.i0 = min((int)f, in_size - 1);
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = __fsub_rn(f, (float)r.i0);
    r.l0 = __fsub_rn(1.f, r.l1);
    return r;
}

constexpr int MERGE_MAX_S = 16;
struct MergePads {
    int v[MERGE_MAX_S][4];  // pad_t, pad_l, h_s, w_s
};

// One thread per output pixel, all classes: mean over the scales of the un-padded, resized, un-flipped,
// label-gated score, BG ** bg_pow, then fg < thresh -> 0 and argmax (first maximum).  grid (tiles_x, tiles_y)
__global__ void __launch_bounds__(EP_BX * EP_BY)
merge_multiscale_kernel(const float* __restrict__ masks, const MergePads pads, const float* __restrict__ labels,
                        float* __restrict__ merged, uint8_t* __restrict__ pred, int S, int C, int Hp, int Wp, int H, int W,
                        int flip, float bg_pow, float thresh) {
    const int x = blockIdx.x * EP_BX + threadIdx.x, y = blockIdx.y * EP_BY + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)Hp * Wp;
    int arg = 0;
    float best = 0.f;
    for (int c = 0; c < C; ++c) {
        const float gate = (c > 0 && labels != nullptr) ? __ldg(labels + c - 1) : 1.f;
        float acc = 0.f;
        for (int s = 0; s < S; ++s) {
            const int xs = (flip && (s & 1)) ? W - 1 - x : x;
            const Lerp ly = make_lerp_half_pixel(y, pads.v[s][2], H), lx = make_lerp_half_pixel(xs, pads.v[s][3], W);
            float v = bilerp(masks + ((size_t)s * C + c) * plane + (size_t)pads.v[s][0] * Wp + pads.v[s][1], Wp, ly, lx);
            if (c > 0 && labels != nullptr) v = __fmul_rn(v, gate);
            acc = __fadd_rn(acc, v);
        }
        float m = __fdiv_rn(acc, (float)S);
        if (c == 0) m = powf(m, bg_pow);
        if (merged != nullptr) merged[((size_t)c * H + y) * W + x] = m;
        const float t = (c > 0 && m < thresh) ? 0.f : m;
        if (c == 0 || t > best) { best = t; arg = c; }
    }
    if (pred != nullptr) pred[(size_t)y * W + x] = (uint8_t)arg;
}

}  // namespace

int launch_resize_bilinear(const float* src, float* dst, int n_planes, int h, int w, int H, int W, cudaStream_t s) {
    if (h == H && w == W) {
        PAMR_CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)n_planes * H * W, cudaMemcpyDeviceToDevice, s));
        return PAMR_OK;
    }
    dim3 block(EP_BX, EP_BY);
    for (int n0 = 0; n0 < n_planes; n0 += 65535) {
        const int nn = min(65535, n_planes - n0);
        dim3 grid((W + EP_BX - 1) / EP_BX, (H + EP_BY - 1) / EP_BY, nn);
        if (grid.y > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "resize: H too large");
        resize_kernel<<<grid, block, 0, s>>>(src + (size_t)n0 * h * w, dst + (size_t)n0 * H * W, h, w, H, W,
                                             scale_of(h, H), scale_of(w, W));
        count_launch();
    }
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_clean(const float* m, const float* labels, float* cleaned, unsigned* cls_max, int B, int C, int h, int w,
                 int H, int W, cudaStream_t s) {
    if (cls_max != nullptr) PAMR_CUDA_TRY(cudaMemsetAsync(cls_max, 0, sizeof(unsigned) * (size_t)B * C, s));
    if (cleaned == nullptr && cls_max == nullptr) return PAMR_OK;
    if ((size_t)B * C > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "clean: B*C must be <= 65535");
    dim3 block(EP_BX, EP_BY);
    dim3 grid((W + EP_BX - 1) / EP_BX, (H + EP_BY - 1) / EP_BY, B * C);
    if (grid.y > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "clean: H too large");
    if (h == H && w == W)
        clean_kernel<false><<<grid, block, 0, s>>>(m, labels, cleaned, cls_max, C, h, w, H, W, 0.f, 0.f);
    else
        clean_kernel<true><<<grid, block, 0, s>>>(m, labels, cleaned, cls_max, C, h, w, H, W, scale_of(h, H),
                                                  scale_of(w, W));
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_pseudo_labels(const float* m, const float* labels, const unsigned* cls_max, uint8_t* label,
                         float* pseudo_gt, int* class_count, int B, int C, int h, int w, int H, int W, float bg_cut,
                         float fg_cut, float low_cut, bool max_is_gated, cudaStream_t s) {
    if (B > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "pseudo_labels: B must be <= 65535");
    if (class_count != nullptr) PAMR_CUDA_TRY(cudaMemsetAsync(class_count, 0, sizeof(int) * (size_t)B * C, s));
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + PL_THREADS - 1) / PL_THREADS), B);
    const size_t smem = sizeof(float) * 3 * (size_t)C;
    if (smem > 48 * 1024) return set_error(PAMR_ERR_INVALID_ARGUMENT, "pseudo_labels: C too large");
    if (h == H && w == W && C > 14)
        pseudo_labels_kernel<false, 21><<<grid, PL_THREADS, smem, s>>>(m, labels, cls_max, label, pseudo_gt, class_count,
                                                                       C, h, w, H, W, 0.f, 0.f, bg_cut, fg_cut, low_cut,
                                                                       max_is_gated);
    else if (h == H && w == W)
        pseudo_labels_kernel<false, 7><<<grid, PL_THREADS, smem, s>>>(m, labels, cls_max, label, pseudo_gt, class_count,
                                                                      C, h, w, H, W, 0.f, 0.f, bg_cut, fg_cut, low_cut,
                                                                      max_is_gated);
    else
        pseudo_labels_kernel<true, 7><<<grid, PL_THREADS, smem, s>>>(m, labels, cls_max, label, pseudo_gt, class_count,
                                                                  C, h, w, H, W, scale_of(h, H), scale_of(w, W),
                                                                  bg_cut, fg_cut, low_cut, max_is_gated);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr

namespace pamr {
int launch_denorm_resize(const float* src, float* dst, const float* mean_host, const float* std_host, int B, int K, int h,
                         int w, int H, int W, cudaStream_t s) {
    if (K < 1 || K > 8) return set_error(PAMR_ERR_INVALID_ARGUMENT, "denorm: 1..8 channels supported, got %d", K);
    ChannelAffine aff;
    for (int k = 0; k < 8; ++k) {
        aff.scale[k] = k < K ? std_host[k] : 1.f;
        aff.shift[k] = k < K ? mean_host[k] : 0.f;
    }
    dim3 grid((W + EP_BX - 1) / EP_BX, (H + EP_BY - 1) / EP_BY, B * K), block(EP_BX, EP_BY);
    if (grid.y > 65535 || grid.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "denorm: H/8 and B*K must be <= 65535");
    if (h != H || w != W)
        denorm_resize_kernel<true><<<grid, block, 0, s>>>(src, dst, aff, K, h, w, H, W, scale_of(h, H), scale_of(w, W));
    else
        denorm_resize_kernel<false><<<grid, block, 0, s>>>(src, dst, aff, K, h, w, H, W, 0.f, 0.f);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_merge_multiscale(const float* masks, const int* pads_host, const float* labels, float* merged, uint8_t* pred,
                            int S, int C, int Hp, int Wp, int H, int W, int flip, float bg_pow, float thresh,
                            cudaStream_t s) {
    if (S < 1 || S > MERGE_MAX_S) return set_error(PAMR_ERR_INVALID_ARGUMENT, "merge: 1..%d scales supported, got %d", MERGE_MAX_S, S);
    MergePads pads;
    for (int i = 0; i < S; ++i) {
        for (int k = 0; k < 4; ++k) pads.v[i][k] = pads_host[4 * i + k];
        const int pt = pads.v[i][0], pl = pads.v[i][1], hs = pads.v[i][2], ws = pads.v[i][3];
        if (pt < 0 || pl < 0 || hs < 1 || ws < 1 || pt + hs > Hp || pl + ws > Wp)
            return set_error(PAMR_ERR_INVALID_ARGUMENT, "merge: pads[%d] = (%d,%d,%d,%d) outside the %dx%d mask", i, pt, pl, hs, ws, Hp, Wp);
    }
    dim3 grid((W + EP_BX - 1) / EP_BX, (H + EP_BY - 1) / EP_BY), block(EP_BX, EP_BY);
    if (grid.y > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "merge: H/8 must be <= 65535");
    merge_multiscale_kernel<<<grid, block, 0, s>>>(masks, pads, labels, merged, pred, S, C, Hp, Wp, H, W, flip, bg_pow, thresh);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}
}  // namespace pamr
