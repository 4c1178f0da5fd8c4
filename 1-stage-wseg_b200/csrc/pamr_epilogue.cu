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

// v = gate * m; optional store; block max -> atomicMax(cls_max[b,c]).  Source and target have the same size.
// grid: (tiles_x, tiles_y, B*C)
__global__ void __launch_bounds__(EP_BX * EP_BY)
clean_kernel(const float* __restrict__ m, const float* __restrict__ labels, float* __restrict__ cleaned,
             unsigned* __restrict__ cls_max, int C, int H, int W) {
    const int x = blockIdx.x * EP_BX + threadIdx.x;
    const int y = blockIdx.y * EP_BY + threadIdx.y;
    const size_t plane = blockIdx.z;
    const int b = (int)(plane / C), c = (int)(plane % C);
    const bool valid = (x < W) && (y < H);
    float v = 0.f;
    if (valid) {
        v = __ldg(m + (plane * H + y) * W + x);
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

// ---- the resized epilogue (stage_net's real call shapes: masks 81x81 / 41x41 -> 321x321, SoftMaxAE.py:266) ----
// Column walks (ColumnWalk, pamr_common.cuh): a thread owns one output column x of a band of output rows; an output
// value costs 3 floating-point instructions instead of 4 loads + 9, and is bit-identical to bilerp()
// (pseudo_labels with resize 0.273 -> 0.054 ms, profiles/r02_real_shape_times.txt).
constexpr int CW_THREADS = 256;
constexpr int CW_ROWS = 32;      // output rows per band in the max / store pass
constexpr int CW_TAB = 512;      // row table: the Lerp of every output row a block touches, computed once per block
constexpr int PW_THREADS = 128;  // label pass: 21 classes per thread (t0 / t1 of every class in registers), 4 blocks per SM
constexpr int PW_CB = 21;        // classes held per thread in the label pass (the reference's 21; more go to the per-pixel kernel)

// grid: (ceil(nbands * W / CW_THREADS), planes); thread t of a plane: band = t / W, x = t % W.
// labels == nullptr: plain resize (C is then irrelevant); cleaned / cls_max optional.  h * w < 2^31.
__global__ void __launch_bounds__(CW_THREADS)
clean_walk_kernel(const float* __restrict__ m, const float* __restrict__ labels, float* __restrict__ cleaned,
                  unsigned* __restrict__ cls_max, int C, int h, int w, int H, int W, float sh, float sw, int nbands) {
    __shared__ float4 tab[CW_TAB];  // (i0, i1, l0, l1) of output row row_lo + k: the same for every thread of a band
    __shared__ unsigned red[CW_THREADS / 32];
    const size_t plane = blockIdx.y;
    const int t_lo = blockIdx.x * CW_THREADS, t = t_lo + threadIdx.x;
    const int row_lo = (t_lo / W) * CW_ROWS;
    const int row_hi = min(H, (min(t_lo + CW_THREADS - 1, nbands * W - 1) / W + 1) * CW_ROWS);
    const bool use_tab = row_hi - row_lo <= CW_TAB;  // narrow maps (a block spans many bands): rows computed in place
    if (use_tab) {
        for (int k = threadIdx.x; k < row_hi - row_lo; k += CW_THREADS) {
            const Lerp ly = make_lerp(row_lo + k, sh, h);
            tab[k] = make_float4(__int_as_float(ly.i0), __int_as_float(ly.i1), ly.l0, ly.l1);
        }
        __syncthreads();
    }
    unsigned best = 0u;
    if (t < nbands * W) {
        const int band = t / W, x = t - band * W;
        const int y1 = min(H, (band + 1) * CW_ROWS);
        const bool gated = labels != nullptr && (plane % C) > 0;
        const float g = gated ? __ldg(labels + (plane / C) * (size_t)(C - 1) + (plane % C - 1)) : 1.f;
        const float* __restrict__ pl = m + plane * h * w;
        float* __restrict__ dst = cleaned != nullptr ? cleaned + plane * H * W + x : nullptr;
        const Lerp lx = make_lerp(x, sw, w);
        ColumnWalk<1> cw;
        cw.reset();
        for (int y = band * CW_ROWS; y < y1; ++y) {
            Lerp ly;
            if (use_tab) {
                const float4 e = tab[y - row_lo];
                ly.i0 = __float_as_int(e.x); ly.i1 = __float_as_int(e.y); ly.l0 = e.z; ly.l1 = e.w;
            } else {
                ly = make_lerp(y, sh, h);
            }
            cw.advance(pl, 1, 0, w, ly, lx);
            float v = cw.value(0, ly);
            if (gated) v = __fmul_rn(v, g);
            if (dst != nullptr) dst[(size_t)y * W] = v;
            best = max(best, ordered_from_float(v));
        }
    }
    if (cls_max != nullptr) {
        best = __reduce_max_sync(0xffffffffu, best);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < CW_THREADS / 32) {
            best = __reduce_max_sync((1u << (CW_THREADS / 32)) - 1u, red[threadIdx.x]);
            if (threadIdx.x == 0 && best != 0u) atomicMax(cls_max + plane, best);
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

// The label pass of the resized epilogue as a column walk (see clean_walk_kernel): a thread keeps t0 / t1 of all
// C <= PW_CB classes in registers and walks `rows` output rows (chosen by the launcher so that the last wave of
// blocks is full).  Everything over the classes is branch-free (classes beyond C re-read class C-1 and carry an
// infinite threshold), so that the 2 x 21 loads of a new source row are all in flight before the first is used.
// grid: (ceil(nbands * W / PW_THREADS), B); C * h * w < 2^31.
// max.NaN: a NaN input gives NaN (torch.max propagates it, and so does the ordered-uint max of the other kernels)
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// The max pass of the resized epilogue for C <= PW_CB classes (_rescale_and_clean at a new size, SoftMaxAE.py:263-268,
// with the cleaned masks stored on request): same walk as the label pass below (one thread, all classes of a column
// band), a running maximum per class in registers, one atomicMax per class and block.  One class per thread
// (clean_walk_kernel, kept for more classes and for the plain resize) pays the row bookkeeping per class: 56 us
// against the 27 us of this kernel for the same values.  grid: (ceil(nbands * W / PW_THREADS), B)
__global__ void __launch_bounds__(PW_THREADS, 4)
class_max_walk_kernel(const float* __restrict__ m, const float* __restrict__ labels, float* __restrict__ cleaned,
                      unsigned* __restrict__ cls_max, int C, int h, int w, int H, int W, float sh, float sw, int rows,
                      int nbands) {
    __shared__ float gate[PW_CB];
    __shared__ unsigned red[PW_CB];
    const int b = blockIdx.y;
    if (threadIdx.x < PW_CB) {
        const int c = threadIdx.x;
        gate[c] = (labels != nullptr && c > 0 && c < C) ? __ldg(labels + (size_t)b * (C - 1) + (c - 1)) : 1.f;
        red[c] = 0u;
    }
    __syncthreads();
    const int hw = h * w;
    const size_t HW = (size_t)H * W;
    const int t = blockIdx.x * PW_THREADS + threadIdx.x;
    const bool valid = t < nbands * W;
    float best[PW_CB];
#pragma unroll
    for (int j = 0; j < PW_CB; ++j) best[j] = __int_as_float(0xff800000);  // -inf; every valid thread sees >= 1 row
    if (valid) {
        const int band = t / W, x = t - band * W;
        const int y1 = min(H, (band + 1) * rows);
        const float* __restrict__ pimg = m + (size_t)b * C * hw;
        const Lerp lx = make_lerp(x, sw, w);
        ColumnWalk<PW_CB> cw;
        cw.reset();
        for (int y = band * rows; y < y1; ++y) {
            const Lerp ly = make_lerp(y, sh, h);
            cw.advance(pimg, C, hw, w, ly, lx);
            float* __restrict__ dst = cleaned != nullptr ? cleaned + (size_t)b * C * HW + (size_t)y * W + x : nullptr;
#pragma unroll
            for (int j = 0; j < PW_CB; ++j) {  // gate 1 (class 0, no labels): fl(v * 1) == v
                const float v = __fmul_rn(cw.value(j, ly), gate[j]);
                if (dst != nullptr && j < C) dst[(size_t)j * HW] = v;
                best[j] = max_nan(best[j], v);
            }
        }
    }
    if (cls_max == nullptr) return;
#pragma unroll
    for (int j = 0; j < PW_CB; ++j) {
        const unsigned u = __reduce_max_sync(0xffffffffu, valid ? ordered_from_float(best[j]) : 0u);
        if ((threadIdx.x & 31) == 0 && u != 0u) atomicMax(&red[j], u);
    }
    __syncthreads();
    if (threadIdx.x < C && threadIdx.x < PW_CB && red[threadIdx.x] != 0u)
        atomicMax(cls_max + (size_t)b * C + threadIdx.x, red[threadIdx.x]);
}

__global__ void __launch_bounds__(PW_THREADS, 4)
pseudo_labels_walk_kernel(const float* __restrict__ m, const float* __restrict__ labels,
                          const unsigned* __restrict__ cls_max, uint8_t* __restrict__ label,
                          float* __restrict__ pseudo_gt, int* __restrict__ class_count, int C, int h, int w, int H,
                          int W, float sh, float sw, float bg_cut, float fg_cut, float low_cut, bool max_is_gated,
                          int rows, int nbands) {
    __shared__ float2 thr_gate[PW_CB];  // (threshold, label gate) per class
    __shared__ int cnt[PW_CB];
    const int b = blockIdx.y;
    if (threadIdx.x < PW_CB) {
        const int c = threadIdx.x;
        float g = 1.f, th = __int_as_float(0x7f800000);  // classes beyond C: never selected
        if (c < C) {  // thresholds exactly as in pseudo_labels_kernel
            g = (labels != nullptr && c > 0) ? __ldg(labels + (size_t)b * (C - 1) + (c - 1)) : 1.f;
            float mx = float_from_ordered(__ldg(cls_max + (size_t)b * C + c));
            if (!max_is_gated) mx = __fmul_rn(mx, g);
            th = fmaxf(__fmul_rn(mx, c == 0 ? bg_cut : fg_cut), low_cut);
        }
        thr_gate[c] = make_float2(th, g);
        cnt[c] = 0;
    }
    __syncthreads();
    const size_t HW = (size_t)H * W;
    const int hw = h * w;
    const int t = blockIdx.x * PW_THREADS + threadIdx.x;
    if (t < nbands * W) {
        const int band = t / W, x = t - band * W;
        const int y1 = min(H, (band + 1) * rows);
        const float* __restrict__ pimg = m + (size_t)b * C * hw;
        const Lerp lx = make_lerp(x, sw, w);
        ColumnWalk<PW_CB> cw;
        cw.reset();
        for (int y = band * rows; y < y1; ++y) {
            const Lerp ly = make_lerp(y, sh, h);
            cw.advance(pimg, C, hw, w, ly, lx);
            // x -> fl(x * 1) is the identity, so the un-gated classes (class 0, or no labels at all: gate 1) go
            // through the same multiply as the gated ones
            unsigned hits = 0u;
#pragma unroll
            for (int j = 0; j < PW_CB; ++j) {
                const float2 tg = thr_gate[j];
                if (__fmul_rn(cw.value(j, ly), tg.y) > tg.x) hits |= 1u << j;
            }
            const bool one = __popc(hits) == 1;
            const int first = __ffs(hits) - 1;
            const size_t i = (size_t)y * W + x;
            if (label != nullptr) label[(size_t)b * HW + i] = one ? (uint8_t)first : (uint8_t)255;
            if (pseudo_gt != nullptr)
                for (int c = 0; c < C; ++c) pseudo_gt[((size_t)b * C + c) * HW + i] = (one && c == first) ? 1.f : 0.f;
            if (class_count != nullptr && one) atomicAdd(&cnt[first], 1);
        }
    }
    if (class_count != nullptr) {
        __syncthreads();
        if (threadIdx.x < C && threadIdx.x < PW_CB && cnt[threadIdx.x] != 0)
            atomicAdd(class_count + (size_t)b * C + threadIdx.x, cnt[threadIdx.x]);
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

bool walk_all_classes_ok(int B, int C, int h, int w, int H, int W) {
    return C <= PW_CB && B <= 65535 && (long long)((H + 7) / 8) * W < (1ll << 30) && (long long)C * h * w < (1ll << 31);
}

}  // namespace

int walk_rows(int B, int H, int W, int threads) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || device_sm_count(dev, &sms) != PAMR_OK || sms <= 0) sms = 148;
    int rows = 8;
    long long best_cost = -1;
    for (int r = 8; r <= 32; ++r) {
        const long long blocks = ((long long)((H + r - 1) / r) * W + threads - 1) / threads * B;
        const long long cost = ((blocks + 4ll * sms - 1) / (4ll * sms)) * (r + 3);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; rows = r; }
    }
    return rows;
}

int launch_resize_bilinear(const float* src, float* dst, int n_planes, int h, int w, int H, int W, cudaStream_t s) {
    if (h == H && w == W) {
        PAMR_CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)n_planes * H * W, cudaMemcpyDeviceToDevice, s));
        return PAMR_OK;
    }
    dim3 block(EP_BX, EP_BY);
    const int nbands = (H + CW_ROWS - 1) / CW_ROWS;
    const bool walk = H >= 2 * h && (long long)nbands * W < (1ll << 30) && (long long)h * w < (1ll << 31);  // enlarging: source rows are re-used
    for (int n0 = 0; n0 < n_planes; n0 += 65535) {
        const int nn = min(65535, n_planes - n0);
        if (walk) {
            dim3 grid((unsigned)((nbands * W + CW_THREADS - 1) / CW_THREADS), nn);
            clean_walk_kernel<<<grid, CW_THREADS, 0, s>>>(src + (size_t)n0 * h * w, nullptr, dst + (size_t)n0 * H * W, nullptr,
                                                          1, h, w, H, W, scale_of(h, H), scale_of(w, W), nbands);
        } else {
            dim3 grid((W + EP_BX - 1) / EP_BX, (H + EP_BY - 1) / EP_BY, nn);
            if (grid.y > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "resize: H too large");
            resize_kernel<<<grid, block, 0, s>>>(src + (size_t)n0 * h * w, dst + (size_t)n0 * H * W, h, w, H, W,
                                                 scale_of(h, H), scale_of(w, W));
        }
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
    if (h == H && w == W) {
        dim3 block(EP_BX, EP_BY);
        dim3 grid((W + EP_BX - 1) / EP_BX, (H + EP_BY - 1) / EP_BY, B * C);
        if (grid.y > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "clean: H too large");
        clean_kernel<<<grid, block, 0, s>>>(m, labels, cleaned, cls_max, C, H, W);
    } else if (walk_all_classes_ok(B, C, h, w, H, W)) {
        const int rows = walk_rows(B, H, W, PW_THREADS), nbands = (H + rows - 1) / rows;
        dim3 grid((unsigned)((nbands * W + PW_THREADS - 1) / PW_THREADS), B);
        class_max_walk_kernel<<<grid, PW_THREADS, 0, s>>>(m, labels, cleaned, cls_max, C, h, w, H, W, scale_of(h, H),
                                                          scale_of(w, W), rows, nbands);
    } else {
        const int nbands = (H + CW_ROWS - 1) / CW_ROWS;
        if ((long long)nbands * W >= (1ll << 30) || (long long)h * w >= (1ll << 31))
            return set_error(PAMR_ERR_INVALID_ARGUMENT, "clean: map too large");
        dim3 grid((unsigned)((nbands * W + CW_THREADS - 1) / CW_THREADS), B * C);
        clean_walk_kernel<<<grid, CW_THREADS, 0, s>>>(m, labels, cleaned, cls_max, C, h, w, H, W, scale_of(h, H),
                                                      scale_of(w, W), nbands);
    }
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
    else if (walk_all_classes_ok(B, C, h, w, H, W)) {
        const int rows = walk_rows(B, H, W, PW_THREADS);
        const int nbands = (H + rows - 1) / rows;
        dim3 wgrid((unsigned)((nbands * W + PW_THREADS - 1) / PW_THREADS), B);
        pseudo_labels_walk_kernel<<<wgrid, PW_THREADS, 0, s>>>(m, labels, cls_max, label, pseudo_gt, class_count, C, h, w,
                                                               H, W, scale_of(h, H), scale_of(w, W), bg_cut, fg_cut,
                                                               low_cut, max_is_gated, rows, nbands);
    } else
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
