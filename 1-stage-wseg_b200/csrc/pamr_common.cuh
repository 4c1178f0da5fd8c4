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

// ---- affinity layout consumed by the tuned sm_100a propagation kernel ----
// Only for the standard 6 dilations (48 taps).  Three regions (AffTiling holds their bases):
//  * tiles, covering [0,Wt) x [0,Ht) (rounded up to whole tiles when a remainder is not split off): the tuned
//    kernel's thread (lane quarter wq, lane) of tile (ty,tx) owns R pixels (rows ty*4R + wq*R + i, column
//    tx*32 + lane) and keeps their 48*R weights in its Tensor Memory lane l = wq*32 + lane, at column s*R + i
//    ("sequence" order s below).  The weights travel global -> shared memory (cp.async.bulk) -> TMEM
//    (tcgen05.cp) without passing through registers, so the global layout IS the shared-memory image
//    tcgen05.cp reads: 16-byte pieces of 4 consecutive columns per lane,
//        [b][ty][tx][column / 4][l = 0..127][column % 4]
//    (128 lanes x 16 bytes = 2 KB per piece; 8 pieces = one 16 KB fill unit of 32 columns).  Pixels of partial
//    tiles that lie outside the image hold zeros;
//  * the column strip x = Wt = W-1 (one column, e.g. W = 321), y < Ht: the tiles on the right image border compute
//    it from their own shared-memory window.  LP = 32/R lanes share a strip pixel (48/LP taps each), and the
//    weights sit in spare Tensor Memory columns of those lanes, so this region is a tcgen05.cp image as well: per
//    (b, ty) a block of [k / 4][l = 0..127][k % 4] with k < 48/LP the lane's tap and l = wq*32 + i*LP + part the
//    lane that owns taps [part*48/LP, (part+1)*48/LP) of strip pixel (row wq*R + i of the tile);
//  * the row strip y in [Ht,H) (at most 8 rows), all x:  [b][y - Ht][s][x]  (coalesced for one warp per 32 pixels).
//
// Tap sequence s (0..47) <-> reference tap p = 8*id + j (pamr.py:25-34):
//   s <  12: centre column (b = 0): id = s/2, j = 1 (dy=-d) or 6 (dy=+d)
//   s >= 12: side group g = (s-12)/3 = 6*bi + id (bi = 0: dx=-d, bi = 1: dx=+d), a = (s-12)%3 -> dy = (a-1)*d
__host__ __device__ constexpr int seq_tap(int s) {
    if (s < 12) return 8 * (s / 2) + ((s % 2 == 0) ? 1 : 6);
    const int t = s - 12, g = t / 3, a = t % 3, bi = g / 6, id = g % 6;
    const int j = (bi == 0) ? (a == 0 ? 0 : a == 1 ? 3 : 5) : (a == 0 ? 2 : a == 1 ? 4 : 7);
    return 8 * id + j;
}
__host__ __device__ constexpr int tap_seq(int p) {
    const int id = p / 8, j = p % 8;
    if (j == 1) return 2 * id;
    if (j == 6) return 2 * id + 1;
    const int bi = (j == 0 || j == 3 || j == 5) ? 0 : 1;
    const int a = (j == 0 || j == 2) ? 0 : (j == 3 || j == 4) ? 1 : 2;
    return 12 + (bi * 6 + id) * 3 + a;
}
struct AffTiling {
    int R, tiles_x, tiles_y;  // R rows per thread (tile = 32 x 4R); R == 0: standard [B,48,H,W] layout
    int Wt, Ht;               // extent covered by tiles; x >= Wt is the column strip, y >= Ht the row strip
    int W, H;
    size_t cs_base, rs_base;  // float offsets of the column-strip and row-strip regions
    size_t floats;            // total size of the layout
};
// offset of TMEM column `col` of a lane relative to the lane's first weight
__host__ __device__ __forceinline__ size_t aff_tiled_col_offset(int col) { return ((size_t)(col >> 2) << 9) + (size_t)(col & 3); }
__host__ __device__ __forceinline__ size_t aff_tile_floats(int R) { return (size_t)48 * R * 128; }
// column strip: lanes per strip pixel, taps per lane, floats of one (b, ty) block
__host__ __device__ constexpr int aff_cs_lanes(int R) { return 32 / R; }
__host__ __device__ constexpr int aff_cs_taps(int R) { return 48 / (32 / R); }
__host__ __device__ constexpr size_t aff_cs_block_floats(int R) { return (size_t)aff_cs_taps(R) * 128; }
// fills the derived fields (region bases, size) for a batch of B
__host__ __device__ __forceinline__ void aff_layout_finish(AffTiling& t, int B) {
    const size_t tiled = (size_t)B * t.tiles_y * t.tiles_x * aff_tile_floats(t.R);
    const size_t cs = (t.W > t.Wt) ? (size_t)B * t.tiles_y * aff_cs_block_floats(t.R) : 0;
    const size_t rs = (size_t)B * (t.H - t.Ht) * 48 * t.W;
    t.cs_base = tiled;
    t.rs_base = tiled + cs;
    t.floats = tiled + cs + rs;
}
// Where the 48 weights of pixel (y, x) live: weight s is at base[s * stride] (row strip, stride > 0), at
// base[aff_tiled_col_offset(s*R + i)] (tile pixels, stride == 0) or in the column strip's image (stride < 0).
struct AffPixel {
    size_t base;
    int stride, i, R;
    __host__ __device__ __forceinline__ size_t at(int s) const {
        if (stride > 0) return base + (size_t)s * stride;
        if (stride == 0) return base + aff_tiled_col_offset(s * R + i);
        const int tpl = aff_cs_taps(R), part = s / tpl, k = s % tpl;  // base = block + lane of part 0
        return base + (size_t)part * 4 + aff_tiled_col_offset(k);
    }
};
__host__ __device__ __forceinline__ AffPixel aff_pixel(const AffTiling& t, int b, int y, int x) {
    AffPixel p;
    p.R = t.R;
    p.i = 0;
    const int ty4 = 4 * t.R;
    if (y >= t.Ht) {  // row strip (including its corner with the column strip)
        p.base = t.rs_base + ((size_t)b * (t.H - t.Ht) + (y - t.Ht)) * 48 * t.W + x;
        p.stride = t.W;
    } else if (x >= t.Wt) {  // column strip (one column)
        const int ry = y % ty4;
        p.base = t.cs_base + ((size_t)b * t.tiles_y + y / ty4) * aff_cs_block_floats(t.R) +
                 (size_t)((ry / t.R) * 32 + (ry % t.R) * aff_cs_lanes(t.R)) * 4;
        p.stride = -1;
    } else {
        const int ry = y % ty4, wq = ry / t.R;
        const size_t tile = ((size_t)b * t.tiles_y + y / ty4) * t.tiles_x + (x >> 5);
        p.base = tile * aff_tile_floats(t.R) + (size_t)(wq * 32 + (x & 31)) * 4;
        p.stride = 0;
        p.i = ry % t.R;
    }
    return p;
}

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

#ifdef __CUDACC__
// torch's align_corners=True source index / weights (area_pixel_compute_scale,
// guard_index_and_lambda), float arithmetic without FMA contraction.
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp make_lerp(int dst, float scale, int in_size) {
    Lerp r;
    const float f = __fmul_rn(scale, (float)dst);
    r.i0 = min((int)f, in_size - 1);
    r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
    r.l1 = fminf(fmaxf(__fsub_rn(f, (float)r.i0), 0.f), 1.f);
    r.l0 = __fsub_rn(1.f, r.l1);
    return r;
}
__device__ __forceinline__ float bilerp(const float* __restrict__ pl, int w, const Lerp& ly, const Lerp& lx) {
    const float p00 = __ldg(pl + (size_t)ly.i0 * w + lx.i0), p01 = __ldg(pl + (size_t)ly.i0 * w + lx.i1);
    const float p10 = __ldg(pl + (size_t)ly.i1 * w + lx.i0), p11 = __ldg(pl + (size_t)ly.i1 * w + lx.i1);
    const float t0 = __fadd_rn(__fmul_rn(lx.l0, p00), __fmul_rn(lx.l1, p01));
    const float t1 = __fadd_rn(__fmul_rn(lx.l0, p10), __fmul_rn(lx.l1, p11));
    return __fadd_rn(__fmul_rn(ly.l0, t0), __fmul_rn(ly.l1, t1));
}
__host__ __device__ __forceinline__ float scale_of(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}
// Column walk over an enlarged map.  bilerp() is  v = ly.l0 * t0 + ly.l1 * t1  with
//     t0 = lx.l0 * p[i0][x0] + lx.l1 * p[i0][x1],   t1 = the same on source row i1,
// and t0 / t1 depend on the SOURCE row only: a thread that owns one output column and walks down its rows recomputes
// them when the source row pair changes (every k-th output row of a k-fold enlargement; the new t0 usually is the old
// t1) and pays one FMUL/FMUL/FADD per output value otherwise.  CB planes (classes) of one image are carried together;
// planes beyond C re-read plane C-1, so that everything over the planes is branch-free and the 2 x CB loads of a new
// source row are in flight together.  value() is the same expression in the same order as bilerp(): bit-identical.
template <int CB>
struct ColumnWalk {
    int i0, i1;  // source rows held in t0 / t1
    float t0[CB], t1[CB];
    __device__ __forceinline__ void reset() {
        i0 = i1 = -1;
#pragma unroll
        for (int j = 0; j < CB; ++j) t0[j] = t1[j] = 0.f;
    }
    // x-interpolated values of source row `row` (off = row * w) of the planes pimg + min(j, C-1) * hw
    static __device__ __forceinline__ void load_row(const float* __restrict__ pimg, int C, int hw, int off, const Lerp& lx,
                                                    float (&tt)[CB]) {
        float a[CB], c[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const int o = min(j, C - 1) * hw + off;
            a[j] = __ldg(pimg + o + lx.i0);
            c[j] = __ldg(pimg + o + lx.i1);
        }
#pragma unroll
        for (int j = 0; j < CB; ++j) tt[j] = __fadd_rn(__fmul_rn(lx.l0, a[j]), __fmul_rn(lx.l1, c[j]));
    }
    // make t0 / t1 the rows ly needs (C * hw < 2^31)
    __device__ __forceinline__ void advance(const float* __restrict__ pimg, int C, int hw, int w, const Lerp& ly, const Lerp& lx) {
        if (ly.i0 == i0 && ly.i1 == i1) return;
        if (ly.i0 == i1) {
#pragma unroll
            for (int j = 0; j < CB; ++j) t0[j] = t1[j];
        } else {
            load_row(pimg, C, hw, ly.i0 * w, lx, t0);
        }
        if (ly.i1 == ly.i0) {
#pragma unroll
            for (int j = 0; j < CB; ++j) t1[j] = t0[j];
        } else {
            load_row(pimg, C, hw, ly.i1 * w, lx, t1);
        }
        i0 = ly.i0;
        i1 = ly.i1;
    }
    __device__ __forceinline__ float value(int j, const Lerp& ly) const {
        return __fadd_rn(__fmul_rn(ly.l0, t0[j]), __fmul_rn(ly.l1, t1[j]));
    }
};
// Rows per band for the walks that carry all classes per thread (4 blocks of `threads` per SM): the candidate with the
// cheapest schedule, waves x (rows + the loads of a band's first row pair, worth about 3 rows).
int walk_rows(int B, int H, int W, int threads);
#endif

#ifdef __CUDACC__
// x / y for y in the normal range, given r = rn(1/y): q = rn(x*r) corrected by the exact residual (Markstein);
// returns the IEEE quotient for the operand ranges of the affinity at a third of the cost of the division routine.
__device__ __forceinline__ float div_markstein(float x, float y, float r) {
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, y, x);
    return __fmaf_rn(e, r, q);
}
#endif

// pamr_resident.cu: affinity + all iterations in one launch for small maps (stage_net's real call shapes)
struct ResidentPlan {
    bool ok;                // the resident kernel applies to this problem
    int images_per_launch;  // images whose row blocks are co-resident in one launch
    size_t scratch_bytes;   // ping-pong buffers + barrier counters
};
ResidentPlan resident_plan(int B, int C, int H, int W, const Dilations& dil, int iters, int dev);
int launch_resident(const float* img, int K, const float* m_in, float* m_out, void* scratch, size_t scratch_bytes, int B,
                    int C, int H, int W, const Dilations& dil, int iters, unsigned* cls_max, int dev, cudaStream_t s);

// pamr_propagate_sm100.cu: 1-D TMA descriptor over a flat fp32 array (map points at a CUtensorMap)
int encode_tensor_map_1d_f32(void* map, const float* base, unsigned long long elems, unsigned box);
int device_sm_count(int dev, int* out);
// pamr_affinity.cu: LocalStDev (pamr.py:77-103) on its own, img [B,K,H,W] -> sd [B,K,H,W]
int launch_local_std(const float* img, float* sd, int B, int K, int H, int W, const Dilations& dil, cudaStream_t s);

// Kernel launchers implemented in the .cu files (all enqueue on `s`, return a PAMR_* code).
int launch_resize_bilinear(const float* src, float* dst, int n_planes, int h, int w, int H, int W, cudaStream_t s);
// img_pitched: nullptr, or device scratch of affinity_pitched_image_bytes() for the tile kernel's copy of an image whose
// rows are not 16-byte aligned (only read when affinity_image_needs_pitching())
int launch_affinity(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil,
                    const AffTiling& tiling, float* img_pitched, cudaStream_t s, cudaStream_t strips_stream = nullptr);
bool affinity_image_needs_pitching(const float* img, int W);
size_t affinity_pitched_image_bytes(int B, int K, int H, int W);
int launch_aff_relayout(const float* aff_std, float* aff_tiled, int B, int H, int W, const AffTiling& tiling,
                        cudaStream_t s);
// Tiling of the tuned propagation kernel for this problem; R == 0 when only the generic kernel applies.
AffTiling tuned_tiling(int B, int C, int H, int W, const Dilations& dil);
size_t propagate_scratch_bytes(int B, int C, int H, int W, const Dilations& dil, int iters, bool aff_is_tiled);
int launch_propagate(const float* aff, bool aff_is_tiled, const float* m_in, float* m_out, void* scratch,
                     size_t scratch_bytes, int B, int C, int H, int W, const Dilations& dil, int iters,
                     unsigned* cls_max, int dev, cudaStream_t s);
int launch_affinity_propagate(const float* img, int K, float* aff_out, float* img_pitched, const float* aff_in, bool aff_is_tiled,
                              const float* m_in, float* m_out, void* scratch, size_t scratch_bytes, int B, int C, int H,
                              int W, const Dilations& dil, int iters, unsigned* cls_max, int dev, cudaStream_t s);
int launch_clean(const float* m, const float* labels, float* cleaned, unsigned* cls_max, int B, int C, int h, int w,
                 int H, int W, cudaStream_t s);
int launch_denorm_resize(const float* src, float* dst, const float* mean_host, const float* std_host, int B, int K, int h,
                         int w, int H, int W, cudaStream_t s);
int launch_merge_multiscale(const float* masks, const int* pads_host, const float* labels, float* merged, uint8_t* pred,
                            int S, int C, int Hp, int Wp, int H, int W, int flip, float bg_pow, float thresh,
                            cudaStream_t s);
// pamr_loss.cu (balanced_mask_loss_ce)
size_t mask_ce_workspace_bytes(int B, int C, int h, int w, int H, int W);
int launch_labels_from_onehot(const float* pseudo_gt, uint8_t* label, int* class_count, int B, int C, int H, int W,
                              cudaStream_t s);
int launch_mask_ce_forward(const float* logits, const uint8_t* label, const int* class_count, const float* gt_labels,
                           float* loss, void* ws, size_t ws_bytes, int B, int C, int h, int w, int H, int W,
                           cudaStream_t s);
int launch_mask_ce_backward(const float* logits, const uint8_t* label, const float* grad_loss, float* grad_logits,
                            const void* ws, size_t ws_bytes, int B, int C, int h, int w, int H, int W, cudaStream_t s);
int launch_pseudo_labels(const float* m, const float* labels, const unsigned* cls_max, uint8_t* label,
                         float* pseudo_gt, int* class_count, int B, int C, int h, int w, int H, int W, float bg_cut,
                         float fg_cut, float low_cut, bool max_is_gated, cudaStream_t s);

}  // namespace pamr
