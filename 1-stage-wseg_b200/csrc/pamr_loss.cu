// Class-balanced cross-entropy on the pseudo-labels and its gradient.
// Replaces (paths relative to the reference repo root)  models/SoftMaxAE.py:52-88  balanced_mask_loss_ce:
//   z     = bilinear(mask logits -> label resolution, align_corners=True)          (:58)  never materialised
//   label = argmax pseudo_gt / 255 where no class is set                           (:61-66) = the uint8 label map
//   n[b,c], cw[b,c] = (tot - n)/(1 + tot), bw[b] = (sum gt_labels + 1 == #present)  (:71-74, :82-84)
//   loss[b] = bw[b] * mean_px( cw[b,label] * (logsumexp_c z - z[label]) )           (:77, :86)
// and d(sum_b g[b] loss[b]) / d logits through the transpose of the interpolation.
//
// Forward (one launch + one small memset): the C interpolated logits of a label pixel go through an online
// log-sum-exp in base 2 (ex2.approx, 2 ulp); the class weights come from the counts (per block, in shared memory);
// per-image sum in double (warp shuffle -> block -> one atomicAdd(double) per block) and the block that arrives last
// at the image's ticket writes loss[b].  lse(px) and coef(px) = cw[label]/(H*W) (0 where ignored) stay in the
// workspace for the backward pass.  One thread per label pixel (ce_forward_kernel), or -- logits enlarged at least
// 2x, the training shapes -- a column walk over the label rows (ce_forward_walk_kernel).
// Backward, deterministic, no atomics:
//   same resolution:  one thread per pixel, grad_c = g * coef * (2^((z_c - lse) log2 e) - [c == label]);
//   with resampling the transposed bilinear interpolation is separable and runs as two passes:
//   * logits enlarged at least 2x: y first (ce_backward_ywalk_kernel, ce_backward_xgather_kernel, described there);
//   * otherwise x first, two gathers --
//       T[b,c,y,j]    = sum_x coef(y,x) * (softmax_c(y,x) - [c == label(y,x)]) * wx(x -> j)    (label rows, logit columns)
//       grad[b,c,i,j] = g_b * sum_y T[b,c,y,j] * wy(y -> i)
//     the first with the six logits a thread needs held in registers (its source column and the two next to it, on
//     the two source rows of label row y), so that the inner loop over the ~2/scale pixels of its footprint has no
//     dependent loads.  (Round 1 visited the whole 2-D footprint per logit element and recomputed bilerp + expf for
//     each of its ~(2/scale)^2 pixels: 0.55 ms at 81x81 -> 321x321.)
#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int LS_BX = 32, LS_BY = 8;
constexpr int CE_BATCH = 7;  // class planes loaded per batch in the forward kernel (21 = 3 x 7)

struct CeWorkspace {
    double* acc;     // [B]    sum_px cw*ce
    unsigned* ticket;// [B]    blocks of image b that have added their sum
    float* bw;       // [B]    batch weight (written by the forward pass)
    float* lse;      // [B,H,W]
    float* coef;     // [B,H,W]
    float* T;        // backward with resampling: the gradient after the first pass, [B,C,H,w] (x first) or [B,C,h,W] (y first)
    size_t head_bytes;  // acc + ticket: zeroed by one memset before the forward kernel
};
__host__ __device__ inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
inline CeWorkspace carve(void* ws, int B, int H, int W) {
    char* p = (char*)ws;
    CeWorkspace r;
    r.acc = (double*)p; p += up256(sizeof(double) * B);
    r.ticket = (unsigned*)p; p += up256(sizeof(unsigned) * B);
    r.head_bytes = (size_t)(p - (char*)ws);
    r.bw = (float*)p; p += up256(sizeof(float) * B);
    r.lse = (float*)p; p += up256(sizeof(float) * (size_t)B * H * W);
    r.coef = (float*)p; p += up256(sizeof(float) * (size_t)B * H * W);
    r.T = (float*)p;
    return r;
}

__device__ __forceinline__ float ce_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float CE_LOG2E = 1.4426950408889634f;

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
        // histogram of the block in shared memory (one shared-memory atomic per distinct label of a warp), then one
        // global atomic per class the block has seen
        __shared__ int hist[256];
        const int tid = threadIdx.y * LS_BX + threadIdx.x;
        hist[tid] = 0;  // LS_BX * LS_BY == 256
        __syncthreads();
        const unsigned peers = __match_any_sync(0xffffffffu, lab);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1 && lab < C) atomicAdd(&hist[lab], __popc(peers));
        __syncthreads();
        if (tid < C && hist[tid] != 0) atomicAdd(count + (size_t)b * C + tid, hist[tid]);
    }
}

template <bool kResize>
__device__ __forceinline__ float logit_at(const float* __restrict__ pl, int w, size_t i, const Lerp& ly, const Lerp& lx) {
    return kResize ? bilerp(pl, w, ly, lx) : __ldg(pl + i);
}

constexpr int CE_MAXC_SMEM = 256;  // labels are uint8: C <= 255

// grid (tiles_x, tiles_y, B)
template <bool kResize>
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_forward_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const int* __restrict__ count,
                  const float* __restrict__ gt_labels, double* __restrict__ acc, unsigned* __restrict__ ticket,
                  float* __restrict__ bw_out, float* __restrict__ loss, float* __restrict__ lse_out,
                  float* __restrict__ coef_out, int C, int h, int w, int H, int W, float sh, float sw, double inv_hw) {
    const int x = blockIdx.x * LS_BX + threadIdx.x, y = blockIdx.y * LS_BY + threadIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W, hw = (size_t)h * w, i = (size_t)y * W + x;
    // class weights of this image (SoftMaxAE.py:71-74), float arithmetic as in the reference
    __shared__ float s_cw[CE_MAXC_SMEM];
    __shared__ float s_tot;
    const int tid = threadIdx.y * LS_BX + threadIdx.x;
    if (tid == 0) {
        float tot = 0.f;
        for (int c = 0; c < C; ++c) tot = __fadd_rn(tot, (float)count[(size_t)b * C + c]);
        s_tot = tot;
    }
    __syncthreads();
    for (int c = tid; c < C; c += LS_BX * LS_BY)
        s_cw[c] = __fdiv_rn(__fsub_rn(s_tot, (float)count[(size_t)b * C + c]), __fadd_rn(1.f, s_tot));
    __syncthreads();
    double term = 0.0;
    if (x < W && y < H) {
        const int lab = label[(size_t)b * HW + i];
        float lse = 0.f, coef = 0.f;
        if (lab < C) {
            const Lerp ly = make_lerp(y, sh, h), lx = make_lerp(x, sw, w);
            const float* __restrict__ base = logits + (size_t)b * C * hw;
            float m = -INFINITY, s = 0.f, zl = 0.f;  // online log-sum-exp, base 2: s = sum 2^((z - m) log2 e)
            // classes in batches of CE_BATCH: the loads of a batch are issued before the dependent chain
            for (int c0 = 0; c0 < C; c0 += CE_BATCH) {
                float vb[CE_BATCH];
#pragma unroll
                for (int j = 0; j < CE_BATCH; ++j)
                    vb[j] = logit_at<kResize>(base + (size_t)min(c0 + j, C - 1) * hw, w, i, ly, lx);
                float bm = vb[0];
#pragma unroll
                for (int j = 1; j < CE_BATCH; ++j) bm = fmaxf(bm, (c0 + j < C) ? vb[j] : vb[0]);
                if (bm > m) {  // one rescale per batch
                    s *= ce_ex2((m - bm) * CE_LOG2E);
                    m = bm;
                }
#pragma unroll
                for (int j = 0; j < CE_BATCH; ++j) {
                    const int c = c0 + j;
                    if (c < C) {
                        if (c == lab) zl = vb[j];
                        s += ce_ex2((vb[j] - m) * CE_LOG2E);
                    }
                }
            }
            lse = m + logf(s);
            const float cw = s_cw[lab];
            term = (double)cw * (double)(lse - zl);
            coef = cw / (float)HW;
        }
        lse_out[(size_t)b * HW + i] = lse;
        coef_out[(size_t)b * HW + i] = coef;
    }
    // block sum in double -> one atomic per block; the image's last block finishes the loss
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    __shared__ double red[LS_BY];
    if (threadIdx.x == 0) red[threadIdx.y] = term;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < LS_BY; ++k) t += red[k];
        if (t != 0.0) atomicAdd(acc + b, t);
        __threadfence();
        const unsigned done = atomicAdd(ticket + b, 1u) + 1u;
        if (done == gridDim.x * gridDim.y) {
            __threadfence();
            const double total = atomicAdd(acc + b, 0.0);  // every block's sum is in
            int present = 0;
            for (int c = 0; c < C; ++c) present += count[(size_t)b * C + c] > 0;
            float gsum = 1.f;  // + BG (SoftMaxAE.py:82-84)
            for (int c = 0; c < C - 1; ++c) gsum = __fadd_rn(gsum, gt_labels[(size_t)b * (C - 1) + c]);
            const float bw = (gsum == (float)present) ? 1.f : 0.f;
            bw_out[b] = bw;
            loss[b] = bw * (float)(total * inv_hw);
        }
    }
}

// Forward with up-sampled logits (the training shapes: logits 81x81 / 41x41, labels 321x321) as a COLUMN WALK, like
// the resized epilogue (ColumnWalk, pamr_common.cuh): a thread owns one label column of a band of label rows and keeps,
// for all C <= CW_CB classes, the x-interpolated logits of the two source rows in registers; a label row whose source
// row pair is unchanged costs one FMUL/FMUL/FADD per class instead of four loads and nine operations.
// Per pixel the values and the log-sum-exp recurrence are those of ce_forward_kernel (same expressions, same batches
// of CE_BATCH classes), so lse / coef are bit-identical; only the order of the double sum differs.
// grid (ceil(nbands * W / CW_THREADS), B); C * h * w < 2^31
constexpr int CW_THREADS = 128;
constexpr int CW_CB = 21;
static_assert(CW_CB % CE_BATCH == 0, "whole batches");
__global__ void __launch_bounds__(CW_THREADS, 4)
ce_forward_walk_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const int* __restrict__ count,
                       const float* __restrict__ gt_labels, double* __restrict__ acc, unsigned* __restrict__ ticket,
                       float* __restrict__ bw_out, float* __restrict__ loss, float* __restrict__ lse_out,
                       float* __restrict__ coef_out, int C, int h, int w, int H, int W, float sh, float sw, double inv_hw,
                       int rows, int nbands) {
    const int b = blockIdx.y, tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    const int hw = h * w;
    __shared__ float s_cw[CW_CB];
    __shared__ float s_tot;
    __shared__ double red[CW_THREADS / 32];
    if (tid == 0) {  // class weights of this image (SoftMaxAE.py:71-74), as in ce_forward_kernel
        float tot = 0.f;
        for (int c = 0; c < C; ++c) tot = __fadd_rn(tot, (float)count[(size_t)b * C + c]);
        s_tot = tot;
    }
    __syncthreads();
    if (tid < C) s_cw[tid] = __fdiv_rn(__fsub_rn(s_tot, (float)count[(size_t)b * C + tid]), __fadd_rn(1.f, s_tot));
    __syncthreads();
    double term = 0.0;
    const int t = blockIdx.x * CW_THREADS + tid;
    if (t < nbands * W) {
        const int band = t / W, x = t - band * W;
        const int y1 = min(H, (band + 1) * rows);
        const float* __restrict__ pimg = logits + (size_t)b * C * hw;
        const Lerp lx = make_lerp(x, sw, w);
        ColumnWalk<CW_CB> cw;
        cw.reset();
        const size_t col = (size_t)b * HW + x;
        int plab[2];  // the label is fetched two rows ahead (a dependent round trip to L2 per row otherwise)
#pragma unroll
        for (int k = 0; k < 2; ++k) plab[k] = (band * rows + k < y1) ? (int)label[col + (size_t)(band * rows + k) * W] : 255;
        for (int y = band * rows; y < y1; ++y) {
            const Lerp ly = make_lerp(y, sh, h);
            cw.advance(pimg, C, hw, w, ly, lx);
            const size_t p = col + (size_t)y * W;
            const int lab = plab[0];
            plab[0] = plab[1];
            if (y + 2 < y1) plab[1] = label[col + (size_t)(y + 2) * W];
            float lse = 0.f, coef = 0.f;
            if (lab < C) {
                float m = -INFINITY, s = 0.f, zl = 0.f;
#pragma unroll
                for (int c0 = 0; c0 < CW_CB; c0 += CE_BATCH) {
                    float vb[CE_BATCH];
#pragma unroll
                    for (int j = 0; j < CE_BATCH; ++j) {  // classes beyond C: -inf, which leaves the maximum and the sum alone
                        vb[j] = (c0 + j < C) ? cw.value(c0 + j, ly) : -INFINITY;
                    }
                    float bm = vb[0];
#pragma unroll
                    for (int j = 1; j < CE_BATCH; ++j) bm = fmaxf(bm, vb[j]);
                    if (bm > m) {  // one rescale per batch
                        s *= ce_ex2((m - bm) * CE_LOG2E);
                        m = bm;
                    }
#pragma unroll
                    for (int j = 0; j < CE_BATCH; ++j) {
                        if (c0 + j == lab) zl = vb[j];
                        s += ce_ex2((vb[j] - m) * CE_LOG2E);
                    }
                }
                lse = m + logf(s);
                const float cw = s_cw[lab];
                term += (double)cw * (double)(lse - zl);
                coef = cw / (float)HW;
            }
            lse_out[p] = lse;
            coef_out[p] = coef;
        }
    }
    // block sum in double -> one atomic per block; the image's last block finishes the loss
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((tid & 31) == 0) red[tid >> 5] = term;
    __syncthreads();
    if (tid == 0) {
        double tsum = 0.0;
        for (int k = 0; k < CW_THREADS / 32; ++k) tsum += red[k];
        if (tsum != 0.0) atomicAdd(acc + b, tsum);
        __threadfence();
        const unsigned done = atomicAdd(ticket + b, 1u) + 1u;
        if (done == gridDim.x) {
            __threadfence();
            const double total = atomicAdd(acc + b, 0.0);  // every block's sum is in
            int present = 0;
            for (int c = 0; c < C; ++c) present += count[(size_t)b * C + c] > 0;
            float gsum = 1.f;  // + BG (SoftMaxAE.py:82-84)
            for (int c = 0; c < C - 1; ++c) gsum = __fadd_rn(gsum, gt_labels[(size_t)b * (C - 1) + c]);
            const float bw = (gsum == (float)present) ? 1.f : 0.f;
            bw_out[b] = bw;
            loss[b] = bw * (float)(total * inv_hw);
        }
    }
}

// Backward without resampling: one thread per pixel.  grid (tiles_x, tiles_y, B)
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_backward_same_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const float* __restrict__ bw,
                        const float* __restrict__ lse_in, const float* __restrict__ coef_in, const float* __restrict__ gout,
                        float* __restrict__ grad, int C, int H, int W) {
    const int x = blockIdx.x * LS_BX + threadIdx.x, y = blockIdx.y * LS_BY + threadIdx.y, b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W, i = (size_t)y * W + x, p = (size_t)b * HW + i;
    const float gc = gout[b] * bw[b] * __ldg(coef_in + p);
    const float nl = -__ldg(lse_in + p) * CE_LOG2E;
    const int lab = label[p];
    const float* __restrict__ zb = logits + (size_t)b * C * HW + i;
    float* __restrict__ gb = grad + (size_t)b * C * HW + i;
    for (int c0 = 0; c0 < C; c0 += CE_BATCH) {
        float vb[CE_BATCH];
#pragma unroll
        for (int j = 0; j < CE_BATCH; ++j) vb[j] = __ldg(zb + (size_t)min(c0 + j, C - 1) * HW);
#pragma unroll
        for (int j = 0; j < CE_BATCH; ++j) {
            const int c = c0 + j;
            if (c < C) gb[(size_t)c * HW] = (gc == 0.f) ? 0.f : gc * (ce_ex2(fmaf(vb[j], CE_LOG2E, nl)) - (c == lab ? 1.f : 0.f));
        }
    }
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

// Backward with resampling, x pass: T[b,c,y,j].  grid (ceil(w/32), ceil(H/8), B * ceil(C/CE_BATCH)): a thread takes
// CE_BATCH classes of one (y, j), so that what does not depend on the class (the pixel's interpolation weights, coef,
// lse, label) is computed once per pixel for all of them.  Per class it holds r[k] = the logit column j-1+k
// interpolated in y (k = 0..2); a pixel of the footprint then needs one FMA pair for its interpolated logit.
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_backward_xpass_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const float* __restrict__ lse_in,
                         const float* __restrict__ coef_in, float* __restrict__ T, int C, int h, int w, int H, int W, float sh,
                         float sw) {
    const int j = blockIdx.x * LS_BX + threadIdx.x, y = blockIdx.y * LS_BY + threadIdx.y;
    const int nbatch = (C + CE_BATCH - 1) / CE_BATCH;
    const int b = blockIdx.z / nbatch, c0 = (blockIdx.z - b * nbatch) * CE_BATCH;
    if (j >= w || y >= H) return;
    const size_t HW = (size_t)H * W, hw = (size_t)h * w;
    const Lerp ly = make_lerp(y, sh, h);
    const int jm = max(j - 1, 0), jp = min(j + 1, w - 1);
    float r[CE_BATCH][3], acc[CE_BATCH];
#pragma unroll
    for (int q = 0; q < CE_BATCH; ++q) {
        const float* __restrict__ pl = logits + ((size_t)b * C + min(c0 + q, C - 1)) * hw;
        const float* __restrict__ p0 = pl + (size_t)ly.i0 * w;
        const float* __restrict__ p1 = pl + (size_t)ly.i1 * w;
        r[q][0] = fmaf(ly.l1, __ldg(p1 + jm), ly.l0 * __ldg(p0 + jm));
        r[q][1] = fmaf(ly.l1, __ldg(p1 + j), ly.l0 * __ldg(p0 + j));
        r[q][2] = fmaf(ly.l1, __ldg(p1 + jp), ly.l0 * __ldg(p0 + jp));
        acc[q] = 0.f;
    }
    int xlo, xhi;
    footprint(j, sw, W, xlo, xhi);
    const size_t row = (size_t)b * HW + (size_t)y * W;
    for (int x = xlo; x <= xhi; ++x) {
        const Lerp lx = make_lerp(x, sw, w);
        const float wx = weight_to(lx, j);
        const float cwx = __ldg(coef_in + row + x) * wx;
        if (cwx == 0.f) continue;
        const float nl = -__ldg(lse_in + row + x) * CE_LOG2E;
        const int lab = (int)label[row + x] - c0;
        // lx.i0 is j-1 or j; lx.i1 is lx.i0 + 1, or lx.i0 on the last column (where r[.][2] == r[.][1] already)
        const bool left = lx.i0 < j;
#pragma unroll
        for (int q = 0; q < CE_BATCH; ++q) {
            const float z = fmaf(lx.l1, left ? r[q][1] : r[q][2], lx.l0 * (left ? r[q][0] : r[q][1]));
            acc[q] = fmaf(cwx, ce_ex2(fmaf(z, CE_LOG2E, nl)) - (lab == q ? 1.f : 0.f), acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < CE_BATCH; ++q)
        if (c0 + q < C) T[(((size_t)b * C + c0 + q) * H + y) * w + j] = acc[q];
}

// y pass: grad[b,c,i,j] = g_b * bw_b * sum_y T[b,c,y,j] * wy(y -> i).  grid (ceil(w/32), ceil(h/8), B*C)
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_backward_ypass_kernel(const float* __restrict__ T, const float* __restrict__ bw, const float* __restrict__ gout,
                         float* __restrict__ grad, int C, int h, int w, int H, float sh) {
    const int j = blockIdx.x * LS_BX + threadIdx.x, i = blockIdx.y * LS_BY + threadIdx.y;
    const int plane = blockIdx.z, b = plane / C;
    if (j >= w || i >= h) return;
    const float g = gout[b] * bw[b];
    float acc = 0.f;
    if (g != 0.f) {
        int ylo, yhi;
        footprint(i, sh, H, ylo, yhi);
        const float* __restrict__ tp = T + (size_t)plane * H * w + j;
        for (int y = ylo; y <= yhi; ++y) {
            const float wy = weight_to(make_lerp(y, sh, h), i);
            if (wy != 0.f) acc = fmaf(wy, __ldg(tp + (size_t)y * w), acc);
        }
    }
    grad[(size_t)plane * h * w + (size_t)i * w + j] = g * acc;
}

// ---- backward with ENLARGED logits (H >= 2h: the training shapes), y first ----
//   U[b,c,i,x]    = sum_y coef(y,x) * (softmax_c(y,x) - [c == label(y,x)]) * wy(y -> i)      (logit rows, label columns)
//   grad[b,c,i,j] = g_b * bw_b * sum_x U[b,c,i,x] * wx(x -> j)
// The first pass is a column walk (ColumnWalk, pamr_common.cuh) like the forward pass: a thread owns one label
// column x, BA_CB classes and a band of `srows` logit rows; it walks down the label rows whose source row pair
// touches its band, computes each pixel's softmax term ONCE (the x-first order above evaluates it for both logit
// columns a pixel feeds, and re-interpolates the logits per pixel) from the same z as the forward pass, and adds
// wy.l0 * G to the accumulator of logit row i0 and wy.l1 * G to that of row i1, both in registers; a row's sum is
// stored when the walk leaves it.  Deterministic without atomics: every U element has one owner, which also walks
// the label rows just above its band (their i1 is the band's first row) -- (srows + 1) / srows of the work; the launcher
// picks srows so that the blocks (4 per SM) fill whole waves.
constexpr int BA_THREADS = 128;
constexpr int BA_CB = 11;  // classes per thread (21 = 11 + 10; with 21 the four register arrays spill): the per-pixel bookkeeping is paid per batch

// grid (ceil(nbands * W / BA_THREADS), ceil(C / BA_CB), B), nbands = ceil(h / srows); BA_CB * h * w < 2^31
__global__ void __launch_bounds__(BA_THREADS, 4)
ce_backward_ywalk_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ label, const float* __restrict__ lse_in,
                         const float* __restrict__ coef_in, float* __restrict__ U, int C, int h, int w, int H, int W, float sh,
                         float sw, int srows, int nbands) {
    const int t = blockIdx.x * BA_THREADS + threadIdx.x;
    if (t >= nbands * W) return;
    const int band = t / W, x = t - band * W;
    const int c0 = blockIdx.y * BA_CB, b = blockIdx.z, nc = min(BA_CB, C - c0);
    const int ia = band * srows, ib = min(h, ia + srows);  // logit rows this thread owns
    const int hw = h * w;
    const size_t HW = (size_t)H * W;
    const float* __restrict__ pimg = logits + ((size_t)b * C + c0) * hw;
    float* __restrict__ ub = U + ((size_t)b * C + c0) * h * W + x;
    const Lerp lx = make_lerp(x, sw, w);
    ColumnWalk<BA_CB> cw;
    cw.reset();
    float acc0[BA_CB], acc1[BA_CB];  // sums of logit rows r0 / r1
#pragma unroll
    for (int q = 0; q < BA_CB; ++q) acc0[q] = acc1[q] = 0.f;
    int r0 = -1, r1 = -1;
    // row r of the band is complete: store it (r0 == r1 on the last logit row: both accumulators are that row's)
    auto store_row = [&](int r, const float (&a)[BA_CB], const float (&a2)[BA_CB], bool both) {
        if (r < ia || r >= ib) return;
#pragma unroll
        for (int q = 0; q < BA_CB; ++q)
            if (q < nc) ub[((size_t)q * h + r) * W] = both ? a[q] + a2[q] : a[q];
    };
    // first label row whose source row pair reaches row ia (its i1 >= ia); the estimate is low by construction
    int y = (sh > 0.f && ia > 0) ? max(0, (int)((float)(ia - 1) / sh) - 1) : 0;
    while (y < H && make_lerp(y, sh, h).i1 < ia) ++y;
    // the per-pixel inputs (coef, lse, label) are fetched two label rows ahead: read where they are used, every row
    // would wait for two dependent round trips to L2 (89 us instead of the time below)
    const size_t col = (size_t)b * HW + x;
    float pc[2] = {0.f, 0.f}, pl[2] = {0.f, 0.f};
    int plab[2] = {255, 255};
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (y + k < H) {
            const size_t p = col + (size_t)(y + k) * W;
            pc[k] = __ldg(coef_in + p); pl[k] = __ldg(lse_in + p); plab[k] = label[p];
        }
    for (; y < H; ++y) {
        const Lerp ly = make_lerp(y, sh, h);
        if (ly.i0 >= ib) break;
        const float coef = pc[0], lse = pl[0];
        const int lab = plab[0] - c0;
        pc[0] = pc[1]; pl[0] = pl[1]; plab[0] = plab[1];
        if (y + 2 < H) {
            const size_t p = col + (size_t)(y + 2) * W;
            pc[1] = __ldg(coef_in + p); pl[1] = __ldg(lse_in + p); plab[1] = label[p];
        }
        if (ly.i0 != r0 || ly.i1 != r1) {
            if (r0 >= 0) {
                if (ly.i0 == r1 && r1 != r0) {  // the pair moved down by one row: r0 is complete, r1 carries on as the upper row
                    store_row(r0, acc0, acc1, false);
#pragma unroll
                    for (int q = 0; q < BA_CB; ++q) { acc0[q] = acc1[q]; acc1[q] = 0.f; }
                } else {  // any other move (not produced by an enlargement): both rows are complete
                    store_row(r0, acc0, acc1, r1 == r0);
                    if (r1 != r0) store_row(r1, acc1, acc0, false);
#pragma unroll
                    for (int q = 0; q < BA_CB; ++q) acc0[q] = acc1[q] = 0.f;
                }
            }
            r0 = ly.i0;
            r1 = ly.i1;
        }
        cw.advance(pimg, nc, hw, w, ly, lx);
        if (coef != 0.f) {
            const float nl = -lse * CE_LOG2E;
#pragma unroll
            for (int q = 0; q < BA_CB; ++q) {
                const float gq = coef * (ce_ex2(fmaf(cw.value(q, ly), CE_LOG2E, nl)) - (lab == q ? 1.f : 0.f));
                acc0[q] = fmaf(ly.l0, gq, acc0[q]);
                acc1[q] = fmaf(ly.l1, gq, acc1[q]);
            }
        }
    }
    if (r0 >= 0) {
        store_row(r0, acc0, acc1, r1 == r0);
        if (r1 != r0) store_row(r1, acc1, acc0, false);
    }
}

// x pass of the y-first order: grad[b,c,i,j] = g_b * bw_b * sum_x U[b,c,i,x] * wx(x -> j).  grid (ceil(w/32), ceil(h/8), B*C)
// The weights wx(x -> j) of a logit column's footprint depend on j only: with kTab the block computes them once into
// shared memory (row pitch 25: conflict-free) instead of every thread re-deriving make_lerp for each of its ~2/scale
// pixels (40 us -> see profiles/r02_loss_times.txt); spans above XG_SPAN (scale < ~0.1) take the direct form.
constexpr int XG_SPAN = 24;
template <bool kTab>
__global__ void __launch_bounds__(LS_BX * LS_BY)
ce_backward_xgather_kernel(const float* __restrict__ U, const float* __restrict__ bw, const float* __restrict__ gout,
                           float* __restrict__ grad, int C, int h, int w, int W, float sw) {
    __shared__ float wt[LS_BX][XG_SPAN + 1];
    const int j = blockIdx.x * LS_BX + threadIdx.x, i = blockIdx.y * LS_BY + threadIdx.y;
    const int plane = blockIdx.z, b = plane / C;
    int xlo = 0, xhi = -1;
    if (j < w) footprint(j, sw, W, xlo, xhi);
    if (kTab) {
        for (int k = threadIdx.y; k <= xhi - xlo; k += LS_BY) wt[threadIdx.x][k] = weight_to(make_lerp(xlo + k, sw, w), j);
        __syncthreads();
    }
    if (j >= w || i >= h) return;
    const float g = gout[b] * bw[b];
    float acc = 0.f;
    if (g != 0.f) {
        const float* __restrict__ up = U + ((size_t)plane * h + i) * W + xlo;
        for (int k = 0; k <= xhi - xlo; ++k) {
            const float wx = kTab ? wt[threadIdx.x][k] : weight_to(make_lerp(xlo + k, sw, w), j);
            if (wx != 0.f) acc = fmaf(wx, __ldg(up + k), acc);
        }
    }
    grad[(size_t)plane * h * w + (size_t)i * w + j] = g * acc;
}

// the y-first backward applies (the same predicate sizes the workspace)
inline bool backward_y_first(int h, int w, int H) { return H >= 2 * h && (long long)BA_CB * h * w < (1ll << 31); }

}  // namespace

size_t mask_ce_workspace_bytes(int B, int C, int h, int w, int H, int W) {
    // intermediate of the backward pass with resampling: T [B,C,H,w] (x first) or U [B,C,h,W] (y first)
    const size_t t_elems = backward_y_first(h, w, H) ? (size_t)h * W : (size_t)H * w;
    const size_t t_bytes = (h != H || w != W) ? up256(sizeof(float) * (size_t)B * C * t_elems) : 0;
    return up256(sizeof(double) * B) + up256(sizeof(unsigned) * B) + up256(sizeof(float) * B) +
           2 * up256(sizeof(float) * (size_t)B * H * W) + t_bytes;
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
    if (ws == nullptr || ws_bytes < mask_ce_workspace_bytes(B, C, h, w, H, W))
        return set_error(PAMR_ERR_WORKSPACE, "mask_ce: workspace of %zu bytes given, %zu needed", ws_bytes,
                         mask_ce_workspace_bytes(B, C, h, w, H, W));
    if (((uintptr_t)ws & 255) != 0) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce: workspace must be 256-byte aligned");
    if (C >= CE_MAXC_SMEM) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce: C=%d does not fit a uint8 label map", C);
    const CeWorkspace k = carve(ws, B, H, W);
    PAMR_CUDA_TRY(cudaMemsetAsync(k.acc, 0, k.head_bytes, s));  // accumulators and tickets
    dim3 grid((W + LS_BX - 1) / LS_BX, (H + LS_BY - 1) / LS_BY, B), block(LS_BX, LS_BY);
    if (grid.y > 65535 || grid.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce: H/8 and B must be <= 65535");
    const float sh = scale_of(h, H), sw = scale_of(w, W);
    const double inv_hw = 1.0 / ((double)H * (double)W);
    if (H >= 2 * h && C <= CW_CB && (long long)C * h * w < (1ll << 31) && (long long)((H + 7) / 8) * W < (1ll << 30)) {
        // enlarged logits: column walk; rows per band so that the blocks (4 per SM) fill whole waves
        const int rows = walk_rows(B, H, W, CW_THREADS);
        const int nbands = (H + rows - 1) / rows;
        dim3 wgrid((unsigned)((nbands * W + CW_THREADS - 1) / CW_THREADS), B);
        ce_forward_walk_kernel<<<wgrid, CW_THREADS, 0, s>>>(logits, label, class_count, gt_labels, k.acc, k.ticket, k.bw, loss,
                                                            k.lse, k.coef, C, h, w, H, W, sh, sw, inv_hw, rows, nbands);
    } else if (h != H || w != W)
        ce_forward_kernel<true><<<grid, block, 0, s>>>(logits, label, class_count, gt_labels, k.acc, k.ticket, k.bw, loss, k.lse, k.coef, C, h, w, H, W, sh, sw, inv_hw);
    else
        ce_forward_kernel<false><<<grid, block, 0, s>>>(logits, label, class_count, gt_labels, k.acc, k.ticket, k.bw, loss, k.lse, k.coef, C, h, w, H, W, sh, sw, inv_hw);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_mask_ce_backward(const float* logits, const uint8_t* label, const float* grad_loss, float* grad_logits,
                            const void* ws, size_t ws_bytes, int B, int C, int h, int w, int H, int W, cudaStream_t s) {
    if (ws == nullptr || ws_bytes < mask_ce_workspace_bytes(B, C, h, w, H, W))
        return set_error(PAMR_ERR_WORKSPACE, "mask_ce backward: workspace of %zu bytes given, %zu needed", ws_bytes,
                         mask_ce_workspace_bytes(B, C, h, w, H, W));
    const CeWorkspace k = carve(const_cast<void*>(ws), B, H, W);
    dim3 block(LS_BX, LS_BY);
    if (h == H && w == W) {
        dim3 grid((W + LS_BX - 1) / LS_BX, (H + LS_BY - 1) / LS_BY, B);
        if (grid.y > 65535 || grid.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce backward: H/8 and B must be <= 65535");
        ce_backward_same_kernel<<<grid, block, 0, s>>>(logits, label, k.bw, k.lse, k.coef, grad_loss, grad_logits, C, H, W);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        return PAMR_OK;
    }
    const float sh = scale_of(h, H), sw = scale_of(w, W);
    int srows = 4;
    {   // logit rows per band: waves of 4 blocks per SM x (srows + 1)
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) != cudaSuccess || device_sm_count(dev, &sms) != PAMR_OK || sms <= 0) sms = 148;
        long long best_cost = -1;
        for (int r = 4; r <= 16; ++r) {
            const long long blocks = ((long long)((h + r - 1) / r) * W + BA_THREADS - 1) / BA_THREADS * ((C + BA_CB - 1) / BA_CB) * B;
            const long long cost = ((blocks + 4ll * sms - 1) / (4ll * sms)) * (r + 1);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; srows = r; }
        }
    }
    const int nbands = (h + srows - 1) / srows;
    if (backward_y_first(h, w, H) && (long long)nbands * W < (1ll << 30) && B <= 65535 && (size_t)B * C <= 65535 &&
        (h + LS_BY - 1) / LS_BY <= 65535) {
        dim3 ga((unsigned)((nbands * W + BA_THREADS - 1) / BA_THREADS), (C + BA_CB - 1) / BA_CB, B);
        ce_backward_ywalk_kernel<<<ga, BA_THREADS, 0, s>>>(logits, label, k.lse, k.coef, k.T, C, h, w, H, W, sh, sw, srows, nbands);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        dim3 gb((w + LS_BX - 1) / LS_BX, (h + LS_BY - 1) / LS_BY, B * C);
        // widest footprint of a logit column (footprint(): floor((j-1)/sw) - 1 .. ceil((j+1)/sw) + 1, clipped to the row)
        const bool tab = sw > 0.f && 2.f / sw + 5.f <= (float)XG_SPAN;
        if (tab)
            ce_backward_xgather_kernel<true><<<gb, block, 0, s>>>(k.T, k.bw, grad_loss, grad_logits, C, h, w, W, sw);
        else
            ce_backward_xgather_kernel<false><<<gb, block, 0, s>>>(k.T, k.bw, grad_loss, grad_logits, C, h, w, W, sw);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        return PAMR_OK;
    }
    dim3 gx((w + LS_BX - 1) / LS_BX, (H + LS_BY - 1) / LS_BY, B * ((C + CE_BATCH - 1) / CE_BATCH));
    dim3 gy((w + LS_BX - 1) / LS_BX, (h + LS_BY - 1) / LS_BY, B * C);
    if (gx.y > 65535 || gy.y > 65535 || gy.z > 65535)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "mask_ce backward: H/8, h/8 and B*C must be <= 65535");
    ce_backward_xpass_kernel<<<gx, block, 0, s>>>(logits, label, k.lse, k.coef, k.T, C, h, w, H, W, sh, sw);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    ce_backward_ypass_kernel<<<gy, block, 0, s>>>(k.T, k.bw, grad_loss, grad_logits, C, h, w, H, sh);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr
