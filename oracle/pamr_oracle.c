/*
 * pamr_oracle.c -- CPU restatement of the reference's PAMR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (the `1-stage-wseg_b200`
 * package or libpamr_b200.so) may import, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and there only as the checker / reported baseline.
 *
 * Parity pinning: this restatement is checked against golden vectors produced by
 * importing the reference's own Python modules in the build container
 * (oracle/gen_golden.py -> tests/golden/ *.npz, tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference repo root).  All tensors are fp32, NCHW, contiguous.
 *
 * Tap order (models/mods/pamr.py:18-38, :48-54): for each dilation d, in list
 * order, the 8 offsets of the 3x3 grid in row-major order skipping the centre:
 *   (-d,-d) (-d,0) (-d,+d) (0,-d) (0,+d) (+d,-d) (+d,0) (+d,+d)
 * tap index p = 8*i_d + j.  Neighbour coordinates are clamped per axis
 * (replicate padding, pamr.py:50).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static const int TAP8_DY[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static const int TAP8_DX[8] = {-1, 0, 1, -1, 1, -1, 0, 1};

int pamr_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void pamr_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/*
 * LocalStDev (pamr.py:77-103): 9 taps per dilation INCLUDING the centre, row-major
 * (:83-93), concatenated over dilations (:48-54) -> 9*nd samples; x.std(2) is the
 * unbiased std (divide by n-1).  torch's CPU kernel accumulates Welford in double
 * and rounds the result to float once.
 * img [B,K,H,W] -> sd [B,K,H,W]
 */
void pamr_oracle_local_std(const float* img, float* sd, int B, int K, int H, int W,
                           const int* dil, int nd) {
    const int n = 9 * nd;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bk = 0; bk < B * K; ++bk) {
        for (int y = 0; y < H; ++y) {
            const float* pl = img + (size_t)bk * H * W;
            for (int x = 0; x < W; ++x) {
                double mean = 0.0, m2 = 0.0;
                int cnt = 0;
                for (int i = 0; i < nd; ++i) {
                    const int d = dil[i];
                    for (int a = -1; a <= 1; ++a)
                        for (int b = -1; b <= 1; ++b) {
                            const int yy = clampi(y + a * d, 0, H - 1);
                            const int xx = clampi(x + b * d, 0, W - 1);
                            const double v = (double)pl[(size_t)yy * W + xx];
                            ++cnt;
                            const double delta = v - mean;
                            mean += delta / (double)cnt;
                            m2 += delta * (v - mean);
                        }
                }
                /* n == 1 cannot happen (n >= 9) */
                sd[(size_t)bk * H * W + (size_t)y * W + x] = (float)sqrt(m2 / (double)(n - 1));
            }
        }
    }
}

/*
 * Affinity half of PAMR.forward (pamr.py:132-136):
 *   a_k[p] = -|I_k(y,x) - I_k(n(y,x,p))| / (1e-8 + 0.1*sd_k)   (:134, LocalAffinityAbs :105-109)
 *   abar[p] = mean_k a_k[p]                                     (:135)
 *   w[p]    = softmax_p(abar)                                   (:136)
 * img [B,K,H,W] -> aff [B,P,H,W], P = 8*nd
 */
void pamr_oracle_affinity(const float* img, float* aff, int B, int K, int H, int W,
                          const int* dil, int nd) {
    const int P = 8 * nd;
    const size_t HW = (size_t)H * W;
    float* sd = (float*)malloc(sizeof(float) * (size_t)B * K * HW);
    pamr_oracle_local_std(img, sd, B, K, H, W, dil, nd);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int y = 0; y < H; ++y) {
            float abar[512];
            for (int x = 0; x < W; ++x) {
                for (int p = 0; p < P; ++p) {
                    const int d = dil[p >> 3];
                    const int yy = clampi(y + TAP8_DY[p & 7] * d, 0, H - 1);
                    const int xx = clampi(x + TAP8_DX[p & 7] * d, 0, W - 1);
                    float s = 0.f;
                    for (int k = 0; k < K; ++k) {
                        const float* pl = img + ((size_t)b * K + k) * HW;
                        const float c = pl[(size_t)y * W + x];
                        const float nb = pl[(size_t)yy * W + xx];
                        const float den = 1e-8f + 0.1f * sd[((size_t)b * K + k) * HW + (size_t)y * W + x];
                        const float a = -fabsf(c - nb) / den;
                        s = (k == 0) ? a : s + a;
                    }
                    abar[p] = s / (float)K;
                }
                float mx = abar[0];
                for (int p = 1; p < P; ++p) mx = abar[p] > mx ? abar[p] : mx;
                float sum = 0.f;
                for (int p = 0; p < P; ++p) {
                    abar[p] = expf(abar[p] - mx);
                    sum += abar[p];
                }
                for (int p = 0; p < P; ++p)
                    aff[((size_t)b * P + p) * HW + (size_t)y * W + x] = abar[p] / sum;
            }
        }
    }
    free(sd);
}

/*
 * Propagation loop (pamr.py:138-140, LocalAffinityCopy :57-75):
 *   repeat iters times: M'[c,y,x] = sum_p w[p,y,x] * M[c, clamp(y+dy_p), clamp(x+dx_p)]
 * aff [B,P,H,W], m_in [B,C,H,W] -> m_out [B,C,H,W].  m_in is not modified.
 */
void pamr_oracle_propagate(const float* aff, const float* m_in, float* m_out, int B, int C,
                           int H, int W, const int* dil, int nd, int iters) {
    const int P = 8 * nd;
    const size_t HW = (size_t)H * W;
    const size_t N = (size_t)B * C * HW;
    if (iters <= 0) { memcpy(m_out, m_in, N * sizeof(float)); return; }
    float* tmp = (float*)malloc(N * sizeof(float));
    const float* src = m_in;
    for (int it = 0; it < iters; ++it) {
        /* ping-pong so that the last iteration lands in m_out */
        float* dst = ((iters - 1 - it) & 1) ? tmp : m_out;
#pragma omp parallel for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b) {
            for (int y = 0; y < H; ++y) {
                int offs[512];
                for (int x = 0; x < W; ++x) {
                    for (int p = 0; p < P; ++p) {
                        const int d = dil[p >> 3];
                        const int yy = clampi(y + TAP8_DY[p & 7] * d, 0, H - 1);
                        const int xx = clampi(x + TAP8_DX[p & 7] * d, 0, W - 1);
                        offs[p] = yy * W + xx;
                    }
                    const float* wv = aff + (size_t)b * P * HW + (size_t)y * W + x;
                    for (int c = 0; c < C; ++c) {
                        const float* pl = src + ((size_t)b * C + c) * HW;
                        /* products are rounded to fp32 as in `m * x` (pamr.py:140); the sum is
                         * carried in double and rounded once, which sits within 1 ulp-ish of
                         * torch's cascaded fp32 sum and of any fp32 summation order */
                        double s = 0.0;
                        for (int p = 0; p < P; ++p) {
                            const float prod = pl[offs[p]] * wv[(size_t)p * HW];
                            s += (double)prod;
                        }
                        dst[((size_t)b * C + c) * HW + (size_t)y * W + x] = (float)s;
                    }
                }
            }
        }
        src = dst;
    }
    free(tmp);
}

/*
 * F.interpolate(mode="bilinear", align_corners=True) as used at pamr.py:125,
 * models/SoftMaxAE.py:177 and :266.  torch (UpSampleKernel / upsample_bilinear2d,
 * float path): scale = (in-1)/(out-1) (0 if out==1), src = scale*dst_index in float,
 * i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1,
 * out = l0h*(l0w*p00 + l1w*p01) + l1h*(l0w*p10 + l1w*p11).
 * src [N,h,w] -> dst [N,H,W]
 */
void pamr_oracle_resize_bilinear(const float* src, float* dst, int N, int h, int w, int H, int W) {
    if (h == H && w == W) { memcpy(dst, src, sizeof(float) * (size_t)N * H * W); return; }
    const float sh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float sw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int y = 0; y < H; ++y) {
            const float fy = sh * (float)y;
            int y0 = (int)fy;
            if (y0 > h - 1) y0 = h - 1;
            const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
            const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
            const float* r0 = src + ((size_t)n * h + y0) * w;
            const float* r1 = src + ((size_t)n * h + y1) * w;
            for (int x = 0; x < W; ++x) {
                const float fx = sw * (float)x;
                int x0 = (int)fx;
                if (x0 > w - 1) x0 = w - 1;
                const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
                const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
                dst[((size_t)n * H + y) * W + x] =
                    ly0 * (lx0 * r0[x0] + lx1 * r0[x1]) + ly1 * (lx0 * r1[x0] + lx1 * r1[x1]);
            }
        }
    }
}

/*
 * Label gate of _rescale_and_clean (SoftMaxAE.py:267): masks[:,1:] *= labels[:,:,None,None]
 * m [B,C,H,W] in place, labels [B,C-1]
 */
void pamr_oracle_gate(float* m, const float* labels, int B, int C, int H, int W) {
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int c = 1; c < C; ++c) {
            const float g = labels[(size_t)b * (C - 1) + (c - 1)];
            float* pl = m + ((size_t)b * C + c) * HW;
            for (size_t i = 0; i < HW; ++i) pl[i] *= g;
        }
}

/*
 * pseudo_gtmask (SoftMaxAE.py:29-50) + argmax/ignore (SoftMaxAE.py:61-67).
 *   mx[b,c] = max over pixels (:35); thr[b,0] = max(fl32(bg_cut*mx), low) (:36,:41-42),
 *   thr[b,c>=1] = max(fl32(fg_cut*mx), low) (:37); pg = m > thr (:44);
 *   pixels with more than one class set are zeroed (:47-48);
 *   label = argmax_c pg, 255 where no class is set (:62-67).
 * m [B,C,H,W] -> pg [B,C,H,W] (may be NULL), label [B,H,W] uint8 (may be NULL)
 */
void pamr_oracle_pseudo_gt(const float* m, float* pg, uint8_t* label, int B, int C, int H, int W,
                           float bg_cut, float fg_cut, float low_cut) {
    const size_t HW = (size_t)H * W;
    float* thr = (float*)malloc(sizeof(float) * (size_t)B * C);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float* pl = m + ((size_t)b * C + c) * HW;
            float mx = pl[0];
            for (size_t i = 1; i < HW; ++i) mx = pl[i] > mx ? pl[i] : mx;
            float t = mx * (c == 0 ? bg_cut : fg_cut);
            thr[b * C + c] = t > low_cut ? t : low_cut;
        }
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        for (size_t i = 0; i < HW; ++i) {
            int cnt = 0, first = -1;
            for (int c = 0; c < C; ++c)
                if (m[((size_t)b * C + c) * HW + i] > thr[b * C + c]) {
                    if (first < 0) first = c;
                    ++cnt;
                }
            if (pg)
                for (int c = 0; c < C; ++c) pg[((size_t)b * C + c) * HW + i] = (cnt == 1 && c == first) ? 1.f : 0.f;
            if (label) label[(size_t)b * HW + i] = (cnt == 1) ? (uint8_t)first : (uint8_t)255;
        }
    }
    free(thr);
}

/*
 * PAMR.forward (pamr.py:124-143): mask is first resized to the image size (:125),
 * then affinity (:132-136) and num_iter propagation steps (:138-140).
 * img [B,K,H,W], mask [B,C,h,w] -> out [B,C,H,W]
 */
void pamr_oracle_forward(const float* img, const float* mask, float* out, int B, int K, int C,
                         int H, int W, int h, int w, const int* dil, int nd, int iters) {
    const size_t HW = (size_t)H * W;
    float* aff = (float*)malloc(sizeof(float) * (size_t)B * 8 * nd * HW);
    float* m0 = (float*)malloc(sizeof(float) * (size_t)B * C * HW);
    pamr_oracle_resize_bilinear(mask, m0, B * C, h, w, H, W);
    pamr_oracle_affinity(img, aff, B, K, H, W, dil, nd);
    pamr_oracle_propagate(aff, m0, out, B, C, H, W, dil, nd, iters);
    free(aff);
    free(m0);
}

/*
 * balanced_mask_loss_ce (models/SoftMaxAE.py:52-88) and its gradient w.r.t. `mask`.
 *   z = bilinear(mask -> size of pseudo_gt), align_corners=True                      (:58)
 *   label = argmax_c pseudo_gt, ignored where sum_c pseudo_gt < 1                    (:61-66)
 *   n[b,c] = sum_px pseudo_gt, tot = sum_c n, cw[b,c] = (tot - n)/(1 + tot) (float)   (:71-74)
 *   ce(px) = logsumexp_c z - z[label]  (0 at ignored pixels)                        (:77)
 *   bw[b] = (sum gt_labels + 1 == #{c : n[b,c] > 0})                                (:82-84)
 *   loss[b] = bw[b] * mean_px(cw[b,label] * ce)   (mean over ALL H*W pixels)         (:86)
 * Gradient of sum_b gout[b]*loss[b]:  dz_c(px) = gout*bw*cw[label]/(H*W) * (softmax_c - [c == label]),
 * pushed through the transpose of the bilinear interpolation.  Sums are accumulated in double.
 * logits [B,C,h,w], pg [B,C,H,W] float one-hot-or-empty, gt_labels [B,C-1], loss [B];
 * gout [B] and grad [B,C,h,w] may both be NULL.
 */
void pamr_oracle_mask_ce(const float* logits, const float* pg, const float* gt_labels, float* loss,
                         const float* gout, float* grad, int B, int C, int h, int w, int H, int W) {
    const size_t HW = (size_t)H * W, hw = (size_t)h * w;
    const float sh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float sw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const int same = (h == H && w == W);
    for (int b = 0; b < B; ++b) {
        float* cw = (float*)malloc(sizeof(float) * C);
        double* gacc = grad ? (double*)calloc((size_t)C * hw, sizeof(double)) : NULL;
        double* z = (double*)malloc(sizeof(double) * C);
        float tot = 0.f;
        int present = 0;
        for (int c = 0; c < C; ++c) {
            double s = 0.0;
            for (size_t i = 0; i < HW; ++i) s += pg[((size_t)b * C + c) * HW + i];
            cw[c] = (float)s;  /* n[b,c], an integer */
            tot += (float)s;
            present += s > 0.0;
        }
        for (int c = 0; c < C; ++c) cw[c] = (tot - cw[c]) / (1.f + tot);
        float gsum = 1.f;  /* + BG */
        for (int c = 0; c < C - 1; ++c) gsum += gt_labels[(size_t)b * (C - 1) + c];
        const float bw = (gsum == (float)present) ? 1.f : 0.f;
        double acc = 0.0;
        for (int y = 0; y < H; ++y) {
            const float fy = sh * (float)y;
            int y0 = (int)fy;
            if (y0 > h - 1) y0 = h - 1;
            const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
            const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
            for (int x = 0; x < W; ++x) {
                const size_t i = (size_t)y * W + x;
                int label = -1;
                float best = 0.f, sum = 0.f;
                for (int c = 0; c < C; ++c) {
                    const float v = pg[((size_t)b * C + c) * HW + i];
                    sum += v;
                    if (label < 0 || v > best) { best = v; label = c; }  /* first maximum, like torch.argmax */
                }
                if (sum < 1.f) continue;  /* ignored: contributes 0 to the sum, still counted in the mean */
                const float fx = sw * (float)x;
                int x0 = (int)fx;
                if (x0 > w - 1) x0 = w - 1;
                const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
                const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
                double m = -1e300;
                for (int c = 0; c < C; ++c) {
                    const float* pl = logits + ((size_t)b * C + c) * hw;
                    const float v = same ? pl[i]
                                         : ly0 * (lx0 * pl[(size_t)y0 * w + x0] + lx1 * pl[(size_t)y0 * w + x1]) +
                                               ly1 * (lx0 * pl[(size_t)y1 * w + x0] + lx1 * pl[(size_t)y1 * w + x1]);
                    z[c] = (double)v;
                    if (z[c] > m) m = z[c];
                }
                double se = 0.0;
                for (int c = 0; c < C; ++c) se += exp(z[c] - m);
                const double lse = m + log(se);
                acc += (double)cw[label] * (lse - z[label]);
                if (gacc) {
                    const double k = (double)gout[b] * bw * cw[label] / (double)HW;
                    for (int c = 0; c < C; ++c) {
                        const double d = k * (exp(z[c] - lse) - (c == label ? 1.0 : 0.0));
                        double* ga = gacc + (size_t)c * hw;
                        if (same) {
                            ga[i] += d;
                        } else {
                            ga[(size_t)y0 * w + x0] += d * ly0 * lx0;
                            ga[(size_t)y0 * w + x1] += d * ly0 * lx1;
                            ga[(size_t)y1 * w + x0] += d * ly1 * lx0;
                            ga[(size_t)y1 * w + x1] += d * ly1 * lx1;
                        }
                    }
                }
            }
        }
        loss[b] = bw * (float)(acc / (double)HW);
        if (gacc) {
            for (size_t e = 0; e < (size_t)C * hw; ++e) grad[(size_t)b * C * hw + e] = (float)gacc[e];
            free(gacc);
        }
        free(cw);
        free(z);
    }
}

/*
 * inference post-processing.
 * MergeMultiScale._merge_masks (utils/inference_tools.py:134-161): for every scale s
 *   cut   masks[s][:, pad_t:pad_t+h_s, pad_l:pad_l+w_s]                      (:136-138, :143)
 *   F.interpolate(cut, (H,W), bilinear, align_corners=False)                 (:146)
 *   flipped back along x for odd s when FLIP                                 (:149-150)
 *   foreground classes multiplied by the image-level labels                  (:154)
 * then the mean over the scales (sequential float sum / S, :158) and mean[0] = mean[0] ** BG_POW (:162).
 * ResultWriter.save, no-CRF path (:85-88): foreground scores < prospect_thresh are zeroed, pred = argmax.
 * torch's align_corners=False source index: src = max(0, (in/out)*(dst + 0.5) - 0.5).
 * masks [S,C,Hp,Wp], pads [S,4] = (pad_t, pad_l, h_s, w_s), labels [C-1] or NULL;
 * merged [C,H,W] (may be NULL), pred [H,W] uint8 (may be NULL).
 */
static inline void lerp_half_pixel(int dst, int in_size, int out_size, int* i0, int* i1, float* l0, float* l1) {
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    int a = (int)src;
    if (a > in_size - 1) a = in_size - 1;
    *i0 = a;
    *i1 = a + (a < in_size - 1 ? 1 : 0);
    *l1 = src - (float)a;
    *l0 = 1.f - *l1;
}

void pamr_oracle_merge_multiscale(const float* masks, const int* pads, const float* labels, float* merged,
                                  uint8_t* pred, int S, int C, int Hp, int Wp, int H, int W, int flip,
                                  float bg_pow, float prospect_thresh) {
    float* mean = (float*)malloc(sizeof(float) * (size_t)C * H * W);
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < C; ++c)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float acc = 0.f;
                for (int s = 0; s < S; ++s) {
                    const int pt = pads[4 * s], pl = pads[4 * s + 1], hs = pads[4 * s + 2], ws = pads[4 * s + 3];
                    const int xs = (flip && (s & 1)) ? W - 1 - x : x;
                    int y0, y1, x0, x1;
                    float ly0, ly1, lx0, lx1;
                    lerp_half_pixel(y, hs, H, &y0, &y1, &ly0, &ly1);
                    lerp_half_pixel(xs, ws, W, &x0, &x1, &lx0, &lx1);
                    const float* pl_ = masks + (((size_t)s * C + c) * Hp + pt) * Wp + pl;
                    float v = ly0 * (lx0 * pl_[(size_t)y0 * Wp + x0] + lx1 * pl_[(size_t)y0 * Wp + x1]) +
                              ly1 * (lx0 * pl_[(size_t)y1 * Wp + x0] + lx1 * pl_[(size_t)y1 * Wp + x1]);
                    if (c > 0 && labels) v *= labels[c - 1];
                    acc += v;
                }
                float m = acc / (float)S;
                if (c == 0) m = powf(m, bg_pow);
                mean[((size_t)c * H + y) * W + x] = m;
            }
    if (merged) memcpy(merged, mean, sizeof(float) * (size_t)C * H * W);
    if (pred)
        for (size_t i = 0; i < (size_t)H * W; ++i) {
            int arg = 0;
            float best = mean[i];
            for (int c = 1; c < C; ++c) {
                float v = mean[(size_t)c * H * W + i];
                if (v < prospect_thresh) v = 0.f;
                if (v > best) { best = v; arg = c; }
            }
            pred[i] = (uint8_t)arg;
        }
    free(mean);
}
