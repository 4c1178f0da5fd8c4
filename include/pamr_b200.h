/*
 * pamr_b200.h -- C ABI of libpamr_b200.so: PAMR (pixel-adaptive mask refinement) and the
 * pseudo-label epilogue of EnchanterXiao/1-stage-wseg, hand-written for NVIDIA B200 (sm_100a).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the functions below are the
 * boundary a maintainer binds with ctypes (INTEGRATION.md shows the stub).  Each entry cites the
 * reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - all tensors are fp32, NCHW, contiguous, DEVICE pointers unless the name says `host`;
 *   - every call takes the CUDA device ordinal and a cudaStream_t (as void*); work is enqueued
 *     on that stream only, no default-stream work, no device-wide synchronisation;
 *   - the library owns no tensor memory: outputs and scratch are allocated by the caller;
 *   - return value 0 = PAMR_OK; otherwise an error code, with a thread-local message available
 *     from pamr_last_error();
 *   - re-entrant: may be called concurrently from several host threads (nn.DataParallel calls
 *     forward from one Python thread per GPU, reference train.py:112).
 *   - there is no CPU fallback: a call on a machine without a usable sm_100 device fails.
 *
 * Tap order (reference models/mods/pamr.py:18-38, :48-54): for each dilation d in list order the
 * 8 offsets (-d,-d) (-d,0) (-d,+d) (0,-d) (0,+d) (+d,-d) (+d,0) (+d,+d); tap p = 8*i_d + j.
 * Neighbour coordinates are clamped per axis (replicate padding, pamr.py:50).
 */
#ifndef PAMR_B200_H
#define PAMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAMR_B200_ABI_VERSION 2

/* The library is built with hidden visibility; only the functions declared here are exported. */
#if defined(__GNUC__)
#define PAMR_API __attribute__((visibility("default")))
#else
#define PAMR_API
#endif

#define PAMR_OK 0
#define PAMR_ERR_INVALID_ARGUMENT 1
#define PAMR_ERR_CUDA 2
#define PAMR_ERR_UNSUPPORTED_DEVICE 3
#define PAMR_ERR_WORKSPACE 4

#define PAMR_MAX_DILATIONS 16

typedef void* pamr_stream_t; /* cudaStream_t */

/* ABI version of the loaded library (== PAMR_B200_ABI_VERSION it was built with). */
PAMR_API int pamr_b200_abi_version(void);

/* Thread-local description of the last error returned to this thread ("" if none). */
PAMR_API const char* pamr_last_error(void);

/* Device facts used by the host side (SM count for persistent grids, L2 size for bench flushes).
 * Any output pointer may be NULL. */
PAMR_API int pamr_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes);

/* Number of kernels this library has launched since load (all threads); bench.py reports the
 * delta over its timed region as `gpu_launches`. */
PAMR_API unsigned long long pamr_launch_count(void);

/*
 * F.interpolate(x, size, mode="bilinear", align_corners=True) on n_planes = B*Ch planes.
 * Replaces: pamr.py:125 (mask -> image size), models/SoftMaxAE.py:177 (image -> mask size).
 * src [n_planes,h,w] -> dst [n_planes,H,W]; same arithmetic as torch's CPU kernel.
 */
PAMR_API int pamr_resize_bilinear_f32(const float* src, float* dst, int n_planes, int h, int w, int H, int W,
                             int dev, pamr_stream_t stream);

/*
 * Local affinity.  Replaces pamr.py:132-136 (LocalStDev :77-103, LocalAffinityAbs :105-109,
 * mean over channels, softmax over the 8*nd neighbours).
 * img [B,K,H,W] -> aff [B,8*nd,H,W]; each pixel's 8*nd weights sum to 1.
 */
PAMR_API int pamr_affinity_f32(const float* img, float* aff, int B, int K, int H, int W, const int* dilations,
                      int nd, int dev, pamr_stream_t stream);

/*
 * LocalStDev.forward (pamr.py:77-103) on its own: unbiased standard deviation over the 9*nd samples
 * (3x3 neighbourhood incl. the centre at every dilation, replicate padding) of every pixel and channel.
 * img [B,K,H,W] -> sd [B,K,H,W]  (the reference returns the same values as [B,K,1,H,W]).  B*K <= 65535.
 */
PAMR_API int pamr_local_std_f32(const float* img, float* sd, int B, int K, int H, int W, const int* dilations,
                       int nd, int dev, pamr_stream_t stream);

/*
 * `iters` propagation steps.  Replaces the loop pamr.py:138-140 (LocalAffinityCopy :57-75):
 *   M'[b,c,y,x] = sum_p aff[b,p,y,x] * M[b,c,clamp(y+dy_p),clamp(x+dx_p)].
 * aff [B,8*nd,H,W]; m_in [B,C,H,W] (never written); m_out [B,C,H,W] receives the result;
 * scratch: device memory of at least pamr_propagate_scratch_bytes(...) bytes, 256-byte aligned,
 * for the ping-pong buffers (rows pitched to 16 bytes for TMA); may be NULL when that is 0.
 * iters == 0 copies m_in to m_out.  m_in, m_out and scratch must not overlap.
 * cls_max: NULL, or [B,C] unsigned that receives the per-(b,c) maximum of the RESULT in the
 * ordered encoding of pamr_ordered_from_float() (fused into the last step; saves the max pass of
 * pseudo_gtmask, SoftMaxAE.py:35, when no resize follows).  The call initialises it.
 */
PAMR_API size_t pamr_propagate_scratch_bytes(int B, int C, int H, int W, const int* dilations, int nd, int iters);
PAMR_API int pamr_propagate_f32(const float* aff, const float* m_in, float* m_out, void* scratch, size_t scratch_bytes,
                       int B, int C, int H, int W, const int* dilations, int nd, int iters,
                       unsigned* cls_max, int dev, pamr_stream_t stream);

/*
 * PAMR(num_iter, dilations).forward(img, mask)  (pamr.py:124-143) in one call:
 * mask [B,C,h,w] is resized to [H,W] first (:125; skipped when equal), then affinity, then
 * `iters` propagation steps into out [B,C,H,W].  workspace: device scratch of at least
 * pamr_forward_workspace_bytes(...) bytes, 256-byte aligned.
 */
PAMR_API size_t pamr_forward_workspace_bytes(int B, int K, int C, int H, int W, int h, int w, const int* dilations,
                                    int nd, int iters);
PAMR_API int pamr_forward_f32(const float* img, const float* mask, float* out, void* workspace,
                     size_t workspace_bytes, int B, int K, int C, int H, int W, int h, int w,
                     const int* dilations, int nd, int iters, unsigned* cls_max, int dev,
                     pamr_stream_t stream);

/*
 * _rescale_and_clean (SoftMaxAE.py:263-268) fused with the class-max of pseudo_gtmask (:35):
 * v = bilinear(m -> [H,W]) (identity when sizes match); v[:,1:] *= labels[b,c-1].
 * m [B,C,h,w]; labels [B,C-1] float 0/1 or NULL (no gate);
 * cleaned: NULL or [B,C,H,W] receiving v;  cls_max: NULL or [B,C] receiving max over pixels of v
 * (ordered encoding; initialised by the call).
 */
PAMR_API int pamr_clean_f32(const float* m, const float* labels, float* cleaned, unsigned* cls_max, int B, int C,
                   int h, int w, int H, int W, int dev, pamr_stream_t stream);

/*
 * pseudo_gtmask (SoftMaxAE.py:29-50) + argmax / ignore-255 (SoftMaxAE.py:61-67), evaluated on
 * v = gate(bilinear(m -> [H,W])) exactly as pamr_clean_f32 computes it:
 *   thr[b,0] = max(bg_cut*mx[b,0], low_cut), thr[b,c>=1] = max(fg_cut*mx[b,c], low_cut)
 *   set = v > thr; pixels with != 1 class set -> label 255 / all-zero pseudo_gt.
 * cls_max [B,C] (ordered encoding): with cls_max_gated != 0 it is the max of v as written by
 * pamr_clean_f32; with cls_max_gated == 0 it is the un-gated max written by pamr_propagate_f32 /
 * pamr_forward_f32 (only valid when h == H and w == W) and the gate (labels >= 0) is applied to it
 * here, which is exact because x -> fl(g*x) is monotone.
 * label: NULL or uint8 [B,H,W] in {0..C-1, 255};  pseudo_gt: NULL or float [B,C,H,W] one-hot/empty;
 * class_count: NULL or int32 [B,C] receiving the number of pixels assigned to each class
 * (num_pixels_per_class of balanced_mask_loss_ce, SoftMaxAE.py:72; initialised by the call).
 */
PAMR_API int pamr_pseudo_labels_f32(const float* m, const float* labels, const unsigned* cls_max, uint8_t* label,
                           float* pseudo_gt, int* class_count, int B, int C, int h, int w, int H, int W,
                           float bg_cut, float fg_cut, float low_cut, int cls_max_gated, int dev,
                           pamr_stream_t stream);

/*
 * SURVEY 8(f) row 2 -- balanced_mask_loss_ce (models/SoftMaxAE.py:52-88), the direct consumer of the
 * labels, and its gradient w.r.t. the mask logits.
 *
 * pamr_labels_from_onehot_f32: the reference passes pseudo_gt as a float one-hot-or-empty tensor
 * [B,C,H,W]; this derives what the loss needs from it: label = argmax_c (first maximum), 255 where
 * sum_c < 1 (SoftMaxAE.py:61-66), class_count[b,c] = pixels labelled c (:71-72).  Callers that hold the
 * uint8 labels / counts of pamr_pseudo_labels_f32 skip this.
 *
 * pamr_mask_ce_forward_f32: logits [B,C,h,w] are interpolated on the fly (bilinear, align_corners=True,
 * SoftMaxAE.py:58) to the label resolution [H,W];
 *   cw[b,c] = (tot_b - n[b,c]) / (1 + tot_b)   bw[b] = (sum_c gt_labels[b,c] + 1 == #{c: n[b,c] > 0})
 *   loss[b] = bw[b] * (1/(H*W)) * sum_px cw[b,label] * (logsumexp_c z - z[label])      (ignored pixels add 0)
 * The workspace (pamr_mask_ce_workspace_bytes for the same B, C, h, w, H, W; 256-byte aligned) keeps per-pixel
 * log-sum-exp and weights for the backward call, and the backward call's intermediate when (h,w) != (H,W).
 *
 * pamr_mask_ce_backward_f32: grad_logits [B,C,h,w] = d(sum_b grad_loss[b]*loss[b]) / d logits, from the
 * same logits / label map and the workspace the forward call filled.  Deterministic (gather, no atomics).
 */
PAMR_API size_t pamr_mask_ce_workspace_bytes(int B, int C, int h, int w, int H, int W);
PAMR_API int pamr_labels_from_onehot_f32(const float* pseudo_gt, uint8_t* label, int* class_count, int B, int C, int H, int W,
                                int dev, pamr_stream_t stream);
PAMR_API int pamr_mask_ce_forward_f32(const float* logits, const uint8_t* label, const int* class_count, const float* gt_labels,
                             float* loss, void* workspace, size_t workspace_bytes, int B, int C, int h, int w, int H,
                             int W, int dev, pamr_stream_t stream);
PAMR_API int pamr_mask_ce_backward_f32(const float* logits, const uint8_t* label, const float* grad_loss, float* grad_logits,
                              const void* workspace, size_t workspace_bytes, int B, int C, int h, int w, int H, int W,
                              int dev, pamr_stream_t stream);

/*
 * SURVEY 8(f) row 4 -- the data set's denorm (datasets/pascal_voc.py:85-101; train.py:120
 * `image_raw = self.denorm(image.clone())`) folded into the image resize of run_pamr (SoftMaxAE.py:177):
 *   dst[b,k] = bilinear(img_norm[b,k] * std[k] + mean[k]  ->  [H,W]),  align_corners=True
 * (multiply and add rounded separately, as mul_().add_() does; with h == H and w == W it is the plain denorm).
 * mean_host / std_host: HOST arrays of K floats, K <= 8.  img_norm [B,K,h,w] -> dst [B,K,H,W].
 */
PAMR_API int pamr_denorm_resize_f32(const float* img_norm, const float* mean_host, const float* std_host, float* dst, int B,
                           int K, int h, int w, int H, int W, int dev, pamr_stream_t stream);

/*
 * SURVEY 8(f) row 3 -- inference post-processing, one image:
 *   MergeMultiScale._merge_masks (utils/inference_tools.py:134-161): every scale s of masks [S,C,Hp,Wp] is
 *   un-padded (pads[s] = pad_t, pad_l, h_s, w_s; HOST array of 4*S ints), resized to [H,W] (bilinear,
 *   align_corners=False), flipped back along x for odd s when flip != 0, its foreground classes gated by
 *   labels [C-1] (device, or NULL); the scales are averaged and the background raised to bg_pow;
 *   ResultWriter.save, no-CRF path (:85-88): foreground scores < prospect_thresh are zeroed, pred = argmax.
 * merged: NULL or float [C,H,W];  pred: NULL or uint8 [H,W].  S <= 16.
 * (The reference moves the [S,C,Hp,Wp] scores to the host for this; here only the uint8 map has to.)
 */
PAMR_API int pamr_merge_multiscale_f32(const float* masks, const int* pads_host, const float* labels, float* merged,
                              uint8_t* pred, int S, int C, int Hp, int Wp, int H, int W, int flip, float bg_pow,
                              float prospect_thresh, int dev, pamr_stream_t stream);

/*
 * End-to-end convenience with HOST buffers (what a non-PyTorch caller binds): copies image, masks
 * and labels to the device, runs run_pamr (SoftMaxAE.py:176-179: image resized to the mask size,
 * PAMR) -> _rescale_and_clean -> pseudo_gtmask -> argmax, copies the uint8 label map back and
 * synchronises the stream it created.  h_img [B,K,H,W], h_mask [B,C,h,w], h_labels [B,C-1] or NULL,
 * h_label [B,H,W] uint8.  Device scratch is allocated and freed inside the call.
 */
PAMR_API int pamr_pseudo_labels_host_f32(const float* h_img, const float* h_mask, const float* h_labels,
                                uint8_t* h_label, int B, int K, int C, int H, int W, int h, int w,
                                const int* dilations, int nd, int iters, float bg_cut, float fg_cut,
                                float low_cut, int dev);

/* Ordered unsigned encoding of a float (monotone: a < b  <=>  enc(a) < enc(b)); 0 is below every
 * float.  Host helpers for reading cls_max. */
PAMR_API unsigned pamr_ordered_from_float(float v);
PAMR_API float pamr_float_from_ordered(unsigned u);

#ifdef __cplusplus
}
#endif
#endif /* PAMR_B200_H */
