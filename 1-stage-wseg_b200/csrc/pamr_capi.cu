// extern "C" boundary of libpamr_b200.so (declared in include/pamr_b200.h).
// Argument validation, device selection, error strings and the composite entry points live
// here; the kernels are in pamr_affinity.cu / pamr_propagate*.cu / pamr_epilogue.cu.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "pamr_common.cuh"

namespace pamr {

namespace {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Makes `dev` current for the calling thread for the duration of one API call and restores
// the previous device afterwards (DataParallel threads each have their own current device).
struct DeviceGuard {
    int prev = -1;
    int rc = PAMR_OK;
    explicit DeviceGuard(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) {
            rc = set_error(PAMR_ERR_UNSUPPORTED_DEVICE, "no usable CUDA device: %s (there is no CPU fallback)",
                           cudaGetErrorString(e));
            prev = -1;
            return;
        }
        if (dev != prev) {
            e = cudaSetDevice(dev);
            if (e != cudaSuccess) {
                rc = set_error(PAMR_ERR_CUDA, "cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(e));
                prev = -1;
            }
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_device_arch(int dev) {
    // cached per device: the library carries sm_100a code only
    static std::atomic<int> ok[64];
    if (dev < 0) return set_error(PAMR_ERR_INVALID_ARGUMENT, "device ordinal %d is negative", dev);
    if (dev < 64 && ok[dev].load(std::memory_order_relaxed) == 1) return PAMR_OK;
    int major = 0, minor = 0;
    PAMR_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PAMR_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    // the library carries sm_100a SASS only, and arch-specific ("a") code does not run on other 10.x parts
    if (major != 10 || minor != 0)
        return set_error(PAMR_ERR_UNSUPPORTED_DEVICE,
                         "device %d is sm_%d%d; libpamr_b200 is built for sm_100a (B200) only and has no fallback",
                         dev, major, minor);
    if (dev < 64) ok[dev].store(1, std::memory_order_relaxed);
    return PAMR_OK;
}

int make_dilations(const int* dilations, int nd, Dilations* out) {
    PAMR_REQUIRE(dilations != nullptr, "dilations is NULL");
    PAMR_REQUIRE(nd >= 1 && nd <= PAMR_MAX_DILATIONS, "nd=%d out of range [1,%d]", nd, PAMR_MAX_DILATIONS);
    out->nd = nd;
    for (int i = 0; i < PAMR_MAX_DILATIONS; ++i) out->d[i] = 0;
    for (int i = 0; i < nd; ++i) {
        PAMR_REQUIRE(dilations[i] >= 1 && dilations[i] <= 4096, "dilation[%d]=%d out of range [1,4096]", i, dilations[i]);
        out->d[i] = dilations[i];
    }
    return PAMR_OK;
}

int check_dims(int B, int C, int H, int W) {
    PAMR_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1, "non-positive dimension (B=%d C=%d H=%d W=%d)", B, C, H, W);
    PAMR_REQUIRE((size_t)H * (size_t)W < ((size_t)1 << 31), "H*W must be < 2^31");
    return PAMR_OK;
}

#define PAMR_ENTER(dev)                     \
    g_err[0] = 0;                           \
    DeviceGuard _guard(dev);                \
    if (_guard.rc != PAMR_OK) return _guard.rc; \
    {                                       \
        int _rc = check_device_arch(dev);   \
        if (_rc != PAMR_OK) return _rc;     \
    }

#define PAMR_TRY(expr)                 \
    do {                               \
        int _rc = (expr);              \
        if (_rc != PAMR_OK) return _rc; \
    } while (0)

}  // namespace

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

}  // namespace pamr

using namespace pamr;

extern "C" {

int pamr_b200_abi_version(void) { return PAMR_B200_ABI_VERSION; }

const char* pamr_last_error(void) { return g_err; }

unsigned long long pamr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

unsigned pamr_ordered_from_float(float v) { return ordered_from_float(v); }
float pamr_float_from_ordered(unsigned u) { return float_from_ordered(u); }

int pamr_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes) {
    g_err[0] = 0;
    int v = 0;
    if (sm_count) { PAMR_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { PAMR_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { PAMR_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    if (l2_bytes) { PAMR_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev)); *l2_bytes = (size_t)v; }
    return PAMR_OK;
}

int pamr_resize_bilinear_f32(const float* src, float* dst, int n_planes, int h, int w, int H, int W, int dev,
                             pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(src && dst, "resize: NULL pointer");
    PAMR_REQUIRE(n_planes >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "resize: non-positive dimension");
    return launch_resize_bilinear(src, dst, n_planes, h, w, H, W, (cudaStream_t)stream);
}

int pamr_affinity_f32(const float* img, float* aff, int B, int K, int H, int W, const int* dilations, int nd, int dev,
                      pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(img && aff, "affinity: NULL pointer");
    PAMR_TRY(check_dims(B, K, H, W));
    Dilations dil;
    PAMR_TRY(make_dilations(dilations, nd, &dil));
    return launch_affinity(img, aff, B, K, H, W, dil, AffTiling{}, nullptr, (cudaStream_t)stream);
}

int pamr_local_std_f32(const float* img, float* sd, int B, int K, int H, int W, const int* dilations, int nd, int dev,
                       pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(img && sd, "local_std: NULL pointer");
    PAMR_REQUIRE(img != sd, "local_std: img and sd must not alias");
    PAMR_TRY(check_dims(B, K, H, W));
    PAMR_REQUIRE((long long)B * K <= 65535, "local_std: B*K must be <= 65535");
    Dilations dil;
    PAMR_TRY(make_dilations(dilations, nd, &dil));
    return launch_local_std(img, sd, B, K, H, W, dil, (cudaStream_t)stream);
}

size_t pamr_propagate_scratch_bytes(int B, int C, int H, int W, const int* dilations, int nd, int iters) {
    Dilations dil;
    if (make_dilations(dilations, nd, &dil) != PAMR_OK) return 0;
    return propagate_scratch_bytes(B, C, H, W, dil, iters, false);
}

int pamr_propagate_f32(const float* aff, const float* m_in, float* m_out, void* scratch, size_t scratch_bytes, int B,
                       int C, int H, int W, const int* dilations, int nd, int iters, unsigned* cls_max, int dev,
                       pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(aff && m_in && m_out, "propagate: NULL pointer");
    PAMR_REQUIRE(iters >= 0, "propagate: iters=%d is negative", iters);
    PAMR_REQUIRE(m_in != m_out, "propagate: m_in and m_out must not alias");
    PAMR_TRY(check_dims(B, C, H, W));
    Dilations dil;
    PAMR_TRY(make_dilations(dilations, nd, &dil));
    return launch_propagate(aff, false, m_in, m_out, scratch, scratch_bytes, B, C, H, W, dil, iters, cls_max, dev,
                            (cudaStream_t)stream);
}

namespace {
// workspace carve-up shared by pamr_forward_workspace_bytes and pamr_forward_f32
struct ForwardPlan {
    AffTiling tiling;
    size_t aff_bytes, scratch_bytes, resize_bytes, img_bytes, total;
};
ForwardPlan plan_forward(int B, int K, int C, int H, int W, int h, int w, const Dilations& dil, int iters) {
    ForwardPlan p;
    const size_t HW = (size_t)H * W;
    p.tiling = tuned_tiling(B, C, H, W, dil);
    const size_t aff_floats = p.tiling.R > 0 ? p.tiling.floats : (size_t)B * 8 * dil.nd * HW;
    p.aff_bytes = align_up(sizeof(float) * aff_floats, 256);
    p.scratch_bytes = align_up(propagate_scratch_bytes(B, C, H, W, dil, iters, p.tiling.R > 0), 256);
    p.resize_bytes = (h != H || w != W) ? align_up(sizeof(float) * (size_t)B * C * HW, 256) : 0;
    // room for a row-pitched copy of the image (the affinity tile kernel's TMA needs 16-byte aligned rows); reserved
    // whenever the tiled path applies, because whether the caller's base pointer is aligned is not known here
    p.img_bytes = p.tiling.R > 0 ? align_up(affinity_pitched_image_bytes(B, K, H, W), 256) : 0;
    p.total = p.aff_bytes + p.scratch_bytes + p.resize_bytes + p.img_bytes;
    return p;
}
}  // namespace

size_t pamr_forward_workspace_bytes(int B, int K, int C, int H, int W, int h, int w, const int* dilations, int nd,
                                    int iters) {
    Dilations dil;
    if (make_dilations(dilations, nd, &dil) != PAMR_OK) return 0;
    return plan_forward(B, K, C, H, W, h, w, dil, iters).total;
}

int pamr_forward_f32(const float* img, const float* mask, float* out, void* workspace, size_t workspace_bytes, int B,
                     int K, int C, int H, int W, int h, int w, const int* dilations, int nd, int iters,
                     unsigned* cls_max, int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(img && mask && out, "forward: NULL pointer");
    PAMR_REQUIRE(iters >= 0, "forward: iters=%d is negative", iters);
    PAMR_REQUIRE(K >= 1 && h >= 1 && w >= 1, "forward: non-positive dimension");
    PAMR_REQUIRE(out != mask && out != img, "forward: out must not alias mask or img");
    PAMR_TRY(check_dims(B, C, H, W));
    Dilations dil;
    PAMR_TRY(make_dilations(dilations, nd, &dil));
    const ForwardPlan plan = plan_forward(B, K, C, H, W, h, w, dil, iters);
    if (workspace == nullptr || workspace_bytes < plan.total)
        return set_error(PAMR_ERR_WORKSPACE, "forward: workspace of %zu bytes given, %zu needed", workspace_bytes,
                         plan.total);
    PAMR_REQUIRE(((uintptr_t)workspace & 255) == 0, "forward: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    float* aff = (float*)ws;
    void* scratch = ws + plan.aff_bytes;
    const float* m0 = mask;
    if (h != H || w != W) {
        float* rs = (float*)(ws + plan.aff_bytes + plan.scratch_bytes);
        PAMR_TRY(launch_resize_bilinear(mask, rs, B * C, h, w, H, W, s));  // pamr.py:125
        m0 = rs;
    }
    // pamr.py:132-136 (affinity) and :138-140 (propagation loop)
    float* img_pitched = plan.img_bytes ? (float*)(ws + plan.aff_bytes + plan.scratch_bytes + plan.resize_bytes) : nullptr;
    return launch_affinity_propagate(img, K, aff, img_pitched, nullptr, false, m0, out, scratch, plan.scratch_bytes, B, C, H, W,
                                     dil, iters, cls_max, dev, s);
}

int pamr_clean_f32(const float* m, const float* labels, float* cleaned, unsigned* cls_max, int B, int C, int h, int w,
                   int H, int W, int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(m != nullptr, "clean: NULL mask");
    PAMR_REQUIRE(h >= 1 && w >= 1, "clean: non-positive dimension");
    PAMR_TRY(check_dims(B, C, H, W));
    return launch_clean(m, labels, cleaned, cls_max, B, C, h, w, H, W, (cudaStream_t)stream);
}

int pamr_pseudo_labels_f32(const float* m, const float* labels, const unsigned* cls_max, uint8_t* label,
                           float* pseudo_gt, int* class_count, int B, int C, int h, int w, int H, int W, float bg_cut,
                           float fg_cut, float low_cut, int cls_max_gated, int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(m && cls_max, "pseudo_labels: NULL pointer");
    PAMR_REQUIRE(h >= 1 && w >= 1, "pseudo_labels: non-positive dimension");
    PAMR_REQUIRE(C <= 255, "pseudo_labels: C=%d does not fit a uint8 label map with 255 = ignore", C);
    PAMR_REQUIRE(cls_max_gated || (h == H && w == W),
                 "pseudo_labels: an un-gated class max from the propagation step is only valid without a resize");
    PAMR_TRY(check_dims(B, C, H, W));
    return launch_pseudo_labels(m, labels, cls_max, label, pseudo_gt, class_count, B, C, h, w, H, W, bg_cut, fg_cut,
                                low_cut, cls_max_gated != 0, (cudaStream_t)stream);
}

int pamr_denorm_resize_f32(const float* img_norm, const float* mean_host, const float* std_host, float* dst, int B,
                           int K, int h, int w, int H, int W, int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(img_norm && mean_host && std_host && dst, "denorm: NULL pointer");
    PAMR_REQUIRE(h >= 1 && w >= 1, "denorm: non-positive dimension");
    PAMR_TRY(check_dims(B, K, H, W));
    return launch_denorm_resize(img_norm, dst, mean_host, std_host, B, K, h, w, H, W, (cudaStream_t)stream);
}

int pamr_merge_multiscale_f32(const float* masks, const int* pads_host, const float* labels, float* merged,
                              uint8_t* pred, int S, int C, int Hp, int Wp, int H, int W, int flip, float bg_pow,
                              float prospect_thresh, int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(masks && pads_host && (merged || pred), "merge: NULL pointer");
    PAMR_REQUIRE(C >= 1 && C <= 255, "merge: C=%d does not fit a uint8 prediction", C);
    PAMR_REQUIRE(Hp >= 1 && Wp >= 1, "merge: non-positive dimension");
    PAMR_TRY(check_dims(1, C, H, W));
    return launch_merge_multiscale(masks, pads_host, labels, merged, pred, S, C, Hp, Wp, H, W, flip, bg_pow,
                                   prospect_thresh, (cudaStream_t)stream);
}

size_t pamr_mask_ce_workspace_bytes(int B, int C, int h, int w, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    return mask_ce_workspace_bytes(B, C, h, w, H, W);
}

int pamr_labels_from_onehot_f32(const float* pseudo_gt, uint8_t* label, int* class_count, int B, int C, int H, int W,
                                int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(pseudo_gt && label, "labels_from_onehot: NULL pointer");
    PAMR_REQUIRE(C <= 255, "labels_from_onehot: C=%d does not fit a uint8 label map with 255 = ignore", C);
    PAMR_TRY(check_dims(B, C, H, W));
    return launch_labels_from_onehot(pseudo_gt, label, class_count, B, C, H, W, (cudaStream_t)stream);
}

int pamr_mask_ce_forward_f32(const float* logits, const uint8_t* label, const int* class_count, const float* gt_labels,
                             float* loss, void* workspace, size_t workspace_bytes, int B, int C, int h, int w, int H,
                             int W, int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(logits && label && class_count && gt_labels && loss, "mask_ce: NULL pointer");
    PAMR_REQUIRE(h >= 1 && w >= 1, "mask_ce: non-positive dimension");
    PAMR_REQUIRE(C >= 2 && C <= 255, "mask_ce: C=%d out of range", C);
    PAMR_TRY(check_dims(B, C, H, W));
    return launch_mask_ce_forward(logits, label, class_count, gt_labels, loss, workspace, workspace_bytes, B, C, h, w, H,
                                  W, (cudaStream_t)stream);
}

int pamr_mask_ce_backward_f32(const float* logits, const uint8_t* label, const float* grad_loss, float* grad_logits,
                              const void* workspace, size_t workspace_bytes, int B, int C, int h, int w, int H, int W,
                              int dev, pamr_stream_t stream) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(logits && label && grad_loss && grad_logits, "mask_ce backward: NULL pointer");
    PAMR_REQUIRE(h >= 1 && w >= 1, "mask_ce backward: non-positive dimension");
    PAMR_TRY(check_dims(B, C, H, W));
    return launch_mask_ce_backward(logits, label, grad_loss, grad_logits, workspace, workspace_bytes, B, C, h, w, H, W,
                                   (cudaStream_t)stream);
}

int pamr_pseudo_labels_host_f32(const float* h_img, const float* h_mask, const float* h_labels, uint8_t* h_label, int B,
                                int K, int C, int H, int W, int h, int w, const int* dilations, int nd, int iters,
                                float bg_cut, float fg_cut, float low_cut, int dev) {
    PAMR_ENTER(dev);
    PAMR_REQUIRE(h_img && h_mask && h_label, "host: NULL pointer");
    PAMR_REQUIRE(K >= 1 && h >= 1 && w >= 1 && iters >= 0, "host: bad dimension");
    PAMR_REQUIRE(C <= 255, "host: C=%d does not fit a uint8 label map", C);
    PAMR_TRY(check_dims(B, C, H, W));
    Dilations dil;
    PAMR_TRY(make_dilations(dilations, nd, &dil));

    const size_t HW = (size_t)H * W, hw = (size_t)h * w;
    const size_t n_img = sizeof(float) * (size_t)B * K * HW, n_ims = sizeof(float) * (size_t)B * K * hw;
    const size_t n_mask = sizeof(float) * (size_t)B * C * hw, n_lab = sizeof(float) * (size_t)B * (C - 1);
    const AffTiling tiling = tuned_tiling(B, C, h, w, dil);
    const size_t n_aff = sizeof(float) * (tiling.R > 0 ? tiling.floats : (size_t)B * 8 * nd * hw);
    const size_t n_out = (size_t)B * HW;
    const size_t n_max = sizeof(unsigned) * (size_t)B * C;
    const bool resize = (h != H || w != W);
    // one allocation, carved up
    size_t off = 0;
    auto carve = [&](size_t n) { size_t o = off; off += align_up(n ? n : 1, 256); return o; };
    const size_t o_img = carve(n_img), o_ims = carve(resize ? n_ims : 0), o_mask = carve(n_mask);
    const size_t n_scr = propagate_scratch_bytes(B, C, h, w, dil, iters, tiling.R > 0);
    const size_t o_lab = carve(n_lab), o_aff = carve(n_aff), o_a = carve(n_mask), o_b = carve(n_scr);
    const size_t o_max = carve(n_max), o_out = carve(n_out);
    const size_t o_pit = carve(tiling.R > 0 ? affinity_pitched_image_bytes(B, K, h, w) : 0);  // pitched copy of the image for the affinity tile kernel
    char* d = nullptr;
    cudaStream_t s = nullptr;
    PAMR_CUDA_TRY(cudaMalloc(&d, off));
    int rc = PAMR_OK;
    cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (ce != cudaSuccess) {
        cudaFree(d);
        return set_error(PAMR_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(ce));
    }
    auto run = [&]() -> int {
        PAMR_CUDA_TRY(cudaMemcpyAsync(d + o_img, h_img, n_img, cudaMemcpyHostToDevice, s));
        PAMR_CUDA_TRY(cudaMemcpyAsync(d + o_mask, h_mask, n_mask, cudaMemcpyHostToDevice, s));
        if (h_labels && C > 1) PAMR_CUDA_TRY(cudaMemcpyAsync(d + o_lab, h_labels, n_lab, cudaMemcpyHostToDevice, s));
        const float* im = (const float*)(d + o_img);
        if (resize) {  // run_pamr: image -> mask size (SoftMaxAE.py:177)
            PAMR_TRY(launch_resize_bilinear(im, (float*)(d + o_ims), B * K, H, W, h, w, s));
            im = (const float*)(d + o_ims);
        }
        unsigned* mx = (unsigned*)(d + o_max);
        PAMR_TRY(launch_affinity_propagate(im, K, (float*)(d + o_aff), tiling.R > 0 ? (float*)(d + o_pit) : nullptr, nullptr, false, (const float*)(d + o_mask),
                                           (float*)(d + o_a), d + o_b, n_scr, B, C, h, w, dil, iters,
                                           resize ? nullptr : mx, dev, s));
        const float* lab = (h_labels && C > 1) ? (const float*)(d + o_lab) : nullptr;
        if (resize) PAMR_TRY(launch_clean((const float*)(d + o_a), lab, nullptr, mx, B, C, h, w, H, W, s));
        PAMR_TRY(launch_pseudo_labels((const float*)(d + o_a), lab, mx, (uint8_t*)(d + o_out), nullptr, nullptr, B, C, h,
                                      w, H, W, bg_cut, fg_cut, low_cut, resize, s));
        PAMR_CUDA_TRY(cudaMemcpyAsync(h_label, d + o_out, n_out, cudaMemcpyDeviceToHost, s));
        PAMR_CUDA_TRY(cudaStreamSynchronize(s));
        return PAMR_OK;
    };
    rc = run();
    cudaStreamDestroy(s);
    cudaFree(d);
    return rc;
}

}  // extern "C"
