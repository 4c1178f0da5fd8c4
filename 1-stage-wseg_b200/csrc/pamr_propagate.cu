// Propagation: replaces the loop at reference models/mods/pamr.py:138-140
//   for _ in range(num_iter):  m = aff_m(mask);  mask = (m * x).sum(2)
// i.e. M'[b,c,y,x] = sum_p w[b,p,y,x] * M[b,c,clamp(y+dy_p),clamp(x+dx_p)].
// The reference materialises the [B,C,48,H,W] unfolded tensor every iteration; here every
// iteration is one stencil pass (read affinity once, mask once, write mask once).
//
// This file holds the iteration driver, the GENERIC kernel (any dilation list, any C, any H/W:
// one thread per pixel, neighbours fetched through L1 with clamped coordinates) and the STRIP
// (the few remainder columns / rows the tuned kernel's tiles leave out are computed inside the tuned launch).  The tuned sm_100a
// kernel for the standard dilation set lives in pamr_propagate_sm100.cu.
#include <cstdlib>

#include "pamr_common.cuh"

namespace pamr {

// pamr_propagate_sm100.cu
int pair_pitch(int W);
int pair_rows_padded(int H);
int launch_repack_pairs(const float* src, float* dst, int planes, int H, int W, cudaStream_t s);
int launch_propagate_tuned(const float* aff_tiled, const AffTiling& tiling, const float* src, float* dst, int dst_pitch,
                           bool dst_pair, int B, int C, int H, int W, unsigned* cls_max, int dev, bool dependent, cudaStream_t s);
bool tuned_fusable(const AffTiling& tiling, int B, int C, int H, int W, int dev);
size_t tuned_fused_epoch_ints(const AffTiling& tiling, int B);
int launch_propagate_tuned_fused(const float* aff_tiled, const AffTiling& tiling, const float* buf0, const float* buf1, float* out,
                                 int B, int C, int H, int W, int iters, unsigned* cls_max, int* epochs, int dev, cudaStream_t s);

namespace {

constexpr int GEN_BX = 32;
constexpr int GEN_BY = 8;
constexpr int GEN_CC = 7;  // classes accumulated per pass over the taps

// src / dst rows may be pitched; aff is the standard [B,P,H,W] layout.
__global__ void __launch_bounds__(GEN_BX * GEN_BY)
propagate_generic_kernel(const float* __restrict__ aff, const float* __restrict__ m_in, int src_pitch,
                         float* __restrict__ m_out, int dst_pitch, int C, int H, int W, Dilations dil,
                         unsigned* __restrict__ cls_max) {
    const int x = blockIdx.x * GEN_BX + threadIdx.x;
    const int y = blockIdx.y * GEN_BY + threadIdx.y;
    const int b = blockIdx.z;
    const bool valid = (x < W) && (y < H);
    const int xc = min(x, W - 1), yc = min(y, H - 1);
    const size_t HW = (size_t)H * W, HPs = (size_t)H * src_pitch, HPd = (size_t)H * dst_pitch;
    const int P = 8 * dil.nd;
    const float* __restrict__ wp = aff + (size_t)b * P * HW + (size_t)yc * W + xc;

    for (int c0 = 0; c0 < C; c0 += GEN_CC) {
        float acc[GEN_CC];
#pragma unroll
        for (int cc = 0; cc < GEN_CC; ++cc) acc[cc] = 0.f;
        const float* __restrict__ mb = m_in + ((size_t)b * C + c0) * HPs;
        for (int i = 0; i < dil.nd; ++i) {
            const int d = dil.d[i];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int yy = clampi(yc + tap_dy(j) * d, 0, H - 1);
                const int xx = clampi(xc + tap_dx(j) * d, 0, W - 1);
                const float wv = __ldg(wp + (size_t)(8 * i + j) * HW);
                const float* __restrict__ q = mb + (size_t)yy * src_pitch + xx;
#pragma unroll
                for (int cc = 0; cc < GEN_CC; ++cc)
                    if (c0 + cc < C) acc[cc] = fmaf(wv, __ldg(q + (size_t)cc * HPs), acc[cc]);
            }
        }
#pragma unroll
        for (int cc = 0; cc < GEN_CC; ++cc) {
            if (c0 + cc < C) {
                if (valid) m_out[((size_t)b * C + c0 + cc) * HPd + (size_t)y * dst_pitch + x] = acc[cc];
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
    // used only when iters == 0 and a max is requested
    const size_t plane = blockIdx.y;
    const float* __restrict__ p = m + plane * HW;
    unsigned u = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x)
        u = max(u, ordered_from_float(p[i]));
    u = __reduce_max_sync(0xffffffffu, u);
    if ((threadIdx.x & 31) == 0 && u != 0u) atomicMax(cls_max + plane, u);
}

__global__ void zero_u32_kernel(unsigned* p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int launch_generic(const float* aff, const float* src, int src_pitch, float* dst, int dst_pitch, int B, int C, int H,
                   int W, const Dilations& dil, unsigned* cls_max, cudaStream_t s) {
    dim3 block(GEN_BX, GEN_BY);
    dim3 grid((W + GEN_BX - 1) / GEN_BX, (H + GEN_BY - 1) / GEN_BY, B);
    if (grid.y > 65535 || grid.z > 65535)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "propagate: H/8 and B must be <= 65535");
    propagate_generic_kernel<<<grid, block, 0, s>>>(aff, src, src_pitch, dst, dst_pitch, C, H, W, dil, cls_max);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

struct ScratchPlan {
    size_t pingpong_each, aff_tiled, epochs, total;
};
ScratchPlan plan_scratch(int B, int C, int H, int W, const Dilations& dil, int iters, bool aff_is_tiled) {
    ScratchPlan p{0, 0, 0, 0};
    if (iters <= 0) return p;
    const AffTiling t = tuned_tiling(B, C, H, W, dil);
    if (t.R > 0) {
        // tuned kernel: two ping-pong buffers in the row-pair layout [B*C][ceil(H/2)][Wp][2], Wp a multiple of 16
        // (TMA reads 64-bit elements and needs 16-byte global strides; whole 128-byte lines per row pair)
        p.pingpong_each = align_up(sizeof(float) * (size_t)B * C * pair_rows_padded(H) * pair_pitch(W) * 2, 256);
        if (!aff_is_tiled) p.aff_tiled = align_up(sizeof(float) * t.floats, 256);
        p.epochs = align_up(sizeof(int) * tuned_fused_epoch_ints(t, B), 256);  // fused iterations: per-tile / per-image progress
    } else {
        p.pingpong_each = align_up(sizeof(float) * (size_t)B * C * H * W, 256);
    }
    p.total = 2 * p.pingpong_each + p.aff_tiled + p.epochs;
    // small maps: the resident kernel (pamr_resident.cu) carves its own ping-pong buffers + barrier counters out of
    // the same scratch
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        const ResidentPlan rp = resident_plan(B, C, H, W, dil, iters, dev);
        if (rp.ok && rp.scratch_bytes > p.total) p.total = rp.scratch_bytes;
    }
    return p;
}

// Per-thread, per-device side stream and events (created once, reused by every call of that thread): small
// independent kernel (the row-pair repack of the input mask) runs on the side stream concurrently with the
// affinity kernel on the caller's stream, so its time disappears from the critical path.
struct SideResources {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
int side_resources(int dev, SideResources** out) {
    static thread_local SideResources cache[64];
    static thread_local SideResources overflow;
    SideResources* r = (dev >= 0 && dev < 64) ? &cache[dev] : &overflow;
    if (r->side == nullptr || r == &overflow) {
        if (r == &overflow && r->side != nullptr) {  // device ordinals >= 64 share one slot: rebuild per call
            cudaStreamDestroy(r->side);
            cudaEventDestroy(r->ev_fork); cudaEventDestroy(r->ev_join);
            *r = SideResources();
        }
        PAMR_CUDA_TRY(cudaStreamCreateWithFlags(&r->side, cudaStreamNonBlocking));
        PAMR_CUDA_TRY(cudaEventCreateWithFlags(&r->ev_fork, cudaEventDisableTiming));
        PAMR_CUDA_TRY(cudaEventCreateWithFlags(&r->ev_join, cudaEventDisableTiming));
    }
    *out = r;
    return PAMR_OK;
}

// Fork/join of the side stream around one API call.  Once forked, the destructor joins: on every return
// path (errors included) the caller's stream waits for whatever was enqueued on the side stream, so that the
// workspace is never handed back to the allocator while side-stream work still touches it.
struct ForkJoin {
    cudaStream_t main_s = nullptr;
    SideResources* r = nullptr;
    bool forked = false;
    int init(int dev, cudaStream_t m) {
        main_s = m;
        return side_resources(dev, &r);
    }
    int fork() {  // work enqueued on the side stream after this sees everything enqueued on main so far
        PAMR_CUDA_TRY(cudaEventRecord(r->ev_fork, main_s));
        PAMR_CUDA_TRY(cudaStreamWaitEvent(r->side, r->ev_fork, 0));
        forked = true;
        return PAMR_OK;
    }
    int join() {  // work enqueued on main after this sees everything enqueued on the side stream so far
        forked = false;
        PAMR_CUDA_TRY(cudaEventRecord(r->ev_join, r->side));
        PAMR_CUDA_TRY(cudaStreamWaitEvent(main_s, r->ev_join, 0));
        return PAMR_OK;
    }
    ~ForkJoin() {
        if (forked && r != nullptr && cudaEventRecord(r->ev_join, r->side) == cudaSuccess)
            cudaStreamWaitEvent(main_s, r->ev_join, 0);
    }
};

}  // namespace

size_t propagate_scratch_bytes(int B, int C, int H, int W, const Dilations& dil, int iters, bool aff_is_tiled) {
    return plan_scratch(B, C, H, W, dil, iters, aff_is_tiled).total;
}

// Affinity (optional) + `iters` propagation steps.  When img != nullptr the affinity is computed
// here into `aff` (tile-major if the tuned kernel applies, see tuned_tiling), concurrently with the
// repack of the input mask; otherwise `aff` is an input (standard layout unless aff_is_tiled).
int launch_affinity_propagate(const float* img, int K, float* aff_out, float* img_pitched, const float* aff_in, bool aff_is_tiled,
                              const float* m_in, float* m_out, void* scratch, size_t scratch_bytes, int B, int C, int H,
                              int W, const Dilations& dil, int iters, unsigned* cls_max, int dev, cudaStream_t s) {
    const size_t N = (size_t)B * C * H * W;
    // small maps with the affinity computed here: one launch does everything (the affinity stays in registers)
    if (img != nullptr && iters >= 1 && resident_plan(B, C, H, W, dil, iters, dev).ok)
        return launch_resident(img, K, m_in, m_out, scratch, scratch_bytes, B, C, H, W, dil, iters, cls_max, dev, s);
    const AffTiling tiling = tuned_tiling(B, C, H, W, dil);
    const bool tuned = tiling.R > 0;
    if (cls_max != nullptr) {
        // zeroed by a kernel, not cudaMemsetAsync: a memset node in front of the fork/join events below
        // cost ~0.13 ms per iteration on B200 (measured), presumably by serialising the two streams
        zero_u32_kernel<<<(unsigned)(((size_t)B * C + 255) / 256), 256, 0, s>>>(cls_max, (size_t)B * C);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
    }
    if (iters <= 0) {
        if (img != nullptr) {
            int rc = launch_affinity(img, aff_out, B, K, H, W, dil, tiling, img_pitched, s);  // the caller asked for it
            if (rc != PAMR_OK) return rc;
        }
        PAMR_CUDA_TRY(cudaMemcpyAsync(m_out, m_in, N * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (cls_max != nullptr) {
            dim3 grid((unsigned)min((size_t)64, ((size_t)H * W + 255) / 256), B * C);
            class_max_kernel<<<grid, 256, 0, s>>>(m_in, cls_max, (size_t)H * W);
            count_launch();
            PAMR_CUDA_TRY(cudaGetLastError());
        }
        return PAMR_OK;
    }
    if (img != nullptr) aff_is_tiled = tuned;
    const ScratchPlan plan = plan_scratch(B, C, H, W, dil, iters, aff_is_tiled);
    if (scratch == nullptr || scratch_bytes < plan.total)
        return set_error(PAMR_ERR_WORKSPACE, "propagate: scratch of %zu bytes given, %zu needed", scratch_bytes,
                         plan.total);
    if (((uintptr_t)scratch & 255) != 0)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "propagate: scratch must be 256-byte aligned");
    float* P[2] = {(float*)scratch, (float*)((char*)scratch + plan.pingpong_each)};
    if (!tuned && aff_is_tiled)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "propagate: tiled affinity without the tuned kernel");

    int rc = PAMR_OK;
    if (!tuned) {  // generic kernel: standard layout throughout, no side stream
        const float* aff = aff_in;
        if (img != nullptr) {
            if ((rc = launch_affinity(img, aff_out, B, K, H, W, dil, tiling, img_pitched, s)) != PAMR_OK) return rc;
            aff = aff_out;
        }
        const float* src = m_in;
        int next = 0;
        for (int it = 0; it < iters; ++it) {
            const bool last = (it == iters - 1);
            float* dst = last ? m_out : P[next];
            if ((rc = launch_generic(aff, src, W, dst, W, B, C, H, W, dil, last ? cls_max : nullptr, s)) != PAMR_OK) return rc;
            src = dst;
            next ^= 1;
        }
        return PAMR_OK;
    }

    // ---- tuned kernel: the caller's mask is repacked into the row-pair layout on the side stream,
    //      concurrently with the affinity / relayout kernel
    ForkJoin fj;
    if ((rc = fj.init(dev, s)) != PAMR_OK) return rc;
    if ((rc = fj.fork()) != PAMR_OK) return rc;
    const float* aff = aff_in;
    if (img != nullptr) {
        // (the strips of the affinity layout go to the side stream, in front of the mask repack)
        if ((rc = launch_affinity(img, aff_out, B, K, H, W, dil, tiling, img_pitched, s, fj.r->side)) != PAMR_OK) return rc;
        aff = aff_out;
    }
    if ((rc = launch_repack_pairs(m_in, P[0], B * C, H, W, fj.r->side)) != PAMR_OK) return rc;
    if (img == nullptr && !aff_is_tiled) {
        float* at = (float*)((char*)scratch + 2 * plan.pingpong_each);
        if ((rc = launch_aff_relayout(aff_in, at, B, H, W, tiling, s)) != PAMR_OK) return rc;
        aff = at;
    }
    if ((rc = fj.join()) != PAMR_OK) return rc;

#ifdef PAMR_FUSED_ITERATIONS  // experiment build: measured slower than one launch per iteration
    if (iters >= 2 && tuned_fusable(tiling, B, C, H, W, dev)) {
        // all iterations in one launch (tile-level dependencies instead of kernel boundaries)
        int* epochs = (int*)((char*)scratch + 2 * plan.pingpong_each + plan.aff_tiled);
        const size_t n = tuned_fused_epoch_ints(tiling, B);
        zero_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((unsigned*)epochs, n);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        return launch_propagate_tuned_fused(aff, tiling, P[0], P[1], m_out, B, C, H, W, iters, cls_max, epochs, dev, s);
    }
#endif
    const float* src = P[0];
    int next = 1;
    const int Wp = pair_pitch(W);
    for (int it = 0; it < iters; ++it) {
        const bool last = (it == iters - 1);
        float* dst = last ? m_out : P[next];
        rc = launch_propagate_tuned(aff, tiling, src, dst, last ? W : Wp, !last, B, C, H, W, last ? cls_max : nullptr, dev, it > 0, s);
        if (rc != PAMR_OK) return rc;
        src = dst;
        next ^= 1;
    }
    return PAMR_OK;
}

int launch_propagate(const float* aff, bool aff_is_tiled, const float* m_in, float* m_out, void* scratch,
                     size_t scratch_bytes, int B, int C, int H, int W, const Dilations& dil, int iters,
                     unsigned* cls_max, int dev, cudaStream_t s) {
    return launch_affinity_propagate(nullptr, 0, nullptr, nullptr, aff, aff_is_tiled, m_in, m_out, scratch, scratch_bytes, B, C, H,
                                     W, dil, iters, cls_max, dev, s);
}

}  // namespace pamr
