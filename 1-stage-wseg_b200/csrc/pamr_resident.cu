// Small-map RESIDENT kernel: affinity + all `iters` propagation steps of PAMR in ONE launch, for the shapes
// stage_net really calls PAMR with (reference models/SoftMaxAE.py:176-179, 251: masks of 41x41 or 81x81 for a
// 321x321 crop, B = 16; models/mods/pamr.py:132-140 for the arithmetic).
//
// The per-iteration path (pamr_propagate_sm100.cu) needs 1 + iters launches of ~25-30 us each at these sizes, because
// a launch holds only a handful of tiles and every one of them pays the weight fill and the pipeline ramp.  Here
//  * every image is cut into G row blocks, one CTA each (G * images <= number of SMs, cooperative launch, so that
//    all CTAs are co-resident); a thread owns ONE pixel of its CTA's block for the whole kernel;
//  * the CTA first stages the image window (block + 24 rows above / below, 24 replicated columns left / right) in
//    shared memory and every thread computes the 48 affinity weights of its pixel (same arithmetic as
//    affinity_smem_kernel) -- they then stay in 48 REGISTERS for all iterations: the affinity never touches HBM;
//  * an iteration is ceil(C/3) passes over 3 class planes each: the window of the planes is copied global (L2) ->
//    shared memory with cp.async (double buffered: the copy of pass p+1 runs behind the FMAs of pass p, replicate
//    padding in x is applied by the copy's clamped source offsets, in y by 13 per-thread row addresses), then
//    48 x 3 LDS + FFMA per thread, one coalesced store per class into the ping-pong buffer (L2 resident);
//  * iterations are separated by a per-IMAGE barrier (a counter in global memory that only the G CTAs of that
//    image touch), not by a kernel boundary; the class max for pseudo_gtmask is fused into the last iteration.
// The 48 products of a pixel are added in the same (tap-sequence) order as on every other path.
//
// Shared-memory bandwidth bounds this kernel like the tiled one: 48 LDS words per pixel-class (no register reuse
// along y: a thread owns one pixel) + the window staging, i.e. >= 1.5 clk per pixel-class per SM.
#include <atomic>
#include <cstddef>

#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int RS_HALO = 24;
constexpr int RS_CPP = 3;                    // class planes per pass
constexpr int RS_NWC = 23;                   // compute warps
constexpr int RS_NTC = RS_NWC * 32;          // compute threads: 736 = the most pixels a CTA owns
constexpr int RS_NT = RS_NTC + 32;           // + the producer warp = 768 threads (80 registers each; 800 threads would get 72)
constexpr int RS_GRP_STRIDE = 256;           // GRP = 3: group g = threads [256 g, 256 g + 224) (the same TMEM lane quarters in every group)
constexpr int RS_GRP_PIXELS = 224;           // blocks of at most this many pixels run 3 thread groups (one plane of a pass each)
constexpr int RS_SMEM_MAX = 227 * 1024;
constexpr int RS_CTRL_BYTES = 256;           // mbarriers and the TMEM base address behind the stages
constexpr int RS_MAX_SEG = 4;                // class segments per iteration, each with its own per-image barrier
constexpr int RS_MAX_STAGES = 6;
// mbarriers of the stage ring: pass n of the sequence uses stage n % NST but barrier pair n % RS_NBAR.  Waits are by
// phase PARITY, which is only sound if a waiter can never be a whole phase ahead of its barrier, i.e. if pass
// n - RS_NBAR has landed by the time a group waits for pass n.  The group has consumed pass n - GRP, and the producer
// issued that one only after pass n - GRP - NST had been consumed; so RS_NBAR >= GRP + NST is sufficient for any
// combination of groups and stages (with one barrier per stage, 3 groups on 2 stages alias at once, and 3 groups on
// 4 stages alias whenever two bulk copies complete out of order).
constexpr int RS_NBAR = 12;

__host__ __device__ constexpr int rs_dil(int id) { return id == 0 ? 1 : id == 1 ? 2 : id == 2 ? 4 : id == 3 ? 8 : id == 4 ? 12 : 24; }
// tap j of the 3x3 neighbourhood without the centre (pamr.py:25-34)
__host__ __device__ constexpr int rs_ty(int j) { return (j < 3) ? -1 : (j < 5 ? 0 : 1); }
__host__ __device__ constexpr int rs_tx(int j) { return (j == 0 || j == 3 || j == 5) ? -1 : ((j == 1 || j == 6) ? 0 : 1); }
// index of the row-address register for dilation id and row step a in {-1,0,1}
__host__ __device__ constexpr int rs_row(int id, int a) { return a == 0 ? 0 : (a < 0 ? 1 + 2 * id : 2 + 2 * id); }
__host__ __device__ inline size_t rs_plane_stride(int H, int W) { return ((size_t)H * W + 3) / 4 * 4; }

// Shared-memory geometry: NST stages of 3 plane slots of SLOT floats each (compile-time, so that the plane offsets of
// the LDS are immediates).  Three instances: tiny windows (41x41 maps), the 81x81 training shape, and the rest.
template <int SLOT_, int NST_>
struct RsGeom {
    static constexpr int SLOT = SLOT_, NST = NST_;
    static constexpr int STAGE_BYTES = RS_CPP * SLOT * 4;
    static constexpr int CTRL_OFF = NST * STAGE_BYTES;
    static constexpr int SMEM_BYTES = CTRL_OFF + RS_CTRL_BYTES;
    static_assert(SLOT % 32 == 0, "slots are 128-byte aligned (bulk-copy destinations)");
    static_assert(NST >= 2 && NST <= RS_MAX_STAGES && SMEM_BYTES <= RS_SMEM_MAX, "shared memory");
};
using RsGeomS = RsGeom<3072, 6>;   // windows of <= 3064 floats (41 x 41 and smaller): 6 stages, 217 KB
using RsGeomM = RsGeom<4704, 4>;   // <= 4696 floats (81 x 81 in blocks of 9 rows: 57 x 81 = 4617): 4 stages, 221.75 KB
using RsGeomL = RsGeom<9216, 2>;   // <= 9208 floats: 2 stages, 216 KB

struct ResidentParams {
    const float* img;    // [B,K,H,W] at the mask's resolution
    const float* m_in;   // [B,C,H,W]
    float* m_out;        // [B,C,H,W]
    float* pp[2];        // ping-pong buffers [B,C,HWp], HWp = H*W rounded up to 4 floats (16-byte aligned planes)
    unsigned* cls_max;   // [B,C] ordered-uint maxima (zeroed by the caller) or nullptr
    unsigned* counters;  // [B][RS_MAX_SEG + 1] barrier counters: one per class segment and one for the pre-pass (zeroed by the caller)
    int K, C, H, W, iters;
    int rows_per;        // rows per CTA
    int b0;              // first image of this launch
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void rs_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok == 0 && ++spins > (1u << 16)) __trap();  // a protocol bug must not hang the device
    } while (ok == 0);
}
__device__ __forceinline__ void rs_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void rs_tmem_ld8(uint32_t taddr, float (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "r"(taddr));
}
// tcgen05.wait::ld with the batch's registers tied through the asm, so that no consumer is scheduled above the wait
__device__ __forceinline__ void rs_tmem_wait_ld8(float (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));
}
__device__ __forceinline__ void rs_tmem_st8(uint32_t taddr, const float (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
                 : "memory");
}
__device__ __forceinline__ void rs_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void rs_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void rs_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

struct RsCtrl {  // behind the stages
    unsigned long long full[RS_NBAR];         // the pass's planes have landed in its stage (bulk copies, bytes counted)
    unsigned long long empty[RS_NBAR];        // every computing warp has read them
    unsigned long long done[RS_MAX_SEG];      // every computing warp has stored its results of the iteration's class segment
    uint32_t tmem_base;
};
static_assert(sizeof(RsCtrl) <= RS_CTRL_BYTES, "control block");
static_assert(RS_NBAR >= 3 + RS_MAX_STAGES, "barrier ring: groups + stages");

// A thread owns a pixel and takes the 3 class planes of a pass.
// GRP = 1: blocks of up to 736 pixels, every warp works on every pass;
// GRP = 3: three groups of 224 threads own the same (at most 224) pixels and take every third pass of the sequence
//          each, so that small blocks still put 21 warps on the shared-memory pipe.
// Warp 23 is the producer: per-image barrier between iterations and the bulk copies of the class-plane windows.
template <int GRP, class GEO>
__global__ void __launch_bounds__(RS_NT, 1)
pamr_resident_kernel(const ResidentParams prm) {
    constexpr int PPT = RS_CPP;  // planes per thread and pass
    constexpr int NST = GEO::NST;
    extern __shared__ __align__(128) float rs_smem[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(rs_smem);
    RsCtrl* ctrl = reinterpret_cast<RsCtrl*>(reinterpret_cast<unsigned char*>(rs_smem) + GEO::CTRL_OFF);
    const uint32_t full0 = sbase + GEO::CTRL_OFF, empty0 = full0 + 8 * RS_NBAR, done0 = empty0 + 8 * RS_NBAR;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool producer = warp == RS_NWC;
    const int grp = (GRP == 1 || producer) ? 0 : t / RS_GRP_STRIDE, tl = (GRP == 1) ? t : t % RS_GRP_STRIDE;
    const int H = prm.H, W = prm.W, C = prm.C;
    const int b = prm.b0 + (int)blockIdx.y;
    const int y0 = (int)blockIdx.x * prm.rows_per, y1 = min(H, y0 + prm.rows_per);
    const int nown = (y1 - y0) * W;                                    // pixels of this CTA
    const int wy0 = max(0, y0 - RS_HALO), wy1 = min(H, y1 + RS_HALO);  // window rows (inside the image), kept DENSE (pitch W)
    const int nwin = (wy1 - wy0) * W;
    const size_t HW = (size_t)H * W, HWp = rs_plane_stride(H, W);
    // bulk copies move the 16-byte aligned range [s0, e1) of a plane; the window starts m0 floats into it
    const int s0 = (wy0 * W) & ~3, m0 = wy0 * W - s0;
    const uint32_t copy_bytes = (uint32_t)((min((wy1 * W + 3) & ~3, (int)HWp) - s0) * 4);
    const int nwarps_g = (nown + 31) / 32;  // warps of a group that compute

    const int npass = (C + RS_CPP - 1) / RS_CPP;
    // class segments: passes [seg_lo(s), seg_lo(s+1)).  With spare warps (GRP = 3) a separate signaller warp does the
    // arrivals and 4 segments fit; otherwise the producer does both jobs and 2 segments is what it can keep up with
    const int nseg = min(npass, GRP == 3 ? RS_MAX_SEG : 2);
    auto seg_lo = [&](int sg) { return sg * npass / nseg; };
    if (t == 0) {
        for (int i = 0; i < RS_NBAR; ++i) {
            rs_mbar_init(full0 + 8 * i, 1);
            rs_mbar_init(empty0 + 8 * i, nwarps_g);  // a pass is read by the warps of one group
        }
        // GRP = 1: a warp arrives once per segment it has stored.  GRP = 3: once per pass it has stored, so that segment
        // sg is done after (its passes) x (warps of a group) arrivals
        for (int i = 0; i < nseg; ++i) rs_mbar_init(done0 + 8 * i, GRP == 1 ? nwarps_g : nwarps_g * (seg_lo(i + 1) - seg_lo(i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // Tensor Memory: the pixels' 48 weights live there, one TMEM lane per thread (see below)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + GEO::CTRL_OFF + (uint32_t)offsetof(RsCtrl, tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }

    // ---- this thread's pixel; row addresses (replicate padding in y) and column offsets (in x) of its neighbours, pamr.py:50
    const bool active = !producer && tl < nown;              // owns a pixel (stores)
    const bool wactive = !producer && (tl & ~31) < nown;     // its warp owns pixels: the whole warp computes (tcgen05.ld / .st are warp-wide)
    const int tp = active ? tl : nown - 1;
    int ly = tp / W, x = tp - ly * W, y = y0 + ly;
    uint32_t rb[13];
    int co[12];
    auto neighbour_addresses = [&]() {
#pragma unroll
        for (int id = 0; id < 6; ++id) {
#pragma unroll
            for (int a = -1; a <= 1; ++a) {
                const int row = min(max(y + a * rs_dil(id), 0), H - 1) - wy0;
                rb[rs_row(id, a)] = sbase + (uint32_t)(m0 + row * W + x) * 4u;
            }
            co[2 * id] = (max(x - rs_dil(id), 0) - x) * 4;
            co[2 * id + 1] = (min(x + rs_dil(id), W - 1) - x) * 4;
        }
    };
    neighbour_addresses();
    // shared-memory address of tap (dilation id, row step a, column step bx) of plane 0, stage 0
#define RS_ADDR(id, a, bx) (rb[rs_row(id, a)] + (uint32_t)((bx) == 0 ? 0 : co[2 * (id) + ((bx) > 0 ? 1 : 0)]))

    // ---- pre-pass: this CTA's rows of the input mask -> pp[1] in the aligned plane layout (what iteration 0 reads)
    {
        const float* __restrict__ mi = prm.m_in + (size_t)b * C * HW + (size_t)y0 * W;
        float* __restrict__ mo = prm.pp[1] + (size_t)b * C * HWp + (size_t)y0 * W;
        // nown <= RS_NTC < RS_NT: one pixel per thread and plane, 8 planes in flight
        for (int c0 = 0; c0 < C; c0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < C && t < nown) v[u] = __ldg(mi + (size_t)(c0 + u) * HW + t);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < C && t < nown) mo[(size_t)(c0 + u) * HWp + t] = v[u];
        }
    }
    unsigned* ctr = prm.counters + (size_t)(RS_MAX_SEG + 1) * b;  // [s < RS_MAX_SEG]: segment s stored; [RS_MAX_SEG]: pre-pass stored
    const unsigned G = gridDim.x;
    rs_tc_fence_before();
    __syncthreads();  // the mask rows are stored; mbarrier init and the TMEM base address are published
    rs_tc_fence_after();
    if (t == RS_NTC) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + RS_MAX_SEG) : "memory");  // pre-pass stored

    // ================= affinity (pamr.py:132-136) =================
    // TMEM layout.  A warp reaches the 32 lanes of quarter warp % 4; thread <-> lane.  Threads that share a quarter
    // use different 48-column blocks (cb).  GRP = 1: block cb = warp / 4 holds the pixel's weights (first the running
    // sum over the image channels, finally the softmax), in tap-SEQUENCE order.  GRP = 3: the three groups own the same
    // pixels (same lanes); region r < 3 (columns 96 r + 48 cb) takes group r's image channel, region 3 the weights.
    const int cb = (GRP == 1) ? warp / 4 : (warp % (RS_GRP_STRIDE / 32)) / 4;
    const uint32_t tq = ctrl->tmem_base + ((uint32_t)((warp % 4) * 32) << 16) + (uint32_t)(cb * 48);
    const uint32_t t_w = tq + (GRP == 1 ? 0u : 3u * 96u);  // where the passes read the weights
    const bool split = (GRP > 1) && prm.K <= GRP;          // group g computes image channel g alone
    constexpr int IMG_ST = NST - 1;                         // the image planes borrow the last stage
#if defined(PAMR_EXPERIMENTS) && defined(RS_X_STOP)        // timing ablation: 1 = stop after the pre-pass, 2 = after the affinity, 3 = affinity without softmax
    if (RS_X_STOP != 1)
#endif
    {
        for (int k0 = 0; k0 < prm.K; k0 += RS_CPP) {
            const int nk = min(RS_CPP, prm.K - k0);
            if (k0 > 0) __syncthreads();  // everybody is done with the previous planes
            for (int kk = 0; kk < nk; ++kk) {  // dense window rows straight from the caller's tensor (read once)
                const float* __restrict__ ip = prm.img + ((size_t)b * prm.K + k0 + kk) * HW + (size_t)wy0 * W;
                float* sp = rs_smem + (IMG_ST * RS_CPP + kk) * GEO::SLOT + m0;
                for (int e0 = t; e0 < nwin; e0 += RS_NT * 4) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (e0 + u * RS_NT < nwin) v[u] = __ldg(ip + e0 + u * RS_NT);
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (e0 + u * RS_NT < nwin) sp[e0 + u * RS_NT] = v[u];
                }
            }
            __syncthreads();
            for (int kk = 0; kk < nk; ++kk) {
                // who computes channel k0 + kk: its group when the channels are split, else group 0
                if (!wactive || grp != (split ? kk : 0)) continue;
                // (keeps ptxas from hoisting the 48 tap addresses rb + co out of this loop -- into spill slots)
#pragma unroll
                for (int i = 0; i < 12; ++i) asm volatile("" : "+r"(co[i]));
                const uint32_t po = (uint32_t)(IMG_ST * GEO::STAGE_BYTES) + (uint32_t)(kk * GEO::SLOT) * 4u;
#define RS_SMP(i, j) lds_f32(RS_ADDR(i, (j) / 3 - 1, (j) % 3 - 1) + po)
                // unbiased std of the 54 samples in fp32, centre-shifted, 6 partial sums (see pamr_affinity.cu)
                const float cc = lds_f32(rb[0] + po);
                float su[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    float s_ = 0.f;
#pragma unroll
                    for (int j = 0; j < 9; ++j) s_ += RS_SMP(i, j) - cc;
                    su[i] = s_;
                }
                const float mean_u = (((su[0] + su[1]) + (su[2] + su[3])) + (su[4] + su[5])) * (1.0f / 54.0f);
                float sq[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    float s_ = 0.f;
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        const float dv = (RS_SMP(i, j) - cc) - mean_u;
                        s_ = fmaf(dv, dv, s_);
                    }
                    sq[i] = s_;
                }
#undef RS_SMP
                const float m2 = ((sq[0] + sq[1]) + (sq[2] + sq[3])) + (sq[4] + sq[5]);
                const float sd = sqrtf(m2 / 53.0f);
                const float den = __fadd_rn(1e-8f, __fmul_rn(0.1f, sd));
                const float rden = __frcp_rn(den);
                const bool first = split || (k0 + kk == 0);
                const uint32_t t_a = tq + (split ? (uint32_t)(kk * 96) : 0u);
#pragma unroll
                for (int ch = 0; ch < 6; ++ch) {  // 8 taps at a time, in tap-sequence order, summed over the channels in TMEM
                    float a8[8], prev[8];
                    if (!first) rs_tmem_ld8(t_a + ch * 8, prev);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int pt = seq_tap(ch * 8 + i), id = pt >> 3, j8 = pt & 7;
                        a8[i] = lds_f32(RS_ADDR(id, rs_ty(j8), rs_tx(j8)) + po);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) a8[i] = div_markstein(-fabsf(__fsub_rn(cc, a8[i])), den, rden);
                    if (!first) {
                        rs_tmem_wait_ld8(prev);
#pragma unroll
                        for (int i = 0; i < 8; ++i) a8[i] = __fadd_rn(prev[i], a8[i]);
                    }
                    rs_tmem_st8(t_a + ch * 8, a8);
                }
                rs_tmem_wait_st();
            }
        }
        if (split) {  // the other groups' channels
            rs_tc_fence_before();
            __syncthreads();
            rs_tc_fence_after();
        }
#if defined(PAMR_EXPERIMENTS) && defined(RS_X_STOP)
        if (RS_X_STOP != 3)
#endif
        if (wactive && grp == 0) {
            // mean over the channels, softmax over the 48 taps (summed in the reference's tap order p, like every other path)
            asm volatile("" : "+r"(x), "+r"(y));  // rb / co are rebuilt afterwards: their 25 registers are free meanwhile
            float e[48];  // e[s], s = tap-sequence index
#pragma unroll
            for (int ch = 0; ch < 6; ++ch) {
                float v[8];
                rs_tmem_ld8(tq + ch * 8, v);
                rs_tmem_wait_ld8(v);
#pragma unroll
                for (int i = 0; i < 8; ++i) e[ch * 8 + i] = v[i];
            }
            if (split) {
                for (int k = 1; k < prm.K; ++k) {
#pragma unroll
                    for (int ch = 0; ch < 6; ++ch) {
                        float v[8];
                        rs_tmem_ld8(tq + (uint32_t)(k * 96) + ch * 8, v);
                        rs_tmem_wait_ld8(v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[ch * 8 + i] = __fadd_rn(e[ch * 8 + i], v[i]);
                    }
                }
            }
            const float kf = (float)prm.K, rk = __frcp_rn(kf);
            float mx = -INFINITY;
#pragma unroll
            for (int s = 0; s < 48; ++s) {
                e[s] = div_markstein(e[s], kf, rk);
                mx = fmaxf(mx, e[s]);
            }
#pragma unroll
            for (int s = 0; s < 48; ++s) e[s] = expf(e[s] - mx);
            float s_ = 0.f;
#pragma unroll
            for (int p = 0; p < 48; ++p) s_ += e[tap_seq(p)];
            const float rs_ = __frcp_rn(s_);
#pragma unroll
            for (int ch = 0; ch < 6; ++ch) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = div_markstein(e[ch * 8 + i], s_, rs_);
                rs_tmem_st8(t_w + ch * 8, v);
            }
            rs_tmem_wait_st();
        }
        // the image planes are consumed (the producer may overwrite the last stage); groups 1, 2 read what group 0 stored
        rs_tc_fence_before();
        __syncthreads();
        rs_tc_fence_after();
        neighbour_addresses();
    }
    // ================= propagation (pamr.py:138-140) =================
    // Passes form one sequence n = it * npass + p over all iterations: stage n % NST, barrier pair n % RS_NBAR, barrier
    // phase (n / RS_NBAR) & 1.
    // Class planes propagate independently: segment s of iteration it+1 needs only segment s of iteration it from the
    // other CTAs of the image.  So every segment has its own per-image barrier (a counter in global memory), and the
    // round trip store -> release -> acquire -> bulk copy from L2 (~2 us) hides behind the passes of the other segments.
    const bool signaller = (GRP == 3) && warp == 7;  // (an idle warp of group 0: its 32 pixel slots are beyond RS_GRP_PIXELS)
#if defined(PAMR_EXPERIMENTS) && defined(RS_X_STOP)
    if (false) {
    } else if (false)
#endif
    if (signaller) {
        // arrivals: once every computing warp of this CTA has stored segment sg of iteration it, tell the image
        if (lane == 0) {
            for (int it = 0; it + 1 < prm.iters; ++it)
                for (int sg = 0; sg < nseg; ++sg) {
                    rs_mbar_wait(done0 + 8 * sg, (uint32_t)it & 1u);
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + sg) : "memory");
                }
        }
    } else if (producer) {
        if (lane == 0) {
            int n = 0;
            for (int it = 0; it < prm.iters; ++it) {
                const float* src = prm.pp[(it + 1) & 1] + (size_t)b * C * HWp + s0;
                for (int sg = 0; sg < nseg; ++sg) {
                    if (GRP != 3 && it > 0) {  // no signaller warp: this thread does the arrival as well
                        rs_mbar_wait(done0 + 8 * sg, (uint32_t)(it - 1) & 1u);
                        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + sg) : "memory");
                    }
#if !(defined(PAMR_EXPERIMENTS) && defined(RS_X_NOBARRIER))
                    if (it > 0 || sg == 0) {  // every CTA of the image has stored segment sg of the previous iteration (it = 0: the pre-pass)
                        const unsigned* c_ = ctr + (it == 0 ? RS_MAX_SEG : sg);
                        const unsigned target = G * (unsigned)(it == 0 ? 1 : it);
                        unsigned seen, spins = 0;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c_) : "memory");
                            if (seen < target && ++spins > (1u << 22)) __trap();
                        } while (seen < target);
                        asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies (async proxy) come after the other CTAs' stores
                    }
#endif
                    for (int p = seg_lo(sg); p < seg_lo(sg + 1); ++p, ++n) {
                        const int st = n % NST;
                        if (n >= NST) {  // the stage's previous planes (pass n - NST) are consumed
                            const int pn = n - NST;
                            rs_mbar_wait(empty0 + 8 * (pn % RS_NBAR), (uint32_t)(pn / RS_NBAR) & 1u);
                        }
                        const int np = min(RS_CPP, C - p * RS_CPP);
                        const uint32_t bar = full0 + 8 * (n % RS_NBAR);
                        rs_mbar_expect_tx(bar, copy_bytes * (uint32_t)np);
                        for (int j = 0; j < np; ++j)
                            rs_bulk_load(sbase + (uint32_t)(st * GEO::STAGE_BYTES) + (uint32_t)(j * GEO::SLOT) * 4u,
                                         src + (size_t)(p * RS_CPP + j) * HWp, copy_bytes, bar);
                    }
                }
            }
        }
    } else if (wactive) {
        const int total = prm.iters * npass;
        int it = 0, p = grp, sg = 0, st = grp % NST, bi = grp;  // pass n = (it, p), its segment, its stage and its barrier pair
        uint32_t ph = 0u;                                        // (n / RS_NBAR) & 1
        while (p >= npass) { p -= npass; ++it; }
        float* obase = (it >= prm.iters - 1 ? prm.m_out + (size_t)b * C * HW : prm.pp[it & 1] + (size_t)b * C * HWp) + (size_t)y0 * W + tl;
        for (int n = grp; n < total; n += GRP) {  // group g takes passes n = g (mod GRP) of the sequence
            const bool last_it = (it == prm.iters - 1);
            const size_t dps = last_it ? HW : HWp;
            const uint32_t sto = (uint32_t)(st * GEO::STAGE_BYTES);
            rs_mbar_wait(full0 + 8 * bi, ph);
            float acc[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) acc[j] = 0.f;
#if !(defined(PAMR_EXPERIMENTS) && defined(RS_X_NOCOMPUTE))
            float wb[2][8];  // the weights stream out of TMEM in six 8-tap batches, double buffered
            rs_tmem_ld8(t_w, wb[0]);
#pragma unroll
            for (int bt = 0; bt < 6; ++bt) {
                float v[8][PPT];  // the batch's neighbours: all loads in flight before the first FMA
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pt = seq_tap(bt * 8 + i), id = pt >> 3, j8 = pt & 7;
                    const uint32_t ad = RS_ADDR(id, rs_ty(j8), rs_tx(j8)) + sto;
#pragma unroll
                    for (int j = 0; j < PPT; ++j) v[i][j] = lds_f32(ad + (uint32_t)(j * GEO::SLOT) * 4u);
                }
                rs_tmem_wait_ld8(wb[bt & 1]);
                if (bt + 1 < 6) rs_tmem_ld8(t_w + (bt + 1) * 8, wb[(bt + 1) & 1]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int j = 0; j < PPT; ++j) acc[j] = fmaf(wb[bt & 1][i], v[i][j], acc[j]);
                }
            }
#endif
            __syncwarp();
            if (lane == 0) rs_mbar_arrive(empty0 + 8 * bi);  // this warp is done with the stage
            const int c0 = p * RS_CPP;
            float* __restrict__ o = obase + (size_t)c0 * dps;
            if (active) {
#pragma unroll
                for (int j = 0; j < PPT; ++j)
                    if (c0 + j < C) o[(size_t)j * dps] = acc[j];
            }
            if (last_it && prm.cls_max != nullptr) {
                unsigned* cm = prm.cls_max + (size_t)b * C + c0;
#pragma unroll
                for (int j = 0; j < PPT; ++j) {
                    const unsigned m = __reduce_max_sync(0xffffffffu, active ? ordered_from_float(acc[j]) : 0u);
                    if (lane == 0 && c0 + j < C && m != 0u) atomicMax(cm + j, m);
                }
            }
            if (!last_it) {  // this warp's rows of the pass are stored
                while (p >= seg_lo(sg + 1)) ++sg;
                if (GRP > 1 || p + 1 == seg_lo(sg + 1)) {
                    __syncwarp();
                    if (lane == 0) rs_mbar_arrive(done0 + 8 * sg);
                }
            }
            // next pass of this group
            st += GRP % NST;
            if (st >= NST) st -= NST;
            bi += GRP;
            if (bi >= RS_NBAR) { bi -= RS_NBAR; ph ^= 1u; }
            p += GRP;
            if (p >= npass) {
                do { p -= npass; ++it; } while (p >= npass);
                sg = 0;
                obase = (it == prm.iters - 1 ? prm.m_out + (size_t)b * C * HW : prm.pp[it & 1] + (size_t)b * C * HWp) + (size_t)y0 * W + tl;
            }
        }
    }
#undef RS_ADDR
    rs_tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(ctrl->tmem_base), "r"(512));
}

bool rs_standard(const Dilations& dil) {
    if (dil.nd != 6) return false;
    for (int i = 0; i < 6; ++i)
        if (dil.d[i] != rs_dil(i)) return false;
    return true;
}

size_t rs_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

int rs_sm_count(int dev) {
    static std::atomic<int> cache[64];
    int n = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

template <int GRP, class GEO>
int rs_launch_one(const ResidentParams& p, int G, int nb, int dev, cudaStream_t s) {
    static std::atomic<int> attr_set[64];
    if (dev < 0 || dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
        PAMR_CUDA_TRY(cudaFuncSetAttribute(pamr_resident_kernel<GRP, GEO>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEO::SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_set[dev].store(1, std::memory_order_release);
    }
    ResidentParams prm = p;
    void* args[] = {(void*)&prm};
    // cooperative: the launch fails instead of dead-locking if the G * nb CTAs cannot all be resident
    PAMR_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)pamr_resident_kernel<GRP, GEO>, dim3((unsigned)G, (unsigned)nb), dim3(RS_NT),
                                              args, (size_t)GEO::SMEM_BYTES, s));
    count_launch();
    return PAMR_OK;
}

__global__ void rs_zero_kernel(unsigned* a, size_t na, unsigned* b, size_t nb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) a[i] = 0u;
    else if (i - na < nb) b[i - na] = 0u;
}

// rows per CTA: at most RS_NT pixels, and the dense window of min(H, rows + 48) rows (+ alignment slack) must fit a plane slot
bool rs_fits(int rows, int H, int W) {
    const int wr = H < rows + 2 * RS_HALO ? H : rows + 2 * RS_HALO;
    return (long long)rows * W <= RS_NTC && (long long)wr * W + 8 <= RsGeomL::SLOT;
}

}  // namespace

// The resident kernel applies when every image can be cut into row blocks of at most 768 pixels whose window
// (block + 24 rows either side) fits a 36 KB plane slot, with all blocks of at least one image resident at once.
// `images_per_launch` images go into one launch.
ResidentPlan resident_plan(int B, int C, int H, int W, const Dilations& dil, int iters, int dev) {
    ResidentPlan r{};
    if (!rs_standard(dil) || iters < 1 || B < 1 || C < 1) return r;
    if ((long long)H * W >= (1 << 24) || W > RS_NTC) return r;
    const int sms = rs_sm_count(dev);
    if (sms <= 0) return r;
    int rows_max = 0;
    while (rows_max < H && rs_fits(rows_max + 1, H, W)) ++rows_max;
    if (rows_max < 1) return r;
    const int g_min = (H + rows_max - 1) / rows_max;
    if (g_min > sms) return r;
    r.ok = true;
    r.images_per_launch = sms / g_min;
    if (r.images_per_launch > 65535) r.images_per_launch = 65535;
    r.scratch_bytes = 2 * rs_align(sizeof(float) * (size_t)B * C * rs_plane_stride(H, W), 256) + rs_align(sizeof(unsigned) * (size_t)B * (RS_MAX_SEG + 1), 256);
    return r;
}

int launch_resident(const float* img, int K, const float* m_in, float* m_out, void* scratch, size_t scratch_bytes, int B,
                    int C, int H, int W, const Dilations& dil, int iters, unsigned* cls_max, int dev, cudaStream_t s) {
    const ResidentPlan plan = resident_plan(B, C, H, W, dil, iters, dev);
    if (!plan.ok) return set_error(PAMR_ERR_INVALID_ARGUMENT, "resident kernel does not apply to this shape");
    if (scratch == nullptr || scratch_bytes < plan.scratch_bytes)
        return set_error(PAMR_ERR_WORKSPACE, "resident: scratch of %zu bytes given, %zu needed", scratch_bytes, plan.scratch_bytes);
    if (((uintptr_t)scratch & 255) != 0) return set_error(PAMR_ERR_INVALID_ARGUMENT, "resident: scratch must be 256-byte aligned");
    const size_t each = rs_align(sizeof(float) * (size_t)B * C * rs_plane_stride(H, W), 256);
    ResidentParams p;
    p.img = img; p.m_in = m_in; p.m_out = m_out;
    p.pp[0] = (float*)scratch;
    p.pp[1] = (float*)((char*)scratch + each);
    p.counters = (unsigned*)((char*)scratch + 2 * each);
    p.cls_max = cls_max;
    p.K = K; p.C = C; p.H = H; p.W = W; p.iters = iters;
    {
        const size_t nmax = cls_max != nullptr ? (size_t)B * C : 0, nctr = (size_t)B * (RS_MAX_SEG + 1), n = nmax + nctr;
        rs_zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cls_max, nmax, p.counters, nctr);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
    }
    const int sms = rs_sm_count(dev);
    for (int b0 = 0; b0 < B; b0 += plan.images_per_launch) {
        const int nb = (B - b0 < plan.images_per_launch) ? B - b0 : plan.images_per_launch;
        int G = sms / nb;
        if (G > H) G = H;
        const int rows_per = (H + G - 1) / G;
        G = (H + rows_per - 1) / rows_per;
        p.rows_per = rows_per;
        p.b0 = b0;
        const int wr = H < rows_per + 2 * RS_HALO ? H : rows_per + 2 * RS_HALO;
        const int win = wr * W + 8;  // window floats incl. alignment slack
        const bool g3 = rows_per * W <= RS_GRP_PIXELS;
        int rc;
        if (win <= RsGeomS::SLOT) rc = g3 ? rs_launch_one<3, RsGeomS>(p, G, nb, dev, s) : rs_launch_one<1, RsGeomS>(p, G, nb, dev, s);
        else if (win <= RsGeomM::SLOT) rc = g3 ? rs_launch_one<3, RsGeomM>(p, G, nb, dev, s) : rs_launch_one<1, RsGeomM>(p, G, nb, dev, s);
        else rc = g3 ? rs_launch_one<3, RsGeomL>(p, G, nb, dev, s) : rs_launch_one<1, RsGeomL>(p, G, nb, dev, s);
        if (rc != PAMR_OK) return rc;
    }
    return PAMR_OK;
}

}  // namespace pamr
