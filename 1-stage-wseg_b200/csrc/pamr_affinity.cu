// Local-affinity kernels: replace reference models/mods/pamr.py:132-136 (LocalStDev :77-103, LocalAffinityAbs
// :105-109, mean over the image channels, softmax over the 8*nd neighbours).  The per-pixel arithmetic is in
// pamr_affinity_math.cuh; this file is about moving the samples in and the weights out.
//
//  * affinity_tile_sm100_kernel<R, K>: the hot one.  Writes the TILES region of the tile-major layout the tuned
//    propagation kernel consumes (pamr_common.cuh).  Persistent, one CTA per SM, one 32 x 4R tile at a time:
//      - a producer warp brings the K image planes of the tile + 24-pixel halo into shared memory with TMA, one
//        80-float row per cp.async.bulk.tensor.1d over a FLAT map of the image (a flat map has no row stride to
//        align; box starts must still be 16-byte aligned, so an image whose W is not a multiple of 4 is first copied
//        to pitched rows), double buffered.  Replicate padding (pamr.py:50) in y is the row the producer asks
//        for (clamped), in x a per-lane clamped column offset of the consumer: nothing is ever patched;
//      - 2R compute warps: warp w owns Tensor-Memory lane quarter w % 4 (= rows [R q, R q + R) of the tile, lane = x)
//        and two of its R rows; a thread computes a pixel's 48 weights in registers and parks them in ITS TMEM lane
//        at column s*R + i -- exactly where the propagation kernel wants them;
//      - the layout in global memory is the Tensor-Memory image in 16-byte pieces [column / 4][lane][column % 4]
//        (what tcgen05.cp reads on the consumer side), so once a quarter's R/2 warps are done (named barrier), they
//        read the quarter back 16 columns at a time (tcgen05.ld.x16) and every warp-wide store is 512 contiguous
//        bytes.  Tensor Memory is the transposer: written pixel-major (one column per tap), read tap-major.
//        (Round 1 wrote this layout with 4-byte stores 16 bytes apart: 25 % sector efficiency, 0.43 ms at config 2.)
//  * affinity_generic_kernel: any dilation list / any K, samples straight from global memory; also computes the two
//    strip regions of the tiled layout (one column, a few rows) next to the tile kernel.
//  * affinity_smem_kernel: standard [B,48,H,W] layout for the public pamr_affinity_f32 (shared-memory staged).
//  * local_std_kernel: LocalStDev alone (row a4).
#include <cuda.h>

#include <atomic>

#include "pamr_affinity_math.cuh"

namespace pamr {

namespace {

constexpr int AFF_BX = 32;
constexpr int AFF_BY = 4;

__host__ __device__ constexpr int sa_dil(int id) { return id == 0 ? 1 : id == 1 ? 2 : id == 2 ? 4 : id == 3 ? 8 : id == 4 ? 12 : 24; }

bool standard_dilations(const Dilations& dil) {
    if (dil.nd != 6) return false;
    for (int i = 0; i < 6; ++i)
        if (dil.d[i] != sa_dil(i)) return false;
    return true;
}

// ---------------------------------------------------------------- generic kernel (global-memory fetch)
struct GlobalFetch {
    const float* __restrict__ img;  // plane 0 of this image
    size_t HW;
    int x, y, H, W;
    const int* d;
    __device__ __forceinline__ float operator()(int k, int i, int j9) const {
        const int yy = clampi(y + (j9 / 3 - 1) * d[i], 0, H - 1);
        const int xx = clampi(x + (j9 % 3 - 1) * d[i], 0, W - 1);
        return __ldg(img + (size_t)k * HW + (size_t)yy * W + xx);
    }
};

// MODE 0: standard [B,P,H,W] layout, every pixel.
// MODE 1: tiled layout, every pixel of the tile grid (pixels of partial tiles outside the image store zeros).
// MODE 2: tiled layout, strip pixels only (column strip x >= Wt, y < Ht; row strip y >= Ht), one thread each.
// NDT: number of dilations the arrays are sized for (6: the standard set; PAMR_MAX_DILATIONS: anything).
template <int NDT, int MODE>
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
affinity_generic_kernel(const float* __restrict__ img, float* __restrict__ aff, int K, int H, int W, Dilations dil,
                        AffTiling tiling) {
    const int nd = dil.nd;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    int x, y;
    if (MODE == 2) {
        const int wc = W - tiling.Wt, ncs = wc * tiling.Ht, nrs = (H - tiling.Ht) * W;
        const int idx = (blockIdx.x * AFF_BY + threadIdx.y) * AFF_BX + threadIdx.x;
        if (idx >= ncs + nrs) return;
        if (idx < ncs) { x = tiling.Wt + idx % wc; y = idx / wc; }
        else { x = (idx - ncs) % W; y = tiling.Ht + (idx - ncs) / W; }
    } else {
        x = blockIdx.x * AFF_BX + threadIdx.x;
        y = blockIdx.y * AFF_BY + threadIdx.y;
        if (x >= W || y >= H) {
            if (MODE == 1 && x < tiling.tiles_x * 32 && y < tiling.tiles_y * 4 * tiling.R) {
                AffTiling full = tiling;  // a partial tile's pixel outside the image: address it as a tile pixel
                full.Wt = tiling.tiles_x * 32; full.Ht = tiling.tiles_y * 4 * tiling.R;
                const AffPixel px = aff_pixel(full, b, y, x);
#pragma unroll
                for (int s = 0; s < 48; ++s) aff[px.at(s)] = 0.f;
            }
            return;
        }
    }
    float w[8 * NDT];
    const GlobalFetch fetch{img + (size_t)b * K * HW, HW, x, y, H, W, dil.d};
    if (K <= AFF_MAXK) {
        affinity_pixel<NDT>(fetch, K, nd, w);
    } else {  // many channels: one loop, u and w live together
        const float l2k = 1.4426950408889634f / (float)K;
#pragma unroll
        for (int p = 0; p < 8 * NDT; ++p) w[p] = 0.f;
        for (int k = 0; k < K; ++k) {
            float u[8 * NDT];
            const float c = fetch(k, 0, 4);
#pragma unroll
            for (int i = 0; i < NDT; ++i)
                if (i < nd) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[8 * i + j] = __fsub_rn(fetch(k, i, j < 4 ? j : j + 1), c);
                }
            const float v = neg_scaled_rcp_den(std_from_shifted<NDT>(u, nd), l2k);
#pragma unroll
            for (int p = 0; p < 8 * NDT; ++p)
                if (p < 8 * nd) w[p] = fmaf(fabsf(u[p]), v, w[p]);
        }
        softmax_base2<NDT>(w, nd);
    }
    if (MODE == 0) {
        float* __restrict__ out = aff + (size_t)b * 8 * nd * HW + (size_t)y * W + x;
#pragma unroll
        for (int p = 0; p < 8 * NDT; ++p)
            if (p < 8 * nd) out[(size_t)p * HW] = w[p];
    } else {
        const AffPixel px = aff_pixel(tiling, b, y, x);
#pragma unroll
        for (int p = 0; p < 8 * NDT; ++p)
            if (p < 48) aff[px.at(tap_seq(p < 48 ? p : 0))] = w[p];
    }
}

// LocalStDev.forward (pamr.py:98-103): one thread per pixel and channel
template <int NDT>
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
local_std_kernel(const float* __restrict__ img, float* __restrict__ sd, int K, int H, int W, Dilations dil) {
    const int x = blockIdx.x * AFF_BX + threadIdx.x, y = blockIdx.y * AFF_BY + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W;
    const GlobalFetch fetch{img + (size_t)blockIdx.z * HW, HW, x, y, H, W, dil.d};  // blockIdx.z = b*K + k
    float u[8 * NDT];
    const float c = fetch(0, 0, 4);
#pragma unroll
    for (int i = 0; i < NDT; ++i)
        if (i < dil.nd) {
#pragma unroll
            for (int j = 0; j < 8; ++j) u[8 * i + j] = __fsub_rn(fetch(0, i, j < 4 ? j : j + 1), c);
        }
    sd[(size_t)blockIdx.z * HW + (size_t)y * W + x] = std_from_shifted<NDT>(u, dil.nd);
}

// ---------------------------------------------------------------- standard layout, shared-memory staged
// A CTA owns a 32 x 8 pixel tile; the K image planes of the tile plus its 24-pixel halo are staged in shared memory
// once, with replicate padding applied while staging (clamped source coordinates), so every neighbour is a
// shared-memory load at an immediate offset.
constexpr int SA_BX = 32, SA_BY = 8, SA_HALO = 24;
constexpr int SA_W = SA_BX + 2 * SA_HALO;  // 80
constexpr int SA_H = SA_BY + 2 * SA_HALO;  // 56

struct SmemFetch {
    const float* c0;  // centre sample of plane 0
    __device__ __forceinline__ float operator()(int k, int i, int j9) const {
        return c0[k * (SA_H * SA_W) + (j9 / 3 - 1) * sa_dil(i) * SA_W + (j9 % 3 - 1) * sa_dil(i)];
    }
};

__global__ void __launch_bounds__(SA_BX * SA_BY, 2)
affinity_smem_kernel(const float* __restrict__ img, float* __restrict__ aff, int K, int H, int W) {
    extern __shared__ float sa_tile[];  // [K][SA_H][SA_W]
    const int x0 = blockIdx.x * SA_BX, y0 = blockIdx.y * SA_BY, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    // staging: thread (tx,ty) covers window columns tx, tx+32, tx+64 and rows ty, ty+8, ... of every plane; all 21
    // loads of a plane are in flight before the first shared-memory store
    for (int k = 0; k < K; ++k) {
        const float* __restrict__ pl = img + ((size_t)b * K + k) * HW;
        float v[3][SA_H / SA_BY];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int wx = threadIdx.x + 32 * cc;
            const int gx = clampi(x0 - SA_HALO + wx, 0, W - 1);
#pragma unroll
            for (int rr = 0; rr < SA_H / SA_BY; ++rr) {
                const int gy = clampi(y0 - SA_HALO + (int)threadIdx.y + SA_BY * rr, 0, H - 1);
                v[cc][rr] = (wx < SA_W) ? __ldg(pl + (size_t)gy * W + gx) : 0.f;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int wx = threadIdx.x + 32 * cc;
#pragma unroll
            for (int rr = 0; rr < SA_H / SA_BY; ++rr)
                if (wx < SA_W) sa_tile[(k * SA_H + threadIdx.y + SA_BY * rr) * SA_W + wx] = v[cc][rr];
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    float w[48];
    const SmemFetch fetch{sa_tile + (threadIdx.y + SA_HALO) * SA_W + threadIdx.x + SA_HALO};
    affinity_pixel<6>(fetch, K, 6, w);
    float* __restrict__ out = aff + (size_t)b * 48 * HW + (size_t)y * W + x;
#pragma unroll
    for (int p = 0; p < 48; ++p) out[(size_t)p * HW] = w[p];
}

// ---------------------------------------------------------------- tile kernel (TMA in, Tensor Memory transposer out)
constexpr int AT_HALO = 24;
constexpr int AT_BOX = 32 + 2 * AT_HALO;  // 80 floats per TMA row
constexpr int AT_PITCH = 96;              // row pitch in shared memory: 384 bytes (TMA destinations are 128-byte aligned)
constexpr int AT_CTRL_BYTES = 128;

template <int R, int K>
struct AtCfg {
    static_assert(R % 2 == 0, "a warp takes rows j and j + R/2 of its quarter");
    static constexpr int NWC = 2 * R;             // compute warps: 4 quarters x R/2
    static constexpr int NT = (NWC + 1) * 32;     // + the producer warp
    static constexpr int WIN_H = 4 * R + 2 * AT_HALO;
    static constexpr int PLANE_BYTES = WIN_H * AT_PITCH * 4;
    static constexpr int BUF_BYTES = K * PLANE_BYTES;
    static constexpr int CTRL_OFF = 2 * BUF_BYTES;
    static constexpr int SMEM_BYTES = CTRL_OFF + AT_CTRL_BYTES;
    static constexpr int NCOLS = 48 * R;          // Tensor Memory columns of a tile
    static constexpr int DUMP_COLS = NCOLS / (R / 2);  // columns a warp writes out: 96
    static_assert(DUMP_COLS % 16 == 0 && NCOLS <= 512, "dump in 16-column batches");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct AtCtrl {
    unsigned long long full[2], empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(AtCtrl) <= AT_CTRL_BYTES, "control block");

struct AtParams {
    float* aff;          // tiles region of the tiled layout
    int B, H, W;
    int pitch;           // row pitch (floats) of the image the TMA map describes: W, or W rounded up to 4
    int tiles_x, tiles_y, ntiles;
};

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float at_lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void at_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void at_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok == 0 && ++spins > (1u << 20)) __trap();  // a protocol bug must not hang the device
    } while (ok == 0);
}
__device__ __forceinline__ void at_tma_row(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0)
                 : "memory");
}
__device__ __forceinline__ void at_tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(v) : "memory");
}
__device__ __forceinline__ void at_tmem_ld16(uint32_t taddr, float (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void at_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// named barrier of the quarter's R/2 warps (ids 1..4; immediate ids so that ptxas does not reserve all 16)
template <int NTHREADS>
__device__ __forceinline__ void at_quarter_sync(int q) {
    if (q == 0) asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
    else if (q == 1) asm volatile("bar.sync 2, %0;" ::"n"(NTHREADS) : "memory");
    else if (q == 2) asm volatile("bar.sync 3, %0;" ::"n"(NTHREADS) : "memory");
    else asm volatile("bar.sync 4, %0;" ::"n"(NTHREADS) : "memory");
}

// shared-memory fetch of the tile kernel: a[0] = address of the pixel itself in plane 0, a[1 + 2 id] / a[2 + 2 id] =
// the same row at the clamped columns x - d / x + d
template <int PLANE_BYTES>
struct TileFetch {
    uint32_t a[13];
    __device__ __forceinline__ float operator()(int k, int i, int j9) const {
        const int tx = j9 % 3 - 1, ty = j9 / 3 - 1;
        const uint32_t base = (tx == 0) ? a[0] : a[(tx < 0 ? 1 : 2) + 2 * i];
        return at_lds(base + (uint32_t)(k * PLANE_BYTES) + (uint32_t)(ty * sa_dil(i) * AT_PITCH * 4));
    }
};

template <int R, int K>
__global__ void __launch_bounds__(AtCfg<R, K>::NT, 1)
affinity_tile_sm100_kernel(const __grid_constant__ CUtensorMap tmap, const AtParams prm) {
    using C_ = AtCfg<R, K>;
    extern __shared__ __align__(1024) unsigned char at_smem[];
    AtCtrl* ctrl = reinterpret_cast<AtCtrl*>(at_smem + C_::CTRL_OFF);
    const uint32_t sbase = at_smem_u32(at_smem);
    const uint32_t full0 = at_smem_u32(&ctrl->full[0]), empty0 = at_smem_u32(&ctrl->empty[0]);
    // (through a shuffle: the compiler then knows the warp index is warp-uniform and keeps what derives from it --
    // Tensor-Memory addresses above all -- in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int H = prm.H, W = prm.W;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            at_mbar_init(full0 + 8 * i, 1);
            at_mbar_init(empty0 + 8 * i, C_::NWC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(&ctrl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    at_fence_before();
    __syncthreads();
    at_fence_after();

    const int my_tiles = (prm.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tiles_per_img = prm.tiles_x * prm.tiles_y;

    if (warp == C_::NWC) {
        // ===================== producer: the K image planes of the tile's window, one TMA per row =====================
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int b = tile / tiles_per_img, tt = tile % tiles_per_img;
            const int x0 = (tt % prm.tiles_x) * 32, y0 = (tt / prm.tiles_x) * (4 * R);
            const int buf = t & 1;
            const uint32_t bar = full0 + 8 * buf;
            if (lane == 0) {
                if (t >= 2) at_mbar_wait(empty0 + 8 * buf, (uint32_t)(t / 2 - 1) & 1u);  // the buffer's previous tile is consumed
                at_mbar_expect_tx(bar, (uint32_t)(K * C_::WIN_H * AT_BOX * 4));
            }
            __syncwarp();
            for (int r = lane; r < K * C_::WIN_H; r += 32) {
                const int k = r / C_::WIN_H, wr = r % C_::WIN_H;
                const int gy = clampi(y0 - AT_HALO + wr, 0, H - 1);  // replicate padding in y
                at_tma_row(sbase + (uint32_t)(buf * C_::BUF_BYTES + r * AT_PITCH * 4), &tmap, bar,
                           ((b * K + k) * H + gy) * prm.pitch + x0 - AT_HALO);
            }
        }
    } else {
        // ===================== compute warps =====================
        const int wq = warp & 3, jrow = warp >> 2;  // lane quarter; rows jrow and jrow + R/2 of the quarter
        const uint32_t tq = ctrl->tmem_base + ((uint32_t)(wq * 32) << 16);
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int tt = tile % tiles_per_img;
            const int x0 = (tt % prm.tiles_x) * 32, y0 = (tt / prm.tiles_x) * (4 * R);
            const int buf = t & 1;
            const int x = x0 + lane;
            at_mbar_wait(full0 + 8 * buf, (uint32_t)(t / 2) & 1u);
            TileFetch<C_::PLANE_BYTES> fetch;
            {
                const uint32_t row0 = sbase + (uint32_t)(buf * C_::BUF_BYTES) +
                                      (uint32_t)(((wq * R + jrow + AT_HALO) * AT_PITCH + AT_HALO + lane) * 4);
                fetch.a[0] = row0;
#pragma unroll
                for (int id = 0; id < 6; ++id) {  // replicate padding in x: clamped column offsets
                    fetch.a[1 + 2 * id] = row0 + (uint32_t)((max(x - sa_dil(id), 0) - x) * 4);
                    fetch.a[2 + 2 * id] = row0 + (uint32_t)((min(x + sa_dil(id), W - 1) - x) * 4);
                }
            }
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int i = jrow + h * (R / 2);
                float w[48];
                const bool inside = x < W && (y0 + wq * R + i) < H;  // pixels of a partial tile outside the image: zeros
                affinity_pixel<6>(fetch, K, 6, w, inside ? 1.0f : 0.0f);
                const uint32_t tcol = tq + (uint32_t)i;
#pragma unroll
                for (int p = 0; p < 48; ++p) at_tmem_st1(tcol + (uint32_t)(tap_seq(p) * R), w[p]);
#pragma unroll
                for (int q = 0; q < 13; ++q) fetch.a[q] += (uint32_t)((R / 2) * AT_PITCH * 4);
            }
            __syncwarp();
            if (lane == 0) at_mbar_arrive(empty0 + 8 * buf);  // this warp has read the window for the last time
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            at_fence_before();
            at_quarter_sync<(R / 2) * 32>(wq);  // the quarter's 48 R columns are complete
            at_fence_after();
            // write the quarter out: this warp takes DUMP_COLS columns, 16 at a time; a 16-byte piece of lane l at
            // piece index c/4 sits at tile + (c/4) * 512 + l * 4 floats (pamr_common.cuh)
            float4* __restrict__ out = reinterpret_cast<float4*>(prm.aff + (size_t)tile * aff_tile_floats(R)) + wq * 32 + lane;
#pragma unroll 1
            for (int bt = 0; bt < C_::DUMP_COLS / 16; ++bt) {
                const int c0 = jrow * C_::DUMP_COLS + bt * 16;
                float v[16];
                at_tmem_ld16(tq + (uint32_t)c0, v);
#pragma unroll
                for (int m = 0; m < 4; ++m) out[(size_t)(c0 / 4 + m) * 128] = make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
            }
            at_fence_before();
            at_quarter_sync<(R / 2) * 32>(wq);  // everybody has read the quarter back: the next tile may overwrite it
            at_fence_after();
        }
    }
    at_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(ctrl->tmem_base), "r"(512));
}

// [planes][H][W] -> [planes][H][pitch] (pitch = W rounded up to 4 floats; the padding is never read).  One thread per
// 16 bytes of the destination: four scalar loads (the source rows have no alignment), one 128-bit store.
__global__ void __launch_bounds__(256) pitch_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int W, int pitch, size_t rows) {
    const int q4 = pitch >> 2;  // float4 per destination row
    const size_t n = rows * (size_t)q4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / q4;
        const int x = (int)(i - r * q4) * 4;
        const float* __restrict__ p = src + r * W + x;
        float4 v;
        v.x = __ldg(p);
        v.y = (x + 1 < W) ? __ldg(p + 1) : 0.f;
        v.z = (x + 2 < W) ? __ldg(p + 2) : 0.f;
        v.w = (x + 3 < W) ? __ldg(p + 3) : 0.f;
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}

template <int R, int K>
int launch_tile_kernel(const float* img, int pitch, float* aff, int B, int H, int W, const AffTiling& tiling, int dev, cudaStream_t s) {
    using C_ = AtCfg<R, K>;
    static std::atomic<int> attr_set[64];
    if (dev < 0 || dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
        PAMR_CUDA_TRY(cudaFuncSetAttribute(affinity_tile_sm100_kernel<R, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_set[dev].store(1, std::memory_order_release);
    }
    alignas(64) CUtensorMap tmap;
    int rc = encode_tensor_map_1d_f32(&tmap, img, (unsigned long long)B * K * H * pitch, AT_BOX);
    if (rc != PAMR_OK) return rc;
    int sms = 0;
    if ((rc = device_sm_count(dev, &sms)) != PAMR_OK) return rc;
    AtParams p;
    p.aff = aff; p.B = B; p.H = H; p.W = W; p.pitch = pitch;
    p.tiles_x = tiling.tiles_x; p.tiles_y = tiling.tiles_y;
    p.ntiles = B * tiling.tiles_x * tiling.tiles_y;
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    affinity_tile_sm100_kernel<R, K><<<grid, C_::NT, C_::SMEM_BYTES, s>>>(tmap, p);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

// standard [B,48,H,W] -> tile-major (for callers of the public propagate API)
__global__ void __launch_bounds__(AFF_BX * AFF_BY)
aff_relayout_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, AffTiling tiling) {
    const int x = blockIdx.x * AFF_BX + threadIdx.x;
    const int y = blockIdx.y * AFF_BY + threadIdx.y;
    const int b = blockIdx.z;
    const bool in = (x < W) && (y < H);
    if (!in && (x >= tiling.tiles_x * 32 || y >= tiling.tiles_y * 4 * tiling.R)) return;
    const size_t HW = (size_t)H * W;
    const float* __restrict__ ip = src + (size_t)b * 48 * HW + (size_t)y * W + x;
    AffTiling full = tiling;
    if (!in) { full.Wt = tiling.tiles_x * 32; full.Ht = tiling.tiles_y * 4 * tiling.R; }  // partial-tile padding: zeros
    const AffPixel px = aff_pixel(full, b, y, x);
#pragma unroll 8
    for (int p = 0; p < 48; ++p) dst[px.at(tap_seq(p))] = in ? __ldg(ip + (size_t)p * HW) : 0.f;
}

int grid_check(const dim3& g, const char* what) {
    if (g.y > 65535 || g.z > 65535) return set_error(PAMR_ERR_INVALID_ARGUMENT, "%s: H/4 and B must be <= 65535", what);
    return PAMR_OK;
}

}  // namespace

// TMA box starts must be 16-byte aligned (measured: a 4-byte granular start raises an illegal-instruction fault), and a
// window row starts at ((plane * H + y) * pitch + x0 - 24) floats with x0 - 24 a multiple of 8: the tile kernel reads
// the caller's tensor directly when W is a multiple of 4 and its base is 16-byte aligned, else a copy with pitched rows.
bool affinity_image_needs_pitching(const float* img, int W) { return (W & 3) != 0 || ((uintptr_t)img & 15) != 0; }
size_t affinity_pitched_image_bytes(int B, int K, int H, int W) { return sizeof(float) * (size_t)B * K * H * ((W + 3) & ~3); }
bool affinity_tile_kernel_applies(int B, int K, int H, int W, const AffTiling& tiling) {
    return (K == 1 || K == 3) && (tiling.R == 8 || tiling.R == 10) && (unsigned long long)B * K * H * ((W + 3) & ~3) < (1ull << 31);
}

// The two strip regions of the tiled layout (one column, a few rows): a few hundred pixels per image, one thread
// each.  The kernel is latency-bound (~35 us at config 2), so the forward path runs it on its side stream next to
// the tile kernel.
int launch_affinity_strips(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil,
                           const AffTiling& tiling, cudaStream_t s) {
    const long long nstrip = (long long)(W - tiling.Wt) * tiling.Ht + (long long)(H - tiling.Ht) * W;
    if (nstrip <= 0) return PAMR_OK;
    dim3 block(AFF_BX, AFF_BY);
    dim3 grid((unsigned)((nstrip + AFF_BX * AFF_BY - 1) / (AFF_BX * AFF_BY)), 1, B);
    int rc;
    if ((rc = grid_check(grid, "affinity")) != PAMR_OK) return rc;
    affinity_generic_kernel<6, 2><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

// strips_stream: nullptr (strips on s, after the tiles) or a stream that has been forked from s and will be joined
// back into it by the caller
int launch_affinity(const float* img, float* aff, int B, int K, int H, int W, const Dilations& dil,
                    const AffTiling& tiling, float* img_pitched, cudaStream_t s, cudaStream_t strips_stream) {
    const bool tiled = tiling.R > 0;
    int dev = 0;
    PAMR_CUDA_TRY(cudaGetDevice(&dev));
    int rc;
    if (tiled) {
        if (!standard_dilations(dil)) return set_error(PAMR_ERR_INVALID_ARGUMENT, "affinity: tiled layout needs the standard dilations");
        const bool pitching = affinity_image_needs_pitching(img, W);
        if (affinity_tile_kernel_applies(B, K, H, W, tiling) && (!pitching || img_pitched != nullptr)) {
            const float* src = img;
            int pitch = W;
            if (pitching) {
                pitch = (W + 3) & ~3;
                const size_t rows = (size_t)B * K * H;
                pitch_rows_kernel<<<(unsigned)((rows * (pitch >> 2) + 255) / 256 < 148 * 16 ? (rows * (pitch >> 2) + 255) / 256 : 148 * 16), 256, 0, s>>>(img, img_pitched, W, pitch, rows);
                count_launch();
                PAMR_CUDA_TRY(cudaGetLastError());
                src = img_pitched;
            }
            if (tiling.R == 10) rc = (K == 3) ? launch_tile_kernel<10, 3>(src, pitch, aff, B, H, W, tiling, dev, s) : launch_tile_kernel<10, 1>(src, pitch, aff, B, H, W, tiling, dev, s);
            else rc = (K == 3) ? launch_tile_kernel<8, 3>(src, pitch, aff, B, H, W, tiling, dev, s) : launch_tile_kernel<8, 1>(src, pitch, aff, B, H, W, tiling, dev, s);
            if (rc != PAMR_OK) return rc;
        } else {  // other channel counts: every pixel of the tile grid through the generic kernel
            dim3 block(AFF_BX, AFF_BY);
            const int gw = max(tiling.tiles_x * 32, W), gh = max(tiling.tiles_y * 4 * tiling.R, H);
            dim3 grid((gw + AFF_BX - 1) / AFF_BX, (gh + AFF_BY - 1) / AFF_BY, B);
            if ((rc = grid_check(grid, "affinity")) != PAMR_OK) return rc;
            affinity_generic_kernel<6, 1><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
            count_launch();
            PAMR_CUDA_TRY(cudaGetLastError());
            return PAMR_OK;
        }
        return launch_affinity_strips(img, aff, B, K, H, W, dil, tiling, strips_stream != nullptr ? strips_stream : s);
    }
    if (standard_dilations(dil) && K <= AFF_MAXK) {
        dim3 sblock(SA_BX, SA_BY);
        dim3 sgrid((W + SA_BX - 1) / SA_BX, (H + SA_BY - 1) / SA_BY, B);
        if ((rc = grid_check(sgrid, "affinity")) != PAMR_OK) return rc;
        const size_t smem = sizeof(float) * (size_t)K * SA_H * SA_W;
        static std::atomic<int> attr_set[64];
        if (dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
            PAMR_CUDA_TRY(cudaFuncSetAttribute(affinity_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(sizeof(float) * AFF_MAXK * SA_H * SA_W)));
            if (dev < 64) attr_set[dev].store(1, std::memory_order_release);
        }
        affinity_smem_kernel<<<sgrid, sblock, smem, s>>>(img, aff, K, H, W);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        return PAMR_OK;
    }
    dim3 block(AFF_BX, AFF_BY);
    dim3 grid((W + AFF_BX - 1) / AFF_BX, (H + AFF_BY - 1) / AFF_BY, B);
    if ((rc = grid_check(grid, "affinity")) != PAMR_OK) return rc;
    if (dil.nd <= 6) affinity_generic_kernel<6, 0><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
    else affinity_generic_kernel<PAMR_MAX_DILATIONS, 0><<<grid, block, 0, s>>>(img, aff, K, H, W, dil, tiling);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_local_std(const float* img, float* sd, int B, int K, int H, int W, const Dilations& dil, cudaStream_t s) {
    dim3 block(AFF_BX, AFF_BY);
    dim3 grid((W + AFF_BX - 1) / AFF_BX, (H + AFF_BY - 1) / AFF_BY, B * K);
    int rc;
    if ((rc = grid_check(grid, "local_std")) != PAMR_OK) return rc;
    if (dil.nd <= 6) local_std_kernel<6><<<grid, block, 0, s>>>(img, sd, K, H, W, dil);
    else local_std_kernel<PAMR_MAX_DILATIONS><<<grid, block, 0, s>>>(img, sd, K, H, W, dil);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

int launch_aff_relayout(const float* aff_std, float* aff_tiled, int B, int H, int W, const AffTiling& tiling,
                        cudaStream_t s) {
    dim3 block(AFF_BX, AFF_BY);
    const int gw = max(tiling.tiles_x * 32, W), gh = max(tiling.tiles_y * 4 * tiling.R, H);
    dim3 grid((gw + AFF_BX - 1) / AFF_BX, (gh + AFF_BY - 1) / AFF_BY, B);
    int rc;
    if ((rc = grid_check(grid, "affinity relayout")) != PAMR_OK) return rc;
    aff_relayout_kernel<<<grid, block, 0, s>>>(aff_std, aff_tiled, H, W, tiling);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace pamr
