// Tuned sm_100a propagation kernel for the standard dilation set [1,2,4,8,12,24]
// (reference models/mods/pamr.py:138-140 with core/config.py:92).
//
//   M'[b,c,y,x] = sum_{p<48} w[b,p,y,x] * M[b,c,clamp(y+dy_p),clamp(x+dx_p)]
//
// Design:
//  * persistent CTAs, one per SM; a CTA owns a 32 x (4*R) pixel tile (R = 8, 9 or 10 rows per thread);
//  * 12 compute warps in 3 groups of 4: lane = x, each thread owns a vertical strip of R pixels, so
//    that one shared-memory load feeds up to 3*R/(R+2d) taps (register reuse along y): ~32 LDS per
//    pixel-class instead of 48.  Group g takes the class planes c = g (mod 3);
//  * the tile's 48 affinity weights per pixel (48*R words per thread) are read once per tile and
//    parked in Tensor Memory (tcgen05.st), 1 TMEM lane per thread, shared by the three groups (warps
//    w, w+4, w+8 address the same lane quarter); every class pass re-reads them with tcgen05.ld -- a
//    data path that does not compete with LDS -- instead of holding them in registers or
//    re-reading them through L1.  The next tile's weights are pulled into L2 with a bulk prefetch
//    three classes before the tile boundary;
//  * a producer warp streams the class planes of the tile (+24 px halo, 80 x (4R+48) floats)
//    through a 4-slot shared-memory ring with TMA (cp.async.bulk.tensor.3d) + mbarriers; TMA
//    zero-fills outside the image; replicate padding (pamr.py:50) in x is a per-lane clamped column
//    offset, in y the consuming group patches the halo rows of top / bottom tiles before it computes;
//  * FP32 math as packed FFMA2 over adjacent rows (a scalar FFMA with three distinct source
//    registers issues only every ~1.8 cycles); results are stored with coalesced 128-byte rows; the
//    per-(b,c) max for pseudo_gtmask is fused into the last iteration (warp reduce + atomicMax);
//  * remainders of at most 8 rows / columns (W = H = 321): the row strip is computed by the CTAs that
//    are idle in the tile kernel's last wave (or a small launch when there are none), the column strip
//    by a small kernel on a second stream, concurrently with the tile kernel.
#include <cuda.h>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "pamr_common.cuh"

namespace pamr {

namespace {

constexpr int TX = 32;
constexpr int HALO = 24;
constexpr int WIN_W = TX + 2 * HALO;  // 80 floats = 320 B rows in shared memory
constexpr int NW = 4;                 // warps per compute group (= TMEM lane quarters)
constexpr int NG = 3;                 // compute groups that share the tile's weights in TMEM (class c -> group c % NG)
constexpr int NWC = NG * NW;          // compute warps
constexpr int NSLOT = 4;              // class-plane slots in the shared-memory ring
// mbarriers of the ring: sequence number n uses slot n % NSLOT but barrier pair n % NBAR.  Waits are by
// phase PARITY, which is only sound if a waiter can never be a whole phase ahead of the barrier.  With one
// barrier per slot and 4 slots, a group that has finished class k-3 tests the barrier of class k while the
// load of class k-4 (same slot, other group) may -- once in ~1e7 passes, when that load straggles -- still
// be in flight; the parity test then reports the OLD phase as "complete" and the group computes on the wrong
// plane (measured: 0.2-2.5 % of the forward calls had one wrong tile-class; 0 of 2500 with 8 slots).  Eight
// barrier pairs put 8 sequence numbers between two uses of a barrier, as in the 8-slot ring.
constexpr int NBAR = 8;
static_assert(NBAR % NSLOT == 0 && NBAR >= 2 * NSLOT, "barrier ring must cover at least two uses of every slot");
#ifndef PAMR_CC
#define PAMR_CC 1
#endif
constexpr int CC = PAMR_CC;           // class planes per compute pass
constexpr int NTHREADS = (NWC + 1) * 32;
constexpr int WB = 16;                // weights per tcgen05.ld batch

__host__ __device__ constexpr int dil_of(int id) { return id == 0 ? 1 : id == 1 ? 2 : id == 2 ? 4 : id == 3 ? 8 : id == 4 ? 12 : 24; }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// non-blocking poll (try_wait may suspend the thread for a while before answering)
__device__ __forceinline__ bool mbar_poll(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking wait: try_wait with a suspend-time hint, so that the waiting thread sleeps in hardware until
// the phase completes instead of re-issuing try_wait / branch pairs that take issue slots from the
// compute warps on its scheduler (producer warp, measured per launch: 4.15 M spin iterations without
// the hint, 0.31 M with it).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
    } while (ok == 0);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
#ifdef PAMR_BODY_NO_TMEM
    for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(taddr + j) * 1e-30f;
    return;
#endif
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
// tcgen05.wait::ld, with the batch's registers tied through the asm so that no consumer of the
// loaded values can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait_ld(float (&r)[16]) {
#ifdef PAMR_BODY_NO_TMEM
    return;
#endif
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]));
}

// ---------------------------------------------------------------- shared-memory layout
template <int R>
struct Cfg {
    static constexpr int TY = NW * R;
    static constexpr int WIN_H = TY + 2 * HALO;
    static constexpr int SLOT_FLOATS = WIN_W * WIN_H;
    static constexpr int SLOT_BYTES = SLOT_FLOATS * 4;
    static constexpr size_t SMEM_BYTES = (size_t)NSLOT * SLOT_BYTES + 1024;
};

struct Ctrl {  // lives in the last 1 KB of dynamic shared memory
    unsigned long long tma_bar[NBAR];
    unsigned long long empty_bar[NBAR];
    uint32_t tmem_base;
};

struct Params {
    const float* aff;  // tile-major affinity (pamr_common.cuh), tiles_x_aff tile columns
    float* dst;        // [B,C,H,dst_pitch]
    unsigned* cls_max; // [B,C] or nullptr
    int stagger_cta_ns, stagger_grp_ns;  // experiment knobs (env PAMR_B200_STAGGER_CTA / _GRP)
    int dbg_cta;       // CTA that records the timeline
    int exp_flags;     // experiments (results invalid): 1 skip weight fill, 2 skip global stores, 4 skip halo patch, 8 skip waits
    int pf_class;      // class index at whose TMA issue the next tile's weights are prefetched into L2 (-1: never)
    long long* dbg;    // nullptr, or timeline buffer (debug hook pamr_debug_set_timeline): 2 x 4096 x {clock, code}
    int dst_pitch;
    int B, C, H, W;
    int tiles_x, tiles_y, ntiles;  // tiles of this launch (tiles_x may exclude the remainder strip)
    int tiles_x_aff, tiles_y_aff;  // tile grid of the affinity layout (covers the whole image)
    const float* src;              // source mask [B,C,H,src_pitch] (for the remainder strips; tiles come through TMA)
    int src_pitch;
    int Wt, Ht;                    // the tiles cover [0,Wt) x [0,Ht); the producer warp computes the remainder strips
    int strip_items;               // number of 32-pixel strip work items (all CTAs together)
    int tail_cta0;                 // row strip inside the tile kernel: CTAs >= tail_cta0 (one tile fewer than the
                                   // others) work through the strip_items row items after their last tile; -1: off
};

// ---------------------------------------------------------------- TMEM weight layout
// Per thread (= TMEM lane) the 48*R weights of its R pixels are laid out in consumption order:
//   columns [0, 12R)            centre column (b = 0): tap sequence s = 2*id + (a>0), R rows each
//   columns SIDE0 + 32*g ...    side group g = 6*bi + id (bi = 0: b = -1, bi = 1: b = +1):
//                               (a+1)*R + i for a = -1,0,+1  (3R <= 30 of the 32 columns used)
// so that every side group is one aligned tcgen05.ld.x32 and the b = -1 / b = +1 halves can share
// one (rolled) copy of the code: the unrolled loop body must stay inside the 32 KB instruction cache.
template <int R>
struct TmemLayout {
    static constexpr int RS = (R + 1) / 2 * 2;                 // columns per tap: R rounded up to even, so
                                                               // that row pairs (i, i+1) are aligned register pairs
    static constexpr int CPAD = (12 * RS + WB - 1) / WB * WB;  // centre columns padded to a batch
    static constexpr int SIDE0 = CPAD;
    static constexpr int NCOLS = CPAD + 12 * 32;              // <= 512
    static_assert(NCOLS <= 512, "TMEM columns");
};

template <int R>
__host__ __device__ constexpr int seq_col(int s) {
    return (s < 12) ? s * TmemLayout<R>::RS
                    : TmemLayout<R>::SIDE0 + ((s - 12) / 3) * 32 + ((s - 12) % 3) * TmemLayout<R>::RS;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
#ifdef PAMR_BODY_NO_TMEM
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(taddr + j) * 1e-30f;
    return;
#endif
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),
          "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),
          "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld32(float (&r)[32]) {
#ifdef PAMR_BODY_NO_TMEM
    return;
#endif
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]),
                   "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]),
                   "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]),
                   "+f"(r[30]), "+f"(r[31]));
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float r0) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(r0));
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float r0, float r1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "f"(r0), "f"(r1));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(r[0]),
                 "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]));
}

// ---------------------------------------------------------------- compute body
// One pass over the 48 taps for N (<= CC) class planes resident in the ring.
// sp[n] points at this thread's pixel (row R*warp, column lane) of plane n inside its slot,
// i.e. slot + (R*warp + HALO)*WIN_W + lane + HALO; neighbours are immediate offsets.
// Two FMAs on adjacent rows as one packed FFMA2 (fma.rn.f32x2).  A scalar FFMA whose three source
// registers are all distinct issues only every ~1.8 cycles per SM sub-partition (measured,
// tools/ubench3.cu); the packed form retires two FMAs per ~2.4 cycles.  The mov.b64 packs are
// free when ptxas allocates the operands as aligned register pairs (TMEM batches, the row strip
// and the accumulators all are, for even R and even row shifts).
__device__ __forceinline__ void fma2(float& a0, float& a1, float w0, float w1, float v0, float v1) {
    unsigned long long A, W2, V2;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(W2) : "f"(w0), "f"(w1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(V2) : "f"(v0), "f"(v1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(W2), "l"(V2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(A));
}

// xg = this lane's image column, W = image width: the side columns are addressed with a per-lane
// offset clamp(xg +- d, 0, W-1) - xg, which implements replicate padding in x for free (the offset is
// a register either way), so only rows above / below the image ever need patching in shared memory.
template <int R, int N>
__device__ __forceinline__ void compute_pass(const float* (&sp)[CC], uint32_t tbase, float (&acc)[CC][R], int xg,
                                             int W) {
    using L = TmemLayout<R>;
#ifdef PAMR_NO_FFMA2
    constexpr bool kPacked = false;
#else
    constexpr bool kPacked = true;  // odd R: the last row stays scalar
#endif
    constexpr int NCB = L::CPAD / WB;  // centre batches
    float wc[2][WB];
    float ws[2][32];
    tmem_ld16(tbase, wc[0]);

    // ---- centre column (b = 0): rows y+-d for all dilations, merged into one register strip
    {
        float v[CC][R + 2 * HALO];
#pragma unroll
        for (int r = -HALO; r < R + HALO; ++r) {
            bool need = false;
#pragma unroll
            for (int id = 0; id < 6; ++id) {
                const int d = dil_of(id);
                need = need || (r >= -d && r < R - d) || (r >= d && r < R + d);
            }
            if (need) {
#pragma unroll
                for (int n = 0; n < N; ++n) v[n][r + HALO] = sp[n][r * WIN_W];
            }
        }
        int bcur = -1;  // weight batch currently held (compile-time after full unrolling)
#pragma unroll
        for (int id = 0; id < 6; ++id) {
            const int d = dil_of(id);
#pragma unroll
            for (int a = -1; a <= 1; a += 2) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const bool pair = kPacked && (d % 2 == 0) && (i + 1 < R || i % 2 == 1);  // rows (i, i+1) as one FFMA2
                    if (pair && (i % 2 == 1)) continue;  // odd row handled with its even partner
                    const int q = (2 * id + (a > 0 ? 1 : 0)) * L::RS + i;
                    if (q / WB != bcur) {  // batch boundary: wait for this batch, prefetch the next one
                        bcur = q / WB;
                        tmem_wait_ld(wc[bcur & 1]);
                        if (bcur + 1 < NCB) tmem_ld16(tbase + (bcur + 1) * WB, wc[(bcur + 1) & 1]);
                        else tmem_ld32(tbase + L::SIDE0, ws[0]);  // first side group
                    }
                    if (pair) {
#pragma unroll
                        for (int n = 0; n < N; ++n)
                            fma2(acc[n][i], acc[n][i + 1], wc[bcur & 1][q % WB], wc[bcur & 1][q % WB + 1],
                                 v[n][i + a * d + HALO], v[n][i + 1 + a * d + HALO]);
                    } else {
                        const float w = wc[bcur & 1][q % WB];
#pragma unroll
                        for (int n = 0; n < N; ++n) acc[n][i] = fmaf(w, v[n][i + a * d + HALO], acc[n][i]);
                    }
                }
            }
        }
    }
    // ---- side columns: bi = 0 (b = -1), bi = 1 (b = +1); rolled so that both share one code copy
#pragma unroll 1
    for (int bi = 0; bi < 2; ++bi) {
        const int sgn = bi * 2 - 1;
#pragma unroll
        for (int id = 0; id < 6; ++id) {
            const int d = dil_of(id);
            float (&w)[32] = ws[id & 1];
            tmem_wait_ld32(w);
            // prefetch the next side group (the last one of b = +1 has no successor)
            if (id < 5) tmem_ld32(tbase + L::SIDE0 + (bi * 6 + id + 1) * 32, ws[(id + 1) & 1]);
            else if (bi == 0) tmem_ld32(tbase + L::SIDE0 + 6 * 32, ws[0]);
            float v[CC][R + 2 * HALO];
            const int coff = min(max(xg + sgn * d, 0), W - 1) - xg;
#pragma unroll
            for (int r = -d; r < R + d; ++r) {
                const bool need = (r < R - d) || (r >= 0 && r < R) || (r >= d);
                if (need) {
#pragma unroll
                    for (int n = 0; n < N; ++n) v[n][r + HALO] = sp[n][r * WIN_W + coff];
                }
            }
#pragma unroll
            for (int a = -1; a <= 1; ++a) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const bool pair = kPacked && ((a * d) % 2 == 0) && (i + 1 < R || i % 2 == 1);  // rows (i, i+1)
                    if (pair && (i % 2 == 1)) continue;
                    if (pair) {
#pragma unroll
                        for (int n = 0; n < N; ++n)
                            fma2(acc[n][i], acc[n][i + 1], w[(a + 1) * L::RS + i], w[(a + 1) * L::RS + i + 1],
                                 v[n][i + a * d + HALO], v[n][i + 1 + a * d + HALO]);
                    } else {
                        const float wv = w[(a + 1) * L::RS + i];
#pragma unroll
                        for (int n = 0; n < N; ++n) acc[n][i] = fmaf(wv, v[n][i + a * d + HALO], acc[n][i]);
                    }
                }
            }
        }
    }
}

template <int R>
__device__ __forceinline__ bool needs_patch(int x0, int y0, int H, int W) {
    (void)x0; (void)W;  // replicate padding in x is handled by the per-lane column offsets
    return y0 < HALO || y0 + Cfg<R>::TY + HALO > H;
}

// Border tiles: TMA zero-filled everything outside the image; overwrite it with the clamped
// (replicate-padded, pamr.py:50) value.  Window element (wy,wx) <-> image pixel (y0-24+wy, x0-24+wx).
// Executed by the NW warps of the compute group that is about to read the slot (wq = warp in group).
// Every store targets an out-of-image element and every load an in-image one, so the four warps
// need no ordering among themselves; the caller synchronises the group afterwards.
template <int R>
__device__ __forceinline__ void patch_window(float* slot, int x0, int y0, int H, int W, int wq, int lane) {
    constexpr int WIN_H = Cfg<R>::WIN_H;
    const int vx0 = max(0, HALO - x0), vx1 = min(WIN_W, W - x0 + HALO);  // in-image columns [vx0, vx1)
    const int vy0 = max(0, HALO - y0), vy1 = min(WIN_H, H - y0 + HALO);  // in-image rows    [vy0, vy1)
    // rows above / below the image: full width; the source row (first / last in-image row, columns
    // clamped) is read once and then stored to every row this warp owns
#pragma unroll
    for (int k = 0; k < (WIN_W + 31) / 32; ++k) {
        const int wx = 32 * k + lane;
        if (wx < WIN_W) {
            const int sx = min(max(wx, vx0), vx1 - 1);
            if (vy0 > 0) {
                const float v = slot[vy0 * WIN_W + sx];
                for (int wy = wq; wy < vy0; wy += NW) slot[wy * WIN_W + wx] = v;
            }
            if (vy1 < WIN_H) {
                const float v = slot[(vy1 - 1) * WIN_W + sx];
                for (int wy = vy1 + wq; wy < WIN_H; wy += NW) slot[wy * WIN_W + wx] = v;
            }
        }
    }
}

__device__ __forceinline__ void compute_bar_sync() {  // the NWC compute warps only (named barrier 1)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(NWC * 32) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Remainder strips.  The tiles cover [0,Wt) x [0,Ht); when W or H leaves a remainder of at most 8
// pixels (W = H = 321 -> one column and one row) that remainder is not worth a tile row/column of
// its own.  It is cut into work items of 32 pixels -- item < n_col: 32 rows of one strip column,
// otherwise 32 columns of one strip row -- which the compute warps pick up at tile boundaries
// (one pixel per lane, neighbours and weights straight from global memory / L2, clamped coordinates).
template <int R>
__device__ __forceinline__ void strip_item(const Params& prm, const int Wt, int item, int lane) {
    const int C = prm.C, H = prm.H, W = prm.W;
    const int wcols = W - Wt, hrows = H - prm.Ht;
    const int yblocks = (H + 31) / 32, xblocks = (Wt + 31) / 32;
    const int per_plane = wcols * yblocks + hrows * xblocks;
    const int plane = item / per_plane, r = item % per_plane;  // plane = b*C + c
    int x, y;
    if (r < wcols * yblocks) {
        x = Wt + r / yblocks;
        y = (r % yblocks) * 32 + lane;
    } else {
        const int q = r - wcols * yblocks;
        y = prm.Ht + q / xblocks;
        x = (q % xblocks) * 32 + lane;
        if (x >= Wt) x = W;  // beyond the row strip (those columns belong to the column strip)
    }
    const bool valid = (y < H) && (x < W);
    const int yc = min(y, H - 1), xc = min(x, W - 1);
    const int b = plane / C;
    const float* __restrict__ pl = prm.src + (size_t)plane * H * prm.src_pitch;
    const AffTiling tl{R, prm.tiles_x_aff, prm.tiles_y_aff, 0, 0};
    const float* __restrict__ wp = prm.aff + aff_tiled_index(tl, b, 0, yc, xc);
    float acc = 0.f;  // one FMA chain in tap-sequence order: bit-identical to the tile kernel's result
#pragma unroll
    for (int s = 0; s < 48; ++s) {
        const int p = seq_tap(s), d = dil_of(p >> 3), j = p & 7;
        const int yy = clampi(yc + tap_dy(j) * d, 0, H - 1);
        const int xx = clampi(xc + tap_dx(j) * d, 0, W - 1);
        acc = fmaf(__ldg(wp + s * (R * 32)), __ldg(pl + (size_t)yy * prm.src_pitch + xx), acc);
    }
    if (valid) prm.dst[((size_t)plane * H + y) * prm.dst_pitch + x] = acc;
    if (prm.cls_max != nullptr) {
        const unsigned m = __reduce_max_sync(0xffffffffu, valid ? ordered_from_float(acc) : 0u);
        if (lane == 0 && m != 0u) atomicMax(prm.cls_max + plane, m);
    }
}

template <int R>
__global__ void __launch_bounds__(NTHREADS, 1)
propagate_sm100_kernel(const __grid_constant__ CUtensorMap tmap, const Params prm) {
    using C_ = Cfg<R>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* slots = reinterpret_cast<float*>(smem_raw);
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + (size_t)NSLOT * C_::SLOT_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = prm.C, H = prm.H, W = prm.W;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBAR; ++s) {
            mbar_init(smem_u32(&ctrl->tma_bar[s]), 1);
            mbar_init(smem_u32(&ctrl->empty_bar[s]), NW);  // the NW warps of the group that read the slot
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctrl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int my_tiles = (prm.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tiles_per_img = prm.tiles_x * prm.tiles_y;

    if (warp == NWC) {
        // ===================== producer warp: TMA issue =====================
        // Sequence number n = (tile_iter, class) -> slot n % NSLOT, barrier pair n % NBAR.  A consumer
        // group waits for tma_bar (bytes landed), patches the halo itself if the tile touches the image
        // border, computes, and releases the slot through empty_bar; the producer may refill slot
        // n % NSLOT once sequence number n - NSLOT has been released.
        const long long total = (long long)my_tiles * C;
        int pn = 0;
        for (long long n_issue = 0; n_issue < total; ++n_issue) {
            const int s = (int)(n_issue % NSLOT), bi = (int)(n_issue % NBAR);
            const int ti = (int)(n_issue / C), c = (int)(n_issue % C);
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int x0 = (t % prm.tiles_x) * TX, y0 = (t / prm.tiles_x) * C_::TY;
            if (lane == 0 && n_issue >= NSLOT) {  // the previous occupant of this slot has been consumed
                const long long prev = n_issue - NSLOT;
                mbar_wait_sleep(smem_u32(&ctrl->empty_bar[prev % NBAR]), (uint32_t)(prev / NBAR) & 1u);
            }
            if (lane == 0) {
                const uint32_t bar = smem_u32(&ctrl->tma_bar[bi]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(bar, C_::SLOT_BYTES);
                tma_load_3d(smem_u32(slots + (size_t)s * C_::SLOT_FLOATS), &tmap, bar, x0 - HALO, y0 - HALO, b * C + c);
                if (prm.dbg != nullptr && (int)blockIdx.x == prm.dbg_cta && pn < 4096) {  // timeline: issue time of class n
                    prm.dbg[(2 * 4096 + pn) * 2] = clock64();
                    prm.dbg[(2 * 4096 + pn) * 2 + 1] = 1000 + n_issue;
                    ++pn;
                }
            }
            // Late in a tile, pull the NEXT tile's affinity weights (one contiguous 48*R*128-byte block
            // per lane quarter of the tile-major layout) from HBM into L2, so that the TMEM fill at the
            // tile boundary -- when every SM asks for its weights at once -- mostly hits L2.
            if (c == prm.pf_class && ti + 1 < my_tiles && lane < 4) {
                const int ntile = (int)blockIdx.x + (ti + 1) * (int)gridDim.x;
                const int nb = ntile / tiles_per_img, nt = ntile % tiles_per_img;
                const float* wp = prm.aff + (((((size_t)nb * prm.tiles_y_aff + nt / prm.tiles_x) * prm.tiles_x_aff +
                                               nt % prm.tiles_x) * 4 + lane) * 48 * R) * 32;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wp), "r"(48 * R * 32 * 4) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ===================== compute warps: NG groups x NW warps =====================
        const int grp = warp / NW, wq = warp % NW;  // wq = TMEM lane quarter; all groups share the weights
        const uint32_t tbase = ctrl->tmem_base + ((uint32_t)(wq * 32) << 16);
        int dn = 0;  // timeline events written by this group's leader lane of CTA 0
        if ((blockIdx.x & 1) && prm.stagger_cta_ns > 0) __nanosleep(prm.stagger_cta_ns);
#define PAMR_EV(code)                                                                         \
    do {                                                                                      \
        if (prm.dbg != nullptr && (int)blockIdx.x == prm.dbg_cta && grp < 2 && wq == 0 && lane == 0 && dn < 4096) {      \
            prm.dbg[((grp & 1) * 4096 + dn) * 2] = clock64();                                        \
            prm.dbg[((grp & 1) * 4096 + dn) * 2 + 1] = (code);                                       \
            ++dn;                                                                             \
        }                                                                                     \
    } while (0)
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int x0 = (t % prm.tiles_x) * TX, y0 = (t / prm.tiles_x) * C_::TY;
            const int x = x0 + lane, yw = y0 + wq * R;
            const bool xok = x < W;
            const long long seq0 = (long long)ti * C;
            const bool border = needs_patch<R>(x0, y0, H, W);

            // ---- park the tile's 48*R weights per thread in TMEM (layout: TmemLayout); each group
            //      loads its share (48/NG) of the taps for the lanes it shares with its sibling warps
            PAMR_EV(1);
            if (ti > 0) compute_bar_sync();  // nobody still reads the previous tile's weights
            PAMR_EV(2);
            {
                // tile-major layout: this thread's weights are ap[(s*R + i)*32], s = tap sequence index
                const int ty = t / prm.tiles_x, tx = t % prm.tiles_x;
                const float* __restrict__ ap =
                    prm.aff + (((((size_t)b * prm.tiles_y_aff + ty) * prm.tiles_x_aff + tx) * 4 + wq) * 48 * R) * 32 + lane;
                // FILL_TAPS*R loads are in flight before the first TMEM store: memory-level
                // parallelism is what bounds this phase (the accumulators are not live here)
                constexpr int TAPS_PER_GROUP = 48 / NG, FILL_TAPS = 8;
#pragma unroll
                for (int h = 0; h < ((prm.exp_flags & 1) ? 0 : TAPS_PER_GROUP / FILL_TAPS); ++h) {
                    const int s0 = TAPS_PER_GROUP * grp + h * FILL_TAPS;  // grp is warp-uniform; offsets below are immediates
                    const float* __restrict__ bp = ap + (size_t)s0 * R * 32;
                    float r[FILL_TAPS][R];
#pragma unroll
                    for (int k = 0; k < FILL_TAPS; ++k)
#pragma unroll
                        for (int i = 0; i < R; ++i) r[k][i] = __ldg(bp + (k * R + i) * 32);
#pragma unroll
                    for (int k = 0; k < FILL_TAPS; ++k) {
                        const uint32_t col = tbase + seq_col<R>(s0 + k);
                        tmem_st8(col, r[k]);
                        if (R == 9) tmem_st1(col + 8, r[k][R - 1]);
                        if (R == 10) tmem_st2(col + 8, r[k][R - 2], r[k][R - 1]);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            PAMR_EV(3);
            compute_bar_sync();  // both halves of the weights are visible to both groups
            PAMR_EV(4);
            if (grp != 0 && prm.stagger_grp_ns > 0) __nanosleep(prm.stagger_grp_ns * grp);

            int probe = 0;  // 1: the barriers of this group's next class were already seen complete
            for (int k = grp; k < C; k += NG) {
                const int c0 = k, n = 1;
                const long long sq = seq0 + k;
                const int s = (int)(sq % NSLOT), bi = (int)(sq % NBAR);
                const uint32_t par = (uint32_t)(sq / NBAR) & 1u;
                PAMR_EV(100 + k);
                if (!probe && !(prm.exp_flags & 8)) mbar_wait_sleep(smem_u32(&ctrl->tma_bar[bi]), par);  // bytes landed
                if (border && !(prm.exp_flags & 4)) {  // replicate padding: the group patches the halo of its own slot
                    patch_window<R>(slots + (size_t)s * C_::SLOT_FLOATS, x0, y0, H, W, wq, lane);
                    // immediate barrier ids: with a register id ptxas reserves all 16 named barriers and no
                    // other CTA (the concurrent column strip) could share the SM
                    if (grp == 0) asm volatile("bar.sync 2, %0;" ::"n"(NW * 32) : "memory");
                    else if (grp == 1) asm volatile("bar.sync 3, %0;" ::"n"(NW * 32) : "memory");
                    else asm volatile("bar.sync 4, %0;" ::"n"(NW * 32) : "memory");
                }
                const float* sp[CC];
#pragma unroll
                for (int j = 0; j < CC; ++j)
                    sp[j] = slots + (size_t)s * C_::SLOT_FLOATS + (wq * R + HALO) * WIN_W + lane + HALO;
                PAMR_EV(6);
                float acc[CC][R];
#pragma unroll
                for (int j = 0; j < CC; ++j)
#pragma unroll
                    for (int i = 0; i < R; ++i) acc[j][i] = 0.f;
                compute_pass<R, 1>(sp, tbase, acc, x, W);
                PAMR_EV(9);
                // release the slot as early as possible
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctrl->empty_bar[bi]));
                // Probe the next class's barriers now, without blocking: an mbarrier test takes a few
                // hundred cycles when the LSU queues are full of LDS, and that latency then overlaps the
                // stores below instead of sitting at the head of the next pass.
                probe = 0;
                if (k + NG < C) {
                    const long long sq2 = sq + NG;
                    const uint32_t par2 = (uint32_t)(sq2 / NBAR) & 1u;
                    probe = (int)mbar_poll(smem_u32(&ctrl->tma_bar[sq2 % NBAR]), par2);
                }
                PAMR_EV(7);
                // ---- store (coalesced 128 B per row) and optional class max
#pragma unroll
                for (int j = 0; j < CC; ++j) {
                    if (j < n) {
                        float* __restrict__ op = prm.dst + ((size_t)(b * C + c0 + j) * H + yw) * prm.dst_pitch + x;
                        const int nrow = xok ? min(R, H - yw) : 0;  // rows of this thread inside the image
                        if (!(prm.exp_flags & 2)) {
                            // running pointer: one 64-bit add per row instead of a 64-bit multiply-add chain
                            const size_t pitch = (size_t)prm.dst_pitch;
                            if (nrow == R) {
#pragma unroll
                                for (int i = 0; i < R; ++i, op += pitch) *op = acc[j][i];
                            } else {
#pragma unroll
                                for (int i = 0; i < R; ++i, op += pitch)
                                    if (i < nrow) *op = acc[j][i];
                            }
                        }
                        if (prm.cls_max != nullptr) {  // last iteration only: keep its ALU work out of the other nine
                            unsigned mx = 0u;
#pragma unroll
                            for (int i = 0; i < R; ++i)
                                if (i < nrow) mx = max(mx, ordered_from_float(acc[j][i]));
                            mx = __reduce_max_sync(0xffffffffu, mx);
                            if (lane == 0 && mx != 0u) atomicMax(prm.cls_max + (size_t)b * C + c0 + j, mx);
                        }
                    }
                }
                PAMR_EV(8);
            }
        }
#undef PAMR_EV
        // ---- row strip in the tail: the CTAs that own one tile fewer than the rest would idle during
        //      the last wave; they compute the row strip y in [Ht,H) instead (no separate launch)
        if (prm.tail_cta0 >= 0 && (int)blockIdx.x >= prm.tail_cta0) {
            const int nw = ((int)gridDim.x - prm.tail_cta0) * NWC;
            for (int item = ((int)blockIdx.x - prm.tail_cta0) * NWC + warp; item < prm.strip_items; item += nw)
                strip_item<R>(prm, W, item, lane);
        }
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(ctrl->tmem_base), "r"(512));
}

// ---- remainder strips as small kernels (row strip: only when the tile kernel's last wave has no idle CTAs) ----
// (Alternatives measured on B200 and rejected: the same work inside the persistent kernel by the
//  producer warp -- its ~100 global loads per item queue behind the compute warps' LDS traffic -- or by
//  the compute warps at every tile boundary -- it lengthens every boundary; the remainder column fused
//  into the last full tile column -- a bank-conflicted extra pass; partial tiles cost a whole tile column.)

// Row strip: y in [Ht,H), all x.  One warp per 32 consecutive pixels of a row (coalesced).
template <int R>
__global__ void __launch_bounds__(128) strip_rows_kernel(const Params prm) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < prm.strip_items; item += gridDim.x * wpb)
        strip_item<R>(prm, prm.Wt, item, lane);
}

// Column strip: x in [Wt,W), all y.  Neighbours of a column of pixels lie in different rows, i.e.
// in different cache lines, so one CTA stages the last 24 + (W-Wt) columns of 32 + 48 rows of all C
// class planes in shared memory with coalesced row reads (one warp per class plane), then every
// warp computes 32 pixels (lane = row) x (W-Wt) columns of its class from shared memory.
constexpr int SC_ROWS = 32;                      // pixel rows per CTA
constexpr int SC_WIN_H = SC_ROWS + 2 * HALO;     // 80
constexpr int SC_WIN_W = 32;                     // staged columns [W-32, W)  (needs W >= 32)
constexpr int SC_LOADS = 20;                     // row loads in flight per thread while staging (SC_WIN_H % SC_LOADS == 0)
constexpr int SC_CG = 8;                         // class planes staged at a time (one warp each)
template <int R>
__global__ void __launch_bounds__(SC_CG * 32, 5) strip_cols_kernel(const Params prm) {  // (.,5): <= 48 registers, so that a CTA fits next to a resident tile CTA
    extern __shared__ float sc_smem[];  // [SC_CG][SC_WIN_H][SC_WIN_W + 1] then weights [wc][48][32]
    const int C = prm.C, H = prm.H, W = prm.W, wc = W - prm.Wt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int yb = blockIdx.x * SC_ROWS, b = blockIdx.y;
    constexpr int PITCH = SC_WIN_W + 1;
    float* wsm = sc_smem + (size_t)SC_CG * SC_WIN_H * PITCH;
    const int xs0 = W - SC_WIN_W;  // first staged column
    // the weights of the strip pixels: wsm[(xi*48 + s)*32 + row]
    const AffTiling tl{R, prm.tiles_x_aff, prm.tiles_y_aff, 0, 0};
    for (int e = threadIdx.x; e < wc * 48 * SC_ROWS; e += blockDim.x) {
        const int row = e % SC_ROWS, s = (e / SC_ROWS) % 48, xi = e / (SC_ROWS * 48);
        wsm[e] = __ldg(prm.aff + aff_tiled_index(tl, b, s, min(yb + row, H - 1), prm.Wt + xi));
    }
    const int y = yb + lane;
    float* mywin = sc_smem + (size_t)warp * SC_WIN_H * PITCH;
    for (int c0 = 0; c0 < C; c0 += SC_CG) {
        const int c = c0 + warp;
        if (c < C) {  // stage this warp's class plane window (rows clamped: replicate padding)
            const float* __restrict__ pl = prm.src + ((size_t)b * C + c) * H * prm.src_pitch;
            // SC_LOADS row loads in flight at a time: the kernel overlaps the tile kernel on a second stream and is
            // kept at <= 48 registers so that its CTAs can be placed next to a resident tile CTA (12 K registers
            // are free there).  321x321 B=16 forward: 8 loads 2.83 ms, 16: 2.82, 20: 2.77, 40 (64 registers): 2.80
#pragma unroll 1
            for (int r0 = 0; r0 < SC_WIN_H; r0 += SC_LOADS) {
                float v[SC_LOADS];
#pragma unroll
                for (int q = 0; q < SC_LOADS; ++q)
                    v[q] = __ldg(pl + (size_t)clampi(yb - HALO + r0 + q, 0, H - 1) * prm.src_pitch + xs0 + lane);
#pragma unroll
                for (int q = 0; q < SC_LOADS; ++q) mywin[(r0 + q) * PITCH + lane] = v[q];
            }
        }
        __syncthreads();  // also covers the weights on the first round
        if (c < C) {
            const float* win = mywin + (lane + HALO) * PITCH;
            for (int xi = 0; xi < wc; ++xi) {
                const int xl = prm.Wt + xi - xs0;  // column inside the staged window
                float acc = 0.f;  // tap-sequence order, like the tile kernel (bit-identical results)
#pragma unroll
                for (int sq = 0; sq < 48; ++sq) {
                    const int p = seq_tap(sq), d = dil_of(p >> 3), j = p & 7;
                    const int xx = min(xl + tap_dx(j) * d, SC_WIN_W - 1);  // clamp at the right image border
                    acc = fmaf(wsm[(xi * 48 + sq) * SC_ROWS + lane], win[tap_dy(j) * d * PITCH + xx], acc);
                }
                const bool valid = y < H;
                if (valid) prm.dst[(((size_t)b * C + c) * H + y) * prm.dst_pitch + prm.Wt + xi] = acc;
                if (prm.cls_max != nullptr) {
                    const unsigned m = __reduce_max_sync(0xffffffffu, valid ? ordered_from_float(acc) : 0u);
                    if (lane == 0 && m != 0u) atomicMax(prm.cls_max + (size_t)b * C + c, m);
                }
            }
        }
        __syncthreads();
    }
}

// Copy [planes,H,W] -> [planes,H,Wp] (Wp % 4 == 0) so that TMA's 16-byte stride rule holds.
__global__ void repack_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Wp, size_t rows) {
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const float* __restrict__ s = src + row * W;
        float* __restrict__ d = dst + row * Wp;
        for (int x = threadIdx.x; x < W; x += blockDim.x) d[x] = s[x];
    }
    (void)H;
}

std::atomic<long long*> g_timeline{nullptr};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap(CUtensorMap* map, const float* base, int planes, int H, int W, int pitch, int win_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return set_error(PAMR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * 4 * (cuuint64_t)H};
    cuuint32_t box[3] = {(cuuint32_t)WIN_W, (cuuint32_t)win_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PAMR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PAMR_OK;
}

// Row strip placement: when the tile count leaves the last wave partly empty, the CTAs without a
// tile in that wave take the row-strip items (one warp per item, ~2.5 us each, TAIL_ITEMS_MAX in a
// row still fit inside one tile time); otherwise the strip is a launch of its own.
constexpr int TAIL_ITEMS_MAX = 8;
inline bool row_strip_in_tail(long long items, int ntiles, int grid) {
    static const bool off = getenv("PAMR_B200_NO_TAIL") != nullptr;
    if (off || grid <= 0 || ntiles <= grid || ntiles % grid == 0) return false;
    const long long warps = (long long)(grid - ntiles % grid) * NWC;
    return (items + warps - 1) / warps <= TAIL_ITEMS_MAX;
}

template <int R>
int launch_one(const float* aff, const AffTiling& tiling, const float* src, int src_pitch, float* dst, int dst_pitch,
               int B, int C, int H, int W, int Wt, int Ht, unsigned* cls_max, int sm_count, cudaStream_t s,
               SideLane* lane) {
    using C_ = Cfg<R>;
    // function attributes are per device: set once per (kernel, device)
    static std::atomic<int> attr_set[64];
    int dev = 0;
    PAMR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
        PAMR_CUDA_TRY(cudaFuncSetAttribute(propagate_sm100_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C_::SMEM_BYTES));
        if (dev < 64) attr_set[dev].store(1, std::memory_order_release);
    }
    alignas(64) CUtensorMap tmap;
    int rc = make_tmap(&tmap, src, B * C, H, W, src_pitch, C_::WIN_H);
    if (rc != PAMR_OK) return rc;
    Params p;
    p.aff = aff; p.dst = dst; p.cls_max = cls_max; p.dst_pitch = dst_pitch;
    p.dbg = g_timeline.load(std::memory_order_relaxed);
    static const int knob_cta = getenv("PAMR_B200_STAGGER_CTA") ? atoi(getenv("PAMR_B200_STAGGER_CTA")) : 0;
    // groups 1 and 2 start their first pass of a tile 0.5 / 1.0 us after group 0: keeps the three groups out of
    // phase, so that one group's wait -> release -> store latency overlaps the others' LDS-bound passes
    // (320x320 B=16 forward: 2.541 ms without, 2.522 / 2.511 / 2.517 ms with 200 / 500 / 1000 ns)
    static const int knob_grp = getenv("PAMR_B200_STAGGER_GRP") ? atoi(getenv("PAMR_B200_STAGGER_GRP")) : 500;
    p.stagger_cta_ns = knob_cta; p.stagger_grp_ns = knob_grp;
    // default: prefetch the next tile's weights when the third-last class is issued (PAMR_B200_PF_CLASS overrides)
    static const int knob_pf = getenv("PAMR_B200_PF_CLASS") ? atoi(getenv("PAMR_B200_PF_CLASS")) : -2;
    p.pf_class = (knob_pf == -2) ? (C >= 3 ? C - 3 : 0) : (knob_pf < C ? knob_pf : -1);
    p.dbg_cta = getenv("PAMR_B200_DBG_CTA") ? atoi(getenv("PAMR_B200_DBG_CTA")) : 0;
    p.exp_flags = getenv("PAMR_B200_EXPERIMENT") ? atoi(getenv("PAMR_B200_EXPERIMENT")) : 0;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.tiles_x = (Wt + TX - 1) / TX;
    p.tiles_y = (Ht + C_::TY - 1) / C_::TY;
    p.tiles_x_aff = tiling.tiles_x;
    p.tiles_y_aff = tiling.tiles_y;
    p.src = src; p.src_pitch = src_pitch; p.Wt = Wt; p.Ht = Ht;
    p.strip_items = 0;
    p.tail_cta0 = -1;
    p.ntiles = p.tiles_x * p.tiles_y * B;
    const int grid = p.ntiles < sm_count ? p.ntiles : sm_count;
    const bool skip_strips = getenv("PAMR_B200_EXPERIMENT") && (atoi(getenv("PAMR_B200_EXPERIMENT")) & 16);
    const bool col_strip = Wt < W && !skip_strips;
    if (col_strip && lane != nullptr && lane->strip_pending) {  // this iteration reads what strip(t-1) wrote
        PAMR_CUDA_TRY(cudaStreamWaitEvent(s, lane->strip_done, 0));
        lane->strip_pending = false;
    }
    const long long row_items = (long long)B * C * (H - Ht) * ((W + 31) / 32);
    if (row_items > 0x7fffffffLL) return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: row strip too large");
    if (Ht < H && !skip_strips && row_strip_in_tail(row_items, p.ntiles, grid)) {
        p.strip_items = (int)row_items;  // the tile kernel's short CTAs do the row strip
        p.tail_cta0 = p.ntiles % grid;
    } else if (Ht < H && !skip_strips) {  // row strip y in [Ht,H), all columns, as a launch of its own
        Params pr = p;
        pr.Wt = W;  // strip_item: no column part, the row part spans [0,W)
        pr.strip_items = (int)row_items;
        const long long items = row_items;
        const int blocks = (int)((items + 3) / 4);
        strip_rows_kernel<R><<<blocks < 8 * sm_count ? blocks : 8 * sm_count, 128, 0, s>>>(pr);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
    }
    // The column strip (x in [Wt,W), all rows) runs on the side lane concurrently with the tile kernel:
    // both only read iteration t-1.  Measured on B200 its CTAs mostly get placed as tile CTAs retire in
    // the last wave (side-stream CTAs are scheduled only sluggishly next to resident persistent CTAs,
    // tools/coresidency.py), which still hides about half of it: 321x321 B=16 forward 3.085 ms vs
    // 3.204 ms with the strip serialised.  Ordering: the main stream waited for strip(t-1) above
    // (iteration t reads and overwrites what that strip wrote / read); strip(t) waits for tiles(t-1).
    Params pm = p;
    propagate_sm100_kernel<R><<<grid, NTHREADS, C_::SMEM_BYTES, s>>>(tmap, pm);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    if (col_strip) {
        const size_t smem = sizeof(float) * ((size_t)SC_CG * SC_WIN_H * (SC_WIN_W + 1) + (size_t)(W - Wt) * 48 * SC_ROWS);
        static std::atomic<int> sc_attr[64];
        if (dev >= 64 || sc_attr[dev].load(std::memory_order_acquire) == 0) {
            PAMR_CUDA_TRY(cudaFuncSetAttribute(strip_cols_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            if (dev < 64) sc_attr[dev].store(1, std::memory_order_release);
        }
        dim3 sgrid((H + SC_ROWS - 1) / SC_ROWS, B);
        cudaStream_t ss = s;
        if (lane != nullptr) {
            ss = lane->stream;
            PAMR_CUDA_TRY(cudaStreamWaitEvent(ss, lane->tiles_done, 0));  // tiles(t-1) (or the prologue) finished
        }
        strip_cols_kernel<R><<<sgrid, SC_CG * 32, smem, ss>>>(p);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
        if (lane != nullptr) {
            PAMR_CUDA_TRY(cudaEventRecord(lane->strip_done, ss));
            lane->strip_pending = true;
            PAMR_CUDA_TRY(cudaEventRecord(lane->tiles_done, s));  // tiles(t), for strip(t+1)
        }
    }
    return PAMR_OK;
}

}  // namespace

// Debug hook (not part of the public ABI): device buffer of 2 x 4096 x 2 int64 that CTA 0 of every
// subsequent tuned launch fills with {clock64, event code} pairs; nullptr switches it off.
extern "C" void pamr_debug_set_timeline(long long* dev_buf) { g_timeline.store(dev_buf); }

// Tiling of the tuned kernel (R rows per thread, tile = 32 x 4R, extent [0,Wt) x [0,Ht) covered by
// tiles) or R == 0 when the tuned kernel does not apply.  Remainders of at most 8 rows / columns have
// an alternative to a padded tile row / partial tile column:
//   rows:    the row strip -- inside the tile kernel's last wave when that wave has idle CTAs
//            (row_strip_in_tail), else a launch of its own;
//   columns: the column-strip launch.
// All combinations with R in {8,9,10} are priced with a time model fitted to measurements on B200
// (profiles/r01_strip_times.txt): waves of tiles over the SMs at ~2.8 us per row-per-thread, row-strip
// launch ~ 8 + 1.3 us per 1000 items, column-strip launch ~ (12 + 2.6 wc) * max(1, CTAs/90)^0.8 us minus what
// hides in the last wave -- and the cheapest wins.  E.g. 321 x 321, B=16: R=10, 10 x 8 tiles, row 320 in the
// tail of the same launch, column 320 as a strip launch on the side lane.  All paths add the 48 products of a pixel in the same order, so the result
// does not depend on the tiling (nor, therefore, on how a batch is sharded).
// Experiment overrides: PAMR_B200_ROWS=8|9|10, PAMR_B200_STRIP_MAX=<pixels> (0 disables strips),
// PAMR_B200_FORCE_STRIPS (bit 0 rows, bit 1 columns: take the strip whenever it qualifies),
// PAMR_B200_NO_TAIL=1.
AffTiling tuned_tiling(int B, int H, int W, const Dilations& dil) {
    static const int want[6] = {1, 2, 4, 8, 12, 24};
    // debugging / A-B aid: PAMR_B200_FORCE_GENERIC=1 routes everything to the generic CUDA kernel
    static const bool force_generic = []() {
        const char* e = getenv("PAMR_B200_FORCE_GENERIC");
        return e != nullptr && e[0] == '1';
    }();
    static const int force_rows = getenv("PAMR_B200_ROWS") ? atoi(getenv("PAMR_B200_ROWS")) : 0;
    static const int strip_max = getenv("PAMR_B200_STRIP_MAX") ? atoi(getenv("PAMR_B200_STRIP_MAX")) : 8;
    static const int force_strips = getenv("PAMR_B200_FORCE_STRIPS") ? atoi(getenv("PAMR_B200_FORCE_STRIPS")) : 0;
    AffTiling t{0, 0, 0, 0, 0};
    if (force_generic || dil.nd != 6 || W < TX || H < 8) return t;
    for (int i = 0; i < 6; ++i)
        if (dil.d[i] != want[i]) return t;
    const int sms = 148, C = 21;  // the model prices the reference's 21 classes
    double best_cost = 1e30;
    for (int r = 8; r <= 10; ++r) {
        if (force_rows >= 8 && force_rows <= 10 && r != force_rows) continue;
        const int ty = NW * r;
        // R = 9 (odd: scalar tail row, spills at 128 registers) measured ~2x slower per tile: kept for
        // experiments, effectively never chosen
        const double tile_us = 2.8 * r * (r == 10 ? 1.00 : r == 9 ? 2.2 : 1.07);
        const int hrem = H % ty, wrem = W % TX;
        const bool rs_ok = hrem != 0 && hrem <= strip_max && H > ty;
        const bool cs_ok = wrem != 0 && wrem <= strip_max && W > TX;
        for (int rs = 0; rs < 2; ++rs) {      // rs: row remainder as a strip
            if (rs ? !rs_ok : (rs_ok && (force_strips & 1))) continue;
            const int ht = rs ? H - hrem : H;
            for (int cs = 0; cs < 2; ++cs) {  // cs: column remainder as a strip launch
                if (cs ? !cs_ok : (cs_ok && (force_strips & 2))) continue;
                const int wt = cs ? W - wrem : W;
                const long long ntiles = (long long)B * ((wt + TX - 1) / TX) * ((ht + ty - 1) / ty);
                if (ntiles > 0x7fffffffLL) continue;
                double cost = (double)((ntiles + sms - 1) / sms) * tile_us;
                if (rs) {
                    const long long items = (long long)B * C * hrem * ((W + 31) / 32);
                    const int grid = ntiles < sms ? (int)ntiles : sms;
                    if (!row_strip_in_tail(items, (int)ntiles, grid)) cost += 8.0 + 1.3e-3 * (double)items;
                }
                if (cs) {
                    const double ctas = (double)B * ((H + SC_ROWS - 1) / SC_ROWS);
                    // stand-alone duration, less what hides in the idle part of the tile kernel's last wave
                    // (the strip runs on the side lane and its CTAs land on SMs whose tile CTA has retired)
                    const double alone = (12.0 + 2.6 * wrem) * pow(ctas > 90.0 ? ctas / 90.0 : 1.0, 0.8);
                    const double waves = (double)((ntiles + sms - 1) / sms);
                    const double idle_us = (waves * sms - (double)ntiles) / sms * tile_us;
                    cost += fmax(0.25 * alone, alone - 0.85 * idle_us);
                }
                if (cost < best_cost) {
                    best_cost = cost;
                    t.R = r; t.Wt = wt; t.Ht = ht;
                }
            }
        }
    }
    t.tiles_x = (W + TX - 1) / TX;
    t.tiles_y = (H + NW * t.R - 1) / (NW * t.R);
    return t;
}

int launch_repack(const float* src, float* dst, int planes, int H, int W, int Wp, cudaStream_t s) {
    const size_t rows = (size_t)planes * H;
    const unsigned grid = (unsigned)(rows < 148 * 16 ? rows : 148 * 16);
    repack_kernel<<<grid, 128, 0, s>>>(src, dst, H, W, Wp, rows);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

// One propagation step src -> dst: the strip kernels (if the tiling has remainders) followed by the
// persistent tile kernel.  src must have a pitch that is a multiple of 4 floats and a 16-byte aligned base.
int launch_propagate_tuned(const float* aff_tiled, const AffTiling& tiling, const float* src, int src_pitch, float* dst,
                           int dst_pitch, int B, int C, int H, int W, unsigned* cls_max, int dev, cudaStream_t s,
                           SideLane* lane) {
    static int sm_counts[64] = {0};
    int sm_count = (dev >= 0 && dev < 64) ? sm_counts[dev] : 0;
    if (sm_count == 0) {
        PAMR_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_counts[dev] = sm_count;
    }
    if ((src_pitch & 3) != 0 || ((uintptr_t)src & 15) != 0)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: source pitch/base not 16-byte aligned");
    const int Wt = tiling.Wt, Ht = tiling.Ht;
    if (tiling.R == 8) return launch_one<8>(aff_tiled, tiling, src, src_pitch, dst, dst_pitch, B, C, H, W, Wt, Ht, cls_max, sm_count, s, lane);
    if (tiling.R == 9) return launch_one<9>(aff_tiled, tiling, src, src_pitch, dst, dst_pitch, B, C, H, W, Wt, Ht, cls_max, sm_count, s, lane);
    return launch_one<10>(aff_tiled, tiling, src, src_pitch, dst, dst_pitch, B, C, H, W, Wt, Ht, cls_max, sm_count, s, lane);
}

}  // namespace pamr
