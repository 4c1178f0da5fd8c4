// Tuned sm_100a propagation kernel for the standard dilation set [1,2,4,8,12,24]
// (reference models/mods/pamr.py:138-140 with core/config.py:92).
//
//   M'[b,c,y,x] = sum_{p<48} w[b,p,y,x] * M[b,c,clamp(y+dy_p),clamp(x+dx_p)]
//
// Design:
//  * persistent CTAs, one per SM; a CTA owns a 32 x (4*R) pixel tile (R = 8 or 10 rows per thread);
//  * 12 compute warps in 3 groups of 4: lane = x, each thread owns a vertical strip of R pixels, so that
//    one shared-memory load feeds up to 3 taps (register reuse along y).  Group g takes the class planes
//    c = g (mod 3);
//  * the mask planes between iterations live in a ROW-PAIR INTERLEAVED layout [plane][y/2][x][y&1]
//    (this library owns the ping-pong buffers), so that a thread's vertical strip is read with 64-bit
//    shared-memory loads: LDS.64 reaches the 128 B/clk/SM of the shared-memory pipe with 8 warps where
//    LDS.32 needs 16 and then tops out at ~92 % (profiles/r01_ubench_lds_ffma_l2.txt), the loaded pairs
//    are the aligned register pairs the packed FFMA2 wants, and the LSU issues half the instructions;
//  * the tile's 48 affinity weights per pixel (48*R words per thread) are parked in Tensor Memory,
//    1 TMEM lane per thread, shared by the three groups (warps w, w+4, w+8 address the same lane
//    quarter); every class pass re-reads them with tcgen05.ld, a data path that does not compete with
//    LDS.  The NEXT tile's weights are streamed in behind the last class passes of the current tile by
//    the copy engines alone, 32 TMEM columns ("unit", 16 KB for the 128 lanes) at a time and in the order
//    the passes consume them: a loader thread brings the unit from L2 into a shared-memory staging ring
//    with cp.async.bulk, and an issuer thread moves it on with tcgen05.cp (UTCCP: shared memory -> TMEM) as
//    soon as the twelve compute warps have read the unit for the last time (mbarrier `free`); tcgen05.commit
//    signals the mbarrier `filled` the first class pass of the next tile waits on, per unit.  No compute warp
//    touches the weights on their way in, no LSU instruction is spent on them, and there is no tile-wide barrier;
//  * a producer warp streams the class planes of the tile (+24 px halo, 80 x (4R+48) floats) through a
//    4-slot shared-memory ring with TMA (cp.async.bulk.tensor.3d over 64-bit elements = row pairs) +
//    mbarriers; TMA zero-fills outside the image; replicate padding (pamr.py:50) in x is a per-lane
//    clamped column offset, in y it is DATA: the planes carry 24 replicated rows above and below the image, written by
//    whoever writes row 0 / row H-1 (no patching of the window in shared memory);
//  * FP32 math as packed FFMA2 over adjacent rows; the per-(b,c) max for pseudo_gtmask is fused into
//    the last iteration (warp reduce + atomicMax);
//  * remainders (W = H = 321 -> one column, one row): a column strip of at most 2 columns is computed by the
//    tiles on the right image border from the window they hold anyway (one pixel per lane after every class
//    pass, weights of the tile's strip pixels in shared memory); a row strip of at most 8 rows is computed by
//    the CTAs that are idle in the kernel's last wave (or a small launch when there are none);
//  * iterations are separate launches chained with programmatic dependent launch.  The kernel is written over
//    "tile visits" (struct Visit) so that the same code also runs ALL iterations in one launch with per-tile
//    epochs instead of kernel boundaries (template parameter FUSED, -DPAMR_FUSED_ITERATIONS): that variant is
//    bit-identical and slower, so the product does not instantiate it.
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
constexpr int WIN_W = TX + 2 * HALO;  // 80 columns
// Pitch of the window in shared memory, in columns: 82, so that a row pair is 164 floats = 4 banks past a
// multiple of 32 and the column strip's lanes (same column, different rows) fall on different banks (with 80
// every row pair starts on the same bank).  TMA simply loads two more columns; nobody reads them.
#ifndef PAMR_WIN_PAD
#define PAMR_WIN_PAD 2
#endif
constexpr int WIN_P = WIN_W + PAMR_WIN_PAD;
constexpr int ROWP = WIN_P * 2;       // floats per row PAIR of the window in shared memory
constexpr int NW = 4;                 // warps per compute group (= TMEM lane quarters)
#ifndef PAMR_NG
#define PAMR_NG 3
#endif
constexpr int NG = PAMR_NG;           // compute groups that share the tile's weights in TMEM (class c -> group c % NG)
// Classes per pass.  2: a group takes two class planes per pass and applies every weight it has fetched from Tensor
// Memory to both, which halves the Tensor-Memory read traffic per pixel-class -- tcgen05.ld and LDS share the path
// back into the register file, and that path (~205 B/clk/SM for this mix, profiles/r01_ubench_tmem.txt), not either
// pipe alone, is what bounds the one-class kernel.
#ifndef PAMR_CPP
#define PAMR_CPP 1
#endif
constexpr int CPP = PAMR_CPP;
constexpr int NWC = NG * NW;          // compute warps
#ifndef PAMR_NSLOT
#define PAMR_NSLOT 4
#endif
#ifndef PAMR_NSTG
#define PAMR_NSTG 3
#endif
#ifndef PAMR_PF_PLANES
#define PAMR_PF_PLANES 0
#endif
constexpr int NSLOT = PAMR_NSLOT;     // class-plane slots in the shared-memory ring
// mbarriers of the ring: sequence number n uses slot n % NSLOT but barrier pair n % NBAR.  Waits are by
// phase PARITY, which is only sound if a waiter can never be a whole phase ahead of the barrier.  With one
// barrier per slot and 4 slots, a group that has finished class k-3 tests the barrier of class k while the
// load of class k-4 (same slot, other group) may -- once in ~1e7 passes -- still be in flight; the parity
// test then reports the OLD phase as complete and the group computes on the wrong plane (round 1: 0.2-2.5 %
// of the forward calls had one wrong tile-class; 0 of 5500 with 8 barrier pairs).
constexpr int NBAR = 2 * NSLOT;
static_assert(NBAR % NSLOT == 0 && NBAR >= 2 * NSLOT, "barrier ring must cover at least two uses of every slot");
// Warp roles: 0..11 compute, 12 class-plane producer (TMA), 13 weight loader (cp.async.bulk into the staging
// ring), 14 and 15 weight issuers (tcgen05.cp staging -> TMEM; even / odd units).
constexpr int W_PRODUCER = NWC, W_LOADER = NWC + 1, W_ISSUER = NWC + 2;
constexpr int NTHREADS = (NWC + 4) * 32;
#ifndef PAMR_CHUNK_UNITS
#define PAMR_CHUNK_UNITS 2
#endif
constexpr int CHUNK_UNITS = PAMR_CHUNK_UNITS;  // units handed over together ("chunk": CHUNK_UNITS x 32 TMEM columns = CHUNK_UNITS x 16 KB; the last chunk of a tile may be short)
constexpr int NSTG = PAMR_NSTG;       // staging ring for the weights: NSTG chunks of 32 KB
constexpr int PF_PLANES = PAMR_PF_PLANES;  // class planes pulled into L2 this many classes ahead of their TMA load (0: off)
constexpr int WB = 16;                // weights per tcgen05.ld batch of the centre column
#ifndef PAMR_UNIT
#define PAMR_UNIT 32
#endif
constexpr int UNIT = PAMR_UNIT;       // TMEM columns per fill unit (a multiple of the 16-column batch)
constexpr int UNIT_BYTES = UNIT * 128 * 4;  // 128 TMEM lanes
constexpr int MAX_CHUNKS = (480 / (CHUNK_UNITS * UNIT) + 2) / 2 * 2;  // >= chunks per tile (R = 10: 480 columns), even
constexpr int CTRL_BYTES = 1024;
constexpr int CS_MAX_W = 1;           // widest column strip the border tiles take on
constexpr int RS_MAX_H = 8;           // tallest row strip

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
// non-blocking poll
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
// Blocking wait: try_wait with a suspend-time hint, so that the waiting thread sleeps in hardware until the
// phase completes instead of re-issuing try_wait / branch pairs that take issue slots from the compute warps
// on its scheduler.  A wait that does not complete within ~1e6 retries (seconds) is a protocol bug: trap
// instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok == 0 && ++spins > (1u << 20)) __trap();
    } while (ok == 0);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
// tcgen05.wait::ld, with the batch's registers tied through the asm so that no consumer of the
// loaded values can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait_ld(float (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]));
}
#if PAMR_CPP == 2
// 32 / 8 / 4 consecutive columns (two-classes-per-pass kernel: one load per tap segment)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),
          "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),
          "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, float* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4p(uint32_t taddr, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
                 : "r"(taddr));
}
// tcgen05.wait::ld with a 32-register segment buffer tied through the asm
__device__ __forceinline__ void tmem_wait_ld32(float* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]),
                   "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]),
                   "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]), "+f"(r[30]), "+f"(r[31]));
}
#endif  // PAMR_CPP == 2
// plain bulk copy global -> shared memory, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// Shared-memory matrix descriptor for tcgen05.cp, no swizzle: 128 rows (= TMEM lanes) of 16-byte pieces;
// row r, piece k (4 columns each) sits at start + (r/8)*sbo + k*lbo + (r%8)*16 (measured: tools/utccp_test.cu).
__device__ __forceinline__ uint64_t utccp_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// 128 lanes x 8 columns (two 16-byte pieces per lane) from shared memory to TMEM, asynchronously
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
// the mbarrier receives one arrival once every tcgen05.cp issued by this thread so far has completed
__device__ __forceinline__ void utccp_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- shared-memory layout
template <int R>
struct Cfg {
    static_assert(R % 2 == 0, "row pairs: R must be even");
    static constexpr int TY = NW * R;
    static constexpr int WIN_H = TY + 2 * HALO;  // even
    static constexpr int SLOT_BYTES = WIN_P * WIN_H * 4;                      // what one TMA load brings
    static constexpr int SLOT_FLOATS = (SLOT_BYTES + 127) / 128 * 32;         // slot stride: TMA destinations are 128-byte aligned
    static constexpr size_t STAGE_OFF = (size_t)NSLOT * SLOT_FLOATS * 4;        // weight staging ring
    static constexpr size_t CSW_OFF = STAGE_OFF + (size_t)NSTG * CHUNK_UNITS * UNIT_BYTES;  // staging of the column-strip weights
    static constexpr int CS_LP = aff_cs_lanes(R), CS_TPL = aff_cs_taps(R);   // lanes per strip pixel, taps per lane
    static constexpr int CSW_BYTES = (int)aff_cs_block_floats(R) * 4;
    static constexpr int CS_COL0 = 48 * R;                                     // their TMEM columns [CS_COL0, CS_COL0 + CS_TPL)
    static_assert(CS_TPL % 4 == 0 && CS_TPL <= WB && CS_COL0 + WB <= 512, "column-strip weights in spare TMEM columns");
    static constexpr size_t CTRL_OFF = CSW_OFF + CSW_BYTES;
    static constexpr size_t SMEM_BYTES = CTRL_OFF + CTRL_BYTES;
    static_assert(STAGE_OFF % 128 == 0, "TMA / bulk destinations");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct Ctrl {  // lives in the last CTRL_BYTES of dynamic shared memory
    unsigned long long tma_bar[NBAR];           // class plane landed in its slot
    unsigned long long empty_bar[NBAR];         // class plane consumed
    unsigned long long filled_bar[MAX_CHUNKS];  // the tile's weights of that chunk are in TMEM (tcgen05.commit)
    unsigned long long free_bar[MAX_CHUNKS];    // every compute warp has read the chunk for the last time in this tile
    unsigned long long csw_staged_bar;          // column-strip weights of the next border tile have landed in shared memory
    unsigned long long csw_full_bar;            // ... are in TMEM (tcgen05.commit)
    unsigned long long csw_free_bar;            // ... and every compute warp is done with them
    unsigned long long staged_bar[2 * NSTG];    // the chunk's bytes have landed in the staging ring (chunk n: stage n % NSTG,
                                                // barrier n % (2 NSTG), so that a barrier's consecutive phases belong to ONE issuer)
    uint32_t tmem_base;
    int done_cnt[4];                            // fused iterations: compute warps that have stored the tile (by tile visit % 4)
};
static_assert(sizeof(Ctrl) <= CTRL_BYTES, "control block");

struct Params {
    const float* aff;  // affinity layout of pamr_common.cuh (tiles, column strip, row strip)
    const float* src;  // source mask, row-pair layout [B*C][Hp2][src_pitch][2] (the tiles read it through TMA, the row strip directly)
    float* dst;        // dst_pair: row-pair layout [B*C][Hp2][dst_pitch][2]; else standard [B*C][H][dst_pitch]
    unsigned* cls_max; // [B,C] or nullptr
    int src_pitch, dst_pitch, dst_pair, Hp2;
    int pf_class;      // class index at whose TMA issue the prefetch of the next tile's weights into L2 starts
    int B, C, H, W;
    int tiles_x, tiles_y, ntiles;  // tile grid (= the affinity layout's)
    int Wt, Ht;                    // the tiles cover [0,Wt) x [0,Ht)
    size_t cs_base, rs_base;       // column-strip / row-strip weights inside aff
    int strip_items;               // number of 32-pixel row-strip work items
    int tail_cta0;                 // row strip inside the tile kernel: CTAs >= tail_cta0 (one tile fewer than the
                                   // others) work through the strip_items row items after their last tile; -1: off
    // ---- fused iterations (tile_epoch != nullptr): ONE launch runs all `iters` propagation steps.  Iteration it reads
    // buf[it & 1] (tensor map it & 1) and writes buf[(it + 1) & 1], the last one writes `out` / `out_cls_max`.  There is no
    // grid-wide barrier: a tile of iteration it starts as soon as its 3 x 3 tile neighbourhood (and, for a bottom-row
    // tile, the image's row strip) of iteration it - 1 is complete, tracked in tile_epoch / strip_count.
    int iters;
    const float* buf[2];
    float* out;
    unsigned* out_cls_max;
    int out_pitch;
    int* tile_epoch;               // [ntiles] iterations completed per tile (zeroed before the launch)
    int* strip_count;              // [B] row-strip items completed per image, over all iterations
    int strip_items_per_image;
#ifdef PAMR_EXPERIMENTS
    long long* dbg;                // timeline buffer [5 streams][4096][2] = {clock64, code} or nullptr (tools/timeline.py)
    int dbg_cta;
#endif
};

// In-kernel timeline, compiled only into experiment builds (-DPAMR_EXPERIMENTS, tools/build_variant.sh).
#ifdef PAMR_EXPERIMENTS
std::atomic<long long*> g_timeline{nullptr};
std::atomic<int> g_timeline_cta{0}, g_timeline_skip{0};
#define PAMR_EV(stream, cond, code)                                                          \
    do {                                                                                     \
        if (prm.dbg != nullptr && (int)blockIdx.x == prm.dbg_cta && (cond) && ev_n < 4096) { \
            prm.dbg[((stream) * 4096 + ev_n) * 2] = clock64();                               \
            prm.dbg[((stream) * 4096 + ev_n) * 2 + 1] = (code);                              \
            ++ev_n;                                                                          \
        }                                                                                    \
    } while (0)
#else
#define PAMR_EV(stream, cond, code) do { } while (0)
#endif

// element (y, x) of a window / an unpadded plane in the row-pair layout
__host__ __device__ __forceinline__ size_t pair_index(int pitch, int y, int x) {
    return ((size_t)(y >> 1) * pitch + x) * 2 + (y & 1);
}
// Planes of the ping-pong buffers carry PADP replicated row pairs above row 0 and below row H-1 (replicate padding in
// y, pamr.py:50, as DATA: whoever writes row 0 / row H-1 also writes the 24 rows beyond it), so that the TMA window of
// a tile on the top / bottom image border needs no patching.  Hp2 in Params counts the padded row pairs.
constexpr int PADP = HALO / 2;
__host__ __device__ __forceinline__ int padded_pairs(int H) { return (H + 1) / 2 + 2 * PADP; }
__host__ __device__ __forceinline__ size_t plane_index(int pitch, int y, int x) {  // y in [-HALO, H + HALO)
    return ((size_t)((y + HALO) >> 1) * pitch + x) * 2 + ((y + HALO) & 1);
}
__device__ __forceinline__ size_t src_plane_stride(const Params& p) { return (size_t)p.Hp2 * p.src_pitch * 2; }
// what one iteration reads and writes
struct IterIO {
    const float* src;
    float* dst;
    unsigned* cls_max;
    int dst_pitch, dst_pair;
};
template <bool FUSED>
__device__ __forceinline__ IterIO iter_io(const Params& p, int it) {
    IterIO io;
    if (!FUSED) {  // plain launch: one iteration
        io.src = p.src; io.dst = p.dst; io.cls_max = p.cls_max; io.dst_pitch = p.dst_pitch; io.dst_pair = p.dst_pair;
    } else {
        const bool last = it == p.iters - 1;
        io.src = p.buf[it & 1];
        io.dst = last ? p.out : const_cast<float*>(p.buf[(it + 1) & 1]);
        io.cls_max = last ? p.out_cls_max : nullptr;
        io.dst_pitch = last ? p.out_pitch : p.src_pitch;
        io.dst_pair = last ? 0 : 1;
    }
    return io;
}
__device__ __forceinline__ float* dst_pixel(const Params& p, const IterIO& io, int plane, int y, int x) {
    return io.dst_pair ? io.dst + (size_t)plane * p.Hp2 * io.dst_pitch * 2 + plane_index(io.dst_pitch, y, x)
                       : io.dst + ((size_t)plane * p.H + y) * io.dst_pitch + x;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spins until *p >= target (another CTA's release); a wait that never ends is a protocol bug: trap instead of hanging
__device__ __forceinline__ void wait_at_least(const int* p, int target) {
    unsigned spins = 0;
    while (ld_acquire(p) < target)
        if (++spins > (1u << 26)) __trap();
}
// rows beyond the image that repeat row y (y == 0: the 24 rows above; y == H-1: the rows below, including the second
// half of the last row pair when H is odd); none for any other row
__device__ __forceinline__ void store_replicas(const Params& p, const IterIO& io, int plane, int y, int x, float v) {
    if (!io.dst_pair) return;
    if (y == 0)
        for (int r = -HALO; r < 0; ++r) *dst_pixel(p, io, plane, r, x) = v;
    if (y == p.H - 1)
        for (int r = p.H; r < 2 * (((p.H + 1) >> 1)) + HALO; ++r) *dst_pixel(p, io, plane, r, x) = v;
}

// ---------------------------------------------------------------- TMEM weight layout
// Per thread (= TMEM lane) the 48*R weights of its R pixels sit in consumption order, which is the dense order
// column = s*R + i (s = tap sequence index of pamr_common.cuh: the 12 taps of the centre column b = 0 first,
// then the side groups g = 6*bi + id with a = -1,0,+1; i = row within the thread's strip).  The passes read
// them as a stream of 16-column batches; a fill unit is two batches (32 columns x 128 lanes = 16 KB), which is
// one contiguous block of the tile-major affinity layout (pamr_common.cuh) and four tcgen05.cp.128x256b.
template <int R>
struct TmemLayout {
    static constexpr int NCOLS = 48 * R;          // 480 (R = 10) or 384 (R = 8) of the 512 columns
    static constexpr int NB = NCOLS / WB;         // batches per pass
    static constexpr int NU = NCOLS / UNIT;       // fill units per tile
    static constexpr int NCH = (NU + CHUNK_UNITS - 1) / CHUNK_UNITS;  // chunks per tile
    static constexpr int BPC = CHUNK_UNITS * UNIT / WB;               // batches per (full) chunk
    static_assert(NCOLS % UNIT == 0 && NCOLS <= 512, "TMEM columns");
    static_assert(NCH <= MAX_CHUNKS, "chunks");
};

// ---------------------------------------------------------------- compute body
// Two FMAs on adjacent rows as one packed FFMA2 (fma.rn.f32x2).  A scalar FFMA whose three source
// registers are all distinct issues only every ~1.8 cycles per SM sub-partition (tools/ubench3.cu);
// the packed form retires two FMAs per ~2.4 cycles.  The mov.b64 packs are free when ptxas allocates the
// operands as aligned register pairs (TMEM batches, the LDS.64 row pairs and the accumulators all are).
__device__ __forceinline__ void fma2(float& a0, float& a1, float w0, float w1, float v0, float v1) {
    unsigned long long A, W2, V2;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(W2) : "f"(w0), "f"(w1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(V2) : "f"(v0), "f"(v1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(W2), "l"(V2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(A));
}
__device__ __forceinline__ void lds_pair(const float* p, float& lo, float& hi) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    lo = t.x;
    hi = t.y;
}
// every group has read this unit of the weights for the last time in this tile
__device__ __forceinline__ void signal_free(uint32_t bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// One pass over the 48 taps for one class plane resident in the ring.
// sp points at this thread's pixel pair row (rows R*wq, R*wq+1; column lane) inside the slot, i.e.
// slot + ((R*wq + HALO)/2)*ROWP + (lane + HALO)*2; neighbours are immediate offsets.
// xg = this lane's image column, W = image width: the side columns are addressed with a per-lane offset
// clamp(xg +- d, 0, W-1) - xg, which implements replicate padding in x for free, so only rows above /
// below the image are ever an issue (they are replicated rows of the padded planes).
// filled_bar != 0: first pass of the tile, wait per unit for the fill warp (phase parity par);
// free_bar != 0: last pass of the tile, release every unit after its last read.
#ifdef PAMR_PASS_SPECIALISED
template <int R, bool FIRST, bool LAST>
#else
template <int R>
#endif
__device__ __forceinline__ void compute_pass(const float* __restrict__ sp, uint32_t tbase, float (&acc)[R], int xg, int W,
                                             uint32_t filled_bar_, uint32_t free_bar_, uint32_t par, int lane) {
#ifdef PAMR_PASS_SPECIALISED
    // first / last pass of a tile as separate instantiations: the five middle passes carry no hand-over branches
    const uint32_t filled_bar = FIRST ? filled_bar_ : 0u, free_bar = LAST ? free_bar_ : 0u;
#else
    const uint32_t filled_bar = filled_bar_, free_bar = free_bar_;
#endif
    using L = TmemLayout<R>;
    // The weights stream through two 16-register buffers in consumption order: batch b (TMEM columns
    // [16b, 16b+16)) lives in wb[b & 1] and belongs to chunk b / BPC.  All indices are compile-time.
    float wb[2][WB];
    // entering batch b (its load was issued one batch earlier): wait for it, release its chunk if this was
    // the chunk's last batch, and issue the load of batch b + 1 (first pass of a tile: once its chunk is filled)
#define PAMR_ENTER_BATCH(b)                                                                      \
    do {                                                                                         \
        tmem_wait_ld(wb[(b) & 1]);                                                               \
        if (free_bar && (((b) % L::BPC) == L::BPC - 1 || (b) == L::NB - 1))                      \
            signal_free(free_bar + 8 * ((b) / L::BPC), lane);                                    \
        if ((b) + 1 < L::NB) {                                                                   \
            if (filled_bar && (((b) + 1) % L::BPC) == 0) {                                       \
                mbar_wait(filled_bar + 8 * (((b) + 1) / L::BPC), par);                           \
                tc_fence_after();                                                                \
            }                                                                                    \
            tmem_ld16(tbase + ((b) + 1) * WB, wb[((b) + 1) & 1]);                                \
        }                                                                                        \
    } while (0)
    if (filled_bar) {
        mbar_wait(filled_bar, par);
        tc_fence_after();
    }
    tmem_ld16(tbase, wb[0]);
    int bcur = -1;  // weight batch currently held (compile-time after full unrolling)

    // ---- centre column (b = 0): rows y+-d for all dilations share one register strip; every row pair is
    //      loaded once, right before the first dilation that needs it (short live ranges: the strip of
    //      all 54 rows would not fit the register budget next to the weights)
    {
        float v[R + 2 * HALO];
#pragma unroll
        for (int id = 0; id < 6; ++id) {
            const int d = dil_of(id);
#pragma unroll
            for (int r = -HALO; r < R + HALO; r += 2) {
                bool need = false, had = false;  // needed by this dilation / already loaded for a smaller one
#pragma unroll
                for (int rr = r; rr < r + 2; ++rr) {
                    need = need || (rr >= -d && rr < R - d) || (rr >= d && rr < R + d);
#pragma unroll
                    for (int jd = 0; jd < 6; ++jd) {
                        const int e = dil_of(jd);
                        if (jd < id) had = had || (rr >= -e && rr < R - e) || (rr >= e && rr < R + e);
                    }
                }
                if (need && !had) lds_pair(sp + (r / 2) * ROWP, v[r + HALO], v[r + 1 + HALO]);
            }
#pragma unroll
            for (int a = -1; a <= 1; a += 2) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const bool pair = (d % 2 == 0);  // rows (i, i+1) as one FFMA2
                    if (pair && (i % 2 == 1)) continue;  // odd row handled with its even partner
                    const int q = (2 * id + (a > 0 ? 1 : 0)) * R + i;
                    if (q / WB != bcur) {
                        bcur = q / WB;
                        PAMR_ENTER_BATCH(bcur);
                    }
                    if (pair) {
                        fma2(acc[i], acc[i + 1], wb[bcur & 1][q % WB], wb[bcur & 1][q % WB + 1], v[i + a * d + HALO],
                             v[i + 1 + a * d + HALO]);
                    } else {
                        acc[i] = fmaf(wb[bcur & 1][q % WB], v[i + a * d + HALO], acc[i]);
                    }
                }
            }
        }
    }
    // ---- side columns: bi = 0 (b = -1), bi = 1 (b = +1)
#pragma unroll
    for (int bi = 0; bi < 2; ++bi) {
        const int sgn = bi * 2 - 1;
#pragma unroll
        for (int id = 0; id < 6; ++id) {
            const int d = dil_of(id);
            float v[R + 2 * HALO];
            const int coff = (min(max(xg + sgn * d, 0), W - 1) - xg) * 2;
#pragma unroll
            for (int r = -HALO; r < R + HALO; r += 2) {
                bool need = false;
#pragma unroll
                for (int rr = r; rr < r + 2; ++rr)
                    need = need || (rr >= -d && rr < R + d && ((rr < R - d) || (rr >= 0 && rr < R) || (rr >= d)));
                if (need) lds_pair(sp + (r / 2) * ROWP + coff, v[r + HALO], v[r + 1 + HALO]);
            }
#pragma unroll
            for (int a = -1; a <= 1; ++a) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const bool pair = ((a * d) % 2 == 0);  // rows (i, i+1)
                    if (pair && (i % 2 == 1)) continue;
                    const int q = (12 + 3 * (bi * 6 + id) + (a + 1)) * R + i;
                    if (q / WB != bcur) {
                        bcur = q / WB;
                        PAMR_ENTER_BATCH(bcur);
                    }
                    if (pair) {
                        fma2(acc[i], acc[i + 1], wb[bcur & 1][q % WB], wb[bcur & 1][q % WB + 1], v[i + a * d + HALO],
                             v[i + 1 + a * d + HALO]);
                    } else {
                        acc[i] = fmaf(wb[bcur & 1][q % WB], v[i + a * d + HALO], acc[i]);
                    }
                }
            }
        }
    }
#undef PAMR_ENTER_BATCH
}

#if PAMR_CPP == 2
// ---- two classes per pass (CPP == 2) ----
// The 48 taps are walked in 18 SEGMENTS of weights that share one register strip of class-plane rows: 6 for the
// centre column (dilation id: taps dy = -d, +d; 2R TMEM columns) and 12 for the side columns ((bi, id): taps
// dy = -d, 0, +d at dx = -+d; 3R columns).  A segment's weights are fetched from Tensor Memory ONCE (double
// buffered: the next segment's load is in flight while this one is used) and applied first to plane 0, then to
// plane 1; the strip registers are reused between the planes.  Tap-sequence order is unchanged, so every pixel's 48
// products are still added in the same order as on every other path.
// (Unlike the one-class pass the centre column's rows are not shared between dilations: 36.4 instead of 31.8 LDS
// words per pixel-class, against 24 + instead of 48 Tensor-Memory words.)
template <int R>
struct Seg {
    // segment index s: 0..5 centre (id = s), 6..17 side (g = s - 6 = 6 bi + id)
    __host__ __device__ static constexpr int col0(int s) { return s < 6 ? 2 * s * R : (12 + 3 * (s - 6)) * R; }
    __host__ __device__ static constexpr int ncol(int s) { return s < 6 ? 2 * R : 3 * R; }
};
constexpr int CHUNK_COLS = CHUNK_UNITS * UNIT;

// loads the weights of segment s into w[0 .. ncol): one or two power-of-two loads; reading a few columns past the
// segment is harmless (they belong to the next segment of the same tile, or to spare columns)
template <int R, int S>
__device__ __forceinline__ void seg_load(uint32_t tbase, float* w) {
    constexpr int c0 = Seg<R>::col0(S), n = Seg<R>::ncol(S);
    if constexpr (n <= 16) {
        tmem_ld16p(tbase + c0, w);
    } else if constexpr (n <= 20) {
        tmem_ld16p(tbase + c0, w);
        tmem_ld4p(tbase + c0 + 16, w + 16);
    } else if constexpr (n <= 24) {
        tmem_ld16p(tbase + c0, w);
        tmem_ld8p(tbase + c0 + 16, w + 16);
    } else {
        static_assert(n <= 32 && c0 + 32 <= 512, "segment");
        tmem_ld32(tbase + c0, w);
    }
}

// FMAs of one centre segment (dilation id) for one plane: rows y -+ d of the thread's own column
template <int R, int ID>
__device__ __forceinline__ void seg_centre_plane(const float* __restrict__ sp, const float* w, float (&acc)[R]) {
    constexpr int d = dil_of(ID);
    float v[R + 2 * HALO];
#pragma unroll
    for (int r = -HALO; r < R + HALO; r += 2) {
        bool need = false;
#pragma unroll
        for (int rr = r; rr < r + 2; ++rr) need = need || (rr >= -d && rr < R - d) || (rr >= d && rr < R + d);
        if (need) lds_pair(sp + (r / 2) * ROWP, v[r + HALO], v[r + 1 + HALO]);
    }
#pragma unroll
    for (int a = -1; a <= 1; a += 2) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const bool pair = (d % 2 == 0);
            if (pair && (i % 2 == 1)) continue;
            const int q = (a > 0 ? R : 0) + i;
            if (pair) fma2(acc[i], acc[i + 1], w[q], w[q + 1], v[i + a * d + HALO], v[i + 1 + a * d + HALO]);
            else acc[i] = fmaf(w[q], v[i + a * d + HALO], acc[i]);
        }
    }
}
// FMAs of one side segment (bi, id) for one plane
template <int R, int BI, int ID>
__device__ __forceinline__ void seg_side_plane(const float* __restrict__ sp, const float* w, float (&acc)[R], int xg, int W) {
    constexpr int d = dil_of(ID), sgn = BI * 2 - 1;
    float v[R + 2 * HALO];
    const int coff = (min(max(xg + sgn * d, 0), W - 1) - xg) * 2;
#pragma unroll
    for (int r = -HALO; r < R + HALO; r += 2) {
        bool need = false;
#pragma unroll
        for (int rr = r; rr < r + 2; ++rr)
            need = need || (rr >= -d && rr < R + d && ((rr < R - d) || (rr >= 0 && rr < R) || (rr >= d)));
        if (need) lds_pair(sp + (r / 2) * ROWP + coff, v[r + HALO], v[r + 1 + HALO]);
    }
#pragma unroll
    for (int a = -1; a <= 1; ++a) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const bool pair = ((a * d) % 2 == 0);
            if (pair && (i % 2 == 1)) continue;
            const int q = (a + 1) * R + i;
            if (pair) fma2(acc[i], acc[i + 1], w[q], w[q + 1], v[i + a * d + HALO], v[i + 1 + a * d + HALO]);
            else acc[i] = fmaf(w[q], v[i + a * d + HALO], acc[i]);
        }
    }
}

// the same for two planes at once: both strips are loaded before the first FMA and the FMAs of the two planes
// alternate (ten independent accumulator pairs instead of five: the kernel runs only two warps per scheduler)
template <int R, int ID>
__device__ __forceinline__ void seg_centre_both(const float* __restrict__ sp0, const float* __restrict__ sp1, const float* w,
                                                float (&acc0)[R], float (&acc1)[R]) {
    constexpr int d = dil_of(ID);
    float v0[R + 2 * HALO], v1[R + 2 * HALO];
#pragma unroll
    for (int r = -HALO; r < R + HALO; r += 2) {
        bool need = false;
#pragma unroll
        for (int rr = r; rr < r + 2; ++rr) need = need || (rr >= -d && rr < R - d) || (rr >= d && rr < R + d);
        if (need) {
            lds_pair(sp0 + (r / 2) * ROWP, v0[r + HALO], v0[r + 1 + HALO]);
            lds_pair(sp1 + (r / 2) * ROWP, v1[r + HALO], v1[r + 1 + HALO]);
        }
    }
#pragma unroll
    for (int a = -1; a <= 1; a += 2) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const bool pair = (d % 2 == 0);
            if (pair && (i % 2 == 1)) continue;
            const int q = (a > 0 ? R : 0) + i;
            if (pair) {
                fma2(acc0[i], acc0[i + 1], w[q], w[q + 1], v0[i + a * d + HALO], v0[i + 1 + a * d + HALO]);
                fma2(acc1[i], acc1[i + 1], w[q], w[q + 1], v1[i + a * d + HALO], v1[i + 1 + a * d + HALO]);
            } else {
                acc0[i] = fmaf(w[q], v0[i + a * d + HALO], acc0[i]);
                acc1[i] = fmaf(w[q], v1[i + a * d + HALO], acc1[i]);
            }
        }
    }
}
template <int R, int BI, int ID>
__device__ __forceinline__ void seg_side_both(const float* __restrict__ sp0, const float* __restrict__ sp1, const float* w,
                                              float (&acc0)[R], float (&acc1)[R], int xg, int W) {
    constexpr int d = dil_of(ID), sgn = BI * 2 - 1;
    float v0[R + 2 * HALO], v1[R + 2 * HALO];
    const int coff = (min(max(xg + sgn * d, 0), W - 1) - xg) * 2;
#pragma unroll
    for (int r = -HALO; r < R + HALO; r += 2) {
        bool need = false;
#pragma unroll
        for (int rr = r; rr < r + 2; ++rr)
            need = need || (rr >= -d && rr < R + d && ((rr < R - d) || (rr >= 0 && rr < R) || (rr >= d)));
        if (need) {
            lds_pair(sp0 + (r / 2) * ROWP + coff, v0[r + HALO], v0[r + 1 + HALO]);
            lds_pair(sp1 + (r / 2) * ROWP + coff, v1[r + HALO], v1[r + 1 + HALO]);
        }
    }
#pragma unroll
    for (int a = -1; a <= 1; ++a) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const bool pair = ((a * d) % 2 == 0);
            if (pair && (i % 2 == 1)) continue;
            const int q = (a + 1) * R + i;
            if (pair) {
                fma2(acc0[i], acc0[i + 1], w[q], w[q + 1], v0[i + a * d + HALO], v0[i + 1 + a * d + HALO]);
                fma2(acc1[i], acc1[i + 1], w[q], w[q + 1], v1[i + a * d + HALO], v1[i + 1 + a * d + HALO]);
            } else {
                acc0[i] = fmaf(w[q], v0[i + a * d + HALO], acc0[i]);
                acc1[i] = fmaf(w[q], v1[i + a * d + HALO], acc1[i]);
            }
        }
    }
}

// entering segment S (its load was issued during segment S-1): wait for it, release the chunks it completes (last
// pass of the tile), issue the load of segment S+1 (first pass of the tile: once its chunks are filled)
template <int R, int S>
__device__ __forceinline__ void seg_enter(uint32_t tbase, float* wcur, float* wnext, uint32_t filled_bar, uint32_t free_bar,
                                          uint32_t par, int lane) {
    using L = TmemLayout<R>;
    tmem_wait_ld32(wcur);
    if (free_bar) {
        // chunks whose last column has now been read: [first chunk not released by segment S-1, last chunk fully inside columns < end(S))
        constexpr int end_prev = S == 0 ? 0 : Seg<R>::col0(S - 1) + Seg<R>::ncol(S - 1);
        constexpr int end_cur = Seg<R>::col0(S) + Seg<R>::ncol(S);
        constexpr int lo = end_prev / CHUNK_COLS;                                   // chunks [0, lo) were released before
        constexpr int hi = (S == 17) ? L::NCH : end_cur / CHUNK_COLS;               // chunks [lo, hi) are complete now
#pragma unroll
        for (int c = lo; c < hi; ++c) signal_free(free_bar + 8 * c, lane);
    }
    if constexpr (S + 1 < 18) {
        if (filled_bar) {
            // the next segment reads columns [col0, col0 + ncol): wait for the chunks it touches that no earlier segment waited for
            constexpr int end_cur = Seg<R>::col0(S) + Seg<R>::ncol(S);
            constexpr int end_next = Seg<R>::col0(S + 1) + Seg<R>::ncol(S + 1);
            constexpr int lo = (end_cur - 1) / CHUNK_COLS + 1;
            constexpr int hi = (end_next - 1) / CHUNK_COLS;
#pragma unroll
            for (int c = lo; c <= hi; ++c) mbar_wait(filled_bar + 8 * c, par);
            if (lo <= hi) tc_fence_after();
        }
        seg_load<R, S + 1>(tbase, wnext);
    }
}

template <int R>
__device__ __forceinline__ void compute_pass2(const float* __restrict__ sp0, const float* __restrict__ sp1, bool two, uint32_t tbase,
                                              float (&acc0)[R], float (&acc1)[R], int xg, int W, uint32_t filled_bar,
                                              uint32_t free_bar, uint32_t par, int lane) {
    float wa[32], wb[32];
    if (filled_bar) {
        constexpr int hi = (Seg<R>::col0(0) + Seg<R>::ncol(0) - 1) / CHUNK_COLS;
#pragma unroll
        for (int c = 0; c <= hi; ++c) mbar_wait(filled_bar + 8 * c, par);
        tc_fence_after();
    }
    seg_load<R, 0>(tbase, wa);
#define PAMR_CENTRE(S, WC, WN)                                                    \
    seg_enter<R, S>(tbase, WC, WN, filled_bar, free_bar, par, lane);              \
    if (two) seg_centre_both<R, S>(sp0, sp1, WC, acc0, acc1);                     \
    else seg_centre_plane<R, S>(sp0, WC, acc0);
#define PAMR_SIDE(S, WC, WN)                                                      \
    seg_enter<R, S>(tbase, WC, WN, filled_bar, free_bar, par, lane);              \
    if (two) seg_side_both<R, ((S) - 6) / 6, ((S) - 6) % 6>(sp0, sp1, WC, acc0, acc1, xg, W); \
    else seg_side_plane<R, ((S) - 6) / 6, ((S) - 6) % 6>(sp0, WC, acc0, xg, W);
    PAMR_CENTRE(0, wa, wb) PAMR_CENTRE(1, wb, wa) PAMR_CENTRE(2, wa, wb) PAMR_CENTRE(3, wb, wa) PAMR_CENTRE(4, wa, wb) PAMR_CENTRE(5, wb, wa)
    PAMR_SIDE(6, wa, wb) PAMR_SIDE(7, wb, wa) PAMR_SIDE(8, wa, wb) PAMR_SIDE(9, wb, wa) PAMR_SIDE(10, wa, wb) PAMR_SIDE(11, wb, wa)
    PAMR_SIDE(12, wa, wb) PAMR_SIDE(13, wb, wa) PAMR_SIDE(14, wa, wb) PAMR_SIDE(15, wb, wa) PAMR_SIDE(16, wa, wb) PAMR_SIDE(17, wb, wa)
#undef PAMR_CENTRE
#undef PAMR_SIDE
}
#endif  // PAMR_CPP == 2

// Row strip.  The tiles cover [0,Wt) x [0,Ht); a remainder of at most RS_MAX_H rows (H = 321 -> one row) is not
// worth a tile row of its own.  It is cut into work items of 32 consecutive pixels of one row of one plane (one
// pixel per lane, neighbours straight from global memory / L2 with clamped coordinates, weights from the
// row-strip region of the affinity layout: coalesced).  The strip spans ALL columns, including its corner with
// the column strip.
template <bool FUSED>
__device__ __forceinline__ void strip_item(const Params& prm, const IterIO& io, int it, int item, int lane) {
    const int C = prm.C, H = prm.H, W = prm.W;
    const int hrows = H - prm.Ht, xblocks = (W + 31) / 32;
    const int per_plane = hrows * xblocks;
    const int plane = item / per_plane, r = item % per_plane;  // plane = b*C + c
    const int yi = r / xblocks, y = prm.Ht + yi, x = (r % xblocks) * 32 + lane;
    const bool valid = x < W;
    const int xc = min(x, W - 1);
    const int b = plane / C;
    if (FUSED && it > 0) {
        // fused iterations: the item reads rows >= Ht - 24 of iteration it - 1 (and overwrites what iteration it - 1 read
        // there): the image's bottom tile row and its row strip of that iteration must be complete
        if (lane == 0) {
            const int t0 = (b * prm.tiles_y + prm.tiles_y - 1) * prm.tiles_x;
            for (int tx = 0; tx < prm.tiles_x; ++tx) wait_at_least(prm.tile_epoch + t0 + tx, it);
            wait_at_least(prm.strip_count + b, it * prm.strip_items_per_image);
        }
        __syncwarp();
    }
    const float* __restrict__ pl = io.src + (size_t)plane * src_plane_stride(prm);
    const float* __restrict__ wp = prm.aff + prm.rs_base + ((size_t)b * hrows + yi) * 48 * W + xc;
    float acc = 0.f;  // one FMA chain in tap-sequence order: bit-identical to the tile kernel's result
#pragma unroll
    for (int s = 0; s < 48; ++s) {
        const int p = seq_tap(s), d = dil_of(p >> 3), j = p & 7;
        const int yy = clampi(y + tap_dy(j) * d, 0, H - 1);
        const int xx = clampi(xc + tap_dx(j) * d, 0, W - 1);
        acc = fmaf(__ldg(wp + (size_t)s * W), __ldg(pl + plane_index(prm.src_pitch, yy, xx)), acc);
    }
    if (valid) {
        *dst_pixel(prm, io, plane, y, x) = acc;
        store_replicas(prm, io, plane, y, x, acc);
    }
    if (io.cls_max != nullptr) {
        const unsigned m = __reduce_max_sync(0xffffffffu, valid ? ordered_from_float(acc) : 0u);
        if (lane == 0 && m != 0u) atomicMax(io.cls_max + plane, m);
    }
    if (FUSED) {  // this item of iteration `it` is stored
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(prm.strip_count + b) : "memory");
    }
}

// Offsets (floats) of the 48 neighbours of a column-strip pixel relative to the pixel itself inside the window, in
// tap-sequence order, for even / odd window rows: dy even moves whole row pairs, dy = +-1 depends on the row's
// parity; dx < 0 moves left; replicate padding clamps every dx > 0 tap to the strip column itself (it is image
// column W-1).  Constant-initialised: no upload.
struct CsOffsets {
    int off[2][48];
};
constexpr CsOffsets make_cs_offsets() {
    CsOffsets t{};
    for (int par = 0; par < 2; ++par)
        for (int s = 0; s < 48; ++s) {
            const int p = seq_tap(s), d = dil_of(p >> 3), j = p & 7;
            const int ty = (j < 3) ? -1 : (j < 5 ? 0 : 1);
            const int tx = (j == 0 || j == 3 || j == 5) ? -1 : ((j == 1 || j == 6) ? 0 : 1);
            int o = 0;
            if (ty != 0) o += (d != 1) ? ty * (d / 2) * ROWP : (ty < 0 ? (par ? -1 : 1 - ROWP) : (par ? ROWP - 1 : 1));
            if (tx < 0) o -= 2 * d;
            t.off[par][s] = o;
        }
    return t;
}
__constant__ CsOffsets c_cs = make_cs_offsets();

// Column strip x = Wt = W-1, computed by the tiles on the right image border (x0 + 32 == Wt) after every class
// pass from the window the warp just used (rows outside the image are the planes' replicated rows).  LP = 32/R lanes share a
// strip pixel: lane i*LP + part of warp wq takes taps [part*TPL, (part+1)*TPL) of pixel (row wq*R + i of the tile);
// the parts run one after the other on the same accumulator (handed on by shuffle), i.e. ONE FMA chain in
// tap-sequence order, bit-identical to every other path.  The lane's TPL weights come from its spare TMEM columns
// (one tcgen05.ld), so the pass costs only TPL shared-memory loads per warp: LSU instructions, not bytes, are
// what a warp on this kernel's saturated shared-memory pipe waits for.
#ifndef PAMR_CS_INLINE
#define PAMR_CS_INLINE __forceinline__
#endif
template <int R>
__device__ PAMR_CS_INLINE void column_strip_pass(const Params& prm, const IterIO& io, const float* __restrict__ slot, uint32_t tbase,
                                                  int plane, int y0, int wq, int lane) {
    constexpr int LP = Cfg<R>::CS_LP, TPL = Cfg<R>::CS_TPL;
    const int part = lane % LP, i = min(lane / LP, R - 1);
    const bool active = lane < R * LP;
    const int row = wq * R + i, wy = row + HALO;
    const float* __restrict__ v0 = slot + pair_index(WIN_P, wy, TX + HALO);
    const int* __restrict__ offs = c_cs.off[wy & 1] + part * TPL;
    float w[WB], v[TPL];
    tmem_ld16(tbase + Cfg<R>::CS_COL0, w);
#pragma unroll
    for (int k = 0; k < TPL; ++k) v[k] = v0[offs[k]];
    tmem_wait_ld(w);
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < LP; ++r) {
        if (part == r) {
#pragma unroll
            for (int k = 0; k < TPL; ++k) acc = fmaf(w[k], v[k], acc);
        }
        if (r + 1 < LP) {
            const float t = __shfl_up_sync(0xffffffffu, acc, 1);
            if (part == r + 1) acc = t;
        }
    }
    const int y = y0 + row;
    const bool valid = active && part == LP - 1 && y < prm.H;
    if (valid) {
        *dst_pixel(prm, io, plane, y, prm.Wt) = acc;
        store_replicas(prm, io, plane, y, prm.Wt, acc);
    }
    if (io.cls_max != nullptr) {
        const unsigned m = __reduce_max_sync(0xffffffffu, valid ? ordered_from_float(acc) : 0u);
        if (lane == 0 && m != 0u) atomicMax(io.cls_max + plane, m);
    }
}

// The schedule of one CTA: tile visits (iteration it, slot ti) in order.  In iteration it the CTA plays the role of
// "virtual CTA" v = (blockIdx.x + it * rot) % grid and takes the tiles v, v + grid, v + 2 grid, ...: the roles with one
// tile more than the others rotate over the CTAs (rot = number of short roles), so that a fused launch gives every CTA
// the same work over its iterations.  Plain launch: one iteration, v = blockIdx.x.
struct Visit {
    int it, ti, v, mt;  // iteration, tile slot in the iteration, virtual CTA index, tiles of this iteration
    int cta, grid, rot, ntiles;
    __device__ __forceinline__ int tiles_of(int vv) const { return (ntiles - vv + grid - 1) / grid; }
    __device__ __forceinline__ void start(int cta_, int grid_, int rot_, int ntiles_) {
        cta = cta_; grid = grid_; rot = rot_; ntiles = ntiles_;
        it = 0; ti = 0; v = cta; mt = tiles_of(v);
    }
    __device__ __forceinline__ void next() {
        if (++ti == mt) {
            ti = 0; ++it;
            v = (cta + it * rot) % grid;
            mt = tiles_of(v);
        }
    }
    __device__ __forceinline__ int tile() const { return v + ti * grid; }
};

template <int R, bool FUSED>
__global__ void __launch_bounds__(NTHREADS, 1)
propagate_sm100_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap1, const Params prm) {
    using C_ = Cfg<R>;
    using L = TmemLayout<R>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* slots = reinterpret_cast<float*>(smem_raw);
    unsigned char* stage_ring = smem_raw + C_::STAGE_OFF;
    const int wc = prm.W - prm.Wt;  // column strip width (0: none)
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + C_::CTRL_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = prm.C, H = prm.H, W = prm.W;
    const int npass_units = (C + CPP - 1) / CPP;                      // class passes per tile
    const int active_groups = npass_units < NG ? npass_units : NG;  // groups that have at least one class pass per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < NBAR; ++s) {
            mbar_init(smem_u32(&ctrl->tma_bar[s]), 1);
            mbar_init(smem_u32(&ctrl->empty_bar[s]), NW);  // the NW warps of the group that read the slot
        }
        for (int u = 0; u < MAX_CHUNKS; ++u) {
            mbar_init(smem_u32(&ctrl->filled_bar[u]), 1);                  // tcgen05.commit of an issuer
            mbar_init(smem_u32(&ctrl->free_bar[u]), NW * active_groups);   // every compute warp that reads weights
        }
        for (int g = 0; g < 2 * NSTG; ++g) mbar_init(smem_u32(&ctrl->staged_bar[g]), 1);
        mbar_init(smem_u32(&ctrl->csw_staged_bar), 1);
        mbar_init(smem_u32(&ctrl->csw_full_bar), 1);
        mbar_init(smem_u32(&ctrl->csw_free_bar), NW * active_groups);
        ctrl->done_cnt[0] = ctrl->done_cnt[1] = ctrl->done_cnt[2] = ctrl->done_cnt[3] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctrl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // Programmatic dependent launch (iterations 2..T are launched with the attribute): the next iteration's CTAs may
    // become resident as soon as SMs free up in this one's last wave, run their prologue and pull their first tile's
    // weights -- which do not depend on this iteration -- and only their producer warp waits for this grid to finish
    // (griddepcontrol.wait below) before it reads the class planes.  Both instructions are no-ops in a plain launch.
    asm volatile("griddepcontrol.launch_dependents;");

    const int tiles_per_img = prm.tiles_x * prm.tiles_y;
    // Tile visits of this CTA (struct Visit): gt counts them over all iterations of the launch (one iteration in a plain
    // launch).  Every ring, barrier phase and counter below runs on gt, so that iteration it + 1 follows iteration it
    // like one more tile: no cold start, no tail.
    constexpr bool fused = FUSED;
    const int n_iters = fused ? prm.iters : 1;
    const int rot = (fused && prm.ntiles > (int)gridDim.x) ? ((int)gridDim.x - prm.ntiles % (int)gridDim.x) % (int)gridDim.x : 0;
    int GT = 0;
    for (int i = 0; i < n_iters; ++i)
        GT += (prm.ntiles - ((int)blockIdx.x + i * rot) % (int)gridDim.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // tile-major affinity layout: the 48*R*128 weights of a tile are one contiguous block
    auto tile_weights = [&](int tile) -> const float* {
        const int b = tile / tiles_per_img, t = tile % tiles_per_img;
        return prm.aff + (((size_t)b * prm.tiles_y + t / prm.tiles_x) * prm.tiles_x + t % prm.tiles_x) * aff_tile_floats(R);
    };

    if (warp == W_PRODUCER) {
        // ===================== producer warp: TMA issue of the class planes =====================
        // Sequence number n = (tile_iter, class) -> slot n % NSLOT, barrier pair n % NBAR.  A consumer
        // group waits for tma_bar (bytes landed), computes, and releases the slot through empty_bar; the producer may refill slot
        // n % NSLOT once sequence number n - NSLOT has been released.
        const long long total = (long long)GT * C;
        [[maybe_unused]] int ev_n = 0;
        // the class planes are the previous iteration's output (and this iteration's stores go to the buffer the previous
        // one reads: every store depends on a plane that is loaded after this wait)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        Visit vis;
        vis.start((int)blockIdx.x, (int)gridDim.x, rot, prm.ntiles);
        int c = 0, gt = 0;
        for (long long n_issue = 0; n_issue < total; ++n_issue, ++c) {
            if (c == C) { c = 0; ++gt; vis.next(); }
            const int s = (int)(n_issue % NSLOT), bi = (int)(n_issue % NBAR);
            const int it = vis.it;
            const int tile = vis.tile();
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int x0 = (t % prm.tiles_x) * TX, y0 = (t / prm.tiles_x) * C_::TY;
            const CUtensorMap* tm = (fused && (it & 1)) ? &tmap1 : &tmap;
            if (fused && it > 0 && c == 0) {
                // Fused iterations: this tile reads a 24-pixel halo of iteration it - 1, and its stores overwrite what
                // iteration it - 1 read there: the tiles around it (same image) must have finished that iteration --
                // and, under a bottom-row tile, the image's row strip.  Lanes 0..8 watch one neighbour each, lane 9 the
                // strip (one L2 round trip for all of them instead of ten in a row).
                const int tx = t % prm.tiles_x, ty = t / prm.tiles_x;
                const int* flag = nullptr;
                int target = it;
                if (lane < 9) {
                    const int nx = tx + lane % 3 - 1, ny = ty + lane / 3 - 1;
                    if (nx >= 0 && nx < prm.tiles_x && ny >= 0 && ny < prm.tiles_y) flag = prm.tile_epoch + b * tiles_per_img + ny * prm.tiles_x + nx;
                } else if (lane == 9 && prm.strip_items_per_image > 0 && ty == prm.tiles_y - 1) {
                    flag = prm.strip_count + b;
                    target = it * prm.strip_items_per_image;
                }
                if (flag != nullptr) wait_at_least(flag, target);
                __syncwarp();
                asm volatile("fence.proxy.async;" ::: "memory");  // TMA (async proxy) reads what the other CTAs stored
            }
            if (lane == 0) {
                if (n_issue >= NSLOT) {  // the previous occupant of this slot has been consumed
                    const long long prev = n_issue - NSLOT;
                    mbar_wait(smem_u32(&ctrl->empty_bar[prev % NBAR]), (uint32_t)(prev / NBAR) & 1u);
                }
                const uint32_t bar = smem_u32(&ctrl->tma_bar[bi]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(bar, C_::SLOT_BYTES);
                // 64-bit elements = row pairs: coordinates (column, row pair, plane)
                tma_load_3d(smem_u32(slots + (size_t)s * C_::SLOT_FLOATS), tm, bar, x0 - HALO, y0 / 2, b * C + c);  // (y0 - HALO) / 2 + PADP
                if (PF_PLANES > 0 && n_issue + PF_PLANES < total) {  // pull a later plane window from HBM into L2
                    const long long np = n_issue + PF_PLANES;
                    const int pc = (int)(np % C);
                    const int ptile = tile;  // (experiment knob: same tile only)
                    const int pb = ptile / tiles_per_img, pt = ptile % tiles_per_img;
                    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                                 ::"l"(&tmap), "r"((pt % prm.tiles_x) * TX - HALO), "r"(((pt / prm.tiles_x) * C_::TY) / 2), "r"(pb * C + pc)
                                 : "memory");
                }
            }
            // Late in a tile, pull the NEXT tile's affinity weights (one contiguous 48*R*128-byte block
            // per lane quarter of the tile-major layout) from HBM into L2, one quarter per class, so that
            // the fill warps' loads hit L2.
            const int pq = c - prm.pf_class;
            if (pq >= 0 && pq < 4 && gt + 1 < GT && lane == 0) {
                Visit nv = vis;
                nv.next();
                const float* wp = tile_weights(nv.tile()) + (size_t)pq * (12 * R * 128);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wp), "r"(12 * R * 128 * 4) : "memory");
            }
            __syncwarp();
        }
    } else if (warp == W_LOADER) {
        // ===================== weight loader: L2 -> staging ring with cp.async.bulk =====================
        // The tiles' chunks form one stream n = tile_iter * NCH + chunk; chunk n uses stage n % NSTG.  A stage is
        // reusable once the tcgen05.cp set that read it has completed, which is what filled_bar of that chunk
        // reports.  (A control thread on a saturated SM pays a few hundred cycles per mbarrier round trip, hence
        // 32 KB per loop iteration.)
        if (lane == 0) {
            const int total = GT * L::NCH;
            [[maybe_unused]] int ev_n = 0;
            int cs_uses = 0;  // border tiles of this CTA so far
            Visit vis;
            vis.start((int)blockIdx.x, (int)gridDim.x, rot, prm.ntiles);
            for (int n = 0; n < total; ++n) {
                const int c = n % L::NCH, g = n % NSTG;
                if (c == 0 && n > 0) vis.next();
                const int tile = vis.tile();
                if (c == 0 && wc > 0) {  // a tile on the right image border: its column-strip weights -> shared memory
                    const int b = tile / tiles_per_img, t = tile % tiles_per_img;
                    if (t % prm.tiles_x == prm.tiles_x - 1) {
                        if (cs_uses > 0) mbar_wait(smem_u32(&ctrl->csw_full_bar), (uint32_t)(cs_uses - 1) & 1u);  // the staging area has been read
                        const uint32_t bar = smem_u32(&ctrl->csw_staged_bar);
                        mbar_arrive_expect_tx(bar, C_::CSW_BYTES);
                        bulk_load(smem_u32(smem_raw + C_::CSW_OFF),
                                  prm.aff + prm.cs_base + ((size_t)b * prm.tiles_y + t / prm.tiles_x) * aff_cs_block_floats(R), C_::CSW_BYTES, bar);
                        ++cs_uses;
                    }
                }
                if (n >= NSTG) {
                    const int pn = n - NSTG;
                    mbar_wait(smem_u32(&ctrl->filled_bar[pn % L::NCH]), (uint32_t)(pn / L::NCH) & 1u);
                }
                const int units = min(CHUNK_UNITS, L::NU - c * CHUNK_UNITS);
                const uint32_t bar = smem_u32(&ctrl->staged_bar[n % (2 * NSTG)]);
                mbar_arrive_expect_tx(bar, units * UNIT_BYTES);
                bulk_load(smem_u32(stage_ring + (size_t)g * (CHUNK_UNITS * UNIT_BYTES)),
                          tile_weights(tile) + (size_t)c * (CHUNK_UNITS * UNIT * 128), units * UNIT_BYTES, bar);
                PAMR_EV(3, true, 2000 + c);
            }
        }
    } else if (warp >= W_ISSUER) {
        // ===================== weight issuers: staging ring -> TMEM with tcgen05.cp =====================
        // two control threads (warps 14 and 15) take the even and the odd chunks of the stream
        if (lane == 0) {
            const int total = GT * L::NCH;
            const uint32_t tb = ctrl->tmem_base;
            [[maybe_unused]] int ev_n = 0;
            int n = warp - W_ISSUER;
            // Waits are by phase parity, so a waiter must never be a whole phase early on its barrier.  With an even number
            // of chunks per tile a chunk index always meets the same issuer; with an odd number the two issuers alternate
            // on free_bar[c], which is still sound because the loader throttles the stream: an issuer can be at most
            // NSTG chunks ahead of the other one, i.e. less than a tile (NCH chunks) apart.
            static_assert(L::NCH % 2 == 0 || NSTG < L::NCH, "the issuers must stay less than a tile apart");
            if (n < total) mbar_wait(smem_u32(&ctrl->staged_bar[n % (2 * NSTG)]), (uint32_t)(n / (2 * NSTG)) & 1u);  // bytes landed
            int cs_uses = 0;  // border tiles of this CTA so far
            Visit vis;
            vis.start((int)blockIdx.x, (int)gridDim.x, rot, prm.ntiles);
            int vgt = 0;
            for (; n < total; n += 2) {
                const int gt = n / L::NCH, c = n % L::NCH, g = n % NSTG;
                while (vgt < gt) { vis.next(); ++vgt; }
                if (c == 0 && wc > 0) {  // (chunk 0 always meets issuer 0) a border tile: its strip weights -> spare TMEM columns
                    const int tile = vis.tile();
                    if ((tile % tiles_per_img) % prm.tiles_x == prm.tiles_x - 1) {
                        mbar_wait(smem_u32(&ctrl->csw_staged_bar), (uint32_t)cs_uses & 1u);
                        if (cs_uses > 0) mbar_wait(smem_u32(&ctrl->csw_free_bar), (uint32_t)(cs_uses - 1) & 1u);  // previous border tile done
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem_raw + C_::CSW_OFF);
#pragma unroll
                        for (int k = 0; k < C_::CS_TPL / 4; ++k)  // 4 columns each: 128 lanes x 16 bytes
                            asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tb + C_::CS_COL0 + k * 4), "l"(utccp_desc(sa + k * 2048, 0, 128)) : "memory");
                        utccp_commit(smem_u32(&ctrl->csw_full_bar));
                        ++cs_uses;
                    }
                }
                if (gt > 0) mbar_wait(smem_u32(&ctrl->free_bar[c]), (uint32_t)(gt - 1) & 1u);  // the previous tile's last passes have read the chunk
                tc_fence_after();
                PAMR_EV(4, warp == W_ISSUER, 2100 + c);
                const int units = min(CHUNK_UNITS, L::NU - c * CHUNK_UNITS);
                const uint32_t sa = smem_u32(stage_ring + (size_t)g * (CHUNK_UNITS * UNIT_BYTES));
#pragma unroll
                for (int k = 0; k < CHUNK_UNITS * UNIT / 8; ++k)  // 8 columns each: pieces 2k, 2k+1 (2 KB apart), 8-lane groups 128 B apart
                    if (k < units * (UNIT / 8)) utccp_128x256b(tb + c * (CHUNK_UNITS * UNIT) + k * 8, utccp_desc(sa + k * 4096, 2048, 128));
                utccp_commit(smem_u32(&ctrl->filled_bar[c]));
                // while the next chunk is still being read: make sure its bytes have landed
                if (n + 2 < total) mbar_wait(smem_u32(&ctrl->staged_bar[(n + 2) % (2 * NSTG)]), (uint32_t)((n + 2) / (2 * NSTG)) & 1u);
            }
        }
    } else if (warp < NWC) {
        // ===================== compute warps: NG groups x NW warps =====================
        const int grp = warp / NW, wq = warp % NW;  // wq = TMEM lane quarter; all groups share the weights
        const uint32_t tbase = ctrl->tmem_base + ((uint32_t)(wq * 32) << 16);
        const uint32_t filled0 = smem_u32(&ctrl->filled_bar[0]), free0 = smem_u32(&ctrl->free_bar[0]);
        int cs_seen = 0;  // border tiles of this CTA so far
        [[maybe_unused]] int ev_n = 0;
        // Fused iterations: a finished tile is announced (epoch in global memory) one class pass LATE, when the warp's
        // stores have long drained and the fences that order them before the announcement cost nothing; issued right
        // after the stores they would wait ~2 k cycles per tile.  Each warp orders its stores before its count at CTA
        // scope; the warp that completes the count pays the one gpu-scope fence and publishes the epoch (the
        // block-barrier + one-thread-fence idiom of a grid-wide barrier, per tile).
        int pend_tile = -1, pend_it = 0, pend_slot = 0;
        auto announce = [&]() {
            if (pend_tile < 0) return;
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.acq_rel.cta;" ::: "memory");
                const int old = atomicAdd(&ctrl->done_cnt[pend_slot], 1);
                if (old == NW * active_groups - 1) {
                    ctrl->done_cnt[pend_slot] = 0;  // (the groups are never two tiles apart: the slot is not in use again yet)
                    __threadfence();
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(prm.tile_epoch + pend_tile), "r"(pend_it + 1) : "memory");
                }
            }
            pend_tile = -1;
        };
        Visit vis;
        vis.start((int)blockIdx.x, (int)gridDim.x, rot, prm.ntiles);
        for (int gt = 0; gt < GT; ++gt, vis.next()) {
            const int it = vis.it, ti = vis.ti;
            const int tile = vis.tile();
            // ---- row strip y in [Ht,H): the roles that own one tile fewer than the others take it.  In a plain launch
            //      after their last tile (they would idle in the last wave); in a fused launch BEFORE their first tile of
            //      the iteration: the strip of iteration it depends on iteration it - 1 only, and the bottom tiles of
            //      iteration it + 1 wait for it
            if (fused && ti == 0 && prm.tail_cta0 >= 0 && vis.v >= prm.tail_cta0) {
                announce();  // (the strip may wait for the very tile this warp has not announced yet)
                const int nw = ((int)gridDim.x - prm.tail_cta0) * NWC;
                for (int item = (vis.v - prm.tail_cta0) * NWC + warp; item < prm.strip_items; item += nw)
                    strip_item<FUSED>(prm, iter_io<FUSED>(prm, it), it, item, lane);
            }
            const int b = tile / tiles_per_img, t = tile % tiles_per_img;
            const int x0 = (t % prm.tiles_x) * TX, y0 = (t / prm.tiles_x) * C_::TY;
            const int x = x0 + lane, yw = y0 + wq * R;
            const bool xok = x < W;
            const long long seq0 = (long long)gt * C;
            const int nrow = xok ? max(0, min(R, H - yw)) : 0;  // rows of this thread inside the image
            const bool cs_tile = wc > 0 && (t % prm.tiles_x == prm.tiles_x - 1);  // this tile also computes the column strip
            if (cs_tile && grp < active_groups) {
                mbar_wait(smem_u32(&ctrl->csw_full_bar), (uint32_t)cs_seen & 1u);  // its strip weights are in TMEM
                tc_fence_after();
                ++cs_seen;
            }

            // ---- store of one class plane's results (+ optional class max)
            auto store_plane = [&](int k, const float (&acc)[R]) {
                const IterIO io = iter_io<FUSED>(prm, it);  // (derived here, not held across the pass: the pass needs every register)
                const int plane = b * C + k;
                [[maybe_unused]] float acc_fix = 0.f;
                if (io.dst_pair) {  // row pairs: 8-byte stores, 256 contiguous bytes per warp
                    float2* __restrict__ op = reinterpret_cast<float2*>(io.dst) +
                                              ((size_t)plane * prm.Hp2 + PADP + (yw >> 1)) * io.dst_pitch + x;
                    const size_t pitch = (size_t)io.dst_pitch;
                    // replicate padding as data: the owner of row 0 / row H-1 also fills the 24 rows beyond it (a warp-uniform
                    // case: only the first warp of the top tiles and the warp that holds row H-1)
                    const int il = H - 1 - yw;  // index of row H-1 in this thread's strip, if it is there
                    if (yw == 0 && xok) {
                        const float2 v = make_float2(acc[0], acc[0]);
                        float2* __restrict__ pp = op - (size_t)PADP * pitch;
#pragma unroll
                        for (int r = 0; r < PADP; ++r, pp += pitch) *pp = v;
                    }
                    if (il >= 0 && il < R && xok) {
                        float vl = acc[0];
#pragma unroll
                        for (int i = 1; i < R; ++i)
                            if (i == il) vl = acc[i];
                        float2* __restrict__ pp = op + (size_t)((il >> 1) + 1) * pitch;  // the pair after the one that holds row H-1
#pragma unroll
                        for (int r = 0; r < PADP; ++r, pp += pitch) *pp = make_float2(vl, vl);
#pragma unroll
                        for (int i = 0; i < R; i += 2)  // odd H: the second half of the last pair repeats row H-1 as well
                            if (i == il) acc_fix = vl;
                    }
#pragma unroll
                    for (int i = 0; i < R; i += 2, op += pitch)
                        if (i < nrow) *op = make_float2(acc[i], (i == il) ? acc_fix : acc[i + 1]);
                } else {  // last iteration: the caller's [B,C,H,W] tensor, coalesced 128-byte rows
                    float* __restrict__ op = io.dst + ((size_t)plane * H + yw) * io.dst_pitch + x;
                    const size_t pitch = (size_t)io.dst_pitch;
                    if (nrow == R) {
#pragma unroll
                        for (int i = 0; i < R; ++i, op += pitch) *op = acc[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < R; ++i, op += pitch)
                            if (i < nrow) *op = acc[i];
                    }
                }
                if (io.cls_max != nullptr) {  // last iteration only: keep its ALU work out of the other nine
                    unsigned mx = 0u;
#pragma unroll
                    for (int i = 0; i < R; ++i)
                        if (i < nrow) mx = max(mx, ordered_from_float(acc[i]));
                    mx = __reduce_max_sync(0xffffffffu, mx);
                    if (lane == 0 && mx != 0u) atomicMax(io.cls_max + plane, mx);
                }
            };
            int probe = 0;  // 1: the barriers of this group's next class were already seen complete
#if PAMR_CPP == 2
            {
                // ---- two class planes per pass: pair p = classes 2p, 2p+1 -> group p % NG
                const int npairs = (C + 1) / 2;
                for (int pp = grp; pp < npairs; pp += NG) {
                    const int k0 = 2 * pp, k1 = k0 + 1;
                    const bool two = k1 < C;
                    const long long sq0 = seq0 + k0, sq1 = sq0 + 1;
                    const int bi0 = (int)(sq0 % NBAR), bi1 = (int)(sq1 % NBAR);
                    PAMR_EV(grp, wq == 0 && lane == 0, 100 + k0);
                    if (!probe) mbar_wait(smem_u32(&ctrl->tma_bar[bi0]), (uint32_t)(sq0 / NBAR) & 1u);
                    if (two) mbar_wait(smem_u32(&ctrl->tma_bar[bi1]), (uint32_t)(sq1 / NBAR) & 1u);
                    PAMR_EV(grp, wq == 0 && lane == 0, 6);
                    float* slot0 = slots + (size_t)(sq0 % NSLOT) * C_::SLOT_FLOATS;
                    float* slot1 = slots + (size_t)(sq1 % NSLOT) * C_::SLOT_FLOATS;
                    const int spo = ((wq * R + HALO) / 2) * ROWP + (lane + HALO) * 2;
                    float acc0[R], acc1[R];
#pragma unroll
                    for (int i = 0; i < R; ++i) acc0[i] = acc1[i] = 0.f;
#if defined(PAMR_EXPERIMENTS) && defined(PAMR_X_NOCHASE)    // timing experiment: no weight hand-over at all (wrong values)
                    const bool first = false, last = false;
#else
                    const bool first = (pp == grp), last = (pp + NG >= npairs);
#endif
                    compute_pass2<R>(slot0 + spo, slot1 + spo, two, tbase, acc0, acc1, x, W, first ? filled0 : 0u, last ? free0 : 0u,
                                     (uint32_t)gt & 1u, lane);
                    if (cs_tile) {
                        column_strip_pass<R>(prm, iter_io<FUSED>(prm, it), slot0, tbase, b * C + k0, y0, wq, lane);
                        if (two) column_strip_pass<R>(prm, iter_io<FUSED>(prm, it), slot1, tbase, b * C + k1, y0, wq, lane);
                        if (last) {  // this warp is done with the tile's strip weights
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&ctrl->csw_free_bar));
                        }
                    }
                    PAMR_EV(grp, wq == 0 && lane == 0, 9);
                    __syncwarp();
                    if (lane == 0) {  // release the slots as early as possible
                        mbar_arrive(smem_u32(&ctrl->empty_bar[bi0]));
                        if (two) mbar_arrive(smem_u32(&ctrl->empty_bar[bi1]));
                    }
                    probe = 0;
                    if (pp + NG < npairs) {
                        const long long sq2 = sq0 + 2 * NG;
                        probe = (int)mbar_poll(smem_u32(&ctrl->tma_bar[sq2 % NBAR]), (uint32_t)(sq2 / NBAR) & 1u);
                    }
                    store_plane(k0, acc0);
                    if (two) store_plane(k1, acc1);
                    if (fused && first) announce();  // the previous tile of this warp
                    PAMR_EV(grp, wq == 0 && lane == 0, 8);
                }
            }
#else
            {
            for (int k = grp; k < C; k += NG) {
                const long long sq = seq0 + k;
                const int s = (int)(sq % NSLOT), bi = (int)(sq % NBAR);
                const uint32_t par = (uint32_t)(sq / NBAR) & 1u;
                PAMR_EV(grp, wq == 0 && lane == 0, 100 + k);
                if (!probe) mbar_wait(smem_u32(&ctrl->tma_bar[bi]), par);  // bytes landed
                PAMR_EV(grp, wq == 0 && lane == 0, 6);
                float* slot = slots + (size_t)s * C_::SLOT_FLOATS;
                const float* sp = slot + ((wq * R + HALO) / 2) * ROWP + (lane + HALO) * 2;
                float acc[R];
#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] = 0.f;
#if defined(PAMR_EXPERIMENTS) && defined(PAMR_X_NOCHASE)    // timing experiment: no weight hand-over at all (wrong values)
                const bool first = false, last = false;
#else
                const bool first = (k == grp), last = (k + NG >= C);
#endif
#ifdef PAMR_PASS_SPECIALISED
                if (first && last) compute_pass<R, true, true>(sp, tbase, acc, x, W, filled0, free0, (uint32_t)gt & 1u, lane);
                else if (first) compute_pass<R, true, false>(sp, tbase, acc, x, W, filled0, free0, (uint32_t)gt & 1u, lane);
                else if (last) compute_pass<R, false, true>(sp, tbase, acc, x, W, filled0, free0, (uint32_t)gt & 1u, lane);
                else compute_pass<R, false, false>(sp, tbase, acc, x, W, filled0, free0, (uint32_t)gt & 1u, lane);
#else
                compute_pass<R>(sp, tbase, acc, x, W, first ? filled0 : 0u, last ? free0 : 0u, (uint32_t)gt & 1u, lane);
#endif
                if (cs_tile) {
                    column_strip_pass<R>(prm, iter_io<FUSED>(prm, it), slot, tbase, b * C + k, y0, wq, lane);
                    if (last) {  // this warp is done with the tile's strip weights
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&ctrl->csw_free_bar));
                    }
                }
                PAMR_EV(grp, wq == 0 && lane == 0, 9);
                // release the slot as early as possible
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctrl->empty_bar[bi]));
                // Probe the next class's barrier now, without blocking: its latency overlaps the stores below
                // instead of sitting at the head of the next pass.
                probe = 0;
                if (k + NG < C) {
                    const long long sq2 = sq + NG;
                    probe = (int)mbar_poll(smem_u32(&ctrl->tma_bar[sq2 % NBAR]), (uint32_t)(sq2 / NBAR) & 1u);
                }
                store_plane(k, acc);
                if (fused && first) announce();  // the previous tile of this warp
                PAMR_EV(grp, wq == 0 && lane == 0, 8);
            }
            }
#endif
            if (fused && grp < active_groups) {  // this warp has stored its share of the tile: tell the others -- later (below)
                pend_tile = tile;
                pend_it = it;
                pend_slot = gt & 3;
                // (with a single tile per iteration the next visit is the same tile one iteration later, which waits for
                // neighbours that wait for this very announcement: no deferral then)
                if (vis.mt < 2) announce();
            }
            if (!fused && ti == vis.mt - 1 && prm.tail_cta0 >= 0 && (int)blockIdx.x >= prm.tail_cta0) {
                const int nw = ((int)gridDim.x - prm.tail_cta0) * NWC;
                for (int item = ((int)blockIdx.x - prm.tail_cta0) * NWC + warp; item < prm.strip_items; item += nw)
                    strip_item<FUSED>(prm, iter_io<FUSED>(prm, it), it, item, lane);
            }
        }
        if (fused) announce();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(ctrl->tmem_base), "r"(512));
}

// Row strip as a launch of its own: only when the tile kernel's last wave has no idle CTAs for it.
__global__ void __launch_bounds__(128) strip_rows_kernel(const Params prm) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const IterIO io = iter_io<false>(prm, 0);
    for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < prm.strip_items; item += gridDim.x * wpb)
        strip_item<false>(prm, io, 0, item, lane);
}

// Copy [planes,H,W] -> padded row-pair layout [planes,Hpp,Wp,2], Hpp = padded_pairs(H) (TMA reads 64-bit elements; Wp
// a multiple of 16 keeps its global strides multiples of 16 bytes).  Row pair p holds image rows 2(p - PADP) and
// 2(p - PADP) + 1, clamped to [0, H-1]: the replicated rows above / below the image and, for odd H, the repeat of row H-1.
__global__ void __launch_bounds__(128) repack_pairs_kernel(const float* __restrict__ src, float2* __restrict__ dst, int H, int W,
                                                           int Hpp, int Wp, size_t pair_rows) {
    for (size_t pr = blockIdx.x; pr < pair_rows; pr += gridDim.x) {
        const size_t plane = pr / Hpp;
        const int p = (int)(pr % Hpp) - PADP;
        const float* __restrict__ s0 = src + (plane * H + min(max(2 * p, 0), H - 1)) * W;
        const float* __restrict__ s1 = src + (plane * H + min(max(2 * p + 1, 0), H - 1)) * W;
        float2* __restrict__ d = dst + pr * Wp;
        for (int x = threadIdx.x; x < W; x += blockDim.x) d[x] = make_float2(__ldg(s0 + x), __ldg(s1 + x));
    }
}

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

// 3-D map over 64-bit elements (one element = the two rows of a row pair at one column)
int make_tmap(CUtensorMap* map, const float* base, int planes, int Hp2, int Wp, int win_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return set_error(PAMR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)Hp2, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)Wp * 8, (cuuint64_t)Wp * 8 * (cuuint64_t)Hp2};
    cuuint32_t box[3] = {(cuuint32_t)WIN_P, (cuuint32_t)(win_h / 2), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PAMR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PAMR_OK;
}

// Row strip placement: when the tile count leaves the last wave partly empty, the CTAs without a
// tile in that wave take the row-strip items (one warp per item, ~2.5 us each, TAIL_ITEMS_MAX in a
// row still fit inside one tile time); otherwise the strip is a launch of its own.
constexpr int TAIL_ITEMS_MAX = 8;
inline bool row_strip_in_tail(long long items, long long ntiles, int grid) {
    if (grid <= 0 || ntiles <= grid || ntiles % grid == 0) return false;
    const long long warps = (long long)(grid - ntiles % grid) * NWC;
    return (items + warps - 1) / warps <= TAIL_ITEMS_MAX;
}

// fused: nullptr (one iteration src -> dst) or the description of a launch that runs all iterations
struct FusedArgs {
    int iters;
    const float* buf0;
    const float* buf1;
    int* epochs;  // [ntiles + B] zeroed ints
};
template <int R>
int launch_one(const float* aff, const AffTiling& tiling, const float* src, int src_pitch, float* dst, int dst_pitch,
               bool dst_pair, int B, int C, int H, int W, unsigned* cls_max, int sm_count, int dev, bool dependent,
               const FusedArgs* fused, cudaStream_t s) {
    using C_ = Cfg<R>;
    // function attributes are per device: set once per (kernel, device)
    static std::atomic<int> attr_set[64];
    if (dev < 0 || dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
        PAMR_CUDA_TRY(cudaFuncSetAttribute(propagate_sm100_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C_::SMEM_BYTES));
#ifdef PAMR_FUSED_ITERATIONS
        PAMR_CUDA_TRY(cudaFuncSetAttribute(propagate_sm100_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)C_::SMEM_BYTES));
#endif
        if (dev >= 0 && dev < 64) attr_set[dev].store(1, std::memory_order_release);
    }
    const int Hp2 = padded_pairs(H), Wt = tiling.Wt, Ht = tiling.Ht;
    alignas(64) CUtensorMap tmap, tmap1;
    int rc = make_tmap(&tmap, fused ? fused->buf0 : src, B * C, Hp2, src_pitch, C_::WIN_H);
    if (rc != PAMR_OK) return rc;
    if ((rc = make_tmap(&tmap1, fused ? fused->buf1 : src, B * C, Hp2, src_pitch, C_::WIN_H)) != PAMR_OK) return rc;
    Params p;
    p.aff = aff; p.src = src; p.dst = dst; p.cls_max = cls_max;
    p.src_pitch = src_pitch; p.dst_pitch = dst_pitch; p.dst_pair = dst_pair ? 1 : 0; p.Hp2 = Hp2;
    p.pf_class = C >= 8 ? C - 8 : 0;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.tiles_x = tiling.tiles_x;
    p.tiles_y = tiling.tiles_y;
    p.Wt = Wt; p.Ht = Ht;
    p.cs_base = tiling.cs_base; p.rs_base = tiling.rs_base;
    p.strip_items = 0;
    p.tail_cta0 = -1;
    p.iters = 1; p.buf[0] = p.buf[1] = nullptr; p.out = nullptr; p.out_cls_max = nullptr; p.out_pitch = 0;
    p.tile_epoch = nullptr; p.strip_count = nullptr; p.strip_items_per_image = 0;
#ifdef PAMR_EXPERIMENTS
    p.dbg = nullptr;
    p.dbg_cta = g_timeline_cta.load(std::memory_order_relaxed);
    if (g_timeline.load() != nullptr && g_timeline_skip.fetch_sub(1) == 0) p.dbg = g_timeline.exchange(nullptr);  // record one launch
#endif
    const long long ntiles = (long long)p.tiles_x * p.tiles_y * B;
    if (ntiles > 0x7fffffffLL) return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: too many tiles");
    p.ntiles = (int)ntiles;
    const int grid = p.ntiles < sm_count ? p.ntiles : sm_count;
    const long long row_items = (long long)B * C * (H - Ht) * ((W + 31) / 32);
    if (row_items > 0x7fffffffLL) return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: row strip too large");
    if (fused != nullptr) {
        if (Ht < H && !row_strip_in_tail(row_items, ntiles, grid))
            return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: fused launch needs the row strip in the tail");
        p.iters = fused->iters;
        p.buf[0] = fused->buf0; p.buf[1] = fused->buf1;
        p.out = dst; p.out_cls_max = cls_max; p.out_pitch = dst_pitch;
        p.src = nullptr; p.dst = nullptr; p.cls_max = nullptr;
        p.tile_epoch = fused->epochs;
        p.strip_count = fused->epochs + p.ntiles;
        p.strip_items_per_image = (int)(row_items / B);
    }
    if (Ht < H && row_strip_in_tail(row_items, ntiles, grid)) {
        p.strip_items = (int)row_items;  // the tile kernel's short CTAs do the row strip
        p.tail_cta0 = p.ntiles % grid;
    } else if (Ht < H) {  // row strip y in [Ht,H) as a launch of its own (reads iteration t-1 like the tiles)
        Params pr = p;
        pr.strip_items = (int)row_items;
        const int blocks = (int)((row_items + 3) / 4);
        strip_rows_kernel<<<blocks < 8 * sm_count ? blocks : 8 * sm_count, 128, 0, s>>>(pr);
        count_launch();
        PAMR_CUDA_TRY(cudaGetLastError());
    }
#ifdef PAMR_NO_PDL
    dependent = false;
#endif
    if (dependent && !(Ht < H && p.tail_cta0 < 0)) {  // (no separate row-strip launch in between)
        // launched behind the previous iteration's tile kernel with programmatic stream serialization
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = C_::SMEM_BYTES;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PAMR_CUDA_TRY(cudaLaunchKernelEx(&cfg, propagate_sm100_kernel<R, false>, tmap, tmap1, p));
    } else {
#ifdef PAMR_FUSED_ITERATIONS
        if (fused != nullptr) propagate_sm100_kernel<R, true><<<grid, NTHREADS, C_::SMEM_BYTES, s>>>(tmap, tmap1, p);
        else
#else
        if (fused != nullptr) return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: built without PAMR_FUSED_ITERATIONS");
#endif
        propagate_sm100_kernel<R, false><<<grid, NTHREADS, C_::SMEM_BYTES, s>>>(tmap, tmap1, p);
    }
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

}  // namespace

int device_sm_count(int dev, int* out) {
    static std::atomic<int> cache[64];
    int n = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (n == 0) {
        PAMR_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
    }
    *out = n;
    return PAMR_OK;
}

// 1-D map over a flat fp32 array (pamr_affinity.cu: image rows are fetched one box of `box` floats at a time, at any
// 4-byte offset -- W needs no alignment; only the base must be 16-byte aligned).  `map` points at a CUtensorMap.
int encode_tensor_map_1d_f32(void* map, const float* base, unsigned long long elems, unsigned box) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return set_error(PAMR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[1] = {(cuuint64_t)elems};
    cuuint64_t strides[1] = {0};  // rank - 1 = 0 entries are read
    cuuint32_t bx[1] = {(cuuint32_t)box};
    cuuint32_t estr[1] = {1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, const_cast<float*>(base), dims,
                    strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PAMR_ERR_CUDA, "cuTensorMapEncodeTiled (1-D) failed with CUresult %d", (int)r);
    return PAMR_OK;
}

#ifdef PAMR_EXPERIMENTS
// Debug hook of experiment builds only (not part of the ABI): device buffer of 5 x 4096 x 2 int64 that CTA `cta`
// of ONE tuned launch (after skipping `skip` launches) fills with {clock64, event code} pairs.
extern "C" void pamr_debug_set_timeline(long long* dev_buf, int cta, int skip) {
    g_timeline_cta.store(cta);
    g_timeline_skip.store(skip);
    g_timeline.store(dev_buf);
}
#endif

// Tiling of the tuned kernel (R rows per thread, tile = 32 x 4R, extent [0,Wt) x [0,Ht) covered by tiles) and
// the affinity layout that goes with it, or R == 0 when the tuned kernel does not apply.  Remainders have an
// alternative to a padded tile row / partial tile column:
//   rows (<= RS_MAX_H):    the row strip -- inside the tile kernel's last wave when that wave has idle CTAs
//                          (row_strip_in_tail), else a launch of its own;
//   columns (<= CS_MAX_W): the column strip, computed by the tiles on the right image border.
// All combinations with R in {8,10} are priced with a time model fitted to measurements on B200 -- waves of
// tiles over the SMs at a cost per row-per-thread, border tiles a little dearer per strip column, row-strip
// launch ~ 8 + 1.3 us per 1000 items -- and the cheapest wins.  E.g. 321 x 321, B=16: R=10, 10 x 8 tiles, column
// 320 by the border tiles, row 320 in the tail of the same launch.  All paths add the 48 products of a pixel in
// the same order, so the result does not depend on the tiling (nor, therefore, on how a batch is sharded).  The
// row-strip placement is priced for the caller's class count; the SM count is the current device's.
AffTiling tuned_tiling(int B, int C, int H, int W, const Dilations& dil) {
    static const int want[6] = {1, 2, 4, 8, 12, 24};
    AffTiling t{};
    if (dil.nd != 6 || W < TX || H < 8) return t;
    for (int i = 0; i < 6; ++i)
        if (dil.d[i] != want[i]) return t;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || device_sm_count(dev, &sms) != PAMR_OK || sms <= 0) sms = 148;
    double best_cost = 1e30;
    for (int r = 8; r <= 10; r += 2) {
        const int ty = NW * r;
        const double tile_us = 1.8 * r * (r == 10 ? 1.00 : 1.07);
        const int hrem = H % ty, wrem = W % TX;
        const bool rs_ok = hrem != 0 && hrem <= RS_MAX_H && H > ty;
        const bool cs_ok = wrem != 0 && wrem <= CS_MAX_W && W > TX;
        for (int rs = 0; rs < 2; ++rs) {      // rs: row remainder as a strip
            if (rs && !rs_ok) continue;
            const int ht = rs ? H - hrem : H;
            for (int cs = 0; cs < 2; ++cs) {  // cs: column remainder as a strip of the border tiles
                if (cs && !cs_ok) continue;
                const int wt = cs ? W - wrem : W;
                const int txs = (wt + TX - 1) / TX, tys = (ht + ty - 1) / ty;
                const long long ntiles = (long long)B * txs * tys;
                if (ntiles > 0x7fffffffLL) continue;
                double cost = (double)((ntiles + sms - 1) / sms) * tile_us;
                if (rs) {
                    const long long items = (long long)B * C * hrem * ((W + 31) / 32);
                    const int grid = ntiles < sms ? (int)ntiles : sms;
                    if (!row_strip_in_tail(items, ntiles, grid)) cost += 8.0 + 1.3e-3 * (double)items;
                }
                if (cs) cost *= 1.0 + 0.1 / txs;  // one tile in txs also does the strip column
                if (cost < best_cost) {
                    best_cost = cost;
                    t.R = r; t.Wt = wt; t.Ht = ht; t.tiles_x = txs; t.tiles_y = tys;
                }
            }
        }
    }
    if (t.R == 0) return t;  // nothing fits (gigantic B*H*W): the generic kernel takes it
    t.W = W; t.H = H;
    aff_layout_finish(t, B);
    return t;
}

// pitch (in 64-bit elements) of the row-pair layout for image width W: whole 128-byte lines per row pair, so that
// the rows TMA fetches and the rows the warps store start on line boundaries for any W
int pair_pitch(int W) { return (W + 15) / 16 * 16; }
// row pairs of a plane of the ping-pong buffers, including the replicated pairs above and below the image
int pair_rows_padded(int H) { return padded_pairs(H); }

int launch_repack_pairs(const float* src, float* dst, int planes, int H, int W, cudaStream_t s) {
    const int Hp2 = padded_pairs(H), Wp = pair_pitch(W);
    const size_t pair_rows = (size_t)planes * Hp2;
    const unsigned grid = (unsigned)(pair_rows < 148 * 16 ? pair_rows : 148 * 16);
    repack_pairs_kernel<<<grid, 128, 0, s>>>(src, reinterpret_cast<float2*>(dst), H, W, Hp2, Wp, pair_rows);
    count_launch();
    PAMR_CUDA_TRY(cudaGetLastError());
    return PAMR_OK;
}

// One propagation step src -> dst: the persistent tile kernel (which also takes the column strip and, when its
// last wave has room, the row strip) and otherwise a small row-strip launch.  src is in the row-pair layout
// (pitch pair_pitch(W), 16-byte aligned base); dst is in the row-pair layout (dst_pair) or the caller's standard layout.
int launch_propagate_tuned(const float* aff_tiled, const AffTiling& tiling, const float* src, float* dst, int dst_pitch,
                           bool dst_pair, int B, int C, int H, int W, unsigned* cls_max, int dev, bool dependent, cudaStream_t s) {
    int sm_count = 0;
    int rc = device_sm_count(dev, &sm_count);
    if (rc != PAMR_OK) return rc;
    if (((uintptr_t)src & 15) != 0 || ((uintptr_t)aff_tiled & 15) != 0)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: source / affinity base not 16-byte aligned");
    const int sp = pair_pitch(W);
    if (tiling.R == 8)
        return launch_one<8>(aff_tiled, tiling, src, sp, dst, dst_pitch, dst_pair, B, C, H, W, cls_max, sm_count, dev, dependent, nullptr, s);
    if (tiling.R == 10)
        return launch_one<10>(aff_tiled, tiling, src, sp, dst, dst_pitch, dst_pair, B, C, H, W, cls_max, sm_count, dev, dependent, nullptr, s);
    return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: unsupported tiling R=%d", tiling.R);
}

// All `iters` (>= 2) propagation steps in ONE launch of the tile kernel: iteration it reads buf[it & 1] and writes
// buf[(it + 1) & 1] (both in the padded row-pair layout; the input is in buf0), the last one writes `out` ([B,C,H,W]).
// Tiles of iteration it + 1 start as soon as the tiles around them have finished iteration it (epochs in global
// memory), so the launch has one cold start and one tail instead of `iters` of each.  Applies when the row strip, if
// any, fits the kernel's tail (tuned_fusable); `epochs` are tuned_fused_epoch_ints() zeroed ints.
bool tuned_fusable(const AffTiling& tiling, int B, int C, int H, int W, int dev) {
    int sm_count = 0;
    if (device_sm_count(dev, &sm_count) != PAMR_OK || tiling.R == 0) return false;
    const long long ntiles = (long long)tiling.tiles_x * tiling.tiles_y * B;
    if (ntiles > 0x7fffffffLL) return false;
    const int grid = ntiles < sm_count ? (int)ntiles : sm_count;
    const long long row_items = (long long)B * C * (H - tiling.Ht) * ((W + 31) / 32);
    if (row_items > 0x7fffffffLL) return false;
    return tiling.Ht == H || row_strip_in_tail(row_items, ntiles, grid);
}
size_t tuned_fused_epoch_ints(const AffTiling& tiling, int B) { return (size_t)tiling.tiles_x * tiling.tiles_y * B + (size_t)B; }
int launch_propagate_tuned_fused(const float* aff_tiled, const AffTiling& tiling, const float* buf0, const float* buf1, float* out,
                                 int B, int C, int H, int W, int iters, unsigned* cls_max, int* epochs, int dev, cudaStream_t s) {
    int sm_count = 0;
    int rc = device_sm_count(dev, &sm_count);
    if (rc != PAMR_OK) return rc;
    if (((uintptr_t)buf0 & 15) != 0 || ((uintptr_t)buf1 & 15) != 0 || ((uintptr_t)aff_tiled & 15) != 0)
        return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: buffer / affinity base not 16-byte aligned");
    const int sp = pair_pitch(W);
    const FusedArgs fa{iters, buf0, buf1, epochs};
    if (tiling.R == 8)
        return launch_one<8>(aff_tiled, tiling, nullptr, sp, out, W, false, B, C, H, W, cls_max, sm_count, dev, false, &fa, s);
    if (tiling.R == 10)
        return launch_one<10>(aff_tiled, tiling, nullptr, sp, out, W, false, B, C, H, W, cls_max, sm_count, dev, false, &fa, s);
    return set_error(PAMR_ERR_INVALID_ARGUMENT, "tuned propagate: unsupported tiling R=%d", tiling.R);
}

}  // namespace pamr
