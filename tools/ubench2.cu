// TMEM read bandwidth (tcgen05.ld 32x32b.xN), alone and overlapped with LDS + FFMA, 4 or 8 warps per SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int N> struct Regs { uint32_t r[N]; };

__device__ __forceinline__ void ld16(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(t) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(t) : "memory");
}
__device__ __forceinline__ void st16(uint32_t t, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(t), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: LDTM.x16 only  1: LDTM.x32 only  2: x16 + FFMA consume (weights*const)  3: x16 + LDS + FFMA (stencil-like mix)
// 4: LDS + FFMA only (same mix, weights from registers)
// NW warps; warps w and w+4 share a lane quarter and use different column halves.
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32) k(float* out, int iters, long long* cyc) {
    __shared__ uint32_t tb;
    __shared__ float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)(i & 255) * 1e-3f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tb);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int ncol = (NW == 8) ? 256 : 512;
    const uint32_t tbase = tb + ((uint32_t)((warp & 3) * 32) << 16) + ((NW == 8 && warp >= 4) ? 256 : 0);
    for (int c0 = 0; c0 < ncol; c0 += 16) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(1.0f + 1e-4f * (float)(c0 + j + lane));
        st16(tbase + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    uint32_t sum = 0;
    int base = lane + warp * 96;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int c0 = 0; c0 < ncol; c0 += 32) {
            uint32_t w[32];
            if (MODE == 1) ld32(tbase + c0, w);
            else if (MODE != 4) { ld16(tbase + c0, w); ld16(tbase + c0 + 16, w + 16); }
            float m[24];
            if (MODE >= 3) {
#pragma unroll
                for (int j = 0; j < 24; ++j) m[j] = sm[(base + j * 80) & 8191];  // 24 LDS per 32 weights (~0.65 LDS/FMA)
                base = (base + 7) & 8191;
            }
            if (MODE != 4) wait_ld();
            if (MODE <= 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sum += w[j];
            } else if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j & 15] = fmaf(__uint_as_float(w[j]), 1.0001f, acc[j & 15]);
            } else if (MODE == 3) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j & 15] = fmaf(__uint_as_float(w[j]), m[(j * 7) % 24], acc[j & 15]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j & 15] = fmaf(1.0001f + acc[(j + 1) & 15] * 0.f, m[(j * 7) % 24], acc[j & 15]);
            }
        }
    }
    long long t1 = clock64();
    float s = (float)sum;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 12345.0f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}

template <int MODE, int NW>
void run(const char* name, float* out, long long* cyc) {
    const int iters = 2000;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k<MODE, NW><<<148, NW * 32>>>(out, iters, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k<MODE, NW><<<148, NW * 32>>>(out, iters, cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    long long hc; CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    const int ncol = (NW == 8) ? 256 : 512;
    double tm = (MODE != 4) ? (double)NW * 32 * iters * ncol * 4 : 0;
    double lds = (MODE >= 3) ? (double)NW * 32 * iters * (ncol / 32) * 24 * 4 : 0;
    double fma = (MODE >= 2) ? (double)NW * 32 * iters * ncol : 0;
    printf("%-34s warps %d: TMEM %7.1f B/clk/SM  LDS %6.1f B/clk/SM  FMA %6.1f /clk/SM  (%.3f ms, %lld clk)\n", name, NW,
           tm / hc, lds / hc, fma / hc, ms, hc);
}

int main() {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 1024)); CK(cudaMalloc(&cyc, 64));
    run<0, 4>("LDTM.x16 only", out, cyc);   run<0, 8>("LDTM.x16 only", out, cyc);
    run<1, 4>("LDTM.x32 only", out, cyc);   run<1, 8>("LDTM.x32 only", out, cyc);
    run<2, 4>("LDTM.x16 + FFMA", out, cyc); run<2, 8>("LDTM.x16 + FFMA", out, cyc);
    run<3, 4>("LDTM.x16 + LDS + FFMA", out, cyc); run<3, 8>("LDTM.x16 + LDS + FFMA", out, cyc);
    run<4, 4>("LDS + FFMA (no TMEM)", out, cyc);  run<4, 8>("LDS + FFMA (no TMEM)", out, cyc);
    return 0;
}
