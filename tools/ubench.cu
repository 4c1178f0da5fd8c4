// Micro-benchmarks that fix the design constants of the propagation kernel on B200:
// shared-memory load bandwidth, FFMA / FFMA2 issue rate with the kernel's operand pattern,
// TMEM (tcgen05.ld) read bandwidth alone and concurrently with LDS, L2-hit and HBM read bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench.bin tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static int g_sms = 148;
static float g_mhz = 1965.f;

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

// ---------------------------------------------------------------- LDS bandwidth
template <int VEC>
__global__ void lds_kernel(float* out, int iters, long long* cyc) {
    extern __shared__ float sm[];
    const int n = 8192;  // floats
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int base = (threadIdx.x * VEC) & (n - 1);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = (base + j * 1056) & (n - 1);
            if (VEC == 1) acc[j] += sm[idx];
            else if (VEC == 2) { float2 v = *reinterpret_cast<float2*>(sm + idx); acc[j] += v.x + v.y; }
            else { float4 v = *reinterpret_cast<float4*>(sm + idx); acc[j] += v.x + v.y + v.z + v.w; }
        }
        base = (base + 32 * VEC) & (n - 1);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 8; ++j) s += acc[j];
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// ---------------------------------------------------------------- FFMA with the stencil's operand pattern
__global__ void ffma_kernel(float* out, int iters, long long* cyc) {
    float acc[56];
#pragma unroll
    for (int j = 0; j < 56; ++j) acc[j] = (float)j;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = (float)(threadIdx.x + j) * 1e-3f;
    float w = 1.0001f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 56; ++j) acc[j] = fmaf(w, m[j & 7], acc[j]);
        w += 1e-7f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 56; ++j) s += acc[j];
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void ffma2_kernel(float* out, int iters, long long* cyc) {
    unsigned long long acc[28];
#pragma unroll
    for (int j = 0; j < 28; ++j) acc[j] = (unsigned long long)j * 0x3f80000000000001ull;
    unsigned long long m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = 0x3a83126f3a83126full + j + threadIdx.x;
    unsigned long long w = 0x3f8003473f800347ull;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 28; ++j)
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(w), "l"(m[j & 3]));
        w += 1;
    }
    long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 28; ++j) s ^= acc[j];
    if (s == 123456ull) out[0] = 1.f;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// ---------------------------------------------------------------- TMEM read bandwidth (tcgen05.ld), optional concurrent LDS
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}

// MODE 0: LDTM only; 1: LDTM + LDS interleaved; 2: LDS only (same structure); block = 128 threads
template <int MODE>
__global__ void __launch_bounds__(128) tmem_kernel(float* out, int iters, long long* cyc, int* check) {
    __shared__ uint32_t tmem_base_s;
    __shared__ float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base_s + ((uint32_t)(warp * 32) << 16);
    // fill: column c of lane l = l*1000 + c
    for (int c0 = 0; c0 < 512; c0 += 16) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = threadIdx.x * 1000 + c0 + j;
        tmem_st16(tbase + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    // correctness probe
    {
        uint32_t v[16];
        tmem_ld16(tbase + 48, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        if (v[5] != threadIdx.x * 1000 + 53) atomicAdd(check, 1);
    }
    uint32_t accu = 0;
    float accf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int base = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c0 = 0; c0 < 512; c0 += 64) {
            if (MODE != 2) {
                uint32_t a[16], b[16], c[16], d[16];
                tmem_ld16(tbase + c0, a);
                tmem_ld16(tbase + c0 + 16, b);
                tmem_ld16(tbase + c0 + 32, c);
                tmem_ld16(tbase + c0 + 48, d);
                if (MODE == 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) accf[j] += sm[(base + j * 1056) & 8191];
#pragma unroll
                    for (int j = 0; j < 8; ++j) accf[j] += sm[(base + 32 + j * 1056) & 8191];
                    base = (base + 64) & 8191;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
                for (int j = 0; j < 16; ++j) accu += a[j] ^ b[j] ^ c[j] ^ d[j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) accf[j] += sm[(base + j * 1056) & 8191];
#pragma unroll
                for (int j = 0; j < 8; ++j) accf[j] += sm[(base + 32 + j * 1056) & 8191];
                base = (base + 64) & 8191;
            }
        }
    }
    long long t1 = clock64();
    float s = (float)accu;
    for (int j = 0; j < 8; ++j) s += accf[j];
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(512));
}

// ---------------------------------------------------------------- global read bandwidth (L2-resident or HBM)
__global__ void read_kernel(const float4* __restrict__ p, size_t n4, int reps, float* out) {
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = __ldg(p + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    g_mhz = khz / 1000.f;
    printf("device %s  SMs %d  max clock %.0f MHz  L2 %d MB  smem/SM %zu KB\n", prop.name, g_sms, g_mhz,
           prop.l2CacheSize >> 20, prop.sharedMemPerMultiprocessor >> 10);
    float* out; long long* cyc; int* check;
    CK(cudaMalloc(&out, 1024)); CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&check, 4));
    CK(cudaMemset(check, 0, 4));
    long long hc = 0;

    // LDS
    {
        const int iters = 4000;
        for (int threads : {128, 256, 512, 1024}) {
            for (int vec : {1, 2, 4}) {
                auto f = [&]() {
                    if (vec == 1) lds_kernel<1><<<g_sms, threads, 32768>>>(out, iters, cyc);
                    else if (vec == 2) lds_kernel<2><<<g_sms, threads, 32768>>>(out, iters, cyc);
                    else lds_kernel<4><<<g_sms, threads, 32768>>>(out, iters, cyc);
                };
                float ms = time_ms(f);
                CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
                double bytes = (double)threads * iters * 8 * 4 * vec;
                printf("LDS.%-3d threads/SM %4d : %6.1f B/clk/SM (clock64)  %7.2f TB/s chip (events)  eff clock %.0f MHz\n",
                       32 * vec, threads, bytes / hc, bytes * g_sms / (ms * 1e-3) / 1e12, hc / (ms * 1e-3) / 1e6);
            }
        }
    }
    // FFMA
    {
        const int iters = 20000;
        for (int threads : {128, 256, 512}) {
            float ms = time_ms([&]() { ffma_kernel<<<g_sms, threads>>>(out, iters, cyc); });
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            double fmas = (double)threads * iters * 56;
            printf("FFMA   threads/SM %4d : %6.1f FMA/clk/SM  %6.2f TFMA/s chip  eff clock %.0f MHz\n", threads, fmas / hc,
                   fmas * g_sms / (ms * 1e-3) / 1e12, hc / (ms * 1e-3) / 1e6);
            ms = time_ms([&]() { ffma2_kernel<<<g_sms, threads>>>(out, iters, cyc); });
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            printf("FFMA2  threads/SM %4d : %6.1f FMA/clk/SM  %6.2f TFMA/s chip  eff clock %.0f MHz\n", threads, fmas / hc,
                   fmas * g_sms / (ms * 1e-3) / 1e12, hc / (ms * 1e-3) / 1e6);
        }
    }
    // TMEM
    {
        const int iters = 2000;
        const char* names[3] = {"LDTM only      ", "LDTM + LDS     ", "LDS only (same)"};
        for (int mode = 0; mode < 3; ++mode) {
            auto f = [&]() {
                if (mode == 0) tmem_kernel<0><<<g_sms, 128>>>(out, iters, cyc, check);
                else if (mode == 1) tmem_kernel<1><<<g_sms, 128>>>(out, iters, cyc, check);
                else tmem_kernel<2><<<g_sms, 128>>>(out, iters, cyc, check);
            };
            float ms = time_ms(f, 3);
            CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
            double tm_bytes = (mode != 2) ? 128.0 * iters * 512 * 4 : 0;
            double lds_bytes = (mode != 0) ? 128.0 * iters * 8 * 16 * 4 : 0;
            printf("%s 128 thr/SM: TMEM %6.1f B/clk/SM  LDS %6.1f B/clk/SM  (%.3f ms)\n", names[mode], tm_bytes / hc,
                   lds_bytes / hc, ms);
        }
        int hcheck = 0;
        CK(cudaMemcpy(&hcheck, check, 4, cudaMemcpyDeviceToHost));
        printf("TMEM st/ld round-trip mismatches: %d\n", hcheck);
    }
    // global reads
    {
        size_t big = (size_t)4 << 30;
        float4* buf;
        CK(cudaMalloc(&buf, big));
        CK(cudaMemset(buf, 0, big));
        for (size_t mb : {16, 32, 64, 96, 4096}) {
            size_t n4 = mb * 1024 * 1024 / 16;
            int reps = mb >= 4096 ? 2 : 200;
            for (int mult : {4, 8}) {
                float ms = time_ms([&]() { read_kernel<<<g_sms * mult, 512>>>(buf, n4, reps, out); }, 3);
                printf("global read  %5zu MB x%3d  grid %4d x512: %8.1f GB/s\n", mb, reps, g_sms * mult,
                       (double)n4 * 16 * reps / (ms * 1e-3) / 1e9);
            }
        }
        CK(cudaFree(buf));
    }
    return 0;
}
