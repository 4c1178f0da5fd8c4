// FFMA issue cost vs operand pattern: distinct (w,v,acc) per FMA vs shared multiplicand; scalar vs packed f32x2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc, float seed) {
    float w[32], v[32], acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { w[j] = seed + j; v[j] = seed * (float)(threadIdx.x + j); acc[j] = 0.f; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // all distinct
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(w[j], v[(j + 1) & 31], acc[j]);
        } else if (MODE == 1) {  // shared multiplicand over 2 (like N=2 classes)
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(w[j >> 1], v[(j + 1) & 31], acc[j]);
        } else if (MODE == 2) {  // shared over 4
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(w[j >> 2], v[(j + 1) & 31], acc[j]);
        } else {  // packed f32x2, all distinct pairs
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                unsigned long long a, b, c;
                asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(w[j]), "f"(w[j + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(v[(j + 2) & 31]), "f"(v[(j + 3) & 31]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc[j]), "f"(acc[j + 1]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[j]), "=f"(acc[j + 1]) : "l"(c));
            }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) v[j] += 1e-9f;  // keep the loop from being folded
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += acc[j];
    if (s == 12345.f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    const int iters = 20000;
    const char* names[4] = {"FFMA all-distinct operands", "FFMA multiplicand shared x2", "FFMA multiplicand shared x4", "FFMA2 (f32x2) all-distinct"};
    for (int mode = 0; mode < 4; ++mode)
        for (int threads : {128, 256, 512}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(out, iters, cyc, 1.5f);
                if (mode == 1) k<1><<<148, threads>>>(out, iters, cyc, 1.5f);
                if (mode == 2) k<2><<<148, threads>>>(out, iters, cyc, 1.5f);
                if (mode == 3) k<3><<<148, threads>>>(out, iters, cyc, 1.5f);
                cudaDeviceSynchronize();
            }
            long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-30s warps/SMSP %d: %6.1f FMA/clk/SM  (%.2f cycles per warp-FFMA-equivalent per SMSP)\n", names[mode], threads / 128,
                   (double)threads * iters * 32 / hc, (double)hc / ((double)iters * 32 * (threads / 128)));
        }
    return 0;
}
