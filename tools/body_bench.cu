// Isolated throughput of the propagation kernel's compute pass (no TMA, no mbarriers, no fill):
// how many cycles does one pass over the 48 taps take per SM with 4 / 8 / 12 compute warps?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I 1-stage-wseg_b200/csrc \
//        -o tools/body_bench.bin tools/body_bench.cu
#include "../1-stage-wseg_b200/csrc/pamr_propagate_sm100.cu"

namespace pamr {
int set_error(int code, const char*, ...) { return code; }
void count_launch(int) {}
int launch_aff_relayout(const float*, float*, int, int, int, const AffTiling&, cudaStream_t) { return 0; }
namespace {

template <int R, int N>
__global__ void __launch_bounds__(416, 1) body_kernel(float* out, int passes, long long* cyc, int nwarps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* slots = reinterpret_cast<float*>(smem_raw);
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NSLOT * Cfg<R>::SLOT_FLOATS; i += blockDim.x) slots[i] = (float)(i & 1023) * 1e-3f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int wq = warp & 3, grp = warp >> 2;
    const uint32_t tbase = tmem_base_s + ((uint32_t)(wq * 32) << 16);
    if (warp < 4) {
        for (int c = 0; c < 512; c += 8) {
            float r[8];
            for (int j = 0; j < 8; ++j) r[j] = 1.0f / 48.f + 1e-5f * (float)(c + j + lane);
            tmem_st8(tbase + c, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float total = 0.f;
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int p = 0; p < passes; ++p) {
            const float* sp[CC];
            for (int j = 0; j < CC; ++j)
                sp[j] = slots + (size_t)((p * N + j + grp * 2) % NSLOT) * Cfg<R>::SLOT_FLOATS + (wq * R + HALO) * WIN_W + lane + HALO;
            float acc[CC][R];
            for (int j = 0; j < CC; ++j)
                for (int i = 0; i < R; ++i) acc[j][i] = 0.f;
            compute_pass<R, N>(sp, tbase, acc, 64 + lane, 1 << 20);
            for (int j = 0; j < N; ++j)
                for (int i = 0; i < R; ++i) total += acc[j][i];
        }
    }
    long long t1 = clock64();
    if (total == 12345.f) out[0] = total;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512));
}

}  // namespace
}  // namespace pamr

template <int R, int N>
void run(float* out, long long* cyc) {
    using namespace pamr;
    const int passes = 200;
    cudaFuncSetAttribute(body_kernel<R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<R>::SMEM_BYTES);
    for (int nw : {4, 8, 12}) {
        body_kernel<R, N><<<148, 416, Cfg<R>::SMEM_BYTES>>>(out, passes, cyc, nw);
        cudaDeviceSynchronize();
        body_kernel<R, N><<<148, 416, Cfg<R>::SMEM_BYTES>>>(out, passes, cyc, nw);
        cudaError_t e = cudaDeviceSynchronize();
        long long hc = 0; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_round = (double)hc / passes, planes = (double)N * nw / 4;
        printf("R=%d N=%d warps %2d: %8.0f cycles per round (%s) -> %6.0f cycles per class-plane per SM (LDS-bound %d), LDS pipe %.0f%%\n",
               R, N, nw, per_round, cudaGetErrorString(e), per_round / planes, R == 10 ? 1256 : R == 8 ? 1072 : 1164,
               100.0 * planes * (R == 10 ? 1256 : R == 8 ? 1072 : 1164) / per_round);
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    run<10, 1>(out, cyc);
#if PAMR_CC >= 2
    run<10, 2>(out, cyc);
#endif
#if PAMR_CC >= 3
    run<10, 3>(out, cyc);
#endif
    run<8, 1>(out, cyc);
    return 0;
}
