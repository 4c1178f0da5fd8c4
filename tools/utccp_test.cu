// Micro-test: semantics of tcgen05.cp (shared memory -> Tensor Memory) with a no-swizzle descriptor.
// Fills shared memory with its own float index, copies with tcgen05.cp.128x128b / 128x256b, reads TMEM back
// with tcgen05.ld.32x32b and prints which shared-memory word landed in (lane, column).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/utccp_test.bin tools/utccp_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // layout_type (bits 61-63) = 0: no swizzle
}

// mode 0: 128x128b (4 columns), mode 1: 128x256b (8 columns)
__global__ void __launch_bounds__(128) utccp_kernel(float* out, int mode, uint32_t lbo, uint32_t sbo, long long* cyc) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned long long bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = (float)i;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    // clear the columns first
    {
        const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 64; c += 1) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(ta + c), "f"(-1.0f));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of sm[] -> async proxy
        const uint64_t desc = make_desc(smem_u32(sm), lbo, sbo);
        t0 = clock64();
        if (mode == 0) {
            asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tb), "l"(desc) : "memory");
        } else if (mode == 1) {
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb), "l"(desc) : "memory");
        } else {  // throughput: 8 x 128x256b = 128 lanes x 64 columns = 32 KB
            for (int k = 0; k < 8; ++k) {
                const uint64_t dk = make_desc(smem_u32(sm) + k * 4096, lbo, sbo);
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb + k * 8), "l"(dk) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
    }
    if (threadIdx.x == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 64; ++c) {
            float v;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v) : "r"(ta + c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            out[(warp * 32 + lane) * 64 + c] = v;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 128 * 64 * 4); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(utccp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    static float h[128 * 64];
    struct { int mode; uint32_t lbo, sbo; } cases[] = {{0, 0, 128}, {0, 16, 128}, {0, 128, 256}, {1, 128, 256}, {1, 2048, 128}, {1, 16, 32}, {2, 2048, 128}};
    for (auto& cs : cases) {
        cudaMemset(out, 0, sizeof(h));
        utccp_kernel<<<1, 128, 65536>>>(out, cs.mode, cs.lbo, cs.sbo, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long hc = 0;
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d lbo %u sbo %u: %s, %lld cycles issue->complete\n", cs.mode, cs.lbo, cs.sbo, cudaGetErrorString(e), hc);
        if (e != cudaSuccess) return 1;
        const int lanes[] = {0, 1, 7, 8, 9, 16, 31, 32, 33, 64, 127};
        for (int l : lanes) {
            printf("  lane %3d:", l);
            for (int c = 0; c < 10; ++c) printf(" %6.0f", h[l * 64 + c]);
            printf("\n");
        }
    }
    return 0;
}
