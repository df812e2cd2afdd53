// How fast does one SM retire back-to-back tcgen05.mma (kind::f16, M = 128, K = 16, both operands K-major / SWIZZLE_NONE in
// shared memory) as a function of N?  The fused tail (k_tz_tail) issues 27 such MMAs per 128-row tile with N = 48 .. 128; this
// measures what one of them costs.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ofighters_b200/csrc scripts/mma_pace.cu -o /tmp/mma_pace
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ofb_tc_ptx.cuh"

__global__ void __launch_bounds__(128, 1) k_pace(int n, int n_mma, int lbo_a, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        const bool leader = elect_one();
        const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 128 * 1024) >> 4;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long best = 1ll << 60;
        for (int rep = 0; rep < 5; rep++) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // A: a different 4 KB (LBO = 128 rows) or strided (LBO = lbo_a) operand every MMA; B: [2 chunks][n rows][16 B]
                    const uint64_t ad = smem_desc(a16 + (uint32_t)(((i + j) & 7) * 256), (uint32_t)lbo_a, 8);
                    const uint64_t bd = smem_desc(b16 + (uint32_t)(((i + j) & 3) * 2 * n), (uint32_t)n, 8);
                    if (leader) tc_mma(tmem, ad, bd, idesc, (i + j) ? 1u : 0u);
                }
            }
            if (leader) tc_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, (uint32_t)(rep & 1));
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (threadIdx.x == 0) out[blockIdx.x] = best;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

// The same with CTA pairs (cta_group::2): M = 256 over the two SMs of a cluster, each CTA holding its own 128 A rows and N / 2 of the
// B columns; the leader issues, the commit is multicast to both CTAs' barriers.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_pace2(int n, int n_mma, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        const bool leader = elect_one() && rank == 0;
        const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 128 * 1024) >> 4;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        long long best = 1ll << 60;
        for (int rep = 0; rep < 5; rep++) {
            const long long t0 = clock64();
            if (rank == 0) {
#pragma unroll 1
                for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint64_t ad = smem_desc(a16 + (uint32_t)(((i + j) & 7) * 256), 128u, 8);
                        const uint64_t bd = smem_desc(b16 + (uint32_t)(((i + j) & 3) * n), (uint32_t)(n / 2), 8);
                        if (leader)
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc),
                                "r"((uint32_t)((i + j) ? 1 : 0)) : "memory");
                    }
                }
                if (leader)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                                     smem_u32(&bar)), "h"((uint16_t)3) : "memory");
                __syncwarp();
            }
            mbar_wait(&bar, (uint32_t)(rep & 1));
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (threadIdx.x == 0) out[blockIdx.x] = best;
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

int main() {
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k_pace, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int n_mma = 256;
    printf("cycles per tcgen05.mma (M = 128, K = 16, bf16, SWIZZLE_NONE K-major operands in shared memory), %d back to back, all 148 SMs busy\n", n_mma);
    for (int lbo : {128, 448}) {
        for (int n : {16, 32, 48, 64, 80, 96, 128, 192, 256}) {
            k_pace<<<148, 128, 200 * 1024>>>(n, n_mma, lbo, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mn = h[0], mx = h[0];
            for (int i = 1; i < 148; i++) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
            printf("A LBO %3d rows  N = %3d : %6.1f cycles per MMA (CTA min) %6.1f (CTA max)   [math floor N/2 = %d]\n", lbo, n, (double)mn / n_mma,
                   (double)mx / n_mma, n / 2);
        }
    }
    cudaFuncSetAttribute(k_pace2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("CTA pairs (cta_group::2, M = 256 = 128 rows per SM, each CTA holds N / 2 columns of B), 74 clusters\n");
    for (int n : {32, 48, 64, 80, 96, 128, 192, 256}) {
        k_pace2<<<148, 128, 200 * 1024>>>(n, n_mma, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mn = h[0], mx = h[0];
        for (int i = 2; i < 148; i += 2) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
        printf("pair  N = %3d : %6.1f cycles per MMA (leader min) %6.1f (leader max)   [math floor N/2 = %d]\n", n, (double)mn / n_mma, (double)mx / n_mma,
               n / 2);
    }
    return 0;
}
