// Micro-benchmark (GPU box): issue rate of tcgen05.mma.cta_group::2 kind::f16 for M=128/256 (pair), various N.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(int M, int N, int n_mma, long long* out, int grp, int mcast) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t dummy[8];
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = make_idesc(M, N);
        const uint64_t ad = make_desc(smem_u32(smem), 16, 1024);
        const uint64_t bd = make_desc(smem_u32(smem) + 32768, 16, 1024);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            if (leader) {
                uint32_t acc = i > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (uint32_t)((i & 1) * 128)), "l"(ad + 2 * (i & 3)), "l"(bd + 2 * (i & 3)), "r"(idesc), "r"(acc) : "memory");
                if (grp > 0 && (i & (grp - 1)) == grp - 1) {
                    if (mcast) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&dummy[(i >> 2) & 7])), "h"((uint16_t)3) : "memory");
                    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[(i >> 2) & 7])) : "memory");
                }
            }
            __syncwarp();
        }
        if (leader) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0 && leader) out[0] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* out; CK(cudaMallocManaged(&out, 32));
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    const int n = 4096;
    struct { int M, N, grp, mc; } cfgs[] = {{128, 256, 0, 0}, {128, 256, 8, 1}, {128, 256, 4, 1}, {128, 256, 4, 0}, {128, 256, 1, 1}, {128, 256, 1, 0}, {128, 64, 1, 1}, {128, 32, 1, 1}};
    for (auto c : cfgs) {
        probe2<<<148, 128, 96 * 1024>>>(c.M, c.N, n, out, c.grp, c.mc);
        CK(cudaDeviceSynchronize());
        printf("cta_group::2  M=%3d N=%3d commit every %d MMAs (%s): %.1f cyc/MMA\n", c.M, c.N, c.grp, c.mc ? "multicast" : "local", (double)out[0] / n);
    }
    return 0;
}
