// Micro-benchmark (GPU box): issue rate of tcgen05.mma kind::f16, M=128, K=16, for different N and accumulator patterns.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// mode: 0 = all MMAs accumulate into one D; 1 = round-robin over `nd` D tiles (independent chains)
__global__ void __launch_bounds__(128, 1) probe(int N, int nd, int n_mma, int bmn, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        // converged warp, one elected lane issues: lets the compiler keep descriptors in uniform registers
        const uint32_t idesc = make_idesc(128, N, bmn);
        const uint64_t ad = make_desc(smem_u32(smem), 16, 1024);
        const uint64_t bd = make_desc(smem_u32(smem) + 16384, bmn ? 0 : 16, 1024);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t d = tmem + (uint32_t)((i & (nd - 1)) * N);
            if (leader) umma(d, ad + 2 * (i & 3), bd + (bmn ? 128 * (i & 3) : 2 * (i & 3)), idesc, i >= nd);
            __syncwarp();
        }
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0 && leader) { out[2] = t1 - t0; out[3] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t idesc = make_idesc(128, N, bmn);
        const uint64_t ad = make_desc(smem_u32(smem), 16, 1024);
        const uint64_t bd = make_desc(smem_u32(smem) + 16384, bmn ? 0 : 16, 1024);
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t d = tmem + (uint32_t)((i & (nd - 1)) * N);
            umma(d, ad + 2 * (i & 3), bd + (bmn ? 128 * (i & 3) : 2 * (i & 3)), idesc, i >= nd);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* out; CK(cudaMallocManaged(&out, 32));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int n = 4096;
    struct { int N, nd, bmn; } cfgs[] = {{64, 1, 1}, {64, 2, 1}, {64, 4, 1}, {64, 8, 1}, {128, 1, 1}, {128, 2, 1}, {128, 4, 1}, {256, 1, 1}, {256, 2, 1},
                                         {64, 1, 0}, {64, 4, 0}, {32, 1, 1}, {32, 4, 1}, {16, 1, 1}, {16, 8, 1}};
    for (auto c : cfgs) {
        for (int grid : {148}) {
            probe<<<grid, 128, 64 * 1024>>>(c.N, c.nd, n, c.bmn, out);
            CK(cudaDeviceSynchronize());
            printf("N=%3d  %d accumulator(s)  B %s-major  grid %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)\n", c.N, c.nd, c.bmn ? "MN" : "K ", grid,
                   (double)out[0] / n, (double)out[1] / n, c.N / 2);
            printf("        elected-lane/converged-warp variant:             issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", (double)out[2] / n, (double)out[3] / n);
        }
    }
    return 0;
}
