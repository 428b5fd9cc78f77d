// Micro-benchmark (GPU box): how fast can every SM stream the SAME L2-resident weight buffer into shared memory?
// Variants: 1-D bulk TMA with different stage sizes / depths / cluster multicast, 2-D tensor-map TMA, cp.async, LDG+STS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/bw_probe.cu -o gpurun_out/bw_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// ---- 1-D bulk TMA ring: producer thread + consumer thread (consumer just releases the stage)
template <int CL>
__global__ void __launch_bounds__(256, 1) bulk_ring(const uint8_t* __restrict__ src, size_t bytes, int stage_bytes, int nst, int iters, int pieces) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bars = base + (uint32_t)nst * stage_bytes;
    const uint32_t rank = CL > 1 ? ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) { mbar_init(bars + 16 * i, 1); mbar_init(bars + 16 * i + 8, CL); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int n_stage = (int)(bytes / stage_bytes);
    const long long total = (long long)n_stage * iters;
    if (threadIdx.x == 0) {
        const uint32_t slice = stage_bytes / CL, piece = slice / pieces;
        for (long long u = 0; u < total; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            mbar_wait(bars + 16 * st + 8, ph ^ 1);
            mbar_expect(bars + 16 * st, stage_bytes);
            const uint8_t* g = src + (size_t)(u % n_stage) * stage_bytes + rank * slice;
            const uint32_t d = base + st * stage_bytes + rank * slice;
            for (int p = 0; p < pieces; ++p) {
                if (CL == 1)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d + p * piece), "l"(g + p * piece), "r"(piece), "r"(bars + 16 * st) : "memory");
                else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(d + p * piece), "l"(g + p * piece), "r"(piece), "r"(bars + 16 * st), "h"((uint16_t)((1u << CL) - 1)) : "memory");
            }
        }
    } else if (threadIdx.x == 32) {
        for (long long u = 0; u < total; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            mbar_wait(bars + 16 * st, ph);
            if (CL == 1) mbar_arrive(bars + 16 * st + 8);
            else for (uint32_t c = 0; c < CL; ++c) mbar_arrive_cluster(bars + 16 * st + 8, c);
        }
    }
    __syncthreads();
    if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- single-thread ring: the producer itself waits for the slot's previous copy (no consumer thread hand-shake)
template <int POLL>
__global__ void __launch_bounds__(256, 1) bulk_self(const uint8_t* __restrict__ src, size_t bytes, int stage_bytes, int nst, int iters, int pieces) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bars = base + (uint32_t)nst * stage_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(bars + 16 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_stage = (int)(bytes / stage_bytes);
    const long long total = (long long)n_stage * iters;
    if (threadIdx.x == 0) {
        for (long long u = 0; u < total + nst; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            if (u >= nst) {
                if (POLL) {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bars + 16 * st), "r"(ph ^ 1) : "memory");
                } else mbar_wait(bars + 16 * st, ph ^ 1);
            }
            if (u < total) {
                mbar_expect(bars + 16 * st, stage_bytes);
                size_t off;
                if (pieces == 2) off = (size_t)((u + 37ull * blockIdx.x) % n_stage) * stage_bytes;                  // phase-shifted per CTA
                else if (pieces == 3) off = (size_t)blockIdx.x * 65536 + (size_t)(u % (65536 / stage_bytes)) * stage_bytes;  // private 64 KiB region
                else off = (size_t)(u % n_stage) * stage_bytes;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + st * stage_bytes), "l"(src + off), "r"(stage_bytes), "r"(bars + 16 * st) : "memory");
            }
        }
    }
    __syncthreads();
}

// ---- several TMA producer threads (one per warp), each owning stages st % NP == p
__global__ void __launch_bounds__(256, 1) bulk_multi(const uint8_t* __restrict__ src, size_t bytes, int stage_bytes, int nst, int iters, int np) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bars = base + (uint32_t)nst * stage_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(bars + 16 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_stage = (int)(bytes / stage_bytes);
    const long long total = (long long)n_stage * iters;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < np) {
        for (long long u = w; u < total + nst; u += np) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            if (u >= nst) mbar_wait(bars + 16 * st, ph ^ 1);
            if (u < total) {
                mbar_expect(bars + 16 * st, stage_bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + st * stage_bytes), "l"(src + (size_t)(u % n_stage) * stage_bytes), "r"(stage_bytes), "r"(bars + 16 * st) : "memory");
            }
        }
    }
    __syncthreads();
}

// ---- cp.async (LDGSTS) loader threads + mbarrier arrive.noinc, one consumer thread releasing stages
__global__ void __launch_bounds__(256, 1) cpasync_ring(const uint8_t* __restrict__ src, size_t bytes, int stage_bytes, int nst, int iters, int nload) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bars = base + (uint32_t)nst * stage_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) { mbar_init(bars + 16 * i, nload); mbar_init(bars + 16 * i + 8, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_stage = (int)(bytes / stage_bytes);
    const long long total = (long long)n_stage * iters;
    if ((int)threadIdx.x < nload) {
        for (long long u = 0; u < total; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            mbar_wait(bars + 16 * st + 8, ph ^ 1);
            const uint8_t* g = src + (size_t)(u % n_stage) * stage_bytes;
            const uint32_t d = base + st * stage_bytes;
            for (int o = threadIdx.x * 16; o < stage_bytes; o += nload * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + o), "l"(g + o) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bars + 16 * st) : "memory");
        }
    } else if (threadIdx.x == 255) {
        for (long long u = 0; u < total; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            mbar_wait(bars + 16 * st, ph);
            mbar_arrive(bars + 16 * st + 8);
        }
    }
    __syncthreads();
}

// ---- 2-D tensor-map TMA ring (box = 64 x rows of bf16, 128-byte rows, SWIZZLE_128B like a GEMM operand load)
__global__ void __launch_bounds__(256, 1) tmap_ring(const __grid_constant__ CUtensorMap tm, int n_stage, int stage_bytes, int nst, int iters, int box_rows) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bars = base + (uint32_t)nst * stage_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) { mbar_init(bars + 16 * i, 1); mbar_init(bars + 16 * i + 8, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long total = (long long)n_stage * iters;
    if (threadIdx.x == 0) {
        for (long long u = 0; u < total; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            mbar_wait(bars + 16 * st + 8, ph ^ 1);
            mbar_expect(bars + 16 * st, stage_bytes);
            const int row = (int)(u % n_stage) * box_rows;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(base + st * stage_bytes), "l"(&tm), "r"(0), "r"(row), "r"(bars + 16 * st) : "memory");
        }
    } else if (threadIdx.x == 32) {
        for (long long u = 0; u < total; ++u) {
            const uint32_t st = u % nst, ph = (u / nst) & 1;
            mbar_wait(bars + 16 * st, ph);
            mbar_arrive(bars + 16 * st + 8);
        }
    }
    __syncthreads();
}

// ---- generic paths: all threads copy global -> smem
template <int MODE>   // 0: LDG.128 + STS.128, 1: cp.async 16 B
__global__ void __launch_bounds__(256, 1) generic_copy(const uint8_t* __restrict__ src, size_t bytes, int chunk, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int n_chunk = (int)(bytes / chunk);
    for (int it = 0; it < iters; ++it)
        for (int c = 0; c < n_chunk; ++c) {
            const uint8_t* g = src + (size_t)c * chunk;
            uint8_t* d = smem + (c & 1) * chunk;
            if (MODE == 0) {
                for (int o = threadIdx.x * 16; o < chunk; o += 256 * 16 * 4) {
                    uint4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (o + j * 4096 < chunk) v[j] = __ldg((const uint4*)(g + o + j * 4096));
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (o + j * 4096 < chunk) *(uint4*)(d + o + j * 4096) = v[j];
                }
            } else {
                for (int o = threadIdx.x * 16; o < chunk; o += 256 * 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + o)), "l"(g + o) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
        }
    if (MODE == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const size_t bytes = 292ull * 2 * 16384;   // the PRE kernel's weight stream: 9.57 MB
    uint8_t* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20;
    auto report = [&](const char* name, float ms, int grid) {
        const double per_sm = (double)bytes * iters;
        printf("%-52s grid %3d  %8.2f ms  %7.1f GB/s per SM  %6.2f TB/s total  (%5.1f B/clk/SM at %d MHz nominal)\n", name, grid, ms,
               per_sm / ms / 1e6, per_sm * grid / ms / 1e9, per_sm / (ms * 1e-3) / (khz * 1e3), khz / 1000);
    };
    int grid_override = 0;
    auto run_bulk = [&](auto kern, int cl, int stage, int nst, int pieces, const char* name) {
        const int smem = stage * nst + 16 * nst + 64;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        cfg.gridDim = dim3(cl * 32);
        CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
        int grid = ncl * cl; if (grid > (sms / cl) * cl) grid = (sms / cl) * cl;
        if (grid_override) grid = grid_override;
        cfg.gridDim = dim3(grid);
        for (int w = 0; w < 2; ++w) {
            if (w == 1) CK(cudaEventRecord(e0));
            CK(cudaLaunchKernelEx(&cfg, kern, (const uint8_t*)buf, bytes, stage, nst, iters, pieces));
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        report(name, ms, grid);
    };
    run_bulk(bulk_ring<1>, 1, 16384, 6, 1, "bulk1d 16K x6 (current parity ring)");
    run_bulk(bulk_ring<1>, 1, 16384, 8, 1, "bulk1d 16K x8 (current fast ring)");
    run_bulk(bulk_ring<1>, 1, 16384, 12, 1, "bulk1d 16K x12");
    run_bulk(bulk_ring<1>, 1, 32768, 6, 1, "bulk1d 32K x6");
    run_bulk(bulk_ring<1>, 1, 65536, 3, 1, "bulk1d 64K x3");
    run_bulk(bulk_ring<1>, 1, 16384, 6, 4, "bulk1d 16K x6, 4 pieces of 4K");
    run_bulk(bulk_ring<1>, 1, 16384, 12, 8, "bulk1d 16K x12, 8 pieces of 2K");
    run_bulk(bulk_ring<1>, 1, 4096, 24, 1, "bulk1d 4K x24");
    run_bulk(bulk_self<0>, 1, 16384, 6, 1, "self-ring try_wait 16K x6");
    run_bulk(bulk_self<1>, 1, 16384, 6, 1, "self-ring test_wait(poll) 16K x6");
    run_bulk(bulk_self<1>, 1, 16384, 12, 1, "self-ring test_wait(poll) 16K x12");
    run_bulk(bulk_self<1>, 1, 8192, 24, 1, "self-ring test_wait(poll) 8K x24");
    run_bulk(bulk_self<1>, 1, 32768, 6, 1, "self-ring test_wait(poll) 32K x6");
    run_bulk(bulk_multi, 1, 16384, 6, 2, "bulk1d 16K x6, 2 producer warps");
    run_bulk(bulk_multi, 1, 16384, 6, 3, "bulk1d 16K x6, 3 producer warps");
    run_bulk(bulk_multi, 1, 16384, 12, 6, "bulk1d 16K x12, 6 producer warps");
    run_bulk(cpasync_ring, 1, 16384, 6, 32, "cp.async ring 16K x6, 32 loader threads");
    run_bulk(cpasync_ring, 1, 16384, 6, 64, "cp.async ring 16K x6, 64 loader threads");
    run_bulk(cpasync_ring, 1, 16384, 6, 96, "cp.async ring 16K x6, 96 loader threads");
    run_bulk(cpasync_ring, 1, 16384, 6, 128, "cp.async ring 16K x6, 128 loader threads");
    run_bulk(cpasync_ring, 1, 16384, 8, 96, "cp.async ring 16K x8, 96 loader threads");
    run_bulk(cpasync_ring, 1, 16384, 3, 96, "cp.async ring 16K x3, 96 loader threads");
    run_bulk(bulk_self<0>, 1, 16384, 6, 2, "self-ring 16K x6 PHASE-SHIFTED per CTA");
    run_bulk(bulk_self<0>, 1, 16384, 6, 3, "self-ring 16K x6 PRIVATE 64K region per CTA");
    for (int g : {1, 8, 32, 74}) { grid_override = g; char nm[64]; snprintf(nm, 64, "self-ring 16K x6 same data, %d CTAs", g); run_bulk(bulk_self<0>, 1, 16384, 6, 1, nm); }
    grid_override = 0;
    run_bulk(bulk_ring<2>, 2, 16384, 6, 1, "bulk1d 16K x6 multicast cluster 2");
    run_bulk(bulk_ring<4>, 4, 16384, 6, 1, "bulk1d 16K x6 multicast cluster 4");
    run_bulk(bulk_ring<4>, 4, 16384, 12, 1, "bulk1d 16K x12 multicast cluster 4");
    run_bulk(bulk_ring<8>, 8, 16384, 12, 1, "bulk1d 16K x12 multicast cluster 8");
    {   // tensor-map TMA
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        EncodeFn enc = (EncodeFn)fn;
        for (int variant = 0; variant < 3; ++variant) {
            const int box_rows = variant == 0 ? 128 : 256, nst = variant == 2 ? 6 : (variant == 0 ? 6 : 3);
            const int stage = box_rows * 128;
            CUtensorMap tm;
            cuuint64_t gdim[2] = {64, bytes / 128};
            cuuint64_t gstr[1] = {128};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)r); break; }
            const int smem = stage * nst + 16 * nst + 64;
            CK(cudaFuncSetAttribute(tmap_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const int n_stage = (int)(bytes / stage);
            for (int w = 0; w < 2; ++w) {
                if (w == 1) CK(cudaEventRecord(e0));
                tmap_ring<<<sms, 128, smem>>>(tm, n_stage, stage, nst, iters, box_rows);
            }
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            char nm[96]; snprintf(nm, sizeof(nm), "tensormap 2D box 64x%d (%dK) x%d", box_rows, stage / 1024, nst);
            report(nm, ms, sms);
        }
    }
    for (int mode = 0; mode < 2; ++mode) {
        const int chunk = 65536, smem = 2 * chunk;
        auto kern = mode == 0 ? generic_copy<0> : generic_copy<1>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int w = 0; w < 2; ++w) {
            if (w == 1) CK(cudaEventRecord(e0));
            kern<<<sms, 256, smem>>>(buf, bytes, chunk, iters);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        report(mode == 0 ? "LDG.128+STS.128, 256 threads" : "cp.async 16B, 256 threads", ms, sms);
    }
    printf("done\n");
    return 0;
}
