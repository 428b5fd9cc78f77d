// Shared pieces of the tcgen05 (5th-gen tensor core) path of the per-sample network: PTX wrappers, the fp16 hi/lo operand
// split, weight / bias packing.  The kernel itself is the CTA-pair kernel in mlp_tc2.cu.
//
// Arithmetic of the PARITY mode (the 1e-4 mode; measured ~5e-7 on rendered rgb / depth, tests/test_gpu_parity.py): every
// product a*w is evaluated as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp16 hi/lo splits of both operands and fp32 accumulation
// (3 MMAs).  fp16 carries 11 significant bits, so hi+lo keeps 22 of fp32's 24 bits and the dropped lo*lo term is 2^-22
// relative -- 30x more accurate than the same three passes in bf16 (8+8 bits) at the same tensor-pipe cost.  fp16's narrow
// exponent range is handled by scaling: weights are packed as W_SCALE * w (a power of two: exact), so that their lo parts
// stay in the normal range, every accumulator in TMEM therefore holds W_SCALE * value, and the epilogues fold 1/W_SCALE
// into the multiply-add they already do; conversions saturate (cvt.satfinite) instead of producing inf.
// FAST mode is a single fp16 pass.
//
// Reference semantics: src/models/resnetfc.py:61-69,129-159; src/models/pixelnerf.py:91-143.
#include "mlp_tc.h"

#include <cuda_fp16.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#define TCK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) return _e; } while (0)

namespace tc {

constexpr int HID = 512;                    // d_hidden served by this path
constexpr int MT = HID / 128;               // 128-row weight tiles per layer
constexpr int KBLK = 64;                    // K elements per weight tile (one 128-byte swizzle atom)
constexpr int WTILE_BYTES = 128 * KBLK * 2; // 16 KiB
constexpr uint32_t SPIN_LIMIT = 1u << 23;   // a protocol bug traps within a fraction of a second instead of hanging the box
constexpr float W_SCALE = 64.0f;            // packed weights = W_SCALE * w; TMEM accumulators carry the same factor
constexpr float W_INV = 1.0f / W_SCALE;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// latency-critical wait: pure polling (no hardware suspend), bounded like mbar_wait
__device__ __forceinline__ void mbar_poll(uint32_t bar, uint32_t parity, int* err, int code) {
    uint32_t spins = 0;
    while (!mbar_test_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) {
            atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}
// bounded wait: a protocol bug must surface as an error code, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) {
            atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}
// one lane of a converged warp (the instruction operands of UTCHMMA / UBLKCP live in uniform registers: issuing them from
// inside a lane-divergent branch makes the compiler emit a uniformisation loop per instruction, ~100 cycles each)
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    return leader != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* v) {     // no wait: pair with tmem_ld_wait()
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st32_issue(uint32_t taddr, const uint32_t* v) {   // no wait: pair with tmem_st_wait()
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

// fp32 -> fp16 pair, round-to-nearest-even, finite saturation (an activation beyond fp16's range must not turn into inf/NaN)
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ float2 f16x2_to_f32(uint32_t p) {
    return __half22float2(*(const __half2*)&p);
}
// fp16 hi/lo split of 8 consecutive k elements of one row -> two 16-byte K-major chunks (lo = fp16(x - hi): 22 significant bits)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = cvt_f16x2_sat(x[2 * i], x[2 * i + 1]);
        const float2 f = f16x2_to_f32(h[i]);
        l[i] = cvt_f16x2_sat(x[2 * i] - f.x, x[2 * i + 1] - f.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// single element (lin_in features, weight packing)
__device__ __forceinline__ void split1(float x, __half& hi, __half& lo) {
    const uint32_t h = cvt_f16x2_sat(x, 0.0f);
    hi = __ushort_as_half((unsigned short)(h & 0xFFFFu));
    const uint32_t l = cvt_f16x2_sat(x - __half2float(hi), 0.0f);
    lo = __ushort_as_half((unsigned short)(l & 0xFFFFu));
}

// sample -> world point + direction (explicit points, or ray + depth: nerf_renderer.py:304)
__device__ __forceinline__ void sample_point(const QueryArgs& q, long long smp, float& px, float& py, float& pz,
                                             float& dx, float& dy, float& dz) {
    if (q.xyz) {
        px = q.xyz[smp * 3]; py = q.xyz[smp * 3 + 1]; pz = q.xyz[smp * 3 + 2];
        dx = q.viewdirs[smp * 3]; dy = q.viewdirs[smp * 3 + 1]; dz = q.viewdirs[smp * 3 + 2];
    } else {
        const long long ray = smp / q.K;
        const float* r = q.rays + ray * 8;
        const float z = q.z[smp];
        dx = r[3]; dy = r[4]; dz = r[5];
        px = __fadd_rn(r[0], __fmul_rn(z, dx));
        py = __fadd_rn(r[1], __fmul_rn(z, dy));
        pz = __fadd_rn(r[2], __fmul_rn(z, dz));
    }
}

// ------------------------------------------------------------------------------------------------
// weight / bias packing
// ------------------------------------------------------------------------------------------------
// One CTA per 16 KiB tile pair: W (out,in) fp32 row-major -> [hi tile][lo tile] of W_SCALE * W in fp16, each 128 rows x 64 k in
// K-major SWIZZLE_128B order: byte(r,k) = (r/8)*1024 + (r%8)*128 + (((k/8) ^ (r%8)) * 16) + (k%8)*2.
__global__ void pack_weight_kernel(const float* __restrict__ W, int out_dim, int in_dim, int nkb, uint8_t* __restrict__ dst) {
    const int m = blockIdx.x / nkb, kb = blockIdx.x % nkb;
    uint8_t* hi = dst + (size_t)blockIdx.x * 2 * WTILE_BYTES;
    uint8_t* lo = hi + WTILE_BYTES;
    for (int i = threadIdx.x; i < 128 * KBLK; i += blockDim.x) {
        const int r = i / KBLK, k = i % KBLK;
        const int o = 128 * m + r, c = KBLK * kb + k;
        const float w = (o < out_dim && c < in_dim) ? W_SCALE * W[(size_t)o * in_dim + c] : 0.0f;
        __half h, l;
        split1(w, h, l);
        const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((k >> 3) ^ (r & 7)) << 4) + (uint32_t)(k & 7) * 2u;
        *(__half*)(hi + off) = h;
        *(__half*)(lo + off) = l;
    }
}

// Bias rows.  PRE (n_pre = combine_layer blocks): rows [0,n_pre): bias to add when reading X before fc_0 of block b
//   = b_in + sum_{j<=b} b_z[j] + sum_{j<b} b_fc1[j];  rows [n_pre,2n_pre): b_fc0[b];  row 2n_pre: total after the last block.
// POST (n_post blocks, X starts as the true x_c): row 0 unused (zeros); rows [1,n_post]: sum_{j<=b} b_fc1[pre+j];
//   rows [n_post+1, 2n_post]: b_fc0[pre+b];  row 2n_post+1: lin_out bias (first 4 entries).
// PAIR (pair kernel, lin_z hoisted into the Y maps): rows [0,n_pre): bias folded into Y_b = b_z[b] + (b == 0 ? b_in : b_fc1[b-1]);
//   row n_pre: b_fc1[n_pre-1], added when the combined x_c is written.
__global__ void pack_bias_kernel(MlpDev m, float* __restrict__ pre, float* __restrict__ post, float* __restrict__ pair) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= HID) return;
    const int n_pre = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks, n_post = m.n_blocks - n_pre;
    float acc = m.b_in[h];
    for (int b = 0; b < n_pre; ++b) {
        pair[(size_t)b * HID + h] = m.b_z[b][h] + (b == 0 ? m.b_in[h] : m.b_fc1[b - 1][h]);
        acc += m.b_z[b][h];
        pre[(size_t)b * HID + h] = acc;
        pre[(size_t)(n_pre + b) * HID + h] = m.b_fc0[b][h];
        acc += m.b_fc1[b][h];
    }
    pre[(size_t)(2 * n_pre) * HID + h] = acc;
    pair[(size_t)n_pre * HID + h] = m.b_fc1[n_pre - 1][h];
    post[h] = 0.0f;
    float acc2 = 0.0f;
    for (int b = 0; b < n_post; ++b) {
        acc2 += m.b_fc1[n_pre + b][h];
        post[(size_t)(1 + b) * HID + h] = acc2;
        post[(size_t)(n_post + 1 + b) * HID + h] = m.b_fc0[n_pre + b][h];
    }
    post[(size_t)(2 * n_post + 1) * HID + h] = h < m.d_out ? m.b_out[h] : 0.0f;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void tc_release(TcState& t) {
    if (t.wpack) cudaFree(t.wpack);
    if (t.bias) cudaFree(t.bias);
    if (t.scratch) cudaFree(t.scratch);
    if (t.err_flag) cudaFreeHost(t.err_flag);
    if (t.table2) cudaFree(t.table2);
    t.table2 = nullptr; t.table2_parity = -1;
    if (t.zmap) cudaFree(t.zmap);
    t.zmap = nullptr; t.zmap_bytes = 0; t.zmap_valid = false;
    for (int i = 0; i < 4; ++i) if (t.ev[i]) { cudaEventDestroy(t.ev[i]); t.ev[i] = nullptr; }
    t.wpack = nullptr; t.bias = nullptr; t.scratch = nullptr; t.err_flag = nullptr;
    t.wpack_bytes = t.scratch_bytes = 0;
    t.ready = false;
}

cudaError_t tc_pack_weights(TcState& t, const MlpDev& m, cudaStream_t st) {
    using namespace tc;
    t.ready = false;
    t.zmap_valid = false;         // the hoisted lin_z maps depend on the weights
    t.table2_parity = -1;         // tile tables depend on the layer shapes
    const int n_pre = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks, n_post = m.n_blocks - n_pre;
    if (m.d_hidden != HID) { snprintf(t.why, sizeof(t.why), "d_hidden=%d (tcgen05 path serves 512)", m.d_hidden); return cudaSuccess; }
    if (m.d_in > KBLK) { snprintf(t.why, sizeof(t.why), "d_in=%d > 64", m.d_in); return cudaSuccess; }
    if (m.d_latent > HID || (m.d_latent % KBLK)) { snprintf(t.why, sizeof(t.why), "d_latent=%d must be a multiple of 64, <= 512", m.d_latent); return cudaSuccess; }
    if (n_pre < 1 || n_post < 1) { snprintf(t.why, sizeof(t.why), "need >=1 block before and after combine_layer (n_blocks=%d combine_layer=%d)", m.n_blocks, m.combine_layer); return cudaSuccess; }
    const int kbz = m.d_latent / KBLK, kbh = HID / KBLK;
    t.n_pre = n_pre; t.n_post = n_post;
    t.pairs_pre = MT * 1 + n_pre * (MT * kbz + 2 * MT * kbh);
    t.pairs_post = n_post * 2 * MT * kbh;                 // (lin_out is not packed: the POST epilogue computes it from the fp32 weights)
    const size_t bytes = (size_t)(t.pairs_pre + t.pairs_post) * 2 * WTILE_BYTES;
    if (bytes > t.wpack_bytes) {
        if (t.wpack) cudaFree(t.wpack);
        t.wpack = nullptr; t.wpack_bytes = 0;
        TCK(cudaMalloc(&t.wpack, bytes));
        t.wpack_bytes = bytes;
    }
    if (!t.bias) TCK(cudaMalloc((void**)&t.bias, (size_t)(5 * DINER_MAX_BLOCKS + 6) * HID * sizeof(float)));
    if (!t.err_flag) {   // host-mapped so the watchdog code survives a trapped kernel
        TCK(cudaHostAlloc((void**)&t.err_flag, sizeof(int), cudaHostAllocMapped));
        *t.err_flag = 0;
    }
    uint8_t* p = (uint8_t*)t.wpack;
    auto pack = [&](const float* W, int out_dim, int in_dim, int nkb, int n_mt) -> cudaError_t {
        pack_weight_kernel<<<n_mt * nkb, 256, 0, st>>>(W, out_dim, in_dim, nkb, p);
        g_launches++;
        p += (size_t)n_mt * nkb * 2 * WTILE_BYTES;
        return cudaGetLastError();
    };
    TCK(pack(m.w_in, HID, m.d_in, 1, MT));
    for (int b = 0; b < n_pre; ++b) {
        TCK(pack(m.w_z[b], HID, m.d_latent, kbz, MT));
        TCK(pack(m.w_fc0[b], HID, HID, kbh, MT));
        TCK(pack(m.w_fc1[b], HID, HID, kbh, MT));
    }
    for (int b = n_pre; b < m.n_blocks; ++b) {
        TCK(pack(m.w_fc0[b], HID, HID, kbh, MT));
        TCK(pack(m.w_fc1[b], HID, HID, kbh, MT));
    }
    t.bias_post_off = (size_t)(2 * DINER_MAX_BLOCKS + 2) * HID;
    t.bias_pair_off = (size_t)(4 * DINER_MAX_BLOCKS + 4) * HID;
    pack_bias_kernel<<<(HID + 127) / 128, 128, 0, st>>>(m, t.bias, t.bias + t.bias_post_off, t.bias + t.bias_pair_off);
    g_launches++;
    TCK(cudaGetLastError());
    {   // tensor map over the packed stream for the pair kernel's cta_group::2 TMA loads (plain linear 16 KiB boxes)
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        t.wmap_ok = false;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
            cuuint64_t gdim[2] = {64, (cuuint64_t)(t.pairs_pre + t.pairs_post) * 2 * 128};
            cuuint64_t gstr[1] = {128};
            cuuint32_t box[2] = {64, 128};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = ((EncodeFn)fn)(&t.wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, t.wpack, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            t.wmap_ok = (r == CUDA_SUCCESS);
        }
        (void)cudaGetLastError();
    }
    t.ready = true;
    t.why[0] = 0;
    return cudaSuccess;
}
