// tcgen05 path of the per-sample network (PixelNeRF.forward + ResnetFC.forward), sm_100a.
//
// Formulation (DESIGN.md §"tcgen05 MLP"): every Linear is evaluated TRANSPOSED, D^T[h][n] = W[h][:] . x[n][:],
//   A operand = weights  (M = 128 hidden units per M-tile, K-major == the reference's (out,in) row-major layout)
//   B operand = activations of a 64-row tile (N = 64 sample-views or samples), MN-major in shared memory
//   D         = fp32 accumulators in TMEM: lane = hidden unit, column = row of the tile
// so that (i) UMMA_M = 128 runs the tensor core at full rate with only 64 rows per CTA, (ii) the residual stream
// x (512 x 64 fp32 = 256 TMEM columns) and the block-internal `net` (another 256 columns) both stay in TMEM for the
// whole network, (iii) the epilogue thread that owns hidden unit h writes 8 consecutive rows as one 16-byte
// shared-memory store into the next layer's B operand, and (iv) the mean over the views of a sample is a sum of
// adjacent TMEM columns inside one thread.
//
// Two persistent kernels per batch of samples:
//   PRE  (per sample-view rows): lin_in, then per block b < combine_layer: x += lin_z[b](latent); x += fc_1(relu(fc_0(relu(x))))
//        the bilinear latent gather, the positional encodings and the nearest depth lookup are computed in the
//        kernel straight into the B operand; the combined (view-averaged) x_c is the only intermediate written out
//   POST (per sample): remaining blocks, lin_out (M padded to 128), sigmoid / relu -> (n,4)
// Weights stream from L2 through a ring of 16 KiB stages with 1-D bulk TMA (cp.async.bulk), optionally
// multicast across a thread-block cluster so that one L2 read feeds CL CTAs.
//
// PARITY mode evaluates every product as hi*hi + hi*lo + lo*hi with bf16 hi/lo splits of both operands and fp32
// accumulation (3 MMAs, |err| ~ 1e-5 vs fp32); FAST mode is single-pass bf16.
//
// Reference semantics: src/models/resnetfc.py:61-69,129-159; src/models/pixelnerf.py:91-143.
#include "mlp_tc.h"

#include <cuda_bf16.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#define TCK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) return _e; } while (0)

namespace tc {

constexpr int TILE_N = 64;                  // rows per CTA tile (UMMA N)
constexpr int HID = 512;                    // d_hidden served by this path
constexpr int MT = HID / 128;               // M-tiles
constexpr int KBLK = 64;                    // K elements per weight tile (one 128-byte swizzle atom)
constexpr int WTILE_BYTES = 128 * KBLK * 2; // 16 KiB
constexpr int NUM_THREADS = 512;                // 16 warps: 1 MMA issuer, 8 workers, 7 weight producers
constexpr int NUM_PRODUCERS = 7;               // bulk-TMA ops of one warp are serialised (~0.4 us each, measured):
                                               // bandwidth scales with the number of issuing warps
constexpr int WORKER_WARP0 = 4;
constexpr int NUM_WORKER_WARPS = 8;
constexpr int NUM_WORKERS = NUM_WORKER_WARPS * 32;
constexpr int B_BYTES = HID * TILE_N * 2;   // 64 KiB per bf16 copy of the B operand (K = 512 rows of 128 B)
constexpr int TMEM_COLS = 512;
constexpr int COL_X = 0, COL_NET = 256;
constexpr int MAX_STEPS = 3 * DINER_MAX_BLOCKS + 2;
constexpr uint32_t SPIN_LIMIT = 1u << 23;   // a protocol bug traps within a fraction of a second instead of hanging the box

struct GemmStep {
    short nkb;        // K blocks of 64
    short n_mt;       // M tiles
    short dst_col;    // TMEM column base of D
    short accumulate; // accumulate onto the existing D
};

struct RowTap {       // bilinear tap set of one sample-view row (written by the prep phase)
    int pix00;        // pixel index (view base included); -1 dx/dy packed below
    int dxy;          // bit0: x1 = x0 + 1, bit1: y1 = y0 + 1 (0 when clamped at the border)
    float ex, wx, ey, wy;
};

struct Args {
    SceneDev s;
    QueryArgs q;
    const uint8_t* wstream;   // weight tiles of this kernel in execution order: [hi 16 KiB][lo 16 KiB] per chunk
    const float* bias;        // bias rows (HID floats each), see pack_bias_kernel
    GemmStep steps[MAX_STEPS];
    int n_steps, n_blocks;
    int tiles_per_layerset;   // 16 KiB tiles consumed per CTA tile (hi+lo in parity, hi only in fast)
    long long s_begin;        // first sample of this sub-batch
    long long n_samples;      // samples in this sub-batch
    long long n_total;        // total samples (for clamping)
    long long n_tiles;        // CTA tiles of this kernel in this sub-batch
    int NV, spv;              // views, samples per PRE tile (= 64 / NV)
    float* xc;                // [tileB][HID][64] fp32 combined activations
    float* out;               // (n_total, 4)
    int* err;
};

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

template <int CL>
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    if constexpr (CL == 1) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    } else {
        const uint16_t mask = (uint16_t)((1u << CL) - 1);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
    }
}
template <int CL>
__device__ __forceinline__ void umma_commit_stage(uint32_t bar) {     // frees a weight stage in every CTA of the cluster
    if constexpr (CL == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        const uint16_t mask = (uint16_t)((1u << CL) - 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(mask) : "memory");
    }
}
__device__ __forceinline__ void umma_commit_local(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
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
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), A K-major (bit15=0),
// B MN-major (bit16=1), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// bf16 hi/lo split of 8 consecutive rows (n) of one hidden unit / channel -> two 16-byte vectors
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        const float r0 = x[2 * i] - __bfloat162float(p.x), r1 = x[2 * i + 1] - __bfloat162float(p.y);
        const __nv_bfloat162 ql = __floats2bfloat162_rn(r0, r1);
        h[i] = *(const uint32_t*)&p;
        l[i] = *(const uint32_t*)&ql;
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of (k-row, 8-row chunk c) inside an MN-major SWIZZLE_128B B operand (128 B per k-row)
__device__ __forceinline__ uint32_t b_off(int k, int c) { return (uint32_t)k * 128u + (uint32_t)((c ^ (k & 7)) << 4); }

// ------------------------------------------------------------------------------------------------
// worker-side building blocks
// ------------------------------------------------------------------------------------------------
// TMEM region (4 M-tiles x 64 cols at colbase) + bias -> relu -> bf16 (hi/lo) B operand
template <bool PARITY>
__device__ __forceinline__ void epilogue_to_B(uint32_t tmem, int colbase, const float* __restrict__ bias,
                                              uint8_t* Bhi, uint8_t* Blo, int q, int lane, int hf) {
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(colbase + 64 * m + 32 * hf), v);
        const int h = 128 * m + 32 * q + lane;
        const float bv = __ldg(bias + h);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[8 * c4 + i]) + bv, 0.0f);
            uint4 hi, lo;
            split8(x, hi, lo);
            const uint32_t off = b_off(h, 4 * hf + c4);
            *(uint4*)(Bhi + off) = hi;
            if (PARITY) *(uint4*)(Blo + off) = lo;
        }
    }
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

// PRE prep: 4 threads per row.  Camera transform, projection, nearest depth, positional encodings -> the
// lin_in B operand (K rows 0..63), plus the row's bilinear tap set for the latent gather.
template <bool PARITY>
__device__ __forceinline__ void prep_rows(const Args& a, long long tile, int wt, uint8_t* Bhi, uint8_t* Blo,
                                          RowTap* taps) {
    const SceneDev& s = a.s;
    const int r = wt & 63, part = wt >> 6;
    long long smp = a.s_begin + tile * a.spv + r / a.NV;
    if (smp >= a.n_total) smp = a.n_total - 1;
    const int v = r % a.NV;
    const int sb = (int)(smp / a.q.n_per_sb);
    float px, py, pz, dx, dy, dz;
    sample_point(a.q, smp, px, py, pz, dx, dy, dz);
    const int sv = sb * s.NV + v;
    const float* P = s.poses + (size_t)sv * 16;
    float p[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) p[k] = __ldg(P + k);
    float xc, yc, zc, dxc, dyc, dzc;
    world_to_cam(p, px, py, pz, xc, yc, zc);
    rotate_to_cam(p, dx, dy, dz, dxc, dyc, dzc);
    const float u = project_axis(xc, zc, __ldg(s.focal + sv * 2), __ldg(s.cxy + sv * 2), s.imgW);
    const float w = project_axis(yc, zc, __ldg(s.focal + sv * 2 + 1), __ldg(s.cxy + sv * 2 + 1), s.imgH);
    const float dd = __fsub_rn(lookup_depth(s, sv, u, w), zc);
    if (part == 0) {
        const LatTaps t = latent_taps(s, u, w);
        RowTap rt;
        rt.pix00 = sv * s.Hl * s.Wl + t.o00;
        rt.dxy = (t.o01 != t.o00 ? 1 : 0) | (t.o10 != t.o00 ? 2 : 0);
        // latent_taps' weights are products of these four factors
        float x = unnormalize(__fmul_rn(u, s.lat_sx), (float)s.Wl), y = unnormalize(__fmul_rn(w, s.lat_sy), (float)s.Hl);
        x = fminf(fmaxf(x, 0.0f), (float)(s.Wl - 1));
        y = fminf(fmaxf(y, 0.0f), (float)(s.Hl - 1));
        if (!(x == x)) x = 0.0f;
        if (!(y == y)) y = 0.0f;
        const float xf = floorf(x), yf = floorf(y);
        rt.ex = (xf + 1.0f) - x; rt.wx = x - xf; rt.ey = (yf + 1.0f) - y; rt.wy = y - yf;
        taps[r] = rt;
    }
    const int d_in = 3 + 6 * s.num_freqs + 3 + 1 + 2 * s.num_freqs;
    const int c = r >> 3;
#pragma unroll 1
    for (int e = part; e < KBLK; e += 4) {
        const float val = e < d_in ? feature_elem(e, s.num_freqs, s.freqs, xc, yc, zc, dxc, dyc, dzc, dd) : 0.0f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(val);
        const uint32_t off = b_off(e, c) + (uint32_t)(r & 7) * 2u;
        *(__nv_bfloat16*)(Bhi + off) = hi;
        if (PARITY) *(__nv_bfloat16*)(Blo + off) = __float2bfloat16_rn(val - __bfloat162float(hi));
    }
}

// PRE gather: bilinear latent (NHWC fp32) of the tile's 64 rows -> B operand (K = L channels).
// Warp task = 8 rows x 32 channels: coalesced 128-byte taps, 16-byte swizzled stores.
template <bool PARITY>
__device__ __forceinline__ void gather_latent(const Args& a, int wwarp, int lane, uint8_t* Bhi, uint8_t* Blo,
                                              const RowTap* taps) {
    const SceneDev& s = a.s;
    const int n_cp = s.L >> 6;                 // pairs of 32-channel groups
    const int n_tasks = 8 * n_cp;
#pragma unroll 1
    for (int t = wwarp; t < n_tasks; t += NUM_WORKER_WARPS) {
        const int rg = t / n_cp, cp = t % n_cp;
        const int k = 64 * cp + lane;
        float acc0[8], acc1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const RowTap rt = taps[8 * rg + i];
            const float* b00 = s.latent + (size_t)rt.pix00 * s.L + k;
            const size_t ox = (rt.dxy & 1) ? (size_t)s.L : 0, oy = (rt.dxy & 2) ? (size_t)s.Wl * s.L : 0;
            const float v00 = __ldg(b00), v01 = __ldg(b00 + ox), v10 = __ldg(b00 + oy), v11 = __ldg(b00 + oy + ox);
            const float u00 = __ldg(b00 + 32), u01 = __ldg(b00 + ox + 32), u10 = __ldg(b00 + oy + 32), u11 = __ldg(b00 + oy + ox + 32);
            const float w00 = rt.ex * rt.ey, w01 = rt.wx * rt.ey, w10 = rt.ex * rt.wy, w11 = rt.wx * rt.wy;
            acc0[i] = v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11;
            acc1[i] = u00 * w00 + u01 * w01 + u10 * w10 + u11 * w11;
        }
        uint4 hi, lo;
        split8(acc0, hi, lo);
        uint32_t off = b_off(k, rg);
        *(uint4*)(Bhi + off) = hi;
        if (PARITY) *(uint4*)(Blo + off) = lo;
        split8(acc1, hi, lo);
        off = b_off(k + 32, rg);
        *(uint4*)(Bhi + off) = hi;
        if (PARITY) *(uint4*)(Blo + off) = lo;
    }
}

// mean over the NV adjacent columns of each sample (resnetfc.py:148-151); NV is a compile-time constant so that
// the 32 accumulator registers are indexed statically (dynamic indexing would spill them to local memory)
template <int NV>
__device__ __forceinline__ void combine_store(const uint32_t* v, float bv, float* dst) {
    constexpr int PER = 32 / NV;
    float o[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        float acc = __uint_as_float(v[j * NV]);
#pragma unroll
        for (int vv = 1; vv < NV; ++vv) acc += __uint_as_float(v[j * NV + vv]);
        o[j] = acc * (1.0f / (float)NV) + bv;
    }
    if constexpr (PER >= 4) {
#pragma unroll
        for (int j = 0; j < PER; j += 4) *(float4*)(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < PER; ++j) dst[j] = o[j];
    }
}

// worker -> MMA hand-off: make generic-proxy smem writes visible to the async proxy, order TMEM accesses, arrive once per warp
__device__ __forceinline__ void worker_arrive(uint32_t bar, int lane) {
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <bool PARITY> struct Cfg {
    static constexpr int NST = PARITY ? 6 : 8;                          // weight stages (16 KiB each)
    static constexpr int OFF_B_HI = NST * WTILE_BYTES;
    static constexpr int OFF_B_LO = OFF_B_HI + B_BYTES;
    static constexpr int OFF_TAPS = OFF_B_LO + (PARITY ? B_BYTES : 0);
    static constexpr int OFF_BARS = OFF_TAPS + TILE_N * (int)sizeof(RowTap);
    static constexpr int SMEM_BYTES = OFF_BARS + 256;
};

template <bool PARITY, bool POST, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_tc_kernel(const __grid_constant__ Args a) {
    using C = Cfg<PARITY>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0;
    const uint32_t smem_base = smem_u32(smem);
    uint8_t* Bhi = smem + C::OFF_B_HI;
    uint8_t* Blo = smem + C::OFF_B_LO;
    RowTap* taps = (RowTap*)(smem + C::OFF_TAPS);
    const uint32_t bar_full = smem_base + C::OFF_BARS;             // NST x 8 B
    const uint32_t bar_empty = bar_full + 8 * C::NST;              // NST x 8 B
    const uint32_t bar_opnd = bar_empty + 8 * C::NST;              // workers -> MMA: B operand ready
    const uint32_t bar_acc = bar_opnd + 8;                         // MMA -> workers: accumulators ready / B free
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + C::OFF_BARS + 8 * (2 * C::NST + 2));

    if ((smem_base & 1023u) != 0) { if (threadIdx.x == 0) atomicExch(a.err, 90); __trap(); }
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, CL); }
        mbar_init(bar_opnd, NUM_WORKER_WARPS);
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // every CTA of a cluster runs the same number of rounds so the multicast weight ring stays in lockstep;
    // tiles past the end are computed on clamped inputs and write nothing
    const long long first = (long long)blockIdx.x, stride = (long long)gridDim.x;
    const long long n_rounds = (a.n_tiles + stride - 1) / stride;

    const int prod_idx = warp == 0 ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : (warp >= 12 ? warp - 9 : -1)));
    if (prod_idx >= 0) {
        // ===== weight producers: 1-D bulk TMA of 16 KiB tiles.  Ring stage st is always filled by producer
        //       st % NUM_PRODUCERS, so every barrier sees its uses in order from one thread (a parity wait cannot
        //       tell "one phase ahead" from "two phases ahead").  Each CTA of a cluster multicasts 1/CL of a tile.
        {
            constexpr uint32_t SLICE = WTILE_BYTES / CL;
            const bool leader = elect_one();
            const long long total_uses = n_rounds * a.tiles_per_layerset;
            for (long long base = 0; base < total_uses; base += C::NST) {
                for (int st = prod_idx; st < C::NST; st += NUM_PRODUCERS) {
                    const long long use = base + st;
                    if (use >= total_uses) break;
                    const int t = (int)(use % a.tiles_per_layerset);
                    const uint32_t ph = (uint32_t)((use / C::NST) & 1);
                    mbar_wait(bar_empty + 8 * st, ph ^ 1, a.err, 10);
                    if (leader) {
                        mbar_arrive_expect_tx(bar_full + 8 * st, WTILE_BYTES);
                        const size_t gt = PARITY ? (size_t)t : (size_t)2 * t;      // fast mode skips the lo tiles
                        bulk_g2s<CL>(smem_base + st * WTILE_BYTES + crank * SLICE, a.wstream + gt * WTILE_BYTES + crank * SLICE,
                                     SLICE, bar_full + 8 * st);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the (warp-uniform) control flow, one elected lane issues
        {
            constexpr uint32_t IDESC = make_idesc(128, TILE_N);
            const bool leader = elect_one();
            const uint64_t bdesc_hi = make_desc(smem_base + C::OFF_B_HI, 0, 1024);
            const uint64_t bdesc_lo = make_desc(smem_base + C::OFF_B_LO, 0, 1024);
            uint32_t use = 0, it = 0;
            for (long long rd = 0; rd < n_rounds; ++rd) {
                for (int sidx = 0; sidx < a.n_steps; ++sidx, ++it) {
                    const GemmStep gs = a.steps[sidx];
                    mbar_wait(bar_opnd, it & 1, a.err, 20);
                    tc_fence_after();
                    for (int m = 0; m < gs.n_mt; ++m) {
                        const uint32_t d = tmem + (uint32_t)(gs.dst_col + 64 * m);
                        for (int kb = 0; kb < gs.nkb; ++kb) {
                            {   // hi weight tile: hi*hi (+ hi*lo)
                                const uint32_t st = use % C::NST, ph = (use / C::NST) & 1;
                                mbar_wait(bar_full + 8 * st, ph, a.err, 30);
                                tc_fence_after();
                                if (leader) {
                                    const uint64_t adesc = make_desc(smem_base + st * WTILE_BYTES, 16, 1024);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const uint64_t bo = (uint64_t)(((kb * 4 + j) * 2048) >> 4);
                                        umma_bf16(d, adesc + 2 * j, bdesc_hi + bo, IDESC, (gs.accumulate | kb | j) ? 1u : 0u);
                                        if (PARITY) umma_bf16(d, adesc + 2 * j, bdesc_lo + bo, IDESC, 1u);
                                    }
                                    umma_commit_stage<CL>(bar_empty + 8 * st);
                                }
                                __syncwarp();
                                ++use;
                            }
                            if (PARITY) {   // lo weight tile: lo*hi
                                const uint32_t st = use % C::NST, ph = (use / C::NST) & 1;
                                mbar_wait(bar_full + 8 * st, ph, a.err, 31);
                                tc_fence_after();
                                if (leader) {
                                    const uint64_t adesc = make_desc(smem_base + st * WTILE_BYTES, 16, 1024);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const uint64_t bo = (uint64_t)(((kb * 4 + j) * 2048) >> 4);
                                        umma_bf16(d, adesc + 2 * j, bdesc_hi + bo, IDESC, 1u);
                                    }
                                    umma_commit_stage<CL>(bar_empty + 8 * st);
                                }
                                __syncwarp();
                                ++use;
                            }
                        }
                    }
                    if (leader) umma_commit_local(bar_acc);
                    __syncwarp();
                }
            }
        }
    } else if (warp >= WORKER_WARP0 && warp < WORKER_WARP0 + NUM_WORKER_WARPS) {
        // ===== workers: operand producers + epilogues.  TMEM lane quarter q = warp % 4, column half hf
        const int wwarp = warp - WORKER_WARP0, wt = threadIdx.x - WORKER_WARP0 * 32;
        const int q = warp & 3, hf = wwarp >> 2;
        uint32_t it = 0;   // counts bar_acc completions consumed
        for (long long rd = 0; rd < n_rounds; ++rd) {
            const long long tile_raw = first + rd * stride;
            const bool live = tile_raw < a.n_tiles;
            const long long tile = live ? tile_raw : a.n_tiles - 1;
            if constexpr (!POST) {
                // (the previous tile's last wait on bar_acc guarantees B and the tap table are free)
                prep_rows<PARITY>(a, tile, wt, Bhi, Blo, taps);
                worker_arrive(bar_opnd, lane);                                   // -> lin_in
                for (int b = 0; b < a.n_blocks; ++b) {
                    mbar_wait(bar_acc, it & 1, a.err, 40); ++it;                 // previous GEMM done: B free
                    tc_fence_after();
                    gather_latent<PARITY>(a, wwarp, lane, Bhi, Blo, taps);
                    worker_arrive(bar_opnd, lane);                               // -> lin_z[b]
                    mbar_wait(bar_acc, it & 1, a.err, 41); ++it;
                    tc_fence_after();
                    epilogue_to_B<PARITY>(tmem, COL_X, a.bias + (size_t)b * HID, Bhi, Blo, q, lane, hf);
                    worker_arrive(bar_opnd, lane);                               // -> fc_0[b]
                    mbar_wait(bar_acc, it & 1, a.err, 42); ++it;
                    tc_fence_after();
                    epilogue_to_B<PARITY>(tmem, COL_NET, a.bias + (size_t)(a.n_blocks + b) * HID, Bhi, Blo, q, lane, hf);
                    worker_arrive(bar_opnd, lane);                               // -> fc_1[b]
                }
                mbar_wait(bar_acc, it & 1, a.err, 43); ++it;
                tc_fence_after();
                // combine: mean over the NV adjacent columns of each sample (resnetfc.py:148-151) -> x_c
                const float* cb = a.bias + (size_t)(2 * a.n_blocks) * HID;
                const int per = 32 / a.NV;                                       // samples in this thread's 32 columns
                const long long tileB = tile / a.NV;
                const int n0 = (int)(tile % a.NV) * a.spv + hf * per;
#pragma unroll 1
                for (int m = 0; m < MT; ++m) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(COL_X + 64 * m + 32 * hf), v);
                    const int h = 128 * m + 32 * q + lane;
                    const float bv = __ldg(cb + h);
                    if (live) {
                        float* dst = a.xc + ((size_t)tileB * HID + h) * TILE_N + n0;
                        switch (a.NV) {
                            case 1: combine_store<1>(v, bv, dst); break;
                            case 2: combine_store<2>(v, bv, dst); break;
                            case 4: combine_store<4>(v, bv, dst); break;
                            case 8: combine_store<8>(v, bv, dst); break;
                            case 16: combine_store<16>(v, bv, dst); break;
                            default: combine_store<32>(v, bv, dst); break;
                        }
                    }
                }
                tc_fence_before();
            } else {
                // load x_c: fp32 residual into TMEM (X region) and relu(x_c) into the B operand
#pragma unroll 1
                for (int m = 0; m < MT; ++m) {
                    const int h = 128 * m + 32 * q + lane;
                    const float4* src = (const float4*)(a.xc + ((size_t)tile * HID + h) * TILE_N + 32 * hf);
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 f = __ldg(src + i);
                        v[4 * i] = __float_as_uint(f.x); v[4 * i + 1] = __float_as_uint(f.y);
                        v[4 * i + 2] = __float_as_uint(f.z); v[4 * i + 3] = __float_as_uint(f.w);
                    }
                    tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(COL_X + 64 * m + 32 * hf), v);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        float x[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[8 * c4 + i]), 0.0f);
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        const uint32_t off = b_off(h, 4 * hf + c4);
                        *(uint4*)(Bhi + off) = hi;
                        if (PARITY) *(uint4*)(Blo + off) = lo;
                    }
                }
                worker_arrive(bar_opnd, lane);                                   // -> fc_0 of the first post block
                for (int b = 0; b < a.n_blocks; ++b) {
                    mbar_wait(bar_acc, it & 1, a.err, 50); ++it;
                    tc_fence_after();
                    epilogue_to_B<PARITY>(tmem, COL_NET, a.bias + (size_t)(a.n_blocks + 1 + b) * HID, Bhi, Blo, q, lane, hf);
                    worker_arrive(bar_opnd, lane);                               // -> fc_1[b]
                    mbar_wait(bar_acc, it & 1, a.err, 51); ++it;
                    tc_fence_after();
                    // relu(x + accumulated fc_1 biases) feeds the next block's fc_0, or lin_out after the last block
                    epilogue_to_B<PARITY>(tmem, COL_X, a.bias + (size_t)(b + 1) * HID, Bhi, Blo, q, lane, hf);
                    worker_arrive(bar_opnd, lane);
                }
                mbar_wait(bar_acc, it & 1, a.err, 52); ++it;                     // lin_out done (NET M-tile 0, lanes 0..3)
                tc_fence_after();
                if (q == 0) {
                    uint32_t v[32];
                    tmem_ld32(tmem + (uint32_t)(COL_NET + 32 * hf), v);
                    if (lane < 4 && live) {
                        const float bo = __ldg(a.bias + (size_t)(2 * a.n_blocks + 1) * HID + lane);
                        for (int i = 0; i < 32; ++i) {
                            const long long smp = tile * TILE_N + 32 * hf + i;      // sample within the sub-batch
                            if (smp < a.n_samples) {
                                const float x = __uint_as_float(v[i]) + bo;
                                a.out[(a.s_begin + smp) * 4 + lane] = lane < 3 ? 1.0f / (1.0f + expf(-x)) : fmaxf(x, 0.0f);
                            }
                        }
                    }
                }
                tc_fence_before();
                // all 8 worker warps must be done with X / NET / B before the next tile's loads overwrite them
                asm volatile("bar.sync 1, %0;" ::"n"(NUM_WORKERS) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// weight / bias packing
// ------------------------------------------------------------------------------------------------
// One CTA per 16 KiB tile pair: W (out,in) fp32 row-major -> [hi tile][lo tile], each 128 rows x 64 k in
// K-major SWIZZLE_128B order: byte(r,k) = (r/8)*1024 + (r%8)*128 + (((k/8) ^ (r%8)) * 16) + (k%8)*2.
__global__ void pack_weight_kernel(const float* __restrict__ W, int out_dim, int in_dim, int nkb, uint8_t* __restrict__ dst) {
    const int m = blockIdx.x / nkb, kb = blockIdx.x % nkb;
    uint8_t* hi = dst + (size_t)blockIdx.x * 2 * WTILE_BYTES;
    uint8_t* lo = hi + WTILE_BYTES;
    for (int i = threadIdx.x; i < 128 * KBLK; i += blockDim.x) {
        const int r = i / KBLK, k = i % KBLK;
        const int o = 128 * m + r, c = KBLK * kb + k;
        const float w = (o < out_dim && c < in_dim) ? W[(size_t)o * in_dim + c] : 0.0f;
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((k >> 3) ^ (r & 7)) << 4) + (uint32_t)(k & 7) * 2u;
        *(__nv_bfloat16*)(hi + off) = h;
        *(__nv_bfloat16*)(lo + off) = l;
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
    t.pairs_post = n_post * 2 * MT * kbh + 2 * kbh;       // lin_out packed as 2 M-tiles (the second is zeros, for the pair kernel)
    t.pairs_post_v1 = t.pairs_post - kbh;
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
    TCK(pack(m.w_out, m.d_out, HID, kbh, 2));
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
        t.wmap_small_ok = false;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
            cuuint64_t gdim[2] = {64, (cuuint64_t)(t.pairs_pre + t.pairs_post) * 2 * 128};
            cuuint64_t gstr[1] = {128};
            cuuint32_t box[2] = {64, 128};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = ((EncodeFn)fn)(&t.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, t.wpack, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            t.wmap_ok = (r == CUDA_SUCCESS);
            cuuint32_t box16[2] = {64, 16};
            r = ((EncodeFn)fn)(&t.wmap_small, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, t.wpack, gdim, gstr, box16, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            t.wmap_small_ok = t.wmap_ok && (r == CUDA_SUCCESS);
        }
        (void)cudaGetLastError();
    }
    t.ready = true;
    t.why[0] = 0;
    return cudaSuccess;
}

namespace tc {

template <bool PARITY, bool POST, int CL>
cudaError_t launch_one(const Args& a, int grid, cudaStream_t st) {
    auto kern = mlp_tc_kernel<PARITY, POST, CL>;
    constexpr int smem = Cfg<PARITY>::SMEM_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    g_launches++;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

// co-resident CTAs for a cluster size (cluster size 4 cannot use all 148 SMs: GPCs hold 16/18/20 SMs)
template <int CL>
cudaError_t max_resident(int* out) {
    auto kern = mlp_tc_kernel<true, false, CL>;
    constexpr int smem = Cfg<true>::SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * 64);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    *out = n * CL;
    return e;
}

template <bool PARITY, bool POST>
cudaError_t launch_cl(const Args& a, int grid, int cl, cudaStream_t st) {
    switch (cl) {
        case 1: return launch_one<PARITY, POST, 1>(a, grid, st);
        case 2: return launch_one<PARITY, POST, 2>(a, grid, st);
        case 4: return launch_one<PARITY, POST, 4>(a, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tc

cudaError_t tc_query(TcState& t, const SceneDev& s, const MlpDev& m, const QueryArgs& q, bool parity,
                     int num_sms, cudaStream_t st) {
    using namespace tc;
    const int NV = s.NV;
    if (NV > 32 || (32 % NV)) { snprintf(t.why, sizeof(t.why), "NV=%d views (tcgen05 path needs NV in {1,2,4,8,16,32})", NV); return cudaErrorNotSupported; }
    if (s.L != m.d_latent || (s.L % 32)) { snprintf(t.why, sizeof(t.why), "latent channels %d != d_latent %d", s.L, m.d_latent); return cudaErrorNotSupported; }
    const int d_in = 3 + 6 * s.num_freqs + 3 + 1 + 2 * s.num_freqs;
    if (d_in != m.d_in) { snprintf(t.why, sizeof(t.why), "positional code gives d_in=%d but lin_in expects %d", d_in, m.d_in); return cudaErrorNotSupported; }
    const long long total = (long long)q.SB * q.n_per_sb;
    const int cl = t.cluster > 0 ? t.cluster : 1;
    // usable CTAs: clusters of 4 cannot cover all 148 SMs (GPC sizes), the occupancy query tells how many fit
    const int ci = cl == 1 ? 0 : (cl == 2 ? 1 : 2);
    if (cl != 1 && cl != 2 && cl != 4) { snprintf(t.why, sizeof(t.why), "cluster size %d not in {1,2,4}", cl); return cudaErrorNotSupported; }
    if (t.max_grid[ci] == 0) {
        int n = 0;
        TCK(cl == 1 ? max_resident<1>(&n) : (cl == 2 ? max_resident<2>(&n) : max_resident<4>(&n)));
        t.max_grid[ci] = n > 0 ? n : cl;
    }
    const int grid_cap = t.max_grid[ci] < (num_sms / cl) * cl ? t.max_grid[ci] : (num_sms / cl) * cl;
    const long long sub = t.sub_batch > 0 ? t.sub_batch : 524288;    // samples per sub-batch (x_c scratch = 2 KiB / sample)
    const size_t need = (size_t)((sub + TILE_N - 1) / TILE_N) * HID * TILE_N * sizeof(float);
    if (need > t.scratch_bytes) {
        if (t.scratch) cudaFree(t.scratch);
        t.scratch = nullptr; t.scratch_bytes = 0;
        TCK(cudaMalloc(&t.scratch, need));
        t.scratch_bytes = need;
    }
    const int kbz = m.d_latent / KBLK, kbh = HID / KBLK;
    Args pre{}, post{};
    pre.s = s; pre.q = q; post.s = s; post.q = q;
    pre.wstream = (const uint8_t*)t.wpack;
    post.wstream = (const uint8_t*)t.wpack + (size_t)t.pairs_pre * 2 * WTILE_BYTES;
    pre.bias = t.bias; post.bias = t.bias + t.bias_post_off;
    pre.n_blocks = t.n_pre; post.n_blocks = t.n_post;
    int n = 0;
    pre.steps[n++] = GemmStep{1, MT, COL_X, 0};
    for (int b = 0; b < t.n_pre; ++b) {
        pre.steps[n++] = GemmStep{(short)kbz, MT, COL_X, 1};
        pre.steps[n++] = GemmStep{(short)kbh, MT, COL_NET, 0};
        pre.steps[n++] = GemmStep{(short)kbh, MT, COL_X, 1};
    }
    pre.n_steps = n;
    n = 0;
    for (int b = 0; b < t.n_post; ++b) {
        post.steps[n++] = GemmStep{(short)kbh, MT, COL_NET, 0};
        post.steps[n++] = GemmStep{(short)kbh, MT, COL_X, 1};
    }
    post.steps[n++] = GemmStep{(short)kbh, 1, COL_NET, 0};
    post.n_steps = n;
    pre.tiles_per_layerset = t.pairs_pre * (parity ? 2 : 1);
    post.tiles_per_layerset = t.pairs_post_v1 * (parity ? 2 : 1);
    pre.NV = post.NV = NV;
    pre.spv = post.spv = TILE_N / NV;
    pre.xc = post.xc = (float*)t.scratch;
    pre.out = post.out = q.out;
    pre.err = post.err = t.err_flag;
    pre.n_total = post.n_total = total;
    t.ms_pre = t.ms_post = 0.f;
    if (t.timing && !t.ev[0]) for (int i = 0; i < 4; ++i) TCK(cudaEventCreate(&t.ev[i]));
    for (long long s0 = 0; s0 < total; s0 += sub) {
        const long long ns = total - s0 < sub ? total - s0 : sub;
        pre.s_begin = post.s_begin = s0;
        pre.n_samples = post.n_samples = ns;
        pre.n_tiles = (ns + pre.spv - 1) / pre.spv;
        post.n_tiles = (ns + TILE_N - 1) / TILE_N;
        long long g1 = ((pre.n_tiles + cl - 1) / cl) * cl, g2 = ((post.n_tiles + cl - 1) / cl) * cl;
        const int grid1 = (int)(g1 < grid_cap ? g1 : grid_cap), grid2 = (int)(g2 < grid_cap ? g2 : grid_cap);
        if (t.timing) TCK(cudaEventRecord(t.ev[0], st));
        if (parity) TCK((launch_cl<true, false>(pre, grid1, cl, st)));
        else TCK((launch_cl<false, false>(pre, grid1, cl, st)));
        if (t.timing) TCK(cudaEventRecord(t.ev[1], st));
        if (parity) TCK((launch_cl<true, true>(post, grid2, cl, st)));
        else TCK((launch_cl<false, true>(post, grid2, cl, st)));
        if (t.timing) {
            TCK(cudaEventRecord(t.ev[2], st));
            TCK(cudaEventSynchronize(t.ev[2]));
            float a = 0.f, b = 0.f;
            TCK(cudaEventElapsedTime(&a, t.ev[0], t.ev[1]));
            TCK(cudaEventElapsedTime(&b, t.ev[1], t.ev[2]));
            t.ms_pre += a; t.ms_post += b;
        }
    }
    return cudaSuccess;
}
