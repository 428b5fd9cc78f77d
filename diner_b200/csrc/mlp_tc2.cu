// tcgen05 path, second generation: CTA-PAIR kernel (cta_group::2).
//
// Why (profiles/r1_ncu_mlp_pre_parity.md, tools/mma_probe.cu): in the single-CTA kernel (mlp_tc.cu) every MMA is
// 128x64x16 and reads 6 KiB of shared memory for 32 cycles of math, and every SM streams the complete weight set per
// 64 rows -- shared-memory bandwidth, not the tensor pipe, bounds it.  Here two CTAs of a cluster form one UMMA:
//   D[128 rows x 256 hidden] += A[128 rows x 16] . B[256 hidden x 16]^T          (tcgen05.mma.cta_group::2, M=128, N=256)
//   A = activations, K-major: each CTA holds its own 64 rows          (2 KiB per MMA per SM)
//   B = weights, K-major (the reference's (out,in) layout): each CTA holds 128 of the 256 hidden rows (4 KiB per MMA per SM)
//   D = fp32 in TMEM, "2x2" layout per CTA: lane = row + 64*(n >= 128), column = n % 128
// so each SM does 64x256x16 MACs per MMA (64 cycles of math) for the same 6 KiB of operand reads, and loads only HALF of
// every weight tile.  Each CTA still owns all 512 hidden units of its own 64 rows (x: 256 TMEM columns, net: 256), so no
// activation ever crosses the pair: the only cross-CTA traffic is barrier signalling.
//
// Row-major orientation: the epilogue thread owns one ROW (sample-view / sample) and 32 consecutive hidden units per
// TMEM load, writes 16-byte K-major chunks into the next layer's A operand; the mean over views is a shuffle over
// adjacent lanes; lin_out is an N=32 MMA whose 4 valid outputs land in the row's own thread.
//
// Reference semantics: src/models/resnetfc.py:61-69,129-159; src/models/pixelnerf.py:91-143.
#include "mlp_tc.h"

namespace tc2 {

using tc::RowTap;
using tc::smem_u32;
using tc::mbar_init;
using tc::mbar_arrive_expect_tx;
using tc::mbar_wait;
using tc::mbar_poll;
using tc::fence_barrier_init;
using tc::fence_proxy_async;
using tc::tc_fence_before;
using tc::tc_fence_after;
using tc::elect_one;
using tc::cluster_ctarank;
using tc::cluster_sync_all;
using tc::tmem_ld32;
using tc::tmem_ld32_issue;
using tc::tmem_ld_wait;
using tc::tmem_st32;
using tc::make_desc;
using tc::split8;
using tc::sample_point;

constexpr int ROWS = 64;                    // rows per CTA (128 per pair)
constexpr int HID = 512;
constexpr int KBLK = 64;
constexpr int WTILE_BYTES = 128 * KBLK * 2; // 16 KiB: 128 hidden rows x 64 k, K-major SWIZZLE_128B (same packing as mlp_tc.cu)
constexpr int ACT_KB_BYTES = ROWS * 128;    // 8 KiB per 64-wide K block of the activation operand
constexpr int ACT_BYTES = ACT_KB_BYTES * (HID / KBLK);   // 64 KiB per bf16 copy
constexpr int NUM_THREADS = 512;
constexpr int NUM_PRODUCERS = 3;               // warps 0,2,3: each CTA streams only half of every weight tile (~33 B/clk needed)
constexpr int WORKER_WARP0 = 4;
constexpr int NUM_WORKER_WARPS = 8;
constexpr int NUM_WORKERS = NUM_WORKER_WARPS * 32;
constexpr int NUM_HELPER_WARPS = 4;             // warps 12..15: extra hands for the operand-producing phases (prep, gather); no TMEM access
constexpr int NUM_OPND_WARPS = NUM_WORKER_WARPS + NUM_HELPER_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int COL_X = 0, COL_NET = 256;
constexpr int MAX_STEPS = 3 * DINER_MAX_BLOCKS + 2;

struct GemmStep {
    short nkb;         // K blocks of 64
    short n_tiles;     // N tiles (of n_width hidden units) = weight tiles per CTA per K block
    short n_width;     // UMMA N: 256, or 32 for lin_out
    short dst_col;     // TMEM column base
    short accumulate;
};

struct Args {
    CUtensorMap wmap;           // packed weight stream as rows of 128 B; one box = one 16 KiB tile
    SceneDev s;
    QueryArgs q;
    const uint8_t* wstream;     // packed weight tiles (16 KiB units), shared with mlp_tc.cu's packing
    const int* tile_table;      // [2][uses_per_tile]: 16 KiB tile index of the i-th ring use of CTA rank r
    const float* bias;
    GemmStep steps[MAX_STEPS];
    int n_steps, n_blocks, uses_per_tile;
    long long s_begin, n_samples, n_total, n_tiles;   // n_tiles counts 64-row CTA tiles
    int NV, spv;
    float* xc;                  // [sample][512] fp32 view-combined activations (sub-batch relative)
    float* out;
    int* err;
    long long* dbg_ts;          // profiling: clock64 stamps of pair 0 in round 1 ([cta][role][slot])
    int dbg_skip;               // profiling experiments only: 1 skip gather, 2 skip epilogues, 4 skip prep, 8 skip MMA issue
};

// ---- cluster / pair PTX ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int* err, int code) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > tc::SPIN_LIMIT) { atomicExch(err, code); __threadfence_system(); __trap(); }
    }
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_commit_pair(uint32_t bar) {     // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// A K-major, B K-major, D f32, bf16 inputs; M = 128 over the pair
__host__ __device__ constexpr uint32_t make_idesc2(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of (row r, 8-wide k chunk kc = k/8) in the K-major SWIZZLE_128B activation operand
__device__ __forceinline__ uint32_t act_off(int r, int kc) {
    return (uint32_t)(kc >> 3) * ACT_KB_BYTES + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((kc & 7) ^ (r & 7)) << 4);
}

template <bool PARITY> struct Cfg {
    static constexpr int NST = PARITY ? 6 : 8;
    static constexpr int OFF_A_HI = NST * WTILE_BYTES;
    static constexpr int OFF_A_LO = OFF_A_HI + ACT_BYTES;
    static constexpr int OFF_TAPS = OFF_A_LO + (PARITY ? ACT_BYTES : 0);
    static constexpr int OFF_BARS = OFF_TAPS + ROWS * (int)sizeof(RowTap);
    static constexpr int SMEM_BYTES = OFF_BARS + 256;
};

// ---- worker building blocks ----------------------------------------------------------------------
// TMEM region (this warp's N tile: 128 columns) + per-column bias -> relu -> bf16 hi/lo chunks of the A operand
// 32 accumulator columns of row r (hidden h0..h0+31) + bias -> relu -> four 16-byte K-major chunks (hi / lo)
template <bool PARITY>
__device__ __forceinline__ void convert32(const uint32_t* v, const float* __restrict__ bias, int h0, int r, uint8_t* Ahi, uint8_t* Alo) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        const float4 b0 = __ldg((const float4*)(bias + h0 + 8 * c8)), b1 = __ldg((const float4*)(bias + h0 + 8 * c8 + 4));
        float x[8];
        x[0] = fmaxf(__uint_as_float(v[8 * c8 + 0]) + b0.x, 0.0f); x[1] = fmaxf(__uint_as_float(v[8 * c8 + 1]) + b0.y, 0.0f);
        x[2] = fmaxf(__uint_as_float(v[8 * c8 + 2]) + b0.z, 0.0f); x[3] = fmaxf(__uint_as_float(v[8 * c8 + 3]) + b0.w, 0.0f);
        x[4] = fmaxf(__uint_as_float(v[8 * c8 + 4]) + b1.x, 0.0f); x[5] = fmaxf(__uint_as_float(v[8 * c8 + 5]) + b1.y, 0.0f);
        x[6] = fmaxf(__uint_as_float(v[8 * c8 + 6]) + b1.z, 0.0f); x[7] = fmaxf(__uint_as_float(v[8 * c8 + 7]) + b1.w, 0.0f);
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = act_off(r, (h0 >> 3) + c8);
        *(uint4*)(Ahi + off) = hi;
        if (PARITY) *(uint4*)(Alo + off) = lo;
    }
}
// TMEM region (this warp's N tile: 128 columns) + per-column bias -> relu -> bf16 hi/lo chunks of the A operand.
// Two TMEM loads are kept in flight (the second 64 columns load while the first are converted).
template <bool PARITY>
__device__ __forceinline__ void epilogue_to_A(uint32_t tmem, int colbase, const float* __restrict__ bias, uint8_t* Ahi,
                                              uint8_t* Alo, int q, int lane, int n2) {
    const int r = 32 * (q & 1) + lane;
    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(colbase + 128 * n2);
    const int hb = 256 * n2 + 128 * (q >> 1);
    uint32_t va[32], vb[32];
    tmem_ld32_issue(t0, va);
    tmem_ld32_issue(t0 + 32, vb);
    tmem_ld_wait();
    convert32<PARITY>(va, bias, hb, r, Ahi, Alo);
    tmem_ld32_issue(t0 + 64, va);
    convert32<PARITY>(vb, bias, hb + 32, r, Ahi, Alo);
    tmem_ld32_issue(t0 + 96, vb);
    tmem_ld_wait();
    convert32<PARITY>(va, bias, hb + 64, r, Ahi, Alo);
    convert32<PARITY>(vb, bias, hb + 96, r, Ahi, Alo);
}

// PRE prep: 4 threads per row -> lin_in A operand (K block 0) + bilinear tap set
template <bool PARITY>
__device__ __forceinline__ void prep_rows(const Args& a, long long tile, int wt, uint8_t* Ahi, uint8_t* Alo, RowTap* taps) {
    const SceneDev& s = a.s;
    const int r = wt & 63, part = wt >> 6;            // NUM_OPND_WARPS * 32 / 64 threads per row
    constexpr int PARTS = NUM_OPND_WARPS * 32 / 64;
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
#pragma unroll 1
    for (int e = part; e < KBLK; e += PARTS) {
        const float val = e < d_in ? feature_elem(e, s.num_freqs, s.freqs, xc, yc, zc, dxc, dyc, dzc, dd) : 0.0f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(val);
        const uint32_t off = act_off(r, e >> 3) + (uint32_t)(e & 7) * 2u;
        *(__nv_bfloat16*)(Ahi + off) = hi;
        if (PARITY) *(__nv_bfloat16*)(Alo + off) = __float2bfloat16_rn(val - __bfloat162float(hi));
    }
}

// PRE gather: bilinear latent of this warp's 8 rows -> A operand.  One row x 256 channels per pass: lane = 8 channels,
// 32-byte loads per tap (1 KiB coalesced per warp), one 16-byte chunk store.
template <bool PARITY>
__device__ __forceinline__ void gather_latent(const Args& a, int wwarp, int lane, uint8_t* Ahi, uint8_t* Alo, const RowTap* taps) {
    const SceneDev& s = a.s;
    const int passes = s.L >> 8;
    const int n_units = ROWS * passes;              // unit = (row, 256-channel pass); lane = 8 channels
    auto issue = [&](int u, float4 (&f)[8], float (&w)[4], uint32_t& off) {
        const int r = u / passes, p = u % passes;
        const RowTap rt = taps[r];
        const size_t ox = (rt.dxy & 1) ? (size_t)s.L : 0, oy = (rt.dxy & 2) ? (size_t)s.Wl * s.L : 0;
        w[0] = rt.ex * rt.ey; w[1] = rt.wx * rt.ey; w[2] = rt.ex * rt.wy; w[3] = rt.wx * rt.wy;
        const int k0 = 256 * p + 8 * lane;
        const float* b00 = s.latent + (size_t)rt.pix00 * s.L + k0;
        f[0] = __ldg((const float4*)b00); f[1] = __ldg((const float4*)(b00 + 4));
        f[2] = __ldg((const float4*)(b00 + ox)); f[3] = __ldg((const float4*)(b00 + ox + 4));
        f[4] = __ldg((const float4*)(b00 + oy)); f[5] = __ldg((const float4*)(b00 + oy + 4));
        f[6] = __ldg((const float4*)(b00 + oy + ox)); f[7] = __ldg((const float4*)(b00 + oy + ox + 4));
        off = act_off(r, k0 >> 3);
    };
    auto finish = [&](const float4 (&f)[8], const float (&w)[4], uint32_t off) {
        float x[8];
        x[0] = f[0].x * w[0] + f[2].x * w[1] + f[4].x * w[2] + f[6].x * w[3]; x[1] = f[0].y * w[0] + f[2].y * w[1] + f[4].y * w[2] + f[6].y * w[3];
        x[2] = f[0].z * w[0] + f[2].z * w[1] + f[4].z * w[2] + f[6].z * w[3]; x[3] = f[0].w * w[0] + f[2].w * w[1] + f[4].w * w[2] + f[6].w * w[3];
        x[4] = f[1].x * w[0] + f[3].x * w[1] + f[5].x * w[2] + f[7].x * w[3]; x[5] = f[1].y * w[0] + f[3].y * w[1] + f[5].y * w[2] + f[7].y * w[3];
        x[6] = f[1].z * w[0] + f[3].z * w[1] + f[5].z * w[2] + f[7].z * w[3]; x[7] = f[1].w * w[0] + f[3].w * w[1] + f[5].w * w[2] + f[7].w * w[3];
        uint4 hi, lo;
        split8(x, hi, lo);
        *(uint4*)(Ahi + off) = hi;
        if (PARITY) *(uint4*)(Alo + off) = lo;
    };
#pragma unroll 1
    for (int u = wwarp; u < n_units; u += 2 * NUM_OPND_WARPS) {
        float4 fa[8], fb[8];
        float wa[4], wb[4];
        uint32_t oa, ob = 0;
        const bool two = u + NUM_OPND_WARPS < n_units;
        issue(u, fa, wa, oa);
        if (two) issue(u + NUM_OPND_WARPS, fb, wb, ob);
        finish(fa, wa, oa);
        if (two) finish(fb, wb, ob);
    }
}

// View-combine of 32 accumulator columns: reduce-scatter over the NV adjacent lanes (rows) of a sample.  After it, lane j of the
// group holds 32/NV consecutive columns starting at the returned offset (summed over the NV views, pairwise order).
template <int NV>
__device__ __forceinline__ int combine_lanes(float (&v)[32], int lane) {
    int cnt = 32, offset = 0;
#pragma unroll
    for (int m = NV / 2; m >= 1; m >>= 1) {
        cnt >>= 1;
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < cnt) {
                const float send = upper ? v[i] : v[cnt + i];
                const float keep = upper ? v[cnt + i] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        }
        offset += upper ? cnt : 0;
    }
    return offset;
}
template <int NV>
__device__ __forceinline__ void combine_store(uint32_t* raw, const float* __restrict__ cb, int h0, int lane, float* dst_sample, bool write) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
    const int off = combine_lanes<NV>(v, lane);
    constexpr int CNT = 32 / NV;
    if (write) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) v[i] = v[i] * (1.0f / (float)NV) + __ldg(cb + h0 + off + i);
        float* dst = dst_sample + h0 + off;
        if constexpr (CNT >= 4) {
#pragma unroll
            for (int i = 0; i < CNT; i += 4) *(float4*)(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < CNT; ++i) dst[i] = v[i];
        }
    }
}

// Operand hand-off to the MMA issuer (leader CTA).  Remote mbarrier arrives are slow (~1 us each and they serialise),
// so the peer CTA first joins its 12 operand warps on a named barrier and sends ONE remote arrive; the leader's own
// warps arrive locally.  Leader barrier count = NUM_OPND_WARPS + 1.
__device__ __forceinline__ void worker_arrive(uint32_t bar_local, uint32_t bar_leader_remote, bool is_leader_cta, int wwarp, int lane) {
    fence_proxy_async();
    tc_fence_before();
    if (is_leader_cta) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_local);
    } else {
        asm volatile("bar.sync 2, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
        if (wwarp == 0 && lane == 0) mbar_arrive_remote(bar_leader_remote);
    }
}

#define TS(role, slot) do { if (a.dbg_ts && blockIdx.x < 2 && rd == 1 && lane == 0) a.dbg_ts[(blockIdx.x * 4 + (role)) * 64 + (slot)] = clock64(); } while (0)
#define TSW() do { if (a.dbg_ts && blockIdx.x < 2 && rd == 1 && wwarp == 0 && lane == 0 && tsn < 64) a.dbg_ts[(blockIdx.x * 4 + 1) * 64 + tsn++] = clock64(); } while (0)
// ---- the kernel ----------------------------------------------------------------------------------
template <bool PARITY, bool POST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) mlp_pair_kernel(const __grid_constant__ Args a) {
    using C = Cfg<PARITY>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool is_leader_cta = crank == 0;
    const uint32_t smem_base = smem_u32(smem);
    uint8_t* Ahi = smem + C::OFF_A_HI;
    uint8_t* Alo = smem + C::OFF_A_LO;
    RowTap* taps = (RowTap*)(smem + C::OFF_TAPS);
    const uint32_t bar_full = smem_base + C::OFF_BARS;             // NST: weight stage landed in THIS CTA
    const uint32_t bar_empty = bar_full + 8 * C::NST;              // NST: stage free (pair commit)
    const uint32_t bar_pfull = bar_empty + 8 * C::NST;             // NST: (leader) peer's stage landed
    const uint32_t bar_opnd = bar_pfull + 8 * C::NST;              // (leader) A operands of both CTAs ready
    const uint32_t bar_acc = bar_opnd + 8;                         // accumulators ready / operand buffers free
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + C::OFF_BARS + 8 * (3 * C::NST + 2));

    if ((smem_base & 1023u) != 0) { if (threadIdx.x == 0) atomicExch(a.err, 90); __trap(); }
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); mbar_init(bar_pfull + 8 * i, 1); }
        mbar_init(bar_opnd, NUM_OPND_WARPS + 1);
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t leader_opnd = map_to_cta(bar_opnd, 0);

    // both CTAs of a pair run the same number of rounds; CTA tile = 2 * pair_tile + rank
    const long long first = (long long)blockIdx.x, stride = (long long)gridDim.x;
    const long long n_rounds = (a.n_tiles + stride - 1) / stride;
    const long long total_uses = n_rounds * a.uses_per_tile;

    const int prod_idx = warp == 0 ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : -1));
    if (prod_idx >= 0) {
        // ===== weight producers (this CTA's half of every weight tile); stage st is always filled by producer st % NUM_PRODUCERS.
        //       2-SM TMA: both CTAs' copies complete_tx on the LEADER's full barrier, so the MMA issuer needs no software relay.
        const bool leader = elect_one();
        const int* table = a.tile_table + (size_t)crank * a.uses_per_tile;
        const uint32_t leader_full = map_to_cta(bar_full, 0);
        for (long long base = 0; base < ((a.dbg_skip & 16) ? 0 : total_uses); base += C::NST) {
            for (int st = prod_idx; st < C::NST; st += NUM_PRODUCERS) {
                const long long use = base + st;
                if (use >= total_uses) break;
                const int t = (int)(use % a.uses_per_tile);
                const uint32_t ph = (uint32_t)((use / C::NST) & 1);
                mbar_wait(bar_empty + 8 * st, ph ^ 1, a.err, 10);
                if (leader) {
                    if (is_leader_cta) mbar_arrive_expect_tx(bar_full + 8 * st, 2 * WTILE_BYTES);
                    const int row = __ldg(table + t) * 128;
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_base + st * WTILE_BYTES), "l"(&a.wmap), "r"(0), "r"(row), "r"(leader_full + 8 * st) : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 && is_leader_cta) {
        // ===== leader CTA: MMA issuer for the pair (converged warp, one elected lane)
        const bool leader = elect_one();
        uint32_t use = 0, it = 0;
        for (long long rd = 0; rd < n_rounds; ++rd) {
            for (int sidx = 0; sidx < a.n_steps; ++sidx, ++it) {
                const GemmStep gs = a.steps[sidx];
                const uint32_t idesc = make_idesc2(gs.n_width);
                if (a.dbg_skip & 64) mbar_poll(bar_opnd, it & 1, a.err, 20); else mbar_wait(bar_opnd, it & 1, a.err, 20);
                tc_fence_after();
                TS(0, 2 * sidx);
                for (int n2 = 0; n2 < gs.n_tiles; ++n2) {
                    const uint32_t d = tmem + (uint32_t)(gs.dst_col + 128 * n2);
                    for (int kb = 0; kb < gs.nkb; ++kb) {
                        {   // W_hi tile: A_hi*W_hi (+ A_lo*W_hi)
                            const uint32_t st = use % C::NST, ph = (use / C::NST) & 1;
                            if (!(a.dbg_skip & 16)) mbar_wait(bar_full + 8 * st, ph, a.err, 30);
                            tc_fence_after();
                            if (leader) {
                                const uint64_t bdesc = make_desc(smem_base + st * WTILE_BYTES, 16, 1024);
                                const uint64_t ahi = make_desc(smem_base + C::OFF_A_HI + kb * ACT_KB_BYTES, 16, 1024);
                                const uint64_t alo = make_desc(smem_base + C::OFF_A_LO + kb * ACT_KB_BYTES, 16, 1024);
#pragma unroll
                                for (int j = 0; j < ((a.dbg_skip & 8) ? 0 : 4); ++j) {
                                    umma2_bf16(d, ahi + 2 * j, bdesc + 2 * j, idesc, (gs.accumulate | kb | j) ? 1u : 0u);
                                    if (PARITY) umma2_bf16(d, alo + 2 * j, bdesc + 2 * j, idesc, 1u);
                                }
                                if (!(a.dbg_skip & 16)) umma2_commit_pair(bar_empty + 8 * st);
                            }
                            __syncwarp();
                            ++use;
                        }
                        if (PARITY) {   // W_lo tile: A_hi*W_lo
                            const uint32_t st = use % C::NST, ph = (use / C::NST) & 1;
                            if (!(a.dbg_skip & 16)) mbar_wait(bar_full + 8 * st, ph, a.err, 31);
                            tc_fence_after();
                            if (leader) {
                                const uint64_t bdesc = make_desc(smem_base + st * WTILE_BYTES, 16, 1024);
                                const uint64_t ahi = make_desc(smem_base + C::OFF_A_HI + kb * ACT_KB_BYTES, 16, 1024);
#pragma unroll
                                for (int j = 0; j < ((a.dbg_skip & 8) ? 0 : 4); ++j) umma2_bf16(d, ahi + 2 * j, bdesc + 2 * j, idesc, 1u);
                                umma2_commit_pair(bar_empty + 8 * st);
                            }
                            __syncwarp();
                            ++use;
                        }
                    }
                }
                if (leader) {
                    if (a.dbg_skip & 32) { tc::mbar_arrive(bar_acc); mbar_arrive_remote(map_to_cta(bar_acc, 1)); }   // experiment: software signal
                    else umma2_commit_pair(bar_acc);
                }
                __syncwarp();
                TS(0, 2 * sidx + 1);
            }
        }
    } else if (warp >= WORKER_WARP0) {
        // ===== workers (warps 4..11) + helpers (warps 12..15; prep and gather only).
        //       worker TMEM lanes 32q..32q+31: row = 32*(q&1)+lane, hidden half (q>>1) of this warp's N tile n2
        const int wwarp = warp - WORKER_WARP0, wt = threadIdx.x - WORKER_WARP0 * 32;
        const bool helper = wwarp >= NUM_WORKER_WARPS;
        const int q = warp & 3, n2 = (wwarp >> 2) & 1;
        const int r = 32 * (q & 1) + lane;
        uint32_t it = 0;
        for (long long rd = 0; rd < n_rounds; ++rd) {
            int tsn = 0; (void)tsn;
            const long long tile_raw = first + rd * stride;
            const bool live = tile_raw < a.n_tiles;
            const long long tile = live ? tile_raw : a.n_tiles - 1;
            if constexpr (!POST) {
                if (!(a.dbg_skip & 4)) prep_rows<PARITY>(a, tile, wt, Ahi, Alo, taps);
                TSW(); worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                                // -> lin_in
                for (int b = 0; b < a.n_blocks; ++b) {
                    if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 40); else mbar_wait(bar_acc, it & 1, a.err, 40); ++it; TSW();
                    tc_fence_after();
                    if (!(a.dbg_skip & 1)) gather_latent<PARITY>(a, wwarp, lane, Ahi, Alo, taps);
                    TSW(); worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                            // -> lin_z[b]
                    if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 41); else mbar_wait(bar_acc, it & 1, a.err, 41); ++it; TSW();
                    tc_fence_after();
                    if (!helper && !(a.dbg_skip & 2)) epilogue_to_A<PARITY>(tmem, COL_X, a.bias + (size_t)b * HID, Ahi, Alo, q, lane, n2);
                    TSW(); worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                            // -> fc_0[b]
                    if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 42); else mbar_wait(bar_acc, it & 1, a.err, 42); ++it; TSW();
                    tc_fence_after();
                    if (!helper && !(a.dbg_skip & 2)) epilogue_to_A<PARITY>(tmem, COL_NET, a.bias + (size_t)(a.n_blocks + b) * HID, Ahi, Alo, q, lane, n2);
                    TSW(); worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                            // -> fc_1[b]
                }
                if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 43); else mbar_wait(bar_acc, it & 1, a.err, 43); ++it; TSW();
                tc_fence_after();
                // combine: mean over the NV adjacent rows (lanes) of each sample, sequential like torch.mean (resnetfc.py:148-151)
                const float* cb = a.bias + (size_t)(2 * a.n_blocks) * HID;
                const long long smp = tile * a.spv + r / a.NV;                   // sample within the sub-batch
                const bool wr = live && smp < a.n_samples;
                float* dst_sample = a.xc + (size_t)(wr ? smp : 0) * HID;
#pragma unroll 1
                for (int c32 = 0; c32 < ((helper || (a.dbg_skip & 128)) ? 0 : 4); ++c32) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(COL_X + 128 * n2 + 32 * c32), v);
                    const int h0 = 256 * n2 + 128 * (q >> 1) + 32 * c32;
                    switch (a.NV) {
                        case 1: combine_store<1>(v, cb, h0, lane, dst_sample, wr); break;
                        case 2: combine_store<2>(v, cb, h0, lane, dst_sample, wr); break;
                        case 4: combine_store<4>(v, cb, h0, lane, dst_sample, wr); break;
                        case 8: combine_store<8>(v, cb, h0, lane, dst_sample, wr); break;
                        case 16: combine_store<16>(v, cb, h0, lane, dst_sample, wr); break;
                        default: combine_store<32>(v, cb, h0, lane, dst_sample, wr); break;
                    }
                }
                tc_fence_before();
            } else {
                // load x_c: fp32 residual -> TMEM X, relu(x_c) -> A operand
                long long smp = tile * ROWS + r;
                if (smp >= a.n_samples) smp = a.n_samples - 1;
#pragma unroll 1
                for (int c32 = 0; c32 < (helper ? 0 : 4); ++c32) {
                    const int h0 = 256 * n2 + 128 * (q >> 1) + 32 * c32;
                    const float4* src = (const float4*)(a.xc + (size_t)smp * HID + h0);
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 f = __ldg(src + i);
                        v[4 * i] = __float_as_uint(f.x); v[4 * i + 1] = __float_as_uint(f.y);
                        v[4 * i + 2] = __float_as_uint(f.z); v[4 * i + 3] = __float_as_uint(f.w);
                    }
                    tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(COL_X + 128 * n2 + 32 * c32), v);
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float x[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[8 * c8 + i]), 0.0f);
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        const uint32_t off = act_off(r, (h0 >> 3) + c8);
                        *(uint4*)(Ahi + off) = hi;
                        if (PARITY) *(uint4*)(Alo + off) = lo;
                    }
                }
                worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                                // -> fc_0 of the first post block
                for (int b = 0; b < a.n_blocks; ++b) {
                    if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 50); else mbar_wait(bar_acc, it & 1, a.err, 50); ++it;
                    tc_fence_after();
                    if (!helper) epilogue_to_A<PARITY>(tmem, COL_NET, a.bias + (size_t)(a.n_blocks + 1 + b) * HID, Ahi, Alo, q, lane, n2);
                    TSW(); worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                            // -> fc_1[b]
                    if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 51); else mbar_wait(bar_acc, it & 1, a.err, 51); ++it;
                    tc_fence_after();
                    if (!helper) epilogue_to_A<PARITY>(tmem, COL_X, a.bias + (size_t)(b + 1) * HID, Ahi, Alo, q, lane, n2);
                    worker_arrive(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                            // -> next fc_0 / lin_out
                }
                if (a.dbg_skip & 64) mbar_poll(bar_acc, it & 1, a.err, 52); else mbar_wait(bar_acc, it & 1, a.err, 52); ++it;                     // lin_out (N=32): outputs 0..3 in columns COL_NET..+3, lanes 0..63
                tc_fence_after();
                if (!helper && q < 2 && n2 == 0) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)COL_NET, v);
                    const long long s_loc = tile * ROWS + r;
                    if (live && s_loc < a.n_samples) {
                        const float4 bo = __ldg((const float4*)(a.bias + (size_t)(2 * a.n_blocks + 1) * HID));
                        const float x0 = __uint_as_float(v[0]) + bo.x, x1 = __uint_as_float(v[1]) + bo.y;
                        const float x2 = __uint_as_float(v[2]) + bo.z, x3 = __uint_as_float(v[3]) + bo.w;
                        ((float4*)a.out)[a.s_begin + s_loc] = make_float4(1.0f / (1.0f + expf(-x0)), 1.0f / (1.0f + expf(-x1)),
                                                                         1.0f / (1.0f + expf(-x2)), fmaxf(x3, 0.0f));
                    }
                }
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <bool PARITY, bool POST>
cudaError_t launch(const Args& a, int grid, cudaStream_t st) {
    auto kern = mlp_pair_kernel<PARITY, POST>;
    constexpr int smem = Cfg<PARITY>::SMEM_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    g_launches++;
    kern<<<grid, NUM_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tc2

// host: the pair kernel shares tc_pack_weights' tile stream; this builds the per-rank tile tables and launches
cudaError_t tc2_prepare(TcState& t, const MlpDev& m, bool parity_table, cudaStream_t st);

cudaError_t tc2_query(TcState& t, const SceneDev& s, const MlpDev& m, const QueryArgs& q, bool parity, int num_sms,
                      cudaStream_t st) {
    using namespace tc2;
    const int NV = s.NV;
    if (NV > 32 || (32 % NV)) { snprintf(t.why, sizeof(t.why), "NV=%d views (tcgen05 path needs NV in {1,2,4,8,16,32})", NV); return cudaErrorNotSupported; }
    if (s.L != m.d_latent || (s.L % 256)) { snprintf(t.why, sizeof(t.why), "pair kernel needs d_latent %% 256 == 0 (got %d)", s.L); return cudaErrorNotSupported; }
    const int d_in = 3 + 6 * s.num_freqs + 3 + 1 + 2 * s.num_freqs;
    if (d_in != m.d_in) { snprintf(t.why, sizeof(t.why), "positional code gives d_in=%d but lin_in expects %d", d_in, m.d_in); return cudaErrorNotSupported; }
    const long long total = (long long)q.SB * q.n_per_sb;
    const int kbz = m.d_latent / KBLK, kbh = HID / KBLK;

    // per-rank tile tables: ring-use order of CTA rank r = for each step, for n2, for kb: tile (2*n2 + r) of that layer
    if (!t.table2 || t.table2_parity != (int)parity) {
        std::vector<int> tab[2];
        int layer_pair0 = 0;
        auto layer = [&](int nkb, int n_mt, int n_tiles) {
            for (int r = 0; r < 2; ++r)
                for (int n2 = 0; n2 < n_tiles; ++n2)
                    for (int kb = 0; kb < nkb; ++kb) {
                        const int pair = layer_pair0 + (2 * n2 + r) * nkb + kb;
                        tab[r].push_back(2 * pair);
                        if (parity) tab[r].push_back(2 * pair + 1);
                    }
            layer_pair0 += n_mt * nkb;
        };
        layer(1, 4, 2);                                              // lin_in
        for (int b = 0; b < t.n_pre; ++b) { layer(kbz, 4, 2); layer(kbh, 4, 2); layer(kbh, 4, 2); }
        t.uses2_pre = (int)tab[0].size();
        for (int b = 0; b < t.n_post; ++b) { layer(kbh, 4, 2); layer(kbh, 4, 2); }
        layer(kbh, 2, 1);                                            // lin_out packed as 2 M-tiles (second is zeros)
        t.uses2_post = (int)tab[0].size() - t.uses2_pre;
        std::vector<int> flat;                                       // [pre r0][pre r1][post r0][post r1]
        for (int r = 0; r < 2; ++r) flat.insert(flat.end(), tab[r].begin(), tab[r].begin() + t.uses2_pre);
        for (int r = 0; r < 2; ++r) flat.insert(flat.end(), tab[r].begin() + t.uses2_pre, tab[r].end());
        if (!t.table2) TCK(cudaMalloc((void**)&t.table2, 4096 * sizeof(int)));
        if (flat.size() > 4096) return cudaErrorInvalidValue;
        TCK(cudaMemcpyAsync(t.table2, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        TCK(cudaStreamSynchronize(st));                              // `flat` is a stack-lifetime host buffer
        t.table2_parity = (int)parity;
    }
    if (t.max_grid2 == 0) {
        auto kern = mlp_pair_kernel<true, false>;
        TCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM_BYTES));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = Cfg<true>::SMEM_BYTES;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = num_sms / 2; }
        t.max_grid2 = n > 0 ? 2 * n : 2;
    }
    const int grid_cap = t.max_grid2 < (num_sms / 2) * 2 ? t.max_grid2 : (num_sms / 2) * 2;
    const long long sub = t.sub_batch > 0 ? t.sub_batch : 524288;
    const size_t need = (size_t)(((sub + ROWS - 1) / ROWS) * ROWS) * HID * sizeof(float);
    if (need > t.scratch_bytes) {
        if (t.scratch) cudaFree(t.scratch);
        t.scratch = nullptr; t.scratch_bytes = 0;
        TCK(cudaMalloc(&t.scratch, need));
        t.scratch_bytes = need;
    }
    Args pre{}, post{};
    if (!t.wmap_ok) { snprintf(t.why, sizeof(t.why), "cuTensorMapEncodeTiled unavailable"); return cudaErrorNotSupported; }
    pre.wmap = post.wmap = t.wmap;
    pre.s = s; pre.q = q; post.s = s; post.q = q;
    pre.wstream = post.wstream = (const uint8_t*)t.wpack;
    pre.tile_table = t.table2; post.tile_table = t.table2 + 2 * t.uses2_pre;
    pre.uses_per_tile = t.uses2_pre; post.uses_per_tile = t.uses2_post;
    pre.bias = t.bias; post.bias = t.bias + t.bias_post_off;
    pre.n_blocks = t.n_pre; post.n_blocks = t.n_post;
    int n = 0;
    pre.steps[n++] = GemmStep{1, 2, 256, COL_X, 0};
    for (int b = 0; b < t.n_pre; ++b) {
        pre.steps[n++] = GemmStep{(short)kbz, 2, 256, COL_X, 1};
        pre.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_NET, 0};
        pre.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_X, 1};
    }
    pre.n_steps = n;
    n = 0;
    for (int b = 0; b < t.n_post; ++b) {
        post.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_NET, 0};
        post.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_X, 1};
    }
    post.steps[n++] = GemmStep{(short)kbh, 1, 32, COL_NET, 0};
    post.n_steps = n;
    pre.NV = post.NV = NV;
    pre.spv = post.spv = ROWS / NV;
    pre.xc = post.xc = (float*)t.scratch;
    pre.out = post.out = q.out;
    pre.err = post.err = t.err_flag;
    pre.n_total = post.n_total = total;
    pre.dbg_skip = t.dbg_skip; post.dbg_skip = 0;
    static long long* dbg_ts = nullptr;
    if ((t.dbg_skip & 512) && !dbg_ts) { TCK(cudaMallocManaged((void**)&dbg_ts, 8 * 64 * sizeof(long long))); memset(dbg_ts, 0, 8 * 64 * sizeof(long long)); }
    pre.dbg_ts = (t.dbg_skip & 512) ? dbg_ts : nullptr; post.dbg_ts = nullptr;
    t.ms_pre = t.ms_post = 0.f;
    if (t.timing && !t.ev[0]) for (int i = 0; i < 4; ++i) TCK(cudaEventCreate(&t.ev[i]));
    for (long long s0 = 0; s0 < total; s0 += sub) {
        const long long ns = total - s0 < sub ? total - s0 : sub;
        pre.s_begin = post.s_begin = s0;
        pre.n_samples = post.n_samples = ns;
        pre.n_tiles = (ns + pre.spv - 1) / pre.spv;
        post.n_tiles = (ns + ROWS - 1) / ROWS;
        const long long g1 = ((pre.n_tiles + 1) / 2) * 2, g2 = ((post.n_tiles + 1) / 2) * 2;
        const int grid1 = (int)(g1 < grid_cap ? g1 : grid_cap), grid2 = (int)(g2 < grid_cap ? g2 : grid_cap);
        if (t.timing) TCK(cudaEventRecord(t.ev[0], st));
        if (parity) TCK((launch<true, false>(pre, grid1, st))); else TCK((launch<false, false>(pre, grid1, st)));
        if (t.timing) TCK(cudaEventRecord(t.ev[1], st));
        if (parity) TCK((launch<true, true>(post, grid2, st))); else TCK((launch<false, true>(post, grid2, st)));
        if (t.timing) {
            TCK(cudaEventRecord(t.ev[2], st));
            TCK(cudaEventSynchronize(t.ev[2]));
            float x = 0.f, y = 0.f;
            TCK(cudaEventElapsedTime(&x, t.ev[0], t.ev[1]));
            TCK(cudaEventElapsedTime(&y, t.ev[1], t.ev[2]));
            t.ms_pre += x; t.ms_post += y;
        }
    }
    if (pre.dbg_ts) {
        TCK(cudaStreamSynchronize(st));
        for (int cta = 0; cta < 2; ++cta) {
            const long long t0 = pre.dbg_ts[(cta * 4 + 1) * 64];
            fprintf(stderr, "[ts] cta %d worker stamps (arriving, woke, arriving, ...) cycles since first:", cta);
            for (int i = 0; i < 22; ++i) fprintf(stderr, " %lld", pre.dbg_ts[(cta * 4 + 1) * 64 + i] - t0);
            fprintf(stderr, "\n");
            if (cta == 0) {
                fprintf(stderr, "[ts] cta 0 mma stamps (opnd-ready, committed per step), same origin:");
                for (int i = 0; i < 20; ++i) fprintf(stderr, " %lld", pre.dbg_ts[i] - t0);
                fprintf(stderr, "\n");
            }
        }
    }
    return cudaSuccess;
}
