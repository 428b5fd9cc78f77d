// tcgen05 path of the per-sample network (PixelNeRF.forward + ResnetFC.forward): the CTA-PAIR kernel (cta_group::2).
//
// Kinds of one kernel template: FUSED (default: per 64 samples, NV tiles of 64 sample-view rows -- lin_in, per block the Y_b
// gather, fc_0, fc_1, view mean -- then one tile of 64 samples -- remaining blocks, lin_out -- with the view-combined
// activations in an L2-resident per-CTA slab), PRE / POST (the same two halves as separate launches with an HBM scratch,
// bit-identical: A/B reference), ZMAP (the once-per-scene Y maps).
//
// Why a CTA pair (tools/mma_probe.cu, round 1): in a single-CTA kernel with 64 rows per CTA every MMA is 128x64x16 and reads
// 6 KiB of shared memory for 32 cycles of math, and every SM streams the complete weight set per 64 rows -- shared-memory
// bandwidth, not the tensor pipe, bounds it.  Here two CTAs of a cluster form one UMMA:
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
// adjacent lanes; lin_out (4 outputs) runs on the CUDA cores in the POST epilogue: 512 FMAs per thread on the last fc_1's
// accumulator, partial sums of the four warps that hold a row added in a fixed order through shared memory.
//
// lin_z is hoisted out of the per-sample work: grid_sample(bilinear) and lin_z are both linear, so
//   lin_z[b](bilinear(latent, uv)) == bilinear(lin_z[b](latent), uv)      (the four tap weights sum to 1)
// and Y_b = W_z[b] . latent is computed ONCE per (scene, weights) for every latent pixel by the ZMAP variant of this kernel
// (fp16x3, fp32 maps [b][pixel][512]).  The PRE kernel then gathers Y_b bilinearly and adds it to the fp32 residual in TMEM:
// one third of the per-sample-view GEMM work and weight streaming disappears.  The biases that enter a block (b_in / b_fc1[b-1],
// b_z[b]) are folded into Y_b; b_fc0 is added by the net epilogue, b_fc1 of the last block when the combined x_c is written.
//
// Software pipeline (everything in place: parity mode has neither spare shared memory nor spare TMEM):
//   * GEMM steps run K-block-outer; the issuer commits bar_afree[kb] after the last MMA that reads operand K block kb;
//   * the gathered Y_b rows are staged as fp32 INSIDE the operand buffers (hi slot + lo slot of each 8-channel chunk),
//     K block by K block behind bar_afree while fc_1 of the previous block still runs (loads are issued before the wait);
//   * epilogues run in two halves (K blocks 0..3 / 4..7), one operand barrier each, so the next GEMM starts on half 0;
//   * across tiles: under the last fc_1 the helper warps compute the next tile's taps and lin_in features and everyone gathers
//     its Y_0 (K blocks 1..7); lin_in of the next tile is handed off right after the view-combine -- or (early_lin) issued
//     straight behind that fc_1 into the OTHER TMEM half: x and net swap halves from tile to tile, so the next tile's lin_in does
//     not have to wait for this tile's x to be combined;
//   * across rounds (warm_rounds, default): the POST tile's last fc_1 hands off to the next round's first PRE tile in the same way
//     (lin_out off the tensor pipe is what frees the operand buffers for it), so only the first tile of a launch starts cold;
//   * accumulator halves: the first epilogue half reads only N tile 0 of the accumulator (hidden units 0..255 = K blocks 0..3 of
//     the next layer), so the last `tail` K blocks of every full-width step are issued N-TILE-OUTER -- (kb, n0) for the tail,
//     commit bar_acc0, then (kb, n1), commit bar_acc -- and that epilogue half runs under the n1 tail of the same GEMM.  It writes
//     operand K blocks 0..3, which the running GEMM released long ago (tail <= 4); only the second half waits for the full step.
// mbarrier parity waits are only safe if the waiter cannot be a whole phase late -- see the notes at the helper-warp code.
// Operands are fp16 hi/lo pairs of W_SCALE-scaled weights and unscaled activations (mlp_tc.cu): every accumulator in TMEM holds
// W_SCALE * value; the Y maps and the x_c load carry the same factor, the epilogues multiply by W_INV where they add their bias.
//
// Reference semantics: src/models/resnetfc.py:61-69,129-159; src/models/pixelnerf.py:91-143; src/models/image_encoder.py:97-146.
#include "mlp_tc.h"

#include <algorithm>

namespace tc2 {

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
using tc::tmem_st32_issue;
using tc::tmem_st_wait;
using tc::make_desc;
using tc::split8;
using tc::split1;
using tc::W_SCALE;
using tc::W_INV;
using tc::sample_point;

constexpr int ROWS = 64;                    // rows per CTA (128 per pair)
constexpr int HID = 512;
constexpr int KBLK = 64;
constexpr int WTILE_BYTES = 128 * KBLK * 2; // 16 KiB: 128 hidden rows x 64 k, K-major SWIZZLE_128B (same packing as mlp_tc.cu)
constexpr int ACT_KB_BYTES = ROWS * 128;    // 8 KiB per 64-wide K block of the activation operand
constexpr int ACT_BYTES = ACT_KB_BYTES * (HID / KBLK);   // 64 KiB per fp16 copy
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

// Bilinear tap set of one sample-view row (16 B, double-buffered: the helper warps compute the next tile's taps during the
// current tile's last block).  ex = 1 - wx and ey = 1 - wy bit-exactly equal image_encoder's (x0 + 1) - x: x - floor(x) is exact.
struct Tap {
    int pix_dxy;       // pixel index of tap (0,0) (view base included) | dx << 30 | dy << 31 (dx/dy = 0 when clamped at the border)
    float wx, wy;
    int pad;
};

struct GemmStep {
    short nkb;         // K blocks of 64
    short n_tiles;     // N tiles (of n_width hidden units) = weight tiles per CTA per K block
    short n_width;     // UMMA N: 256
    short dst_col;     // TMEM column base
    short accumulate;
    short release;     // commit the per-K-block "A operand free" barriers (a gather into A overlaps / follows the step); 2 = lin_in;
                       // 3 = fc_1 of the last PRE block: only when another PRE tile follows (its early gather consumes the phases)
    short tail;        // the last `tail` K blocks are issued N-tile-outer (accumulator N tile 0 completes early -> bar_acc0)
};

struct Args {
    CUtensorMap wmap;           // packed weight stream as rows of 128 B; one box = one 16 KiB tile
    SceneDev s;
    QueryArgs q;
    const uint8_t* wstream;     // packed weight tiles (16 KiB units), shared with mlp_tc.cu's packing
    const int* tile_table;      // [2][uses_per_tile]: 16 KiB tile index of the i-th ring use of CTA rank r
    const float* bias;
    const float* bias2;         // pair-kernel PRE rows: [0,n_pre) bias folded into Y_b, row n_pre = b_fc1[n_pre-1] (added at the combine)
    GemmStep steps[MAX_STEPS];
    int n_steps, n_blocks, uses_per_tile;
    // FUSED launch: a round = ppr PRE tiles (ppr * spv = 64 * pts samples) followed by pts POST tiles over the same samples; the view-combined
    // activations x_c go through a per-CTA slab of `xc` (64 * pts rows x 512 fp32, L2 resident) instead of a sub-batch sized HBM scratch.
    // Only the first PRE tile of a round starts cold (the POST tiles own the operand buffers), so pts > 1 amortises that start.
    GemmStep steps_post[2 * DINER_MAX_BLOCKS + 1];
    int n_steps_post, n_blocks_post, uses_post, ppr, pts;
    const int* tile_table_post;
    const float* bias_post;
    long long s_begin, n_samples, n_total, n_tiles;   // n_tiles counts 64-row CTA tiles
    int NV, spv;                // NV = views per sample padded to a power of two (rows of a sample are NV adjacent rows), spv = 64 / NV
    int NV_real;                // the scene's view count: rows of the padding views repeat view 0 and are masked out of the combine
    float* xc;                  // [sample][512] fp32 view-combined activations (sub-batch relative)
    float* out;
    float* zmap;                // Y maps [block][pixel][512] fp32: PRE reads them, ZMAP writes them
    long long zmap_stride;      // floats per block map = n_pix * 512
    long long n_pix;            // ZMAP: latent pixels (SB*NV*Hl*Wl)
    int* err;
    long long* dbg_ts;          // profiling: clock64 stamps of pair 0 in round 1 ([cta][role][slot])
    int perm_w, perm_h;         // FUSED: the rays of a scene are a row-major perm_h x perm_w image -> process them in 16 x 16 pixel tiles
                                // (the Y-map lines gathered by neighbouring rays stay in L2); 0 = process in the caller's order
    int early_worker_kb_hi;     // next-tile Y_0 gather under the last fc_1: K blocks 1..this on the workers, the rest on the helpers
    int worker_kb_hi;           // Y_b gather inside a tile: K blocks 0..this on the workers (released before the N-outer tail), the rest on the helpers
    const float* w_out;         // lin_out weights, fp32 [4][512] (the output layer runs on the CUDA cores in the POST epilogue)
    const float* b_out;
    int warm_rounds;            // FUSED: the first PRE tile of the next round is prepared under the POST tile's last fc_1 (needs early_lin, even ppr)
    int early_lin;              // PRE tiles: issue the next tile's lin_in right behind the last fc_1 into the other TMEM half (x / net ping-pong)
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
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_commit_pair(uint32_t bar) {     // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// A K-major, B K-major, D f32 (1 << 4), fp16 inputs (a_format = b_format = 0; bf16 would be 1 << 7 | 1 << 10); M = 128 over the pair
__host__ __device__ constexpr uint32_t make_idesc2(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of (row r, 8-wide k chunk kc = k/8) in the K-major SWIZZLE_128B activation operand
__device__ __forceinline__ uint32_t act_off(int r, int kc) {
    return (uint32_t)(kc >> 3) * ACT_KB_BYTES + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((kc & 7) ^ (r & 7)) << 4);
}

// Processing order -> sample index.  Identity unless the caller declared the ray list a row-major image (Args::perm_w): then the
// n-th ray processed is the n-th pixel in 16 x 16 tile order.  Only the ORDER of the work changes; every buffer keeps its layout.
constexpr int PERM_T = 16;
__device__ __forceinline__ long long map_sample(const Args& a, long long s) {
    if (a.perm_w == 0) return s;
    const long long K = a.q.K, ray = s / K, k = s - ray * K;
    const long long img = (long long)a.perm_w * a.perm_h, sc = ray / img, i = ray - sc * img;
    const int tiles_x = a.perm_w / PERM_T;
    const long long tile = i / (PERM_T * PERM_T);
    const int within = (int)(i - tile * (PERM_T * PERM_T));
    const long long ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const long long pix = (ty * PERM_T + within / PERM_T) * a.perm_w + tx * PERM_T + within % PERM_T;
    return (sc * img + pix) * K + k;
}

// The lo-operand region exists in both modes: the gathered fp32 Y rows are staged across (A_hi, A_lo) chunk slots in place
// (8 floats of (row, k-chunk) = 16 B in the hi slot + 16 B in the lo slot) before the epilogue turns them into operands.
template <bool PARITY> struct Cfg {
    static constexpr int NST = 6;
    static constexpr int OFF_A_HI = NST * WTILE_BYTES;
    static constexpr int OFF_A_LO = OFF_A_HI + ACT_BYTES;
    static constexpr int OFF_TAPS = OFF_A_LO + ACT_BYTES;
    static constexpr int OFF_BARS = OFF_TAPS + 2 * ROWS * (int)sizeof(Tap);
    static constexpr int SMEM_BYTES = OFF_BARS + 256;
};

// ---- worker building blocks ----------------------------------------------------------------------
// TMEM region (this warp's N tile: 128 columns) + per-column bias -> relu -> fp16 hi/lo chunks of the A operand
// 32 accumulator columns of row r (hidden h0..h0+31): relu(acc * W_INV + bias) -> four 16-byte K-major chunks (hi / lo)
template <bool PARITY>
__device__ __forceinline__ void convert32(const uint32_t* v, const float* __restrict__ bias, int h0, int r, uint8_t* Ahi, uint8_t* Alo) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        const float4 b0 = __ldg((const float4*)(bias + h0 + 8 * c8)), b1 = __ldg((const float4*)(bias + h0 + 8 * c8 + 4));
        float x[8];
        x[0] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 0]), W_INV, b0.x), 0.0f); x[1] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 1]), W_INV, b0.y), 0.0f);
        x[2] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 2]), W_INV, b0.z), 0.0f); x[3] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 3]), W_INV, b0.w), 0.0f);
        x[4] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 4]), W_INV, b1.x), 0.0f); x[5] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 5]), W_INV, b1.y), 0.0f);
        x[6] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 6]), W_INV, b1.z), 0.0f); x[7] = fmaxf(fmaf(__uint_as_float(v[8 * c8 + 7]), W_INV, b1.w), 0.0f);
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = act_off(r, (h0 >> 3) + c8);
        *(uint4*)(Ahi + off) = hi;
        if (PARITY) *(uint4*)(Alo + off) = lo;
    }
}
// Epilogues run in two HALVES so that the next GEMM (K-block-outer order) starts on K blocks 0..3 while the workers still
// convert K blocks 4..7.  In half h every worker warp (q = TMEM lane quarter, j = 0/1) owns 64 accumulator columns:
//   TMEM lanes 32q..32q+31 (row = 32*(q&1) + lane), columns colbase + 128*h + 64*j + [0,64)
//   = hidden units 256*h + 128*(q>>1) + 64*j + [0,64) = K block 4*h + 2*(q>>1) + j of the next layer's A operand.
template <bool PARITY>
__device__ __forceinline__ void epilogue_half(uint32_t tmem, int colbase, const float* __restrict__ bias, uint8_t* Ahi,
                                              uint8_t* Alo, int q, int lane, int j, int h) {
    const int r = 32 * (q & 1) + lane;
    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(colbase + 128 * h + 64 * j);
    const int hb = 256 * h + 128 * (q >> 1) + 64 * j;
    uint32_t va[32], vb[32];
    tmem_ld32_issue(t0, va);
    tmem_ld32_issue(t0 + 32, vb);
    tmem_ld_wait();
    convert32<PARITY>(va, bias, hb, r, Ahi, Alo);
    convert32<PARITY>(vb, bias, hb + 32, r, Ahi, Alo);
}

// Output layer on the CUDA cores (POST epilogue): this thread's 64 columns of x in half h -> relu(x) . W_out[0..3], accumulated
// in column order into acc.  (The layer is 4 outputs wide: as a GEMM step it kept the whole A operand and the tensor pipe busy for
// ~1/30 of a tile's work and put a staging + hand-off between the last fc_1 and the result; as 512 FMAs per thread it leaves the
// operand buffers free for the next round's first PRE tile as soon as the last fc_1 has read them.)
__device__ __forceinline__ void lin_out_dot32(const uint32_t* v, const float* __restrict__ bias, const float* __restrict__ w_out, int h0, float (&acc)[4]) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        const float4 b = __ldg((const float4*)(bias + h0 + 4 * c4));
        float x[4];
        x[0] = fmaxf(fmaf(__uint_as_float(v[4 * c4 + 0]), W_INV, b.x), 0.0f); x[1] = fmaxf(fmaf(__uint_as_float(v[4 * c4 + 1]), W_INV, b.y), 0.0f);
        x[2] = fmaxf(fmaf(__uint_as_float(v[4 * c4 + 2]), W_INV, b.z), 0.0f); x[3] = fmaxf(fmaf(__uint_as_float(v[4 * c4 + 3]), W_INV, b.w), 0.0f);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const float4 w = __ldg((const float4*)(w_out + o * HID + h0 + 4 * c4));
            acc[o] = fmaf(x[3], w.w, fmaf(x[2], w.z, fmaf(x[1], w.y, fmaf(x[0], w.x, acc[o]))));
        }
    }
}
__device__ __forceinline__ void lin_out_half(uint32_t tmem, int colx, const float* __restrict__ bias, const float* __restrict__ w_out, int q, int lane,
                                             int j, int h, float (&acc)[4]) {
    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(colx + 128 * h + 64 * j);
    const int hb = 256 * h + 128 * (q >> 1) + 64 * j;
    uint32_t va[32], vb[32];
    tmem_ld32_issue(t0, va);
    tmem_ld32_issue(t0 + 32, vb);
    tmem_ld_wait();
    lin_out_dot32(va, bias, w_out, hb, acc);
    lin_out_dot32(vb, bias, w_out, hb + 32, acc);
}

// PRE prep: PARTS threads per row (wt = row + 64 * part).  Camera transform, projection, nearest depth; TAPS: the row's
// bilinear tap set for the Y-map gathers; FEAT: positional encodings etc. -> the lin_in A operand (K block 0).
template <bool PARITY, int PARTS, bool TAPS, bool FEAT>
__device__ __noinline__ void prep_rows(const Args& a, long long tile, int wt, uint8_t* Ahi, uint8_t* Alo, Tap* taps) {
    const SceneDev& s = a.s;
    const int r = wt & 63, part = wt >> 6;
    long long smp = a.s_begin + tile * a.spv + r / a.NV;
    if (smp >= a.n_total) smp = a.n_total - 1;
    smp = map_sample(a, smp);
    const int v = (r % a.NV) < a.NV_real ? (r % a.NV) : 0;          // padding views (view count not a power of two) repeat view 0
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
    const float u = project_axis(xc, zc, __ldg(s.focal + sv * 2), __ldg(s.cxy + sv * 2), s.imgW);
    const float w = project_axis(yc, zc, __ldg(s.focal + sv * 2 + 1), __ldg(s.cxy + sv * 2 + 1), s.imgH);
    if (TAPS && part == 0) {
        const LatTaps t = latent_taps(s, u, w);
        Tap rt;
        rt.pix_dxy = (sv * s.Hl * s.Wl + t.o00) | (t.o01 != t.o00 ? (1 << 30) : 0) | (t.o10 != t.o00 ? (int)(1u << 31) : 0);
        float x = unnormalize(__fmul_rn(u, s.lat_sx), (float)s.Wl), y = unnormalize(__fmul_rn(w, s.lat_sy), (float)s.Hl);
        x = fminf(fmaxf(x, 0.0f), (float)(s.Wl - 1));
        y = fminf(fmaxf(y, 0.0f), (float)(s.Hl - 1));
        if (!(x == x)) x = 0.0f;
        if (!(y == y)) y = 0.0f;
        rt.wx = x - floorf(x); rt.wy = y - floorf(y);
        rt.pad = 0;
        taps[r] = rt;
    }
    if (FEAT) {
        // lin_in input columns of this row (same values and order as feature_elem, mlp_simt.cu):
        //   [x_c y_c z_c | per frequency k: sin(f_k xyz) cos(f_k xyz) | dir_c xyz | dd | per frequency k: sin(f_k dd) cos(f_k dd) | 0 ...]
        // The PARTS threads of a row split the FREQUENCIES: the 8 sin evaluations of one frequency are independent, so they pipeline
        // (one element at a time through a branchy selector cost ~20 k cycles per tile on the four helper warps).
        const int F = s.num_freqs, npe = 6 * F, d_in = 3 + npe + 3 + 1 + 2 * F;
        auto put = [&](int e, float val) {
            __half hi, lo;
            split1(val, hi, lo);
            const uint32_t off = act_off(r, e >> 3) + (uint32_t)(e & 7) * 2u;
            *(__half*)(Ahi + off) = hi;
            if (PARITY) *(__half*)(Alo + off) = lo;
        };
        const float dd = __fsub_rn(lookup_depth(s, sv, u, w), zc);
#pragma unroll 1
        for (int k = part; k < F; k += PARTS) {
            const float f = s.freqs[k];
            const float v0 = pe_sin(xc, f), v1 = pe_sin(yc, f), v2 = pe_sin(zc, f), v3 = pe_cos(xc, f), v4 = pe_cos(yc, f), v5 = pe_cos(zc, f);
            const float v6 = pe_sin(dd, f), v7 = pe_cos(dd, f);
            const int e0 = 3 + 6 * k, e1 = 7 + npe + 2 * k;
            put(e0, v0); put(e0 + 1, v1); put(e0 + 2, v2); put(e0 + 3, v3); put(e0 + 4, v4); put(e0 + 5, v5);
            put(e1, v6); put(e1 + 1, v7);
        }
        if (part == PARTS - 1) {
            rotate_to_cam(p, dx, dy, dz, dxc, dyc, dzc);
            put(0, xc); put(1, yc); put(2, zc);
            put(3 + npe, dxc); put(4 + npe, dyc); put(5 + npe, dzc);
            put(6 + npe, dd);
            for (int e = d_in; e < KBLK; ++e) put(e, 0.0f);
        }
    }
}

// PRE gather: bilinear Y_b (= lin_z[b] of the latent map) rows of a tile -> fp32 staging in the (A_hi, A_lo) chunk slots, for
// the K blocks [kb_lo, kb_hi], in K-BLOCK ORDER and overlapped with the GEMM that is still reading the operand buffers: before
// touching K block kb the warp waits on bar_afree[kb], which the MMA issuer commits right after the last MMA that reads it.
// A warp waits on the barriers of the K blocks up to its last own one, in order: par0 is the parity of
// bar_afree[0], par1 that of the others (K block 0 has one more release per tile: it also carries the lin_in features).
// One pass = 4 rows x 64 channels (one K block): lane -> row 4*(p%16) + lane/8, 8 channels (lane%8): 32-byte loads per tap;
// channels 0..3 of the chunk go to the hi slot, 4..7 to the lo slot (the epilogue thread that owns the row reads them back).
// Work split: the 8 worker warps take K blocks <= WORKER_KB_HI (released in the K-block-outer part of the running GEMM), the 4
// helper warps the rest, which are released in its N-tile-outer tail -- the workers run the first epilogue half meanwhile.
__device__ __noinline__ void gather_y(const Args& a, const float* __restrict__ ymap, int wwarp, int lane, uint8_t* Ahi,
                                         uint8_t* Alo, const Tap* taps, int kb_lo, int kb_hi, uint32_t bar_afree, uint32_t par0,
                                         uint32_t par1, int WORKER_KB_HI = HID / KBLK - 1) {
    const SceneDev& s = a.s;
    constexpr int PASSES_PER_KB = ROWS / 4;         // 16
    auto issue = [&](int p, float4 (&f)[8], float (&w)[4], uint32_t& off) {
        const int kb = p / PASSES_PER_KB, r = 4 * (p % PASSES_PER_KB) + (lane >> 3);
        const Tap rt = taps[r];
        const size_t ox = (rt.pix_dxy & (1 << 30)) ? (size_t)HID : 0, oy = (rt.pix_dxy < 0) ? (size_t)s.Wl * HID : 0;
        const float ex = 1.0f - rt.wx, ey = 1.0f - rt.wy;
        w[0] = __fmul_rn(ex, ey); w[1] = __fmul_rn(rt.wx, ey); w[2] = __fmul_rn(ex, rt.wy); w[3] = __fmul_rn(rt.wx, rt.wy);
        const int k0 = KBLK * kb + 8 * (lane & 7);
        const float* b00 = ymap + (size_t)(rt.pix_dxy & 0x3FFFFFFF) * HID + k0;
        f[0] = __ldg((const float4*)b00); f[1] = __ldg((const float4*)(b00 + 4));
        f[2] = __ldg((const float4*)(b00 + ox)); f[3] = __ldg((const float4*)(b00 + ox + 4));
        f[4] = __ldg((const float4*)(b00 + oy)); f[5] = __ldg((const float4*)(b00 + oy + 4));
        f[6] = __ldg((const float4*)(b00 + oy + ox)); f[7] = __ldg((const float4*)(b00 + oy + ox + 4));
        off = act_off(r, k0 >> 3);
    };
    // explicit fma chain: the expression is instantiated twice below and must round identically in both (the result must not
    // depend on which warp / which pass of a warp stages a row)
    auto bl = [](float a, float b, float c, float d, const float (&w)[4]) { return fmaf(d, w[3], fmaf(c, w[2], fmaf(b, w[1], __fmul_rn(a, w[0])))); };
    auto finish = [&](const float4 (&f)[8], const float (&w)[4], uint32_t off) {
        float4 c03, c47;
        c03.x = bl(f[0].x, f[2].x, f[4].x, f[6].x, w); c03.y = bl(f[0].y, f[2].y, f[4].y, f[6].y, w);
        c03.z = bl(f[0].z, f[2].z, f[4].z, f[6].z, w); c03.w = bl(f[0].w, f[2].w, f[4].w, f[6].w, w);
        c47.x = bl(f[1].x, f[3].x, f[5].x, f[7].x, w); c47.y = bl(f[1].y, f[3].y, f[5].y, f[7].y, w);
        c47.z = bl(f[1].z, f[3].z, f[5].z, f[7].z, w); c47.w = bl(f[1].w, f[3].w, f[5].w, f[7].w, w);
        *(float4*)(Ahi + off) = c03;                // channels k0..k0+3
        *(float4*)(Alo + off) = c47;                // channels k0+4..k0+7
    };
    const bool helper = wwarp >= NUM_WORKER_WARPS;
    const int my_lo = helper ? (kb_lo > WORKER_KB_HI + 1 ? kb_lo : WORKER_KB_HI + 1) : kb_lo;
    const int my_hi = helper ? kb_hi : (kb_hi < WORKER_KB_HI ? kb_hi : WORKER_KB_HI);
    const int p_step = helper ? NUM_HELPER_WARPS : NUM_WORKER_WARPS;
    const int p_begin = my_lo * PASSES_PER_KB + (helper ? wwarp - NUM_WORKER_WARPS : wwarp);
    const int p_end = (my_hi + 1) * PASSES_PER_KB;   // <= p_begin when this warp class has no K block in the range
    int waited = kb_lo - 1;                         // highest K block whose "free" barrier this warp has passed
#pragma unroll 1
    for (int p = p_begin; p < p_end; p += 2 * p_step) {
        float4 fa[8], fb[8];
        float wa[4], wb[4];
        uint32_t oa, ob = 0;
        const int p2 = p + p_step;
        const bool two = p2 < p_end;
        const int kb_need = (two ? p2 : p) / PASSES_PER_KB;
        // the global loads do not touch the operand buffers: issue them BEFORE waiting for the K block's release, so that
        // their latency overlaps the wait and only the staging stores are gated by the barrier
        issue(p, fa, wa, oa);
        if (two) issue(p2, fb, wb, ob);
        while (waited < kb_need) { ++waited; mbar_wait(bar_afree + 8 * waited, waited == 0 ? par0 : par1, a.err, 45); }
        finish(fa, wa, oa);
        if (two) finish(fb, wb, ob);
    }
    // Helpers pass every barrier of the range (they own its last K blocks anyway).  Workers do NOT wait for the K blocks beyond their
    // own share: those are released in the N-tile-outer tail of the running GEMM, i.e. at its very end, and the point of the tail
    // is that the workers' first epilogue half runs under it.  Skipping a phase of bar_afree[kb] is safe for them: before their
    // next wait on that barrier they pass bar_acc of the step that completes the skipped phase (commits complete in order).
    if (helper)
        while (waited < kb_hi) { ++waited; mbar_wait(bar_afree + 8 * waited, waited == 0 ? par0 : par1, a.err, 46); }
}

// Block entry epilogue (replaces the lin_z GEMM): x' = x (TMEM, fp32) + g (staged bilinear Y_b row; both carry W_SCALE), written
// back to TMEM as the residual; relu(x') * W_INV -> fp16 hi/lo chunks of the fc_0 operand.  The biases that precede this point (b_in / b_fc1[b-1] and
// b_z[b]) are folded into Y_b when the maps are built (the four bilinear weights sum to 1), so the residual in TMEM carries them.  Each thread reads and then overwrites
// only its own (row, k-chunk) slots, so the staging can live in the operand buffers.
template <bool PARITY>
__device__ __forceinline__ void add32_convert(uint32_t* v, int h0, int r, uint8_t* Ahi, uint8_t* Alo) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = act_off(r, (h0 >> 3) + c8);
        const float4 g0 = *(const float4*)(Ahi + off), g1 = *(const float4*)(Alo + off);
        float xs[8], x[8];
        xs[0] = __uint_as_float(v[8 * c8 + 0]) + g0.x; xs[1] = __uint_as_float(v[8 * c8 + 1]) + g0.y;
        xs[2] = __uint_as_float(v[8 * c8 + 2]) + g0.z; xs[3] = __uint_as_float(v[8 * c8 + 3]) + g0.w;
        xs[4] = __uint_as_float(v[8 * c8 + 4]) + g1.x; xs[5] = __uint_as_float(v[8 * c8 + 5]) + g1.y;
        xs[6] = __uint_as_float(v[8 * c8 + 6]) + g1.z; xs[7] = __uint_as_float(v[8 * c8 + 7]) + g1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * c8 + i] = __float_as_uint(xs[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaxf(xs[i], 0.0f) * W_INV;
        uint4 hi, lo;
        split8(x, hi, lo);
        *(uint4*)(Ahi + off) = hi;
        if (PARITY) *(uint4*)(Alo + off) = lo;
    }
}
template <bool PARITY>
__device__ __forceinline__ void epilogue_add_y_half(uint32_t tmem, int colx, uint8_t* Ahi, uint8_t* Alo, int q, int lane, int j, int h) {
    const int r = 32 * (q & 1) + lane;
    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(colx + 128 * h + 64 * j);
    const int hb = 256 * h + 128 * (q >> 1) + 64 * j;
    uint32_t va[32], vb[32];
    tmem_ld32_issue(t0, va);
    tmem_ld32_issue(t0 + 32, vb);
    tmem_ld_wait();
    add32_convert<PARITY>(va, hb, r, Ahi, Alo);
    tmem_st32_issue(t0, va);
    add32_convert<PARITY>(vb, hb + 32, r, Ahi, Alo);
    tmem_st32_issue(t0 + 32, vb);
    tmem_st_wait();
}

// ZMAP: 64 latent pixels (rows) x L channels -> fp16 hi/lo A operand; consecutive threads take consecutive 8-channel chunks
__device__ __forceinline__ void load_latent_rows(const Args& a, long long tile, int wt, uint8_t* Ahi, uint8_t* Alo) {
    const int cpr = a.s.L >> 3;                     // 8-channel chunks per row
#pragma unroll 2
    for (int i = wt; i < ROWS * cpr; i += NUM_OPND_WARPS * 32) {
        const int r = i / cpr, kc = i % cpr;
        long long pix = tile * ROWS + r;
        if (pix >= a.n_pix) pix = a.n_pix - 1;
        const float4* src = (const float4*)(a.s.latent + (size_t)pix * a.s.L + 8 * kc);
        const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
        const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = act_off(r, kc);
        *(uint4*)(Ahi + off) = hi;
        *(uint4*)(Alo + off) = lo;
    }
}
// ZMAP: accumulator rows -> Y_b[pixel][512] fp32, stored WITH the W_SCALE factor of the accumulator (thread = row, 128 contiguous bytes per TMEM load)
__device__ __forceinline__ void store_y_rows(uint32_t tmem, float* __restrict__ ymap, const float* __restrict__ bias, long long pix, bool write,
                                             int q, int lane, int n2) {
    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(COL_X + 128 * n2);
    const int hb = 256 * n2 + 128 * (q >> 1);
#pragma unroll 1
    for (int c32 = 0; c32 < 4; ++c32) {
        uint32_t v[32];
        tmem_ld32(t0 + 32 * c32, v);
        if (write) {
            float4* dst = (float4*)(ymap + (size_t)pix * HID + hb + 32 * c32);
            const float4* bs = (const float4*)(bias + hb + 32 * c32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 bv = __ldg(bs + i);
                dst[i] = make_float4(fmaf(bv.x, W_SCALE, __uint_as_float(v[4 * i])), fmaf(bv.y, W_SCALE, __uint_as_float(v[4 * i + 1])),
                                     fmaf(bv.z, W_SCALE, __uint_as_float(v[4 * i + 2])), fmaf(bv.w, W_SCALE, __uint_as_float(v[4 * i + 3])));
            }
        }
    }
}

// L2 residency of the FUSED launch's x_c slab (148 x 128 KiB): the Y-map gather streams gigabytes through the L2 and would evict the
// slab's dirty lines to HBM (ncu: 0.47 GB of write-backs per 524 288 samples without the hint); evict_last keeps them.
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_f4(float* p, float4 v, uint64_t pol) {
    if (pol) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    else *(float4*)p = v;
}
__device__ __forceinline__ float4 ld_f4_l2(const float4* p, uint64_t pol) {     // data written by this kernel: L2 only, never the read-only path
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
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
__device__ __forceinline__ void combine_store(uint32_t* raw, const float* __restrict__ cb, int h0, int lane, float* dst_sample, bool write,
                                              int nv_real, uint64_t pol) {
    float v[32];
    const bool pad_row = (lane & (NV - 1)) >= nv_real;          // row of a padding view: contributes nothing to the mean
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = pad_row ? 0.0f : __uint_as_float(raw[i]);
    const int off = combine_lanes<NV>(v, lane);
    constexpr int CNT = 32 / NV;
    if (write) {
        const float scale = W_INV / (float)nv_real;
#pragma unroll
        for (int i = 0; i < CNT; ++i) v[i] = v[i] * scale + __ldg(cb + h0 + off + i);
        float* dst = dst_sample + h0 + off;
        if constexpr (CNT >= 4) {
#pragma unroll
            for (int i = 0; i < CNT; i += 4) st_f4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]), pol);
        } else {
#pragma unroll
            for (int i = 0; i < CNT; ++i) dst[i] = v[i];
        }
    }
}

// Operand hand-off to the MMA issuer (leader CTA), one barrier per operand HALF (K blocks 0..3 / 4..7); called by the 8 WORKER
// warps only (helper warps hand their prep / gather results to the workers through named barriers).  Remote mbarrier arrives
// are slow (~1 us each and they serialise), so the peer CTA joins its worker warps on a named barrier (non-blocking bar.arrive
// for 7 of them) and warp 0 sends ONE remote arrive; the leader's own warps arrive locally.  Leader barrier count =
// NUM_WORKER_WARPS + 1.  Two uses of the same half are always separated by a wait on bar_acc.
template <int HALF>
__device__ __forceinline__ void worker_arrive(uint32_t bar_opnd, uint32_t leader_opnd, bool is_leader_cta, int wwarp, int lane) {
    fence_proxy_async();
    tc_fence_before();
    if (is_leader_cta) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_opnd + 8 * HALF);
    } else if (wwarp == 0) {
        asm volatile("bar.sync %0, %1;" ::"n"(2 + 2 * HALF), "n"(NUM_WORKERS) : "memory");
        if (lane == 0) mbar_arrive_remote(leader_opnd + 8 * HALF);
    } else {
        asm volatile("bar.arrive %0, %1;" ::"n"(2 + 2 * HALF), "n"(NUM_WORKERS) : "memory");
    }
}
// all 12 operand warps (prep / operand loads that the helpers take part in): generic-proxy writes -> async proxy, then join
__device__ __forceinline__ void opnd_warps_join() {
    fence_proxy_async();
    asm volatile("bar.sync 6, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
}

#define TS_ROUND 21     // a steady-state tile (the first rounds of a launch gather cold Y-map lines from HBM); fused launch, 4 views: a warm PRE tile
#define TS(role, slot) do { if (a.dbg_ts && blockIdx.x < 2 && rd == TS_ROUND && lane == 0) a.dbg_ts[(blockIdx.x * 4 + (role)) * 64 + (slot)] = clock64(); } while (0)
#define TSH(slot) do { if (a.dbg_ts && blockIdx.x < 2 && rd == TS_ROUND && wwarp == NUM_WORKER_WARPS && lane == 0) a.dbg_ts[(blockIdx.x * 4 + 2) * 64 + (slot)] = clock64(); } while (0)
#define TSW() do { if (a.dbg_ts && blockIdx.x < 2 && rd == TS_ROUND && wwarp == 0 && lane == 0 && tsn < 64) a.dbg_ts[(blockIdx.x * 4 + 1) * 64 + tsn++] = clock64(); } while (0)
// ---- the kernel ----------------------------------------------------------------------------------
constexpr int KIND_PRE = 0, KIND_POST = 1, KIND_ZMAP = 2, KIND_FUSED = 3;
template <bool PARITY, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) mlp_pair_kernel(const __grid_constant__ Args a) {
    using C = Cfg<PARITY>;
    constexpr bool FUSED = KIND == KIND_FUSED;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool is_leader_cta = crank == 0;
    const uint32_t smem_base = smem_u32(smem);
    uint8_t* Ahi = smem + C::OFF_A_HI;
    uint8_t* Alo = smem + C::OFF_A_LO;
    Tap* taps = (Tap*)(smem + C::OFF_TAPS);                        // [2][ROWS]
    const uint32_t bar_full = smem_base + C::OFF_BARS;             // NST: weight stage landed in THIS CTA
    const uint32_t bar_empty = bar_full + 8 * C::NST;              // NST: stage free (pair commit)
    const uint32_t bar_opnd = bar_empty + 8 * C::NST;              // 2: (leader) A operand half h of both CTAs ready
    const uint32_t bar_acc = bar_opnd + 16;                        // accumulators of a GEMM step complete (pair commit)
    const uint32_t bar_acc0 = bar_acc + 8;                         // N tile 0 of the step's accumulator complete (pair commit; before bar_acc)
    const uint32_t bar_afree = bar_acc0 + 8;                       // 8: K block kb of the A operand no longer read (pair commit)
    const uint32_t bar_lin0 = bar_afree + 8 * (HID / KBLK);          // lin_in: N tile 0 / the whole step complete.  Its own pair of barriers:
    const uint32_t bar_lin = bar_lin0 + 8;                          // lin_in of the NEXT tile is issued right behind the last fc_1 (early_lin),
                                                                   // and two phases of ONE barrier must never complete within a waiter's reach
    const uint32_t bar_feat = bar_lin + 8;                          // (leader) lin_in features of the next tile in place in both CTAs (helper warps)
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + C::OFF_BARS + 8 * (2 * C::NST + 4 + HID / KBLK + 3));

    if ((smem_base & 1023u) != 0) { if (threadIdx.x == 0) atomicExch(a.err, 90); __trap(); }
    const long long clk_start = a.dbg_ts ? clock64() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_opnd, NUM_WORKER_WARPS + 1);
        mbar_init(bar_opnd + 8, NUM_WORKER_WARPS + 1);
        mbar_init(bar_acc, 1);
        mbar_init(bar_acc0, 1);
        mbar_init(bar_lin0, 1);
        mbar_init(bar_lin, 1);
        mbar_init(bar_feat, NUM_HELPER_WARPS + 1);
        for (int i = 0; i < HID / KBLK; ++i) mbar_init(bar_afree + 8 * i, 1);
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
    const uint32_t leader_opnd = map_to_cta(bar_opnd, 0);      // + 8 * half
    const uint32_t leader_feat = map_to_cta(bar_feat, 0);

    // both CTAs of a pair run the same number of rounds; CTA tile = 2 * pair_tile + rank
    const long long first = (long long)blockIdx.x, stride = (long long)gridDim.x;
    const long long n_rounds = (a.n_tiles + stride - 1) / stride;
    const int uses_per_round = FUSED ? a.ppr * a.uses_per_tile + a.pts * a.uses_post : a.uses_per_tile;
    const long long total_uses = n_rounds * uses_per_round;

    const int prod_idx = warp == 0 ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : -1));
    if (prod_idx >= 0) {
        // ===== weight producers (this CTA's half of every weight tile); stage st is always filled by producer st % NUM_PRODUCERS.
        //       2-SM TMA: both CTAs' copies complete_tx on the LEADER's full barrier, so the MMA issuer needs no software relay.
        const bool leader = elect_one();
        const int* table = a.tile_table + (size_t)crank * a.uses_per_tile;
        const int* table_post = FUSED ? a.tile_table_post + (size_t)crank * a.uses_post : nullptr;
        const uint32_t leader_full = map_to_cta(bar_full, 0);
        for (long long base = 0; base < total_uses; base += C::NST) {
            for (int st = prod_idx; st < C::NST; st += NUM_PRODUCERS) {
                const long long use = base + st;
                if (use >= total_uses) break;
                const int t = (int)(use % uses_per_round);
                const uint32_t ph = (uint32_t)((use / C::NST) & 1);
                mbar_wait(bar_empty + 8 * st, ph ^ 1, a.err, 10);
                if (leader) {
                    const int tix = (!FUSED || t < a.ppr * a.uses_per_tile) ? __ldg(table + t % a.uses_per_tile)
                                                                            : __ldg(table_post + (t - a.ppr * a.uses_per_tile) % a.uses_post);
                    const int row = tix * 128;
                    if (is_leader_cta) mbar_arrive_expect_tx(bar_full + 8 * st, 2 * WTILE_BYTES);
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_base + st * WTILE_BYTES), "l"(&a.wmap), "r"(0), "r"(row), "r"(leader_full + 8 * st) : "memory");
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 && is_leader_cta) {
        // ===== leader CTA: MMA issuer for the pair (converged warp, one elected lane).  K-block-outer order: the A operand is
        //       consumed half by half (K blocks 0..3, then 4..7), each half behind its own operand barrier, and a K block is
        //       released (bar_afree) right after its last MMA when the step is followed by a gather into the operand buffers.
        const bool leader = elect_one();
        uint32_t use = 0, oph[2] = {0, 0};
        // the GEMM steps of one tile; cold = no previous PRE tile prepared this one (the whole Y_0 gather follows lin_in),
        // has_next = another PRE tile follows in the pipeline (its early gather consumes the last fc_1's operand releases)
        uint32_t fph = 0;                        // phases of bar_feat consumed
        // One GEMM step.  par: TMEM half parity of the tile (x / net swap columns from tile to tile, see early_lin); via_feat: the
        // operand is the next tile's lin_in features, signalled by the helper warps on bar_feat (not by the workers on bar_opnd)
        auto issue_step = [&](GemmStep gs, int sidx, bool cold, bool has_next, long long rd, int par, bool via_feat) {
            {
                if (gs.release == 3) gs.release = has_next ? 1 : 0;
                const bool is_lin = gs.release == 2;
                gs.dst_col = (short)(gs.dst_col ^ (par ? 256 : 0));
                const uint32_t idesc = make_idesc2(gs.n_width);
                // K blocks [0, kb_split) K-block-outer over both N tiles; the tail [kb_split, nkb) N-tile-outer: N tile 0 of the
                // accumulator completes `tail` K blocks early (bar_acc0) and its epilogue half overlaps the n1 tail
                const int kb_split = gs.n_tiles == 2 ? gs.nkb - (gs.tail < gs.nkb ? gs.tail : gs.nkb) : gs.nkb;
                auto wait_half = [&](int kb) {
                    if ((kb & 3) == 0) {
                        const int h = kb >> 2;
                        if (via_feat) { mbar_wait(bar_feat, fph & 1, a.err, 22); ++fph; }
                        else { mbar_wait(bar_opnd + 8 * h, oph[h] & 1, a.err, 20 + h); ++oph[h]; }
                        tc_fence_after();
                        TS(0, 4 * sidx + h);
                    }
                };
                auto issue_tile = [&](int kb, int n2) {
                    const uint32_t d = tmem + (uint32_t)(gs.dst_col + 128 * n2);
                    {   // W_hi tile: A_hi*W_hi (+ A_lo*W_hi)
                        const uint32_t st = use % C::NST, ph = (use / C::NST) & 1;
                        mbar_wait(bar_full + 8 * st, ph, a.err, 30);
                        tc_fence_after();
                        if (leader) {
                            const uint64_t bdesc = make_desc(smem_base + st * WTILE_BYTES, 16, 1024);
                            const uint64_t ahi = make_desc(smem_base + C::OFF_A_HI + kb * ACT_KB_BYTES, 16, 1024);
                            const uint64_t alo = make_desc(smem_base + C::OFF_A_LO + kb * ACT_KB_BYTES, 16, 1024);
#pragma unroll
                            for (int j = 0; j < ((a.dbg_skip & 8) ? 0 : 4); ++j) {
                                umma2_f16(d, ahi + 2 * j, bdesc + 2 * j, idesc, (gs.accumulate | kb | j) ? 1u : 0u);
                                if (PARITY) umma2_f16(d, alo + 2 * j, bdesc + 2 * j, idesc, 1u);
                            }
                            umma2_commit_pair(bar_empty + 8 * st);
                        }
                        __syncwarp();
                        ++use;
                    }
                    if (PARITY) {   // W_lo tile: A_hi*W_lo
                        const uint32_t st = use % C::NST, ph = (use / C::NST) & 1;
                        mbar_wait(bar_full + 8 * st, ph, a.err, 31);
                        tc_fence_after();
                        if (leader) {
                            const uint64_t bdesc = make_desc(smem_base + st * WTILE_BYTES, 16, 1024);
                            const uint64_t ahi = make_desc(smem_base + C::OFF_A_HI + kb * ACT_KB_BYTES, 16, 1024);
#pragma unroll
                            for (int j = 0; j < ((a.dbg_skip & 8) ? 0 : 4); ++j) umma2_f16(d, ahi + 2 * j, bdesc + 2 * j, idesc, 1u);
                            umma2_commit_pair(bar_empty + 8 * st);
                        }
                        __syncwarp();
                        ++use;
                    }
                };
                for (int kb = 0; kb < kb_split; ++kb) {
                    wait_half(kb);
                    for (int n2 = 0; n2 < gs.n_tiles; ++n2) issue_tile(kb, n2);
                    if (gs.release && leader) umma2_commit_pair(bar_afree + 8 * kb);
                    __syncwarp();
                }
                for (int n2 = 0; n2 < gs.n_tiles; ++n2) {
                    for (int kb = kb_split; kb < gs.nkb; ++kb) {
                        if (n2 == 0) wait_half(kb);
                        issue_tile(kb, n2);
                        if (n2 + 1 == gs.n_tiles && gs.release && leader) umma2_commit_pair(bar_afree + 8 * kb);
                        __syncwarp();
                    }
                    if (n2 == 0 && leader) umma2_commit_pair(is_lin ? bar_lin0 : bar_acc0);      // N tile 0 (or the whole of a single-tile step)
                    __syncwarp();
                }
                if (leader) {
                    // lin_in (release == 2) reads K block 0 only; in a cold tile the gather of the whole Y_0 row set follows it, so
                    // the other K blocks are released here as well (warm tiles gather them during the previous tile's last fc_1)
                    if (is_lin && cold) for (int kb = gs.nkb; kb < HID / KBLK; ++kb) umma2_commit_pair(bar_afree + 8 * kb);
                    umma2_commit_pair(is_lin ? bar_lin : bar_acc);
                }
                __syncwarp();
                TS(0, 4 * sidx + 2);
            }
        };
        // The steps of one tile.  early_lin: lin_in of a warm tile was issued at the end of the previous tile (right behind its last
        // fc_1, into the TMEM half that tile used for `net`), so the tile starts with its blocks and ends with the next tile's lin_in.
        auto run_steps = [&](const GemmStep* steps, int n_steps, bool cold, bool has_next, long long rd, int par) {
            const bool early = a.early_lin && steps[0].release == 2;
            for (int sidx = (early && !cold) ? 1 : 0; sidx < n_steps; ++sidx) issue_step(steps[sidx], sidx, cold, has_next, rd, par, false);
            if (early && has_next) issue_step(steps[0], 0, false, false, rd + 1, par ^ 1, true);
        };
        for (long long rd = 0; rd < n_rounds; ++rd) {
            if constexpr (FUSED) {
                for (int j = 0; j < a.ppr; ++j)
                    run_steps(a.steps, a.n_steps, j == 0 && (rd == 0 || !a.warm_rounds), j + 1 < a.ppr, rd * a.ppr + j,
                              a.warm_rounds ? ((j + 1) & 1) : a.early_lin ? (j & 1) : 0);
                for (int k = 0; k < a.pts; ++k) {
                    // warm_rounds: the last fc_1 of the round's last POST tile releases its K blocks to the next round's first PRE
                    // tile, whose lin_in (a.steps[0]) follows it into TMEM half 1
                    const bool warm_next = a.warm_rounds && k + 1 == a.pts && rd + 1 < n_rounds;
                    run_steps(a.steps_post, a.n_steps_post, false, warm_next, -1, 0);
                    if (warm_next) issue_step(a.steps[0], 0, false, false, (rd + 1) * a.ppr, 1, true);
                }
            } else {
                run_steps(a.steps, a.n_steps, rd == 0, rd + 1 < n_rounds, rd, (a.early_lin && KIND == KIND_PRE) ? (int)(rd & 1) : 0);
            }
        }
    } else if (warp >= WORKER_WARP0) {
        // ===== workers (warps 4..11) + helpers (warps 12..15; prep and gather only).
        //       worker TMEM lanes 32q..32q+31: row = 32*(q&1)+lane, hidden half (q>>1) of this warp's N tile n2
        const int wwarp = warp - WORKER_WARP0, wt = threadIdx.x - WORKER_WARP0 * 32;
        const bool helper = wwarp >= NUM_WORKER_WARPS;
        const int q = warp & 3, n2 = (wwarp >> 2) & 1;
        const int r = 32 * (q & 1) + lane;
        uint32_t it = 0, ph0 = 0, ph1 = 0, itl = 0;   // phase counters: bar_acc, bar_afree[0], bar_afree[1..7], bar_lin
        const uint64_t slab_pol = FUSED ? l2_evict_last_policy() : 0;
        (void)ph0; (void)ph1;
        // helpers: next tile's lin_in features in place -> straight to the issuer (leader's bar_feat; the peer's helpers join first
        // and send ONE remote arrive, like worker_arrive)
        auto feat_arrive = [&]() {
            const int hw = wwarp - NUM_WORKER_WARPS;
            if (is_leader_cta) {
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bar_feat);
            } else if (hw == 0) {
                asm volatile("bar.sync 11, %0;" ::"n"(NUM_HELPER_WARPS * 32) : "memory");
                if (lane == 0) mbar_arrive_remote(leader_feat);
            } else {
                asm volatile("bar.arrive 11, %0;" ::"n"(NUM_HELPER_WARPS * 32) : "memory");
            }
        };
        // One PRE tile (64 sample-view rows).  cold: nothing was prepared by a previous tile; has_next: tile_next follows in the
        // pipeline (its taps / features / Y_0 are produced under this tile's last block); pt: running PRE tile count (tap buffer
        // parity); xc_row0 (FUSED): first row of this tile's samples in the CTA's x_c slab.
        auto pre_tile = [&](long long tile, bool live, bool cold, bool has_next, long long tile_next, long long pt, long long xc_row0, int par) {
            // Tile pipeline (steady state, !cold): the taps and the lin_in features of this tile, and the Y_0 staging of
            // K blocks 1..7, were produced during the previous tile's last block; its lin_in was handed off after that tile's combine.
            int tsn = 0; (void)tsn;
            const long long rd = pt; (void)rd;
            if (a.dbg_ts && blockIdx.x < 4 && wwarp == 0 && lane == 0) a.dbg_ts[512 + blockIdx.x * 512 + (pt < 511 ? pt : 511)] = clock64();
            Tap* tp = taps + (pt & 1) * ROWS;
            Tap* tn = taps + ((pt + 1) & 1) * ROWS;
            // TMEM halves of this tile: with early_lin the residual x and the fc_0 output swap columns from tile to tile (the next
            // tile's lin_in is written into this tile's `net` half while this tile's x is still being combined)
            const int colX = par ? COL_NET : COL_X, colNET = par ? COL_X : COL_NET;
            if (cold) {
                prep_rows<PARITY, NUM_OPND_WARPS * 32 / 64, true, true>(a, tile, wt, Ahi, Alo, tp);
                opnd_warps_join();
                if (!helper) worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                   // -> lin_in (K block 0 only)
            }
            TSW();
            for (int b = 0; b < a.n_blocks; ++b) {
                const bool last = b + 1 == a.n_blocks;
                // x += lin_z[b](latent)  ==  x += bilinear(Y_b): gathered into the operand buffers K block by K block while the
                // previous GEMM (lin_in / fc_1[b-1]) is still running, then added to the residual in the epilogue
                if (b == 0 && !cold) {       // only K block 0 is left (it held the lin_in features until now)
                    // (early_lin: handing this K block to the otherwise idle helper warps -- 4 warps instead of 8, plus a 384-thread
                    // barrier -- was measured 2 % SLOWER end to end than letting the workers stage it themselves)
                    gather_y(a, a.zmap, wwarp, lane, Ahi, Alo, tp, 0, 0, bar_afree, ph0 & 1, ph1 & 1); ++ph0;
                } else {
                    gather_y(a, a.zmap + (size_t)b * a.zmap_stride, wwarp, lane, Ahi, Alo, tp, 0, HID / KBLK - 1, bar_afree, ph0 & 1, ph1 & 1, a.worker_kb_hi);
                    ++ph0; ++ph1;
                }
                TSW();
                // Helpers never wait on bar_acc in this kernel: they are gated by bar_afree alone.  (A helper that finishes a late
                // gather could reach a bar_acc wait after the barrier has already completed its NEXT phase -- lin_in of the next
                // tile is short -- and a parity wait that is one phase late blocks for good.)
                if (!helper) {                                                                                     // x, N tile 0 (hidden 0..255) complete
                    if (b == 0) mbar_wait(bar_lin0, itl & 1, a.err, 40); else mbar_wait(bar_acc0, it & 1, a.err, 40);
                }
                TSW();
                tc_fence_after();
                if (helper) {
                    asm volatile("bar.arrive 3, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");                        // staging of K blocks 5..7 complete
                    if (last && has_next) {  // taps of the next tile, while fc_0 of the last block runs
                        if (wt - NUM_WORKERS < ROWS) prep_rows<PARITY, 1, true, false>(a, tile_next, wt - NUM_WORKERS, Ahi, Alo, tn);
                        asm volatile("bar.sync 9, %0;" ::"n"(NUM_HELPER_WARPS * 32) : "memory");                    // all helpers read these taps below
                        asm volatile("bar.arrive 7, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
                    }
                } else {
                    asm volatile("bar.sync 5, %0;" ::"n"(NUM_WORKERS) : "memory");                                  // staging of K blocks 0..4 complete
                    if (!(a.dbg_skip & 2)) epilogue_add_y_half<PARITY>(tmem, colX, Ahi, Alo, q, lane, n2, 0);
                    TSW(); worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                     // -> fc_0[b], K blocks 0..3
                    asm volatile("bar.sync 3, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
                    if (b == 0) { mbar_wait(bar_lin, itl & 1, a.err, 41); ++itl; }                                  // x complete
                    else { mbar_wait(bar_acc, it & 1, a.err, 41); ++it; }
                    tc_fence_after();
                    if (!(a.dbg_skip & 2)) epilogue_add_y_half<PARITY>(tmem, colX, Ahi, Alo, q, lane, n2, 1);
                    TSW(); worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                     // -> fc_0[b], K blocks 4..7
                }
                if (!helper) {
                    mbar_wait(bar_acc0, it & 1, a.err, 42); TSW();                                                  // net, N tile 0
                    tc_fence_after();
                    const float* b0 = a.bias + (size_t)(a.n_blocks + b) * HID;
                    if (!(a.dbg_skip & 2)) epilogue_half<PARITY>(tmem, colNET, b0, Ahi, Alo, q, lane, n2, 0);
                    TSW(); worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                     // -> fc_1[b], K blocks 0..3
                    mbar_wait(bar_acc, it & 1, a.err, 39); ++it;                                                    // net complete
                    tc_fence_after();
                    if (!(a.dbg_skip & 2)) epilogue_half<PARITY>(tmem, colNET, b0, Ahi, Alo, q, lane, n2, 1);
                    TSW(); worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                     // -> fc_1[b], K blocks 4..7
                }
                if (last && has_next) {
                    // next tile, under the last fc_1: lin_in features into K block 0 as soon as fc_1 has consumed it (helpers),
                    // Y_0 staging of K blocks 1..7 as they are released (everyone)
                    if (helper) {
                        // (computing the features BEFORE this wait and only storing them after it was measured: slower -- the sin / cos
                        // work then runs under the last block's epilogues and takes issue slots from the worker warps that the MMA waits for)
                        TSH(0);
                        mbar_wait(bar_afree, ph0 & 1, a.err, 47);
                        TSH(1);
                        prep_rows<PARITY, NUM_HELPER_WARPS * 32 / 64, false, true>(a, tile_next, wt - NUM_WORKERS, Ahi, Alo, tn);
                        TSH(2);
                        fence_proxy_async();
                        if (a.early_lin) {
                            feat_arrive();
                        } else {
                            asm volatile("bar.arrive 8, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");                // features in place
                        }
                    } else {
                        asm volatile("bar.sync 7, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");                      // next taps in place
                    }
                    gather_y(a, a.zmap, wwarp, lane, Ahi, Alo, tn, 1, HID / KBLK - 1, bar_afree, ph0 & 1, ph1 & 1, a.early_worker_kb_hi);
                    if (helper) TSH(3); else TSW();
                    // K block 0's phase (released by this last fc_1, consumed by the helpers above) is only COUNTED here, never waited
                    // for: by the time a warp gets here, lin_in of the next tile may already have completed the barrier's NEXT phase
                    // (with early_lin it does not depend on the workers at all), and a parity wait that is one phase late blocks for
                    // good.  Skipping is safe: every warp's next wait on bar_afree[0] is for lin_in's release, and the phase after
                    // that needs all of them (fc_1[0] of the next tile).
                    ++ph0; ++ph1;
                }
            }
            // last fc_1: the view combine reads x.  Every worker warp takes two 32-column chunks of N tile 0 as soon as that tile is
            // complete (bar_acc0, under the n1 tail of the GEMM) and two chunks of N tile 1 after the whole step: both warps of a TMEM
            // lane quarter can read any column, so the part of the combine that follows the GEMM is half of what it would be if
            // each warp kept to its own N tile
            if (!helper) mbar_wait(bar_acc0, it & 1, a.err, 43);
            TSW();
            tc_fence_after();
            // combine: mean over the NV adjacent rows (lanes) of each sample, sequential like torch.mean (resnetfc.py:148-151)
            const float* cb = a.bias2 + (size_t)a.n_blocks * HID;                // b_fc1 of the last block (everything earlier is in x already)
            // destination row of this row's sample: its index in the sub-batch (PRE launch) or in this CTA's slab (FUSED)
            const long long smp = tile * a.spv + r / a.NV;
            const bool wr = FUSED || (live && smp < a.n_samples);
            float* dst_sample = a.xc + (size_t)(FUSED ? xc_row0 + r / a.NV : (wr ? smp : 0)) * HID;
#pragma unroll 1
            for (int ci = 0; ci < ((helper || (a.dbg_skip & 128)) ? 0 : 4); ++ci) {
                const int nt = ci >> 1, c32 = 2 * n2 + (ci & 1);                 // N tile of the chunk, chunk within the tile
                if (ci == 2) {
                    mbar_wait(bar_acc, it & 1, a.err, 38);                        // N tile 1 = the whole step
                    tc_fence_after();
                }
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(colX + 128 * nt + 32 * c32), v);
                const int h0 = 256 * nt + 128 * (q >> 1) + 32 * c32;
                switch (a.NV) {
                    case 1: combine_store<1>(v, cb, h0, lane, dst_sample, wr, a.NV_real, slab_pol); break;
                    case 2: combine_store<2>(v, cb, h0, lane, dst_sample, wr, a.NV_real, slab_pol); break;
                    case 4: combine_store<4>(v, cb, h0, lane, dst_sample, wr, a.NV_real, slab_pol); break;
                    case 8: combine_store<8>(v, cb, h0, lane, dst_sample, wr, a.NV_real, slab_pol); break;
                    case 16: combine_store<16>(v, cb, h0, lane, dst_sample, wr, a.NV_real, slab_pol); break;
                    default: combine_store<32>(v, cb, h0, lane, dst_sample, wr, a.NV_real, slab_pol); break;
                }
            }
            if (!helper) ++it;
            tc_fence_before();
            TSW();
            if (has_next && !helper && !a.early_lin) {
                asm volatile("bar.sync 8, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");                              // next tile's features in place
                TSW();
                worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                                // X read out -> lin_in of the next tile
                TSW();
            }
            if (FUSED && !has_next && !helper) {
                // the POST tile overwrites TMEM X and the whole A operand: the last fc_1 is complete (every worker waited for it in
                // the combine), every worker must be done reading X, and the x_c slab rows visible to the whole CTA
                __threadfence_block();
                asm volatile("bar.sync 10, %0;" ::"n"(NUM_WORKERS) : "memory");
            }
        };
        // One POST tile (64 samples): remaining blocks + lin_out on the view-combined activations.  has_next (FUSED, warm_rounds): the
        // PRE tile tile_next (running PRE tile count pt_next) follows; its taps / features / Y_0 staging are produced under this
        // tile's last fc_1 exactly as under a PRE tile's, and its lin_in is issued behind that fc_1 into this tile's `net` half.
        auto post_tile = [&](long long tile, bool live, const float* bias, int n_blocks, long long xc_row0, bool has_next, long long tile_next,
                             long long pt_next) {
            Tap* tn = taps + (pt_next & 1) * ROWS;
            float4* red = (float4*)(taps + ((pt_next + 1) & 1) * ROWS);       // the other tap buffer is dead during a POST tile: lin_out partial sums
            // load x_c: W_SCALE * x_c -> TMEM X (the residual the fc_1 steps accumulate onto), relu(x_c) -> A operand
            long long smp = FUSED ? xc_row0 + r : tile * ROWS + r;          // row of the CTA's slab (FUSED) / sample of the sub-batch
            if (!FUSED && smp >= a.n_samples) smp = a.n_samples - 1;
            // in operand halves like every epilogue (half h: this warp's 64 columns of N tile h), so that fc_0 starts on K blocks 0..3
            // while K blocks 4..7 are still being loaded
#pragma unroll 1
            for (int c32 = 0; c32 < (helper ? 0 : 4); ++c32) {
                const int tcol = 128 * (c32 >> 1) + 64 * n2 + 32 * (c32 & 1);                                       // column within the 256 of x
                const int h0 = 256 * (tcol >> 7) + 128 * (q >> 1) + (tcol & 127);
                const float4* src = (const float4*)(a.xc + (size_t)smp * HID + h0);
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 f = FUSED ? ld_f4_l2(src + i, slab_pol) : __ldg(src + i);     // FUSED: written by this kernel -> no read-only path
                    v[4 * i] = __float_as_uint(f.x * W_SCALE); v[4 * i + 1] = __float_as_uint(f.y * W_SCALE);
                    v[4 * i + 2] = __float_as_uint(f.z * W_SCALE); v[4 * i + 3] = __float_as_uint(f.w * W_SCALE);
                }
                tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(COL_X + tcol), v);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[8 * c8 + i]), 0.0f) * W_INV;
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    const uint32_t off = act_off(r, (h0 >> 3) + c8);
                    *(uint4*)(Ahi + off) = hi;
                    if (PARITY) *(uint4*)(Alo + off) = lo;
                }
                if (c32 == 1) worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                 // -> fc_0 of the first post block, K blocks 0..3
            }
            if (!helper) worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                       // K blocks 4..7
            for (int b = 0; b < n_blocks; ++b) {
                const bool last = b + 1 == n_blocks;
                if (helper && last && has_next) {    // taps of the next PRE tile (the helpers have nothing else to do in a POST tile)
                    if (wt - NUM_WORKERS < ROWS) prep_rows<PARITY, 1, true, false>(a, tile_next, wt - NUM_WORKERS, Ahi, Alo, tn);
                    asm volatile("bar.sync 9, %0;" ::"n"(NUM_HELPER_WARPS * 32) : "memory");
                    asm volatile("bar.arrive 7, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
                }
                if (!helper) {
                    const float* b0 = bias + (size_t)(n_blocks + 1 + b) * HID;
                    mbar_wait(bar_acc0, it & 1, a.err, 50);                                                         // net, N tile 0
                    tc_fence_after();
                    epilogue_half<PARITY>(tmem, COL_NET, b0, Ahi, Alo, q, lane, n2, 0);
                    worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                            // -> fc_1[b], K blocks 0..3
                    mbar_wait(bar_acc, it & 1, a.err, 53); ++it;
                    tc_fence_after();
                    epilogue_half<PARITY>(tmem, COL_NET, b0, Ahi, Alo, q, lane, n2, 1);
                    worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);
                }
                if (!last) {
                    if (!helper) {
                        const float* b1 = bias + (size_t)(b + 1) * HID;
                        mbar_wait(bar_acc0, it & 1, a.err, 51);                                                     // x, N tile 0
                        tc_fence_after();
                        epilogue_half<PARITY>(tmem, COL_X, b1, Ahi, Alo, q, lane, n2, 0);
                        worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);                        // -> next fc_0
                        mbar_wait(bar_acc, it & 1, a.err, 54); ++it;
                        tc_fence_after();
                        epilogue_half<PARITY>(tmem, COL_X, b1, Ahi, Alo, q, lane, n2, 1);
                        worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);
                    }
                } else if (has_next) {
                    // under the last fc_1: the next PRE tile's lin_in features (helpers, K block 0 once fc_1 has read it) and its
                    // Y_0 staging of K blocks 1..7 as they are released (everyone) -- the same hand-off as between two PRE tiles
                    if (helper) {
                        mbar_wait(bar_afree, ph0 & 1, a.err, 56);
                        prep_rows<PARITY, NUM_HELPER_WARPS * 32 / 64, false, true>(a, tile_next, wt - NUM_WORKERS, Ahi, Alo, tn);
                        fence_proxy_async();
                        feat_arrive();
                    } else {
                        asm volatile("bar.sync 7, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");                      // next taps in place
                    }
                    gather_y(a, a.zmap, wwarp, lane, Ahi, Alo, tn, 1, HID / KBLK - 1, bar_afree, ph0 & 1, ph1 & 1, a.early_worker_kb_hi);
                    ++ph0; ++ph1;                                                                                   // (K block 0: counted, see pre_tile)
                }
            }
            // lin_out on the last fc_1's accumulator, half by half as it completes; x = acc * W_INV + b_fc1[last] like every epilogue
            if (!helper) {
                const float* b1 = bias + (size_t)n_blocks * HID;
                float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                mbar_wait(bar_acc0, it & 1, a.err, 55);
                tc_fence_after();
                lin_out_half(tmem, COL_X, b1, a.w_out, q, lane, n2, 0, acc);
                mbar_wait(bar_acc, it & 1, a.err, 52); ++it;
                tc_fence_after();
                lin_out_half(tmem, COL_X, b1, a.w_out, q, lane, n2, 1, acc);
                // the four warps that hold a row (TMEM lane halves x N-tile column split) add up in a fixed order through shared memory
                const int cls = 2 * (q >> 1) + n2;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    if (cls == c) {
                        float4 s = c ? red[r] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        s.x += acc[0]; s.y += acc[1]; s.z += acc[2]; s.w += acc[3];
                        if (c < 3) red[r] = s;
                        else {
                            const long long s_loc = tile * ROWS + r;
                            if (live && s_loc < a.n_samples) {
                                const float4 bo = __ldg((const float4*)a.b_out);
                                const float x0 = s.x + bo.x, x1 = s.y + bo.y, x2 = s.z + bo.z, x3 = s.w + bo.w;
                                ((float4*)a.out)[map_sample(a, a.s_begin + s_loc)] = make_float4(1.0f / (1.0f + expf(-x0)), 1.0f / (1.0f + expf(-x1)),
                                                                                                 1.0f / (1.0f + expf(-x2)), fmaxf(x3, 0.0f));
                            }
                        }
                    }
                    if (c < 3) asm volatile("bar.sync 12, %0;" ::"n"(NUM_WORKERS) : "memory");
                }
            }
            tc_fence_before();
            // a cold PRE tile (or the next POST tile) rewrites the whole operand region and TMEM X with all twelve warps
            if (!has_next) asm volatile("bar.sync 1, %0;" ::"n"(NUM_OPND_WARPS * 32) : "memory");
        };
        for (long long rd = 0; rd < n_rounds; ++rd) {
            const long long tile_raw = first + rd * stride;
            const bool live = tile_raw < a.n_tiles;
            const long long tile = live ? tile_raw : a.n_tiles - 1;
            if constexpr (KIND == KIND_ZMAP) {
                // Y_b = W_z[b] . latent for the 64 latent pixels of this tile (once per scene x weights)
                const bool two_halves = a.steps[0].nkb > 4;
                load_latent_rows(a, tile, wt, Ahi, Alo);
                opnd_warps_join();
                if (!helper) {                                                                                          // -> lin_z[0]
                    worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);
                    if (two_halves) worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);
                }
                const long long pix = tile * ROWS + r;
                for (int b = 0; b < a.n_blocks; ++b) {
                    mbar_wait(bar_acc0, it & 1, a.err, 49);
                    mbar_wait(bar_acc, it & 1, a.err, 44); ++it;
                    tc_fence_after();
                    if (!helper) {
                        store_y_rows(tmem, a.zmap + (size_t)b * a.zmap_stride, a.bias2 + (size_t)b * HID, pix, live && pix < a.n_pix, q, lane, n2);
                        tc_fence_before();
                        if (b + 1 < a.n_blocks) {                                                                       // X free -> lin_z[b+1]
                            worker_arrive<0>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);
                            if (two_halves) worker_arrive<1>(bar_opnd, leader_opnd, is_leader_cta, wwarp, lane);
                        }
                    }
                }
            } else if constexpr (KIND == KIND_PRE) {
                long long tile_next = first + (rd + 1) * stride;
                if (tile_next >= a.n_tiles) tile_next = a.n_tiles - 1;
                pre_tile(tile, live, rd == 0, rd + 1 < n_rounds, tile_next, rd, 0, a.early_lin ? (int)(rd & 1) : 0);
            } else if constexpr (KIND == KIND_POST) {
                post_tile(tile, live, a.bias, a.n_blocks, 0, false, 0, 0);
            } else {
                // FUSED: a.n_tiles counts rounds of 64 * pts samples; PRE tile j covers samples [64 * pts * tile + j * spv, + spv)
                // warm_rounds: TMEM half parity (j + 1) & 1, so that the last PRE tile (ppr is even) and the POST tile keep x in half 0
                // and the next round's first lin_in goes to half 1 while the POST epilogue still reads x
                const long long slab = (long long)blockIdx.x * ROWS * a.pts;
                const long long next_raw = first + (rd + 1) * stride;
                const long long tile_nr = next_raw < a.n_tiles ? next_raw : a.n_tiles - 1;
                for (int j = 0; j < a.ppr; ++j)
                    pre_tile(tile * a.ppr + j, live, j == 0 && (rd == 0 || !a.warm_rounds), j + 1 < a.ppr, tile * a.ppr + j + 1, rd * a.ppr + j,
                             slab + (long long)j * a.spv, a.warm_rounds ? ((j + 1) & 1) : a.early_lin ? (j & 1) : 0);
                for (int k = 0; k < a.pts; ++k)
                    post_tile(tile * a.pts + k, live, a.bias_post, a.n_blocks_post, slab + (long long)k * ROWS,
                              a.warm_rounds && k + 1 == a.pts && rd + 1 < n_rounds, tile_nr * a.ppr, (rd + 1) * a.ppr);
            }
        }
    }
    if (a.dbg_ts && threadIdx.x == 0 && blockIdx.x < 256) a.dbg_ts[2560 + blockIdx.x] = clock64() - clk_start;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <bool PARITY, int KIND>
cudaError_t launch(const Args& a, int grid, cudaStream_t st) {
    auto kern = mlp_pair_kernel<PARITY, KIND>;
    constexpr int smem = Cfg<PARITY>::SMEM_BYTES;
    // per launch (a few hundred ns): the attribute is per device, and one process may drive several
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    g_launches++;
    kern<<<grid, NUM_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tc2

// Per-rank weight tile tables of the pair kernel.  Ring-use order of CTA rank r = for each GEMM step, for kb, for n2:
// tile (2*n2 + r) of that layer; every entry is a 16 KiB tile index into the packed stream ([hi][lo] per tile pair).
// Layout of t.table2: [zmap r0][zmap r1][pre r0][pre r1][post r0][post r1]; zmap is always fp16x3.
static cudaError_t tc2_build_tables(TcState& t, const MlpDev& m, bool parity, cudaStream_t st) {
    using namespace tc2;
    const int kbz = m.d_latent / KBLK, kbh = HID / KBLK;
    std::vector<int> zm[2], pre[2], post[2];
    int layer_pair0 = 0;
    // ring-use order = the MMA issue order of a step (mlp_pair_kernel): K blocks [0, nkb - tail) K-block-outer, the tail N-tile-outer
    auto layer = [&](std::vector<int>* tab, bool par, int nkb, int n_mt, int n_tiles) {
        const int tail = n_tiles == 2 ? (t.tail_kb < nkb ? t.tail_kb : nkb) : 0, kb_split = nkb - tail;
        if (tab)
            for (int r = 0; r < 2; ++r) {
                auto put = [&](int kb, int n2) {
                    const int pair = layer_pair0 + (2 * n2 + r) * nkb + kb;
                    tab[r].push_back(2 * pair);
                    if (par) tab[r].push_back(2 * pair + 1);
                };
                for (int kb = 0; kb < kb_split; ++kb)
                    for (int n2 = 0; n2 < n_tiles; ++n2) put(kb, n2);
                for (int n2 = 0; n2 < n_tiles; ++n2)
                    for (int kb = kb_split; kb < nkb; ++kb) put(kb, n2);
            }
        layer_pair0 += n_mt * nkb;
    };
    layer(pre, parity, 1, 4, 2);                                     // lin_in
    for (int b = 0; b < t.n_pre; ++b) {
        layer(zm, true, kbz, 4, 2);                                  // lin_z[b]: hoisted into the per-scene Y maps
        layer(pre, parity, kbh, 4, 2);                               // fc_0[b]
        layer(pre, parity, kbh, 4, 2);                               // fc_1[b]
    }
    for (int b = 0; b < t.n_post; ++b) { layer(post, parity, kbh, 4, 2); layer(post, parity, kbh, 4, 2); }
    // (lin_out is not a GEMM step: the POST epilogue computes its 4 outputs on the CUDA cores from the fp32 weights)
    t.uses2_zmap = (int)zm[0].size(); t.uses2_pre = (int)pre[0].size(); t.uses2_post = (int)post[0].size();
    std::vector<int> flat;
    for (int r = 0; r < 2; ++r) flat.insert(flat.end(), zm[r].begin(), zm[r].end());
    for (int r = 0; r < 2; ++r) flat.insert(flat.end(), pre[r].begin(), pre[r].end());
    for (int r = 0; r < 2; ++r) flat.insert(flat.end(), post[r].begin(), post[r].end());
    if (flat.size() > 8192) return cudaErrorInvalidValue;
    if (!t.table2) TCK(cudaMalloc((void**)&t.table2, 8192 * sizeof(int)));
    TCK(cudaMemcpyAsync(t.table2, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    TCK(cudaStreamSynchronize(st));                                  // `flat` is a stack-lifetime host buffer
    t.table2_parity = (int)parity;
    t.table2_tail = t.tail_kb;
    return cudaSuccess;
}

static cudaError_t tc2_grid_cap(TcState& t, int num_sms, int* cap) {
    using namespace tc2;
    if (t.max_grid2 == 0) {
        auto kern = mlp_pair_kernel<true, KIND_PRE>;
        TCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM_BYTES));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = Cfg<true>::SMEM_BYTES;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = num_sms / 2; }
        t.max_grid2 = n > 0 ? 2 * n : 2;
    }
    *cap = t.max_grid2 < (num_sms / 2) * 2 ? t.max_grid2 : (num_sms / 2) * 2;
    return cudaSuccess;
}

// Y_b = W_z[b] . latent for every latent pixel (fp16x3): once per (scene, weights); see the header comment.
static cudaError_t tc2_zmap(TcState& t, const SceneDev& s, const MlpDev& m, int grid_cap, cudaStream_t st) {
    using namespace tc2;
    const long long n_pix = (long long)s.SB * s.NV * s.Hl * s.Wl;
    const size_t need = (size_t)t.n_pre * n_pix * HID * sizeof(float);
    if (need > t.zmap_bytes) {
        if (t.zmap) cudaFree(t.zmap);
        t.zmap = nullptr; t.zmap_bytes = 0;
        TCK(cudaMalloc((void**)&t.zmap, need));
        t.zmap_bytes = need;
    }
    Args z{};
    z.wmap = t.wmap;
    z.s = s;
    z.wstream = (const uint8_t*)t.wpack;
    z.tile_table = t.table2;
    z.uses_per_tile = t.uses2_zmap;
    z.bias = t.bias; z.bias2 = t.bias + t.bias_pair_off;
    z.n_blocks = t.n_pre;
    for (int b = 0; b < t.n_pre; ++b) z.steps[b] = GemmStep{(short)(m.d_latent / KBLK), 2, 256, COL_X, 0, 0, (short)t.tail_kb};
    z.n_steps = t.n_pre;
    z.zmap = t.zmap; z.zmap_stride = n_pix * HID; z.n_pix = n_pix;
    z.n_tiles = (n_pix + ROWS - 1) / ROWS;
    z.NV = z.NV_real = 1; z.spv = ROWS;
    z.err = t.err_flag;
    const long long g = ((z.n_tiles + 1) / 2) * 2;
    if (t.timing) TCK(cudaEventRecord(t.ev[0], st));
    TCK((launch<true, KIND_ZMAP>(z, (int)(g < grid_cap ? g : grid_cap), st)));
    if (t.timing) {
        TCK(cudaEventRecord(t.ev[1], st));
        TCK(cudaEventSynchronize(t.ev[1]));
        TCK(cudaEventElapsedTime(&t.ms_zmap, t.ev[0], t.ev[1]));
    }
    t.zmap_valid = true;
    return cudaSuccess;
}

static cudaError_t tc2_dump_stamps(const long long* dev, cudaStream_t st);

cudaError_t tc2_query(TcState& t, const SceneDev& s, const MlpDev& m, const QueryArgs& q, bool parity, int num_sms,
                      cudaStream_t st) {
    using namespace tc2;
    if (s.NV > 32) { snprintf(t.why, sizeof(t.why), "NV=%d views (the tcgen05 path serves up to 32)", s.NV); return cudaErrorNotSupported; }
    int NV = 1;                       // rows per sample: the view count padded to a power of two (padding rows are masked out of the mean)
    while (NV < s.NV) NV *= 2;
    if (s.L != m.d_latent || (s.L % KBLK) || s.L > HID) { snprintf(t.why, sizeof(t.why), "pair kernel needs d_latent == latent channels, %% 64 == 0, <= 512 (got %d / %d)", m.d_latent, s.L); return cudaErrorNotSupported; }
    const int d_in = 3 + 6 * s.num_freqs + 3 + 1 + 2 * s.num_freqs;
    if (d_in != m.d_in) { snprintf(t.why, sizeof(t.why), "positional code gives d_in=%d but lin_in expects %d", d_in, m.d_in); return cudaErrorNotSupported; }
    if (!t.wmap_ok) { snprintf(t.why, sizeof(t.why), "cuTensorMapEncodeTiled unavailable"); return cudaErrorNotSupported; }
    const long long total = (long long)q.SB * q.n_per_sb;
    const int kbh = HID / KBLK;

    if (!t.table2 || t.table2_parity != (int)parity || t.table2_tail != t.tail_kb) {
        TCK(tc2_build_tables(t, m, parity, st));
        t.zmap_valid = false;            // (the ZMAP launch reads the same tables; nothing else depends on them)
    }
    int grid_cap = 2;
    TCK(tc2_grid_cap(t, num_sms, &grid_cap));
    if (t.timing && !t.ev[0]) for (int i = 0; i < 4; ++i) TCK(cudaEventCreate(&t.ev[i]));
    if (!t.zmap_valid) TCK(tc2_zmap(t, s, m, grid_cap, st));
    const long long sub = t.sub_batch > 0 ? t.sub_batch : 524288;
    const size_t need = t.fused ? 0 : (size_t)(((sub + ROWS - 1) / ROWS) * ROWS) * HID * sizeof(float);
    if (need > t.scratch_bytes) {
        if (t.scratch) cudaFree(t.scratch);
        t.scratch = nullptr; t.scratch_bytes = 0;
        TCK(cudaMalloc(&t.scratch, need));
        t.scratch_bytes = need;
    }
    Args pre{}, post{};
    pre.wmap = post.wmap = t.wmap;
    pre.s = s; pre.q = q; post.s = s; post.q = q;
    pre.wstream = post.wstream = (const uint8_t*)t.wpack;
    pre.tile_table = t.table2 + 2 * t.uses2_zmap; post.tile_table = pre.tile_table + 2 * t.uses2_pre;
    pre.uses_per_tile = t.uses2_pre; post.uses_per_tile = t.uses2_post;
    pre.bias = t.bias; post.bias = t.bias + t.bias_post_off;
    pre.bias2 = post.bias2 = t.bias + t.bias_pair_off;
    pre.n_blocks = t.n_pre; post.n_blocks = t.n_post;
    pre.zmap = t.zmap; pre.zmap_stride = (long long)s.SB * s.NV * s.Hl * s.Wl * HID;
    int n = 0;
    const short tl = (short)t.tail_kb;
    pre.steps[n++] = GemmStep{1, 2, 256, COL_X, 0, 2, tl};                            // lin_in; followed by the gather of Y_0 (K block 0)
    for (int b = 0; b < t.n_pre; ++b) {
        pre.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_NET, 0, 0, tl};             // fc_0[b]
        pre.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_X, 1, (short)(b + 1 < t.n_pre ? 1 : 3), tl};   // fc_1[b]; overlapped by the gather of Y_{b+1} / the next tile's Y_0
    }
    pre.n_steps = n;
    n = 0;
    for (int b = 0; b < t.n_post; ++b) {
        post.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_NET, 0, 0, tl};
        // fc_1[b]; the last one releases its K blocks to the next round's first PRE tile when that follows (FUSED, warm_rounds)
        post.steps[n++] = GemmStep{(short)kbh, 2, 256, COL_X, 1, (short)(b + 1 < t.n_post ? 0 : 3), tl};
    }
    post.n_steps = n;
    pre.w_out = post.w_out = m.w_out; pre.b_out = post.b_out = m.b_out;
    pre.NV = post.NV = NV;
    pre.NV_real = post.NV_real = s.NV;
    pre.spv = post.spv = ROWS / NV;
    pre.xc = post.xc = (float*)t.scratch;
    pre.out = post.out = q.out;
    pre.err = post.err = t.err_flag;
    pre.n_total = post.n_total = total;
    pre.dbg_skip = t.dbg_skip; post.dbg_skip = 0;
    // workers gather the K blocks that are released in the K-block-outer part of the running GEMM, helpers those of its tail
    pre.worker_kb_hi = t.tail_kb > 0 ? HID / KBLK - 1 - t.tail_kb : 4;
    pre.early_worker_kb_hi = t.early_split > 0 ? t.early_split : pre.worker_kb_hi;
    pre.early_lin = t.early_lin;
    static long long* dbg_ts = nullptr;      // device memory (managed memory would page-fault inside the kernel and distort the timeline)
    if ((t.dbg_skip & 512) && !dbg_ts) TCK(cudaMalloc((void**)&dbg_ts, 4096 * sizeof(long long)));
    if (t.dbg_skip & 512) TCK(cudaMemsetAsync(dbg_ts, 0, 4096 * sizeof(long long), st));
    pre.dbg_ts = (t.dbg_skip & 512) ? dbg_ts : nullptr; post.dbg_ts = nullptr;
    t.ms_pre = t.ms_post = 0.f;
    if (t.fused) {
        // ONE launch for the whole call: per 64 samples, NV PRE tiles then the POST tile; x_c goes through a 128 KiB slab per CTA
        // (L2 resident) instead of the sub-batch sized scratch, and the per-sample (rgb, sigma) rows are the only output
        Args f = pre;
        for (int i = 0; i < post.n_steps; ++i) f.steps_post[i] = post.steps[i];
        f.n_steps_post = post.n_steps; f.n_blocks_post = post.n_blocks; f.uses_post = post.uses_per_tile;
        f.tile_table_post = post.tile_table; f.bias_post = post.bias;
        f.pts = t.post_tiles > 0 ? t.post_tiles : 1;
        // 2-D tile order of the rays when the caller declared the ray list of every scene a row-major image (diner_render_image
        // does; diner_set_option("ray_image_width") for ray tensors): needs rays + depths (not explicit points) and whole tiles
        if (t.ray_image_w > 0 && q.rays && q.K > 0 && (q.n_per_sb % q.K) == 0) {
            const long long nr = q.n_per_sb / q.K;
            if (nr % t.ray_image_w == 0 && t.ray_image_w % PERM_T == 0 && (nr / t.ray_image_w) % PERM_T == 0) {
                f.perm_w = t.ray_image_w;
                f.perm_h = (int)(nr / t.ray_image_w);
            }
        }
        f.ppr = NV * f.pts;
        // warm rounds: x / net swap TMEM halves from tile to tile, and the POST tile must find x in half 0 -> even tile count
        f.warm_rounds = t.warm_rounds && (f.ppr % 2 == 0);
        if (f.warm_rounds) f.early_lin = 1;
        f.s_begin = 0; f.n_samples = total;
        f.n_tiles = (total + (long long)ROWS * f.pts - 1) / ((long long)ROWS * f.pts);
        const long long g = ((f.n_tiles + 1) / 2) * 2;
        const int grid = (int)(g < grid_cap ? g : grid_cap);
        const size_t slab = (size_t)grid * ROWS * f.pts * HID * sizeof(float);
        if (slab > t.scratch_bytes) {
            if (t.scratch) cudaFree(t.scratch);
            t.scratch = nullptr; t.scratch_bytes = 0;
            TCK(cudaMalloc(&t.scratch, slab));
            t.scratch_bytes = slab;
        }
        f.xc = (float*)t.scratch;
        if (t.timing) TCK(cudaEventRecord(t.ev[0], st));
        if (parity) TCK((launch<true, KIND_FUSED>(f, grid, st))); else TCK((launch<false, KIND_FUSED>(f, grid, st)));
        if (t.timing) {
            TCK(cudaEventRecord(t.ev[1], st));
            TCK(cudaEventSynchronize(t.ev[1]));
            TCK(cudaEventElapsedTime(&t.ms_pre, t.ev[0], t.ev[1]));
        }
        if (f.dbg_ts) TCK(tc2_dump_stamps(f.dbg_ts, st));
        return cudaSuccess;
    }
    for (long long s0 = 0; s0 < total; s0 += sub) {
        const long long ns = total - s0 < sub ? total - s0 : sub;
        pre.s_begin = post.s_begin = s0;
        pre.n_samples = post.n_samples = ns;
        pre.n_tiles = (ns + pre.spv - 1) / pre.spv;
        post.n_tiles = (ns + ROWS - 1) / ROWS;
        const long long g1 = ((pre.n_tiles + 1) / 2) * 2, g2 = ((post.n_tiles + 1) / 2) * 2;
        const int grid1 = (int)(g1 < grid_cap ? g1 : grid_cap), grid2 = (int)(g2 < grid_cap ? g2 : grid_cap);
        if (t.timing) TCK(cudaEventRecord(t.ev[0], st));
        if (parity) TCK((launch<true, KIND_PRE>(pre, grid1, st))); else TCK((launch<false, KIND_PRE>(pre, grid1, st)));
        if (t.timing) TCK(cudaEventRecord(t.ev[1], st));
        if (parity) TCK((launch<true, KIND_POST>(post, grid2, st))); else TCK((launch<false, KIND_POST>(post, grid2, st)));
        if (t.timing) {
            TCK(cudaEventRecord(t.ev[2], st));
            TCK(cudaEventSynchronize(t.ev[2]));
            float x = 0.f, y = 0.f;
            TCK(cudaEventElapsedTime(&x, t.ev[0], t.ev[1]));
            TCK(cudaEventElapsedTime(&y, t.ev[1], t.ev[2]));
            t.ms_pre += x; t.ms_post += y;
        }
    }
    if (pre.dbg_ts) TCK(tc2_dump_stamps(pre.dbg_ts, st));
    return cudaSuccess;
}

// profiling (DINER_TC_DBG_SKIP=512): timeline of CTA pair 0 in round TS_ROUND, PRE-tile periods of CTAs 0..3, per-CTA kernel cycles
static cudaError_t tc2_dump_stamps(const long long* dev, cudaStream_t st) {
    {
        static long long h[4096];
        TCK(cudaStreamSynchronize(st));
        TCK(cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost));
        for (int cta = 0; cta < 4; ++cta) {
            std::vector<long long> d;
            for (int i = 1; i < 511 && h[512 + cta * 512 + i]; ++i) d.push_back(h[512 + cta * 512 + i] - h[512 + cta * 512 + i - 1]);
            if (d.empty()) continue;
            std::vector<long long> s = d;
            std::sort(s.begin(), s.end());
            double mean = 0;
            for (long long v : d) mean += (double)v;
            fprintf(stderr, "[ts] cta %d PRE-tile periods (cycles, %d tiles): min %lld  p10 %lld  median %lld  mean %.0f  p90 %lld  max %lld | first 6:", cta,
                    (int)d.size(), s.front(), s[s.size() / 10], s[s.size() / 2], mean / d.size(), s[s.size() * 9 / 10], s.back());
            for (size_t i = 0; i < 6 && i < d.size(); ++i) fprintf(stderr, " %lld", d[i]);
            fprintf(stderr, "\n");
        }
        std::vector<long long> tot;
        for (int i = 0; i < 256; ++i) if (h[2560 + i]) tot.push_back(h[2560 + i]);
        if (!tot.empty()) {
            std::sort(tot.begin(), tot.end());
            fprintf(stderr, "[ts] kernel cycles per CTA (%d CTAs): min %lld  median %lld  max %lld  (max/min %.3f)\n", (int)tot.size(), tot.front(),
                    tot[tot.size() / 2], tot.back(), (double)tot.back() / (double)tot.front());
        }
        for (int cta = 0; cta < 2; ++cta) {
            const long long t0 = h[(cta * 4 + 1) * 64];
            fprintf(stderr, "[ts] cta %d worker-warp-0 stamps of round %d, cycles since the first:", cta, TS_ROUND);
            for (int i = 0; i < 40 && h[(cta * 4 + 1) * 64 + i]; ++i) fprintf(stderr, " %lld", h[(cta * 4 + 1) * 64 + i] - t0);
            fprintf(stderr, "\n");
            fprintf(stderr, "[ts] cta %d helper-warp-0 stamps of the last block (at the kb-0 wait, past it, features done, early gather done):", cta);
            for (int i = 0; i < 4; ++i) fprintf(stderr, " %lld", h[(cta * 4 + 2) * 64 + i] ? h[(cta * 4 + 2) * 64 + i] - t0 : 0);
            fprintf(stderr, "\n");
            if (cta == 0) {
                fprintf(stderr, "[ts] cta 0 mma stamps per step (half 0 ready, half 1 ready, committed, -), same origin:");
                for (int i = 0; i < 28; ++i) fprintf(stderr, " %lld", h[i] ? h[i] - t0 : 0);
                fprintf(stderr, "\n");
            }
        }
    }
    return cudaSuccess;
}
