// tcgen05 (5th-gen tensor core) path of the per-sample network: packed weights + launch entry points.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "diner_internal.h"

struct TcState {
    bool ready = false;
    char why[160] = "diner_set_mlp not called";
    void* wpack = nullptr;        // fp16 hi/lo weight tiles (scaled by tc::W_SCALE) in UMMA K-major SWIZZLE_128B layout
    size_t wpack_bytes = 0;
    float* bias = nullptr;        // per-phase cumulative bias vectors
    void* scratch = nullptr;      // combined activations x_c between the pre- and post-combine kernels
    size_t scratch_bytes = 0;
    int* err_flag = nullptr;      // device-side watchdog flag (mbarrier timeouts)
    int n_pre = 0, n_post = 0;    // ResnetFC blocks before / after the view-combine
    int pairs_pre = 0, pairs_post = 0;   // (hi,lo) weight tile pairs per CTA tile of the PRE / POST kernel
    size_t bias_post_off = 0, bias_pair_off = 0;   // float offsets of the POST rows / the pair kernel's folded PRE rows in `bias`
    long long sub_batch = 0;      // samples per PRE/POST launch pair (0 = default)
    int* table2 = nullptr;        // pair kernel: per-rank weight tile tables
    int table2_parity = -1, table2_tail = -1, uses2_zmap = 0, uses2_pre = 0, uses2_post = 0, max_grid2 = 0;
    int fused = 1;                // 1 = one FUSED launch per call (PRE tiles + POST tile per 64 samples, x_c in an L2-resident slab); 0 = PRE / POST launches
    int post_tiles = 1;           // FUSED: POST tiles (of 64 samples) per round = per cold start of the PRE pipeline; slab = 128 KiB x this per CTA
    int ray_image_w = 0;          // > 0: the ray list of each scene is a row-major image of this width -> the fused launch walks it in 16 x 16 pixel tiles
    int tail_kb = 3;              // last K blocks of every full-width GEMM step issued N-tile-outer (see mlp_tc2.cu "accumulator halves")
    float* zmap = nullptr;        // pair kernel: Y_b = lin_z[b](latent) maps, [block][pixel][512] fp32 (hoisted lin_z, see mlp_tc2.cu)
    size_t zmap_bytes = 0;
    bool zmap_valid = false;      // cleared by diner_set_mlp / diner_set_scene; rebuilt lazily by the next query
    float ms_zmap = 0.f;          // device time of the last Y-map build (timing enabled)
    CUtensorMap wmap;             // 2-D view of wpack: rows of 128 B, box = one 16 KiB tile (pair kernel: cta_group::2 TMA)
    bool wmap_ok = false;
    int warm_rounds = 1;          // fused launch: next round's first PRE tile prepared under the POST tile's last fc_1 (implies early_lin)
    int early_lin = 0;            // pair kernel PRE tiles: next tile's lin_in issued behind the last fc_1 into the other TMEM half
    int early_split = 0;          // pair kernel PRE: worker/helper split of the next tile's early Y_0 gather (0 = same as inside a tile)
    int dbg_skip = 0;             // profiling experiments (pair kernel PRE): see tc2::Args::dbg_skip
    bool timing = false;          // per-kernel CUDA-event timing (adds one sync per sub-batch)
    float ms_pre = 0.f, ms_post = 0.f;   // accumulated over the last tc_query call
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

cudaError_t tc_pack_weights(TcState& t, const MlpDev& m, cudaStream_t st);
void tc_release(TcState& t);
cudaError_t tc2_query(TcState& t, const SceneDev& s, const MlpDev& m, const QueryArgs& q, bool parity,
                      int num_sms, cudaStream_t st);
