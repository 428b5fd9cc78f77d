// Internal (non-ABI) declarations shared by the .cu files of libdiner_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

struct SamplerArgs {
    const float* rays;       // (SB*NR, 8)
    int SB, NR, K, C, G;
    const float* u_coarse;   // (SB*NR, C)  or nullptr -> counter-based noise from `seed`
    const float* g_noise;    // (SB*NR, G)  or nullptr
    const float* u_fill;     // (SB*NR, K)  or nullptr
    uint64_t seed, ray_offset;   // counter-based noise: key and logical index of the call's first ray (per scene)
    float lin_step, lin_end, cstep;   // torch.linspace(0, 1-1/C, C) parameters, fp32(1/C)
    float depth_diff_max;    // |d_ref - z_c| gate of the likelihood mask (nerf_renderer.py:66,121; default 0.05)
    float* z_out;            // (SB*NR, K) ascending
    float* z_dgs;            // optional (SB*NR, K): depth-guided samples before fill-up, ascending, 0 = empty
};

// ResnetFC parameters as the reference lays them out (resnetfc.py:92-118); fp32 device pointers
// owned by the context (copied at diner_set_mlp).
#define DINER_MAX_BLOCKS 8
struct MlpDev {
    int d_in, d_latent, d_hidden, d_out, n_blocks, combine_layer;
    float beta;              // > 0: Softplus(beta) activations instead of ReLU (resnetfc.py:124-127); fp32 mode only
    const float *w_in, *b_in, *w_out, *b_out;
    const float *w_fc0[DINER_MAX_BLOCKS], *b_fc0[DINER_MAX_BLOCKS];
    const float *w_fc1[DINER_MAX_BLOCKS], *b_fc1[DINER_MAX_BLOCKS];
    const float *w_z[DINER_MAX_BLOCKS], *b_z[DINER_MAX_BLOCKS];
};

cudaError_t launch_sampler(const SceneDev& s, const SamplerArgs& a, int num_sms, cudaStream_t st);

// --- per-sample network query, fp32 CUDA-core path (mlp_simt.cu) ---------------------------------
// out (n_samples, 4) = [sigmoid rgb, relu sigma] for samples given as (ray, z) pairs or explicit points.
struct QueryArgs {
    int SB;
    long long n_per_sb;        // samples per scene (B of PixelNeRF.forward; NR*K for the renderer)
    // either explicit points ...
    const float* xyz;          // (SB*n_per_sb, 3) or nullptr
    const float* viewdirs;     // (SB*n_per_sb, 3) or nullptr
    // ... or rays + depths
    const float* rays;         // (SB*NR, 8)
    const float* z;            // (SB*NR, K)
    int K;
    float* out;                // (SB*n_per_sb, 4)
};
struct SimtWorkspace {
    float *xin, *zlat, *x, *net, *xc;
    long long rows_cap;        // sample-view rows the buffers hold
};
cudaError_t query_simt(const SceneDev& s, const MlpDev& m, const QueryArgs& q, SimtWorkspace& ws,
                       cudaStream_t st);
size_t simt_workspace_bytes(const MlpDev& m, long long rows);

// --- backward pass, fp32 CUDA cores (backward_simt.cu; experimental) ------------------------------------------------------
size_t backward_param_count(const MlpDev& m);
size_t backward_workspace_bytes(const MlpDev& m, const SceneDev& s, long long chunk_samples);
struct TcState;
size_t backward_tc_workspace_bytes(const MlpDev& m, const SceneDev& s, long long chunk_samples);
cudaError_t backward_simt(const SceneDev& s, const MlpDev& m, const QueryArgs& q, const float* d_pre, float* grad_params,
                          float* d_latent, float* ws, long long chunk, cudaStream_t st, TcState* tcs = nullptr, uint8_t* tcws = nullptr,
                          int num_sms = 148);
cudaError_t launch_composite_backward(const float* rays, const float* z, const float* net_out, long long n_rays, int K, int white,
                                      const float* g_rgb, const float* g_depth, float* d_pre, cudaStream_t st);

// --- alpha compositing (composite.cu) -------------------------------------------------------------
cudaError_t launch_composite(const float* rays, const float* z, const float* net_out, long long n_rays,
                             int K, int white_bkgd, float* rgb, float* depth, float* weights,
                             cudaStream_t st, float* rgbd = nullptr);

// --- scene re-layout (scene.cu) ---------------------------------------------------------------------
cudaError_t launch_nchw_to_nhwc(const float* src, float* dst, int N, int C, int HW, cudaStream_t st);
cudaError_t launch_gen_rays(const float* ext, const float* intr, int SB, int H, int W, float z_near, float z_far, float* rays,
                            int num_sms, cudaStream_t st);
cudaError_t launch_depth2normal(const float* depth, const float* intr, int N, int H, int W, float* normals, int num_sms, cudaStream_t st);
cudaError_t upload_std_ring_gain();
extern long long g_launches;   // bumped at every kernel launch site
