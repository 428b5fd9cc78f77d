/* libdiner_b200 -- C ABI of the B200-native DINER volumetric-rendering hot path.
 *
 * The reference (malteprinzler/diner) is pure Python/PyTorch and has no FFI of its own; its plugin
 * boundary for this path is the pair of nn.Modules resolved by import_obj (src/util/import_helper.py:16-24)
 * at src/models/diner.py:47-48:
 *     src.models.nerf_renderer.NeRFRendererDGS   (configs/train_dtu.yaml:53)
 *     src.models.pixelnerf.PixelNeRF             (configs/train_dtu.yaml:32)
 * The entry points below are what those modules' methods bind through ctypes (diner_b200/capi.py);
 * each one names the reference method / lines it replaces.  Plain pointers and sizes only, no torch
 * types.  Unless a function is suffixed _host, every pointer is caller-owned DEVICE memory
 * (contiguous fp32) that must stay valid until the work queued on `stream` has completed; `stream`
 * is a cudaStream_t passed as void* (0 = default stream).
 *
 * Every function returns 0 on success and a negative DINER_E_* code on failure; the message is
 * available from diner_last_error() (thread-local).  Nothing throws across the boundary.
 * One call in flight per context (no internal locking).
 */
#ifndef DINER_B200_H
#define DINER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DINER_OK 0
#define DINER_E_INVALID (-1)      /* bad argument / unsupported shape */
#define DINER_E_CUDA (-2)         /* CUDA runtime error (message has the cudaError string) */
#define DINER_E_STATE (-3)        /* call order: weights / scene not set */
#define DINER_E_UNSUPPORTED (-4)  /* valid request the selected mode cannot serve */

/* Arithmetic mode of the per-sample MLP (ResnetFC). */
#define DINER_MODE_FP32 0     /* CUDA-core fp32 FMA: reference arithmetic, slow; parity anchor            */
#define DINER_MODE_PARITY 1   /* tcgen05 fp16x3 split (hi*hi + lo*hi + hi*lo, fp32 accum): ~1e-6 vs fp32  */
#define DINER_MODE_FAST 2     /* tcgen05 single-pass fp16, fp32 accum: PSNR-level agreement only          */

typedef struct diner_ctx diner_ctx;

const char* diner_last_error(void);
int diner_version(void);

/* Creates a context on CUDA device `device` (must be sm_100). */
int diner_create(diner_ctx** out, int device);
void diner_destroy(diner_ctx* ctx);

/* ResnetFC parameters -- replaces reading nerf.mlp_fine.* inside ResnetFC.forward (src/models/resnetfc.py:129-159).
 * Layout exactly as the reference state_dict: weight (out,in) row-major + bias (out).
 *   lin_in (d_hidden,d_in)  lin_out (d_out,d_hidden)
 *   fc0[b], fc1[b] (d_hidden,d_hidden) for b < n_blocks      (blocks.b.fc_0 / fc_1)
 *   lin_z[b] (d_hidden,d_latent)       for b < min(combine_layer,n_blocks)
 * The library keeps its own packed copies; call again whenever the parameters change. */
int diner_set_mlp(diner_ctx* ctx, int d_in, int d_latent, int d_hidden, int d_out, int n_blocks,
                  int combine_layer, const float* lin_in_w, const float* lin_in_b, const float* lin_out_w,
                  const float* lin_out_b, const float* const* fc0_w, const float* const* fc0_b,
                  const float* const* fc1_w, const float* const* fc1_b, const float* const* lin_z_w,
                  const float* const* lin_z_b, void* stream);

/* Scene state -- replaces what PixelNeRF.encode leaves on the modules (src/models/pixelnerf.py:44-51,
 * src/models/image_encoder.py:232-237,290-291):
 *   latent  (SB,NV,L,Hl,Wl) NCHW fp32 as the reference stores it (re-laid out to NHWC internally); with
 *           diner_set_option("latent_layout", 1 | 2) set beforehand it is read as channels-last (SB,NV,Hl,Wl,L) instead --
 *           1 = copied, 2 = borrowed (no copy: the pointer must then stay valid until the next diner_set_scene)
 *   depths, depths_std (SB,NV,1,H,W); normals (SB,NV,3,H,W)
 *   poses (SB,NV,4,4) world->cam; focal, c (SB,NV,2); image is W x H pixels
 *   feature_padding = image_padding / conv1 stride (image_encoder.py:58); num_freqs / freq_factor of
 *   the PositionalEncoding (src/models/positional_encoding.py:14-31). */
int diner_set_scene(diner_ctx* ctx, int SB, int NV, int L, int Hl, int Wl, int H, int W, const float* latent,
                    const float* depths, const float* depths_std, const float* normals, const float* poses,
                    const float* focal, const float* c, float feature_padding, int num_freqs,
                    float freq_factor, void* stream);

/* Optional injected noise (dense, indexed by logical position); any pointer may be NULL, in which
 * case counter-based noise derived from `seed` is used for that draw.
 *   u_coarse (SB,NR,C) U[0,1)  -> torch.rand_like  at nerf_renderer.py:57
 *   g_noise  (SB,NR,G) N(0,1)  -> torch.randn_like at nerf_renderer.py:188
 *   u_fill   (SB,NR,K) U[0,1)  -> torch.rand_like  at nerf_renderer.py:390 (by column after the sort) */
typedef struct diner_noise {
    const float* u_coarse;
    const float* g_noise;
    const float* u_fill;
    uint64_t seed;
    uint64_t ray_offset;   /* counter-based noise only: logical index (within its scene) of the call's first ray, so that a shard /
                              chunk of a ray list draws the same noise as the whole list would (multi-GPU ray sharding) */
} diner_noise;

/* NeRFRendererDGS.forward (src/models/nerf_renderer.py:399-424): rays (SB,NR,8) = [o3,d3,near,far] ->
 * rgb (SB,NR,3), depth (SB,NR); weights (SB,NR,K) and z (SB,NR,K) are optional outputs (NULL to skip).
 * K = n_samples, C = n_depth_candidates, G = n_gaussian are read per call because the reference CLI
 * mutates them on the live module (python_scripts/create_prediction_folder.py:44-47). */
int diner_render(diner_ctx* ctx, const float* rays, int SB, int NR, int K, int C, int G, int white_bkgd,
                 int mode, const diner_noise* noise, float* rgb, float* depth, float* weights, float* z,
                 void* stream);

/* Same render with the outputs packed as rgbd (SB,NR,4) = [r,g,b,depth] per ray (16-byte aligned): the layout the multi-GPU
 * image all-gather sends (diner_b200/multi_gpu.py; SURVEY 8(e): one ncclAllGather of rgb|depth per image), written in place into
 * the rank's slice of the gather buffer so that no pack / unpack pass runs between the compositing kernel and the collective. */
int diner_render_rgbd(diner_ctx* ctx, const float* rays, int SB, int NR, int K, int C, int G, int white_bkgd, int mode,
                      const diner_noise* noise, float* rgbd, void* stream);

/* Whole-image entry: ray generation (gen_rays, src/util/cam_geometry.py:5-48) + the ray_batch_size chunk loop and torch.cat of
 * DINER.predict_imgs_from_batch (src/models/diner.py:79-92) in one call.  target_extrinsics (SB,4,4) world->cam, target_intrinsics
 * (SB,3,3), device fp32; rays are generated on the device for every pixel centre of the H x W target view (row-major) and never
 * leave the library; rgb (SB,H*W,3), depth (SB,H*W). */
int diner_render_image(diner_ctx* ctx, const float* target_extrinsics, const float* target_intrinsics, int SB, int H, int W,
                       float z_near, float z_far, int K, int C, int G, int white_bkgd, int mode, const diner_noise* noise,
                       float* rgb, float* depth, void* stream);
/* gen_rays alone (src/util/cam_geometry.py:5-48): rays (SB,H*W,8) = [origin3, dir3, near, far]. */
int diner_gen_rays(diner_ctx* ctx, const float* target_extrinsics, const float* target_intrinsics, int SB, int H, int W,
                   float z_near, float z_far, float* rays, void* stream);

/* Scene prepare: depth2normal (src/util/depth2normal.py:6-87, called by PixelNeRF.encode, pixelnerf.py:41-42):
 * depths (N,1,H,W), intrinsics (N,3,3) -> normals (N,3,H,W), zero where there is no depth. */
int diner_depth2normal(diner_ctx* ctx, const float* depths, const float* intrinsics, int N, int H, int W, float* normals,
                       void* stream);

/* Backward of the render path for the training step (src/models/diner.py:257-266: MSE on the rendered colours, autograd
 * through NeRFRendererDGS.composite / PixelNeRF.forward / ResnetFC; the sampler is @torch.no_grad).  fp32 CUDA cores (the
 * correctness anchor for a tcgen05 version), checked against the reference's own autograd gradients in
 * tests/test_gpu_parity.py::test_backward_*.  Given the sample depths z (SB,NR,K) of the forward call and the upstream gradients
 * g_rgb (SB,NR,3), g_depth (SB,NR) or NULL, ACCUMULATES
 *   grad_params: diner_mlp_param_count() floats in the order of diner_set_mlp's arguments (lin_in w,b; lin_out w,b; per block
 *                fc_0 w,b, fc_1 w,b; per lin_z block w,b)
 *   d_latent:    (SB,NV,L,Hl,Wl) NCHW like the latent passed to diner_set_scene, or NULL. */
int diner_render_backward(diner_ctx* ctx, const float* rays, const float* z, int SB, int NR, int K, int white_bkgd,
                          const float* g_rgb, const float* g_depth, float* grad_params, float* d_latent, void* stream);
long long diner_mlp_param_count(diner_ctx* ctx);

/* Same call with HOST buffers (rays in, rgb/depth out): copies host->device, renders, copies back and
 * synchronises the stream.  This is the end-to-end entry a non-torch caller uses. */
int diner_render_host(diner_ctx* ctx, const float* rays_host, int SB, int NR, int K, int C, int G,
                      int white_bkgd, int mode, uint64_t seed, float* rgb_host, float* depth_host,
                      void* stream);

/* Stage entry points (used by the module methods and the stage-wise parity tests). */
/* sample_depthguided + fill_up_uniform_samples (nerf_renderer.py:65-190, :367-397) -> z (SB,NR,K) ascending;
 * z_dgs (optional): depth-guided result before the fill, ascending, 0 = empty slot. */
int diner_sample(diner_ctx* ctx, const float* rays, int SB, int NR, int K, int C, int G,
                 const diner_noise* noise, float* z, float* z_dgs, void* stream);
/* PixelNeRF.forward (src/models/pixelnerf.py:55-145): xyz, viewdirs (SB,B,3) -> out (SB,B,4). */
int diner_query(diner_ctx* ctx, const float* xyz, const float* viewdirs, int SB, long long B, int mode,
                float* out, void* stream);
/* NeRFRendererDGS.composite (nerf_renderer.py:286-365) on given sample depths z (SB,NR,K). */
int diner_composite(diner_ctx* ctx, const float* rays, const float* z, int SB, int NR, int K, int white_bkgd,
                    int mode, float* rgb, float* depth, float* weights, void* stream);

/* Tuning knobs of the tcgen05 path (not part of the reference API): key = "fused" (1, default: one launch per call that runs the
 * per sample-view layers AND the per sample layers, the view-combined activations staying in an L2-resident slab; 0: separate
 * PRE / POST launches with an HBM scratch between them), "post_tiles" (1..8: 64-sample tiles per round of the fused launch),
 * "ray_image_width" (W > 0 declares that the rays of every scene passed to the render calls form a row-major image of that width,
 * e.g. the reference's gen_rays output: the MLP launch then walks them in 16 x 16 pixel tiles, which keeps the gathered feature
 * map lines in L2; results do not change; 0 = off; diner_render_image sets it by itself), "tail_kb" (0..4: K blocks of every GEMM step issued
 * N-tile-outer so that the first epilogue half overlaps the step's tail), "early_split" (worker/helper split of the next tile's
 * early gather), "warm_rounds" (1, default: in the fused launch the next round's first per-sample-view tile is prepared under the last
 * GEMM of the per-sample tile, so only the first tile of a launch starts cold; needs an even number of tiles per round, i.e. more
 * than one view), "early_lin" (the next tile's lin_in is issued behind the current tile's last GEMM into the other TMEM half; implied by
 * warm_rounds; results do not change with either), "sub_batch" (samples per PRE/POST launch pair), "rebuild_maps" (forces the next query to rebuild the hoisted
 * lin_z maps) or "latent_layout" (see diner_set_scene). */
int diner_set_option(diner_ctx* ctx, const char* key, long long value);
/* Reference arguments that are not per-call sizes: key = "depth_diff_max" (sample_depthguided(..., depth_diff_max=0.05),
 * src/models/nerf_renderer.py:66,121) or "softplus_beta" (ResnetFC(beta=...): Softplus(beta) activations instead of ReLU,
 * src/models/resnetfc.py:124-127; 0 = ReLU; served by DINER_MODE_FP32 only, the tcgen05 modes and the backward return
 * DINER_E_UNSUPPORTED). */
int diner_set_float_option(diner_ctx* ctx, const char* key, double value);

/* cudaDeviceSynchronize + decoded tcgen05 watchdog code on failure (debugging aid). */
int diner_debug_sync(diner_ctx* ctx);

/* Number of kernels this library launched on behalf of ctx since creation (bench.py's gpu_launches). */
long long diner_launch_count(diner_ctx* ctx);
/* Device time in ms of the MLP kernels of the last render/composite/query call when timing was
 * enabled with diner_set_timing(ctx, 1) (CUDA events on the caller's stream; forces a sync). */
int diner_set_timing(diner_ctx* ctx, int enabled);
float diner_last_mlp_ms(diner_ctx* ctx);
/* Per-stage device time of the last call (timing enabled): 0 sampler, 1 MLP PRE kernel(s) (per sample-view layers),
 * 2 MLP POST kernel(s) (per sample layers), 3 compositing, 4 the once-per-(scene, weights) build of the hoisted lin_z maps
 * (runs inside the first query after diner_set_scene / diner_set_mlp). */
float diner_last_stage_ms(diner_ctx* ctx, int stage);

#ifdef __cplusplus
}
#endif
#endif /* DINER_B200_H */
