// extern "C" boundary of libdiner_b200.so (see include/diner_b200.h for the contract).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

#include "../../include/diner_b200.h"
#include "diner_internal.h"
#include "mlp_tc.h"

static thread_local char g_err[512] = "";
long long g_launches = 0;   // bumped by every launch wrapper (single-threaded use per the contract)

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(DINER_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                        __FILE__, __LINE__);                                                    \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return (T*)p; }
};

struct diner_ctx {
    int device = 0, num_sms = 0;
    bool has_mlp = false, has_scene = false;
    MlpDev mlp{};
    DevBuf mlp_store;                // all fp32 parameters, contiguous
    SceneDev scene{};
    DevBuf latent, maps, cams;       // library-owned scene copies
    DevBuf zbuf, netbuf, simt_ws, rays_dev, out_dev, rays_img, bwd_ws, bwd_tc, dpre;
    int backward_tc = 1;             // training-step backward: 1 = tcgen05 GEMMs for the 512-wide layers (gemm_tc3.cu), 0 = fp32 CUDA cores
    void* host_pin = nullptr; size_t host_pin_cap = 0;
    TcState tc;                      // packed weights + scratch of the tcgen05 path
    long long launches = 0;
    int latent_layout = 0;           // how the next diner_set_scene reads `latent`: 0 NCHW (transposed into a library copy), 1 NHWC (copied), 2 NHWC borrowed
    float depth_diff_max = 0.05f;    // nerf_renderer.py:66 default
    float softplus_beta = 0.0f;      // resnetfc.py:124-127: > 0 selects Softplus(beta) activations (fp32 mode only)
    int timing = 0;
    float last_mlp_ms = 0.f, last_sampler_ms = 0.f, last_composite_ms = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
};

extern "C" const char* diner_last_error(void) { return g_err; }
extern "C" int diner_version(void) { return 1; }

extern "C" int diner_create(diner_ctx** out, int device) {
    if (!out) return fail(DINER_E_INVALID, "out is NULL");
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(DINER_E_INVALID, "device %d out of range (%d devices)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DINER_E_UNSUPPORTED, "device %d is sm_%d%d; libdiner_b200 is built for sm_100a only", device,
                    prop.major, prop.minor);
    diner_ctx* c = new (std::nothrow) diner_ctx();
    if (!c) return fail(DINER_E_INVALID, "out of host memory");
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    CUDA_TRY(upload_std_ring_gain());
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaEventCreate(&c->ev2));
    CUDA_TRY(cudaEventCreate(&c->ev3));
    *out = c;
    return DINER_OK;
}

extern "C" void diner_destroy(diner_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    c->mlp_store.release(); c->latent.release(); c->maps.release(); c->cams.release();
    c->zbuf.release(); c->netbuf.release(); c->simt_ws.release(); c->rays_dev.release(); c->out_dev.release(); c->rays_img.release(); c->bwd_ws.release(); c->bwd_tc.release(); c->dpre.release();
    tc_release(c->tc);
    if (c->host_pin) cudaFreeHost(c->host_pin);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    delete c;
}

extern "C" int diner_set_mlp(diner_ctx* c, int d_in, int d_latent, int d_hidden, int d_out, int n_blocks,
                             int combine_layer, const float* lin_in_w, const float* lin_in_b,
                             const float* lin_out_w, const float* lin_out_b, const float* const* fc0_w,
                             const float* const* fc0_b, const float* const* fc1_w, const float* const* fc1_b,
                             const float* const* lin_z_w, const float* const* lin_z_b, void* stream) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    if (n_blocks < 1 || n_blocks > DINER_MAX_BLOCKS) return fail(DINER_E_INVALID, "n_blocks %d not in [1,%d]", n_blocks, DINER_MAX_BLOCKS);
    if (d_out != 4) return fail(DINER_E_INVALID, "d_out must be 4 (rgb+sigma), got %d", d_out);
    if (d_in < 1 || d_latent < 1 || d_hidden < 1 || (d_latent % 4)) return fail(DINER_E_INVALID, "bad dims d_in=%d d_latent=%d d_hidden=%d", d_in, d_latent, d_hidden);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(c->device));
    const int nz = combine_layer < n_blocks ? combine_layer : n_blocks;
    size_t total = (size_t)d_hidden * d_in + d_hidden + (size_t)d_out * d_hidden + d_out +
                   (size_t)n_blocks * 2 * ((size_t)d_hidden * d_hidden + d_hidden) +
                   (size_t)nz * ((size_t)d_hidden * d_latent + d_hidden);
    CUDA_TRY(c->mlp_store.reserve(total * sizeof(float)));
    float* p = c->mlp_store.as<float>();
    auto put = [&](const float* src, size_t n, const float*& dst) -> cudaError_t {
        if (!src) return cudaErrorInvalidValue;
        dst = p;
        cudaError_t e = cudaMemcpyAsync(p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
        p += n;
        return e;
    };
    MlpDev m{};
    m.d_in = d_in; m.d_latent = d_latent; m.d_hidden = d_hidden; m.d_out = d_out;
    m.n_blocks = n_blocks; m.combine_layer = combine_layer;
    CUDA_TRY(put(lin_in_w, (size_t)d_hidden * d_in, m.w_in));
    CUDA_TRY(put(lin_in_b, d_hidden, m.b_in));
    CUDA_TRY(put(lin_out_w, (size_t)d_out * d_hidden, m.w_out));
    CUDA_TRY(put(lin_out_b, d_out, m.b_out));
    for (int b = 0; b < n_blocks; ++b) {
        CUDA_TRY(put(fc0_w[b], (size_t)d_hidden * d_hidden, m.w_fc0[b]));
        CUDA_TRY(put(fc0_b[b], d_hidden, m.b_fc0[b]));
        CUDA_TRY(put(fc1_w[b], (size_t)d_hidden * d_hidden, m.w_fc1[b]));
        CUDA_TRY(put(fc1_b[b], d_hidden, m.b_fc1[b]));
    }
    for (int b = 0; b < nz; ++b) {
        CUDA_TRY(put(lin_z_w[b], (size_t)d_hidden * d_latent, m.w_z[b]));
        CUDA_TRY(put(lin_z_b[b], d_hidden, m.b_z[b]));
    }
    c->mlp = m;
    c->has_mlp = true;
    // tensor-core packing (fp16 hi/lo tiles of 64 w in UMMA layout); shapes it cannot serve leave tc.ready = false
    cudaError_t e = tc_pack_weights(c->tc, m, st);
    if (e != cudaSuccess) return fail(DINER_E_CUDA, "tc_pack_weights: %s", cudaGetErrorString(e));
    return DINER_OK;
}

extern "C" int diner_set_scene(diner_ctx* c, int SB, int NV, int L, int Hl, int Wl, int H, int W,
                               const float* latent, const float* depths, const float* depths_std,
                               const float* normals, const float* poses, const float* focal, const float* cc,
                               float feature_padding, int num_freqs, float freq_factor, void* stream) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    if (SB < 1 || NV < 1 || L < 4 || (L % 4) || Hl < 2 || Wl < 2 || H < 1 || W < 1)
        return fail(DINER_E_INVALID, "bad scene dims SB=%d NV=%d L=%d Hl=%d Wl=%d H=%d W=%d", SB, NV, L, Hl, Wl, H, W);
    if (num_freqs < 0 || num_freqs > DINER_MAX_FREQS) return fail(DINER_E_INVALID, "num_freqs %d > %d", num_freqs, DINER_MAX_FREQS);
    if (!latent || !depths || !depths_std || !normals || !poses || !focal || !cc) return fail(DINER_E_INVALID, "NULL scene pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t nimg = (size_t)SB * NV, lat_n = nimg * L * Hl * Wl, px = nimg * H * W;
    const float* lat_dev = nullptr;
    if (c->latent_layout == 2) {                    // channels-last, borrowed: valid until the next diner_set_scene (caller keeps it alive)
        c->latent.release();
        lat_dev = latent;
    } else {
        CUDA_TRY(c->latent.reserve(lat_n * sizeof(float)));
        if (c->latent_layout == 1) {
            CUDA_TRY(cudaMemcpyAsync(c->latent.p, latent, lat_n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else {
            CUDA_TRY(launch_nchw_to_nhwc(latent, c->latent.as<float>(), (int)nimg, L, Hl * Wl, st));
            g_launches++;
        }
        lat_dev = c->latent.as<float>();
    }
    CUDA_TRY(c->maps.reserve(px * 5 * sizeof(float)));
    float* mp = c->maps.as<float>();
    CUDA_TRY(cudaMemcpyAsync(mp, depths, px * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(mp + px, depths_std, px * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(mp + 2 * px, normals, 3 * px * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(c->cams.reserve(nimg * 20 * sizeof(float)));
    float* cp = c->cams.as<float>();
    CUDA_TRY(cudaMemcpyAsync(cp, poses, nimg * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(cp + nimg * 16, focal, nimg * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(cp + nimg * 18, cc, nimg * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SceneDev s{};
    s.SB = SB; s.NV = NV; s.L = L; s.Hl = Hl; s.Wl = Wl; s.H = H; s.W = W;
    s.latent = lat_dev;
    s.depth = mp; s.dstd = mp + px; s.normal = mp + 2 * px;
    s.poses = cp; s.focal = cp + nimg * 16; s.cxy = cp + nimg * 18;
    s.imgW = (float)W; s.imgH = (float)H;
    s.lat_sx = ((float)Wl - feature_padding * 2.0f) / (float)Wl;
    s.lat_sy = ((float)Hl - feature_padding * 2.0f) / (float)Hl;
    s.std_sx = (float)W / ((float)W + 2.0f * STD_PAD);
    s.std_sy = (float)H / ((float)H + 2.0f * STD_PAD);
    s.num_freqs = num_freqs;
    float pw = 1.0f;
    for (int i = 0; i < num_freqs; ++i) { s.freqs[i] = freq_factor * pw; pw *= 2.0f; }
    c->scene = s;
    c->has_scene = true;
    c->tc.zmap_valid = false;        // hoisted lin_z maps are per scene
    return DINER_OK;
}

static int check_ready(diner_ctx* c) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    if (!c->has_mlp) return fail(DINER_E_STATE, "diner_set_mlp has not been called");
    if (!c->has_scene) return fail(DINER_E_STATE, "diner_set_scene has not been called (PixelNeRF.encode first)");
    // the MLP and the scene arrive independently through the ABI: every mode sizes its buffers from both, so they must agree
    // before anything is launched (pixelnerf.py:23-24: d_latent = encoder.latent_size, d_in = poscode + depthcode + 3)
    if (c->scene.L != c->mlp.d_latent)
        return fail(DINER_E_INVALID, "the scene has %d latent channels but the MLP expects d_latent=%d", c->scene.L, c->mlp.d_latent);
    const int d_in = 3 + 6 * c->scene.num_freqs + 3 + 1 + 2 * c->scene.num_freqs;
    if (d_in != c->mlp.d_in)
        return fail(DINER_E_INVALID, "the positional code (num_freqs=%d, include_input) gives d_in=%d but lin_in expects %d",
                    c->scene.num_freqs, d_in, c->mlp.d_in);
    return DINER_OK;
}

// Runs the network on (ray, z) samples or explicit points into `out` (n,4) in the requested mode.
static int run_query(diner_ctx* c, const QueryArgs& q, int mode, cudaStream_t st) {
    const long long total = (long long)q.SB * q.n_per_sb;
    if (total == 0) return DINER_OK;
    if (q.SB != c->scene.SB) return fail(DINER_E_INVALID, "SB=%d but the encoded scene has %d objects", q.SB, c->scene.SB);
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev0, st));
    c->mlp.beta = c->softplus_beta;
    if (c->softplus_beta > 0.0f && mode != DINER_MODE_FP32)
        return fail(DINER_E_UNSUPPORTED, "Softplus activations (beta=%g) are served by DINER_MODE_FP32 only", (double)c->softplus_beta);
    if (mode == DINER_MODE_FP32) {
        long long rows = total * c->scene.NV;
        const long long cap_rows = 32768LL * c->scene.NV;
        if (rows > cap_rows) rows = cap_rows;
        CUDA_TRY(c->simt_ws.reserve(simt_workspace_bytes(c->mlp, rows)));
        SimtWorkspace ws;
        const int ld_in = (c->mlp.d_in + 7) & ~7;
        float* p = c->simt_ws.as<float>();
        ws.xin = p; p += rows * ld_in;
        ws.zlat = p; p += rows * c->mlp.d_latent;
        ws.x = p; p += rows * c->mlp.d_hidden;
        ws.net = p; p += rows * c->mlp.d_hidden;
        ws.xc = p;
        ws.rows_cap = rows;
        CUDA_TRY(query_simt(c->scene, c->mlp, q, ws, st));
    } else if (mode == DINER_MODE_PARITY || mode == DINER_MODE_FAST) {
        if (!c->tc.ready)
            return fail(DINER_E_UNSUPPORTED, "tensor-core path unavailable for this MLP shape: %s", c->tc.why);
        cudaError_t e = tc2_query(c->tc, c->scene, c->mlp, q, mode == DINER_MODE_PARITY, c->num_sms, st);
        if (e == cudaErrorNotSupported) return fail(DINER_E_UNSUPPORTED, "tensor-core path: %s", c->tc.why);
        if (e != cudaSuccess) return fail(DINER_E_CUDA, "tc2_query: %s (watchdog code %d)", cudaGetErrorString(e), c->tc.err_flag ? *c->tc.err_flag : -1);
    } else {
        return fail(DINER_E_INVALID, "unknown mode %d", mode);
    }
    if (c->timing) {
        CUDA_TRY(cudaEventRecord(c->ev1, st));
        CUDA_TRY(cudaEventSynchronize(c->ev1));
        CUDA_TRY(cudaEventElapsedTime(&c->last_mlp_ms, c->ev0, c->ev1));
    }
    return DINER_OK;
}

static int check_render_args(diner_ctx* c, int SB, int NR, int K, int C, int G) {
    if (SB < 1 || NR < 0 || K < 1) return fail(DINER_E_INVALID, "bad SB=%d NR=%d K=%d", SB, NR, K);
    if (K > 1024) return fail(DINER_E_INVALID, "n_samples %d > 1024 not supported", K);
    if (C >= 0 && (C < 1 || C > 4096)) return fail(DINER_E_INVALID, "n_depth_candidates %d not in [1,4096]", C);
    if (G >= 0 && G > K) return fail(DINER_E_INVALID, "n_gaussian %d > n_samples %d (reference asserts n_samples >= n_gaussian)", G, K);
    if (SB != c->scene.SB) return fail(DINER_E_INVALID, "rays have SB=%d but the encoded scene has %d objects", SB, c->scene.SB);
    return DINER_OK;
}

static int do_sample(diner_ctx* c, const float* rays, int SB, int NR, int K, int C, int G,
                     const diner_noise* noise, float* z, float* z_dgs, cudaStream_t st) {
    SamplerArgs a{};
    a.rays = rays; a.SB = SB; a.NR = NR; a.K = K; a.C = C; a.G = G;
    a.u_coarse = noise ? noise->u_coarse : nullptr;
    a.g_noise = noise ? noise->g_noise : nullptr;
    a.u_fill = noise ? noise->u_fill : nullptr;
    a.seed = noise ? noise->seed : 0;
    a.ray_offset = noise ? noise->ray_offset : 0;
    const float end = (float)(1.0 - 1.0 / (double)C);          // torch.linspace(0, 1 - step, C)
    a.lin_end = end;
    a.lin_step = C > 1 ? end / (float)(C - 1) : 0.0f;
    a.cstep = (float)(1.0 / (double)C);
    a.depth_diff_max = c->depth_diff_max;
    a.z_out = z; a.z_dgs = z_dgs;
    if (NR == 0) return DINER_OK;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev2, st));
    CUDA_TRY(launch_sampler(c->scene, a, c->num_sms, st));
    g_launches++;
    if (c->timing) {
        CUDA_TRY(cudaEventRecord(c->ev3, st));
        CUDA_TRY(cudaEventSynchronize(c->ev3));
        CUDA_TRY(cudaEventElapsedTime(&c->last_sampler_ms, c->ev2, c->ev3));
    }
    return DINER_OK;
}

extern "C" int diner_sample(diner_ctx* c, const float* rays, int SB, int NR, int K, int C, int G,
                            const diner_noise* noise, float* z, float* z_dgs, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_render_args(c, SB, NR, K, C, G))) return rc;
    if ((long long)SB * NR == 0) return DINER_OK;
    if (!rays || !z) return fail(DINER_E_INVALID, "NULL rays / z");
    CUDA_TRY(cudaSetDevice(c->device));
    const long long l0 = g_launches;
    rc = do_sample(c, rays, SB, NR, K, C, G, noise, z, z_dgs, (cudaStream_t)stream);
    c->launches += g_launches - l0;
    return rc;
}

extern "C" int diner_query(diner_ctx* c, const float* xyz, const float* viewdirs, int SB, long long B, int mode,
                           float* out, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (!xyz || !viewdirs || !out) return fail(DINER_E_INVALID, "NULL xyz / viewdirs / out");
    if (SB < 1 || B < 0) return fail(DINER_E_INVALID, "bad SB=%d B=%lld", SB, B);
    CUDA_TRY(cudaSetDevice(c->device));
    QueryArgs q{};
    q.SB = SB; q.n_per_sb = B; q.xyz = xyz; q.viewdirs = viewdirs; q.out = out; q.K = 1;
    const long long l0 = g_launches;
    rc = run_query(c, q, mode, (cudaStream_t)stream);
    c->launches += g_launches - l0;
    return rc;
}

static int do_composite(diner_ctx* c, const float* rays, const float* z, int SB, int NR, int K, int white,
                        int mode, float* rgb, float* depth, float* weights, cudaStream_t st, float* rgbd = nullptr) {
    const long long n_rays = (long long)SB * NR;
    if (n_rays == 0) return DINER_OK;
    CUDA_TRY(c->netbuf.reserve((size_t)n_rays * K * 4 * sizeof(float)));
    QueryArgs q{};
    q.SB = SB; q.n_per_sb = (long long)NR * K; q.rays = rays; q.z = z; q.K = K;
    q.out = c->netbuf.as<float>();
    int rc = run_query(c, q, mode, st);
    if (rc) return rc;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev2, st));
    CUDA_TRY(launch_composite(rays, z, q.out, n_rays, K, white, rgb, depth, weights, st, rgbd));
    g_launches++;
    if (c->timing) {
        CUDA_TRY(cudaEventRecord(c->ev3, st));
        CUDA_TRY(cudaEventSynchronize(c->ev3));
        CUDA_TRY(cudaEventElapsedTime(&c->last_composite_ms, c->ev2, c->ev3));
    }
    return DINER_OK;
}

extern "C" int diner_composite(diner_ctx* c, const float* rays, const float* z, int SB, int NR, int K,
                               int white_bkgd, int mode, float* rgb, float* depth, float* weights, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_render_args(c, SB, NR, K, -1, -1))) return rc;
    if ((long long)SB * NR == 0) return DINER_OK;
    if (!rays || !z || !rgb || !depth) return fail(DINER_E_INVALID, "NULL pointer argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const long long l0 = g_launches;
    rc = do_composite(c, rays, z, SB, NR, K, white_bkgd, mode, rgb, depth, weights, (cudaStream_t)stream);
    c->launches += g_launches - l0;
    return rc;
}

extern "C" int diner_render(diner_ctx* c, const float* rays, int SB, int NR, int K, int C, int G, int white_bkgd,
                            int mode, const diner_noise* noise, float* rgb, float* depth, float* weights,
                            float* z, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_render_args(c, SB, NR, K, C, G))) return rc;
    if ((long long)SB * NR == 0) return DINER_OK;
    if (!rays || !rgb || !depth) return fail(DINER_E_INVALID, "NULL pointer argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    float* zz = z;
    if (!zz) {
        CUDA_TRY(c->zbuf.reserve((size_t)SB * NR * K * sizeof(float)));
        zz = c->zbuf.as<float>();
    }
    const long long l0 = g_launches;
    rc = do_sample(c, rays, SB, NR, K, C, G, noise, zz, nullptr, st);
    if (!rc) rc = do_composite(c, rays, zz, SB, NR, K, white_bkgd, mode, rgb, depth, weights, st);
    c->launches += g_launches - l0;
    return rc;
}

extern "C" int diner_render_rgbd(diner_ctx* c, const float* rays, int SB, int NR, int K, int C, int G, int white_bkgd,
                                 int mode, const diner_noise* noise, float* rgbd, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_render_args(c, SB, NR, K, C, G))) return rc;
    if ((long long)SB * NR == 0) return DINER_OK;
    if (!rays || !rgbd) return fail(DINER_E_INVALID, "NULL pointer argument");
    if (((uintptr_t)rgbd & 15) != 0) return fail(DINER_E_INVALID, "rgbd must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(c->zbuf.reserve((size_t)SB * NR * K * sizeof(float)));
    const long long l0 = g_launches;
    rc = do_sample(c, rays, SB, NR, K, C, G, noise, c->zbuf.as<float>(), nullptr, st);
    if (!rc) rc = do_composite(c, rays, c->zbuf.as<float>(), SB, NR, K, white_bkgd, mode, nullptr, nullptr, nullptr, st, rgbd);
    c->launches += g_launches - l0;
    return rc;
}

extern "C" long long diner_mlp_param_count(diner_ctx* c) {
    return (c && c->has_mlp) ? (long long)backward_param_count(c->mlp) : 0;
}

// Training-step backward (fp32 CUDA cores; checked against the reference's autograd gradients by the GPU tests): gradients of sum(g_rgb . rgb) + sum(g_depth . depth) through
// composite -> PixelNeRF.forward -> ResnetFC for given sample depths z.
extern "C" int diner_render_backward(diner_ctx* c, const float* rays, const float* z, int SB, int NR, int K, int white_bkgd,
                                     const float* g_rgb, const float* g_depth, float* grad_params, float* d_latent, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_render_args(c, SB, NR, K, -1, -1))) return rc;
    if ((long long)SB * NR == 0) return DINER_OK;
    if (!rays || !z || !g_rgb || !grad_params) return fail(DINER_E_INVALID, "NULL pointer argument");
    if (c->softplus_beta > 0.0f) return fail(DINER_E_UNSUPPORTED, "the backward pass implements ReLU activations only (softplus_beta=%g)", (double)c->softplus_beta);
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long l0 = g_launches;
    const long long n_rays = (long long)SB * NR, n = n_rays * K;
    // forward for the per-sample outputs the compositing derivative needs: the fused tcgen05 kernel when the backward runs on
    // the tensor cores (same arithmetic as the training forward), else fp32 CUDA cores
    CUDA_TRY(c->netbuf.reserve((size_t)n * 4 * sizeof(float)));
    QueryArgs q{};
    q.SB = SB; q.n_per_sb = (long long)NR * K; q.rays = rays; q.z = z; q.K = K;
    q.out = c->netbuf.as<float>();
    const bool tcb = c->backward_tc && c->tc.ready && c->mlp.d_hidden == 512 && c->mlp.d_latent == 512 && c->softplus_beta == 0.0f &&
                     c->scene.NV <= 32;
    rc = run_query(c, q, tcb ? DINER_MODE_PARITY : DINER_MODE_FP32, st);
    if (rc) return rc;
    CUDA_TRY(c->dpre.reserve((size_t)n * 4 * sizeof(float)));
    CUDA_TRY(launch_composite_backward(rays, z, q.out, n_rays, K, white_bkgd, g_rgb, g_depth, c->dpre.as<float>(), st));
    const long long chunk = 32768;
    CUDA_TRY(c->bwd_ws.reserve(backward_workspace_bytes(c->mlp, c->scene, chunk)));
    if (tcb) CUDA_TRY(c->bwd_tc.reserve(backward_tc_workspace_bytes(c->mlp, c->scene, chunk)));
    cudaError_t e = backward_simt(c->scene, c->mlp, q, c->dpre.as<float>(), grad_params, d_latent, c->bwd_ws.as<float>(), chunk, st,
                                  tcb ? &c->tc : nullptr, tcb ? c->bwd_tc.as<uint8_t>() : nullptr, c->num_sms);
    c->launches += g_launches - l0;
    if (e == cudaErrorNotSupported) return fail(DINER_E_UNSUPPORTED, "backward needs >= 1 block before and after combine_layer");
    if (e != cudaSuccess) return fail(DINER_E_CUDA, "backward: %s (watchdog code %d)", cudaGetErrorString(e), c->tc.err_flag ? *c->tc.err_flag : -1);
    return DINER_OK;
}

extern "C" int diner_depth2normal(diner_ctx* c, const float* depths, const float* intrinsics, int N, int H, int W, float* normals,
                                  void* stream) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    if (N < 1 || H < 1 || W < 1) return fail(DINER_E_INVALID, "bad N=%d H=%d W=%d", N, H, W);
    if (!depths || !intrinsics || !normals) return fail(DINER_E_INVALID, "NULL pointer argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(launch_depth2normal(depths, intrinsics, N, H, W, normals, c->num_sms, (cudaStream_t)stream));
    g_launches++; c->launches++;
    return DINER_OK;
}

extern "C" int diner_gen_rays(diner_ctx* c, const float* target_extrinsics, const float* target_intrinsics, int SB, int H, int W,
                              float z_near, float z_far, float* rays, void* stream) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    if (SB < 1 || H < 1 || W < 1) return fail(DINER_E_INVALID, "bad SB=%d H=%d W=%d", SB, H, W);
    if (!target_extrinsics || !target_intrinsics || !rays) return fail(DINER_E_INVALID, "NULL pointer argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(launch_gen_rays(target_extrinsics, target_intrinsics, SB, H, W, z_near, z_far, rays, c->num_sms, (cudaStream_t)stream));
    g_launches++; c->launches++;
    return DINER_OK;
}

extern "C" int diner_render_image(diner_ctx* c, const float* target_extrinsics, const float* target_intrinsics, int SB, int H, int W,
                                  float z_near, float z_far, int K, int C, int G, int white_bkgd, int mode,
                                  const diner_noise* noise, float* rgb, float* depth, void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (H < 1 || W < 1 || (long long)H * W > 0x7fffffffLL) return fail(DINER_E_INVALID, "bad image size %dx%d", W, H);
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(c->rays_img.reserve((size_t)SB * H * W * 8 * sizeof(float)));
    rc = diner_gen_rays(c, target_extrinsics, target_intrinsics, SB, H, W, z_near, z_far, c->rays_img.as<float>(), stream);
    if (rc) return rc;
    const int keep = c->tc.ray_image_w;
    c->tc.ray_image_w = W;               // the library generated the rays itself: row-major H x W per scene -> 2-D tile order in the MLP launch
    rc = diner_render(c, c->rays_img.as<float>(), SB, H * W, K, C, G, white_bkgd, mode, noise, rgb, depth, nullptr, nullptr, stream);
    c->tc.ray_image_w = keep;
    return rc;
}

extern "C" int diner_render_host(diner_ctx* c, const float* rays_host, int SB, int NR, int K, int C, int G,
                                 int white_bkgd, int mode, uint64_t seed, float* rgb_host, float* depth_host,
                                 void* stream) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_render_args(c, SB, NR, K, C, G))) return rc;
    if (!rays_host || !rgb_host || !depth_host) return fail(DINER_E_INVALID, "NULL pointer argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)SB * NR;
    CUDA_TRY(c->rays_dev.reserve(n * 8 * sizeof(float)));
    CUDA_TRY(c->out_dev.reserve(n * 4 * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(c->rays_dev.p, rays_host, n * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
    float* rgb = c->out_dev.as<float>();
    float* dep = rgb + n * 3;
    diner_noise nz{nullptr, nullptr, nullptr, seed, 0};
    rc = diner_render(c, c->rays_dev.as<float>(), SB, NR, K, C, G, white_bkgd, mode, &nz, rgb, dep, nullptr, nullptr, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(rgb_host, rgb, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(depth_host, dep, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return DINER_OK;
}

extern "C" int diner_set_option(diner_ctx* c, const char* key, long long value) {
    if (!c || !key) return fail(DINER_E_INVALID, "NULL ctx / key");
    if (!strcmp(key, "backward_tc")) {
        if (value != 0 && value != 1) return fail(DINER_E_INVALID, "backward_tc must be 0 or 1");
        c->backward_tc = (int)value;
    } else if (!strcmp(key, "fused")) {
        if (value != 0 && value != 1) return fail(DINER_E_INVALID, "fused must be 0 or 1");
        c->tc.fused = (int)value;
    } else if (!strcmp(key, "ray_image_width")) {
        if (value < 0 || value > (1 << 20)) return fail(DINER_E_INVALID, "ray_image_width out of range");
        c->tc.ray_image_w = (int)value;
    } else if (!strcmp(key, "post_tiles")) {
        if (value < 1 || value > 8) return fail(DINER_E_INVALID, "post_tiles must be in [1,8]");
        c->tc.post_tiles = (int)value;
    } else if (!strcmp(key, "tail_kb")) {
        if (value < 0 || value > 4) return fail(DINER_E_INVALID, "tail_kb must be in [0,4]");
        c->tc.tail_kb = (int)value;
    } else if (!strcmp(key, "warm_rounds")) {
        c->tc.warm_rounds = value != 0;
    } else if (!strcmp(key, "early_lin")) {
        c->tc.early_lin = value != 0;
    } else if (!strcmp(key, "early_split")) {
        if (value < 0 || value > 7) return fail(DINER_E_INVALID, "early_split must be in [0,7]");
        c->tc.early_split = (int)value;
    } else if (!strcmp(key, "latent_layout")) {
        if (value < 0 || value > 2) return fail(DINER_E_INVALID, "latent_layout must be 0 (NCHW), 1 (NHWC, copied) or 2 (NHWC, borrowed)");
        c->latent_layout = (int)value;
    } else if (!strcmp(key, "rebuild_maps")) {
        c->tc.zmap_valid = false;        // next query rebuilds the hoisted lin_z maps (bench: times the scene-prepare step)
    } else if (!strcmp(key, "dbg_skip")) {
        c->tc.dbg_skip = (int)value;
    } else if (!strcmp(key, "sub_batch")) {
        if (value < 64) return fail(DINER_E_INVALID, "sub_batch must be >= 64");
        c->tc.sub_batch = value;
    } else {
        return fail(DINER_E_INVALID, "unknown option '%s'", key);
    }
    return DINER_OK;
}

extern "C" int diner_set_float_option(diner_ctx* c, const char* key, double value) {
    if (!c || !key) return fail(DINER_E_INVALID, "NULL ctx / key");
    if (!strcmp(key, "depth_diff_max")) {
        if (!(value > 0.0)) return fail(DINER_E_INVALID, "depth_diff_max must be > 0");
        c->depth_diff_max = (float)value;
    } else if (!strcmp(key, "softplus_beta")) {
        if (!(value >= 0.0)) return fail(DINER_E_INVALID, "softplus_beta must be >= 0 (0 = ReLU)");
        c->softplus_beta = (float)value;
    } else {
        return fail(DINER_E_INVALID, "unknown float option '%s'", key);
    }
    return DINER_OK;
}

extern "C" int diner_debug_sync(diner_ctx* c) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
        return fail(DINER_E_CUDA, "device sync: %s (tcgen05 watchdog code %d: 10 producer/empty, 20 mma/operand, 30-31 mma/weights, 4x-5x worker/accumulator, 90 smem misaligned)",
                    cudaGetErrorString(e), c->tc.err_flag ? *c->tc.err_flag : -1);
    return DINER_OK;
}

extern "C" long long diner_launch_count(diner_ctx* c) { return c ? c->launches : 0; }
extern "C" int diner_set_timing(diner_ctx* c, int enabled) {
    if (!c) return fail(DINER_E_INVALID, "ctx is NULL");
    c->timing = enabled;
    c->tc.timing = enabled != 0;
    return DINER_OK;
}
extern "C" float diner_last_mlp_ms(diner_ctx* c) { return c ? c->last_mlp_ms : 0.f; }
extern "C" float diner_last_stage_ms(diner_ctx* c, int stage) {
    if (!c) return 0.f;
    switch (stage) {
        case 0: return c->last_sampler_ms;
        case 1: return c->tc.ms_pre;
        case 2: return c->tc.ms_post;
        case 3: return c->last_composite_ms;
        case 4: return c->tc.ms_zmap;
        default: return c->last_mlp_ms;
    }
}
