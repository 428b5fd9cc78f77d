// Shared device helpers for the DINER render path (sm_100a).
//
// Geometry / lookup arithmetic is written with explicit single-rounding intrinsics (__fmul_rn,
// __fadd_rn, fmaf ...) so that nvcc cannot contract it differently from the reference's torch-CPU
// op sequence: the nearest-neighbour lookups downstream are discontinuous, so a 1-ulp difference in
// uv would flip pixels.  Each helper cites the reference line it mirrors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DINER_MAX_FREQS 8

struct SceneDev {
    int SB, NV, L, Hl, Wl, H, W;
    const float* latent;   // NHWC fp32 (SB*NV, Hl, Wl, L)   (re-laid out from the reference's NCHW once per scene)
    const float* depth;    // (SB*NV, H, W)
    const float* dstd;     // (SB*NV, H, W)
    const float* normal;   // (SB*NV, 3, H, W)
    const float* poses;    // (SB*NV, 4, 4) world->cam, rows 0..2 used
    const float* focal;    // (SB*NV, 2)
    const float* cxy;      // (SB*NV, 2)
    float imgW, imgH;      // model.image_shape = [W, H]            (pixelnerf.py:50-51)
    float lat_sx, lat_sy;  // (Wl - 2p)/Wl, (Hl - 2p)/Hl           (image_encoder.py:113-114)
    float std_sx, std_sy;  // W/(W+2*100), H/(H+2*100)             (torch_helpers.py:157-158)
    int num_freqs;
    float freqs[DINER_MAX_FREQS];   // freq_factor * 2^i            (positional_encoding.py:18)
};

#define STD_PAD 100
__constant__ float c_std_ring_gain[STD_PAD];  // 2^(ring/12), ring = 0..99  (torch_helpers.py:110-120)

// ---------------------------------------------------------------------------------------------
// x_c = R x + t : torch.matmul on CPU evaluates the K=3 dot as fma(r2,z, fma(r1,y, r0*x))
// (probed, see DESIGN.md "arithmetic pinning"), then the translation is a separate add.
// (pixelnerf.py:91-93, nerf_renderer.py:99-101)
__device__ __forceinline__ float dot3_rm(const float* r, float x, float y, float z) {
    return fmaf(r[2], z, fmaf(r[1], y, __fmul_rn(r[0], x)));
}
__device__ __forceinline__ void world_to_cam(const float* P, float x, float y, float z,
                                             float& cx, float& cy, float& cz) {
    cx = __fadd_rn(dot3_rm(P + 0, x, y, z), P[3]);
    cy = __fadd_rn(dot3_rm(P + 4, x, y, z), P[7]);
    cz = __fadd_rn(dot3_rm(P + 8, x, y, z), P[11]);
}
__device__ __forceinline__ void rotate_to_cam(const float* P, float x, float y, float z,
                                              float& cx, float& cy, float& cz) {
    cx = dot3_rm(P + 0, x, y, z);
    cy = dot3_rm(P + 4, x, y, z);
    cz = dot3_rm(P + 8, x, y, z);
}

// uv = ((xy / z) * f + c) / image_shape * 2 - 1     (pixelnerf.py:105-108, nerf_renderer.py:107-110)
__device__ __forceinline__ float project_axis(float a, float z, float f, float c, float size) {
    float u = __fdiv_rn(a, z);
    u = __fmul_rn(u, f);
    u = __fadd_rn(u, c);
    u = __fdiv_rn(u, size);
    u = __fmul_rn(u, 2.0f);
    return __fsub_rn(u, 1.0f);
}

// grid_sample(align_corners=False) un-normalisation on CPU: (u + 1) * (size/2) - 0.5
__device__ __forceinline__ float unnormalize(float u, float size) {
    return __fsub_rn(__fmul_rn(__fadd_rn(u, 1.0f), __fmul_rn(size, 0.5f)), 0.5f);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// nearest + border: clip in float, then round-half-even  (F.grid_sample nearest/border)
__device__ __forceinline__ int nearest_border(float u, int size) {
    float v = unnormalize(u, (float)size);
    v = fminf(fmaxf(v, 0.0f), (float)(size - 1));
    return clampi(__float2int_rn(v), 0, size - 1);
}
// nearest + zeros: round, then in-range test
__device__ __forceinline__ bool nearest_zeros(float u, int size, int& idx) {
    float v = unnormalize(u, (float)size);
    float r = rintf(v);
    idx = clampi(__float2int_rn(v), 0, size - 1);
    return (r >= 0.0f) && (r <= (float)(size - 1));
}

// depth lookup (image_encoder.py:148-170)
__device__ __forceinline__ float lookup_depth(const SceneDev& s, int sv, float u, float v) {
    int x = nearest_border(u, s.W), y = nearest_border(v, s.H);
    return __ldg(s.depth + ((size_t)sv * s.H + y) * s.W + x);
}
// depth-std lookup through the analytic exponential padding (image_encoder.py:172-199,
// torch_helpers.py:99-159): map padded by 100 px, ring r scaled by 2^(r/12), zeros outside.
__device__ __forceinline__ float lookup_std(const SceneDev& s, int sv, float u, float v) {
    int Wp = s.W + 2 * STD_PAD, Hp = s.H + 2 * STD_PAD, xi, yi;
    bool okx = nearest_zeros(__fmul_rn(u, s.std_sx), Wp, xi);
    bool oky = nearest_zeros(__fmul_rn(v, s.std_sy), Hp, yi);
    if (!(okx && oky)) return 0.0f;
    int rx = xi < STD_PAD ? (STD_PAD - 1 - xi) : (xi >= s.W + STD_PAD ? xi - (s.W + STD_PAD) : 0);
    int ry = yi < STD_PAD ? (STD_PAD - 1 - yi) : (yi >= s.H + STD_PAD ? yi - (s.H + STD_PAD) : 0);
    int ring = rx > ry ? rx : ry;
    int x = clampi(xi - STD_PAD, 0, s.W - 1), y = clampi(yi - STD_PAD, 0, s.H - 1);
    float base = __ldg(s.dstd + ((size_t)sv * s.H + y) * s.W + x);
    return __fmul_rn(base, c_std_ring_gain[ring]);
}
// normal lookup (image_encoder.py:201-223): zeros outside
__device__ __forceinline__ void lookup_normal(const SceneDev& s, int sv, float u, float v,
                                              float& nx, float& ny, float& nz) {
    int xi, yi;
    bool okx = nearest_zeros(u, s.W, xi), oky = nearest_zeros(v, s.H, yi);
    if (!(okx && oky)) { nx = ny = nz = 0.0f; return; }
    const float* p = s.normal + (size_t)sv * 3 * s.H * s.W + (size_t)yi * s.W + xi;
    nx = __ldg(p); ny = __ldg(p + (size_t)s.H * s.W); nz = __ldg(p + 2 * (size_t)s.H * s.W);
}

// Bilinear tap set for the latent gather (image_encoder.py:97-146: uv rescaled for the feature
// padding, bilinear, border, align_corners=False).
struct LatTaps {
    int o00, o01, o10, o11;   // pixel offsets (in pixels, multiply by L for the NHWC element offset)
    float w00, w01, w10, w11;
};
__device__ __forceinline__ LatTaps latent_taps(const SceneDev& s, float u, float v) {
    float x = unnormalize(__fmul_rn(u, s.lat_sx), (float)s.Wl);
    float y = unnormalize(__fmul_rn(v, s.lat_sy), (float)s.Hl);
    x = fminf(fmaxf(x, 0.0f), (float)(s.Wl - 1));
    y = fminf(fmaxf(y, 0.0f), (float)(s.Hl - 1));
    if (!(x == x)) x = 0.0f;
    if (!(y == y)) y = 0.0f;
    float xf = floorf(x), yf = floorf(y);
    int x0 = (int)xf, y0 = (int)yf;
    int x1 = x0 + 1 < s.Wl ? x0 + 1 : s.Wl - 1, y1 = y0 + 1 < s.Hl ? y0 + 1 : s.Hl - 1;
    float ex = (xf + 1.0f) - x, wx = x - xf, ey = (yf + 1.0f) - y, wy = y - yf;
    LatTaps t;
    t.o00 = y0 * s.Wl + x0; t.o01 = y0 * s.Wl + x1; t.o10 = y1 * s.Wl + x0; t.o11 = y1 * s.Wl + x1;
    t.w00 = ex * ey; t.w01 = wx * ey; t.w10 = ex * wy; t.w11 = wx * wy;
    return t;
}

// PositionalEncoding argument: torch.addcmul(phase, x, freq) is one fma on CPU
// (positional_encoding.py:46); cos is sin(x + fp32(pi/2)).
#define DINER_HALF_PI_F 1.5707963705062866f
__device__ __forceinline__ float pe_sin(float x, float f) { return sinf(__fmul_rn(x, f)); }
__device__ __forceinline__ float pe_cos(float x, float f) { return sinf(fmaf(x, f, DINER_HALF_PI_F)); }

// Counter-based fallback noise (used only when the caller passes no explicit noise arrays).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint32_t stream, uint64_t idx) {
    uint64_t b = mix64(idx ^ mix64(seed * 0x1000003ull + stream));
    return (float)(b >> 40) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float rng_normal(uint64_t seed, uint32_t stream, uint64_t idx) {
    float u1 = rng_uniform(seed, stream, 2 * idx), u2 = rng_uniform(seed, stream, 2 * idx + 1);
    u1 = fmaxf(u1, 5.9604645e-8f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
