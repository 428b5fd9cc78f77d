// Alpha compositing along rays with a warp product-scan: one warp per ray.
// Reference: NeRFRendererDGS.composite, nerf_renderer.py:299-301 (deltas), :344-360 (alpha, T, weights,
// rgb, depth, white background).  net_out holds [sigmoid rgb, relu sigma] per sample.
#include "common.cuh"
#include "diner_internal.h"

namespace {

__global__ void __launch_bounds__(128)
composite_kernel(const float* __restrict__ rays, const float* __restrict__ z, const float* __restrict__ net,
                 long long n_rays, int K, int white, float* __restrict__ rgb, float* __restrict__ depth,
                 float* __restrict__ weights, float4* __restrict__ rgbd) {
    const int lane = threadIdx.x & 31;
    const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= n_rays) return;
    const float far = rays[ray * 8 + 7];
    const float* zr = z + ray * K;
    float T_run = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        float alpha = 0.0f, zk = 0.0f;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
            zk = zr[k];
            const float znext = (k + 1 < K) ? zr[k + 1] : far;          // last delta = far - z_last (:300)
            const float delta = __fsub_rn(znext, zk);
            c = ((const float4*)net)[ray * K + k];
            alpha = __fsub_rn(1.0f, expf(__fmul_rn(-delta, fmaxf(c.w, 0.0f))));   // (:344)
        }
        float f = (k < K) ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;     // (:347-348)
        float inc = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc *= t;
        }
        float excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 1.0f;
        const float w = alpha * (T_run * excl);                                   // (:350-351)
        T_run *= __shfl_sync(0xffffffffu, inc, 31);
        if (k < K && weights) weights[ray * K + k] = w;
        ar += w * c.x; ag += w * c.y; ab += w * c.z; ad += w * zk; aw += w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, o);
        ag += __shfl_xor_sync(0xffffffffu, ag, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
        ad += __shfl_xor_sync(0xffffffffu, ad, o);
        aw += __shfl_xor_sync(0xffffffffu, aw, o);
    }
    if (lane == 0) {
        if (white) { ar = (ar + 1.0f) - aw; ag = (ag + 1.0f) - aw; ab = (ab + 1.0f) - aw; }   // (:357-360)
        if (rgbd) {                       // packed rgb|depth: the layout the image all-gather sends (diner_render_rgbd)
            rgbd[ray] = make_float4(ar, ag, ab, ad);
        } else {
            rgb[ray * 3 + 0] = ar; rgb[ray * 3 + 1] = ag; rgb[ray * 3 + 2] = ab;
            depth[ray] = ad;
        }
    }
}

}  // namespace

cudaError_t launch_composite(const float* rays, const float* z, const float* net_out, long long n_rays,
                             int K, int white_bkgd, float* rgb, float* depth, float* weights,
                             cudaStream_t st, float* rgbd) {
    if (n_rays <= 0) return cudaSuccess;
    const long long threads = n_rays * 32;
    composite_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(rays, z, net_out, n_rays, K, white_bkgd,
                                                                        rgb, depth, weights, (float4*)rgbd);
    return cudaGetLastError();
}
