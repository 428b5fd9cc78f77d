// Once-per-scene re-layouts: the reference keeps the latent as NCHW fp32 (image_encoder.py:290-291),
// which makes a 512-channel bilinear tap touch 512 scattered sectors; NHWC makes it one contiguous 2 KiB read.
#include "common.cuh"
#include "diner_internal.h"
#include <math.h>

namespace {
// (N, C, HW) -> (N, HW, C), 32x32 tile transpose through shared memory
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* s = src + (size_t)n * C * HW;
    float* d = dst + (size_t)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? s[(size_t)c * HW + p] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) d[(size_t)p * C + c] = tile[threadIdx.x][i];
    }
}

// Target-view rays [o3, d3, near, far] for every pixel centre (src/util/cam_geometry.py:5-48), one thread per pixel.
// Op order follows the reference's torch sequence: (screen - c) / focal, / sqrt((px^2 + py^2) + 1), then the K=3 matmul form
// fma(r2,z, fma(r1,y, r0*x)) with R^T (see common.cuh dot3_rm); origin = (-R^T) t.
__global__ void gen_rays_kernel(const float* __restrict__ ext, const float* __restrict__ intr, int SB, int H, int W, float z_near,
                                float z_far, float* __restrict__ rays) {
    const long long n = (long long)SB * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int sb = (int)(i / ((long long)H * W));
        const int pix = (int)(i % ((long long)H * W));
        const int y = pix / W, x = pix % W;
        const float* E = ext + (size_t)sb * 16;
        const float* Kp = intr + (size_t)sb * 9;
        const float px = __fdiv_rn(__fsub_rn((float)x + 0.5f, __ldg(Kp + 2)), __ldg(Kp + 0));
        const float py = __fdiv_rn(__fsub_rn((float)y + 0.5f, __ldg(Kp + 5)), __ldg(Kp + 4));
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), 1.0f));
        const float dx = __fdiv_rn(px, nrm), dy = __fdiv_rn(py, nrm), dz = __fdiv_rn(1.0f, nrm);
        float o[3], d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float r0 = __ldg(E + k), r1 = __ldg(E + 4 + k), r2 = __ldg(E + 8 + k);      // row k of R^T = column k of R
            d[k] = fmaf(r2, dz, fmaf(r1, dy, __fmul_rn(r0, dx)));
            o[k] = fmaf(-r2, __ldg(E + 11), fmaf(-r1, __ldg(E + 7), __fmul_rn(-r0, __ldg(E + 3))));
        }
        float4* dst = (float4*)(rays + i * 8);
        dst[0] = make_float4(o[0], o[1], o[2], d[0]);
        dst[1] = make_float4(d[1], d[2], z_near, z_far);
    }
}
}  // namespace

cudaError_t launch_gen_rays(const float* ext, const float* intr, int SB, int H, int W, float z_near, float z_far, float* rays,
                            int num_sms, cudaStream_t st) {
    gen_rays_kernel<<<num_sms * 8, 256, 0, st>>>(ext, intr, SB, H, W, z_near, z_far, rays);
    return cudaGetLastError();
}

cudaError_t launch_nchw_to_nhwc(const float* src, float* dst, int N, int C, int HW, cudaStream_t st) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, st>>>(src, dst, C, HW);
    return cudaGetLastError();
}

// 2^(ring/12): torch evaluates exp(ring / 12 * log(2)) in fp32 (torch_helpers.py:120)
cudaError_t upload_std_ring_gain() {
    float h[STD_PAD];
    for (int i = 0; i < STD_PAD; ++i) {
        const float e = (float)i / 12.0f;
        const float a = e * 0.6931471805599453f;
        h[i] = (float)exp((double)a);
    }
    return cudaMemcpyToSymbol(c_std_ring_gain, h, sizeof(h));
}
