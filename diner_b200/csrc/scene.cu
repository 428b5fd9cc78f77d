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
// Op order follows the reference's torch-CPU sequence (probed with a numpy emulation): (screen - c) / focal,
// / sqrt((px^2 + py^2) + 1), then the (B,3,3)@(B,3,N) matmul as (r0*x + r2*z) + r1*y without contraction; origin = (-R^T) t.
// Bit-exact except where torch-CPU's vectorised sqrt is not correctly rounded (~2 % of the pixels, 1 ulp).
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
            d[k] = __fadd_rn(__fadd_rn(__fmul_rn(r0, dx), __fmul_rn(r2, dz)), __fmul_rn(r1, dy));
            o[k] = __fadd_rn(__fadd_rn(__fmul_rn(-r0, __ldg(E + 3)), __fmul_rn(-r2, __ldg(E + 11))), __fmul_rn(-r1, __ldg(E + 7)));
        }
        float4* dst = (float4*)(rays + i * 8);
        dst[0] = make_float4(o[0], o[1], o[2], d[0]);
        dst[1] = make_float4(d[1], d[2], z_near, z_far);
    }
}

// depth2normal (src/util/depth2normal.py:6-87): back-project with the intrinsics, central differences on the replicate-padded
// point map, cross product, normalise; a pixel next to a hole (neighbour point with x == 0) copies the un-cleaned normal of the
// neighbour on the opposite side; pixels without depth get 0.  Arithmetic pinned to torch-CPU by emulation: the cross product
// components are fma(a1, b2, -(a2*b1)), the norm is sqrt(fma(c2,c2, fma(c1,c1, c0*c0))), then a true division.
struct D2NCam { float fx, fy, cx, cy; };
__device__ __forceinline__ void d2n_point(const float* __restrict__ d, const D2NCam& k, int H, int W, int y, int x, float& px, float& py, float& pz) {
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    const float z = __ldg(d + (size_t)y * W + x);
    px = __fmul_rn(__fdiv_rn(__fsub_rn((float)x + 0.5f, k.cx), k.fx), z);
    py = __fmul_rn(__fdiv_rn(__fsub_rn((float)y + 0.5f, k.cy), k.fy), z);
    pz = z;
}
__device__ __forceinline__ void d2n_raw(const float* __restrict__ d, const D2NCam& k, int H, int W, int y, int x, float (&n)[3], int& dy, int& dx) {
    float dn[3], up[3], rt[3], lf[3];
    d2n_point(d, k, H, W, y + 1, x, dn[0], dn[1], dn[2]);
    d2n_point(d, k, H, W, y - 1, x, up[0], up[1], up[2]);
    d2n_point(d, k, H, W, y, x + 1, rt[0], rt[1], rt[2]);
    d2n_point(d, k, H, W, y, x - 1, lf[0], lf[1], lf[2]);
    float a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = __fsub_rn(dn[i], up[i]); b[i] = __fsub_rn(rt[i], lf[i]); }
    const float c0 = fmaf(a[1], b[2], -__fmul_rn(a[2], b[1]));
    const float c1 = fmaf(a[2], b[0], -__fmul_rn(a[0], b[2]));
    const float c2 = fmaf(a[0], b[1], -__fmul_rn(a[1], b[0]));
    const float nn = __fsqrt_rn(fmaf(c2, c2, fmaf(c1, c1, __fmul_rn(c0, c0))));
    n[0] = __fdiv_rn(c0, nn); n[1] = __fdiv_rn(c1, nn); n[2] = __fdiv_rn(c2, nn);
    dy = (up[0] == 0.0f ? 1 : 0) - (dn[0] == 0.0f ? 1 : 0);
    dx = (lf[0] == 0.0f ? 1 : 0) - (rt[0] == 0.0f ? 1 : 0);
}
__global__ void depth2normal_kernel(const float* __restrict__ depth, const float* __restrict__ intr, int N, int H, int W,
                                    float* __restrict__ normals) {
    const long long total = (long long)N * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / ((long long)H * W));
        const int pix = (int)(i % ((long long)H * W));
        const int y = pix / W, x = pix % W;
        const float* d = depth + (size_t)n * H * W;
        const float* Kp = intr + (size_t)n * 9;
        const D2NCam k{__ldg(Kp + 0), __ldg(Kp + 4), __ldg(Kp + 2), __ldg(Kp + 5)};
        float nr[3] = {0.0f, 0.0f, 0.0f};
        if (__ldg(d + pix) != 0.0f) {
            int dy, dx;
            d2n_raw(d, k, H, W, y, x, nr, dy, dx);
            if (dy != 0 || dx != 0) {
                int y2 = y + dy, x2 = x + dx, dy2, dx2;
                y2 = y2 < 0 ? 0 : (y2 > H - 1 ? H - 1 : y2);
                x2 = x2 < 0 ? 0 : (x2 > W - 1 ? W - 1 : x2);
                d2n_raw(d, k, H, W, y2, x2, nr, dy2, dx2);
            }
        }
        float* o = normals + (size_t)n * 3 * H * W + pix;
        o[0] = nr[0]; o[(size_t)H * W] = nr[1]; o[2 * (size_t)H * W] = nr[2];
    }
}
}  // namespace

cudaError_t launch_depth2normal(const float* depth, const float* intr, int N, int H, int W, float* normals, int num_sms, cudaStream_t st) {
    depth2normal_kernel<<<num_sms * 8, 256, 0, st>>>(depth, intr, N, H, W, normals);
    return cudaGetLastError();
}

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
