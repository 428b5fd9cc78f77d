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
}  // namespace

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
