// fp32 CUDA-core path for the per-sample network query (PixelNeRF.forward + ResnetFC.forward).
//
// This is the *exact-arithmetic* mode of the library (fp32 FMA accumulation like the reference's
// SGEMM): it anchors parity for the tensor-core kernels in mlp_tc.cu and serves MLP shapes the
// tcgen05 path does not cover.  Layer by layer with activations in a global workspace -- simple,
// not fast; the fused tcgen05 path is the product hot path.
//
// Reference: pixelnerf.py:55-145 (feature assembly), resnetfc.py:61-69,129-159 (network).
#include "common.cuh"
#include "diner_internal.h"

namespace {

// One warp per (sample, view) row: builds the d_in network inputs and gathers the latent.
// Row order is [sample][view] (views of a sample adjacent).
// xin columns (pixelnerf.py:96,102,116,128):
//   [x_c y_c z_c | sin/cos(f_j * xyz) j-major | dir_c xyz | dd | sin/cos(f_j * dd)]
__device__ __forceinline__ float feature_elem(int e, int F, const float* fr, float xc, float yc, float zc,
                                              float dxc, float dyc, float dzc, float dd) {
    const int npe = 2 * F * 3;
    if (e < 3) return e == 0 ? xc : (e == 1 ? yc : zc);
    if (e < 3 + npe) {
        const int q = e - 3, j = q / 3, i = q % 3;
        const float x = i == 0 ? xc : (i == 1 ? yc : zc);
        return (j & 1) ? pe_cos(x, fr[j >> 1]) : pe_sin(x, fr[j >> 1]);
    }
    if (e < 6 + npe) { const int i = e - 3 - npe; return i == 0 ? dxc : (i == 1 ? dyc : dzc); }
    if (e == 6 + npe) return dd;
    const int j = e - 7 - npe;
    return (j & 1) ? pe_cos(dd, fr[j >> 1]) : pe_sin(dd, fr[j >> 1]);
}

__global__ void __launch_bounds__(256)
features_kernel(SceneDev s, QueryArgs q, long long s_begin, long long n_samples, int d_in, int ld_in,
                float* __restrict__ xin, float* __restrict__ zlat) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_rows = n_samples * s.NV;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = warp_global; row < n_rows; row += n_warps) {
        const long long smp = s_begin + row / s.NV;
        const int v = (int)(row % s.NV);
        const int sb = (int)(smp / q.n_per_sb);
        float px, py, pz, dx, dy, dz;
        if (q.xyz) {
            px = q.xyz[smp * 3]; py = q.xyz[smp * 3 + 1]; pz = q.xyz[smp * 3 + 2];
            dx = q.viewdirs[smp * 3]; dy = q.viewdirs[smp * 3 + 1]; dz = q.viewdirs[smp * 3 + 2];
        } else {
            const long long ray = smp / q.K;
            const float* r = q.rays + ray * 8;
            const float z = q.z[smp];
            dx = r[3]; dy = r[4]; dz = r[5];
            px = __fadd_rn(r[0], __fmul_rn(z, dx));        // nerf_renderer.py:304
            py = __fadd_rn(r[1], __fmul_rn(z, dy));
            pz = __fadd_rn(r[2], __fmul_rn(z, dz));
        }
        const int sv = sb * s.NV + v;
        const float* P = s.poses + (size_t)sv * 16;
        float p[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) p[k] = __ldg(P + k);
        float xc, yc, zc, dxc, dyc, dzc;
        world_to_cam(p, px, py, pz, xc, yc, zc);
        rotate_to_cam(p, dx, dy, dz, dxc, dyc, dzc);
        const float u = project_axis(xc, zc, __ldg(s.focal + sv * 2), __ldg(s.cxy + sv * 2), s.imgW);
        const float w = project_axis(yc, zc, __ldg(s.focal + sv * 2 + 1), __ldg(s.cxy + sv * 2 + 1), s.imgH);
        const float dd = __fsub_rn(lookup_depth(s, sv, u, w), zc);     // pixelnerf.py:114-115
        for (int e = lane; e < ld_in; e += 32)
            xin[row * ld_in + e] = e < d_in ? feature_elem(e, s.num_freqs, s.freqs, xc, yc, zc, dxc, dyc, dzc, dd) : 0.0f;
        const LatTaps t = latent_taps(s, u, w);
        const float* base = s.latent + (size_t)sv * s.Hl * s.Wl * s.L;
        for (int c = lane * 4; c < s.L; c += 128) {
            const float4 a = __ldg((const float4*)(base + (size_t)t.o00 * s.L + c));
            const float4 b = __ldg((const float4*)(base + (size_t)t.o01 * s.L + c));
            const float4 g = __ldg((const float4*)(base + (size_t)t.o10 * s.L + c));
            const float4 h = __ldg((const float4*)(base + (size_t)t.o11 * s.L + c));
            float4 o;
            o.x = a.x * t.w00 + b.x * t.w01 + g.x * t.w10 + h.x * t.w11;
            o.y = a.y * t.w00 + b.y * t.w01 + g.y * t.w10 + h.y * t.w11;
            o.z = a.z * t.w00 + b.z * t.w01 + g.z * t.w10 + h.z * t.w11;
            o.w = a.w * t.w00 + b.w * t.w01 + g.w * t.w10 + h.w * t.w11;
            *(float4*)(zlat + row * s.L + c) = o;
        }
    }
}

// Y[r][o] (+)= sum_k act(X[r][k]) * W[o][k] + b[o]   (64x64 tile per CTA, 4x4 per thread, K step 16)
// nn.Softplus(beta) with torch's threshold of 20 (resnetfc.py:124-125), or ReLU for beta == 0 (:127)
__device__ __forceinline__ float activation(float x, float beta) {
    if (beta > 0.0f) {
        const float bx = beta * x;
        return bx > 20.0f ? x : log1pf(expf(bx)) / beta;
    }
    return fmaxf(x, 0.0f);
}

template <bool RELU_IN, bool ACCUM>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
              const float* __restrict__ bias, float* __restrict__ Y, int ldy, long long rows, int in_dim,
              int out_dim, float beta) {
    __shared__ float Xs[16][68];
    __shared__ float Ws[16][68];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long r0 = (long long)blockIdx.x * 64;
    const int o0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < in_dim; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int rr = i >> 4, kk = i & 15;
            const long long r = r0 + rr;
            float xv = (r < rows && k0 + kk < in_dim) ? X[r * ldx + k0 + kk] : 0.0f;
            if (RELU_IN) xv = activation(xv, beta);
            Xs[kk][rr] = xv;
            const int o = o0 + rr;
            Ws[kk][rr] = (o < out_dim && k0 + kk < in_dim) ? W[(size_t)o * ldw + k0 + kk] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float xr[4], wr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { xr[i] = Xs[kk][ty * 4 + i]; wr[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + ty * 4 + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + tx * 4 + j;
            if (o >= out_dim) continue;
            float v = acc[i][j] + (bias ? bias[o] : 0.0f);
            if (ACCUM) v = Y[r * ldy + o] + v;
            Y[r * ldy + o] = v;
        }
    }
}

// mean over the NV view rows of each sample (resnetfc.py:9-14 via torch.mean: sequential sum / NV)
__global__ void combine_kernel(const float* __restrict__ x, float* __restrict__ xc, long long n_samples,
                               int NV, int Hd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples * Hd) return;
    const long long smp = i / Hd;
    const int h = (int)(i % Hd);
    float acc = x[(smp * NV) * Hd + h];
    for (int v = 1; v < NV; ++v) acc = __fadd_rn(acc, x[(smp * NV + v) * Hd + h]);
    xc[i] = __fdiv_rn(acc, (float)NV);
}

// sigmoid(rgb), relu(sigma)  (pixelnerf.py:139-143)
__global__ void activate_kernel(float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = ((float4*)out)[i];
    v.x = 1.0f / (1.0f + expf(-v.x));
    v.y = 1.0f / (1.0f + expf(-v.y));
    v.z = 1.0f / (1.0f + expf(-v.z));
    v.w = fmaxf(v.w, 0.0f);
    ((float4*)out)[i] = v;
}

template <bool RELU_IN, bool ACCUM>
cudaError_t linear(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy,
                   long long rows, int in_dim, int out_dim, cudaStream_t st, float beta = 0.0f) {
    dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((out_dim + 63) / 64));
    linear_kernel<RELU_IN, ACCUM><<<grid, 256, 0, st>>>(X, ldx, W, ldw, b, Y, ldy, rows, in_dim, out_dim, beta);
    g_launches++;
    return cudaGetLastError();
}

constexpr long long SIMT_CHUNK_SAMPLES = 32768;

}  // namespace

size_t simt_workspace_bytes(const MlpDev& m, long long rows) {
    const int ld_in = (m.d_in + 7) & ~7;
    return (size_t)rows * (ld_in + (size_t)m.d_latent + 3 * (size_t)m.d_hidden) * sizeof(float);
}

#define CK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t query_simt(const SceneDev& s, const MlpDev& m, const QueryArgs& q, SimtWorkspace& ws,
                       cudaStream_t st) {
    const long long total = (long long)q.SB * q.n_per_sb;
    const int Hd = m.d_hidden, ld_in = (m.d_in + 7) & ~7;
    const long long chunk = ws.rows_cap / s.NV < SIMT_CHUNK_SAMPLES ? ws.rows_cap / s.NV : SIMT_CHUNK_SAMPLES;
    if (chunk <= 0) return cudaErrorInvalidValue;
    for (long long s0 = 0; s0 < total; s0 += chunk) {
        const long long ns = total - s0 < chunk ? total - s0 : chunk;
        const long long rows = ns * s.NV;
        const int fgrid = (int)((rows * 32 + 255) / 256 < 148 * 64 ? (rows * 32 + 255) / 256 : 148 * 64);
        features_kernel<<<fgrid, 256, 0, st>>>(s, q, s0, ns, m.d_in, ld_in, ws.xin, ws.zlat);
        g_launches++;
        CK(cudaGetLastError());
        CK((linear<false, false>(ws.xin, ld_in, m.w_in, m.d_in, m.b_in, ws.x, Hd, rows, m.d_in, Hd, st)));
        float* x = ws.x;
        long long r = rows;
        for (int b = 0; b < m.n_blocks; ++b) {
            if (b == m.combine_layer) {
                const long long n = ns * Hd;
                combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws.x, ws.xc, ns, s.NV, Hd);
                g_launches++;
                CK(cudaGetLastError());
                x = ws.xc;
                r = ns;
            }
            if (b < m.combine_layer)
                CK((linear<false, true>(ws.zlat, m.d_latent, m.w_z[b], m.d_latent, m.b_z[b], x, Hd, r, m.d_latent, Hd, st)));
            CK((linear<true, false>(x, Hd, m.w_fc0[b], Hd, m.b_fc0[b], ws.net, Hd, r, Hd, Hd, st, m.beta)));
            CK((linear<true, true>(ws.net, Hd, m.w_fc1[b], Hd, m.b_fc1[b], x, Hd, r, Hd, Hd, st, m.beta)));
        }
        if (m.combine_layer >= m.n_blocks && s.NV > 1) return cudaErrorNotSupported;
        CK((linear<true, false>(x, Hd, m.w_out, Hd, m.b_out, q.out + s0 * 4, 4, r, Hd, m.d_out, st, m.beta)));
        activate_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(q.out + s0 * 4, ns);
        g_launches++;
        CK(cudaGetLastError());
    }
    return cudaSuccess;
}
