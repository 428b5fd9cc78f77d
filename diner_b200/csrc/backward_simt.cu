// Backward pass of the render path on fp32 CUDA cores: the training step of BASELINE config 3 (loss + backward through the
// renderer), checked against the reference's own autograd gradients (tests/test_gpu_parity.py::test_backward_*).
//
// What the reference differentiates (src/models/diner.py:257-266 -> nerf_renderer.py:286-365 -> pixelnerf.py:55-145 ->
// resnetfc.py:129-159): rendered colours w.r.t. the ResnetFC parameters and the latent maps.  The sample depths carry no
// gradient (sample_depthguided is @torch.no_grad, nerf_renderer.py:65; rays / cameras / depth maps are data).
//
// Structure: (1) compositing backward, one thread per ray: d(rgb, depth) -> pre-activation gradients of the per-sample
// network output; (2) per chunk of samples the fp32 forward is recomputed with every block input kept, then the chain is
// walked backwards with the same 64x64-tile GEMM kernel (data gradients use transposed weight copies) and a split-K
// weight-gradient kernel that accumulates with atomics; (3) the gradient of the gathered latent rows is scattered back
// through the bilinear taps.
//
// Two arithmetic paths share this chain: fp32 CUDA cores (the 64x64-tile kernels below; parity anchor, any MLP shape), and --
// for the shipped 512-wide network -- the tcgen05 GEMM of gemm_tc3.cu (fp16 hi/lo split operands, fp32 accumulation: the same
// arithmetic as the forward parity mode) for every 512x512 GEMM: forward recompute, data gradients (packed transposed weights)
// and weight gradients (G^T against packed activations, split over the rows, atomics).  The few narrow GEMMs (lin_in: 55
// inputs, lin_out: 4 outputs) stay on the CUDA cores.
#include "diner_internal.h"
#include "mlp_tc.h"

namespace {

// ---- compositing backward (nerf_renderer.py:299-360) ---------------------------------------------------------------
// alpha_k = 1 - exp(-delta_k sigma_k), t_k = 1 - alpha_k + 1e-10, T_k = prod_{j<k} t_j, w_k = alpha_k T_k,
// rgb = sum w_k c_k (+ 1 - sum w_k), depth = sum w_k z_k.  With G_k = g_rgb.c_k + g_depth z_k - [white] sum(g_rgb):
//   dL/dc_k = w_k g_rgb;  dL/dalpha_k = G_k T_k - (sum_{j>k} G_j w_j) / t_k;  dL/dsigma_k = dL/dalpha_k delta_k (1 - alpha_k).
// net_out holds sigmoid(rgb) and relu(sigma); the outputs are gradients w.r.t. the PRE-activation values.
__global__ void composite_backward_kernel(const float* __restrict__ rays, const float* __restrict__ z,
                                          const float* __restrict__ net_out, long long n_rays, int K, int white,
                                          const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
                                          float* __restrict__ d_pre) {
    const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float far = rays[ray * 8 + 7];
    const float* zr = z + ray * K;
    const float4* o = (const float4*)net_out + ray * K;
    float4* d = (float4*)d_pre + ray * K;
    const float gr = g_rgb[ray * 3], gg = g_rgb[ray * 3 + 1], gb = g_rgb[ray * 3 + 2];
    const float gd = g_depth ? g_depth[ray] : 0.0f;
    const float gw = white ? gr + gg + gb : 0.0f;
    // forward sweep: total of G_k w_k, then a second sweep carries the prefix product and the exclusive suffix sum
    float T = 1.0f, tot = 0.0f;
    for (int k = 0; k < K; ++k) {
        const float4 v = o[k];
        const float delta = (k + 1 < K ? zr[k + 1] : far) - zr[k];
        const float a = 1.0f - expf(-delta * v.w);
        const float G = gr * v.x + gg * v.y + gb * v.z + gd * zr[k] - gw;
        tot += G * a * T;
        T *= 1.0f - a + 1e-10f;
    }
    T = 1.0f;
    float pre = 0.0f;   // sum_{j<=k} G_j w_j
    for (int k = 0; k < K; ++k) {
        const float4 v = o[k];
        const float delta = (k + 1 < K ? zr[k + 1] : far) - zr[k];
        const float a = 1.0f - expf(-delta * v.w);
        const float t = 1.0f - a + 1e-10f;
        const float w = a * T;
        const float G = gr * v.x + gg * v.y + gb * v.z + gd * zr[k] - gw;
        pre += G * w;
        const float dA = G * T - (tot - pre) / t;
        const float dS = v.w > 0.0f ? dA * delta * (1.0f - a) : 0.0f;
        d[k] = make_float4(w * gr * v.x * (1.0f - v.x), w * gg * v.y * (1.0f - v.y), w * gb * v.z * (1.0f - v.z), dS);
        T *= t;
    }
}

// ---- small element-wise pieces ------------------------------------------------------------------------------------------
// g[i] = (act[i] > 0) ? g[i] : 0
__global__ void relu_mask_kernel(float* __restrict__ g, const float* __restrict__ act, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(act[i] > 0.0f)) g[i] = 0.0f;
}
// dst[i] += (act[i] > 0) ? g[i] : 0
__global__ void relu_mask_add_kernel(float* __restrict__ dst, const float* __restrict__ g, const float* __restrict__ act, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && act[i] > 0.0f) dst[i] += g[i];
}
// mean over views backwards: every view row of a sample receives g / NV
__global__ void combine_backward_kernel(const float* __restrict__ gc, float* __restrict__ gv, long long n_samples, int NV, int Hd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples * NV * Hd) return;
    const long long row = i / Hd;
    gv[i] = gc[(row / NV) * Hd + i % Hd] / (float)NV;
}
// (rows, cols) -> (cols, rows)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

// ---- weight gradient: dW[o][i] += sum_r G[r][o] * act(A[r][i]),  db[o] += sum_r G[r][o] ------------------------------------
// 64x64 tile of dW per CTA over a slice of the rows (blockIdx.z), 4x4 per thread, atomics at the end.
template <bool RELU_A>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ A, int lda, long long rows, int out_dim, int in_dim,
             long long rows_per_split, float* __restrict__ dW, int ldw, float* __restrict__ db) {
    __shared__ float Gs[16][68];
    __shared__ float As[16][68];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int o0 = blockIdx.x * 64, i0 = blockIdx.y * 64;
    const long long r_begin = (long long)blockIdx.z * rows_per_split;
    const long long r_end = r_begin + rows_per_split < rows ? r_begin + rows_per_split : rows;
    float acc[4][4] = {};
    float bsum = 0.0f;      // bias gradient: thread c < 64 of the CTAs with blockIdx.y == 0 sums column o0 + c of G over this row slice
    for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int rr = t >> 6, cc = t & 63;
            const long long r = r0 + rr;
            const bool ok = r < r_end;
            Gs[rr][cc] = (ok && o0 + cc < out_dim) ? G[r * ldg + o0 + cc] : 0.0f;
            float a = (ok && i0 + cc < in_dim) ? A[r * lda + i0 + cc] : 0.0f;
            if (RELU_A) a = fmaxf(a, 0.0f);
            As[rr][cc] = a;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            float gr[4], ar[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { gr[i] = Gs[rr][ty * 4 + i]; ar[i] = As[rr][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gr[i], ar[j], acc[i][j]);
        }
        if (db && blockIdx.y == 0 && threadIdx.x < 64) {
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) bsum += Gs[rr][threadIdx.x];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = o0 + ty * 4 + i;
        if (o >= out_dim) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = i0 + tx * 4 + j;
            if (c < in_dim) atomicAdd(dW + (size_t)o * ldw + c, acc[i][j]);
        }
    }
    if (db && blockIdx.y == 0 && threadIdx.x < 64 && o0 + threadIdx.x < out_dim) atomicAdd(db + o0 + threadIdx.x, bsum);
}

// ---- latent gradient: rows of d(gathered latent) scattered through the bilinear taps into NCHW d_latent ------------------
__global__ void scatter_latent_kernel(SceneDev s, QueryArgs q, long long s_begin, long long n_samples,
                                      const float* __restrict__ gz, float* __restrict__ d_latent) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_rows = n_samples * s.NV;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const size_t plane = (size_t)s.Hl * s.Wl;
    for (long long row = warp_global; row < n_rows; row += n_warps) {
        const long long smp = s_begin + row / s.NV;
        const int v = (int)(row % s.NV);
        const int sb = (int)(smp / q.n_per_sb);
        float px, py, pz;
        if (q.xyz) {
            px = q.xyz[smp * 3]; py = q.xyz[smp * 3 + 1]; pz = q.xyz[smp * 3 + 2];
        } else {
            const float* r = q.rays + (smp / q.K) * 8;
            const float zz = q.z[smp];
            px = __fadd_rn(r[0], __fmul_rn(zz, r[3]));
            py = __fadd_rn(r[1], __fmul_rn(zz, r[4]));
            pz = __fadd_rn(r[2], __fmul_rn(zz, r[5]));
        }
        const int sv = sb * s.NV + v;
        const float* P = s.poses + (size_t)sv * 16;
        float p[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) p[k] = __ldg(P + k);
        float xc, yc, zc;
        world_to_cam(p, px, py, pz, xc, yc, zc);
        const float u = project_axis(xc, zc, __ldg(s.focal + sv * 2), __ldg(s.cxy + sv * 2), s.imgW);
        const float w = project_axis(yc, zc, __ldg(s.focal + sv * 2 + 1), __ldg(s.cxy + sv * 2 + 1), s.imgH);
        const LatTaps t = latent_taps(s, u, w);
        float* base = d_latent + (size_t)sv * s.L * plane;
        for (int c = lane; c < s.L; c += 32) {
            const float g = gz[row * s.L + c];
            float* ch = base + (size_t)c * plane;
            atomicAdd(ch + t.o00, g * t.w00);
            atomicAdd(ch + t.o01, g * t.w01);
            atomicAdd(ch + t.o10, g * t.w10);
            atomicAdd(ch + t.o11, g * t.w11);
        }
    }
}

#define BK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) return _e; } while (0)

inline unsigned grid1d(long long n) { return (unsigned)((n + 255) / 256); }

template <bool RELU_A>
cudaError_t wgrad(const float* G, int ldg, const float* A, int lda, long long rows, int out_dim, int in_dim, float* dW, float* db,
                  cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    const long long per = 4096;                              // rows per split: bounds the atomics per weight to rows / 4096
    const unsigned splits = (unsigned)((rows + per - 1) / per);
    dim3 grid((unsigned)((out_dim + 63) / 64), (unsigned)((in_dim + 63) / 64), splits);
    wgrad_kernel<RELU_A><<<grid, 256, 0, st>>>(G, ldg, A, lda, rows, out_dim, in_dim, per, dW, in_dim, db);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace

// Parameter-gradient buffer layout = the order diner_set_mlp stores the parameters (capi.cu): lin_in w,b; lin_out w,b;
// per block fc_0 w,b, fc_1 w,b; per lin_z block w,b.
size_t backward_param_count(const MlpDev& m) {
    const int nz = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks;
    return (size_t)m.d_hidden * m.d_in + m.d_hidden + (size_t)m.d_out * m.d_hidden + m.d_out +
           (size_t)m.n_blocks * 2 * ((size_t)m.d_hidden * m.d_hidden + m.d_hidden) +
           (size_t)nz * ((size_t)m.d_hidden * m.d_latent + m.d_hidden);
}

size_t backward_workspace_bytes(const MlpDev& m, const SceneDev& s, long long chunk_samples) {
    const long long R = chunk_samples * s.NV;
    const int Hd = m.d_hidden, ld_in = (m.d_in + 7) & ~7;
    const int n_pre = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks, n_post = m.n_blocks - n_pre;
    size_t f = (size_t)R * (ld_in + 2 * (size_t)m.d_latent);                  // xin, zlat, gz
    f += (size_t)R * Hd * (2 * (size_t)n_pre + 3);                            // xa[b], net[b]; running x, gx, gnet (view rows)
    f += (size_t)chunk_samples * Hd * (2 * (size_t)n_post + 4);               // xc_in[b], netc[b]; xc, gxc, gnetc, tmp
    f += (size_t)chunk_samples * 4;                                           // (unused slack)
    // transposed weights
    f += (size_t)m.d_out * Hd + (size_t)m.n_blocks * 2 * Hd * Hd + (size_t)n_pre * Hd * m.d_latent;
    return f * sizeof(float);
}

// workspace of the tcgen05 path: packed transposed weights, G^T (512 x R fp32) and two packed activation buffers (R x 2 KiB each)
size_t backward_tc_workspace_bytes(const MlpDev& m, const SceneDev& s, long long chunk_samples) {
    const long long R = chunk_samples * s.NV;
    const long long nkb = (R + 63) / 64;
    const int n_pre = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks;
    const size_t wt = (size_t)(2 * m.n_blocks + n_pre) * 4 * 8 * 2 * tc::WTILE_BYTES;
    return wt + (size_t)512 * (size_t)(nkb * 64) * sizeof(float) + 2 * (size_t)(4 * nkb) * 2 * tc::WTILE_BYTES + 4096;
}

// d_pre (n,4): pre-activation gradients of the per-sample outputs (composite_backward_kernel).  grad_params / d_latent are
// ACCUMULATED into (the caller zeroes them).  ws = workspace of backward_workspace_bytes(chunk); tcs / tcws != nullptr selects
// the tcgen05 GEMMs (tcws = workspace of backward_tc_workspace_bytes(chunk), needs the 512-wide network packed in tcs).
cudaError_t backward_simt(const SceneDev& s, const MlpDev& m, const QueryArgs& q, const float* d_pre, float* grad_params,
                          float* d_latent, float* ws, long long chunk, cudaStream_t st, TcState* tcs, uint8_t* tcws, int num_sms) {
    const long long total = (long long)q.SB * q.n_per_sb;
    const int Hd = m.d_hidden, L = m.d_latent, ld_in = (m.d_in + 7) & ~7;
    const int n_pre = m.combine_layer < m.n_blocks ? m.combine_layer : m.n_blocks, n_post = m.n_blocks - n_pre;
    if (n_post < 1 || n_pre < 1) return cudaErrorNotSupported;
    const bool use_tc = tcs && tcws && tcs->ready && tcs->wmap_ok && Hd == 512 && L == 512;
    // ---- gradient buffer slices (same order as the parameter store)
    float* gp = grad_params;
    auto take = [&](size_t n) { float* p = gp; gp += n; return p; };
    float* g_w_in = take((size_t)Hd * m.d_in); float* g_b_in = take(Hd);
    float* g_w_out = take((size_t)m.d_out * Hd); float* g_b_out = take(m.d_out);
    float *g_w0[DINER_MAX_BLOCKS], *g_b0[DINER_MAX_BLOCKS], *g_w1[DINER_MAX_BLOCKS], *g_b1[DINER_MAX_BLOCKS];
    float *g_wz[DINER_MAX_BLOCKS], *g_bz[DINER_MAX_BLOCKS];
    for (int b = 0; b < m.n_blocks; ++b) {
        g_w0[b] = take((size_t)Hd * Hd); g_b0[b] = take(Hd);
        g_w1[b] = take((size_t)Hd * Hd); g_b1[b] = take(Hd);
    }
    for (int b = 0; b < n_pre; ++b) { g_wz[b] = take((size_t)Hd * L); g_bz[b] = take(Hd); }
    // ---- workspace slices
    const long long Rcap = chunk * s.NV;
    float* w = ws;
    auto wtake = [&](size_t n) { float* p = w; w += n; return p; };
    float* xin = wtake((size_t)Rcap * ld_in);
    float* zlat = wtake((size_t)Rcap * L);
    float* gz = wtake((size_t)Rcap * L);
    float *xa[DINER_MAX_BLOCKS], *net[DINER_MAX_BLOCKS], *xci[DINER_MAX_BLOCKS], *netc[DINER_MAX_BLOCKS];
    for (int b = 0; b < n_pre; ++b) { xa[b] = wtake((size_t)Rcap * Hd); net[b] = wtake((size_t)Rcap * Hd); }
    float* x = wtake((size_t)Rcap * Hd);
    float* gx = wtake((size_t)Rcap * Hd);
    float* gnet = wtake((size_t)Rcap * Hd);
    for (int b = 0; b < n_post; ++b) { xci[b] = wtake((size_t)chunk * Hd); netc[b] = wtake((size_t)chunk * Hd); }
    float* xc = wtake((size_t)chunk * Hd);
    float* gxc = wtake((size_t)chunk * Hd);
    float* gnetc = wtake((size_t)chunk * Hd);
    float* tmpc = wtake((size_t)chunk * Hd);
    (void)wtake((size_t)chunk * 4);
    float* wT_out = wtake((size_t)m.d_out * Hd);                      // (Hd, d_out)
    float *wT0[DINER_MAX_BLOCKS], *wT1[DINER_MAX_BLOCKS], *wTz[DINER_MAX_BLOCKS];
    for (int b = 0; b < m.n_blocks; ++b) { wT0[b] = wtake((size_t)Hd * Hd); wT1[b] = wtake((size_t)Hd * Hd); }
    for (int b = 0; b < n_pre; ++b) wTz[b] = wtake((size_t)Hd * L);  // (L, Hd)
    auto transpose = [&](const float* src, float* dst, int rows, int cols) -> cudaError_t {
        dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
        transpose_kernel<<<grid, block, 0, st>>>(src, dst, rows, cols);
        g_launches++;
        return cudaGetLastError();
    };
    BK(transpose(m.w_out, wT_out, m.d_out, Hd));
    for (int b = 0; b < m.n_blocks; ++b) { BK(transpose(m.w_fc0[b], wT0[b], Hd, Hd)); BK(transpose(m.w_fc1[b], wT1[b], Hd, Hd)); }
    for (int b = 0; b < n_pre; ++b) BK(transpose(m.w_z[b], wTz[b], Hd, L));

    // ---- tcgen05 path: packed transposed weights, tile-pair offsets of the forward weights in tcs->wpack, GEMM wrappers
    const long long Rmax = chunk * s.NV, nkb_max = (Rmax + 63) / 64;
    const int grid_pairs = num_sms / 2;
    uint8_t* wt_pack = tcws;                                                       // [2 n_blocks + n_pre][4 x 8 tile pairs]
    const size_t layer_bytes = (size_t)4 * 8 * 2 * tc::WTILE_BYTES;
    float* GT = use_tc ? (float*)(tcws + (size_t)(2 * m.n_blocks + n_pre) * layer_bytes) : nullptr;      // (512, nkb_max * 64)
    uint8_t* AP = use_tc ? (uint8_t*)(GT + (size_t)512 * nkb_max * 64) : nullptr;                         // packed activations
    uint8_t* ZP = use_tc ? AP + (size_t)(4 * nkb_max) * 2 * tc::WTILE_BYTES : nullptr;                   // packed zlat (shared by the lin_z blocks)
    float* gscale = use_tc ? (float*)(ZP + (size_t)(4 * nkb_max) * 2 * tc::WTILE_BYTES) : nullptr;        // {s, 1/s, amax bits}: loss scaling of the gradient operands
    if (use_tc) {
        BK(cudaMemsetAsync(gscale, 0, 16, st));
        tc3::absmax_kernel<<<296, 256, 0, st>>>(d_pre, total * 4, (unsigned int*)(gscale + 2));
        tc3::make_scale_kernel<<<1, 1, 0, st>>>((const unsigned int*)(gscale + 2), gscale);
        g_launches += 2;
        BK(cudaGetLastError());
    }
    static Tc3Map map_wt, map_ap, map_zp;
    long long pair_z[DINER_MAX_BLOCKS], pair_0[DINER_MAX_BLOCKS], pair_1[DINER_MAX_BLOCKS];   // offsets in tcs->wpack (tc_pack_weights order)
    {
        long long p = tc::MT * 1;                                                  // lin_in
        for (int b = 0; b < n_pre; ++b) { pair_z[b] = p; p += tc::MT * (L / 64); pair_0[b] = p; p += tc::MT * 8; pair_1[b] = p; p += tc::MT * 8; }
        for (int b = n_pre; b < m.n_blocks; ++b) { pair_0[b] = p; p += tc::MT * 8; pair_1[b] = p; p += tc::MT * 8; }
    }
    auto wt_pair = [&](int idx) { return (long long)idx * 4 * 8; };                // idx: 2b = W0^T of block b, 2b+1 = W1^T, 2 n_blocks + b = Wz^T
    if (use_tc) {
        auto packT = [&](const float* wT, int idx) -> cudaError_t {              // wT is (in, out) row-major = the "weight" of the data-gradient GEMM
            tc::pack_weight_kernel<<<4 * 8, 256, 0, st>>>(wT, 512, 512, 8, wt_pack + (size_t)idx * layer_bytes);
            g_launches++;
            return cudaGetLastError();
        };
        for (int b = 0; b < m.n_blocks; ++b) { BK(packT(wT0[b], 2 * b)); BK(packT(wT1[b], 2 * b + 1)); }
        for (int b = 0; b < n_pre; ++b) BK(packT(wTz[b], 2 * m.n_blocks + b));
        BK(tc3_make_map(map_wt, wt_pack, (size_t)(2 * m.n_blocks + n_pre) * layer_bytes));
        BK(tc3_make_map(map_ap, AP, (size_t)(4 * nkb_max) * 2 * tc::WTILE_BYTES));
        BK(tc3_make_map(map_zp, ZP, (size_t)(4 * nkb_max) * 2 * tc::WTILE_BYTES));
    }
    // Y (rows x 512) (+)= act(X (rows x 512)) . W^T + bias   [forward weights packed in tcs->wpack]
    // (`add` != nullptr: Y = add + ...; lets the residual recurrence write each block input into its own buffer without a copy)
    auto tc_fwd = [&](const float* X, bool relu_in, long long pair0, const float* bias, float* Y, const float* add, long long rows) -> cudaError_t {
        tc3::Args a{};
        a.bmap = tcs->wmap; a.A = X; a.lda = 512; a.rows = rows; a.K = 512; a.nkb_total = 8; a.relu_a = relu_in; a.b_pair0 = pair0;
        a.n_slices = 1; a.kb_per_slice = 8; a.C = Y; a.ldc = 512; a.mode = 0; a.scale = tc::W_INV; a.bias = bias; a.add = add; a.ldd = 512;
        return tc3_gemm(a, grid_pairs, tcs->err_flag, st);
    };
    // GX (rows x 512) (+)= (G (rows x 512) . W) * (mask > 0)   [W^T packed in wt_pack]
    auto tc_dgrad = [&](const float* G, int idx, const float* mask, float* GX, bool accum, long long rows) -> cudaError_t {
        tc3::Args a{};
        a.bmap = map_wt.map; a.A = G; a.lda = 512; a.rows = rows; a.K = 512; a.nkb_total = 8; a.relu_a = 0; a.b_pair0 = wt_pair(idx);
        a.n_slices = 1; a.kb_per_slice = 8; a.C = GX; a.ldc = 512; a.mode = accum ? 1 : 0; a.scale = tc::W_INV; a.mask = mask; a.ldm = 512;
        a.a_scale = gscale;
        return tc3_gemm(a, grid_pairs, tcs->err_flag, st);
    };
    // packs act(X)^T (rows x 512) into `dst` as the B operand of a weight-gradient GEMM
    auto tc_pack_act = [&](const float* X, bool relu, long long rows, uint8_t* dst) -> cudaError_t {
        const long long nkb = (rows + 63) / 64;
        tc3::pack_rows_kernel<<<(unsigned)(4 * nkb), 256, 0, st>>>(X, 512, rows, (int)nkb, relu ? 1 : 0, dst);
        g_launches++;
        return cudaGetLastError();
    };
    // dW (512 x 512) += G^T . act(X);  db += column sums of G.  `packed` = tc_pack_act(X) (AP or ZP)
    auto tc_wgrad = [&](const float* G, const Tc3Map& pmap, long long rows, float* dW, float* db) -> cudaError_t {
        const long long nkb = (rows + 63) / 64;
        dim3 tg(512 / 32, (unsigned)((nkb * 64 + 255) / 256)), tb(32, 8);
        tc3::transpose_pad_kernel<<<tg, tb, 0, st>>>(G, GT, rows, 512, nkb * 64, db);   // (rows, 512) -> (512, nkb * 64), zero padded; db += column sums
        g_launches++;
        tc3::Args a{};
        a.bmap = pmap.map; a.A = GT; a.lda = nkb * 64; a.rows = 512; a.K = rows; a.nkb_total = (int)nkb; a.relu_a = 0; a.b_pair0 = 0;
        a.a_scale = gscale;
        int slices = grid_pairs / 4 > 0 ? grid_pairs / 4 : 1;                      // 4 row tiles of 128 outputs x K slices ~ one item per CTA pair
        if (slices > nkb) slices = (int)nkb;
        a.kb_per_slice = (int)((nkb + slices - 1) / slices);
        a.n_slices = (int)((nkb + a.kb_per_slice - 1) / a.kb_per_slice);
        a.C = dW; a.ldc = 512; a.mode = 2; a.scale = 1.0f;
        BK(tc3_gemm(a, grid_pairs, tcs->err_flag, st));
        return cudaGetLastError();
    };

    for (long long s0 = 0; s0 < total; s0 += chunk) {
        const long long ns = total - s0 < chunk ? total - s0 : chunk;
        const long long R = ns * s.NV;
        // ================= forward recompute, keeping every block input =================
        const int fgrid = (int)((R * 32 + 255) / 256 < 148 * 64 ? (R * 32 + 255) / 256 : 148 * 64);
        features_kernel<<<fgrid, 256, 0, st>>>(s, q, s0, ns, m.d_in, ld_in, xin, zlat);
        g_launches++;
        BK(cudaGetLastError());
        BK((linear<false, false>(xin, ld_in, m.w_in, m.d_in, m.b_in, x, Hd, R, m.d_in, Hd, st)));
        for (int b = 0; b < n_pre; ++b) {
            if (use_tc) {
                BK(tc_fwd(zlat, false, pair_z[b], m.b_z[b], xa[b], x, R));               // xa[b] = x + lin_z[b](z)
                BK(tc_fwd(xa[b], true, pair_0[b], m.b_fc0[b], net[b], nullptr, R));
                BK(tc_fwd(net[b], true, pair_1[b], m.b_fc1[b], x, xa[b], R));              // x = xa[b] + fc_1(relu(net))
                continue;
            }
            BK((linear<false, true>(zlat, L, m.w_z[b], L, m.b_z[b], x, Hd, R, L, Hd, st)));
            BK(cudaMemcpyAsync(xa[b], x, (size_t)R * Hd * sizeof(float), cudaMemcpyDeviceToDevice, st));
            BK((linear<true, false>(x, Hd, m.w_fc0[b], Hd, m.b_fc0[b], net[b], Hd, R, Hd, Hd, st)));
            BK((linear<true, true>(net[b], Hd, m.w_fc1[b], Hd, m.b_fc1[b], x, Hd, R, Hd, Hd, st)));
        }
        combine_kernel<<<grid1d(ns * Hd), 256, 0, st>>>(x, xc, ns, s.NV, Hd);
        g_launches++;
        BK(cudaGetLastError());
        for (int b = 0; b < n_post; ++b) {
            const int B = n_pre + b;
            BK(cudaMemcpyAsync(xci[b], xc, (size_t)ns * Hd * sizeof(float), cudaMemcpyDeviceToDevice, st));
            if (use_tc) {
                BK(tc_fwd(xci[b], true, pair_0[B], m.b_fc0[B], netc[b], nullptr, ns));
                BK(tc_fwd(netc[b], true, pair_1[B], m.b_fc1[B], xc, xci[b], ns));
                continue;
            }
            BK((linear<true, false>(xc, Hd, m.w_fc0[B], Hd, m.b_fc0[B], netc[b], Hd, ns, Hd, Hd, st)));
            BK((linear<true, true>(netc[b], Hd, m.w_fc1[B], Hd, m.b_fc1[B], xc, Hd, ns, Hd, Hd, st)));
        }
        // ================= backward =================
        const float* dpre = d_pre + s0 * 4;
        // lin_out: out = W_out relu(xc) + b
        BK((wgrad<true>(dpre, 4, xc, Hd, ns, m.d_out, Hd, g_w_out, g_b_out, st)));
        BK((linear<false, false>(dpre, 4, wT_out, m.d_out, nullptr, gxc, Hd, ns, m.d_out, Hd, st)));
        relu_mask_kernel<<<grid1d(ns * Hd), 256, 0, st>>>(gxc, xc, ns * Hd);
        g_launches++;
        for (int b = n_post - 1; b >= 0; --b) {
            const int B = n_pre + b;
            // x_out = x_in + fc_1(relu(net)),  net = fc_0(relu(x_in));  gxc = dL/dx_out
            if (use_tc) {
                BK(tc_pack_act(netc[b], true, ns, AP));
                BK(tc_wgrad(gxc, map_ap, ns, g_w1[B], g_b1[B]));
                BK(tc_dgrad(gxc, 2 * B + 1, netc[b], gnetc, false, ns));                    // gnetc = (gxc W1) * (net > 0)
                BK(tc_pack_act(xci[b], true, ns, AP));
                BK(tc_wgrad(gnetc, map_ap, ns, g_w0[B], g_b0[B]));
                BK(tc_dgrad(gnetc, 2 * B, xci[b], gxc, true, ns));                           // gxc += (gnetc W0) * (x_in > 0)
                continue;
            }
            BK((wgrad<true>(gxc, Hd, netc[b], Hd, ns, Hd, Hd, g_w1[B], g_b1[B], st)));
            BK((linear<false, false>(gxc, Hd, wT1[B], Hd, nullptr, gnetc, Hd, ns, Hd, Hd, st)));
            relu_mask_kernel<<<grid1d(ns * Hd), 256, 0, st>>>(gnetc, netc[b], ns * Hd);
            g_launches++;
            BK((wgrad<true>(gnetc, Hd, xci[b], Hd, ns, Hd, Hd, g_w0[B], g_b0[B], st)));
            BK((linear<false, false>(gnetc, Hd, wT0[B], Hd, nullptr, tmpc, Hd, ns, Hd, Hd, st)));
            relu_mask_add_kernel<<<grid1d(ns * Hd), 256, 0, st>>>(gxc, tmpc, xci[b], ns * Hd);
            g_launches++;
        }
        combine_backward_kernel<<<grid1d(R * Hd), 256, 0, st>>>(gxc, gx, ns, s.NV, Hd);
        g_launches++;
        if (!use_tc) BK(cudaMemsetAsync(gz, 0, (size_t)R * L * sizeof(float), st));     // (the tcgen05 path stores the first block's contribution)
        for (int b = n_pre - 1; b >= 0; --b) {
            // x_out = xa + fc_1(relu(net)),  net = fc_0(relu(xa)),  xa = x_prev + lin_z[b](zlat);  gx = dL/dx_out
            if (use_tc) {
                if (b == n_pre - 1) BK(tc_pack_act(zlat, false, R, ZP));                     // shared by the lin_z weight gradients of all blocks
                BK(tc_pack_act(net[b], true, R, AP));
                BK(tc_wgrad(gx, map_ap, R, g_w1[b], g_b1[b]));
                BK(tc_dgrad(gx, 2 * b + 1, net[b], gnet, false, R));                         // gnet = (gx W1) * (net > 0)
                BK(tc_pack_act(xa[b], true, R, AP));
                BK(tc_wgrad(gnet, map_ap, R, g_w0[b], g_b0[b]));
                BK(tc_dgrad(gnet, 2 * b, xa[b], gx, true, R));                               // gx += (gnet W0) * (xa > 0)  = dL/dxa
                BK(tc_wgrad(gx, map_zp, R, g_wz[b], g_bz[b]));
                BK(tc_dgrad(gx, 2 * m.n_blocks + b, nullptr, gz, b != n_pre - 1, R));        // gz (+)= gx W_z
                continue;
            }
            BK((wgrad<true>(gx, Hd, net[b], Hd, R, Hd, Hd, g_w1[b], g_b1[b], st)));
            BK((linear<false, false>(gx, Hd, wT1[b], Hd, nullptr, gnet, Hd, R, Hd, Hd, st)));
            relu_mask_kernel<<<grid1d(R * Hd), 256, 0, st>>>(gnet, net[b], R * Hd);
            g_launches++;
            BK((wgrad<true>(gnet, Hd, xa[b], Hd, R, Hd, Hd, g_w0[b], g_b0[b], st)));
            BK((linear<false, false>(gnet, Hd, wT0[b], Hd, nullptr, x, Hd, R, Hd, Hd, st)));       // x reused as scratch
            relu_mask_add_kernel<<<grid1d(R * Hd), 256, 0, st>>>(gx, x, xa[b], R * Hd);               // gx = dL/dxa
            g_launches++;
            BK((wgrad<false>(gx, Hd, zlat, L, R, Hd, L, g_wz[b], g_bz[b], st)));
            BK((linear<false, true>(gx, Hd, wTz[b], Hd, nullptr, gz, L, R, Hd, L, st)));             // gz += gx W_z
        }
        BK((wgrad<false>(gx, Hd, xin, ld_in, R, Hd, m.d_in, g_w_in, g_b_in, st)));
        if (d_latent) {
            const int sgrid = (int)((R * 32 + 255) / 256 < 148 * 32 ? (R * 32 + 255) / 256 : 148 * 32);
            scatter_latent_kernel<<<sgrid, 256, 0, st>>>(s, q, s0, ns, gz, d_latent);
            g_launches++;
        }
        BK(cudaGetLastError());
    }
    return cudaSuccess;
}

cudaError_t launch_composite_backward(const float* rays, const float* z, const float* net_out, long long n_rays, int K, int white,
                                      const float* g_rgb, const float* g_depth, float* d_pre, cudaStream_t st) {
    composite_backward_kernel<<<grid1d(n_rays), 256, 0, st>>>(rays, z, net_out, n_rays, K, white, g_rgb, g_depth, d_pre);
    g_launches++;
    return cudaGetLastError();
}
