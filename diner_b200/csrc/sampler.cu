// Depth-guided ray sampler: one warp per ray, everything on-chip.
//
// Fuses reference NeRFRendererDGS.sample_coarse (nerf_renderer.py:39-63), sample_depthguided
// (:65-190), fill_up_uniform_samples (:367-397) and the helper weighted_mean_n_std
// (torch_helpers.py:215-223).  The reference materialises ~10 tensors of shape (SB,NV,NR*C,3) in
// HBM per ray batch and argsorts 1000-wide rows; here the C candidates of a ray live in shared
// memory (2 floats each), depth/std/normal maps are read through L2, and the only HBM traffic is
// the 8-float ray in and the K sorted sample depths out.
#include "common.cuh"
#include "diner_internal.h"

namespace {

constexpr int WARPS_PER_CTA = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ascending rank sort of n floats (ties keep index order); src/dst are per-warp shared arrays
__device__ __forceinline__ void warp_rank_sort(const float* src, float* dst, int n, int lane) {
    for (int e = lane; e < n; e += 32) {
        float v = src[e];
        int rank = 0;
        for (int x = 0; x < n; ++x) {
            float o = src[x];
            rank += (o < v) || (o == v && x < e);
        }
        dst[rank] = v;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
sampler_kernel(SceneDev s, SamplerArgs a) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cpad = (a.C + 31) & ~31;
    float* zc = smem + (size_t)warp * (2 * Cpad + 2 * a.K);   // candidate depths
    float* lk = zc + Cpad;                                    // likelihood (max over views)
    float* buf0 = lk + Cpad;                                  // K slots
    float* buf1 = buf0 + a.K;

    const long long total = (long long)a.SB * a.NR;
    for (long long ray = (long long)blockIdx.x * WARPS_PER_CTA + warp; ray < total;
         ray += (long long)gridDim.x * WARPS_PER_CTA) {
        const int sb = (int)(ray / a.NR);
        // counter-based fallback noise is keyed by the ray's LOGICAL index (scene, ray + ray_offset), so that a shard of a ray list
        // (multi-GPU, chunked calls) draws what the whole list would draw; dense injected noise is indexed by the local ray
        const long long rkey = ((long long)sb << 40) + (ray - (long long)sb * a.NR) + (long long)a.ray_offset;
        const float* r = a.rays + ray * 8;
        const float ox = r[0], oy = r[1], oz = r[2], dx = r[3], dy = r[4], dz = r[5];
        const float near = r[6], far = r[7];
        const int C = a.C, K = a.K, G = a.G;

        // ---- sample_coarse (nerf_renderer.py:51-59); torch.linspace on CPU is evaluated
        //      symmetrically with one fma per element (probed)
        for (int i = lane; i < Cpad; i += 32) {
            float z = 0.0f;
            if (i < C) {
                float lin = (i < C / 2) ? __fmul_rn(a.lin_step, (float)i)
                                        : fmaf(-a.lin_step, (float)(C - 1 - i), a.lin_end);
                float u = a.u_coarse ? a.u_coarse[ray * C + i] : rng_uniform(a.seed, 1, rkey * C + i);
                float sfrac = __fadd_rn(lin, __fmul_rn(u, a.cstep));
                z = __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, sfrac)), __fmul_rn(far, sfrac));
            }
            zc[i] = z;
            lk[i] = 0.0f;
        }
        __syncwarp();

        // ---- per candidate x view likelihood (nerf_renderer.py:95-129)
        const float step = __fdiv_rn(__fsub_rn(far, near), (float)C);
        const float half = __fmul_rn(step, 0.5f);
        for (int v = 0; v < s.NV; ++v) {
            const int sv = sb * s.NV + v;
            const float* P = s.poses + (size_t)sv * 16;
            float p[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) p[q] = __ldg(P + q);
            const float fx = __ldg(s.focal + sv * 2), fy = __ldg(s.focal + sv * 2 + 1);
            const float cx = __ldg(s.cxy + sv * 2), cy = __ldg(s.cxy + sv * 2 + 1);
            float rdx, rdy, rdz;
            rotate_to_cam(p, dx, dy, dz, rdx, rdy, rdz);
            for (int i = lane; i < C; i += 32) {
                const float z = zc[i];
                const float wx = __fadd_rn(ox, __fmul_rn(z, dx));
                const float wy = __fadd_rn(oy, __fmul_rn(z, dy));
                const float wz = __fadd_rn(oz, __fmul_rn(z, dz));
                float xc, yc, zcam;
                world_to_cam(p, wx, wy, wz, xc, yc, zcam);
                const float u = project_axis(xc, zcam, fx, cx, s.imgW);
                const float w = project_axis(yc, zcam, fy, cy, s.imgH);
                const float sd = lookup_std(s, sv, u, w);
                if (sd == 0.0f) continue;                                         // bg_mask      (:122)
                const float dref = lookup_depth(s, sv, u, w);
                if (!(fabsf(__fsub_rn(dref, zcam)) < a.depth_diff_max)) continue;            // depth_dist   (:121)
                float nx, ny, nz;
                lookup_normal(s, sv, u, w, nx, ny, nz);
                const float cosd = __fadd_rn(__fadd_rn(__fmul_rn(rdx, nx), __fmul_rn(rdy, ny)),
                                             __fmul_rn(rdz, nz));
                if (!(cosd <= 0.0f)) continue;                                    // cosdist mask (:120)
                const float den = __fmul_rn(sd, 1.41421354f);
                const float ea = erff(__fdiv_rn(__fsub_rn(__fadd_rn(zcam, half), dref), den));
                const float eb = erff(__fdiv_rn(__fsub_rn(__fsub_rn(zcam, half), dref), den));
                const float l = fabsf(__fmul_rn(0.5f, __fsub_rn(ea, eb)));        // (:125-128)
                lk[i] = fmaxf(lk[i], l);                                          // max over views (:129)
            }
            __syncwarp();
        }

        // ---- occlusion-aware likelihood = lik * exclusive cumprod(1 - lik)   (:131-132)
        //      and its weighted mean / std over the candidates (torch_helpers.py:215-223)
        float run = 1.0f, wsum = 0.0f;
        bool any_nz = false;
        for (int base = 0; base < Cpad; base += 32) {
            const float l = lk[base + lane];
            float inc = 1.0f - l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc *= t;
            }
            float excl = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) excl = 1.0f;
            const float opq = l * (run * excl);
            run *= __shfl_sync(0xffffffffu, inc, 31);
            any_nz |= (opq != 0.0f);
            wsum += opq;
            // stash opaque weight in place of nothing: recomputed in the second pass (cheap)
        }
        any_nz = __any_sync(0xffffffffu, any_nz);
        wsum = warp_sum(wsum);
        float mean = 0.0f, sdev = 0.0f;
        if (G > 0 && any_nz) {
            // second + third pass recompute opaque (registers are cheaper than another C floats of smem)
            for (int pass = 0; pass < 2; ++pass) {
                float acc = 0.0f;
                run = 1.0f;
                for (int base = 0; base < Cpad; base += 32) {
                    const float l = lk[base + lane];
                    float inc = 1.0f - l;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        float t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc *= t;
                    }
                    float excl = __shfl_up_sync(0xffffffffu, inc, 1);
                    if (lane == 0) excl = 1.0f;
                    const float wn = (l * (run * excl)) / wsum;
                    run *= __shfl_sync(0xffffffffu, inc, 31);
                    const float z = zc[base + lane];
                    if (pass == 0) acc += z * wn;
                    else { const float dlt = z - mean; acc += dlt * dlt * wn; }
                }
                acc = warp_sum(acc);
                if (pass == 0) mean = acc; else sdev = sqrtf(acc);
            }
        }

        // ---- compact the non-zero-likelihood candidates (zero ones become empty slots anyway, :176-178)
        int nnz = 0;
        for (int base = 0; base < Cpad; base += 32) {
            const float l = lk[base + lane], z = zc[base + lane];
            const unsigned m = __ballot_sync(0xffffffffu, l != 0.0f);
            __syncwarp();
            if (l != 0.0f) {
                const int pos = nnz + __popc(m & ((1u << lane) - 1));
                lk[pos] = l;
                zc[pos] = z;
            }
            nnz += __popc(m);
            __syncwarp();
        }

        // ---- top-(K-G) by likelihood (descending; ties -> lower candidate index)   (:172-177)
        const int NT = K - G;
        for (int t = 0; t < NT; ++t) {
            float bv = 0.0f;
            int bi = 0x7fffffff;
            for (int i = lane; i < nnz; i += 32) {
                const float l = lk[i];
                if (l > bv) { bv = l; bi = i; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bv == 0.0f) {                     // nothing left: remaining slots stay empty
                for (int q = t + lane; q < NT; q += 32) buf0[q] = 0.0f;
                break;
            }
            if (lane == 0) { buf0[t] = zc[bi]; lk[bi] = 0.0f; }
            __syncwarp();
        }
        // ---- Gaussian samples overwrite the last G slots (:181-190); not clamped, may be < near
        for (int g = lane; g < G; g += 32) {
            float val = 0.0f;
            if (any_nz) {
                const float n = a.g_noise ? a.g_noise[ray * G + g] : rng_normal(a.seed, 2, rkey * G + g);
                val = __fadd_rn(__fmul_rn(n, sdev), mean);
            }
            buf0[NT + g] = val;
        }
        __syncwarp();

        // ---- fill_up_uniform_samples (:367-397): sort, fill zero slots by their column, sort again
        warp_rank_sort(buf0, buf1, K, lane);
        if (a.z_dgs) for (int e = lane; e < K; e += 32) a.z_dgs[ray * K + e] = buf1[e];
        int nmiss = 0;
        for (int e = lane; e < K; e += 32) nmiss += (buf1[e] == 0.0f);
#pragma unroll
        for (int o = 16; o; o >>= 1) nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
        if (nmiss > 0) {
            const float fstep = __fdiv_rn(__fsub_rn(far, near), (float)nmiss);
            for (int e = lane; e < K; e += 32) {
                if (buf1[e] == 0.0f) {
                    const float u = a.u_fill ? a.u_fill[ray * K + e] : rng_uniform(a.seed, 3, rkey * K + e);
                    float z = __fadd_rn(near, __fmul_rn((float)e, fstep));
                    buf1[e] = __fadd_rn(z, __fmul_rn(u, fstep));
                }
            }
        }
        __syncwarp();
        warp_rank_sort(buf1, buf0, K, lane);
        for (int e = lane; e < K; e += 32) a.z_out[ray * K + e] = buf0[e];
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_sampler(const SceneDev& s, const SamplerArgs& a, int num_sms, cudaStream_t st) {
    const int Cpad = (a.C + 31) & ~31;
    const size_t smem = (size_t)WARPS_PER_CTA * (2 * Cpad + 2 * a.K) * sizeof(float);
    if (smem > 48 * 1024) {     // per launch: the attribute is per device and one process may drive several (a few hundred ns)
        cudaError_t e = cudaFuncSetAttribute(sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long total = (long long)a.SB * a.NR;
    long long want = (total + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const long long cap = (long long)num_sms * 8;   // grid-stride: a multiple of the SM count
    const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    sampler_kernel<<<grid, WARPS_PER_CTA * 32, smem, st>>>(s, a);
    return cudaGetLastError();
}
