// nlm_filter.cuh -- NonLocalMeansFilter (Kernel/ImagePipeline/Filter/NonLocalMeansFilter.{h,cu}), the second ImageSamplesFilter of the reference's
// image pipeline: variance-guided non-local means over a 13x13 search window (R = 6) with 7x7 patches (F = 3).
//
//   reference                                                       here
//   ---------                                                       ----
//   copyToCached: PixelData -> RGBE, one image                      k_nlm_prepare: same + computeVariance() rounded to half, kept as float (the reference
//                                                                    rounds when it fills its shared-memory cache, NonLocalMeansFilter.cu:14,32)
//   computeWeights: 64x64 RGBE/half tile per 16x16 block, colours   k_nlm_weights: 34x34 tile (exactly the 16 + 2 (R + F) pixels a block reads) of DECODED colours
//     decoded from RGBE at each of the 2 x 49 x 169 loads per pixel,   and scaled variances in shared memory: one decode per tile cell instead of 16 562 per pixel;
//     launched in 200-pixel super-blocks                               one launch
//   weight buffer [pixel][169] floats                               [169][pixel]: a warp's 32 pixels write / read 128 consecutive bytes per slot
//   applyWeights                                                    k_nlm_apply<TO_OUTPUT>: 28x28 colour tile; fused with copyFilteredToOutput when no tone mapper follows
//
// Arithmetic is the reference's, operation by operation (patch loop order, reciprocal multiply of Spectrum / float, 0.05 weight cut-off, NaN skip);
// exp is taken in double and rounded, which agrees with the host libm's expf the oracle is pinned with except for rare last-bit ties.
#pragma once
#include "image_pipeline.cuh"
#include <cuda_fp16.h>

namespace ctld {

constexpr int NLM_R = 6, NLM_F = 3, NLM_WIN = 2 * NLM_R + 1, NLM_NW = NLM_WIN * NLM_WIN, NLM_B = 16;
constexpr int NLM_TW = NLM_B + 2 * (NLM_R + NLM_F);   // 34: weights tile
constexpr int NLM_TA = NLM_B + 2 * NLM_R;             // 28: apply tile

// copyToCached (NonLocalMeansFilter.cu:150-158) + PixelVarianceInfo::computeVariance (Kernel/PixelVarianceBuffer.h:44-47, Math/VarAccumulator.h:7-11) -> half -> float
__global__ void __launch_bounds__(256) k_nlm_prepare(const float* __restrict__ accum, const ctl_pixel_variance_info* __restrict__ var, int n_pixels, float splat_scale,
                                                     uchar4* __restrict__ cached, float* __restrict__ varh) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        float c[3]; px_to_spectrum(accum + (size_t)i * 7, splat_scale, c);
        cached[i] = to_rgbe(c);
        if (var) {
            const float sum_x = var[i].sum_x, sum_x2 = var[i].sum_x2, invN = 1.0f / (float)var[i].num_samples_var;
            const float v = (sum_x2 - (sum_x * sum_x) * invN) * invN;
            varh[i] = __half2float(__float2half_rn(v));
        }
    }
}

// computeWeights (NonLocalMeansFilter.cu:100-121) with weight (:91-98) and patchDistance (:67-89)
__global__ void __launch_bounds__(NLM_B * NLM_B) k_nlm_weights(const uchar4* __restrict__ cached, const float* __restrict__ varh, int w, int h, float k, float sigma2Scale,
                                                               float* __restrict__ weights) {
    __shared__ float s_r[NLM_TW * NLM_TW], s_g[NLM_TW * NLM_TW], s_b[NLM_TW * NLM_TW], s_v[NLM_TW * NLM_TW];
    const int bx0 = blockIdx.x * NLM_B - (NLM_R + NLM_F), by0 = blockIdx.y * NLM_B - (NLM_R + NLM_F);
    for (int cidx = threadIdx.y * NLM_B + threadIdx.x; cidx < NLM_TW * NLM_TW; cidx += NLM_B * NLM_B) {
        const int gx = bx0 + cidx % NLM_TW, gy = by0 + cidx / NLM_TW;
        float c[3] = {0.0f, 0.0f, 0.0f}, v = 0.0f;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) { from_rgbe(cached[(size_t)gy * w + gx], c); v = varh[(size_t)gy * w + gx] * sigma2Scale; }
        s_r[cidx] = c[0]; s_g[cidx] = c[1]; s_b[cidx] = c[2]; s_v[cidx] = v;
    }
    __syncthreads();
    const int x = blockIdx.x * NLM_B + threadIdx.x, y = blockIdx.y * NLM_B + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t n_pixels = (size_t)w * h, pix = (size_t)y * w + x;
    const int lx = threadIdx.x + NLM_R + NLM_F, ly = threadIdx.y + NLM_R + NLM_F;   // tile coordinates of p
    const float eps = 1e-10f, kk = k * k;
    for (int xo = -NLM_R; xo <= NLM_R; xo++)
        for (int yo = -NLM_R; yo <= NLM_R; yo++) {
            const int qx = x + xo, qy = y + yo;
            if (qx < 0 || qx >= w || qy < 0 || qy >= h) continue;
            // patch offsets with p + d and q + d inside the image (the reference tests every tap, :75-77): a rectangle, visited in the same dx-major order
            const int dx0 = max(-NLM_F, max(-x, -qx)), dx1 = min(NLM_F, min(w - 1 - x, w - 1 - qx));
            const int dy0 = max(-NLM_F, max(-y, -qy)), dy1 = min(NLM_F, min(h - 1 - y, h - 1 - qy));
            float d_range = 0.0f, cnt = 0.0f;
            for (int dx = dx0; dx <= dx1; dx++)
                for (int dy = dy0; dy <= dy1; dy++) {
                    const int ip = (ly + dy) * NLM_TW + (lx + dx), iq = (ly + yo + dy) * NLM_TW + (lx + xo + dx);
                    const float var_p = s_v[ip], var_q = s_v[iq];
                    const float er = s_r[ip] - s_r[iq], eg = s_g[ip] - s_g[iq], eb = s_b[ip] - s_b[iq];
                    const float u_diff = ((er * er + eg * eg) + eb * eb) * (1.0f / 3);
                    const float d = (u_diff - (var_p + fminf(var_p, var_q))) / (eps + kk * (var_p + var_q));
                    d_range += d; cnt += 1.0f;
                }
            const float dist = cnt != 0.0f ? d_range / cnt : 0.0f;
            const float we = (float)exp((double)-fmaxf(0.0f, dist));
            weights[(size_t)((yo + NLM_R) * NLM_WIN + (xo + NLM_R)) * n_pixels + pix] = we < 0.05f ? 0.0f : we;
        }
}

// applyWeights (NonLocalMeansFilter.cu:123-148); TO_OUTPUT fuses copyFilteredToOutput (ImagePipeline.cu:32-41)
template <bool TO_OUTPUT>
__global__ void __launch_bounds__(NLM_B * NLM_B) k_nlm_apply(const uchar4* __restrict__ cached, const float* __restrict__ weights, int w, int h, uchar4* __restrict__ out) {
    __shared__ float s_r[NLM_TA * NLM_TA], s_g[NLM_TA * NLM_TA], s_b[NLM_TA * NLM_TA];
    const int bx0 = blockIdx.x * NLM_B - NLM_R, by0 = blockIdx.y * NLM_B - NLM_R;
    for (int cidx = threadIdx.y * NLM_B + threadIdx.x; cidx < NLM_TA * NLM_TA; cidx += NLM_B * NLM_B) {
        const int gx = bx0 + cidx % NLM_TA, gy = by0 + cidx / NLM_TA;
        float c[3] = {0.0f, 0.0f, 0.0f};
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) from_rgbe(cached[(size_t)gy * w + gx], c);
        s_r[cidx] = c[0]; s_g[cidx] = c[1]; s_b[cidx] = c[2];
    }
    __syncthreads();
    const int x = blockIdx.x * NLM_B + threadIdx.x, y = blockIdx.y * NLM_B + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t n_pixels = (size_t)w * h, pix = (size_t)y * w + x;
    const int lx = threadIdx.x + NLM_R, ly = threadIdx.y + NLM_R;
    float acc[3] = {0.0f, 0.0f, 0.0f}, C_p = 0.0f;
    for (int xo = -NLM_R; xo <= NLM_R; xo++)
        for (int yo = -NLM_R; yo <= NLM_R; yo++) {
            const int qx = x + xo, qy = y + yo;
            if (qx < 0 || qx >= w || qy < 0 || qy >= h) continue;
            const float we = __ldg(weights + (size_t)((yo + NLM_R) * NLM_WIN + (xo + NLM_R)) * n_pixels + pix);
            if (we != we) continue;
            const int iq = (ly + yo) * NLM_TA + (lx + xo);
            C_p += we;
            acc[0] += we * s_r[iq]; acc[1] += we * s_g[iq]; acc[2] += we * s_b[iq];
        }
    float c[3];
    if (C_p > 1e-4f) { const float r = 1.0f / C_p; c[0] = acc[0] * r; c[1] = acc[1] * r; c[2] = acc[2] * r; }
    else { const int ip = ly * NLM_TA + lx; c[0] = s_r[ip]; c[1] = s_g[ip]; c[2] = s_b[ip]; }
    const uchar4 q = to_rgbe(c);
    if (TO_OUTPUT) { from_rgbe(q, c); out[pix] = gamma_rgba8(c[0], c[1], c[2]); }
    else out[pix] = q;
}

// copyFilteredToOutput alone is never needed: k_nlm_apply<true> covers "filter, no process"; with a tone mapper the stage feeds k_lum_blocks / k_reinhard.

} // namespace ctld
