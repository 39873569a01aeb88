// image_pipeline.cuh -- the step after the path: applyImagePipeline (Kernel/ImagePipeline/ImagePipeline.cu:14-84) and the
// per-pass PixelVarianceBuffer update (Kernel/PixelVarianceBuffer.h:19-42, .cu:10-36).  Bandwidth kernels: 28 B PixelData in, 4 B out per
// pixel (+ 4 B RGBE stage when a tone mapper runs).
//
//   no filter, no process : PixelData::toSpectrum -> sRGB -> RGBA8                               (copySamplesToOutput)
//   filter                : CanonicalFilter::Apply = evalFilter over Box / Gaussian / Triangle / Mitchell / LanczosSinc (SceneTypes/Filter.h)
//                           -> RGBE stage -> fromRGBE -> sRGB -> RGBA8                            (rtm_Copy + copyFilteredToOutput)
//   process (tone map)    : RGBE stage -> Image::ComputeLuminanceInfo -> Reinhard05Kernel -> RGBA8 -> fromRGBCOL -> sRGB -> RGBA8
#pragma once
#include "device/dmath.cuh"
#include "../../include/ctl_b200.h"
#include <cfloat>

namespace ctld {

CTL_DEV float to_srgb_component(float v) { return v <= 0.0031308f ? 12.92f * v : 1.055f * powf(v, (float)(1.0 / 2.4)) - 0.055f; } // Math/Spectrum.cu:229-235
CTL_DEV unsigned to_u8(float x) { return (unsigned)(unsigned char)(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f); }                        // Float3ToCOLORREF, Math/Spectrum.h:521-526
CTL_DEV uchar4 gamma_rgba8(float r, float g, float b) { return make_uchar4((unsigned char)to_u8(to_srgb_component(r)), (unsigned char)to_u8(to_srgb_component(g)), (unsigned char)to_u8(to_srgb_component(b)), 255); }
CTL_DEV void px_to_spectrum(const float* __restrict__ p, float splat_scale, float c[3]) { // PixelData::toSpectrum, Engine/Image.h:20-27 (Spectrum / float = * reciprocal)
    const float ws = __ldg(p + 6), weight = ws != 0.0f ? ws : 1.0f, recip = 1.0f / weight;
#pragma unroll
    for (int k = 0; k < 3; k++) c[k] = __ldg(p + k) * recip + __ldg(p + 3 + k) * splat_scale;
}
CTL_DEV uchar4 to_rgbe(const float c[3]) { // Spectrum::toRGBE, Math/Spectrum.h:534-555
    const float mx = fmaxf(c[0], fmaxf(c[1], c[2]));
    if (mx < 1e-32f) return make_uchar4(0, 0, 0, 0);
    int e;
    const float scale = (float)frexp((double)mx, &e) * 256.0f / mx;
    return make_uchar4((unsigned char)(c[0] * scale), (unsigned char)(c[1] * scale), (unsigned char)(c[2] * scale), (unsigned char)(e + 128));
}
CTL_DEV void from_rgbe(uchar4 q, float c[3]) { // Math/Spectrum.h:557-565
    if (!q.w) { c[0] = c[1] = c[2] = 0.0f; return; }
    const float ex = ldexpf(1.0f, (int)q.w - (128 + 8));
    c[0] = (float)q.x * ex; c[1] = (float)q.y * ex; c[2] = (float)q.z * ex;
}

struct PipeFilter { int type; float xw, yw, p0, p1, ix, iy, expx, expy; };
CTL_DEV float mitchell1d(const PipeFilter& f, float x) { // SceneTypes/Filter.h:101-110
    const float B = f.p0, C = f.p1;
    x = fabsf(2.f * x);
    if (x > 1.f) return ((-B - 6 * C) * x * x * x + (6 * B + 30 * C) * x * x + (-12 * B - 48 * C) * x + (8 * B + 24 * C)) * (1.f / 6.f);
    return ((12 - 9 * B - 6 * C) * x * x * x + (-18 + 12 * B + 6 * C) * x * x + (6 - 2 * B)) * (1.f / 6.f);
}
CTL_DEV float sinc1d(const PipeFilter& f, float x) { // :133-141
    x = fabsf(x);
    if (x < 1e-5f) return 1.f;
    if (x > 1.f) return 0.f;
    x *= PI_F;
    const float sinc = sinf(x) / x, lanczos = sinf(x * f.p0) / (x * f.p0);
    return sinc * lanczos;
}
CTL_DEV float filter_eval(const PipeFilter& f, float x, float y) {
    switch (f.type) {
    case 0: return 1.0f;
    case 1: return fmaxf(0.0f, expf(-f.p0 * x * x) - f.expx) * fmaxf(0.0f, expf(-f.p0 * y * y) - f.expy);
    case 2: return fmaxf(0.0f, f.xw - fabsf(x)) * fmaxf(0.0f, f.yw - fabsf(y));
    case 3: return mitchell1d(f, x * f.ix) * mitchell1d(f, y * f.iy);
    default: return sinc1d(f, x * f.ix) * sinc1d(f, y * f.iy);
    }
}

// copySamplesToOutput (ImagePipeline.cu:14-21)
__global__ void __launch_bounds__(256) k_pipe_direct(const float* __restrict__ accum, int n_pixels, float splat_scale, uchar4* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        float c[3]; px_to_spectrum(accum + (size_t)i * 7, splat_scale, c);
        out[i] = gamma_rgba8(c[0], c[1], c[2]);
    }
}
// Stage 2: copySamplesToFiltered (FILTER = false, ImagePipeline.cu:23-30) or rtm_Copy / evalFilter (FILTER = true, CanonicalFilter.cu:6-36).
// TO_OUTPUT fuses copyFilteredToOutput (ImagePipeline.cu:32-41): the RGBE value is decoded and gamma-corrected in the same thread.
template <bool FILTER, bool TO_OUTPUT>
__global__ void __launch_bounds__(256) k_pipe_stage2(const float* __restrict__ accum, int w, int h, float splat_scale, const __grid_constant__ PipeFilter F, uchar4* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * h; i += gridDim.x * blockDim.x) {
        float c[3] = {0.0f, 0.0f, 0.0f};
        if (FILTER) {
            const int _x = i % w, _y = i / w;
            const int x0 = max(0, (int)ceilf((float)_x - F.xw)), x1 = min(w - 1, (int)floorf((float)_x + F.xw));
            const int y0 = max(0, (int)ceilf((float)_y - F.yw)), y1 = min(h - 1, (int)floorf((float)_y + F.yw));
            if ((x1 - x0) >= 0 && (y1 - y0) >= 0) {
                float acc[3] = {0.0f, 0.0f, 0.0f}, acc_w = 0.0f;
                for (int y = y0; y <= y1; ++y)
                    for (int x = x0; x <= x1; ++x) {
                        const float wt = filter_eval(F, (float)abs(x - _x), (float)abs(y - _y));
                        float pc[3]; px_to_spectrum(accum + ((size_t)y * w + x) * 7, splat_scale, pc);
                        for (int k = 0; k < 3; k++) acc[k] += pc[k] * wt;
                        acc_w += wt;
                    }
                const float recip = 1.0f / acc_w;
                for (int k = 0; k < 3; k++) c[k] = acc[k] * recip;
            }
        } else px_to_spectrum(accum + (size_t)i * 7, splat_scale, c);
        const uchar4 q = to_rgbe(c);
        if (TO_OUTPUT) { from_rgbe(q, c); out[i] = gamma_rgba8(c[0], c[1], c[2]); }
        else out[i] = q;
    }
}

// Image::ComputeLuminanceInfo (Engine/Image.cu:88-173).  The reference adds with float atomics (order = scheduler's); here the sums are taken
// in the fixed order of the CPU restatement: row-major inside each 16x16 pixel block (one CUDA block per pixel block, thread 0 adds the 256
// staged values), then block by block (k_lum_final) -- min / max / average are bit-identical to the oracle's.
__global__ void __launch_bounds__(256) k_lum_blocks(const uchar4* __restrict__ rgbe, int w, int h, int blocks_x, float4* __restrict__ partial) {
    __shared__ float sY[256], sL[256];
    const int bx = blockIdx.x % blocks_x, by = blockIdx.x / blocks_x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, x = bx * 16 + tx, y = by * 16 + ty;
    const bool in = x < w && y < h;
    float Y = 0.0f, lg = 0.0f;
    if (in) {
        float c[3]; from_rgbe(rgbe[(size_t)y * w + x], c);
        Y = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f; // Spectrum::getLuminance, Math/Spectrum.cu:174-177
        lg = logf(2.3e-5f + Y);
    }
    sY[threadIdx.x] = in ? Y : -1.0f; sL[threadIdx.x] = lg;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mn = FLT_MAX, mx = 0.0f, s = 0.0f, sl = 0.0f;
        for (int k = 0; k < 256; k++) { const float v = sY[k]; if (v >= 0.0f) { mn = fminf(mn, v); mx = fmaxf(mx, v); s += v; sl += sL[k]; } }
        partial[blockIdx.x] = make_float4(mn, mx, s, sl);
    }
}
// lum: [0] min [1] max [2] avg [3] exp(avg log) [4] scale [5] invWp2   (ToneMapPostProcess::Apply, ToneMapPostProcess.cu:27-39)
__global__ void k_lum_final(const float4* __restrict__ partial, int n_blocks, int n_pixels, float key, float burn_param, float* __restrict__ lum) {
    if (threadIdx.x || blockIdx.x) return;
    float mn = FLT_MAX, mx = 0.0f, s = 0.0f, sl = 0.0f;
    for (int b = 0; b < n_blocks; b++) { const float4 p = partial[b]; mn = fminf(mn, p.x); mx = fmaxf(mx, p.y); s += p.z; sl += p.w; }
    const float logAvg = expf(sl / (float)n_pixels);
    const float scale = key / logAvg, Lwhite = mx * scale;
    const float burn = fminf(1.0f, fmaxf(1e-8f, 1.0f - burn_param));
    lum[0] = mn; lum[1] = mx; lum[2] = s / (float)n_pixels; lum[3] = logAvg; lum[4] = scale; lum[5] = 1 / (Lwhite * Lwhite * powf(burn, 4.0f));
}
// Reinhard05Kernel (ToneMapPostProcess.cu:6-25) + applyGammaCorrectureToOutput (ImagePipeline.cu:43-52), both quantisations kept
__global__ void __launch_bounds__(256) k_reinhard(const uchar4* __restrict__ rgbe, int n_pixels, const float* __restrict__ lum, uchar4* __restrict__ out) {
    const float scale = lum[4], invWp2 = lum[5];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        float c[3]; from_rgbe(rgbe[i], c);
        float X = c[0] * 0.412453f + c[1] * 0.357580f + c[2] * 0.180423f, Y = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f, Z = c[0] * 0.019334f + c[1] * 0.119193f + c[2] * 0.950227f;
        const float sxyz = fminf(fmaxf(X + Y + Z, 0.001f), 100000.0f); // toYxy, Math/Spectrum.cu:286-294
        const float x = X / sxyz, y = Y / sxyz;
        const float Lp = scale * Y;
        Y = Lp * (1.0f + Lp * invWp2) / (1.0f + Lp);
        const float yc = fminf(fmaxf(y, 0.001f), 100000.0f);           // fromYxy, :296-302
        X = Y / yc * x; Z = Y / yc * (1 - x - y);
        const float r = 3.240479f * X + -1.537150f * Y + -0.498535f * Z, g = -0.969256f * X + 1.875991f * Y + 0.041556f * Z, b = 0.055648f * X + -0.204043f * Y + 1.057311f * Z;
        const float lr = (float)to_u8(r) / 255.0f, lg = (float)to_u8(g) / 255.0f, lb = (float)to_u8(b) / 255.0f; // toRGBCOL -> fromRGBCOL
        out[i] = gamma_rgba8(lr, lg, lb);
    }
}

// PixelVarianceInfo::updateMoments on every pixel (uniform block sampler: samplerPerformed = 1)
__global__ void __launch_bounds__(256) k_variance_update(ctl_pixel_variance_info* __restrict__ var, const float* __restrict__ accum, int n_pixels, float splat_scale) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        ctl_pixel_variance_info V = var[i];
        const float* p = accum + (size_t)i * 7;
        const float recip = 1.0f / 1.0f;
        float est[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { const float nps = p[k] + p[3 + k] * splat_scale; est[k] = (nps - V.prev_I[k]) * recip; V.prev_I[k] = nps; }
        V.weight = p[6];
        if (V.iterations_done++ % 2 == 1) { V.half_buffer[0] += est[0]; V.half_buffer[1] += est[1]; V.half_buffer[2] += est[2]; }
        const float Y = est[0] * 0.212671f + est[1] * 0.715160f + est[2] * 0.072169f;
        V.sum_x += Y; V.sum_x2 += Y * Y; V.num_samples_var++;
        var[i] = V;
    }
}

} // namespace ctld
