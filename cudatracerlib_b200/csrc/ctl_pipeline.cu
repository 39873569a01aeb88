// ctl_pipeline.cu -- image pipeline entry points of the C ABI: reconstruction filters, tone mapping, NonLocalMeansFilter,
// PixelVarianceBuffer (kernels: csrc/image_pipeline.cuh, csrc/nlm_filter.cuh).
#include "ctl_internal.h"
#include "image_pipeline.cuh"
#include "nlm_filter.cuh"

extern "C" {

// == applyImagePipeline (Kernel/ImagePipeline/ImagePipeline.cu:54-84): optional reconstruction filter, optional tone mapper, gamma
int ctl_apply_image_pipeline(ctl_ctx* c, float splat_scale, const ctl_image_pipeline* P, void* d_rgba8, void* host_rgba8, float lum_info[6]) {
    if (!c || !P || (!d_rgba8 && !host_rgba8)) return set_err("null argument");
    if (P->filter_type < -1 || P->filter_type > 5) return set_err("filter_type must be -1 (none), 0 (box), 1 (Gaussian), 2 (triangle), 3 (Mitchell), 4 (Lanczos-sinc) or 5 (non-local means)");
    const bool nlm = P->filter_type == 5;
    if (P->filter_type >= 0 && !nlm && (!(P->x_width > 0) || !(P->y_width > 0) || P->x_width > 16 || P->y_width > 16)) return set_err("filter widths out of range (0, 16]");
    if (P->tonemap < 0 || P->tonemap > 1) return set_err("tonemap must be 0 (none) or 1 (Reinhard05)");
    if (nlm) {
        if (!(P->param0 >= 0.0f) || !(P->param1 >= 0.0f)) return set_err("NonLocalMeansFilter: k (param0) and sigma2Scale (param1) must be >= 0");
        if (!(P->x_width >= 1.0f) || P->x_width > 1e9f || P->x_width != floorf(P->x_width)) return set_err("NonLocalMeansFilter: UpdateWeightPeriodicity (x_width) must be an integer >= 1");
        if (!c->variance_buffer || !c->d_var.p || c->d_var.n < (size_t)c->w * c->h) return set_err("NonLocalMeansFilter reads the PixelVarianceBuffer: ctl_set_param_i(ctx, \"PixelVarianceBuffer\", 1) before the passes");
    }
    CK(cudaSetDevice(c->device));
    const int n = c->w * c->h;
    uchar4* dst = (uchar4*)d_rgba8;
    if (!dst) { CK(c->resolve_tmp.ensure((size_t)n)); dst = c->resolve_tmp.p; }
    const int grid = grid_for(c, 8);
    if (nlm) { // NonLocalMeansFilter::Apply (Kernel/ImagePipeline/Filter/NonLocalMeansFilter.cu:184-228), then the rest of applyImagePipeline (ImagePipeline.cu:71-82)
        const long long numPasses = (long long)c->passes_done; const int n_update = (int)P->x_width;
        bool force_update = false;
        if (c->nlm_pixels != (size_t)n) { // Resize / first use: new cache and weight buffer (adaptBuffer), last_iter_weight_update = -1
            CK(c->nlm_cached.ensure((size_t)n)); CK(c->nlm_varh.ensure((size_t)n)); CK(c->nlm_weights.ensure((size_t)n * NLM_NW));
            c->nlm_pixels = (size_t)n; c->nlm_last_update = -1; force_update = true;
        }
        k_nlm_prepare<<<grid, 256, 0, c->stream>>>(c->accum, c->d_var.p, n, splat_scale, c->nlm_cached.p, c->nlm_varh.p);
        const dim3 nb((c->w + NLM_B - 1) / NLM_B, (c->h + NLM_B - 1) / NLM_B), nt(NLM_B, NLM_B);
        if (c->nlm_last_update + 1 != numPasses || (numPasses % n_update) == 0 || force_update) {
            CK(cudaMemsetAsync(c->nlm_weights.p, 0, (size_t)n * NLM_NW * sizeof(float), c->stream));   // m_weightBuffer.ClearBuffer()
            k_nlm_weights<<<nb, nt, 0, c->stream>>>(c->nlm_cached.p, c->nlm_varh.p, c->w, c->h, P->param0, P->param1, c->nlm_weights.p);
        }
        c->nlm_last_update = numPasses;
        if (!P->tonemap) k_nlm_apply<true><<<nb, nt, 0, c->stream>>>(c->nlm_cached.p, c->nlm_weights.p, c->w, c->h, dst);
        else {
            const int bx = (c->w + 15) / 16, by = (c->h + 15) / 16;
            CK(c->pipe_rgbe.ensure((size_t)n)); CK(c->pipe_partial.ensure((size_t)bx * by)); CK(c->pipe_lum.ensure(8));
            k_nlm_apply<false><<<nb, nt, 0, c->stream>>>(c->nlm_cached.p, c->nlm_weights.p, c->w, c->h, c->pipe_rgbe.p);
            k_lum_blocks<<<bx * by, 256, 0, c->stream>>>(c->pipe_rgbe.p, c->w, c->h, bx, c->pipe_partial.p);
            k_lum_final<<<1, 32, 0, c->stream>>>(c->pipe_partial.p, bx * by, n, P->key, P->burn, c->pipe_lum.p);
            k_reinhard<<<grid, 256, 0, c->stream>>>(c->pipe_rgbe.p, n, c->pipe_lum.p, dst);
            if (lum_info) { CK(cudaMemcpyAsync(lum_info, c->pipe_lum.p, 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
        }
        CK(cudaGetLastError());
        if (host_rgba8) { CK(cudaMemcpyAsync(host_rgba8, dst, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
        return 0;
    }
    PipeFilter F = {P->filter_type, P->x_width, P->y_width, P->param0, P->param1, 1.f / P->x_width, 1.f / P->y_width,
                    expf(-P->param0 * P->x_width * P->x_width), expf(-P->param0 * P->y_width * P->y_width)}; // FilterBase / GaussianFilter ctor, SceneTypes/Filter.h:15-19, 60-66
    if (P->filter_type < 0 && !P->tonemap) k_pipe_direct<<<grid, 256, 0, c->stream>>>(c->accum, n, splat_scale, dst);
    else if (!P->tonemap) k_pipe_stage2<true, true><<<grid, 256, 0, c->stream>>>(c->accum, c->w, c->h, splat_scale, F, dst);
    else {
        const int bx = (c->w + 15) / 16, by = (c->h + 15) / 16;
        CK(c->pipe_rgbe.ensure((size_t)n)); CK(c->pipe_partial.ensure((size_t)bx * by)); CK(c->pipe_lum.ensure(8));
        if (P->filter_type >= 0) k_pipe_stage2<true, false><<<grid, 256, 0, c->stream>>>(c->accum, c->w, c->h, splat_scale, F, c->pipe_rgbe.p);
        else k_pipe_stage2<false, false><<<grid, 256, 0, c->stream>>>(c->accum, c->w, c->h, splat_scale, F, c->pipe_rgbe.p);
        k_lum_blocks<<<bx * by, 256, 0, c->stream>>>(c->pipe_rgbe.p, c->w, c->h, bx, c->pipe_partial.p);
        k_lum_final<<<1, 32, 0, c->stream>>>(c->pipe_partial.p, bx * by, n, P->key, P->burn, c->pipe_lum.p);
        k_reinhard<<<grid, 256, 0, c->stream>>>(c->pipe_rgbe.p, n, c->pipe_lum.p, dst);
        if (lum_info) { CK(cudaMemcpyAsync(lum_info, c->pipe_lum.p, 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
    }
    CK(cudaGetLastError());
    if (host_rgba8) { CK(cudaMemcpyAsync(host_rgba8, dst, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
    return 0;
}
// NonLocalMeansFilter's weight buffer of the last ctl_apply_image_pipeline(filter_type 5), in the reference's layout [pixel][169] (slot (yo + 6) * 13 + xo + 6,
// NonLocalMeansFilter.h:13-31, 63-66); kept on the device as [169][pixel]
int ctl_read_nlm_weights(ctl_ctx* c, float* host_out) {
    if (!c || !host_out) return set_err("null argument");
    if (!c->nlm_pixels || c->nlm_pixels != (size_t)c->w * c->h || c->nlm_last_update < 0) return set_err("no NonLocalMeansFilter weights: apply a pipeline with filter_type 5 first");
    CK(cudaSetDevice(c->device));
    const size_t n = c->nlm_pixels;
    std::vector<float> soa(n * NLM_NW);
    CK(cudaMemcpyAsync(soa.data(), c->nlm_weights.p, n * NLM_NW * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t p = 0; p < n; p++) for (int s = 0; s < NLM_NW; s++) host_out[p * NLM_NW + s] = soa[(size_t)s * n + p];
    return 0;
}
// == applyImagePipeline(tracer, img, 0, 0): the default resolve
int ctl_resolve_srgb8(ctl_ctx* c, float splat_scale, void* d_rgba8, void* host_rgba8) {
    ctl_image_pipeline P; memset(&P, 0, sizeof(P)); P.filter_type = -1;
    return ctl_apply_image_pipeline(c, splat_scale, &P, d_rgba8, host_rgba8, nullptr);
}
// == applyImagePipeline(tracer, img, filter, 0) with box / Gaussian / triangle (kept for callers of the first f3 slice)
int ctl_resolve_filtered_srgb8(ctl_ctx* c, float splat_scale, int filter_type, float x_width, float y_width, float alpha, void* d_rgba8, void* host_rgba8) {
    if (filter_type < 0 || filter_type > 2) return set_err("filter_type must be 0 (box), 1 (Gaussian) or 2 (triangle)");
    ctl_image_pipeline P; memset(&P, 0, sizeof(P)); P.filter_type = filter_type; P.x_width = x_width; P.y_width = y_width; P.param0 = alpha;
    return ctl_apply_image_pipeline(c, splat_scale, &P, d_rgba8, host_rgba8, nullptr);
}
// == PixelVarianceBuffer (Kernel/PixelVarianceBuffer.h)
} // extern "C"
int ctl_variance_after_pass(ctl_ctx* c, bool new_trace) { // Tracer<true>::DoPass: Clear on a new trace (Tracer.h:222-226), AddPass after DoRender (:233-237)
    if (!c->variance_buffer) return 0;
    const size_t n = (size_t)c->w * c->h;
    const bool fresh = c->d_var.n < n;
    CK(c->d_var.ensure(n));
    if (new_trace || fresh) CK(cudaMemsetAsync(c->d_var.p, 0, n * sizeof(ctl_pixel_variance_info), c->stream));
    k_variance_update<<<grid_for(c, 8), 256, 0, c->stream>>>(c->d_var.p, c->accum, (int)n, 0.0f /* getSplatScale(): the path tracers never splat */);
    CK(cudaGetLastError());
    return 0;
}
extern "C" {
int ctl_read_variance(ctl_ctx* c, ctl_pixel_variance_info* out) {
    if (!c || !out) return set_err("null argument");
    if (!c->variance_buffer || !c->d_var.p) return set_err("PixelVarianceBuffer is off: ctl_set_param_i(ctx, \"PixelVarianceBuffer\", 1) before the passes");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->d_var.p, (size_t)c->w * c->h * sizeof(ctl_pixel_variance_info), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
} // extern "C"
