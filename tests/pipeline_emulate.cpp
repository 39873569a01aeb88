// tests/pipeline_emulate.cpp -- TEST INFRASTRUCTURE (see tests/nlm_emulate.cpp for the method): the image-pipeline CUDA kernels
// (cudatracerlib_b200/csrc/image_pipeline.cuh), the same source text, compiled for the host and run block by block in the launch sequence of
// ctl_apply_image_pipeline (csrc/ctl_api.cu).  With the host's libm behind expf / sinf / powf / logf the kernels must reproduce the goldens minted from
// the reference's own code byte for byte; on the device only libdevice's last-ulp differences remain (tests/test_gpu_image_pipeline.py).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <vector>
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static inline void __syncthreads() {}
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::max; using std::min;
#include "../cudatracerlib_b200/csrc/image_pipeline.cuh"

template <typename K> static void launch(dim3 grid, dim3 block, int sweeps, K kernel) {
    gridDim = grid; blockDim = block;
    for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx = make_uint3(bx, by, 0);
        for (int s = 0; s < sweeps; s++)
            for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) { threadIdx = make_uint3(tx, ty, 0); kernel(); }
    }
}

extern "C" void emu_apply_image_pipeline(const float* accum, int w, int h, float splat_scale, const ctl_image_pipeline* P, unsigned char* rgba8, float* lum_info) {
    using namespace ctld;
    const int n = w * h;
    uchar4* dst = (uchar4*)rgba8;
    const dim3 grid(5), block(256);
    PipeFilter F = {P->filter_type, P->x_width, P->y_width, P->param0, P->param1, 1.f / P->x_width, 1.f / P->y_width,
                    expf(-P->param0 * P->x_width * P->x_width), expf(-P->param0 * P->y_width * P->y_width)};
    if (P->filter_type < 0 && !P->tonemap) { launch(grid, block, 1, [&]() { k_pipe_direct(accum, n, splat_scale, dst); }); return; }
    if (!P->tonemap) { launch(grid, block, 1, [&]() { k_pipe_stage2<true, true>(accum, w, h, splat_scale, F, dst); }); return; }
    const int bx = (w + 15) / 16, by = (h + 15) / 16;
    std::vector<uchar4> rgbe(n); std::vector<float4> partial((size_t)bx * by); float lum[8];
    if (P->filter_type >= 0) launch(grid, block, 1, [&]() { k_pipe_stage2<true, false>(accum, w, h, splat_scale, F, rgbe.data()); });
    else launch(grid, block, 1, [&]() { k_pipe_stage2<false, false>(accum, w, h, splat_scale, F, rgbe.data()); });
    launch(dim3(bx * by), block, 2, [&]() { k_lum_blocks(rgbe.data(), w, h, bx, partial.data()); });
    launch(dim3(1), dim3(32), 1, [&]() { k_lum_final(partial.data(), bx * by, n, P->key, P->burn, lum); });
    launch(grid, block, 1, [&]() { k_reinhard(rgbe.data(), n, lum, dst); });
    if (lum_info) memcpy(lum_info, lum, 6 * sizeof(float));
}

extern "C" void emu_variance_update(ctl_pixel_variance_info* var, const float* accum, int n, float splat_scale) {
    launch(dim3(3), dim3(256), 1, [&]() { ctld::k_variance_update(var, accum, n, splat_scale); });
}
