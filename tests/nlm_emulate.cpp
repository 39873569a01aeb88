// tests/nlm_emulate.cpp -- TEST INFRASTRUCTURE.  The NonLocalMeansFilter CUDA kernels (cudatracerlib_b200/csrc/nlm_filter.cuh), the same source text, compiled
// for the HOST with g++ and run block by block, thread by thread, so that their indexing and arithmetic can be checked against the oracle where no GPU
// exists (tests/test_golden_cpu.py).  __global__ / __device__ are ignored attributes for g++ (cuda_runtime.h's host_defines.h), __shared__ becomes a static
// array, threadIdx / blockIdx / blockDim / gridDim are globals set by the launcher below.  Kernels with a __syncthreads are swept twice per block (first
// sweep completes the tile, second computes from it; their outputs are plain overwrites).  Not a product path: the library never runs these on the CPU.
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
#include "../cudatracerlib_b200/csrc/nlm_filter.cuh"

template <typename K> static void launch(dim3 grid, dim3 block, int sweeps, K kernel) {
    gridDim = grid; blockDim = block;
    for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx = make_uint3(bx, by, 0);
        for (int s = 0; s < sweeps; s++)
            for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) { threadIdx = make_uint3(tx, ty, 0); kernel(); }
    }
}

// Same sequence as ctl_apply_image_pipeline's filter_type 5 branch (csrc/ctl_api.cu).  weights: [169][w*h] (device layout), in/out.
extern "C" void emu_nlm_filter(const float* accum, const ctl_pixel_variance_info* var, int w, int h, float splat_scale, float k, float sigma2Scale,
                               float* weights, int compute_weights, unsigned char* stage_rgbe, unsigned char* out_rgba8) {
    using namespace ctld;
    const int n = w * h;
    std::vector<uchar4> cached(n); std::vector<float> varh(n);
    launch(dim3(3), dim3(256), 1, [&]() { k_nlm_prepare(accum, var, n, splat_scale, cached.data(), varh.data()); });
    const dim3 nb((w + NLM_B - 1) / NLM_B, (h + NLM_B - 1) / NLM_B), nt(NLM_B, NLM_B);
    if (compute_weights) {
        memset(weights, 0, (size_t)n * NLM_NW * sizeof(float));
        launch(nb, nt, 2, [&]() { k_nlm_weights(cached.data(), varh.data(), w, h, k, sigma2Scale, weights); });
    }
    launch(nb, nt, 2, [&]() { k_nlm_apply<false>(cached.data(), weights, w, h, (uchar4*)stage_rgbe); });
    launch(nb, nt, 2, [&]() { k_nlm_apply<true>(cached.data(), weights, w, h, (uchar4*)out_rgba8); });
}
