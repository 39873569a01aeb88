// Compiled and run by tests/test_abi_cpu.py::test_cpp_adapter_compiles_and_fails_loudly_without_gpu and, on the GPU box,
// by tests/test_gpu_parity.py::test_cpp_adapter_renders (argv[1] = "gpu").
#include <cstdio>
#include <cstring>
#include "../include/b200_path_tracer.hpp"

int main(int argc, char** argv) {
    const bool want_gpu = argc > 1 && !strcmp(argv[1], "gpu");
    ctlb200::Scene scene(/*cornell*/ 0, 64, 64);
    if (scene.view().n_tri_data != 32) { printf("FAIL scene\n"); return 2; }
    ctlb200::PathTracer tracer;
    tracer.setParameter("MaxPathLength", 8);
    try {
        tracer.Resize(64, 64);
    } catch (const std::runtime_error& e) {
        printf("no device: %s\n", e.what());
        return want_gpu ? 3 : 0; // without a GPU the adapter must throw (no CPU fallback)
    }
    tracer.InitializeScene(scene.view());
    std::vector<ctl_pixel_data> img(64 * 64);
    tracer.DoPass(img.data(), true);
    double sum = 0, wsum = 0;
    for (auto& p : img) { sum += p.rgb[0] + p.rgb[1] + p.rgb[2]; wsum += p.weight_sum; }
    printf("passes %u rays %llu mean %.6f weight %.0f\n", tracer.getNumPassesDone(), tracer.getRaysInLastPass(), sum / (3.0 * 64 * 64), wsum);
    try { tracer.setParameter("NoSuchKey", 1); printf("FAIL no throw\n"); return 4; } catch (const std::runtime_error&) {}
    return (wsum == 64 * 64 && sum > 0) ? 0 : 5;
}
