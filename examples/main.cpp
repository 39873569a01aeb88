// examples/main.cpp -- the reference's example driver (main.cpp:135-180) on the B200 backend, in C++ over the C ABI:
//   tracer.Resize -> InitializeScene -> n x DoPass -> applyImagePipeline(BoxFilter(0.5, 0.5)) -> write image.
// Build:  g++ -std=c++17 -O2 examples/main.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_render
// Usage:  examples/ctl_render [scene kind 0..6 = cornell, cornell7, c2, c3, c4, c5, soup] [n_passes] [width] [height] [out.ppm]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "b200_path_tracer.hpp"

int main(int ac, char** av) {
    const int kind = ac > 1 ? atoi(av[1]) : 0, n_passes = ac > 2 ? atoi(av[2]) : 64;
    const int width = ac > 3 ? atoi(av[3]) : 512, height = ac > 4 ? atoi(av[4]) : 512;
    const std::string out = ac > 5 ? av[5] : "result.ppm";
    try {
        ctlb200::Scene scene(kind, width, height);
        ctlb200::PathTracer tracer;                         // == options.tracer (PathTracer)
        tracer.setParameter("MaxPathLength", 8);
        tracer.Resize(width, height);
        tracer.InitializeScene(scene.view());
        for (int i = 0; i < n_passes; i++) {
            tracer.DoPass(nullptr, i == 0);
            printf("\r%3d%%", (i + 1) * 100 / n_passes); fflush(stdout);
        }
        std::vector<unsigned char> rgba((size_t)width * height * 4);
        // applyImagePipeline(*tracer, outImage, BoxFilter(0.5f, 0.5f))  (main.cpp:172)
        ctlb200::check(ctl_resolve_filtered_srgb8(tracer.handle(), tracer.getSplatScale(), 0, 0.5f, 0.5f, 0.0f, nullptr, rgba.data()));
        FILE* f = fopen(out.c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 2; }
        fprintf(f, "P6\n%d %d\n255\n", width, height);
        for (size_t i = 0; i < (size_t)width * height; i++) fwrite(&rgba[4 * i], 1, 3, f);
        fclose(f);
        printf("\n%s: %d passes, %llu rays in the last pass, %.3f s, %.1f Mrays/s\n", out.c_str(), tracer.getNumPassesDone(), tracer.getRaysInLastPass(),
               tracer.getLastTimeSpentRenderingSec(), tracer.getRaysInLastPass() / tracer.getLastTimeSpentRenderingSec() / 1e6);
    } catch (const std::runtime_error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
