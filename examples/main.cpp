// examples/main.cpp -- the reference's example driver (main.cpp:135-180) on the B200 backend, in C++ over the C ABI:
//   tracer.Resize -> InitializeScene -> n x DoPass -> applyImagePipeline(BoxFilter(0.5, 0.5)) -> write image.
// Like the reference's main, arguments are recognised by what they are, in any order: a number = passes, "PT" / "PT_Wave" = tracer
// (PathTracer / WavefrontPathTracer; the reference also offers direct, BDPT, PPPM), an existing file = mesh to import (.obj / .ply / .xmsh; the
// reference takes a Mitsuba scene file there, whose loader is outside the hot path), WxH = resolution, "tonemap" = ToneMapPostProcess, "nlm" = NonLocalMeansFilter instead of the box filter, *.ppm = output,
// a scene name (cornell, cornell7, c2, c3, c4, c5, soup) = one of the built-in synthetic scenes.
// Build:  g++ -std=c++17 -O2 examples/main.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_render
// Usage:  examples/ctl_render cornell 64 PT 512x512 result.ppm        examples/ctl_render tests/golden/obj/room.obj 32 PT_Wave tonemap
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "b200_path_tracer.hpp"

int main(int ac, char** av) {
    const char* kinds[] = {"cornell", "cornell7", "c2", "c3", "c4", "c5", "soup"};
    int kind = 0, n_passes = 64, width = 512, height = 512; bool wave = false, tonemap = false, nlm = false;
    std::string out = "result.ppm", mesh_file;
    for (int i = 1; i < ac; i++) {
        const std::string a = av[i];
        int k = -1; for (int j = 0; j < 7; j++) if (a == kinds[j]) k = j;
        unsigned w2 = 0, h2 = 0;
        if (k >= 0) kind = k;
        else if (a == "PT") wave = false;
        else if (a == "PT_Wave") wave = true;
        else if (a == "tonemap") tonemap = true;
        else if (a == "nlm") nlm = true;
        else if (sscanf(a.c_str(), "%ux%u", &w2, &h2) == 2) { width = (int)w2; height = (int)h2; }
        else if (a.size() > 4 && a.substr(a.size() - 4) == ".ppm") out = a;
        else if (!a.empty() && a.find_first_not_of("0123456789") == std::string::npos) n_passes = atoi(a.c_str());
        else if (FILE* f = fopen(a.c_str(), "rb")) { fclose(f); mesh_file = a; }
        else { fprintf(stderr, "accepts: passes, tracer {PT, PT_Wave}, scene name or mesh file, WxH, tonemap, nlm, out.ppm\n%s could not be used, exiting now\n", a.c_str()); return 2; }
    }
    try {
        std::unique_ptr<ctlb200::Scene> scene;
        if (mesh_file.empty()) scene.reset(new ctlb200::Scene(kind, width, height));
        else { // look at the mesh's bounding box from the -z side
            const float o[3] = {0, 0, -1}, t[3] = {0, 0, 0}, up[3] = {0, 1, 0};
            ctlb200::Scene probe({mesh_file}, o, t, up, 60.0f, width, height);
            const ctl_scene_view& v = probe.view();
            const float c[3] = {0.5f * (v.box_min[0] + v.box_max[0]), 0.5f * (v.box_min[1] + v.box_max[1]), 0.5f * (v.box_min[2] + v.box_max[2])};
            const float pos[3] = {c[0], c[1], v.box_min[2] + 0.02f * (v.box_max[2] - v.box_min[2])};   // just inside the front face (rooms) / at it (objects)
            scene.reset(new ctlb200::Scene({mesh_file}, pos, c, up, 75.0f, width, height));
        }
        ctlb200::PathTracer pt; ctlb200::WavefrontPathTracer wpt;       // == options.tracer: PT / PT_Wave
        auto run = [&](auto& tracer) {
            tracer.setParameter("MaxPathLength", 8);
            if (nlm) tracer.setParameter("PixelVarianceBuffer", 1);   // the filter reads the per-pixel variance of the passes
            tracer.Resize(width, height);
            tracer.InitializeScene(scene->view());
            for (int i = 0; i < n_passes; i++) {
                tracer.DoPass(nullptr, i == 0);
                printf("\r%3d%%", (i + 1) * 100 / n_passes); fflush(stdout);
            }
            std::vector<unsigned char> rgba((size_t)width * height * 4);
            // applyImagePipeline(*tracer, outImage, BoxFilter(0.5f, 0.5f))  (main.cpp:172), optionally with a ToneMapPostProcess behind the filter
            ctl_image_pipeline P; memset(&P, 0, sizeof(P)); P.filter_type = 0; P.x_width = P.y_width = 0.5f; P.tonemap = tonemap ? 1 : 0; P.key = 0.18f;
            if (nlm) { P.filter_type = 5; P.x_width = 25.0f; P.param0 = 0.45f; P.param1 = 1.0f; }   // NonLocalMeansFilter: UpdateWeightPeriodicity, k, sigma2Scale (reference default 0.005 only filters well-converged frames)
            ctlb200::check(ctl_apply_image_pipeline(tracer.handle(), tracer.getSplatScale(), &P, nullptr, rgba.data(), nullptr));
            FILE* f = fopen(out.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot write " + out);
            fprintf(f, "P6\n%d %d\n255\n", width, height);
            for (size_t i = 0; i < (size_t)width * height; i++) fwrite(&rgba[4 * i], 1, 3, f);
            fclose(f);
            printf("\n%s: %s, %d passes, %llu rays in the last pass, %.3f s, %.1f Mrays/s\n", out.c_str(), wave ? "PT_Wave" : "PT", tracer.getNumPassesDone(), tracer.getRaysInLastPass(),
                   tracer.getLastTimeSpentRenderingSec(), tracer.getRaysInLastPass() / tracer.getLastTimeSpentRenderingSec() / 1e6);
        };
        if (wave) run(wpt); else run(pt);
    } catch (const std::runtime_error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
