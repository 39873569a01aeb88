// examples/multi_gpu.cpp -- the north star's multi-GPU step without Python: C++ host code over the C ABI, one process driving N GPUs of one node,
// image tiled across them, a single NCCL reduce of the Spectrum accumulation buffer per frame (csrc/ctl_comm.cu; NCCL is loaded at run time).
// Build:  g++ -std=c++17 -O2 examples/multi_gpu.cpp -Iinclude -Lcudatracerlib_b200 -lctl_b200 -Wl,-rpath,$PWD/cudatracerlib_b200 -o examples/ctl_multi_gpu
// Usage:  examples/ctl_multi_gpu [scene=c4] [gpus=all] [frames=5] [spp=8] [WxH=1920x1080] [inflight=1] [check]
//         inflight=L > 1: the frames as a pipeline with L frames in flight (SubmitFrame / AcquireFrame: ctl_comm_submit_frame_all, ctl_acquire_frame).
//         `check` also renders the frame on device 0 alone and compares the images (same paths => equal up to float summation order).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "b200_path_tracer.hpp"

int main(int ac, char** av) {
    const char* kinds[] = {"cornell", "cornell7", "c2", "c3", "c4", "c5", "soup"};
    int kind = 4, gpus = 0, frames = 5, spp = 8, width = 1920, height = 1080, depth = 8, inflight = 1; bool check_single = false;
    for (int i = 1; i < ac; i++) {
        const std::string a = av[i]; unsigned w2 = 0, h2 = 0; int k = -1;
        for (int j = 0; j < 7; j++) if (a == kinds[j]) k = j;
        if (k >= 0) kind = k;
        else if (a == "check") check_single = true;
        else if (sscanf(a.c_str(), "%ux%u", &w2, &h2) == 2) { width = (int)w2; height = (int)h2; }
        else if (a.rfind("gpus=", 0) == 0) gpus = atoi(a.c_str() + 5);
        else if (a.rfind("frames=", 0) == 0) frames = atoi(a.c_str() + 7);
        else if (a.rfind("spp=", 0) == 0) spp = atoi(a.c_str() + 4);
        else if (a.rfind("depth=", 0) == 0) depth = atoi(a.c_str() + 6);
        else if (a.rfind("inflight=", 0) == 0) inflight = atoi(a.c_str() + 9);
        else { fprintf(stderr, "accepts: scene name, gpus=N, frames=N, spp=N, depth=N, inflight=N, WxH, check\n"); return 2; }
    }
    try {
        if (gpus <= 0) { // all devices: probe by creating contexts until it fails
            for (gpus = 0; gpus < 64; gpus++) { ctl_ctx* c = ctl_create(gpus, 16, 16); if (!c) break; ctl_destroy(c); }
            if (!gpus) { fprintf(stderr, "no CUDA device: %s\n", ctl_last_error()); return 1; }
        }
        const int batch = spp % 8 == 0 ? 8 : spp;
        ctlb200::Scene scene(kind, width, height);
        ctlb200::MultiGpuPathTracer mt(gpus);
        mt.setParameter("MaxPathLength", depth);
        mt.Resize(width, height);
        mt.InitializeScene(scene.view());
        std::vector<ctl_pixel_data> img((size_t)width * height);
        unsigned long long rays = mt.RenderFrame(spp, batch, nullptr);   // warm-up
        if (inflight > 1) {   // warm-up of the pipeline's lanes (their buffers are allocated on first use)
            mt.setParameter("FramesInFlight", inflight);
            for (int f = 0; f < inflight; f++) mt.SubmitFrame(spp, batch);
            while (mt.FramesInFlight()) mt.AcquireFrame();
            for (int d = 0; d < gpus; d++) mt.device(d).Synchronize();
        }
        const auto t0 = std::chrono::steady_clock::now();
        if (inflight > 1) {   // the pipeline: frame f is submitted while frames f-1 .. f-inflight+1 render; images come back in order
            int got = 0;
            for (int f = 0; f < frames; f++) {
                mt.SubmitFrame(spp, batch);
                if (f >= inflight - 1) { got++; mt.AcquireFrame(got == frames ? img.data() : nullptr); }
            }
            while (mt.FramesInFlight()) { got++; mt.AcquireFrame(got == frames ? img.data() : nullptr); }
            for (int d = 0; d < gpus; d++) mt.device(d).Synchronize();
        } else
        for (int f = 0; f < frames; f++) rays = mt.RenderFrame(spp, batch, f + 1 == frames ? img.data() : nullptr);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double mean = 0; for (auto& p : img) mean += (p.rgb[0] + p.rgb[1] + p.rgb[2]) / (3.0 * (p.weight_sum > 0 ? p.weight_sum : 1));
        mean /= (double)img.size();
        printf("{\"scene\": \"%s\", \"gpus\": %d, \"frames\": %d, \"spp\": %d, \"width\": %d, \"height\": %d, \"rays_per_frame\": %llu, \"ms_per_frame\": %.3f, \"mrays_s\": %.1f, \"image_mean\": %.6f, \"frames_in_flight\": %d",
               kinds[kind], gpus, frames, spp, width, height, rays, 1e3 * s / frames, rays * (double)frames / s / 1e6, mean, inflight);
        if (check_single) {
            ctlb200::PathTracer one(0);
            one.setParameter("MaxPathLength", depth); one.Resize(width, height); one.InitializeScene(scene.view());
            for (int p = 0; p < spp; p += batch) one.DoPassesTiled(batch, p == 0);
            std::vector<ctl_pixel_data> ref((size_t)width * height);
            one.Synchronize();
            if (ctl_read_accum(one.handle(), ref.data())) throw std::runtime_error(ctl_last_error());
            double worst = 0; size_t weights_differ = 0;
            for (size_t i = 0; i < ref.size(); i++) {
                if (ref[i].weight_sum != img[i].weight_sum) weights_differ++;
                for (int k = 0; k < 3; k++) { const double d = std::fabs((double)ref[i].rgb[k] - img[i].rgb[k]) / (std::fabs((double)ref[i].rgb[k]) + 1e-3); if (d > worst) worst = d; }
            }
            printf(", \"single_gpu_rays\": %llu, \"weights_differ\": %zu, \"worst_rel_diff\": %.3g", one.getAccRays(), weights_differ, worst);
            if (weights_differ || worst > 1e-4) { printf("}\n"); fprintf(stderr, "multi-GPU image differs from the single-GPU image\n"); return 1; }
        }
        printf("}\n");
    } catch (const std::exception& e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
