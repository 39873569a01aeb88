// b200_path_tracer.hpp -- header-only C++ adapter over the C ABI (ctl_b200.h), shaped like the reference's
// `PathTracer : Tracer<true>` (Integrators/PathTracer.h:7-24, Kernel/Tracer.h:67-294) so that an application written
// against CudaTracerLib's tracer interface can switch this path to the B200 backend:
//
//     reference                                   here
//     ---------                                   ----
//     PathTracer tracer;                          ctlb200::PathTracer tracer;
//     tracer.Resize(w, h);                        tracer.Resize(w, h);
//     tracer.InitializeScene(&scene);             tracer.InitializeScene(view);      // flat ctl_scene_view, see INTEGRATION.md
//     tracer.getParameters() ... KEY_x()          tracer.setParameter("MaxPathLength", 8);
//     tracer.DoPass(&image, newTrace);            tracer.DoPass(image, newTrace);    // image: PixelData[w*h] host buffer or nullptr
//     tracer.getRaysInLastPass() ...              same names
//
// Errors: every failing ABI call is re-raised as std::runtime_error carrying the reference's
// "In file ... at line ... : msg" text (ThrowCudaErrors convention, Defines.cpp:15-29).  No CPU fallback exists.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "ctl_b200.h"

namespace ctlb200 {

inline void check(int rc) { if (rc != 0) throw std::runtime_error(ctl_last_error()); }

// RAII owner of a host scene built by the library's own builder (synthetic scenes / one user mesh).
class Scene {
public:
    Scene(int kind, int width, int height, uint32_t seed = 1234, int n_hint = 0) : h_(ctl_scene_create(kind, width, height, seed, n_hint)) {
        if (!h_) throw std::runtime_error(ctl_last_error());
        check(ctl_scene_get_view(h_, &view_));
    }
    // flat import of compiled / source mesh files (.xmsh, .obj, .ply), one node per file: ctl_scene_create_from_files
    Scene(const std::vector<std::string>& paths, const float cam_pos[3], const float cam_target[3], const float cam_up[3], float fov_deg, int width, int height,
          const float* node_xforms = nullptr) {
        std::vector<const char*> p; for (auto& s : paths) p.push_back(s.c_str());
        h_ = ctl_scene_create_from_files(p.data(), (uint32_t)p.size(), node_xforms, cam_pos, cam_target, cam_up, fov_deg, width, height);
        if (!h_) throw std::runtime_error(ctl_last_error());
        check(ctl_scene_get_view(h_, &view_));
    }
    ~Scene() { ctl_scene_destroy(h_); }
    Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;
    const ctl_scene_view& view() const { return view_; }
private:
    ctl_scene* h_; ctl_scene_view view_;
};

// == CudaTracerLib::PathTracer (progressive tracer; one DoPass = one path per pixel)
class PathTracer {
public:
    explicit PathTracer(int device = 0) : device_(device) {}
    ~PathTracer() { if (ctx_) ctl_destroy(ctx_); }
    PathTracer(const PathTracer&) = delete; PathTracer& operator=(const PathTracer&) = delete;

    // TracerBase::Resize (Kernel/Tracer.h:104-116)
    void Resize(unsigned w, unsigned h) {
        if (!ctx_) { ctx_ = ctl_create(device_, (int)w, (int)h); if (!ctx_) throw std::runtime_error(ctl_last_error()); apply_params(); }
        else check(ctl_resize(ctx_, (int)w, (int)h));
        w_ = w; h_ = h; new_trace_ = true;
    }
    // TracerBase::InitializeScene (Kernel/Tracer.h:100-103) + the scene half of UpdateKernel (Kernel/TraceHelper.cu:182-217)
    void InitializeScene(const ctl_scene_view& view) { need_ctx(); check(ctl_upload_scene(ctx_, &view)); new_trace_ = true; }
    // m_sParameters << KEY_Direct() / KEY_MaxPathLength() / KEY_RRStartDepth() / KEY_Regularization() (Integrators/PathTracer.h:10-20)
    void setParameter(const std::string& key, int value) {
        for (auto& kv : params_) if (kv.first == key) { kv.second = value; if (ctx_) check(ctl_set_param_i(ctx_, key.c_str(), value)); return; }
        params_.emplace_back(key, value);
        if (ctx_) check(ctl_set_param_i(ctx_, key.c_str(), value));
    }
    int getParameter(const std::string& key) { need_ctx(); int v = 0; check(ctl_get_param_i(ctx_, key.c_str(), &v)); return v; }
    // Tracer<true>::DoPass (Kernel/Tracer.h:209-248).  `image` (optional) receives the PixelData accumulator after the pass,
    // as Image::getPixelData would expose it (Engine/Image.h:10-29, 78-81); pass nullptr to keep it on the device.
    void DoPass(ctl_pixel_data* image, bool a_NewTrace) {
        need_ctx();
        check(ctl_render_pass(ctx_, (a_NewTrace || new_trace_) ? 1 : 0, 0, 0, (int)w_, (int)h_));
        new_trace_ = false;
        if (image) check(ctl_read_accum(ctx_, image)); else check(ctl_synchronize(ctx_));
    }
    bool isMultiPass() const { return true; }
    unsigned getNumPassesDone() { uint32_t p = 0; stats(nullptr, nullptr, nullptr, &p); return p; }
    unsigned long long getRaysInLastPass() { uint64_t r = 0; stats(&r, nullptr, nullptr, nullptr); return r; }
    float getLastTimeSpentRenderingSec() { float s = 0; stats(nullptr, &s, nullptr, nullptr); return s; }
    unsigned long long getAccRays() { uint64_t r = 0; stats(nullptr, nullptr, &r, nullptr); return r; }
    float getSplatScale() const { return 0.0f; } // the path tracer never splats (rgbSplat stays 0)
    // __internal__IntersectBuffers (Kernel/TraceHelper.cu:736-746): device pointers, asynchronous on `stream`
    void IntersectBuffers(int n, const ctl_traversal_ray* d_rays, ctl_traversal_result* d_results, bool any_hit, void* stream = nullptr) {
        need_ctx(); check(ctl_intersect(ctx_, n, d_rays, d_results, any_hit ? 1 : 0, stream));
    }
    // TracerBase::TraceSingleRay (Kernel/Tracer.cu:74-78), batched; host buffers
    std::vector<ctl_trace_result> TraceRays(const std::vector<ctl_traversal_ray>& rays) {
        need_ctx(); std::vector<ctl_trace_result> out(rays.size());
        check(ctl_trace_rays_host(ctx_, (int)rays.size(), rays.data(), out.data(), nullptr)); return out;
    }
    // n_passes DoPass calls fused into one wavefront on the tiles of `part` of `n_parts` (ctl_render_passes_tiled); asynchronous
    void DoPassesTiled(int n_passes, bool a_NewTrace, int part = 0, int n_parts = 1, int tile = 64) {
        need_ctx(); check(ctl_render_passes_tiled(ctx_, (a_NewTrace || new_trace_) ? 1 : 0, n_passes, tile, tile, part, n_parts)); new_trace_ = false;
    }
    // one progressive frame (a new trace of spp passes, `batch` fused per wavefront, the wavefronts overlapped on two streams: ctl_render_frame_tiled); asynchronous
    void DoFrameTiled(int spp, int batch = 8, int part = 0, int n_parts = 1, int tile = 64) {
        need_ctx(); check(ctl_render_frame_tiled(ctx_, spp, batch, tile, tile, part, n_parts)); new_trace_ = false;
    }
    // a sequence of frames as a pipeline ("FramesInFlight"): SubmitFrame enqueues one more frame on a lane of its own, AcquireFrame hands back the oldest one
    // (ctl_submit_frame_tiled / ctl_acquire_frame); asynchronous, `image` (optional) reads the acquired frame's PixelData
    void SubmitFrame(int spp, int batch = 8, int part = 0, int n_parts = 1, int tile = 64) {
        need_ctx(); check(ctl_submit_frame_tiled(ctx_, spp, batch, tile, tile, part, n_parts)); new_trace_ = true;
    }
    void AcquireFrame(ctl_pixel_data* image = nullptr) { need_ctx(); check(ctl_acquire_frame(ctx_)); if (image) check(ctl_read_accum(ctx_, image)); }
    int FramesInFlight() const { return ctx_ ? ctl_frames_in_flight(ctx_) : 0; }
    void Synchronize() { need_ctx(); check(ctl_synchronize(ctx_)); }
    ctl_ctx* handle() { return ctx_; }
protected:
    bool take_new_trace(bool a_NewTrace) { const bool nt = a_NewTrace || new_trace_; new_trace_ = false; return nt; }
    void finish_pass(ctl_pixel_data* image) { if (image) check(ctl_read_accum(ctx_, image)); else check(ctl_synchronize(ctx_)); }
    void require_ctx() const { need_ctx(); }
private:
    void need_ctx() const { if (!ctx_) throw std::runtime_error("PathTracer: call Resize(w, h) first"); }
    void apply_params() { for (auto& kv : params_) check(ctl_set_param_i(ctx_, kv.first.c_str(), kv.second)); }
    void stats(uint64_t* r, float* s, uint64_t* t, uint32_t* p) { need_ctx(); check(ctl_stats(ctx_, r, s, t, p)); }
    int device_; ctl_ctx* ctx_ = nullptr; unsigned w_ = 0, h_ = 0; bool new_trace_ = true;
    std::vector<std::pair<std::string, int>> params_;
};

// == CudaTracerLib::WavefrontPathTracer (Integrators/PseudoRealtime/WavefrontPathTracer.h:28-66): the reference's own wavefront
// integrator over DoubleRayBuffer -- same parameter keys, same DoPass contract; one pass = one path per pixel.
class WavefrontPathTracer : public PathTracer {
public:
    explicit WavefrontPathTracer(int device = 0) : PathTracer(device) {
        setParameter("Direct", 1); setParameter("MaxPathLength", 50); setParameter("RRStartDepth", 5); // WavefrontPathTracer.h:38-42
    }
    // Tracer<true>::DoPass -> WavefrontPathTracer::DoRender (WavefrontPathTracer.cu:166-191)
    void DoPass(ctl_pixel_data* image, bool a_NewTrace) {
        require_ctx();
        check(ctl_wavefront_pass(handle(), take_new_trace(a_NewTrace) ? 1 : 0));
        finish_pass(image);
    }
    // a whole progressive frame: StartNewTrace + spp DoPass calls, the passes overlapped on several streams (ctl_wavefront_frame)
    void DoFrame(ctl_pixel_data* image, int spp) {
        require_ctx();
        check(ctl_wavefront_frame(handle(), spp)); take_new_trace(false);
        finish_pass(image);
    }
};

// Several GPUs of one node driven by ONE host process (no reference counterpart: the reference is single-GPU).  One PathTracer per device, the same
// scene uploaded to each, the image split in interleaved 64x64 tiles (tile index % devices == device), one NCCL reduce of the PixelData accumulators
// to device 0 per frame (csrc/ctl_comm.cu).  Every path is the one a single device would trace, so the reduced image equals the single-GPU image up
// to the order of float additions into a pixel.
class MultiGpuPathTracer {
public:
    explicit MultiGpuPathTracer(int n_devices) { for (int d = 0; d < n_devices; d++) t_.emplace_back(new PathTracer(d)); }
    int devices() const { return (int)t_.size(); }
    PathTracer& device(int d) { return *t_[d]; }
    void Resize(unsigned w, unsigned h) {
        for (auto& t : t_) t->Resize(w, h);
        std::vector<ctl_ctx*> c; for (auto& t : t_) c.push_back(t->handle());
        if (c.size() > 1) check(ctl_comm_init_all(c.data(), (int)c.size()));
        w_ = w; h_ = h;
    }
    void InitializeScene(const ctl_scene_view& view) { for (auto& t : t_) t->InitializeScene(view); }
    void setParameter(const std::string& key, int value) { for (auto& t : t_) t->setParameter(key, value); }
    // One progressive frame of `spp` passes (`batch` fused per wavefront): every device renders its tiles, then the one reduce to device 0.
    // image (optional): the reduced PixelData accumulator.  Returns the rays traced by all devices.
    unsigned long long RenderFrame(int spp, int batch, ctl_pixel_data* image = nullptr) {
        if (spp < 1 || batch < 1 || spp % batch) throw std::runtime_error("spp must be a positive multiple of batch");
        const int n = devices();
        for (int p = 0; p < spp; p += batch)
            for (int d = 0; d < n; d++) t_[d]->DoPassesTiled(batch, p == 0, d, n);     // asynchronous: the devices run concurrently
        std::vector<ctl_ctx*> c; for (auto& t : t_) c.push_back(t->handle());
        check(ctl_comm_reduce_accum_all(c.data(), n, 0));
        unsigned long long rays = 0;
        for (int d = 0; d < n; d++) { t_[d]->Synchronize(); rays += t_[d]->getAccRays(); }
        const unsigned long long frame_rays = rays - rays_before_; rays_before_ = rays;
        if (image) check(ctl_read_accum(t_[0]->handle(), image));
        return frame_rays;
    }
    // A sequence of frames as a pipeline (ctl_comm_submit_frame_all / ctl_acquire_frame, "FramesInFlight"): SubmitFrame enqueues one more frame on every device
    // (its own lane and accumulator; the reduce on the communication streams); AcquireFrame waits for the oldest one and, optionally, reads its image.
    void SubmitFrame(int spp, int batch, int tile = 64) {
        std::vector<ctl_ctx*> c; for (auto& t : t_) c.push_back(t->handle());
        check(ctl_comm_submit_frame_all(c.data(), devices(), spp, batch, tile, 0));
    }
    void AcquireFrame(ctl_pixel_data* image = nullptr) {
        for (auto& t : t_) check(ctl_acquire_frame(t->handle()));
        if (image) check(ctl_read_accum(t_[0]->handle(), image));
    }
    int FramesInFlight() const { return ctl_frames_in_flight(t_[0]->handle()); }
private:
    std::vector<std::unique_ptr<PathTracer>> t_; unsigned w_ = 0, h_ = 0; unsigned long long rays_before_ = 0;
};

} // namespace ctlb200
