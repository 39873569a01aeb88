// ctl_api.cu -- C ABI (include/ctl_b200.h) over the sm_100a kernels: context, scene upload, traversal queries, the two integrators' host loops.
//
// Host orchestration replaces Tracer<true>::DoPass / UpdateKernel / __internal__IntersectBuffers
// (Kernel/Tracer.h:209-289, Kernel/TraceHelper.cu:182-217, 736-746): one context per device, one
// stream, no globals, no per-launch cudaDeviceSynchronize, queue sizes stay on the device.
// (Scenes: ctl_scene_api.cpp; image pipeline: ctl_pipeline.cu; GPU BVH build: ctl_bvh_gpu.cu; NCCL: ctl_comm.cu.)
#include "ctl_internal.h"
#include "wavefront.cuh"
#include "wavefront_pt.cuh"


template <bool REGU>
static void launch_shade_t(int cls, int grid, cudaStream_t st, const DScene& S, const ShadeParams& P, const PathState& ps, const Queues& Q, const unsigned* n_in, unsigned* n_out, unsigned* n_shadow, const unsigned* seg_hist) {
    switch (cls) {
    case 0: k_shade<0, REGU><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    case 1: k_shade<1, REGU><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    case 2: k_shade<2, REGU><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    case 3: k_shade<3, REGU><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    default: k_shade<-1, REGU><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, nullptr); break;
    }
}
static void launch_shade(bool regu, int cls, int grid, cudaStream_t st, const DScene& S, const ShadeParams& P, const PathState& ps, const Queues& Q, const unsigned* n_in, unsigned* n_out, unsigned* n_shadow, const unsigned* seg_hist) {
    if (regu) launch_shade_t<true>(cls, grid, st, S, P, ps, Q, n_in, n_out, n_shadow, seg_hist);   // KEY_Regularization: PathTraceRegularization's vertex (Integrators/PathTracer.cu:115-170)
    else launch_shade_t<false>(cls, grid, st, S, P, ps, Q, n_in, n_out, n_shadow, seg_hist);
}

// Launch shape of the staged kernel: blocks of `staged_threads`, as many per SM as 1024 resident threads and the shared memory allow
static size_t staged_smem_bytes(const ctl_ctx* c) {
    return (size_t)c->staged.tl_nodes * 64 + 16 + (size_t)(c->staged.stack_rows + 1) * c->staged_threads * 4 + (c->staged.ray_tma ? (size_t)(c->staged_threads / 32) * (1024 + 8) : 0);
}
static int staged_grid(const ctl_ctx* c) {
    const size_t smem = staged_smem_bytes(c);
    int per_sm = c->staged_resident / c->staged_threads;
    const int fit = (int)((227u * 1024u) / (smem + 1024));
    if (per_sm > fit) per_sm = fit;
    if (per_sm < 1) per_sm = 1;
    return c->n_sm * per_sm;
}
template <int MODE, bool ANY_HIT, bool COUNT>
static void launch_staged(const ctl_ctx* c, cudaStream_t st, const float4* rays, const unsigned* n_ptr, const unsigned* n2_ptr, int n_fixed, unsigned* work, const TravOut& out, unsigned long long* visit) {
    static size_t attr_set_dev[64][2] = {}; // per instantiation and device (the attribute is a per-device property of the function)
    size_t* attr_set = attr_set_dev[c->device & 63];
    const size_t smem = staged_smem_bytes(c);
    if (c->staged.ray_tma) {
        if (smem > attr_set[1]) { cudaFuncSetAttribute(k_intersect_staged<MODE, ANY_HIT, COUNT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set[1] = smem; }
        k_intersect_staged<MODE, ANY_HIT, COUNT, true><<<staged_grid(c), c->staged_threads, smem, st>>>(c->scene, c->staged, c->tune, rays, n_ptr, n2_ptr, n_fixed, work, out, visit);
    } else {
        if (smem > attr_set[0]) { cudaFuncSetAttribute(k_intersect_staged<MODE, ANY_HIT, COUNT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set[0] = smem; }
        k_intersect_staged<MODE, ANY_HIT, COUNT, false><<<staged_grid(c), c->staged_threads, smem, st>>>(c->scene, c->staged, c->tune, rays, n_ptr, n2_ptr, n_fixed, work, out, visit);
    }
}

// "TraversalKernel": 0 = persistent phase-scheduled kernel, 1 = simple ray-batch kernel (A/B baseline), 2 = persistent kernel with shared-memory staging
template <int MODE, bool ANY_HIT, bool COUNT>
static void launch_intersect(const ctl_ctx* c, int grid, cudaStream_t st, const DScene& S, const float4* rays, const unsigned* n_ptr, int n_fixed, unsigned* work_ctr,
                             float4* hit_a, uint32_t* hit_node, const float4* sh_payload, float4* cl, void* api_out, unsigned long long* visit_out, unsigned* cls_hist = nullptr) {
    if (c->trav_kernel == 2 && c->staged_ok) {
        const TravOut out = {hit_a, hit_node, sh_payload, cl, api_out, nullptr, 0, nullptr, cls_hist};
        launch_staged<MODE, ANY_HIT, COUNT>(c, st, rays, n_ptr, nullptr, n_fixed, work_ctr, out, visit_out);
    }
    else if (c->trav_kernel == 1) k_intersect_simple<MODE, ANY_HIT, COUNT><<<grid, 128, 0, st>>>(S, rays, n_ptr, n_fixed, work_ctr, hit_a, hit_node, sh_payload, cl, api_out, visit_out);
    else k_intersect<MODE, ANY_HIT, COUNT><<<grid, 128, 0, st>>>(S, c->tune_p, rays, n_ptr, n_fixed, work_ctr, hit_a, hit_node, sh_payload, cl, api_out, visit_out);
}

extern "C" {

// ------------------------------------------------------------------ context
static void release_frame_slots(ctl_ctx* c) {   // (every stream idle) the accumulators / table state of the frames-in-flight slots
    for (FrameSlot& F : c->fslot) {
        F.accum.release(); F.states.release();
        if (F.h1) cudaFreeHost(F.h1); if (F.h2) cudaFreeHost(F.h2); F.h1 = F.h2 = nullptr; F.h_cap = 0;
        for (cudaEvent_t* e : {&F.begin, &F.done, &F.ready, &F.h_free}) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
    }
    for (int k = 0; k < MAX_LANES; k++) c->lane_accum[k] = nullptr;
    c->f_submitted = c->f_acquired = 0;
}

static int alloc_image(ctl_ctx* c) {
    CK(c->own_accum.ensure((size_t)c->w * c->h * 7));
    c->accum = c->own_accum.p;
    CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream));
    return 0;
}

static int init_ctx(ctl_ctx* c) { // everything of ctl_create that can fail after the context object exists
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    c->n_sm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaEventCreate(&c->ev_start)); CK(cudaEventCreate(&c->ev_stop));
    CK(cudaEventCreateWithFlags(&c->h_tab_free, cudaEventDisableTiming));
    {   // device-side table generator: per-sequence start states of pass 0 + the jump matrix to the next pass
        std::vector<uint32_t> st0((size_t)ctlb::kNumSeq * 6); ctlb::XorwowJump J;
        ctlb::device_generator_data(st0.data(), &J);
        CK(c->d_states0.upload(st0.data(), st0.size())); CK(c->d_states.upload(st0.data(), st0.size()));
        CK(c->d_jump.upload(&J.row[0][0], 160 * 5));
    }
    CK(c->lanes[0].counters.ensure(CTR_TOTAL)); CK(c->api_work.ensure(API_WORK_RING)); CK(c->stats.ensure(16)); CK(c->d_captured_n.ensure(1));
    CK(cudaMemset(c->stats.p, 0, 16 * sizeof(unsigned long long)));
    return alloc_image(c);
}

ctl_ctx* ctl_create(int device, int width, int height) {
    if (width <= 0 || height <= 0) { set_err("invalid resolution"); return nullptr; }
    int n_dev = 0;
    CKP(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) { set_err("no such CUDA device"); return nullptr; }
    CKP(cudaSetDevice(device));
    ctl_ctx* c = new ctl_ctx();
    c->device = device; c->w = width; c->h = height;
    if (init_ctx(c)) { const std::string keep = ctl_last_error(); ctl_destroy(c); ctl_set_err(keep); return nullptr; }
    return c;
}

void ctl_destroy(ctl_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int k = 1; k < MAX_LANES; k++) if (c->lane_stream[k]) cudaStreamSynchronize(c->lane_stream[k]);   // frames still in flight: their kernels and reduces end first
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    ctl_comm_release(c);
    c->d_scene_nodes.release(); c->d_bvh_nodes.release(); c->d_woop.release(); c->d_tri_index.release(); c->d_tri_data.release(); c->d_meshes.release();
    c->d_nodes.release(); c->d_xf.release(); c->d_inv_xf.release(); c->d_materials.release(); c->d_lights.release(); c->d_light_tris.release();
    c->d_light_cdf.release(); c->d_normal_lut.release(); c->d_tri64.release(); c->d_inst.release(); c->d_treelet.release();
    c->d_tab1.release(); c->d_tab2.release(); c->d_states.release(); c->d_states0.release(); c->d_jump.release();
    if (c->h_tab1) cudaFreeHost(c->h_tab1); if (c->h_tab2) cudaFreeHost(c->h_tab2); if (c->h_tab_free) cudaEventDestroy(c->h_tab_free);
    for (int k = 0; k < MAX_LANES; k++) { if (c->ev_cls_fork[k]) cudaEventDestroy(c->ev_cls_fork[k]); for (int j = 0; j < 3; j++) { if (c->cls_stream[k][j]) { cudaStreamSynchronize(c->cls_stream[k][j]); cudaStreamDestroy(c->cls_stream[k][j]); } if (c->ev_cls_done[k][j]) cudaEventDestroy(c->ev_cls_done[k][j]); } }
    if (c->tab_stream) { cudaStreamSynchronize(c->tab_stream); cudaStreamDestroy(c->tab_stream); } if (c->ev_tab) cudaEventDestroy(c->ev_tab);
    for (int k = 1; k < MAX_LANES; k++) { if (c->lane_stream[k]) { cudaStreamSynchronize(c->lane_stream[k]); cudaStreamDestroy(c->lane_stream[k]); } if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->comm_stream) { cudaStreamSynchronize(c->comm_stream); cudaStreamDestroy(c->comm_stream); }
    release_frame_slots(c);
    for (int k = 0; k < MAX_LANES; k++) c->lanes[k].release();
    c->ho_buf[0].release(); c->ho_buf[1].release(); c->ho_cnt.release(); c->df_cnt.release();
    for (int k = 0; k < MAX_LANES; k++) { c->df_sh_rays[k].release(); c->df_sh_payload[k].release(); }
    c->capture.release(); c->api_work.release(); c->stats.release();
    c->own_accum.release(); c->d_captured_n.release(); c->resolve_tmp.release(); c->pipe_rgbe.release(); c->pipe_partial.release(); c->pipe_lum.release(); c->d_var.release(); c->nlm_cached.release(); c->nlm_varh.release(); c->nlm_weights.release(); c->nlm_last_update = -1; c->nlm_pixels = 0; c->d_node_alias.release();
    for (int k = 0; k < MAX_LANES; k++) c->wl[k].release();
    for (auto e : c->stage_ev) cudaEventDestroy(e);
    if (c->ev_start) cudaEventDestroy(c->ev_start); if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int ctl_resize(ctl_ctx* c, int width, int height) {
    if (!c || width <= 0 || height <= 0) return set_err("invalid argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->f_submitted != c->f_acquired) return set_err("ctl_resize: frames are in flight (ctl_acquire_frame them first)");
    for (int k = 1; k < MAX_LANES; k++) if (c->lane_stream[k]) CK(cudaStreamSynchronize(c->lane_stream[k]));
    release_frame_slots(c);
    c->w = width; c->h = height; c->scene.img_w = width; c->scene.img_h = height;
    c->passes_done = 0;
    c->nlm_pixels = 0; c->nlm_last_update = -1;   // NonLocalMeansFilter::Resize (NonLocalMeansFilter.h:131-137)
    return alloc_image(c);
}

int ctl_set_param_i(ctl_ctx* c, const char* key, int v) {
    if (!c || !key) return set_err("null argument");
    std::string k(key);
    if (k == "MaxPathLength") { if (v < 1 || v > MAX_BOUNCES) return set_err("MaxPathLength out of range [1,256]"); c->max_path_length = v; }
    else if (k == "RRStartDepth") { if (v < 0) return set_err("RRStartDepth must be >= 0"); c->rr_start = v; }
    else if (k == "Direct") c->direct = v != 0;
    else if (k == "StopZeroThroughput") c->stop_zero = v != 0;   // 1 (default): a path whose throughput is exactly zero ends; 0: it is traced until Russian roulette ends it, the reference's ray count (Kernel/TraceHelper.cu:176)
    else if (k == "Regularization") c->regularization = v != 0;   // KEY_Regularization (Integrators/PathTracer.h:10-20): PathTraceRegularization instead of PathTrace; needs MaxPathLength <= 255
    else if (k == "SortMode") c->sort_mode = v;
    else if (k == "ShadeMode") { if (v < 0 || v > 1) return set_err("ShadeMode must be 0 (run-time BSDF dispatch) or 1 (one launch per material class)"); c->shade_mode = v; }
    else if (k == "StageTimers") c->stage_timers = v != 0;
    else if (k == "CaptureBounce") c->capture_bounce = v;
    else if (k == "DeviceSampleTables") c->device_tables = v != 0;
    else if (k == "FuseTraversal") c->fuse_traversal = v != 0;
    else if (k == "OverlapWavefronts") { if (v < 0 || v > 2) return set_err("OverlapWavefronts must be 0, 1 or 2"); c->overlap = v; }
    else if (k == "ShadeConcurrent") c->shade_concurrent = v != 0;   // the per-class shade launches of a bounce on their own streams
    else if (k == "DeferStragglers") c->defer = v != 0;   // frames: traversal launches move their unfinished rays into the wavefront's next launch (paths may lag DeferMaxLag bounces)
    else if (k == "DeferMaxLag") { if (v < 1 || v > 3) return set_err("DeferMaxLag out of range [1,3]"); c->defer_max_lag = v; }
    else if (k == "HandOver") c->handover = v != 0;   // ctl_render_frame_tiled: one-wavefront frames as two half-wavefronts with ray hand-over between their launches
    else if (k == "HandOverDrain") { if (v < 1 || v > 4096) return set_err("HandOverDrain out of range [1,4096]"); c->handover_drain = v; }
    else if (k == "OverlapLanes") { if (v < 1 || v > MAX_LANES) return set_err("OverlapLanes out of range [1,8]"); c->n_lanes = v; }   // wavefronts of a frame in flight at once   // ctl_render_frame_tiled / ctl_comm_render_frame: the frame's wavefronts alternate between two streams
    else if (k == "FramesInFlight") {   // ctl_submit_frame_tiled / ctl_acquire_frame: how many frames may be outstanding
        if (v < 1 || v > MAX_LANES - 1) return set_err("FramesInFlight out of range [1,7]");
        if (c->f_submitted != c->f_acquired) return set_err("FramesInFlight cannot change while frames are in flight");
        c->fif = v; c->f_submitted = c->f_acquired = 0;
    }
    else if (k == "PixelVarianceBuffer") c->variance_buffer = v != 0;
    else if (k == "WarpPixelBlocks") c->warp_blocks = v != 0;
    else if (k == "PassStride") { if (v < 1) return set_err("PassStride must be >= 1"); c->pass_stride = v; }   // multi-GPU by pass: this context renders passes PassPhase + k * PassStride
    else if (k == "PassPhase") { if (v < 0) return set_err("PassPhase must be >= 0"); c->pass_phase = v; }
    else if (k == "TraversalKernel") { if (v < 0 || v > 2) return set_err("TraversalKernel must be 0 (persistent), 1 (ray batch) or 2 (staged)"); c->trav_kernel = v; }
    else if (k == "StagedThreads") { if (v < 32 || v > 1024 || (v & 31)) return set_err("StagedThreads must be a multiple of 32 in [32,1024]"); c->staged_threads = v; }
    else if (k == "StagedResidentThreads") { if (v < 32 || v > 2048) return set_err("StagedResidentThreads out of range [32,2048]"); c->staged_resident = v; }
    else if (k == "StagedRayTMA") c->staged.ray_tma = v != 0;   // ray-queue chunks through TMA bulk copies into per-warp shared-memory buffers
    else if (k == "StagedStackRows") { if (v < 0 || v > TP_STACK) return set_err("StagedStackRows out of range [0,64]"); c->staged_rows = v; c->staged.stack_rows = v; }
    else if (k == "StagedTreeletNodes") { if (v < 0 || v > 2048) return set_err("StagedTreeletNodes out of range [0,2048]"); c->staged_treelet = v; }   // takes effect at the next ctl_upload_scene / ctl_update_scene_nodes
    else if (k == "TravThT") c->tune.th_t = c->tune_p.th_t = v; else if (k == "TravThL") c->tune.th_l = c->tune_p.th_l = v; else if (k == "TravThF") c->tune.th_f = c->tune_p.th_f = v;
    else if (k == "TravThNExit") c->tune.th_n_exit = c->tune_p.th_n_exit = v;
    else if (k == "TravChunk") { if (v < 0 || v > 4096 || (v & 31)) return set_err("TravChunk must be 0 (by queue size) or a multiple of 32 up to 4096"); c->tune.chunk = v; }   // rays a warp claims per atomic (staged kernel)
    else if (k == "TravDrainPrefetch") c->tune.drain_prefetch = v != 0;   // staged kernel: prefetch both children per node step once the queue is exhausted
    else if (k == "TravTSteps") { if (v < 1 || v > 8) return set_err("TravTSteps out of range [1,8]"); c->tune.t_steps = v; }
    else if (k == "ShadeBlocksPerSM") { if (v < 1 || v > 16) return set_err("ShadeBlocksPerSM out of range [1,16]"); c->shade_blocks_per_sm = v; }
    else if (k == "TravSmemCarveout") { // experiment: shared-memory carve-out (percent) of the traversal kernels = how much L1 they lose
        c->smem_carveout = v;
        CK(cudaFuncSetAttribute(k_intersect<0, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v));
        CK(cudaFuncSetAttribute(k_intersect<1, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v));
    }
    else if (k == "TraversalBlocksPerSM") { if (v < 1 || v > 16) return set_err("TraversalBlocksPerSM out of range [1,16]"); c->trav_blocks_per_sm = v; }
    else return set_err("unknown parameter key: " + k);
    return 0;
}
int ctl_get_param_i(ctl_ctx* c, const char* key, int* v) {
    if (!c || !key || !v) return set_err("null argument");
    std::string k(key);
    if (k == "MaxPathLength") *v = c->max_path_length; else if (k == "RRStartDepth") *v = c->rr_start; else if (k == "Direct") *v = c->direct;
    else if (k == "StopZeroThroughput") *v = c->stop_zero; else if (k == "Regularization") *v = c->regularization; else if (k == "SortMode") *v = c->sort_mode; else if (k == "StageTimers") *v = c->stage_timers;
    else if (k == "CaptureBounce") *v = c->capture_bounce; else if (k == "TraversalKernel") *v = c->trav_kernel; else if (k == "DeviceSampleTables") *v = c->device_tables; else if (k == "FuseTraversal") *v = c->fuse_traversal; else if (k == "OverlapWavefronts") *v = c->overlap; else if (k == "OverlapLanes") *v = c->n_lanes; else if (k == "HandOver") *v = c->handover; else if (k == "DeferStragglers") *v = c->defer; else if (k == "DeferMaxLag") *v = c->defer_max_lag; else if (k == "ShadeConcurrent") *v = c->shade_concurrent; else if (k == "HandOverDrain") *v = c->handover_drain; else if (k == "PixelVarianceBuffer") *v = c->variance_buffer; else if (k == "PassStride") *v = c->pass_stride; else if (k == "PassPhase") *v = c->pass_phase; else if (k == "FramesInFlight") *v = c->fif;
    else if (k == "TraversalBlocksPerSM") *v = c->trav_blocks_per_sm; else if (k == "StagedThreads") *v = c->staged_threads; else if (k == "StagedStackRows") *v = c->staged_rows;
    else if (k == "ShadeMode") *v = c->shade_mode; else if (k == "MaterialClassMask") *v = (int)c->class_mask;
    else if (k == "StagedRayTMA") *v = c->staged.ray_tma;
    else if (k == "StagedTreeletNodes") *v = c->staged.tl_nodes; else if (k == "StagedUsable") *v = c->staged_ok ? 1 : 0; else return set_err("unknown parameter key: " + k);
    return 0;
}

// Derived records of the staged traversal kernel (csrc/staging.cpp): leaf triangles (with_tris) and the node-level half (instance records, treelet).
static int upload_staging(ctl_ctx* c, const ctl_scene_view* v, bool with_tris) {
    ctlb::StagedHost H;
    if (with_tris) {
        ctlb::build_staging_tris(*v, H);
        c->staged_ok = H.usable; c->staged_why = H.why; c->class_mask = H.class_mask;
        if (H.usable) CK(c->d_tri64.upload((const float4*)H.tri64.data(), H.tri64.size() / 4));
    }
    if (!c->staged_ok) return 0;
    ctlb::build_staging_nodes(*v, c->staged_treelet, H);
    CK(c->d_inst.upload((const float4*)H.inst.data(), H.inst.size() / 4));
    CK(c->d_treelet.upload((const float4*)H.treelet.data(), H.treelet.size() / 4));
    c->staged.tri64 = c->d_tri64.p; c->staged.inst = c->d_inst.p; c->staged.treelet = c->d_treelet.p;
    c->staged.tl_nodes = H.tl_nodes; c->staged.scene_root = H.scene_root; c->staged.stack_rows = c->staged_rows;
    c->class_ok = H.class_ok;
    return 0;
}

int ctl_upload_scene(ctl_ctx* c, const ctl_scene_view* v) {
    if (!c || !v) return set_err("null argument");
    if (c->f_submitted != c->f_acquired) return set_err("ctl_upload_scene: frames are in flight (ctl_acquire_frame them first)");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(c->d_scene_nodes.upload(v->scene_bvh_nodes, v->n_scene_bvh_nodes)); CK(c->d_bvh_nodes.upload(v->bvh_nodes, v->n_bvh_nodes));
    CK(c->d_woop.upload(v->woop, v->n_woop)); CK(c->d_tri_index.upload(v->tri_index, v->n_tri_index)); CK(c->d_tri_data.upload(v->tri_data, v->n_tri_data));
    CK(c->d_meshes.upload(v->meshes, v->n_meshes)); CK(c->d_nodes.upload(v->nodes, v->n_nodes));
    c->n_alias = 0; if (v->node_alias) { CK(c->d_node_alias.upload(v->node_alias, v->n_nodes)); c->n_alias = v->n_nodes; }
    CK(c->d_xf.upload(v->node_xf, (size_t)v->n_nodes * 16)); CK(c->d_inv_xf.upload(v->node_inv_xf, (size_t)v->n_nodes * 16));
    CK(c->d_materials.upload(v->materials, v->n_materials)); CK(c->d_lights.upload(v->lights, v->n_lights_buf));
    CK(c->d_light_tris.upload(v->light_tris, v->n_light_tris)); CK(c->d_light_cdf.upload(v->light_cdf_data, v->n_light_cdf_data));
    // sin/cos tables of the 16-bit spherical normal code (Math/Compression.h:20-31), computed with the host libm
    std::vector<float> lut(1024);
    for (int i = 0; i < 256; i++) {
        float th, ph, th2, ph2;
        ctlb::normal_code_angles((uint16_t)(i << 8), th, ph); ctlb::normal_code_angles((uint16_t)i, th2, ph2);
        lut[i] = sinf(th); lut[256 + i] = cosf(th); lut[512 + i] = sinf(ph2); lut[768 + i] = cosf(ph2);
        (void)ph; (void)th2;
    }
    CK(c->d_normal_lut.upload(lut.data(), 1024));
    DScene& S = c->scene;
    S.scene_nodes = (const float4*)c->d_scene_nodes.p; S.bvh_nodes = (const float4*)c->d_bvh_nodes.p; S.woop = (const float4*)c->d_woop.p;
    S.tri_index = c->d_tri_index.p; S.tri_data = (const uint4*)c->d_tri_data.p; S.meshes = c->d_meshes.p; S.nodes = c->d_nodes.p;
    S.node_xf = (const float4*)c->d_xf.p; S.node_inv_xf = (const float4*)c->d_inv_xf.p; S.materials = c->d_materials.p; S.lights = c->d_lights.p;
    S.light_tris = c->d_light_tris.p; S.light_cdf_data = c->d_light_cdf.p; S.normal_lut = c->d_normal_lut.p;
    S.d1 = c->d_tab1.p; S.d2 = (const float2*)c->d_tab2.p;
    S.num_lights = v->num_lights;
    memcpy(S.light_indices, v->light_indices, sizeof(S.light_indices)); memcpy(S.light_cdf, v->light_cdf, sizeof(S.light_cdf));
    for (int k = 0; k < 3; k++) { S.box_min[k] = v->box_min[k]; const float e = v->box_max[k] - v->box_min[k]; S.box_inv_extent[k] = e > 0 ? 1.0f / e : 0.0f; }
    S.camera = v->camera; S.ray_eps = v->ray_eps; S.scene_start = v->scene_start_node; S.n_nodes = v->n_nodes;
    S.img_w = c->w; S.img_h = c->h;
    c->has_scene = true;
    return upload_staging(c, v, true);
}

// Node-level half of ctl_upload_scene for a view whose meshes are the ones already uploaded: nodes, transforms, scene-level BVH, lights, box, epsilon,
// camera -- what changes when instances move (a few KB instead of the whole scene; the reference re-uploads through Stream<T>::UpdateInvalidated).
int ctl_update_scene_nodes(ctl_ctx* c, const ctl_scene_view* v) {
    if (!c || !v) return set_err("null argument");
    if (!c->has_scene) return set_err("no scene uploaded");
    if (c->f_submitted != c->f_acquired) return set_err("ctl_update_scene_nodes: frames are in flight (ctl_acquire_frame them first)");
    if (v->node_alias || c->n_alias) return ctl_upload_scene(c, v);   // re-braided view (now or before): its mesh-level records follow the node level -> everything is uploaded
    if (v->n_bvh_nodes != c->d_bvh_nodes.n && v->n_bvh_nodes > c->d_bvh_nodes.n) return set_err("the view has other meshes than the uploaded scene: use ctl_upload_scene");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(c->d_scene_nodes.upload(v->scene_bvh_nodes, v->n_scene_bvh_nodes)); CK(c->d_nodes.upload(v->nodes, v->n_nodes));
    CK(c->d_xf.upload(v->node_xf, (size_t)v->n_nodes * 16)); CK(c->d_inv_xf.upload(v->node_inv_xf, (size_t)v->n_nodes * 16));
    CK(c->d_materials.upload(v->materials, v->n_materials)); CK(c->d_lights.upload(v->lights, v->n_lights_buf));
    CK(c->d_light_tris.upload(v->light_tris, v->n_light_tris)); CK(c->d_light_cdf.upload(v->light_cdf_data, v->n_light_cdf_data));
    DScene& S = c->scene;
    S.scene_nodes = (const float4*)c->d_scene_nodes.p; S.nodes = c->d_nodes.p; S.node_xf = (const float4*)c->d_xf.p; S.node_inv_xf = (const float4*)c->d_inv_xf.p;
    S.materials = c->d_materials.p; S.lights = c->d_lights.p; S.light_tris = c->d_light_tris.p; S.light_cdf_data = c->d_light_cdf.p;
    S.num_lights = v->num_lights;
    memcpy(S.light_indices, v->light_indices, sizeof(S.light_indices)); memcpy(S.light_cdf, v->light_cdf, sizeof(S.light_cdf));
    for (int k = 0; k < 3; k++) { S.box_min[k] = v->box_min[k]; const float e = v->box_max[k] - v->box_min[k]; S.box_inv_extent[k] = e > 0 ? 1.0f / e : 0.0f; }
    S.camera = v->camera; S.ray_eps = v->ray_eps; S.scene_start = v->scene_start_node; S.n_nodes = v->n_nodes;
    return upload_staging(c, v, false);
}

static const size_t TAB1 = (size_t)ctlb::kNumSeq * ctlb::kSeqLen, TAB2 = TAB1 * 2;

static int ensure_tables(ctl_ctx* c, int n_passes) {
    if (n_passes <= c->tab_cap) return 0;
    CK(cudaStreamSynchronize(c->stream));
    CK(c->d_tab1.ensure(TAB1 * n_passes)); CK(c->d_tab2.ensure(TAB2 * n_passes));
    c->tab_cap = n_passes;
    return 0;
}
static int ensure_host_tables(ctl_ctx* c, int n_passes) {
    if (n_passes <= c->h_tab_cap) return 0;
    CK(cudaEventSynchronize(c->h_tab_free));
    if (c->h_tab1) cudaFreeHost(c->h_tab1); if (c->h_tab2) cudaFreeHost(c->h_tab2);
    c->h_tab1 = c->h_tab2 = nullptr; c->h_tab_cap = 0;
    CK(cudaMallocHost((void**)&c->h_tab1, TAB1 * 4 * n_passes)); CK(cudaMallocHost((void**)&c->h_tab2, TAB2 * 4 * n_passes));
    c->h_tab_cap = n_passes;
    return 0;
}

int ctl_upload_samples(ctl_ctx* c, const float* d1, const float* d2) {
    if (!c || !d1 || !d2) return set_err("null argument");
    CK(cudaSetDevice(c->device));
    if (ensure_tables(c, 1) || ensure_host_tables(c, 1)) return 1;
    CK(cudaEventSynchronize(c->h_tab_free));
    memcpy(c->h_tab1, d1, TAB1 * 4); memcpy(c->h_tab2, d2, TAB2 * 4);
    CK(cudaMemcpyAsync(c->d_tab1.p, c->h_tab1, TAB1 * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_tab2.p, c->h_tab2, TAB2 * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->h_tab_free, c->stream));
    c->user_tables = true;
    return 0;
}

// Tables of passes [first, first + n) into the device table sets set0 .. set0+n-1 (GenerateNewRandomSequences, Kernel/Sampler.h:36-55), on stream `st`.
// Host mode: the caller has made sure the pinned sets it overwrites are free (h_tab_free).
static int generate_tables(ctl_ctx* c, uint32_t first, int n, int set0 = 0, cudaStream_t st = nullptr, bool wait_free = true) {
    if (!st) st = c->stream;
    if (ensure_tables(c, set0 + n)) return 1;
    float* d1 = c->d_tab1.p + TAB1 * set0; float* d2 = c->d_tab2.p + TAB2 * set0;
    if (c->device_tables) {
        if (c->gen_pos_dev != first) { // re-synchronise the device stream position (mode switch / user tables / strided passes): skip forward, or restart
            uint32_t from = c->gen_pos_dev;
            if (from > first) { CK(cudaMemcpyAsync(c->d_states.p, c->d_states0.p, (size_t)ctlb::kNumSeq * 6 * 4, cudaMemcpyDeviceToDevice, st)); from = 0; }
            for (uint32_t p = from; p < first; p++) k_gen_tables<<<ctlb::kNumSeq / 128, 128, 0, st>>>(c->d_states.p, c->d_jump.p, 1, d1, (float2*)d2);
        }
        k_gen_tables<<<ctlb::kNumSeq / 128, 128, 0, st>>>(c->d_states.p, c->d_jump.p, n, d1, (float2*)d2);
        CK(cudaGetLastError());
        c->gen_pos_dev = first + n;
    } else {
        if (ensure_host_tables(c, set0 + n)) return 1;
        if (wait_free) CK(cudaEventSynchronize(c->h_tab_free));
        float* h1 = c->h_tab1 + TAB1 * set0; float* h2 = c->h_tab2 + TAB2 * set0;
        if (c->gen_pos_host != first) { uint32_t from = c->gen_pos_host; if (from > first) { c->gen.reset(); from = 0; } for (uint32_t p = from; p < first; p++) c->gen.next_pass(h1, h2); }
        ctlb::generate_passes_threaded(c->gen, n, h1, h2, TAB1, TAB2);   // the n passes concurrently (start states by jump-ahead): bit-identical to n next_pass calls
        CK(cudaMemcpyAsync(d1, h1, TAB1 * 4 * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d2, h2, TAB2 * 4 * n, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(c->h_tab_free, st));
        c->gen_pos_host = first + n;
    }
    return 0;
}

int ctl_read_sample_tables(ctl_ctx* c, int table_set, float* d1, float* d2) {
    if (!c || !d1 || !d2) return set_err("null argument");
    if (table_set < 0 || table_set >= c->tab_cap) return set_err("no such table set");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(d1, c->d_tab1.p + TAB1 * table_set, TAB1 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d2, c->d_tab2.p + TAB2 * table_set, TAB2 * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------ intersect API


// Re-braided scenes (ctl_scene_set_rebraid): the traversal reports the pseudo-node it hit; API results name the instance, as the reference's would.
// res: n records of stride_words 32-bit words, node index at node_word; misses hold 0xffffffff and are left alone.
__global__ void __launch_bounds__(256) k_alias_nodes(uint32_t* __restrict__ res, int n, int stride_words, int node_word, const uint32_t* __restrict__ alias, uint32_t n_alias) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t v = res[(size_t)i * stride_words + node_word];
        if (v < n_alias) res[(size_t)i * stride_words + node_word] = alias[v];
    }
}
int ctl_intersect(ctl_ctx* c, int n, const void* d_rays, void* d_results, int any_hit, void* stream) {
    if (!c || !c->has_scene) return set_err("no scene uploaded");
    if (n < 0) return set_err("negative ray count");
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    unsigned* work = c->api_work.p + (c->api_seq++ % API_WORK_RING);
    CK(cudaMemsetAsync(work, 0, sizeof(unsigned), st));
    const int grid = grid_for(c, c->trav_blocks_per_sm);
    if (any_hit) launch_intersect<2, true, false>(c, grid, st, c->scene, (const float4*)d_rays, nullptr, n, work, nullptr, nullptr, nullptr, nullptr, d_results, nullptr);
    else launch_intersect<2, false, false>(c, grid, st, c->scene, (const float4*)d_rays, nullptr, n, work, nullptr, nullptr, nullptr, nullptr, d_results, nullptr);
    if (c->n_alias) k_alias_nodes<<<grid, 256, 0, st>>>((uint32_t*)d_results, n, 4, 1, c->d_node_alias.p, c->n_alias);   // traversalResult: {dist, nodeIdx, triIdx, bary}
    CK(cudaGetLastError());
    return 0;
}

int ctl_intersect_host(ctl_ctx* c, int n, const ctl_traversal_ray* rays, ctl_traversal_result* results, int any_hit) {
    if (!c || !c->has_scene) return set_err("no scene uploaded");
    if (n < 0) return set_err("negative ray count");
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    DevBuf<float4> dr, dres;
    CK(dr.ensure((size_t)n * 2)); CK(dres.ensure((size_t)n));
    CK(cudaMemcpyAsync(dr.p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    int rc = ctl_intersect(c, n, dr.p, dres.p, any_hit, nullptr);
    if (!rc) { cudaError_t e = cudaMemcpyAsync(results, dres.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = set_err(cudaGetErrorString(e)); }
    dr.release(); dres.release();
    return rc;
}

int ctl_trace_rays_host(ctl_ctx* c, int n, const ctl_traversal_ray* rays, ctl_trace_result* results, uint64_t counts[3]) {
    if (!c || !c->has_scene) return set_err("no scene uploaded");
    if (n < 0) return set_err("negative ray count");
    if (n == 0) { if (counts) counts[0] = counts[1] = counts[2] = 0; return 0; }
    CK(cudaSetDevice(c->device));
    DevBuf<float4> dr; DevBuf<float> dres; DevBuf<unsigned long long> dcnt;
    CK(dr.ensure((size_t)n * 2)); CK(dres.ensure((size_t)n * 5)); CK(dcnt.ensure(4));
    CK(cudaMemsetAsync(dcnt.p, 0, 32, c->stream));
    CK(cudaMemcpyAsync(dr.p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    unsigned* work = c->api_work.p + (c->api_seq++ % API_WORK_RING);
    CK(cudaMemsetAsync(work, 0, sizeof(unsigned), c->stream));
    const int grid = grid_for(c, c->trav_blocks_per_sm);
    if (counts) launch_intersect<3, false, true>(c, grid, c->stream, c->scene, dr.p, nullptr, n, work, nullptr, nullptr, nullptr, nullptr, dres.p, dcnt.p);
    else launch_intersect<3, false, false>(c, grid, c->stream, c->scene, dr.p, nullptr, n, work, nullptr, nullptr, nullptr, nullptr, dres.p, nullptr);
    if (c->n_alias) k_alias_nodes<<<grid, 256, 0, c->stream>>>((uint32_t*)dres.p, n, 5, 4, c->d_node_alias.p, c->n_alias);   // TraceResult: {dist, u, v, triIdx, nodeIdx}
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(results, dres.p, (size_t)n * 20, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long hc[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(hc, dcnt.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (counts) { counts[0] = hc[0]; counts[1] = hc[1]; counts[2] = hc[2]; }
    dr.release(); dres.release(); dcnt.release();
    return 0;
}

// ------------------------------------------------------------------ render pass
static int ensure_state(ctl_ctx* c, WaveLane& L, size_t n, cudaStream_t s) {
    CK(L.cf.ensure(n)); CK(L.cl.ensure(n)); CK(L.nor.ensure(n)); CK(L.px.ensure(n));
    CK(L.rays_a.ensure(2 * n)); CK(L.rays_b.ensure(2 * n)); CK(L.hit_a.ensure(n)); CK(L.hit_node.ensure(n));
    const size_t n_sh = n * (size_t)(c->regularization && c->scene.num_lights > 1 ? c->scene.num_lights : 1);   // Regularization: one shadow ray per light and vertex
    CK(L.sh_rays.ensure(2 * n_sh)); CK(L.sh_payload.ensure(n_sh)); CK(L.path_a.ensure(n)); CK(L.path_b.ensure(n));
    if (!c->stop_zero) CK(L.wo_prev.ensure(n));
    if (c->sort_mode == 2) CK(L.mat_cls.ensure(n));
    if (c->sort_mode == 2 || c->shade_mode == 1) { CK(L.mat_order.ensure(n)); CK(L.mat_hist.ensure(2 * MAT_CLASSES * (MAX_BOUNCES + 1))); }
    if (c->sort_mode == 1) {
        CK(L.rays_c.ensure(2 * n)); CK(L.path_c.ensure(n)); CK(L.sort_keys.ensure(n));
        if (!L.sort_hist.p) { CK(L.sort_hist.ensure(SORT_BUCKETS)); CK(L.sort_offsets.ensure(SORT_BUCKETS)); CK(cudaMemsetAsync(L.sort_hist.p, 0, SORT_BUCKETS * sizeof(unsigned), s)); }
    }
    CK(L.counters.ensure(CTR_TOTAL));
    return 0;
}

static void stage_mark(ctl_ctx* c, int kind) {
    if (!c->stage_timers) return;   // (never on together with OverlapWavefronts lanes: see ctl_render_frame_tiled)
    size_t i = c->stage_kind.size();
    if (i >= c->stage_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); c->stage_ev.push_back(e); }
    cudaEventRecord(c->stage_ev[i], c->stream);
    c->stage_kind.push_back(kind);
}


// lane / framed: ctl_comm_render_frame with "OverlapWavefronts" renders the wavefronts of a frame alternately on two lanes (two streams, two sets of
// wavefront buffers) so that the tail of one persistent traversal launch is filled by the head of the other lane's; a framed call leaves the
// accumulator clear, the sample tables (W.tab0) and the timing events to the frame.
static int render_window(ctl_ctx* c, int new_trace, const Window& W, int lane = 0, bool framed = false) {
    WaveLane& L = c->lanes[lane];
    const cudaStream_t s = lane ? c->lane_stream[lane] : c->stream;
    if (!c->has_scene) return set_err("no scene uploaded");
    if (!framed && c->f_submitted != c->f_acquired) return set_err("frames are in flight (ctl_submit_frame_tiled): ctl_acquire_frame them before other render calls");
    if (c->variance_buffer && (W.n_passes != 1 || W.mode != 0 || W.n_slots != c->w * c->h))
        return set_err("PixelVarianceBuffer=1 needs whole-image single-pass renders (ctl_render_pass with the full window, ctl_wavefront_pass)");
    if (W.n_slots <= 0) return 0;
    CK(cudaSetDevice(c->device));
    if (!framed) {
        CK(cudaEventRecord(c->ev_start, s));
        if (new_trace) {
            CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), s));
            c->passes_done = 0;
        }
        // sample tables of these passes: caller-supplied (single pass), generated on the device, or host XORWOW + H2D
        if (c->user_tables) { if (W.n_passes != 1) return set_err("caller-supplied sample tables cover exactly one pass"); }
        else if (generate_tables(c, c->passes_done, W.n_passes)) return 1;
        c->user_tables = false;
    }
    c->scene.d1 = c->d_tab1.p; c->scene.d2 = (const float2*)c->d_tab2.p;
    c->scene.img_w = c->w; c->scene.img_h = c->h;
    const size_t n_paths = (size_t)W.n_slots * W.n_passes;
    if (ensure_state(c, L, n_paths, s)) return 1;
    if (c->capture_bounce > 0) CK(c->capture.ensure(2 * n_paths));

    c->stage_kind.clear();
    CK(cudaMemsetAsync(L.counters.p, 0, CTR_TOTAL * sizeof(unsigned), s));
    // per-class shade launches: the staged kernel tags every hit with its material class; one launch per class present (a sort pass only when there are several)
    const bool by_class = c->shade_mode == 1 && c->sort_mode != 2 && c->trav_kernel == 2 && c->staged_ok && c->class_ok && c->class_mask != 0;
    int n_classes = 0, single_cls = -1;
    for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { n_classes++; single_cls = k; }
    const bool class_sort = by_class && n_classes > 1;
    if (c->sort_mode == 2 || class_sort) CK(cudaMemsetAsync(L.mat_hist.p, 0, 2 * MAT_CLASSES * (MAX_BOUNCES + 1) * sizeof(unsigned), s));
    if (c->instrumented) CK(cudaMemsetAsync(c->stats.p + 2, 0, 8 * sizeof(unsigned long long), s));
    unsigned* ctr = L.counters.p;
    PathState st = {L.cf.p, L.cl.p, L.nor.p, L.px.p, c->stop_zero ? nullptr : L.wo_prev.p};
    const int g_light = grid_for(c, c->shade_blocks_per_sm);
    const int g_trav = grid_for(c, c->trav_blocks_per_sm);
    uint32_t launches = 0;
    stage_mark(c, 0);
    k_generate<<<g_light, 256, 0, s>>>(c->scene, W, st, L.rays_a.p, L.path_a.p, ctr + CTR_Q + 0);
    launches++;
    ShadeParams P = {c->max_path_length, c->rr_start, c->direct, c->stop_zero};
    float4* rin = L.rays_a.p; float4* rout = L.rays_b.p; uint32_t* pin = L.path_a.p; uint32_t* pout = L.path_b.p;
    float4* rspare = L.rays_c.p; uint32_t* pspare = L.path_c.p;
    const bool fuse = c->fuse_traversal && c->direct && !c->instrumented && (c->trav_kernel == 0 || (c->trav_kernel == 2 && c->staged_ok));
    const bool regu = c->regularization != 0;
    if (regu && c->max_path_length > 255) return set_err("Regularization needs MaxPathLength <= 255");
    if (regu && c->sort_mode != 0) return set_err("Regularization with SortMode != 0 is not supported");
    const int n_iter = c->max_path_length + (regu ? 1 : 0);   // Regularization: the ray past the last vertex is traced (and counted), never shaded (Integrators/PathTracer.cu:125)
    for (int b = 0; b < n_iter; b++) {
        const bool shade_this = b < c->max_path_length;
        stage_mark(c, 1);
        if (c->capture_bounce == b + 1) {
            CK(cudaMemcpyAsync(c->capture.p, rin, 32 * n_paths, cudaMemcpyDeviceToDevice, s));
            CK(cudaMemcpyAsync(c->d_captured_n.p, ctr + CTR_Q + b, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
        }
        if (fuse && b > 0) { // shadow rays of bounce b-1 + extension rays of bounce b in one persistent launch
            if (c->trav_kernel == 2 && c->staged_ok) {
                const TravOut out = {L.hit_a.p, L.hit_node.p, L.sh_payload.p, L.cl.p, nullptr, L.sh_rays.p, 0, nullptr, class_sort ? L.mat_hist.p + 2 * MAT_CLASSES * b : nullptr};
                launch_staged<4, false, false>(c, s, rin, ctr + CTR_Q + b, ctr + CTR_SH + b - 1, 0, ctr + CTR_WORK + 2 * b, out, nullptr);
            } else
            k_intersect_fused<<<g_trav, 128, 0, s>>>(c->scene, c->tune_p, rin, ctr + CTR_Q + b, L.sh_rays.p, ctr + CTR_SH + b - 1, ctr + CTR_WORK + 2 * b,
                                                              L.hit_a.p, L.hit_node.p, L.sh_payload.p, L.cl.p);
        }
        else if (c->instrumented) launch_intersect<0, false, true>(c, g_trav, s, c->scene, rin, ctr + CTR_Q + b, 0, ctr + CTR_WORK + 2 * b, L.hit_a.p, L.hit_node.p, nullptr, nullptr, nullptr, c->stats.p + 2,
                                                                   class_sort ? L.mat_hist.p + 2 * MAT_CLASSES * b : nullptr);
        else launch_intersect<0, false, false>(c, g_trav, s, c->scene, rin, ctr + CTR_Q + b, 0, ctr + CTR_WORK + 2 * b, L.hit_a.p, L.hit_node.p, nullptr, nullptr, nullptr, nullptr,
                                               class_sort ? L.mat_hist.p + 2 * MAT_CLASSES * b : nullptr);
        stage_mark(c, 2);
        if (!shade_this) { stage_mark(c, 3); launches++; break; }   // (its shadow queue is empty: shade launch b-1 was the last)
        const bool sort_next = c->sort_mode == 1 && b + 1 < c->max_path_length;
        const uint32_t* order = nullptr;
        if (c->sort_mode == 2) { // group this bounce's hits by material class before shading
            unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            k_matsort_classify<<<g_light, 256, 0, s>>>(c->scene, ctr + CTR_Q + b, L.hit_a.p, L.hit_node.p, L.mat_cls.p, hist);
            k_matsort_scatter<<<g_light, 256, 0, s>>>(ctr + CTR_Q + b, L.mat_cls.p, hist, hist + MAT_CLASSES, L.mat_order.p);
            order = L.mat_order.p; launches += 2;
        }
        if (class_sort) { // group the hit records by the class bits the traversal kernel left in them
            unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            k_class_scatter<<<g_light, 256, 0, s>>>(ctr + CTR_Q + b, L.hit_a.p, hist, hist + MAT_CLASSES, L.mat_order.p);
            order = L.mat_order.p; launches++;
        }
        Queues Q = {rin, pin, rout, pout, L.hit_a.p, L.hit_node.p, L.sh_rays.p, L.sh_payload.p, sort_next ? L.sort_keys.p : nullptr, sort_next ? L.sort_hist.p : nullptr, order};
        if (class_sort) {
            const unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            if (c->shade_concurrent) {   // the class launches are independent (disjoint segments of the hit queue, atomic appends): each on its own stream, joined before the next stage
                if (!c->ev_cls_fork[lane]) {
                    CK(cudaEventCreateWithFlags(&c->ev_cls_fork[lane], cudaEventDisableTiming));
                    for (int j = 0; j < 3; j++) { CK(cudaStreamCreateWithFlags(&c->cls_stream[lane][j], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->ev_cls_done[lane][j], cudaEventDisableTiming)); }
                }
                CK(cudaEventRecord(c->ev_cls_fork[lane], s));
                int j = -1;
                for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) {
                    const cudaStream_t cs = j < 0 ? s : c->cls_stream[lane][j];
                    if (j >= 0) CK(cudaStreamWaitEvent(cs, c->ev_cls_fork[lane], 0));
                    launch_shade(regu, k, g_light, cs, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, hist); launches++;
                    if (j >= 0) { CK(cudaEventRecord(c->ev_cls_done[lane][j], cs)); CK(cudaStreamWaitEvent(s, c->ev_cls_done[lane][j], 0)); }
                    j++;
                }
            } else
            for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { launch_shade(regu, k, g_light, s, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, hist); launches++; }
            launches--;
        }
        else launch_shade(regu, by_class ? single_cls : -1, g_light, s, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, nullptr);
        if (sort_next) { // counting sort of the next bounce's extension queue by (octant, origin cell)
            stage_mark(c, 4);
            k_sort_scan<<<1, 1024, 0, s>>>(L.sort_hist.p, L.sort_offsets.p);
            k_sort_scatter<<<g_light, 256, 0, s>>>(ctr + CTR_Q + b + 1, L.sort_keys.p, rout, pout, L.sort_offsets.p, rspare, pspare);
            std::swap(rout, rspare); std::swap(pout, pspare);
            launches += 2;
        }
        stage_mark(c, 3);
        if (c->direct && (!fuse || b + 1 == n_iter)) {
            if (c->instrumented) launch_intersect<1, true, true>(c, g_trav, s, c->scene, L.sh_rays.p, ctr + CTR_SH + b, 0, ctr + CTR_WORK + 2 * b + 1, nullptr, nullptr, L.sh_payload.p, L.cl.p, nullptr, c->stats.p + 6);
            else launch_intersect<1, true, false>(c, g_trav, s, c->scene, L.sh_rays.p, ctr + CTR_SH + b, 0, ctr + CTR_WORK + 2 * b + 1, nullptr, nullptr, L.sh_payload.p, L.cl.p, nullptr, nullptr);
            launches++;
        }
        launches += 2;
        std::swap(rin, rout); std::swap(pin, pout);
    }
    stage_mark(c, 4);
    k_finish<<<g_light, 256, 0, s>>>((int)n_paths, st, c->lane_accum[lane] ? c->lane_accum[lane] : c->accum, c->w, c->h);   // (lane_accum: a frame in flight owns its accumulator)
    k_tally<<<1, 32, 0, s>>>(ctr + CTR_Q, ctr + CTR_SH, n_iter, c->stats.p, c->stats.p + 1);
    launches += 2;
    stage_mark(c, 5);
    CK(cudaGetLastError());
    c->n_launches = launches;
    if (framed) return 0;
    if (ctl_variance_after_pass(c, new_trace != 0)) return 1;
    CK(cudaEventRecord(c->ev_stop, s));
    c->events_recorded = true;
    c->passes_done += W.n_passes;
    return 0;
}

int ctl_render_pass(ctl_ctx* c, int new_trace, int x0, int y0, int x1, int y1) {
    if (!c) return set_err("null context");
    if (x0 < 0 || y0 < 0 || x1 > c->w || y1 > c->h || x1 < x0 || y1 < y0) return set_err("pixel window outside the image");
    Window W; memset(&W, 0, sizeof(W));
    W.mode = 0; W.x0 = x0; W.y0 = y0; W.x1 = x1; W.y1 = y1; W.n_slots = (x1 - x0) * (y1 - y0); W.n_passes = 1;
    return render_window(c, new_trace, W);
}

static int tiled_window(ctl_ctx* c, Window& W, int n_passes, int tile_w, int tile_h, int part, int n_parts) {
    if (n_passes < 1 || n_passes > 4096) return set_err("n_passes out of range [1,4096]");
    if (tile_w <= 0 || tile_h <= 0 || n_parts <= 0 || part < 0 || part >= n_parts) return set_err("invalid tiling");
    memset(&W, 0, sizeof(W));
    W.mode = 1; W.tile_w = tile_w; W.tile_h = tile_h; W.part = part; W.n_parts = n_parts; W.n_passes = n_passes;
    W.tiles_x = (c->w + tile_w - 1) / tile_w; W.tiles_y = (c->h + tile_h - 1) / tile_h;
    W.warp_blocks = c->warp_blocks && tile_w % 8 == 0 && tile_h % 4 == 0;
    const int n_tiles = W.tiles_x * W.tiles_y;
    const int n_local = n_tiles > part ? (n_tiles - part + n_parts - 1) / n_parts : 0;
    W.n_slots = n_local * tile_w * tile_h;
    if ((size_t)W.n_slots * n_passes > 0x7fffffffull / 2) return set_err("batch too large: reduce n_passes");
    return 0;
}

int ctl_render_passes_tiled(ctl_ctx* c, int new_trace, int n_passes, int tile_w, int tile_h, int part, int n_parts) {
    if (!c) return set_err("null context");
    Window W;
    if (tiled_window(c, W, n_passes, tile_w, tile_h, part, n_parts)) return 1;
    if (W.n_slots == 0) { // nothing to trace on this part: still honour the clear and advance the pass counter / sample stream
        CK(cudaSetDevice(c->device));
        CK(cudaEventRecord(c->ev_start, c->stream));
        if (new_trace) { CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream)); c->passes_done = 0; }
        CK(cudaMemsetAsync(c->stats.p, 0, sizeof(unsigned long long), c->stream));   // rays of the last pass: none
        CK(cudaEventRecord(c->ev_stop, c->stream));
        c->events_recorded = true; c->n_launches = 0;
        c->passes_done += n_passes;
        return 0;
    }
    return render_window(c, new_trace, W);
}

// A frame that is ONE wavefront, rendered as two interleaved half-wavefronts A and B (half the passes each, lanes[0] / lanes[1]) on the context's stream:
//     T(A,0) T(B,0) S(A,0) T(A,1) S(B,0) T(B,1) S(A,1) ... T(B,last) S(A,last) Tsh(A) S(B,last) Tsh(B) finish
// Every traversal launch first finishes the rays its predecessor suspended when its queue ran dry (device/traverse_handover.cuh), so no launch drains
// except the last.  Same paths, same hits, same ray counts as the one-wavefront frame.
static int render_frame_handover(ctl_ctx* c, Window W, int n_half) {
    const cudaStream_t s = c->stream;
    const int mpl = c->max_path_length;
    const size_t n_paths = (size_t)W.n_slots * n_half;
    const size_t smem = staged_smem_bytes(c);
    const int grid = staged_grid(c), lanes_resident = grid * c->staged_threads;
    for (int h = 0; h < 2; h++) { if (ensure_state(c, c->lanes[h], n_paths, s)) return 1; CK(c->ho_buf[h].ensure((size_t)lanes_resident * HO_WORDS)); }
    CK(c->ho_cnt.ensure(2 * (size_t)MAX_BOUNCES + 8));
    static size_t attr_set_dev[64] = {};
    if (smem > attr_set_dev[c->device & 63]) { CK(cudaFuncSetAttribute(k_intersect_handover<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set_dev[c->device & 63] = smem; }
    CK(cudaMemsetAsync(c->ho_cnt.p, 0, (2 * (size_t)MAX_BOUNCES + 8) * sizeof(unsigned), s));
    c->scene.d1 = c->d_tab1.p; c->scene.d2 = (const float2*)c->d_tab2.p; c->scene.img_w = c->w; c->scene.img_h = c->h;
    const bool by_class = c->shade_mode == 1 && c->class_ok && c->class_mask != 0;
    int n_classes = 0, single_cls = -1;
    for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { n_classes++; single_cls = k; }
    const bool class_sort = by_class && n_classes > 1;
    const int g_light = grid_for(c, c->shade_blocks_per_sm);
    const ShadeParams P = {c->max_path_length, c->rr_start, c->direct, c->stop_zero};
    struct Half { WaveLane* L; unsigned* ctr; PathState st; float4 *rin, *rout; uint32_t *pin, *pout; } H2[2];
    uint32_t launches = 0;
    for (int h = 0; h < 2; h++) {
        WaveLane& L = c->lanes[h];
        CK(cudaMemsetAsync(L.counters.p, 0, CTR_TOTAL * sizeof(unsigned), s));
        if (class_sort) CK(cudaMemsetAsync(L.mat_hist.p, 0, 2 * MAT_CLASSES * (MAX_BOUNCES + 1) * sizeof(unsigned), s));
        H2[h] = {&L, L.counters.p, {L.cf.p, L.cl.p, L.nor.p, L.px.p, c->stop_zero ? nullptr : L.wo_prev.p}, L.rays_a.p, L.rays_b.p, L.path_a.p, L.path_b.p};
        Window Wh = W; Wh.n_passes = n_half; Wh.tab0 = h * n_half;
        k_generate<<<g_light, 256, 0, s>>>(c->scene, Wh, H2[h].st, H2[h].rin, H2[h].pin, H2[h].ctr + CTR_Q + 0);
        launches++;
    }
    auto shade = [&](int h, int b) {   // S(h, b): the bounce's hit records -> next extension queue + shadow queue
        Half& X = H2[h]; unsigned* ctr = X.ctr; WaveLane& L = *X.L;
        const uint32_t* order = nullptr;
        if (class_sort) {
            unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            k_class_scatter<<<g_light, 256, 0, s>>>(ctr + CTR_Q + b, L.hit_a.p, hist, hist + MAT_CLASSES, L.mat_order.p);
            order = L.mat_order.p; launches++;
        }
        Queues Q = {X.rin, X.pin, X.rout, X.pout, L.hit_a.p, L.hit_node.p, L.sh_rays.p, L.sh_payload.p, nullptr, nullptr, order};
        if (class_sort) {
            const unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { launch_shade(false, k, g_light, s, c->scene, P, X.st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, hist); launches++; }
        } else { launch_shade(false, by_class ? single_cls : -1, g_light, s, c->scene, P, X.st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, nullptr); launches++; }
        std::swap(X.rin, X.rout); std::swap(X.pin, X.pout);
    };
    TravOut prev_out = {}; const float4* prev_rays = nullptr;
    int seq = 0;
    auto trav = [&](int h, int b, bool ext, bool last) {   // T(h, b): [extension rays of bounce b] + [shadow rays of bounce b-1]; resumes launch seq-1's suspended rays
        Half& X = H2[h]; unsigned* ctr = X.ctr; WaveLane& L = *X.L;
        const TravOut out = {L.hit_a.p, L.hit_node.p, L.sh_payload.p, L.cl.p, nullptr, L.sh_rays.p, 0, nullptr, (class_sort && ext) ? L.mat_hist.p + 2 * MAT_CLASSES * b : nullptr};
        HandOver HO;
        HO.resume = seq > 0 ? c->ho_buf[(seq - 1) & 1].p : nullptr; HO.n_resume = seq > 0 ? c->ho_cnt.p + (seq - 1) : nullptr;
        HO.alt = prev_out; HO.alt_rays = prev_rays;
        HO.suspend = last ? nullptr : c->ho_buf[seq & 1].p; HO.n_suspend = last ? nullptr : c->ho_cnt.p + seq; HO.drain_iters = c->handover_drain;
        const unsigned* n_ext_ptr = ext ? ctr + CTR_Q + b : nullptr;
        const unsigned* n_sh_ptr = ext ? (b > 0 ? ctr + CTR_SH + b - 1 : nullptr) : ctr + CTR_SH + b;
        unsigned* work = ext ? ctr + CTR_WORK + 2 * b : ctr + CTR_WORK + 2 * b + 1;
        k_intersect_handover<true><<<grid, c->staged_threads, smem, s>>>(c->scene, c->staged, c->tune, X.rin, n_ext_ptr, n_sh_ptr, work, out, HO);
        prev_out = out; prev_rays = X.rin; seq++; launches++;
    };
    int item_h = -1, item_b = -1;   // the (half, bounce) of the previous traversal launch: shaded after the current one
    for (int b = 0; b < mpl; b++) for (int h = 0; h < 2; h++) {
        trav(h, b, true, false);
        if (item_h >= 0) shade(item_h, item_b);
        item_h = h; item_b = b;
    }
    // item = (B, mpl-1) is still to be shaded; the shadow rays of the last bounce remain: Tsh(A) resumes B's last stragglers, Tsh(B) resumes A's and drains
    trav(0, mpl - 1, false, false);
    shade(item_h, item_b);
    trav(1, mpl - 1, false, true);
    for (int h = 0; h < 2; h++) {
        k_finish<<<g_light, 256, 0, s>>>((int)n_paths, H2[h].st, c->accum, c->w, c->h);
        k_tally<<<1, 32, 0, s>>>(H2[h].ctr + CTR_Q, H2[h].ctr + CTR_SH, mpl, c->stats.p, c->stats.p + 1, h);
        launches += 2;
    }
    CK(cudaGetLastError());
    c->n_launches = launches;
    return 0;
}

// One wavefront of a frame with straggler deferral ("DeferStragglers", device/traverse_handover.cuh: k_intersect_defer): every traversal launch moves the
// rays it has not finished a few iterations after its queue ran dry into the next launch (front of the next bounce's queues + a record of the lane state);
// their paths run up to max_lag bounces behind, the shade launches skip deferred hit records, and max_lag extra iterations at the end serve the paths that
// lag.  Counters: DC[2b] = deferred entries at the front of extension queue b, DC[2b+1] = at the front of the shadow queue traced by launch b.
__global__ void k_copy2_u32(unsigned* d0, const unsigned* s0, unsigned* d1, const unsigned* s1) { *d0 = *s0; *d1 = *s1; }
static int render_wavefront_deferred(ctl_ctx* c, const Window& W, int lane) {
    WaveLane& L = c->lanes[lane];
    const cudaStream_t s = lane ? c->lane_stream[lane] : c->stream;
    const int mpl = c->max_path_length, n_iter = mpl + c->defer_max_lag;
    const size_t n_paths = (size_t)W.n_slots * W.n_passes;
    if (ensure_state(c, L, n_paths, s)) return 1;
    CK(c->df_sh_rays[lane].ensure(2 * n_paths)); CK(c->df_sh_payload[lane].ensure(n_paths));
    const size_t smem = staged_smem_bytes(c);
    const int grid = staged_grid(c), lanes_resident = grid * c->staged_threads;
    const int rec_lanes = c->n_lanes > 1 ? c->n_lanes : 1;
    CK(c->ho_buf[0].ensure((size_t)lanes_resident * HO_WORDS * rec_lanes)); CK(c->ho_buf[1].ensure((size_t)lanes_resident * HO_WORDS * rec_lanes));
    const size_t DC_LANE = 4 * (size_t)MAX_BOUNCES + 16;   // per lane: [0, 2 MAX_BOUNCES + 8): DC; then n_suspend per launch
    CK(c->df_cnt.ensure(DC_LANE * MAX_LANES));
    uint32_t* rec[2] = {c->ho_buf[0].p + (size_t)lane * lanes_resident * HO_WORDS, c->ho_buf[1].p + (size_t)lane * lanes_resident * HO_WORDS};
    unsigned* DC = c->df_cnt.p + DC_LANE * lane; unsigned* NS = DC + 2 * MAX_BOUNCES + 8;
    static size_t attr_set_dev[64] = {};
    if (smem > attr_set_dev[c->device & 63]) { CK(cudaFuncSetAttribute(k_intersect_defer<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set_dev[c->device & 63] = smem; }
    CK(cudaMemsetAsync(DC, 0, DC_LANE * sizeof(unsigned), s));
    c->scene.d1 = c->d_tab1.p; c->scene.d2 = (const float2*)c->d_tab2.p; c->scene.img_w = c->w; c->scene.img_h = c->h;
    const bool by_class = c->shade_mode == 1 && c->class_ok && c->class_mask != 0;
    int n_classes = 0, single_cls = -1;
    for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { n_classes++; single_cls = k; }
    const bool class_sort = by_class && n_classes > 1;
    const int g_light = grid_for(c, c->shade_blocks_per_sm);
    const ShadeParams P = {mpl, c->rr_start, c->direct, c->stop_zero};
    unsigned* ctr = L.counters.p;
    CK(cudaMemsetAsync(ctr, 0, CTR_TOTAL * sizeof(unsigned), s));
    if (class_sort) CK(cudaMemsetAsync(L.mat_hist.p, 0, 2 * MAT_CLASSES * (MAX_BOUNCES + 1) * sizeof(unsigned), s));
    PathState st = {L.cf.p, L.cl.p, L.nor.p, L.px.p, c->stop_zero ? nullptr : L.wo_prev.p};
    float4 *rin = L.rays_a.p, *rout = L.rays_b.p; uint32_t *pin = L.path_a.p, *pout = L.path_b.p;
    float4* shr[2] = {L.sh_rays.p, c->df_sh_rays[lane].p}; float4* shp[2] = {L.sh_payload.p, c->df_sh_payload[lane].p};   // shadow queue of bounce b lives in buffer b & 1
    uint32_t launches = 0;
    k_generate<<<g_light, 256, 0, s>>>(c->scene, W, st, rin, pin, ctr + CTR_Q + 0);
    launches++;
    for (int b = 0; b <= n_iter; b++) {   // launch b: extension queue b (none at b == n_iter) + shadow queue b-1; the last launch finishes everything
        const bool ext = b < n_iter, last = b == n_iter;
        const TravOut out = {L.hit_a.p, L.hit_node.p, b > 0 ? shp[(b - 1) & 1] : nullptr, L.cl.p, nullptr, b > 0 ? shr[(b - 1) & 1] : nullptr, 0, nullptr, (class_sort && ext) ? L.mat_hist.p + 2 * MAT_CLASSES * b : nullptr};
        Defer D;
        D.resume = rec[(b + 1) & 1]; D.n_resume = b > 0 ? NS + (b - 1) : nullptr; D.k_ext = ext ? DC + 2 * b : nullptr; D.k_sh = b > 0 ? DC + 2 * b + 1 : nullptr;
        D.suspend = last ? nullptr : rec[b & 1]; D.n_suspend = last ? nullptr : NS + b;
        D.paths_in = pin; D.next_rays = rout; D.next_paths = pout; D.next_ext_ctr = ctr + CTR_Q + b + 1;
        D.next_sh_rays = shr[b & 1]; D.next_sh_payload = shp[b & 1]; D.next_sh_ctr = ctr + CTR_SH + b;
        D.drain_iters = c->handover_drain; D.max_lag = (b + 1 < n_iter) ? c->defer_max_lag : 0;   // (an extension ray deferred by launch b is traced by launch b + 1: the last one that takes extension rays is n_iter - 1)
        k_intersect_defer<true><<<grid, c->staged_threads, smem, s>>>(c->scene, c->staged, c->tune, rin, ext ? ctr + CTR_Q + b : nullptr, b > 0 ? ctr + CTR_SH + b - 1 : nullptr, ctr + CTR_WORK + 2 * b, out, D);
        launches++;
        if (last) break;
        // what the launch deferred sits at the front of the next queues: remember how many before the shade launches append behind them
        k_copy2_u32<<<1, 1, 0, s>>>(DC + 2 * (b + 1), ctr + CTR_Q + b + 1, DC + 2 * (b + 1) + 1, ctr + CTR_SH + b);
        launches++;
        const uint32_t* order = nullptr;
        if (class_sort) {
            unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            k_class_scatter<<<g_light, 256, 0, s>>>(ctr + CTR_Q + b, L.hit_a.p, hist, hist + MAT_CLASSES, L.mat_order.p);
            order = L.mat_order.p; launches++;
        }
        Queues Q = {rin, pin, rout, pout, L.hit_a.p, L.hit_node.p, shr[b & 1], shp[b & 1], nullptr, nullptr, order};
        if (class_sort) {
            const unsigned* hist = L.mat_hist.p + 2 * MAT_CLASSES * b;
            for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { launch_shade(false, k, g_light, s, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, hist); launches++; }
        } else { launch_shade(false, by_class ? single_cls : -1, g_light, s, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, nullptr); launches++; }
        std::swap(rin, rout); std::swap(pin, pout);
    }
    k_finish<<<g_light, 256, 0, s>>>((int)n_paths, st, c->accum, c->w, c->h);
    k_tally<<<1, 32, 0, s>>>(ctr + CTR_Q, ctr + CTR_SH, n_iter, c->stats.p, c->stats.p + 1, 0, DC);
    launches += 2;
    CK(cudaGetLastError());
    c->n_launches = launches;
    return 0;
}

// One progressive frame (a new trace): `spp` passes on the tiles of `part`, `batch` passes fused per wavefront.  With "OverlapWavefronts" (default) the
// frame's wavefronts alternate between two lanes -- two streams with their own wavefront buffers -- and a frame that is ONE wavefront is cut into two
// half-batches: every traversal launch is a persistent kernel whose last rays leave most of the SMs idle (nine tails per wavefront; at 1/8 of the image
// per GPU they are ~10 % of the frame), and the other lane's launch moves into the blocks that drain first.  Paths are the same paths (a path only
// depends on its pixel and pass); only the order of the float atomics into PixelData differs, as between any two runs.
int ctl_render_frame_tiled(ctl_ctx* c, int spp, int batch, int tile_w, int tile_h, int part, int n_parts) {
    if (!c) return set_err("null context");
    if (spp < 1 || batch < 1 || spp % batch) return set_err("spp must be a positive multiple of batch");
    if (c->f_submitted != c->f_acquired) return set_err("frames are in flight (ctl_submit_frame_tiled): ctl_acquire_frame them before other render calls");
    // "OverlapWavefronts": 0 = never; 1 (default) = when the frame has several wavefronts anyway (spp > batch: the lanes come for free); 2 = also cut the batches
    // of a frame with fewer wavefronts than lanes (measured: what the overlap gains, the extra launches lose)
    const bool plain = !c->overlap || c->stage_timers || c->instrumented || c->capture_bounce > 0 || c->variance_buffer || c->user_tables || c->sort_mode != 0 || spp > 128 ||
                       c->n_lanes < 2 || (c->overlap == 1 ? spp == batch : (spp == batch && (batch & 1)));
    const bool ho = c->handover && spp == batch && !(batch & 1) && c->has_scene && c->direct && c->fuse_traversal && c->trav_kernel == 2 && c->staged_ok && !c->staged.tl_nodes && !c->staged.ray_tma &&
                    !c->regularization && !c->stage_timers && !c->instrumented && c->capture_bounce <= 0 && !c->variance_buffer && !c->user_tables && c->sort_mode == 0 && c->max_path_length < MAX_BOUNCES;
    const bool df = c->defer && c->has_scene && c->direct && c->fuse_traversal && c->trav_kernel == 2 && c->staged_ok && !c->staged.tl_nodes && !c->staged.ray_tma && !c->regularization &&
                    !c->stage_timers && !c->instrumented && c->capture_bounce <= 0 && !c->variance_buffer && !c->user_tables && c->sort_mode == 0 && c->max_path_length + c->defer_max_lag + 2 < MAX_BOUNCES;
    if (df && plain) {   // the frame's wavefronts one after the other, each with straggler deferral between its traversal launches
        Window W;
        if (tiled_window(c, W, batch, tile_w, tile_h, part, n_parts)) return 1;
        if (W.n_slots > 0) {
            CK(cudaSetDevice(c->device));
            CK(cudaEventRecord(c->ev_start, c->stream));
            CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream));
            c->passes_done = 0;
            if (generate_tables(c, 0, spp)) return 1;
            uint32_t launches = 0;
            for (int p = 0; p < spp; p += batch) { W.tab0 = p; if (render_wavefront_deferred(c, W, 0)) return 1; launches += c->n_launches; }
            c->n_launches = launches;
            CK(cudaEventRecord(c->ev_stop, c->stream));
            c->events_recorded = true; c->passes_done = (uint32_t)spp;
            return 0;
        }
    }
    if (ho) {   // a one-wavefront frame as two half-wavefronts whose traversal launches hand their unfinished rays over
        Window W;
        if (tiled_window(c, W, batch / 2, tile_w, tile_h, part, n_parts)) return 1;
        if (W.n_slots > 0) {
            CK(cudaSetDevice(c->device));
            CK(cudaEventRecord(c->ev_start, c->stream));
            CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream));
            c->passes_done = 0;
            if (generate_tables(c, 0, spp)) return 1;
            if (render_frame_handover(c, W, batch / 2)) return 1;
            CK(cudaEventRecord(c->ev_stop, c->stream));
            c->events_recorded = true; c->passes_done = (uint32_t)spp;
            return 0;
        }
    }
    if (plain) {
        for (int p = 0; p < spp; p += batch)
            if (ctl_render_passes_tiled(c, p == 0, batch, tile_w, tile_h, part, n_parts)) return 1;
        return 0;
    }
    const int n_lanes = c->n_lanes;
    if (c->overlap >= 2) while (spp / batch < n_lanes && batch % 2 == 0) batch /= 2;   // fewer wavefronts than lanes: cut the batches
    Window W;
    if (tiled_window(c, W, batch, tile_w, tile_h, part, n_parts)) return 1;
    if (!c->has_scene) return set_err("no scene uploaded");
    CK(cudaSetDevice(c->device));
    if (!c->ev_fork) CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int k = 1; k < n_lanes; k++) if (!c->lane_stream[k]) {
        CK(cudaStreamCreateWithFlags(&c->lane_stream[k], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    if (!c->tab_stream) { CK(cudaStreamCreateWithFlags(&c->tab_stream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->ev_tab, cudaEventDisableTiming)); }
    CK(cudaEventRecord(c->ev_start, c->stream));
    CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream));
    c->passes_done = 0;
    if (W.n_slots == 0) CK(cudaMemsetAsync(c->stats.p, 0, sizeof(unsigned long long), c->stream));
    else {
        // the frame keeps one table set per pass; each wavefront's sets are produced on the table stream just before the wavefront is enqueued, so that host-generated
        // tables (DeviceSampleTables = 0) of wavefront k+1 are computed while the device renders wavefront k, and no lane waits for another lane's work
        if (ensure_tables(c, spp)) return 1;
        if (!c->device_tables) { if (ensure_host_tables(c, spp)) return 1; CK(cudaEventSynchronize(c->h_tab_free)); }   // the previous frame's copies have left the pinned sets
        CK(cudaEventRecord(c->ev_fork, c->stream));
        CK(cudaStreamWaitEvent(c->tab_stream, c->ev_fork, 0));
        for (int k = 1; k < n_lanes; k++) CK(cudaStreamWaitEvent(c->lane_stream[k], c->ev_fork, 0));
        uint32_t launches = 0;
        for (int p = 0, i = 0; p < spp; p += batch, i++) {
            const int lane = i % n_lanes;
            if (generate_tables(c, (uint32_t)p, batch, p, c->tab_stream, false)) return 1;
            CK(cudaEventRecord(c->ev_tab, c->tab_stream));
            CK(cudaStreamWaitEvent(lane ? c->lane_stream[lane] : c->stream, c->ev_tab, 0));
            W.tab0 = p;
            if (df ? render_wavefront_deferred(c, W, lane) : render_window(c, 0, W, lane, true)) return 1;
            launches += c->n_launches;
        }
        c->n_launches = launches;
        for (int k = 1; k < n_lanes; k++) { CK(cudaEventRecord(c->ev_join[k], c->lane_stream[k])); CK(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0)); }
    }
    CK(cudaEventRecord(c->ev_stop, c->stream));
    c->events_recorded = true;
    c->passes_done = (uint32_t)spp;
    return 0;
}

// ------------------------------------------------------------------ frames in flight
// A renderer that produces a SEQUENCE of frames (animation, interactive progressive display, a render farm's queue) does not have to finish frame k before
// frame k+1 starts: the frames are independent.  ctl_submit_frame_tiled enqueues a whole frame -- accumulator clear, sample tables, every wavefront, and the
// reduce when there is a communicator -- on wavefront lane 1 + k % FramesInFlight (its own stream, wavefront buffers, accumulator, table sets and generator
// state); ctl_acquire_frame makes the context's stream wait for the OLDEST outstanding frame and makes its accumulator the context's.  Why: every traversal
// launch is a persistent kernel that ends with a drain (the scene's longest rays, ~0.45 ms whatever the launch size: DESIGN.md section 5); a frame that is
// one wavefront cannot hide its own drains (cutting it in two doubles the launches), but the NEXT frame's launches fill them -- the launches per frame stay
// the same.  Same paths, same images as ctl_render_frame_tiled (only the order of the float atomics differs, as between any two runs).
static int frame_tables(ctl_ctx* c, FrameSlot& F, int set0, int spp, cudaStream_t st) {
    float* d1 = c->d_tab1.p + TAB1 * set0; float* d2 = c->d_tab2.p + TAB2 * set0;
    if (c->device_tables) {   // a new trace: passes 0 .. spp-1 from the generator's start states, on this slot's own copy of the state
        CK(F.states.ensure((size_t)ctlb::kNumSeq * 6));
        CK(cudaMemcpyAsync(F.states.p, c->d_states0.p, (size_t)ctlb::kNumSeq * 6 * 4, cudaMemcpyDeviceToDevice, st));
        k_gen_tables<<<ctlb::kNumSeq / 128, 128, 0, st>>>(F.states.p, c->d_jump.p, spp, d1, (float2*)d2);
        CK(cudaGetLastError());
    } else {                  // the reference's UpdateKernel behaviour: host XORWOW, pinned sets of this slot, H2D on the lane's stream
        if (F.h_cap < spp) {
            if (F.h_free) CK(cudaEventSynchronize(F.h_free));
            if (F.h1) cudaFreeHost(F.h1); if (F.h2) cudaFreeHost(F.h2); F.h1 = F.h2 = nullptr; F.h_cap = 0;
            CK(cudaMallocHost((void**)&F.h1, TAB1 * 4 * spp)); CK(cudaMallocHost((void**)&F.h2, TAB2 * 4 * spp));
            F.h_cap = spp;
        }
        CK(cudaEventSynchronize(F.h_free));   // the copies of the frame that used this slot last have left the pinned sets
        c->gen.reset();
        ctlb::generate_passes_threaded(c->gen, spp, F.h1, F.h2, TAB1, TAB2);
        c->gen_pos_host = (uint32_t)spp;
        CK(cudaMemcpyAsync(d1, F.h1, TAB1 * 4 * spp, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d2, F.h2, TAB2 * 4 * spp, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(F.h_free, st));
    }
    return 0;
}

int ctl_submit_frame_tiled(ctl_ctx* c, int spp, int batch, int tile_w, int tile_h, int part, int n_parts) {
    if (!c) return set_err("null context");
    if (spp < 1 || batch < 1 || spp % batch || spp > 128) return set_err("spp must be a positive multiple of batch (<= 128)");
    if (!c->has_scene) return set_err("no scene uploaded");
    if (c->stage_timers || c->instrumented || c->capture_bounce > 0 || c->variance_buffer || c->user_tables || c->sort_mode != 0)
        return set_err("ctl_submit_frame_tiled: StageTimers / instrumentation / CaptureBounce / PixelVarianceBuffer / caller-supplied tables / SortMode need ctl_render_frame_tiled");
    if (c->f_submitted - c->f_acquired >= (unsigned long long)c->fif) return set_err("FramesInFlight frames are outstanding: ctl_acquire_frame first");
    Window W;
    if (tiled_window(c, W, batch, tile_w, tile_h, part, n_parts)) return 1;
    CK(cudaSetDevice(c->device));
    const int slot = (int)(c->f_submitted % (unsigned long long)c->fif), lane = slot + 1;
    FrameSlot& F = c->fslot[slot];
    if (!c->ev_fork) CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    if (!c->lane_stream[lane]) { CK(cudaStreamCreateWithFlags(&c->lane_stream[lane], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->ev_join[lane], cudaEventDisableTiming)); }
    if (!F.begin) { CK(cudaEventCreate(&F.begin)); CK(cudaEventCreate(&F.done)); CK(cudaEventCreateWithFlags(&F.ready, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&F.h_free, cudaEventDisableTiming)); }
    const cudaStream_t s = c->lane_stream[lane];
    if (c->fif * spp > c->tab_cap) {   // the table sets of every slot: grown only while no lane reads them
        for (int k = 1; k < MAX_LANES; k++) if (c->lane_stream[k]) CK(cudaStreamSynchronize(c->lane_stream[k]));
        if (ensure_tables(c, c->fif * spp)) return 1;
    }
    CK(F.accum.ensure((size_t)c->w * c->h * 7));
    // the lane starts after what the context's stream holds so far (the consumer of the frame that used this slot last, scene updates, ...)
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(s, c->ev_fork, 0));
    CK(cudaEventRecord(F.begin, s));
    CK(cudaMemsetAsync(F.accum.p, 0, (size_t)c->w * c->h * 7 * sizeof(float), s));
    c->lane_accum[lane] = F.accum.p;
    F.launches = 0; F.spp = (uint32_t)spp;
    if (W.n_slots > 0) {
        if (frame_tables(c, F, slot * spp, spp, s)) return 1;
        for (int p = 0; p < spp; p += batch) {
            W.tab0 = slot * spp + p;
            if (render_window(c, 0, W, lane, true)) { c->lane_accum[lane] = nullptr; return 1; }
            F.launches += c->n_launches;
        }
    }
    c->lane_accum[lane] = nullptr;   // (read at launch time only: the lanes of ctl_render_frame_tiled share the context's accumulator again)
    CK(cudaEventRecord(F.done, s));
    F.reduced = false;
    if (ctl_comm_reduce_slot(c, F)) return 1;   // with a communicator: ncclReduce of this frame's accumulator on the communication stream, after F.done
    c->f_submitted++;
    return 0;
}

int ctl_acquire_frame(ctl_ctx* c) {
    if (!c) return set_err("null context");
    if (c->f_acquired == c->f_submitted) return set_err("ctl_acquire_frame: no frame in flight");
    CK(cudaSetDevice(c->device));
    FrameSlot& F = c->fslot[(int)(c->f_acquired % (unsigned long long)c->fif)];
    CK(cudaStreamWaitEvent(c->stream, F.reduced ? F.ready : F.done, 0));
    c->accum = F.accum.p;                                                  // resolve / image pipeline / read-back calls now see this frame
    std::swap(c->ev_start, F.begin); std::swap(c->ev_stop, F.done);        // ctl_stats: device time of this frame on its lane (its wait on F.done is already enqueued)
    c->events_recorded = true; c->n_launches = F.launches; c->passes_done = F.spp;
    c->f_acquired++;
    return 0;
}

int ctl_frames_in_flight(ctl_ctx* c) { return c ? (int)(c->f_submitted - c->f_acquired) : 0; }

int ctl_render_pass_tiled(ctl_ctx* c, int new_trace, int tile_w, int tile_h, int part, int n_parts) {
    return ctl_render_passes_tiled(c, new_trace, 1, tile_w, tile_h, part, n_parts);
}

// ------------------------------------------------------------------ WavefrontPathTracer
__global__ void k_set_u32(unsigned* p, unsigned v) { *p = v; }

// == Tracer<true>::DoPass + WavefrontPathTracer::DoRender (Kernel/Tracer.h:209-248, Integrators/PseudoRealtime/WavefrontPathTracer.cu:166-191).
// One pass = one path per pixel through the DoubleRayBuffer-shaped queue: create -> { intersect primaries (+ last iteration's secondaries),
// iterate } x MaxPathLength.  Unlike the reference nothing crosses the host per bounce (it copies the queue struct to and from the device
// around every kernel, cu:175-188): the queue sizes stay in device counters and an empty iteration costs three empty launches.
// One WavefrontPathTracer pass on lane `lane` (stream st), drawing from sample-table set `table_set`.  framed: part of ctl_wavefront_frame (the frame owns the
// accumulator clear, the tables, the timing events and the pass counter).
static int wavefront_pass_on(ctl_ctx* c, int new_trace, int lane, cudaStream_t st, uint32_t pass_index, int table_set, bool framed) {
    WptLane& WL = c->wl[lane];
    if (!framed) {
        CK(cudaEventRecord(c->ev_start, st));
        if (new_trace) { CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), st)); c->passes_done = 0; }
        if (c->user_tables) c->user_tables = false;
        else if (generate_tables(c, pass_index, 1)) return 1;
    }
    c->scene.img_w = c->w; c->scene.img_h = c->h;
    DScene scene = c->scene;   // this pass's view of the scene: its own sample-table set
    scene.d1 = c->d_tab1.p + TAB1 * table_set; scene.d2 = (const float2*)(c->d_tab2.p + TAB2 * table_set);
    const size_t n = (size_t)c->w * c->h;
    if (n > 0x3fffffffull) return set_err("image too large for the wavefront queue");
    const int n_tiles = (int)((n + WPT_TILE - 1) / WPT_TILE);
    const int mpl = c->max_path_length;
    CK(WL.w_thr.ensure(n)); CK(WL.w_lxy.ensure(n)); CK(WL.w_df.ensure(n)); CK(WL.w_misc.ensure(n)); CK(WL.w_ray.ensure(2 * n)); CK(WL.w_res.ensure(n));
    for (int k = 0; k < 2; k++) { CK(WL.w_sec[k].ensure(2 * n)); CK(WL.w_sres[k].ensure(n)); }
    CK(WL.w_desc.ensure((size_t)mpl * (n_tiles + 1)));
    c->stage_kind.clear();
    CK(c->lanes[lane].counters.ensure(CTR_TOTAL));
    unsigned* ctr = c->lanes[lane].counters.p;
    CK(cudaMemsetAsync(ctr, 0, CTR_TOTAL * sizeof(unsigned), st));
    CK(cudaMemsetAsync(WL.w_desc.p, 0, (size_t)mpl * (n_tiles + 1) * sizeof(unsigned long long), st));
    if (c->instrumented) CK(cudaMemsetAsync(c->stats.p + 2, 0, 8 * sizeof(unsigned long long), st));
    const int g_light = grid_for(c, c->shade_blocks_per_sm), g_trav = grid_for(c, c->trav_blocks_per_sm);
    uint32_t launches = 0;
    stage_mark(c, 0);
    k_set_u32<<<1, 1, 0, st>>>(ctr + CTR_Q, (unsigned)n);
    WptBuf B = {WL.w_thr.p, WL.w_lxy.p, WL.w_df.p, WL.w_misc.p, WL.w_ray.p, WL.w_res.p, nullptr, nullptr};
    k_wpt_create<<<g_light, 256, 0, st>>>(scene, B, (int)n);
    launches += 2;
    for (int d = 0; d < mpl; d++) {
        stage_mark(c, 1);
        // FinishIteration (DoubleRayBuffer.h:84-112): the primaries and the secondary rays pushed by iteration d-1
        const bool have_sec = d > 0 && c->direct;
        if (c->instrumented) { // visit counts for the roofline (ctl_get_visit_counts: "extension" = primaries, "shadow" = secondaries): unfused, counting builds
            launch_intersect<2, false, true>(c, g_trav, st, scene, (const float4*)WL.w_ray.p, ctr + CTR_Q + d, 0, ctr + CTR_WORK + 2 * d, nullptr, nullptr, nullptr, nullptr, (void*)WL.w_res.p, c->stats.p + 2);
            launches++;
            if (have_sec) {
                stage_mark(c, 3);
                launch_intersect<2, true, true>(c, g_trav, st, scene, (const float4*)WL.w_sec[(d - 1) & 1].p, ctr + CTR_SH + d - 1, 0, ctr + CTR_WORK + 2 * d + 1, nullptr, nullptr, nullptr, nullptr,
                                                (void*)WL.w_sres[(d - 1) & 1].p, c->stats.p + 6);
                launches++;
            }
        } else if (have_sec && c->fuse_traversal && c->trav_kernel == 2 && c->staged_ok) {
            const TravOut out = {nullptr, nullptr, nullptr, nullptr, (void*)WL.w_res.p, (const float4*)WL.w_sec[(d - 1) & 1].p, 0, (void*)WL.w_sres[(d - 1) & 1].p};
            launch_staged<5, false, false>(c, st, (const float4*)WL.w_ray.p, ctr + CTR_Q + d, ctr + CTR_SH + d - 1, 0, ctr + CTR_WORK + 2 * d, out, nullptr);
            launches++;
        } else if (have_sec && c->fuse_traversal && c->trav_kernel == 0) {
            k_intersect_fused_api<<<g_trav, 128, 0, st>>>(scene, c->tune_p, (const float4*)WL.w_ray.p, ctr + CTR_Q + d, (const float4*)WL.w_sec[(d - 1) & 1].p, ctr + CTR_SH + d - 1,
                                                                  ctr + CTR_WORK + 2 * d, (void*)WL.w_res.p, (void*)WL.w_sres[(d - 1) & 1].p);
            launches++;
        } else {
            launch_intersect<2, false, false>(c, g_trav, st, scene, (const float4*)WL.w_ray.p, ctr + CTR_Q + d, 0, ctr + CTR_WORK + 2 * d, nullptr, nullptr, nullptr, nullptr, (void*)WL.w_res.p, nullptr);
            launches++;
            if (have_sec) {
                stage_mark(c, 3);
                launch_intersect<2, true, false>(c, g_trav, st, scene, (const float4*)WL.w_sec[(d - 1) & 1].p, ctr + CTR_SH + d - 1, 0, ctr + CTR_WORK + 2 * d + 1, nullptr, nullptr, nullptr, nullptr,
                                                 (void*)WL.w_sres[(d - 1) & 1].p, nullptr);
                launches++;
            }
        }
        stage_mark(c, 2);
        B.sec_out = WL.w_sec[d & 1].p; B.sec_res = WL.w_sres[(d - 1) & 1].p;
        const WptParams P = {d, (int)pass_index + 1, mpl, c->rr_start}; // m_uPassesDone++ precedes DoRender (Kernel/Tracer.h:231-232)
        unsigned long long* desc = WL.w_desc.p + (size_t)d * (n_tiles + 1);
        if (c->direct) k_wpt_iterate<true><<<n_tiles, WPT_TILE, 0, st>>>(scene, P, B, ctr + CTR_Q + d, ctr + CTR_Q + d + 1, ctr + CTR_SH + d, desc, n_tiles, c->accum);
        else k_wpt_iterate<false><<<n_tiles, WPT_TILE, 0, st>>>(scene, P, B, ctr + CTR_Q + d, ctr + CTR_Q + d + 1, ctr + CTR_SH + d, desc, n_tiles, c->accum);
        launches++;
    }
    stage_mark(c, 4);
    k_tally<<<1, 32, 0, st>>>(ctr + CTR_Q, ctr + CTR_SH, mpl, c->stats.p, c->stats.p + 1);
    launches++;
    stage_mark(c, 5);
    CK(cudaGetLastError());
    c->n_launches = launches;
    if (framed) return 0;
    if (ctl_variance_after_pass(c, new_trace != 0)) return 1;
    CK(cudaEventRecord(c->ev_stop, st));
    c->events_recorded = true;
    c->passes_done += 1;
    return 0;
}

int ctl_wavefront_pass(ctl_ctx* c, int new_trace) {
    if (!c) return set_err("null context");
    if (!c->has_scene) return set_err("no scene uploaded");
    if (c->f_submitted != c->f_acquired) return set_err("frames are in flight (ctl_submit_frame_tiled): ctl_acquire_frame them before other render calls");
    CK(cudaSetDevice(c->device));
    const uint32_t pass_index = (uint32_t)c->pass_phase + (uint32_t)c->pass_stride * (new_trace ? 0u : c->passes_done);   // which pass of the (possibly shared) frame this is
    return wavefront_pass_on(c, new_trace, 0, c->stream, pass_index, 0, false);
}

// One progressive frame of the WavefrontPathTracer: a new trace of `spp` passes.  Its passes are independent given their index (random numbers are keyed by
// pass and queue slot, results go to the accumulator by atomics), so with "OverlapWavefronts" they run on up to "OverlapLanes" streams with their own
// queues -- the drain of one pass's traversal launches is filled by another pass's.  Same passes, same paths as spp ctl_wavefront_pass calls.
int ctl_wavefront_frame(ctl_ctx* c, int spp) {
    if (!c) return set_err("null context");
    if (!c->has_scene) return set_err("no scene uploaded");
    if (c->f_submitted != c->f_acquired) return set_err("frames are in flight (ctl_submit_frame_tiled): ctl_acquire_frame them before other render calls");
    if (spp < 1 || spp > 4096) return set_err("spp out of range [1,4096]");
    CK(cudaSetDevice(c->device));
    const int n_lanes = c->n_lanes < spp ? c->n_lanes : spp;
    const bool plain = !c->overlap || n_lanes < 2 || spp > 128 || c->stage_timers || c->instrumented || c->variance_buffer || c->user_tables;
    if (plain) { for (int p = 0; p < spp; p++) if (ctl_wavefront_pass(c, p == 0)) return 1; return 0; }
    if (!c->ev_fork) CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int k = 1; k < n_lanes; k++) if (!c->lane_stream[k]) {
        CK(cudaStreamCreateWithFlags(&c->lane_stream[k], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    if (!c->tab_stream) { CK(cudaStreamCreateWithFlags(&c->tab_stream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->ev_tab, cudaEventDisableTiming)); }
    CK(cudaEventRecord(c->ev_start, c->stream));
    CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream));
    c->passes_done = 0;
    if (ensure_tables(c, spp)) return 1;
    if (!c->device_tables) { if (ensure_host_tables(c, spp)) return 1; CK(cudaEventSynchronize(c->h_tab_free)); }
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(c->tab_stream, c->ev_fork, 0));
    for (int k = 1; k < n_lanes; k++) CK(cudaStreamWaitEvent(c->lane_stream[k], c->ev_fork, 0));
    uint32_t launches = 0;
    for (int p = 0; p < spp; p++) {
        const int lane = p % n_lanes;
        const cudaStream_t st = lane ? c->lane_stream[lane] : c->stream;
        const uint32_t pass_index = (uint32_t)c->pass_phase + (uint32_t)c->pass_stride * (uint32_t)p;
        if (generate_tables(c, pass_index, 1, p, c->tab_stream, false)) return 1;
        CK(cudaEventRecord(c->ev_tab, c->tab_stream));
        CK(cudaStreamWaitEvent(st, c->ev_tab, 0));
        if (wavefront_pass_on(c, 0, lane, st, pass_index, p, true)) return 1;
        launches += c->n_launches;
    }
    c->n_launches = launches;
    for (int k = 1; k < n_lanes; k++) { CK(cudaEventRecord(c->ev_join[k], c->lane_stream[k])); CK(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0)); }
    CK(cudaEventRecord(c->ev_stop, c->stream));
    c->events_recorded = true;
    c->passes_done = (uint32_t)spp;
    return 0;
}

int ctl_synchronize(ctl_ctx* c) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int ctl_read_accum(ctl_ctx* c, ctl_pixel_data* out) {
    if (!c || !out) return set_err("null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->accum, (size_t)c->w * c->h * 7 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
void* ctl_accum_device_ptr(ctl_ctx* c) { return c ? (void*)c->accum : nullptr; }
int ctl_set_accum_device_ptr(ctl_ctx* c, void* p) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->accum = p ? (float*)p : c->own_accum.p;
    return 0;
}
void* ctl_stream(ctl_ctx* c) { return c ? (void*)c->stream : nullptr; }
int ctl_set_stream(ctl_ctx* c, void* stream) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return 0;
}

int ctl_stats(ctl_ctx* c, uint64_t* rays_last, float* seconds_last, uint64_t* rays_total, uint32_t* passes_done) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpy(h, c->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    float ms = 0.0f;
    if (c->events_recorded) CK(cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop));
    if (rays_last) *rays_last = h[0];
    if (rays_total) *rays_total = h[1];
    if (seconds_last) *seconds_last = ms * 1e-3f;
    if (passes_done) *passes_done = c->passes_done;
    return 0;
}

int ctl_stage_times(ctl_ctx* c, float ms[5], uint32_t* n_launches) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 5; i++) c->stage_ms[i] = 0;
    for (size_t i = 0; i + 1 < c->stage_kind.size(); i++) {
        float t = 0; CK(cudaEventElapsedTime(&t, c->stage_ev[i], c->stage_ev[i + 1]));
        int k = c->stage_kind[i]; if (k >= 0 && k < 5) c->stage_ms[k] += t;
    }
    if (ms) memcpy(ms, c->stage_ms, sizeof(c->stage_ms));
    if (n_launches) *n_launches = c->n_launches;
    return 0;
}

int ctl_set_instrumented(ctl_ctx* c, int on) { if (!c) return set_err("null context"); c->instrumented = on != 0; return 0; }
int ctl_get_visit_counts(ctl_ctx* c, uint64_t ext[4], uint64_t sh[4]) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long h[8]; std::vector<unsigned> ctr(CTR_TOTAL);
    CK(cudaMemcpy(h, c->stats.p + 2, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ctr.data(), c->lanes[0].counters.p, CTR_TOTAL * sizeof(unsigned), cudaMemcpyDeviceToHost));
    unsigned long long ne = 0, ns = 0;
    for (int b = 0; b < c->max_path_length; b++) { ne += ctr[CTR_Q + b]; ns += ctr[CTR_SH + b]; }
    if (ext) { ext[0] = h[0]; ext[1] = h[1]; ext[2] = h[2]; ext[3] = ne; }
    if (sh) { sh[0] = h[4]; sh[1] = h[5]; sh[2] = h[6]; sh[3] = ns; }
    return 0;
}
int ctl_get_queue_sizes(ctl_ctx* c, uint32_t* ext, uint32_t* sh, int n) {
    if (!c) return set_err("null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<unsigned> ctr(CTR_TOTAL);
    CK(cudaMemcpy(ctr.data(), c->lanes[0].counters.p, CTR_TOTAL * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (int b = 0; b < n && b < MAX_BOUNCES; b++) { if (ext) ext[b] = ctr[CTR_Q + b]; if (sh) sh[b] = ctr[CTR_SH + b]; }
    return 0;
}
#ifdef CTL_EXP_DROP_STRAGGLERS
int ctl_debug_counters(ctl_ctx* c, unsigned* out, int n) {   // experiment builds only: the raw per-bounce counter array of lane 0
    if (!c || !out) return set_err("null argument");
    CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, c->lanes[0].counters.p, (size_t)(n < CTR_TOTAL ? n : CTR_TOTAL) * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return 0;
}
#endif
int ctl_get_captured_rays(ctl_ctx* c, ctl_traversal_ray* host_out, int capacity) {
    if (!c) { set_err("null context"); return -1; }
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { set_err("cuda error"); return -1; }
    unsigned n = 0;
    if (cudaMemcpy(&n, c->d_captured_n.p, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) { set_err("cuda error"); return -1; }
    int m = (int)n < capacity ? (int)n : capacity;
    if (m > 0 && host_out && cudaMemcpy(host_out, c->capture.p, (size_t)m * 32, cudaMemcpyDeviceToHost) != cudaSuccess) { set_err("cuda error"); return -1; }
    return m;
}

} // extern "C"
