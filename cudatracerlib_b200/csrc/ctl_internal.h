// ctl_internal.h -- what the translation units behind the C ABI share: error text, device buffers, the context object.
// (ctl_api.cu: context + render + intersect; ctl_pipeline.cu: image pipeline / variance / NLM; ctl_bvh_gpu.cu: GPU BVH build; ctl_comm.cu: NCCL;
//  ctl_scene_api.cpp: host scenes.)  Kernels with external linkage live in exactly one of them each.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <stdexcept>
#include <memory>
#include "../../include/ctl_b200.h"
#include "scene_builder.h"
#include "sampler_tables.h"
#include "staging.h"
#include "device/traverse_handover.cuh"

using namespace ctld;

// error text of the calling thread (ctl_last_error); returns 1 so that `return ctl_set_err(...)` reads as the failure code
int ctl_set_err(const std::string& s);
#define set_err ctl_set_err
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return set_err(b_); } } while (0)
#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); set_err(b_); return nullptr; } } while (0)

struct ctl_scene { ctlb::SceneStorage S; };

const int MAX_BOUNCES = 256;
const int MAX_LANES = 8;   // wavefronts of a frame in flight at once (OverlapWavefronts / OverlapLanes)
const unsigned API_WORK_RING = 256;
enum { CTR_Q = 0, CTR_SH = MAX_BOUNCES + 1, CTR_WORK = 2 * (MAX_BOUNCES + 1), CTR_TOTAL = 4 * (MAX_BOUNCES + 1) };

template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* h, size_t count) {
        cudaError_t e = ensure(count ? count : 1);
        if (e != cudaSuccess || !count) return e;
        return cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct ncclComm;   // ctl_comm.cu (NCCL is loaded at run time, only when a communicator is asked for)

// Buffers of one wavefront in flight (SoA path state, rotating ray queues, hit records, shadow queue, per-bounce device counters).
struct WaveLane {
    DevBuf<float4> wo_prev, cf, cl, nor, px, rays_a, rays_b, rays_c, hit_a, sh_rays, sh_payload;
    DevBuf<uint32_t> path_a, path_b, path_c, hit_node, sort_keys, mat_order; DevBuf<unsigned> sort_hist, sort_offsets, mat_hist, counters; DevBuf<unsigned char> mat_cls;
    void release() {
        wo_prev.release(); cf.release(); cl.release(); nor.release(); px.release(); rays_a.release(); rays_b.release(); rays_c.release(); hit_a.release(); sh_rays.release(); sh_payload.release();
        path_a.release(); path_b.release(); path_c.release(); hit_node.release(); sort_keys.release(); mat_order.release(); sort_hist.release(); sort_offsets.release(); mat_hist.release(); counters.release(); mat_cls.release();
    }
};

// Queue of one WavefrontPathTracer pass in flight (DoubleRayBuffer<WavefrontPTRayData>)
struct WptLane {
    DevBuf<float4> w_thr, w_lxy, w_df, w_ray, w_sec[2]; DevBuf<uint2> w_misc; DevBuf<uint4> w_res, w_sres[2]; DevBuf<unsigned long long> w_desc;
    void release() { w_thr.release(); w_lxy.release(); w_df.release(); w_ray.release(); w_misc.release(); w_res.release(); w_desc.release(); for (int k = 0; k < 2; k++) { w_sec[k].release(); w_sres[k].release(); } }
};

// One frame in flight ("FramesInFlight", ctl_submit_frame_tiled / ctl_acquire_frame): its own PixelData accumulator, sample-table generator state and
// pinned table sets; rendered on wavefront lane slot + 1 (own stream, own wavefront buffers).
struct FrameSlot {
    DevBuf<float> accum; DevBuf<uint32_t> states; float* h1 = nullptr; float* h2 = nullptr; int h_cap = 0;
    cudaEvent_t begin = nullptr, done = nullptr, ready = nullptr, h_free = nullptr;   // begin / done: the lane's kernels (timed); ready: accumulator final (after the reduce, if any)
    uint32_t launches = 0, spp = 0; bool reduced = false;
};

struct ctl_ctx {
    ncclComm* comm = nullptr; int comm_rank = 0, comm_size = 1; unsigned long long* comm_scratch = nullptr;   // ctl_comm_init_* (ctl_comm.cu)
    int device = 0, w = 0, h = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // parameters (Integrators/PathTracer.h:10-20)
    int max_path_length = 50, rr_start = 5, direct = 1, regularization = 0, sort_mode = 0, stage_timers = 0, capture_bounce = 0, trav_kernel = 2, trav_blocks_per_sm = 8, shade_blocks_per_sm = 8, smem_carveout = -1, fuse_traversal = 1, warp_blocks = 0, pass_stride = 1, pass_phase = 0, stop_zero = 1;
    // scene
    DevBuf<ctl_bvh_node> d_scene_nodes, d_bvh_nodes; DevBuf<ctl_woop_tri> d_woop; DevBuf<uint32_t> d_tri_index; DevBuf<ctl_tri_data> d_tri_data;
    DevBuf<ctl_mesh> d_meshes; DevBuf<ctl_node> d_nodes; DevBuf<float> d_xf, d_inv_xf; DevBuf<ctl_material> d_materials; DevBuf<ctl_light> d_lights;
    DevBuf<ctl_light_tri> d_light_tris; DevBuf<float> d_light_cdf, d_normal_lut;
    DScene scene; bool has_scene = false;
    // sampler tables: `tab_cap` consecutive table sets (one per pass of a batch) on the device; generated there
    // (k_gen_tables) or -- "DeviceSampleTables"=0, the reference's UpdateKernel behaviour -- on the host and copied H2D
    int tab_cap = 0; DevBuf<float> d_tab1, d_tab2;
    float* h_tab1 = nullptr; float* h_tab2 = nullptr; int h_tab_cap = 0; cudaEvent_t h_tab_free = nullptr;
    DevBuf<uint32_t> d_states, d_states0, d_jump;
    bool user_tables = false; int device_tables = 1;
    uint32_t gen_pos_host = 0, gen_pos_dev = 0;   // index of the next pass each generator would produce
    ctlb::SamplerTableGenerator gen;
    // wavefront state: lane 0 runs on `stream`; lanes 1.. (own streams) hold the other wavefronts of a frame rendered with "OverlapWavefronts" (ctl_comm_render_frame)
    WaveLane lanes[MAX_LANES]; DevBuf<float4> capture;
    cudaStream_t lane_stream[MAX_LANES] = {}; cudaEvent_t ev_fork = nullptr, ev_join[MAX_LANES] = {}; cudaStream_t tab_stream = nullptr; cudaEvent_t ev_tab = nullptr;   // tab_stream: sample tables of a frame's wavefronts
    int shade_concurrent = 0; cudaStream_t cls_stream[MAX_LANES][3] = {}; cudaEvent_t ev_cls_fork[MAX_LANES] = {}, ev_cls_done[MAX_LANES][3] = {};   // "ShadeConcurrent": the per-class shade launches of a bounce on their own streams
    int defer = 0, defer_max_lag = 3; DevBuf<float4> df_sh_rays[MAX_LANES], df_sh_payload[MAX_LANES]; DevBuf<unsigned> df_cnt;   // "DeferStragglers": second shadow-queue buffer per lane, deferral counters
    int handover = 0, handover_drain = 16; DevBuf<uint32_t> ho_buf[2]; DevBuf<unsigned> ho_cnt;   // "HandOver": one-wavefront frames as two interleaved half-wavefronts whose traversal launches hand their unfinished rays over (device/traverse_handover.cuh)
    int overlap = 1, n_lanes = 4;   // "OverlapWavefronts", "OverlapLanes": see ctl_render_frame_tiled
    // frames in flight: frame k renders on lane 1 + k % fif into fslot[k % fif] while the frames before it drain; ctl_acquire_frame hands them back in order
    int fif = 2; FrameSlot fslot[MAX_LANES - 1]; unsigned long long f_submitted = 0, f_acquired = 0; float* lane_accum[MAX_LANES] = {}; cudaStream_t comm_stream = nullptr; int frame_root = 0; bool comm_defer = false; FrameSlot* comm_pending = nullptr;   // comm_defer: ctl_comm_submit_frame_all groups the reduces itself
    DevBuf<unsigned> api_work; unsigned api_seq = 0;   // ring of work counters of the API traversal launches: calls in flight on different streams never share one
    // WavefrontPathTracer queue (DoubleRayBuffer<WavefrontPTRayData>)
    WptLane wl[MAX_LANES];   // lane 0: ctl_wavefront_pass; lanes 1..: the other passes of a ctl_wavefront_frame in flight
    DevBuf<unsigned long long> stats; // [0] rays_last [1] rays_total [2..4] ext visits [5] ext rays [6..8] shadow visits [9] shadow rays
    DevBuf<float> own_accum; float* accum = nullptr; DevBuf<uchar4> resolve_tmp, pipe_rgbe; DevBuf<float4> pipe_partial; DevBuf<float> pipe_lum;
    DevBuf<ctl_pixel_variance_info> d_var; int variance_buffer = 0;
    DevBuf<uint32_t> d_node_alias; uint32_t n_alias = 0;   // re-braided scene: instance of every (pseudo-)node, for the node indices the API reports
    DevBuf<uchar4> nlm_cached; DevBuf<float> nlm_varh, nlm_weights; long long nlm_last_update = -1; size_t nlm_pixels = 0;   // NonLocalMeansFilter state (m_cachedImg, m_weightBuffer, last_iter_weight_update)
    unsigned captured_n = 0; DevBuf<unsigned> d_captured_n;
    uint32_t passes_done = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr; bool events_recorded = false;
    std::vector<cudaEvent_t> stage_ev; std::vector<int> stage_kind;
    float stage_ms[5] = {0, 0, 0, 0, 0}; uint32_t n_launches = 0;
    bool instrumented = false;
    TravTune tune = {2, 8, 8, 6, 2, 32, 0}, tune_p = {2, 8, 8, 4, 1, 0, 0};   // scheduler parameters of the staged kernel (swept on the device: profiles/r02c_tune_sweep.log) / of the persistent kernel (profiles/r01d_*)
    // staged traversal kernel (device/traverse_staged.cuh): derived records + launch shape
    DevBuf<float4> d_tri64, d_inst, d_treelet; StagedScene staged = {nullptr, nullptr, nullptr, 0, 0, 16, 0}; bool staged_ok = false; std::string staged_why;
    int shade_mode = 1; uint32_t class_mask = 0; bool class_ok = false;   // "ShadeMode": 0 = one k_shade with the run-time BSDF dispatch, 1 = one launch per material class present (staged kernel only)
    int staged_threads = 128, staged_rows = 16, staged_treelet = 0, staged_resident = 1024;   // "StagedThreads", "StagedStackRows", "StagedTreeletNodes", "StagedResidentThreads"
};

inline int grid_for(const ctl_ctx* c, int per_sm) { return c->n_sm * per_sm; }
int ctl_variance_after_pass(ctl_ctx* c, bool new_trace);   // ctl_pipeline.cu: PixelVarianceBuffer::AddPass after a whole-image pass
void ctl_comm_release(ctl_ctx* c);                          // ctl_comm.cu: called by ctl_destroy
int ctl_comm_reduce_slot(ctl_ctx* c, FrameSlot& F);         // ctl_comm.cu: the reduce of a frame in flight on the communication stream (no-op without a communicator)
