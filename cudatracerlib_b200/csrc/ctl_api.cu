// ctl_api.cu -- C ABI (include/ctl_b200.h) over the sm_100a kernels.
//
// Host orchestration replaces Tracer<true>::DoPass / UpdateKernel / __internal__IntersectBuffers
// (Kernel/Tracer.h:209-289, Kernel/TraceHelper.cu:182-217, 736-746): one context per device, one
// stream, no globals, no per-launch cudaDeviceSynchronize, queue sizes stay on the device.
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
#include "wavefront.cuh"
#include "device/traverse_staged.cuh"
#include "wavefront_pt.cuh"
#include "image_pipeline.cuh"
#include "nlm_filter.cuh"
#include "bvh_build.cuh"

using namespace ctld;

static thread_local std::string g_err;
static int set_err(const std::string& s) { g_err = s; return 1; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return set_err(b_); } } while (0)
#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); set_err(b_); return nullptr; } } while (0)

struct ctl_scene { ctlb::SceneStorage S; };

namespace {
const int MAX_BOUNCES = 256;
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
} // namespace

struct ctl_ctx {
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
    // wavefront state
    DevBuf<float4> wo_prev; DevBuf<float4> cf, cl, nor, px, rays_a, rays_b, hit_a, sh_rays, sh_payload, capture;
    DevBuf<uint32_t> path_a, path_b, path_c, hit_node, sort_keys; DevBuf<float4> rays_c; DevBuf<unsigned> sort_hist, sort_offsets, mat_hist; DevBuf<unsigned char> mat_cls; DevBuf<uint32_t> mat_order;
    DevBuf<unsigned> counters;
    DevBuf<unsigned> api_work; unsigned api_seq = 0;   // ring of work counters of the API traversal launches: calls in flight on different streams never share one
    // WavefrontPathTracer queue (DoubleRayBuffer<WavefrontPTRayData>)
    DevBuf<float4> w_thr, w_lxy, w_df, w_ray, w_sec[2]; DevBuf<uint2> w_misc; DevBuf<uint4> w_res, w_sres[2]; DevBuf<unsigned long long> w_desc;
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
    TravTune tune = {2, 8, 8, 6, 2}, tune_p = {2, 8, 8, 4, 1};   // scheduler parameters of the staged kernel (swept on the device: profiles/r02c_tune_sweep.log) / of the persistent kernel (profiles/r01d_*)
    // staged traversal kernel (device/traverse_staged.cuh): derived records + launch shape
    DevBuf<float4> d_tri64, d_inst, d_treelet; StagedScene staged = {nullptr, nullptr, nullptr, 0, 0, 16}; bool staged_ok = false; std::string staged_why;
    int shade_mode = 1; uint32_t class_mask = 0; bool class_ok = false;   // "ShadeMode": 0 = one k_shade with the run-time BSDF dispatch, 1 = one launch per material class present (staged kernel only)
    int staged_threads = 512, staged_rows = 16, staged_treelet = 0, staged_resident = 1024;   // "StagedThreads", "StagedStackRows", "StagedTreeletNodes", "StagedResidentThreads"
};


static void launch_shade(int cls, int grid, cudaStream_t st, const DScene& S, const ShadeParams& P, const PathState& ps, const Queues& Q, const unsigned* n_in, unsigned* n_out, unsigned* n_shadow, const unsigned* seg_hist) {
    switch (cls) {
    case 0: k_shade<0><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    case 1: k_shade<1><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    case 2: k_shade<2><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    case 3: k_shade<3><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, seg_hist); break;
    default: k_shade<-1><<<grid, 128, 0, st>>>(S, P, ps, Q, n_in, n_out, n_shadow, nullptr); break;
    }
}

// Launch shape of the staged kernel: blocks of `staged_threads`, as many per SM as 1024 resident threads and the shared memory allow
static size_t staged_smem_bytes(const ctl_ctx* c) { return (size_t)c->staged.tl_nodes * 64 + 16 + (size_t)(c->staged.stack_rows + 1) * c->staged_threads * 4; }
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
    static size_t attr_set = 0; // per instantiation
    const size_t smem = staged_smem_bytes(c);
    if (smem > attr_set) { cudaFuncSetAttribute(k_intersect_staged<MODE, ANY_HIT, COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = smem; }
    k_intersect_staged<MODE, ANY_HIT, COUNT><<<staged_grid(c), c->staged_threads, smem, st>>>(c->scene, c->staged, c->tune, rays, n_ptr, n2_ptr, n_fixed, work, out, visit);
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

const char* ctl_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ scenes (host)
ctl_scene* ctl_scene_create(int kind, int width, int height, uint32_t seed, int n_hint) {
    try { std::unique_ptr<ctl_scene> s(new ctl_scene()); ctlb::make_scene(kind, width, height, seed, n_hint, s->S); return s.release(); }
    catch (const std::exception& e) { set_err(e.what()); return nullptr; }
}
ctl_scene* ctl_scene_create_from_mesh(const float* verts, uint32_t nv, const uint32_t* indices, uint32_t nt, const uint8_t* mat_index,
                                      const ctl_material* materials, uint32_t nm, const float* emissive, const float* cam_pos,
                                      const float* cam_target, const float* cam_up, float fov_deg, int width, int height) {
    if (!verts || !indices || !mat_index || !materials || !cam_pos || !cam_target || !cam_up) { set_err("null argument"); return nullptr; }
    if (!nv || !nt || !nm) { set_err("empty mesh (no vertices, triangles or materials)"); return nullptr; }
    if (width <= 0 || height <= 0) { set_err("bad image size"); return nullptr; }
    for (size_t i = 0; i < 3 * (size_t)nt; i++) if (indices[i] >= nv) { set_err("triangle " + std::to_string(i / 3) + ": vertex index " + std::to_string(indices[i]) + " out of range (" + std::to_string(nv) + " vertices)"); return nullptr; }
    for (uint32_t i = 0; i < nt; i++) if (mat_index[i] >= nm) { set_err("triangle " + std::to_string(i) + ": material index " + std::to_string((unsigned)mat_index[i]) + " out of range (" + std::to_string(nm) + " materials)"); return nullptr; }
    try {
        ctlb::MeshInput M;
        for (uint32_t i = 0; i < nv; i++) M.verts.push_back(ctlb::V3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
        M.indices.assign(indices, indices + 3 * (size_t)nt);
        M.mat_index.assign(mat_index, mat_index + nt);
        M.materials.assign(materials, materials + nm);
        for (uint32_t i = 0; i < nm; i++) M.emissive.push_back(emissive ? ctlb::V3(emissive[3 * i], emissive[3 * i + 1], emissive[3 * i + 2]) : ctlb::V3(0.0f));
        std::unique_ptr<ctl_scene> s(new ctl_scene());
        std::vector<ctlb::MeshInput> meshes = {M};
        std::vector<ctlb::NodeInput> nodes = {{0, ctlb::M4::identity(), -1}};
        ctlb::assemble_scene(meshes, nodes, ctlb::V3(cam_pos[0], cam_pos[1], cam_pos[2]), ctlb::V3(cam_target[0], cam_target[1], cam_target[2]),
                             ctlb::V3(cam_up[0], cam_up[1], cam_up[2]), fov_deg, width, height, s->S);
        return s.release();
    } catch (const std::exception& e) { set_err(e.what()); return nullptr; }
}
// == DynamicScene::CreateNode(compiled mesh file) per file + the camera (Engine/DynamicScene.cpp:283-345; reader Engine/Mesh.cpp:46-98): SURVEY 8 f4
ctl_scene* ctl_scene_create_from_xmsh(const char* const* paths, uint32_t n_files, const float* node_xforms, const float* cam_pos, const float* cam_target,
                                      const float* cam_up, float fov_deg, int width, int height) {
    if (!paths || !n_files || !cam_pos || !cam_target || !cam_up) { set_err("null / empty argument"); return nullptr; }
    try {
        std::vector<ctlb::MeshInput> meshes(n_files); std::vector<ctlb::NodeInput> nodes;
        for (uint32_t i = 0; i < n_files; i++) {
            if (!paths[i]) throw std::runtime_error("null path");
            const std::string pth(paths[i]);
            if (pth.size() > 4 && (pth.substr(pth.size() - 4) == ".obj" || pth.substr(pth.size() - 4) == ".OBJ")) ctlb::read_obj(paths[i], meshes[i]); // MeshCompilerManager picks the compiler by extension (MeshCompiler.cpp:21-27)
            else if (pth.size() > 4 && (pth.substr(pth.size() - 4) == ".ply" || pth.substr(pth.size() - 4) == ".PLY")) ctlb::read_ply(paths[i], meshes[i]);
            else ctlb::read_xmsh(paths[i], meshes[i]);
            ctlb::M4 xf = ctlb::M4::identity();
            if (node_xforms) memcpy(xf.m, node_xforms + 16 * (size_t)i, 64);
            nodes.push_back({i, xf, -1});
        }
        ctl_scene* s = new ctl_scene();
        try {
            ctlb::assemble_scene(meshes, nodes, ctlb::V3(cam_pos[0], cam_pos[1], cam_pos[2]), ctlb::V3(cam_target[0], cam_target[1], cam_target[2]),
                                 ctlb::V3(cam_up[0], cam_up[1], cam_up[2]), fov_deg, width, height, s->S);
        } catch (...) { delete s; throw; }
        return s;
    } catch (const std::exception& e) { set_err(e.what()); return nullptr; }
}
ctl_scene* ctl_scene_create_from_files(const char* const* paths, uint32_t n_files, const float* node_xforms, const float* cam_pos, const float* cam_target,
                                       const float* cam_up, float fov_deg, int width, int height) {
    return ctl_scene_create_from_xmsh(paths, n_files, node_xforms, cam_pos, cam_target, cam_up, fov_deg, width, height);
}
// == the output sequence of Mesh::CompileMesh (Engine/Mesh.cpp:278-289) for mesh `mesh` of a host scene
int ctl_scene_write_xmsh(const ctl_scene* s, uint32_t mesh, const char* path) {
    if (!s || !path) return set_err("null argument");
    try { ctlb::write_xmsh(path, s->S, mesh); return 0; }
    catch (const std::exception& e) { return set_err(e.what()); }
}
// Source triangles of mesh `mesh` (9 floats each, TriangleData order) for export / rebuild tooling; *n_tris receives the count (verts9_out may be NULL to size).
int ctl_scene_get_mesh_triangles(const ctl_scene* s, uint32_t mesh, float* verts9_out, uint32_t* n_tris) {
    if (!s || !n_tris) return set_err("null argument");
    if (mesh >= s->S.mesh_verts9.size()) return set_err("no such mesh");
    const std::vector<float>& v = s->S.mesh_verts9[mesh];
    if (v.empty()) return set_err("mesh has no source triangles (imported from a compiled file)");
    *n_tris = (uint32_t)(v.size() / 9);
    if (verts9_out) memcpy(verts9_out, v.data(), v.size() * sizeof(float));
    return 0;
}
// == DynamicScene::SetNodeTransform (Engine/DynamicScene.cpp:433-443): new local-to-world matrix of one instance; the node level is re-assembled (scene-level
// BVH = BVHRebuilder's job, inverse matrix, the node's area lights -> RecomputeShape, scene box, ray epsilon).  Mesh BVHs, Woop triangles and TriangleData are
// untouched.  Views obtained before are invalidated: call ctl_scene_get_view and ctl_upload_scene (or ctl_update_scene_nodes) again.
int ctl_scene_set_node_transform(ctl_scene* s, uint32_t node, const float* xf16) {
    if (!s || !xf16) return set_err("null argument");
    if (node >= s->S.node_inputs.size()) return set_err("no such node");
    try { memcpy(s->S.node_inputs[node].xf.m, xf16, 64); ctlb::assemble_nodes(s->S); return 0; }
    catch (const std::exception& e) { return set_err(e.what()); }
}
int ctl_scene_set_rebraid(ctl_scene* s, uint32_t max_entries) {
    if (!s) return set_err("null argument");
    try { s->S.rebraid_entries = max_entries; ctlb::assemble_nodes(s->S); return 0; }
    catch (const std::exception& e) { return set_err(e.what()); }
}
int ctl_scene_get_view(const ctl_scene* s, ctl_scene_view* out) { if (!s || !out) return set_err("null argument"); s->S.fill_view(out); return 0; }
void ctl_scene_destroy(ctl_scene* s) { delete s; }
int ctl_validate_scene_view(const ctl_scene_view* v) {
    if (!v) return set_err("null argument");
    try { ctlb::validate_view(*v); return 0; } catch (const std::exception& e) { return set_err(e.what()); }
}
void ctl_encode_woop(const float v0[3], const float v1[3], const float v2[3], ctl_woop_tri* out) {
    ctlb::encode_woop(ctlb::V3(v0[0], v0[1], v0[2]), ctlb::V3(v1[0], v1[1], v1[2]), ctlb::V3(v2[0], v2[1], v2[2]), out);
}
void ctl_encode_tri_data(const float p[9], const float n[9], const float uv[6], uint32_t mat, ctl_tri_data* out) {
    ctlb::V3 P[3], N[3];
    for (int i = 0; i < 3; i++) { P[i] = ctlb::V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); N[i] = ctlb::V3(n[3 * i], n[3 * i + 1], n[3 * i + 2]); }
    ctlb::encode_tri_data(P, N, uv, mat, out);
}
int ctl_generate_sample_tables(uint32_t pass, float* d1, float* d2) {
    ctlb::SamplerTableGenerator g;
    for (uint32_t p = 0; p <= pass; p++) g.next_pass(d1, d2);
    return 0;
}


// ------------------------------------------------------------------ GPU BVH build
// LBVH of one triangle mesh in the reference layout; host arrays in, host arrays out (nodes_out: capacity >= max(1, n_tris) entries).
int ctl_bvh_build_gpu(int device, const float* verts9, uint32_t n_tris, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms) {
    using namespace ctlbvh;
    if (!verts9 || !n_tris || !nodes_out || !n_nodes_out || !woop_out || !index_out) return set_err("null / empty argument");
    if (n_tris > 0x3fffffffu) return set_err("too many triangles");
    CK(cudaSetDevice(device));
    const int n = (int)n_tris;
    const int nb_sort = (n + SORT_TILE - 1) / SORT_TILE;
    DevBuf<float> d_verts; DevBuf<float4> d_boxes, d_nbox; DevBuf<unsigned> d_sbox, d_counts, d_flags, d_emit; DevBuf<uint32_t> d_k0, d_k1, d_v0, d_v1, d_index;
    DevBuf<int> d_left, d_right, d_pint, d_pleaf, d_first, d_last; DevBuf<ctl_bvh_node> d_nodes; DevBuf<ctl_woop_tri> d_woop; DevBuf<unsigned char> d_lastflag, d_collapse; DevBuf<float> d_cost;
    auto free_all = [&]() { d_verts.release(); d_boxes.release(); d_nbox.release(); d_sbox.release(); d_counts.release(); d_flags.release(); d_emit.release(); d_k0.release(); d_k1.release(); d_v0.release(); d_v1.release();
                            d_index.release(); d_left.release(); d_right.release(); d_pint.release(); d_pleaf.release(); d_first.release(); d_last.release(); d_nodes.release(); d_woop.release(); d_lastflag.release(); d_collapse.release(); d_cost.release(); };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { free_all(); char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return set_err(b_); } } while (0)
    CKF(d_verts.upload(verts9, (size_t)n * 9)); CKF(d_boxes.ensure((size_t)n * 2)); CKF(d_nbox.ensure((size_t)n * 2)); CKF(d_sbox.ensure(6)); CKF(d_counts.ensure((size_t)256 * nb_sort));
    CKF(d_flags.ensure((size_t)n)); CKF(d_emit.ensure((size_t)n + 1)); CKF(d_k0.ensure(n)); CKF(d_k1.ensure(n)); CKF(d_v0.ensure(n)); CKF(d_v1.ensure(n)); CKF(d_index.ensure(n));
    CKF(d_left.ensure(n)); CKF(d_right.ensure(n)); CKF(d_pint.ensure(n)); CKF(d_pleaf.ensure(n)); CKF(d_first.ensure(n)); CKF(d_last.ensure(n)); CKF(d_nodes.ensure((size_t)n)); CKF(d_woop.ensure(n)); CKF(d_lastflag.ensure(n)); CKF(d_collapse.ensure(n)); CKF(d_cost.ensure(n));
    cudaEvent_t e0, e1; CKF(cudaEventCreate(&e0)); CKF(cudaEventCreate(&e1));
    cudaStream_t st = nullptr;
    CKF(cudaEventRecord(e0, st));
    const unsigned sbox_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CKF(cudaMemcpyAsync(d_sbox.p, sbox_init, sizeof(sbox_init), cudaMemcpyHostToDevice, st));
    CKF(cudaMemsetAsync(d_flags.p, 0, (size_t)n * 4, st)); CKF(cudaMemsetAsync(d_lastflag.p, 0, (size_t)n, st));
    const int g = (n + 255) / 256;
    k_tri_boxes<<<g, 256, 0, st>>>(d_verts.p, n_tris, d_boxes.p, d_sbox.p);
    k_morton<<<g, 256, 0, st>>>(d_boxes.p, n_tris, d_sbox.p, d_k0.p, d_v0.p);
    uint32_t *kin = d_k0.p, *kout = d_k1.p, *vin = d_v0.p, *vout = d_v1.p;
    for (int pass = 0; pass < 4; pass++) {
        k_sort_hist<<<nb_sort, SORT_THREADS, 0, st>>>(kin, n_tris, 8 * pass, d_counts.p, nb_sort);
        k_scan_exclusive<<<1, 1024, 0, st>>>(d_counts.p, (uint32_t)(256 * nb_sort));
        k_sort_scatter<<<nb_sort, SORT_THREADS, 0, st>>>(kin, vin, n_tris, 8 * pass, d_counts.p, nb_sort, kout, vout);
        std::swap(kin, kout); std::swap(vin, vout);
    }
    uint32_t n_nodes = 1;
    if (n > MAX_LEAF) {
        k_radix_tree<<<g, 256, 0, st>>>(kin, n, d_left.p, d_right.p, d_pint.p, d_pleaf.p, d_first.p, d_last.p);
        k_fit_boxes<<<g, 256, 0, st>>>(d_boxes.p, vin, n, d_left.p, d_right.p, d_pint.p, d_pleaf.p, d_first.p, d_last.p, d_flags.p, d_nbox.p, d_cost.p, d_collapse.p);
        k_mark_emitted<<<g, 256, 0, st>>>(n, d_pint.p, d_first.p, d_last.p, d_collapse.p, d_emit.p);
        k_scan_exclusive<<<1, 1024, 0, st>>>(d_emit.p, (uint32_t)n);
        k_emit_nodes<<<g, 256, 0, st>>>(n, d_left.p, d_right.p, d_pint.p, d_first.p, d_last.p, d_emit.p, d_boxes.p, vin, d_nbox.p, d_collapse.p, d_nodes.p, d_lastflag.p);
        CKF(cudaMemcpyAsync(&n_nodes, d_emit.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    } else {
        k_single_leaf_root<<<1, 32, 0, st>>>(d_sbox.p, n, d_nodes.p, d_lastflag.p);
    }
    k_emit_tris<<<g, 256, 0, st>>>(d_verts.p, vin, n, d_lastflag.p, d_woop.p, d_index.p);
    CKF(cudaGetLastError());
    CKF(cudaEventRecord(e1, st));
    CKF(cudaStreamSynchronize(st));
    float ms = 0; CKF(cudaEventElapsedTime(&ms, e0, e1));
    CKF(cudaMemcpy(nodes_out, d_nodes.p, (size_t)n_nodes * sizeof(ctl_bvh_node), cudaMemcpyDeviceToHost));
    CKF(cudaMemcpy(woop_out, d_woop.p, (size_t)n * sizeof(ctl_woop_tri), cudaMemcpyDeviceToHost));
    CKF(cudaMemcpy(index_out, d_index.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    free_all();
#undef CKF
    *n_nodes_out = n_nodes;
    if (build_ms) *build_ms = ms;
    return 0;
}

// Rebuild every mesh BVH of a host scene on the GPU (node / Woop / index arrays, mesh offsets, light-triangle slots).
int ctl_scene_rebuild_bvh_gpu(ctl_scene* s, int device, float* build_ms_total) {
    if (!s) return set_err("null scene");
    ctlb::SceneStorage& S = s->S;
    if (S.mesh_verts9.size() != S.meshes.size()) return set_err("scene has no triangle vertices (built by an older builder)");
    std::vector<ctl_bvh_node> all_nodes; std::vector<ctl_woop_tri> all_woop; std::vector<uint32_t> all_index;
    std::vector<ctl_mesh> meshes = S.meshes;
    float total = 0;
    std::vector<std::vector<uint32_t>> slot_of_tri(S.meshes.size());
    for (size_t mi = 0; mi < S.meshes.size(); mi++) {
        const uint32_t nt = (uint32_t)(S.mesh_verts9[mi].size() / 9);
        std::vector<ctl_bvh_node> nodes(nt ? nt : 1); std::vector<ctl_woop_tri> woop(nt); std::vector<uint32_t> index(nt);
        uint32_t nn = 0; float ms = 0;
        if (ctl_bvh_build_gpu(device, S.mesh_verts9[mi].data(), nt, nodes.data(), &nn, woop.data(), index.data(), &ms)) return 1;
        total += ms;
        if (getenv("CTL_LBVH_OPTIMIZE")) { nodes.resize(nn); ctlb::optimize_bvh(nodes); }   // experiment for round 2: the mesh trees' host post-pass (re-insertion + rotations) on the LBVH; node count unchanged
        meshes[mi].bvh_node_offset = (uint32_t)all_nodes.size() * 4;
        meshes[mi].bvh_tri_offset = (uint32_t)all_woop.size() * 3;
        meshes[mi].bvh_idx_offset = (uint32_t)all_index.size();
        slot_of_tri[mi].assign(nt, 0);
        for (uint32_t k = 0; k < nt; k++) slot_of_tri[mi][index[k] >> 1] = meshes[mi].bvh_idx_offset + k;
        all_nodes.insert(all_nodes.end(), nodes.begin(), nodes.begin() + nn);
        all_woop.insert(all_woop.end(), woop.begin(), woop.end());
        all_index.insert(all_index.end(), index.begin(), index.end());
    }
    // light triangles point at Woop slots (ShapeSet::triData::iDat): remap through (mesh, triangle)
    for (auto& lt : S.light_tris) {
        for (size_t mi = 0; mi < S.meshes.size(); mi++) {
            const uint32_t t0 = S.meshes[mi].tri_offset, nt = (uint32_t)slot_of_tri[mi].size();
            if (lt.t_dat >= t0 && lt.t_dat < t0 + nt) { lt.i_dat = slot_of_tri[mi][lt.t_dat - t0]; break; }
        }
    }
    S.bvh_nodes.swap(all_nodes); S.woop.swap(all_woop); S.tri_index.swap(all_index); S.meshes = meshes;
    if (S.rb_active) { try { ctlb::assemble_nodes(S); } catch (const std::exception& e) { return set_err(e.what()); } }   // re-braided entries are copies of the old sub-trees: redo them
    if (build_ms_total) *build_ms_total = total;
    return 0;
}

// ------------------------------------------------------------------ context
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
    CK(c->counters.ensure(CTR_TOTAL)); CK(c->api_work.ensure(API_WORK_RING)); CK(c->stats.ensure(16)); CK(c->d_captured_n.ensure(1));
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
    if (init_ctx(c)) { const std::string keep = g_err; ctl_destroy(c); g_err = keep; return nullptr; }
    return c;
}

void ctl_destroy(ctl_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->d_scene_nodes.release(); c->d_bvh_nodes.release(); c->d_woop.release(); c->d_tri_index.release(); c->d_tri_data.release(); c->d_meshes.release();
    c->d_nodes.release(); c->d_xf.release(); c->d_inv_xf.release(); c->d_materials.release(); c->d_lights.release(); c->d_light_tris.release();
    c->d_light_cdf.release(); c->d_normal_lut.release(); c->d_tri64.release(); c->d_inst.release(); c->d_treelet.release();
    c->d_tab1.release(); c->d_tab2.release(); c->d_states.release(); c->d_states0.release(); c->d_jump.release();
    if (c->h_tab1) cudaFreeHost(c->h_tab1); if (c->h_tab2) cudaFreeHost(c->h_tab2); if (c->h_tab_free) cudaEventDestroy(c->h_tab_free);
    c->wo_prev.release(); c->cf.release(); c->cl.release(); c->nor.release(); c->px.release(); c->rays_a.release(); c->rays_b.release(); c->hit_a.release(); c->sh_rays.release();
    c->sh_payload.release(); c->capture.release(); c->path_a.release(); c->path_b.release(); c->path_c.release(); c->rays_c.release(); c->sort_keys.release(); c->sort_hist.release(); c->sort_offsets.release(); c->mat_hist.release(); c->mat_cls.release(); c->mat_order.release(); c->hit_node.release(); c->counters.release(); c->api_work.release(); c->stats.release();
    c->own_accum.release(); c->d_captured_n.release(); c->resolve_tmp.release(); c->pipe_rgbe.release(); c->pipe_partial.release(); c->pipe_lum.release(); c->d_var.release(); c->nlm_cached.release(); c->nlm_varh.release(); c->nlm_weights.release(); c->nlm_last_update = -1; c->nlm_pixels = 0; c->d_node_alias.release();
    c->w_thr.release(); c->w_lxy.release(); c->w_df.release(); c->w_ray.release(); c->w_misc.release(); c->w_res.release(); c->w_desc.release();
    for (int k = 0; k < 2; k++) { c->w_sec[k].release(); c->w_sres[k].release(); }
    for (auto e : c->stage_ev) cudaEventDestroy(e);
    if (c->ev_start) cudaEventDestroy(c->ev_start); if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int ctl_resize(ctl_ctx* c, int width, int height) {
    if (!c || width <= 0 || height <= 0) return set_err("invalid argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
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
    else if (k == "Regularization") { if (v != 0) return set_err("Regularization=true is not implemented (off by default in the reference)"); c->regularization = 0; }
    else if (k == "SortMode") c->sort_mode = v;
    else if (k == "ShadeMode") { if (v < 0 || v > 1) return set_err("ShadeMode must be 0 (run-time BSDF dispatch) or 1 (one launch per material class)"); c->shade_mode = v; }
    else if (k == "StageTimers") c->stage_timers = v != 0;
    else if (k == "CaptureBounce") c->capture_bounce = v;
    else if (k == "DeviceSampleTables") c->device_tables = v != 0;
    else if (k == "FuseTraversal") c->fuse_traversal = v != 0;
    else if (k == "PixelVarianceBuffer") c->variance_buffer = v != 0;
    else if (k == "WarpPixelBlocks") c->warp_blocks = v != 0;
    else if (k == "PassStride") { if (v < 1) return set_err("PassStride must be >= 1"); c->pass_stride = v; }   // multi-GPU by pass: this context renders passes PassPhase + k * PassStride
    else if (k == "PassPhase") { if (v < 0) return set_err("PassPhase must be >= 0"); c->pass_phase = v; }
    else if (k == "TraversalKernel") { if (v < 0 || v > 2) return set_err("TraversalKernel must be 0 (persistent), 1 (ray batch) or 2 (staged)"); c->trav_kernel = v; }
    else if (k == "StagedThreads") { if (v < 32 || v > 1024 || (v & 31)) return set_err("StagedThreads must be a multiple of 32 in [32,1024]"); c->staged_threads = v; }
    else if (k == "StagedResidentThreads") { if (v < 32 || v > 2048) return set_err("StagedResidentThreads out of range [32,2048]"); c->staged_resident = v; }
    else if (k == "StagedStackRows") { if (v < 0 || v > TP_STACK) return set_err("StagedStackRows out of range [0,64]"); c->staged_rows = v; c->staged.stack_rows = v; }
    else if (k == "StagedTreeletNodes") { if (v < 0 || v > 2048) return set_err("StagedTreeletNodes out of range [0,2048]"); c->staged_treelet = v; }   // takes effect at the next ctl_upload_scene / ctl_update_scene_nodes
    else if (k == "TravThT") c->tune.th_t = c->tune_p.th_t = v; else if (k == "TravThL") c->tune.th_l = c->tune_p.th_l = v; else if (k == "TravThF") c->tune.th_f = c->tune_p.th_f = v;
    else if (k == "TravThNExit") c->tune.th_n_exit = c->tune_p.th_n_exit = v;
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
    else if (k == "CaptureBounce") *v = c->capture_bounce; else if (k == "TraversalKernel") *v = c->trav_kernel; else if (k == "DeviceSampleTables") *v = c->device_tables; else if (k == "FuseTraversal") *v = c->fuse_traversal; else if (k == "PixelVarianceBuffer") *v = c->variance_buffer; else if (k == "PassStride") *v = c->pass_stride; else if (k == "PassPhase") *v = c->pass_phase;
    else if (k == "TraversalBlocksPerSM") *v = c->trav_blocks_per_sm; else if (k == "StagedThreads") *v = c->staged_threads; else if (k == "StagedStackRows") *v = c->staged_rows;
    else if (k == "ShadeMode") *v = c->shade_mode; else if (k == "MaterialClassMask") *v = (int)c->class_mask;
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

// Tables of passes [first, first + n) into the device table sets 0..n-1 (GenerateNewRandomSequences, Kernel/Sampler.h:36-55)
static int generate_tables(ctl_ctx* c, uint32_t first, int n) {
    if (ensure_tables(c, n)) return 1;
    if (c->device_tables) {
        if (c->gen_pos_dev != first) { // re-synchronise the device stream position (mode switch / user tables / strided passes): skip forward, or restart
            uint32_t from = c->gen_pos_dev;
            if (from > first) { CK(cudaMemcpyAsync(c->d_states.p, c->d_states0.p, (size_t)ctlb::kNumSeq * 6 * 4, cudaMemcpyDeviceToDevice, c->stream)); from = 0; }
            for (uint32_t p = from; p < first; p++) k_gen_tables<<<ctlb::kNumSeq / 128, 128, 0, c->stream>>>(c->d_states.p, c->d_jump.p, 1, c->d_tab1.p, (float2*)c->d_tab2.p);
        }
        k_gen_tables<<<ctlb::kNumSeq / 128, 128, 0, c->stream>>>(c->d_states.p, c->d_jump.p, n, c->d_tab1.p, (float2*)c->d_tab2.p);
        CK(cudaGetLastError());
        c->gen_pos_dev = first + n;
    } else {
        if (ensure_host_tables(c, n)) return 1;
        CK(cudaEventSynchronize(c->h_tab_free));
        if (c->gen_pos_host != first) { uint32_t from = c->gen_pos_host; if (from > first) { c->gen.reset(); from = 0; } for (uint32_t p = from; p < first; p++) c->gen.next_pass(c->h_tab1, c->h_tab2); }
        for (int p = 0; p < n; p++) c->gen.next_pass(c->h_tab1 + TAB1 * p, c->h_tab2 + TAB2 * p);
        CK(cudaMemcpyAsync(c->d_tab1.p, c->h_tab1, TAB1 * 4 * n, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_tab2.p, c->h_tab2, TAB2 * 4 * n, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->h_tab_free, c->stream));
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
static int grid_for(const ctl_ctx* c, int per_sm) { return c->n_sm * per_sm; }


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
static int ensure_state(ctl_ctx* c, size_t n) {
    CK(c->cf.ensure(n)); CK(c->cl.ensure(n)); CK(c->nor.ensure(n)); CK(c->px.ensure(n));
    CK(c->rays_a.ensure(2 * n)); CK(c->rays_b.ensure(2 * n)); CK(c->hit_a.ensure(n)); CK(c->hit_node.ensure(n));
    CK(c->sh_rays.ensure(2 * n)); CK(c->sh_payload.ensure(n)); CK(c->path_a.ensure(n)); CK(c->path_b.ensure(n));
    if (!c->stop_zero) CK(c->wo_prev.ensure(n));
    if (c->sort_mode == 2) CK(c->mat_cls.ensure(n));
    if (c->sort_mode == 2 || c->shade_mode == 1) { CK(c->mat_order.ensure(n)); CK(c->mat_hist.ensure(2 * MAT_CLASSES * (MAX_BOUNCES + 1))); }
    if (c->sort_mode == 1) {
        CK(c->rays_c.ensure(2 * n)); CK(c->path_c.ensure(n)); CK(c->sort_keys.ensure(n));
        if (!c->sort_hist.p) { CK(c->sort_hist.ensure(SORT_BUCKETS)); CK(c->sort_offsets.ensure(SORT_BUCKETS)); CK(cudaMemsetAsync(c->sort_hist.p, 0, SORT_BUCKETS * sizeof(unsigned), c->stream)); }
    }
    return 0;
}

static void stage_mark(ctl_ctx* c, int kind) {
    if (!c->stage_timers) return;
    size_t i = c->stage_kind.size();
    if (i >= c->stage_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); c->stage_ev.push_back(e); }
    cudaEventRecord(c->stage_ev[i], c->stream);
    c->stage_kind.push_back(kind);
}

static int variance_after_pass(ctl_ctx* c, bool new_trace);

static int render_window(ctl_ctx* c, int new_trace, const Window& W) {
    if (!c->has_scene) return set_err("no scene uploaded");
    if (c->variance_buffer && (W.n_passes != 1 || W.mode != 0 || W.n_slots != c->w * c->h))
        return set_err("PixelVarianceBuffer=1 needs whole-image single-pass renders (ctl_render_pass with the full window, ctl_wavefront_pass)");
    if (W.n_slots <= 0) return 0;
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev_start, c->stream));
    if (new_trace) {
        CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream));
        c->passes_done = 0;
    }
    // sample tables of these passes: caller-supplied (single pass), generated on the device, or host XORWOW + H2D
    if (c->user_tables) { if (W.n_passes != 1) return set_err("caller-supplied sample tables cover exactly one pass"); }
    else if (generate_tables(c, c->passes_done, W.n_passes)) return 1;
    c->user_tables = false;
    c->scene.d1 = c->d_tab1.p; c->scene.d2 = (const float2*)c->d_tab2.p;
    c->scene.img_w = c->w; c->scene.img_h = c->h;
    const size_t n_paths = (size_t)W.n_slots * W.n_passes;
    if (ensure_state(c, n_paths)) return 1;
    if (c->capture_bounce > 0) CK(c->capture.ensure(2 * n_paths));

    c->stage_kind.clear();
    CK(cudaMemsetAsync(c->counters.p, 0, CTR_TOTAL * sizeof(unsigned), c->stream));
    // per-class shade launches: the staged kernel tags every hit with its material class; one launch per class present (a sort pass only when there are several)
    const bool by_class = c->shade_mode == 1 && c->sort_mode != 2 && c->trav_kernel == 2 && c->staged_ok && c->class_ok && c->class_mask != 0;
    int n_classes = 0, single_cls = -1;
    for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { n_classes++; single_cls = k; }
    const bool class_sort = by_class && n_classes > 1;
    if (c->sort_mode == 2 || class_sort) CK(cudaMemsetAsync(c->mat_hist.p, 0, 2 * MAT_CLASSES * (MAX_BOUNCES + 1) * sizeof(unsigned), c->stream));
    if (c->instrumented) CK(cudaMemsetAsync(c->stats.p + 2, 0, 8 * sizeof(unsigned long long), c->stream));
    unsigned* ctr = c->counters.p;
    PathState st = {c->cf.p, c->cl.p, c->nor.p, c->px.p, c->stop_zero ? nullptr : c->wo_prev.p};
    const int g_light = grid_for(c, c->shade_blocks_per_sm);
    const int g_trav = grid_for(c, c->trav_blocks_per_sm);
    uint32_t launches = 0;
    stage_mark(c, 0);
    k_generate<<<g_light, 256, 0, c->stream>>>(c->scene, W, st, c->rays_a.p, c->path_a.p, ctr + CTR_Q + 0);
    launches++;
    ShadeParams P = {c->max_path_length, c->rr_start, c->direct, c->stop_zero};
    float4* rin = c->rays_a.p; float4* rout = c->rays_b.p; uint32_t* pin = c->path_a.p; uint32_t* pout = c->path_b.p;
    float4* rspare = c->rays_c.p; uint32_t* pspare = c->path_c.p;
    const bool fuse = c->fuse_traversal && c->direct && !c->instrumented && (c->trav_kernel == 0 || (c->trav_kernel == 2 && c->staged_ok));
    for (int b = 0; b < c->max_path_length; b++) {
        stage_mark(c, 1);
        if (c->capture_bounce == b + 1) {
            CK(cudaMemcpyAsync(c->capture.p, rin, 32 * n_paths, cudaMemcpyDeviceToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_captured_n.p, ctr + CTR_Q + b, sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
        }
        if (fuse && b > 0) { // shadow rays of bounce b-1 + extension rays of bounce b in one persistent launch
            if (c->trav_kernel == 2 && c->staged_ok) {
                const TravOut out = {c->hit_a.p, c->hit_node.p, c->sh_payload.p, c->cl.p, nullptr, c->sh_rays.p, 0, nullptr, class_sort ? c->mat_hist.p + 2 * MAT_CLASSES * b : nullptr};
                launch_staged<4, false, false>(c, c->stream, rin, ctr + CTR_Q + b, ctr + CTR_SH + b - 1, 0, ctr + CTR_WORK + 2 * b, out, nullptr);
            } else
            k_intersect_fused<<<g_trav, 128, 0, c->stream>>>(c->scene, c->tune_p, rin, ctr + CTR_Q + b, c->sh_rays.p, ctr + CTR_SH + b - 1, ctr + CTR_WORK + 2 * b,
                                                              c->hit_a.p, c->hit_node.p, c->sh_payload.p, c->cl.p);
        }
        else if (c->instrumented) launch_intersect<0, false, true>(c, g_trav, c->stream, c->scene, rin, ctr + CTR_Q + b, 0, ctr + CTR_WORK + 2 * b, c->hit_a.p, c->hit_node.p, nullptr, nullptr, nullptr, c->stats.p + 2,
                                                                   class_sort ? c->mat_hist.p + 2 * MAT_CLASSES * b : nullptr);
        else launch_intersect<0, false, false>(c, g_trav, c->stream, c->scene, rin, ctr + CTR_Q + b, 0, ctr + CTR_WORK + 2 * b, c->hit_a.p, c->hit_node.p, nullptr, nullptr, nullptr, nullptr,
                                               class_sort ? c->mat_hist.p + 2 * MAT_CLASSES * b : nullptr);
        stage_mark(c, 2);
        const bool sort_next = c->sort_mode == 1 && b + 1 < c->max_path_length;
        const uint32_t* order = nullptr;
        if (c->sort_mode == 2) { // group this bounce's hits by material class before shading
            unsigned* hist = c->mat_hist.p + 2 * MAT_CLASSES * b;
            k_matsort_classify<<<g_light, 256, 0, c->stream>>>(c->scene, ctr + CTR_Q + b, c->hit_a.p, c->hit_node.p, c->mat_cls.p, hist);
            k_matsort_scatter<<<g_light, 256, 0, c->stream>>>(ctr + CTR_Q + b, c->mat_cls.p, hist, hist + MAT_CLASSES, c->mat_order.p);
            order = c->mat_order.p; launches += 2;
        }
        if (class_sort) { // group the hit records by the class bits the traversal kernel left in them
            unsigned* hist = c->mat_hist.p + 2 * MAT_CLASSES * b;
            k_class_scatter<<<g_light, 256, 0, c->stream>>>(ctr + CTR_Q + b, c->hit_a.p, hist, hist + MAT_CLASSES, c->mat_order.p);
            order = c->mat_order.p; launches++;
        }
        Queues Q = {rin, pin, rout, pout, c->hit_a.p, c->hit_node.p, c->sh_rays.p, c->sh_payload.p, sort_next ? c->sort_keys.p : nullptr, sort_next ? c->sort_hist.p : nullptr, order};
        if (class_sort) {
            const unsigned* hist = c->mat_hist.p + 2 * MAT_CLASSES * b;
            for (int k = 0; k < 4; k++) if (c->class_mask & (1u << k)) { launch_shade(k, g_light, c->stream, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, hist); launches++; }
            launches--;
        }
        else launch_shade(by_class ? single_cls : -1, g_light, c->stream, c->scene, P, st, Q, ctr + CTR_Q + b, ctr + CTR_Q + b + 1, ctr + CTR_SH + b, nullptr);
        if (sort_next) { // counting sort of the next bounce's extension queue by (octant, origin cell)
            stage_mark(c, 4);
            k_sort_scan<<<1, 1024, 0, c->stream>>>(c->sort_hist.p, c->sort_offsets.p);
            k_sort_scatter<<<g_light, 256, 0, c->stream>>>(ctr + CTR_Q + b + 1, c->sort_keys.p, rout, pout, c->sort_offsets.p, rspare, pspare);
            std::swap(rout, rspare); std::swap(pout, pspare);
            launches += 2;
        }
        stage_mark(c, 3);
        if (c->direct && (!fuse || b + 1 == c->max_path_length)) {
            if (c->instrumented) launch_intersect<1, true, true>(c, g_trav, c->stream, c->scene, c->sh_rays.p, ctr + CTR_SH + b, 0, ctr + CTR_WORK + 2 * b + 1, nullptr, nullptr, c->sh_payload.p, c->cl.p, nullptr, c->stats.p + 6);
            else launch_intersect<1, true, false>(c, g_trav, c->stream, c->scene, c->sh_rays.p, ctr + CTR_SH + b, 0, ctr + CTR_WORK + 2 * b + 1, nullptr, nullptr, c->sh_payload.p, c->cl.p, nullptr, nullptr);
            launches++;
        }
        launches += 2;
        std::swap(rin, rout); std::swap(pin, pout);
    }
    stage_mark(c, 4);
    k_finish<<<g_light, 256, 0, c->stream>>>((int)n_paths, st, c->accum, c->w, c->h);
    k_tally<<<1, 32, 0, c->stream>>>(ctr + CTR_Q, ctr + CTR_SH, c->max_path_length, c->stats.p, c->stats.p + 1);
    launches += 2;
    stage_mark(c, 5);
    CK(cudaGetLastError());
    if (variance_after_pass(c, new_trace != 0)) return 1;
    CK(cudaEventRecord(c->ev_stop, c->stream));
    c->events_recorded = true;
    c->n_launches = launches;
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

int ctl_render_passes_tiled(ctl_ctx* c, int new_trace, int n_passes, int tile_w, int tile_h, int part, int n_parts) {
    if (!c) return set_err("null context");
    if (n_passes < 1 || n_passes > 4096) return set_err("n_passes out of range [1,4096]");
    if (tile_w <= 0 || tile_h <= 0 || n_parts <= 0 || part < 0 || part >= n_parts) return set_err("invalid tiling");
    Window W; memset(&W, 0, sizeof(W));
    W.mode = 1; W.tile_w = tile_w; W.tile_h = tile_h; W.part = part; W.n_parts = n_parts; W.n_passes = n_passes;
    W.tiles_x = (c->w + tile_w - 1) / tile_w; W.tiles_y = (c->h + tile_h - 1) / tile_h;
    W.warp_blocks = c->warp_blocks && tile_w % 8 == 0 && tile_h % 4 == 0;
    const int n_tiles = W.tiles_x * W.tiles_y;
    const int n_local = n_tiles > part ? (n_tiles - part + n_parts - 1) / n_parts : 0;
    W.n_slots = n_local * tile_w * tile_h;
    if ((size_t)W.n_slots * n_passes > 0x7fffffffull / 2) return set_err("batch too large: reduce n_passes");
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

int ctl_render_pass_tiled(ctl_ctx* c, int new_trace, int tile_w, int tile_h, int part, int n_parts) {
    return ctl_render_passes_tiled(c, new_trace, 1, tile_w, tile_h, part, n_parts);
}

// ------------------------------------------------------------------ WavefrontPathTracer
__global__ void k_set_u32(unsigned* p, unsigned v) { *p = v; }

// == Tracer<true>::DoPass + WavefrontPathTracer::DoRender (Kernel/Tracer.h:209-248, Integrators/PseudoRealtime/WavefrontPathTracer.cu:166-191).
// One pass = one path per pixel through the DoubleRayBuffer-shaped queue: create -> { intersect primaries (+ last iteration's secondaries),
// iterate } x MaxPathLength.  Unlike the reference nothing crosses the host per bounce (it copies the queue struct to and from the device
// around every kernel, cu:175-188): the queue sizes stay in device counters and an empty iteration costs three empty launches.
int ctl_wavefront_pass(ctl_ctx* c, int new_trace) {
    if (!c) return set_err("null context");
    if (!c->has_scene) return set_err("no scene uploaded");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev_start, c->stream));
    if (new_trace) { CK(cudaMemsetAsync(c->accum, 0, (size_t)c->w * c->h * 7 * sizeof(float), c->stream)); c->passes_done = 0; }
    const uint32_t pass_index = (uint32_t)c->pass_phase + (uint32_t)c->pass_stride * c->passes_done;   // which pass of the (possibly shared) frame this is
    if (c->user_tables) c->user_tables = false;
    else if (generate_tables(c, pass_index, 1)) return 1;
    c->scene.d1 = c->d_tab1.p; c->scene.d2 = (const float2*)c->d_tab2.p;
    c->scene.img_w = c->w; c->scene.img_h = c->h;
    const size_t n = (size_t)c->w * c->h;
    if (n > 0x3fffffffull) return set_err("image too large for the wavefront queue");
    const int n_tiles = (int)((n + WPT_TILE - 1) / WPT_TILE);
    const int mpl = c->max_path_length;
    CK(c->w_thr.ensure(n)); CK(c->w_lxy.ensure(n)); CK(c->w_df.ensure(n)); CK(c->w_misc.ensure(n)); CK(c->w_ray.ensure(2 * n)); CK(c->w_res.ensure(n));
    for (int k = 0; k < 2; k++) { CK(c->w_sec[k].ensure(2 * n)); CK(c->w_sres[k].ensure(n)); }
    CK(c->w_desc.ensure((size_t)mpl * (n_tiles + 1)));
    c->stage_kind.clear();
    unsigned* ctr = c->counters.p;
    CK(cudaMemsetAsync(ctr, 0, CTR_TOTAL * sizeof(unsigned), c->stream));
    CK(cudaMemsetAsync(c->w_desc.p, 0, (size_t)mpl * (n_tiles + 1) * sizeof(unsigned long long), c->stream));
    if (c->instrumented) CK(cudaMemsetAsync(c->stats.p + 2, 0, 8 * sizeof(unsigned long long), c->stream));
    const int g_light = grid_for(c, c->shade_blocks_per_sm), g_trav = grid_for(c, c->trav_blocks_per_sm);
    uint32_t launches = 0;
    stage_mark(c, 0);
    k_set_u32<<<1, 1, 0, c->stream>>>(ctr + CTR_Q, (unsigned)n);
    WptBuf B = {c->w_thr.p, c->w_lxy.p, c->w_df.p, c->w_misc.p, c->w_ray.p, c->w_res.p, nullptr, nullptr};
    k_wpt_create<<<g_light, 256, 0, c->stream>>>(c->scene, B, (int)n);
    launches += 2;
    for (int d = 0; d < mpl; d++) {
        stage_mark(c, 1);
        // FinishIteration (DoubleRayBuffer.h:84-112): the primaries and the secondary rays pushed by iteration d-1
        const bool have_sec = d > 0 && c->direct;
        if (c->instrumented) { // visit counts for the roofline (ctl_get_visit_counts: "extension" = primaries, "shadow" = secondaries): unfused, counting builds
            launch_intersect<2, false, true>(c, g_trav, c->stream, c->scene, (const float4*)c->w_ray.p, ctr + CTR_Q + d, 0, ctr + CTR_WORK + 2 * d, nullptr, nullptr, nullptr, nullptr, (void*)c->w_res.p, c->stats.p + 2);
            launches++;
            if (have_sec) {
                stage_mark(c, 3);
                launch_intersect<2, true, true>(c, g_trav, c->stream, c->scene, (const float4*)c->w_sec[(d - 1) & 1].p, ctr + CTR_SH + d - 1, 0, ctr + CTR_WORK + 2 * d + 1, nullptr, nullptr, nullptr, nullptr,
                                                (void*)c->w_sres[(d - 1) & 1].p, c->stats.p + 6);
                launches++;
            }
        } else if (have_sec && c->fuse_traversal && c->trav_kernel == 2 && c->staged_ok) {
            const TravOut out = {nullptr, nullptr, nullptr, nullptr, (void*)c->w_res.p, (const float4*)c->w_sec[(d - 1) & 1].p, 0, (void*)c->w_sres[(d - 1) & 1].p};
            launch_staged<5, false, false>(c, c->stream, (const float4*)c->w_ray.p, ctr + CTR_Q + d, ctr + CTR_SH + d - 1, 0, ctr + CTR_WORK + 2 * d, out, nullptr);
            launches++;
        } else if (have_sec && c->fuse_traversal && c->trav_kernel == 0) {
            k_intersect_fused_api<<<g_trav, 128, 0, c->stream>>>(c->scene, c->tune_p, (const float4*)c->w_ray.p, ctr + CTR_Q + d, (const float4*)c->w_sec[(d - 1) & 1].p, ctr + CTR_SH + d - 1,
                                                                  ctr + CTR_WORK + 2 * d, (void*)c->w_res.p, (void*)c->w_sres[(d - 1) & 1].p);
            launches++;
        } else {
            launch_intersect<2, false, false>(c, g_trav, c->stream, c->scene, (const float4*)c->w_ray.p, ctr + CTR_Q + d, 0, ctr + CTR_WORK + 2 * d, nullptr, nullptr, nullptr, nullptr, (void*)c->w_res.p, nullptr);
            launches++;
            if (have_sec) {
                stage_mark(c, 3);
                launch_intersect<2, true, false>(c, g_trav, c->stream, c->scene, (const float4*)c->w_sec[(d - 1) & 1].p, ctr + CTR_SH + d - 1, 0, ctr + CTR_WORK + 2 * d + 1, nullptr, nullptr, nullptr, nullptr,
                                                 (void*)c->w_sres[(d - 1) & 1].p, nullptr);
                launches++;
            }
        }
        stage_mark(c, 2);
        B.sec_out = c->w_sec[d & 1].p; B.sec_res = c->w_sres[(d - 1) & 1].p;
        const WptParams P = {d, (int)pass_index + 1, mpl, c->rr_start}; // m_uPassesDone++ precedes DoRender (Kernel/Tracer.h:231-232)
        unsigned long long* desc = c->w_desc.p + (size_t)d * (n_tiles + 1);
        if (c->direct) k_wpt_iterate<true><<<n_tiles, WPT_TILE, 0, c->stream>>>(c->scene, P, B, ctr + CTR_Q + d, ctr + CTR_Q + d + 1, ctr + CTR_SH + d, desc, n_tiles, c->accum);
        else k_wpt_iterate<false><<<n_tiles, WPT_TILE, 0, c->stream>>>(c->scene, P, B, ctr + CTR_Q + d, ctr + CTR_Q + d + 1, ctr + CTR_SH + d, desc, n_tiles, c->accum);
        launches++;
    }
    stage_mark(c, 4);
    k_tally<<<1, 32, 0, c->stream>>>(ctr + CTR_Q, ctr + CTR_SH, mpl, c->stats.p, c->stats.p + 1);
    launches++;
    stage_mark(c, 5);
    CK(cudaGetLastError());
    if (variance_after_pass(c, new_trace != 0)) return 1;
    CK(cudaEventRecord(c->ev_stop, c->stream));
    c->events_recorded = true;
    c->n_launches = launches;
    c->passes_done += 1;
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
static int variance_after_pass(ctl_ctx* c, bool new_trace) { // Tracer<true>::DoPass: Clear on a new trace (Tracer.h:222-226), AddPass after DoRender (:233-237)
    if (!c->variance_buffer) return 0;
    const size_t n = (size_t)c->w * c->h;
    const bool fresh = c->d_var.n < n;
    CK(c->d_var.ensure(n));
    if (new_trace || fresh) CK(cudaMemsetAsync(c->d_var.p, 0, n * sizeof(ctl_pixel_variance_info), c->stream));
    k_variance_update<<<grid_for(c, 8), 256, 0, c->stream>>>(c->d_var.p, c->accum, (int)n, 0.0f /* getSplatScale(): the path tracers never splat */);
    CK(cudaGetLastError());
    return 0;
}
int ctl_read_variance(ctl_ctx* c, ctl_pixel_variance_info* out) {
    if (!c || !out) return set_err("null argument");
    if (!c->variance_buffer || !c->d_var.p) return set_err("PixelVarianceBuffer is off: ctl_set_param_i(ctx, \"PixelVarianceBuffer\", 1) before the passes");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->d_var.p, (size_t)c->w * c->h * sizeof(ctl_pixel_variance_info), cudaMemcpyDeviceToHost, c->stream));
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
    CK(cudaMemcpy(ctr.data(), c->counters.p, CTR_TOTAL * sizeof(unsigned), cudaMemcpyDeviceToHost));
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
    CK(cudaMemcpy(ctr.data(), c->counters.p, CTR_TOTAL * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (int b = 0; b < n && b < MAX_BOUNCES; b++) { if (ext) ext[b] = ctr[CTR_Q + b]; if (sh) sh[b] = ctr[CTR_SH + b]; }
    return 0;
}
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
