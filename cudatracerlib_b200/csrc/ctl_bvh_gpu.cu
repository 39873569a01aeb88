// ctl_bvh_gpu.cu -- GPU BVH build entry points of the C ABI: the agglomerative builder of csrc/bvh_ploc.cuh (default) and the LBVH of
// csrc/bvh_build.cuh (algorithm 0: fastest build, 1.1-1.3x slower traversal), both in the reference layout.
#include "ctl_internal.h"
#include "bvh_ploc.cuh"
#include "bvh_presplit.cuh"
#include <cstdlib>

extern "C" {

// ------------------------------------------------------------------ GPU BVH build
// LBVH of one triangle mesh in the reference layout; host arrays in, host arrays out (nodes_out: capacity >= max(1, n_tris) entries).
// depth of an emitted tree (longest root-to-leaf chain of inner nodes); the traversal stack holds 64 entries for scene level + mesh level
static int tree_depth(const ctl_bvh_node* nodes, uint32_t n_nodes) {
    if (!n_nodes || nodes[0].child1 == (int)CTL_SENTINEL) return 1;
    std::vector<std::pair<uint32_t, int>> st; st.push_back({0u, 1});
    int depth = 0;
    while (!st.empty()) {
        const auto [i, d] = st.back(); st.pop_back();
        if (d > depth) depth = d;
        if (d > 4096) break;
        for (const int c : {nodes[i].child0, nodes[i].child1}) if (c >= 0 && (uint32_t)c / 4 < n_nodes) st.push_back({(uint32_t)c / 4, d + 1});
    }
    return depth;
}

struct GpuBuildStats { float sah_cost = 0; int rounds = 0; uint32_t refs = 0; };
// max_growth > 0: triangles are pre-split into at most (1 + max_growth) * n_tris references (bvh_presplit.cuh); the outputs hold `capacity` entries each.
static int build_gpu(int device, const float* verts9, uint32_t n_tris, int algorithm, int radius, float max_growth, uint32_t capacity, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out,
                     uint32_t* index_out, uint32_t* n_slots_out, float* build_ms, GpuBuildStats* stats) {
    using namespace ctlbvh;
    if (!verts9 || !n_tris || !nodes_out || !n_nodes_out || !woop_out || !index_out) return set_err("null / empty argument");
    if (n_tris > 0x3fffffffu / 4) return set_err("too many triangles");
    if (capacity < n_tris) return set_err("output capacity below the triangle count");
    if (radius <= 0) radius = 16;
    if (radius > PLOC_MAX_RADIUS) radius = PLOC_MAX_RADIUS;
    CK(cudaSetDevice(device));
    int n = (int)n_tris;   // references from here on (== triangles unless pre-split)
    // ---- pre-splitting: count the pieces (lowering the scale until the budget holds), scan, emit the references
    DevBuf<float> d_verts; DevBuf<float4> d_tboxes, d_rboxes; DevBuf<unsigned> d_sbox, d_pieces; DevBuf<uint32_t> d_reftri;
    const unsigned sbox_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, eA = nullptr, eB = nullptr;   // device time = (e0 -> eA: boxes + pre-splitting) + (eB -> e1: sort, tree, emission); the allocations in between are host time
    auto free_pre = [&]() { d_verts.release(); d_tboxes.release(); d_rboxes.release(); d_sbox.release(); d_pieces.release(); d_reftri.release(); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (eA) cudaEventDestroy(eA); if (eB) cudaEventDestroy(eB); };
#define CKP0(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { free_pre(); char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return set_err(b_); } } while (0)
    CKP0(d_verts.upload(verts9, (size_t)n_tris * 9)); CKP0(d_tboxes.ensure((size_t)n_tris * 2)); CKP0(d_sbox.ensure(6));
    CKP0(cudaEventCreate(&e0)); CKP0(cudaEventCreate(&e1)); CKP0(cudaEventCreate(&eA)); CKP0(cudaEventCreate(&eB));
    CKP0(cudaEventRecord(e0, st));
    CKP0(cudaMemcpyAsync(d_sbox.p, sbox_init, sizeof(sbox_init), cudaMemcpyHostToDevice, st));
    k_tri_boxes<<<(n_tris + 255) / 256, 256, 0, st>>>(d_verts.p, n_tris, d_tboxes.p, d_sbox.p);
    bool split = false;
    const uint32_t budget = (uint32_t)std::min<double>((double)capacity, (double)n_tris * (1.0 + (double)std::max(0.0f, max_growth)));
    if (max_growth > 0.0f && n_tris > (uint32_t)MAX_LEAF && budget > n_tris) {
        CKP0(d_pieces.ensure((size_t)n_tris + 1));
        // the largest scale of the piece count (<= the requested one) whose references fit the budget: bisection, one count pass per probe
        const char* se = getenv("CTL_GPU_SPLIT_SCALE");
        const float want = se ? std::min(4.0f, std::max(0.05f, (float)atof(se))) : 1.0f;
        auto count_at = [&](float sc, unsigned& tot) -> cudaError_t {
            k_split_count<<<(n_tris + 256) / 256, 256, 0, st>>>(d_verts.p, n_tris, sc, SPLIT_MAX_PIECES, d_pieces.p);
            k_scan_exclusive<<<1, 1024, 0, st>>>(d_pieces.p, n_tris + 1);
            cudaError_t e = cudaMemcpyAsync(&tot, d_pieces.p + n_tris, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
            return e != cudaSuccess ? e : cudaStreamSynchronize(st);
        };
        float scale = want; unsigned total = 0;
        CKP0(count_at(scale, total));
        if (total > budget) {
            float lo_s = 0.0f, hi_s = want;
            for (int it = 0; it < 10; it++) { const float mid = 0.5f * (lo_s + hi_s); unsigned t2 = 0; CKP0(count_at(mid, t2)); if (t2 <= budget) lo_s = mid; else hi_s = mid; }
            scale = lo_s; total = 0;
            if (scale > 0.0f) CKP0(count_at(scale, total));   // leaves the scan of the chosen scale in d_pieces
        }
        if (total > n_tris) {
            CKP0(d_rboxes.ensure((size_t)total * 2)); CKP0(d_reftri.ensure(total));
            k_split_emit<<<(n_tris + 127) / 128, 128, 0, st>>>(d_verts.p, n_tris, scale, SPLIT_MAX_PIECES, d_pieces.p, d_tboxes.p, d_rboxes.p, d_reftri.p);
            n = (int)total; split = true;
        }
    }
    if (stats) stats->refs = (uint32_t)n;
    CKP0(cudaEventRecord(eA, st));
    const int nb_sort = (n + SORT_TILE - 1) / SORT_TILE;
    DevBuf<float4> d_nbox; DevBuf<unsigned> d_counts, d_flags, d_emit; DevBuf<uint32_t> d_k0, d_k1, d_v0, d_v1, d_index;
    struct { float4* p; } d_boxes = {split ? d_rboxes.p : d_tboxes.p};   // boxes of the references the builders work on
    DevBuf<int> d_left, d_right, d_pint, d_pleaf, d_first, d_last; DevBuf<ctl_bvh_node> d_nodes; DevBuf<ctl_woop_tri> d_woop; DevBuf<unsigned char> d_lastflag, d_collapse; DevBuf<float> d_cost;
    DevBuf<int> d_cid0, d_cid1, d_nn, d_count, d_ecount, d_slot, d_pleaf2, d_pst; DevBuf<float4> d_cb0, d_cb1; DevBuf<unsigned long long> d_scan; DevBuf<uint32_t> d_vals2;   // agglomerative builder
    int* h_st_owned = nullptr;
    auto free_all = [&]() { if (h_st_owned) cudaFreeHost(h_st_owned); h_st_owned = nullptr; free_pre(); d_pst.release(); d_cid0.release(); d_cid1.release(); d_nn.release(); d_count.release(); d_ecount.release(); d_slot.release(); d_pleaf2.release(); d_cb0.release(); d_cb1.release(); d_scan.release(); d_vals2.release(); d_nbox.release(); d_counts.release(); d_flags.release(); d_emit.release(); d_k0.release(); d_k1.release(); d_v0.release(); d_v1.release();
                            d_index.release(); d_left.release(); d_right.release(); d_pint.release(); d_pleaf.release(); d_first.release(); d_last.release(); d_nodes.release(); d_woop.release(); d_lastflag.release(); d_collapse.release(); d_cost.release(); };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { free_all(); char b_[512]; snprintf(b_, sizeof(b_), "In file %s at line %d : %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return set_err(b_); } } while (0)
    CKF(d_nbox.ensure((size_t)n * 2)); CKF(d_counts.ensure((size_t)256 * nb_sort));
    CKF(d_flags.ensure((size_t)n)); CKF(d_emit.ensure((size_t)n + 1)); CKF(d_k0.ensure(n)); CKF(d_k1.ensure(n)); CKF(d_v0.ensure(n)); CKF(d_v1.ensure(n)); CKF(d_index.ensure(n));
    CKF(d_left.ensure(n)); CKF(d_right.ensure(n)); CKF(d_pint.ensure(n)); CKF(d_pleaf.ensure(n)); CKF(d_first.ensure(n)); CKF(d_last.ensure(n)); CKF(d_nodes.ensure((size_t)n)); CKF(d_woop.ensure(n)); CKF(d_lastflag.ensure(n)); CKF(d_collapse.ensure(n)); CKF(d_cost.ensure(n));
    if (n > MAX_LEAF && algorithm == 1) {   // buffers of the agglomerative builder
        CKF(d_cid0.ensure(n)); CKF(d_cid1.ensure(n)); CKF(d_nn.ensure(n)); CKF(d_count.ensure(n)); CKF(d_ecount.ensure(n)); CKF(d_slot.ensure(n)); CKF(d_pleaf2.ensure(n));
        CKF(d_cb0.ensure((size_t)n * 2)); CKF(d_cb1.ensure((size_t)n * 2)); CKF(d_scan.ensure((size_t)n + 1)); CKF(d_vals2.ensure(n));
        CKF(d_pst.ensure(4));
    }
    int* h_st = nullptr;   // pinned: round state of the agglomerative builder (per call: builds may run concurrently on several devices)
    CKF(cudaHostAlloc((void**)&h_st, 4 * sizeof(int), cudaHostAllocDefault)); h_st_owned = h_st;
    CKF(cudaEventRecord(eB, st));
    CKF(cudaMemsetAsync(d_flags.p, 0, (size_t)n * 4, st)); CKF(cudaMemsetAsync(d_lastflag.p, 0, (size_t)n, st));
    const int g = (n + 255) / 256;
    const uint32_t n_refs = (uint32_t)n;
    k_morton<<<g, 256, 0, st>>>(d_boxes.p, n_refs, d_sbox.p, d_k0.p, d_v0.p);
    uint32_t *kin = d_k0.p, *kout = d_k1.p, *vin = d_v0.p, *vout = d_v1.p;
    for (int pass = 0; pass < 4; pass++) {
        k_sort_hist<<<nb_sort, SORT_THREADS, 0, st>>>(kin, n_refs, 8 * pass, d_counts.p, nb_sort);
        k_scan_exclusive<<<1, 1024, 0, st>>>(d_counts.p, (uint32_t)(256 * nb_sort));
        k_sort_scatter<<<nb_sort, SORT_THREADS, 0, st>>>(kin, vin, n_refs, 8 * pass, d_counts.p, nb_sort, kout, vout);
        std::swap(kin, kout); std::swap(vin, vout);
    }
    uint32_t n_nodes = 1;
    if (n > MAX_LEAF && algorithm == 1) {
        k_ploc_init<<<g, 256, 0, st>>>(n, vin, d_boxes.p, d_cid0.p, d_cb0.p);
        int *cin = d_cid0.p, *cout = d_cid1.p; float4 *bin = d_cb0.p, *bout = d_cb1.p;
        // rounds (nearest partner in the window, mutual pairs merge, survivors compact in order) in groups of 6 without a host round trip; a round on one
        // cluster is a no-op, so overshooting the end is harmless
        h_st[0] = n; h_st[1] = n - 2; h_st[2] = 0; h_st[3] = 0;
        CKF(cudaMemcpyAsync(d_pst.p, h_st, 4 * sizeof(int), cudaMemcpyHostToDevice, st));
        int nc = n, rounds = 0;
        while (nc > 1) {
            const int gc = (nc + 255) / 256;
            for (int k = 0; k < 6; k++) {
                k_ploc_nn<<<gc, 256, 0, st>>>(d_pst.p, radius, bin, d_nn.p);
                k_ploc_flags<<<(nc + 256) / 256, 256, 0, st>>>(d_pst.p, d_nn.p, d_scan.p);
                k_scan_exclusive64<<<1, 1024, 0, st>>>(d_scan.p, d_pst.p);
                k_ploc_merge<<<gc, 256, 0, st>>>(d_pst.p, d_nn.p, d_scan.p, cin, bin, cout, bout, d_left.p, d_right.p, d_pint.p, d_pleaf.p);
                k_ploc_advance<<<1, 1, 0, st>>>(d_pst.p, d_scan.p);
                std::swap(cin, cout); std::swap(bin, bout);
            }
            CKF(cudaMemcpyAsync(h_st, d_pst.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
            CKF(cudaStreamSynchronize(st));
            if (h_st[0] >= nc || h_st[0] < 1) { free_all(); return set_err("agglomerative build made no progress (internal error)"); }
            nc = h_st[0]; rounds = h_st[2];
        }
        if (h_st[1] != -1) { free_all(); return set_err("agglomerative build: node count mismatch (internal error)"); }
        if (stats) stats->rounds = rounds;
        k_fit_counts<<<g, 256, 0, st>>>(d_boxes.p, vin, n, d_left.p, d_right.p, d_pint.p, d_pleaf.p, d_flags.p, d_nbox.p, d_cost.p, d_collapse.p, d_count.p, d_ecount.p);
        k_tree_order<<<(2 * n - 1 + 255) / 256, 256, 0, st>>>(n, d_left.p, d_right.p, d_pint.p, d_pleaf.p, d_count.p, d_ecount.p, d_first.p, d_last.p, d_emit.p, d_slot.p);
        k_leaf_remap<<<g, 256, 0, st>>>(n, d_slot.p, vin, d_pleaf.p, d_vals2.p, d_pleaf2.p, d_left.p, d_right.p);
        vin = d_vals2.p;
        k_emit_nodes<<<g, 256, 0, st>>>(n, d_left.p, d_right.p, d_pint.p, d_first.p, d_last.p, d_emit.p /* pre-order index */, d_boxes.p, vin, d_nbox.p, d_collapse.p, d_nodes.p, d_lastflag.p);
        CKF(cudaMemcpyAsync(&n_nodes, d_ecount.p, 4, cudaMemcpyDeviceToHost, st));
    } else if (n > MAX_LEAF) {
        k_radix_tree<<<g, 256, 0, st>>>(kin, n, d_left.p, d_right.p, d_pint.p, d_pleaf.p, d_first.p, d_last.p);
        k_fit_boxes<<<g, 256, 0, st>>>(d_boxes.p, vin, n, d_left.p, d_right.p, d_pint.p, d_pleaf.p, d_first.p, d_last.p, d_flags.p, d_nbox.p, d_cost.p, d_collapse.p);
        k_mark_emitted<<<g, 256, 0, st>>>(n, d_pint.p, d_first.p, d_last.p, d_collapse.p, d_emit.p);
        k_scan_exclusive<<<1, 1024, 0, st>>>(d_emit.p, (uint32_t)n);
        k_emit_nodes<<<g, 256, 0, st>>>(n, d_left.p, d_right.p, d_pint.p, d_first.p, d_last.p, d_emit.p, d_boxes.p, vin, d_nbox.p, d_collapse.p, d_nodes.p, d_lastflag.p);
        CKF(cudaMemcpyAsync(&n_nodes, d_emit.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    } else {
        k_single_leaf_root<<<1, 32, 0, st>>>(d_sbox.p, n, d_nodes.p, d_lastflag.p);
    }
    k_emit_tris<<<g, 256, 0, st>>>(d_verts.p, vin, n, d_lastflag.p, d_woop.p, d_index.p, split ? d_reftri.p : nullptr);
    if (stats && n > MAX_LEAF) CKF(cudaMemcpyAsync(&stats->sah_cost, d_cost.p, sizeof(float), cudaMemcpyDeviceToHost, st));   // SAH cost of the root (C_inner 1.2, C_tri 1, unnormalised)
    CKF(cudaGetLastError());
    CKF(cudaEventRecord(e1, st));
    CKF(cudaStreamSynchronize(st));
    float ms = 0, ms_pre = 0; CKF(cudaEventElapsedTime(&ms_pre, e0, eA)); CKF(cudaEventElapsedTime(&ms, eB, e1)); ms += ms_pre;
    CKF(cudaMemcpy(nodes_out, d_nodes.p, (size_t)n_nodes * sizeof(ctl_bvh_node), cudaMemcpyDeviceToHost));
    CKF(cudaMemcpy(woop_out, d_woop.p, (size_t)n * sizeof(ctl_woop_tri), cudaMemcpyDeviceToHost));
    CKF(cudaMemcpy(index_out, d_index.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    free_all();
#undef CKF
#undef CKP0
    *n_nodes_out = n_nodes;
    if (n_slots_out) *n_slots_out = n_refs;
    if (build_ms) *build_ms = ms;
    return 0;
}

// algorithm 0 = LBVH, 1 = agglomerative (PLOC), 2 = both, keep the tree with the lower SAH cost; max_growth > 0: triangle pre-splitting
int ctl_bvh_build_gpu_split(int device, const float* verts9, uint32_t n_tris, int algorithm, int radius, float max_growth, uint32_t capacity, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out,
                            ctl_woop_tri* woop_out, uint32_t* index_out, uint32_t* n_slots_out, float* build_ms) {
    if (algorithm < 0 || algorithm > 2) return set_err("algorithm must be 0 (LBVH), 1 (agglomerative, PLOC) or 2 (both, lower SAH cost wins)");
    if (!(max_growth >= 0.0f) || max_growth > 8.0f) return set_err("max_growth out of range [0, 8]");
    const bool verbose = getenv("CTL_GPU_BUILDER_VERBOSE") != nullptr;
    GpuBuildStats sa, sb; float ms_a = 0, ms_b = 0;
    if (build_gpu(device, verts9, n_tris, algorithm == 0 ? 0 : 1, radius, max_growth, capacity, nodes_out, n_nodes_out, woop_out, index_out, n_slots_out, &ms_a, &sa)) return 1;
    if (build_ms) *build_ms = ms_a;
    if (n_tris <= (uint32_t)ctlbvh::MAX_LEAF) return 0;
    const int depth = tree_depth(nodes_out, *n_nodes_out);
    if (verbose) fprintf(stderr, "[ctl gpu builder] %s: %u triangles, %u references, %d rounds, %u nodes, depth %d, SAH cost %.6g, %.2f ms\n", algorithm == 0 ? "LBVH" : "PLOC", n_tris, sa.refs, sa.rounds, *n_nodes_out, depth, sa.sah_cost, ms_a);
    if (algorithm == 0) return depth <= 60 ? 0 : set_err("GPU-built tree deeper than the traversal stack allows");
    if (algorithm == 1 && depth <= 56) return 0;   // (deeper: pathological input, merge chains -- the LBVH's depth is bounded by the key length)
    std::vector<ctl_bvh_node> nodes2(capacity); std::vector<ctl_woop_tri> woop2(capacity); std::vector<uint32_t> index2(capacity); uint32_t nn2 = 0, ns2 = 0;
    if (build_gpu(device, verts9, n_tris, 0, 0, max_growth, capacity, nodes2.data(), &nn2, woop2.data(), index2.data(), &ns2, &ms_b, &sb)) return 1;
    if (verbose) fprintf(stderr, "[ctl gpu builder] LBVH: %u nodes, SAH cost %.6g, %.2f ms\n", nn2, sb.sah_cost, ms_b);
    if (build_ms) *build_ms = ms_a + ms_b;
    if (depth > 56 || sb.sah_cost < sa.sah_cost) {
        memcpy(nodes_out, nodes2.data(), (size_t)nn2 * sizeof(ctl_bvh_node)); memcpy(woop_out, woop2.data(), (size_t)ns2 * sizeof(ctl_woop_tri)); memcpy(index_out, index2.data(), (size_t)ns2 * 4);
        *n_nodes_out = nn2; if (n_slots_out) *n_slots_out = ns2;
    }
    return 0;
}

int ctl_bvh_build_gpu_ex(int device, const float* verts9, uint32_t n_tris, int algorithm, int radius, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms) {
    return ctl_bvh_build_gpu_split(device, verts9, n_tris, algorithm, radius, 0.0f, n_tris, nodes_out, n_nodes_out, woop_out, index_out, nullptr, build_ms);
}

// Default builder: the agglomerative one; CTL_GPU_BUILDER=lbvh selects the LBVH, =auto builds both and keeps the tree with the lower SAH cost;
// CTL_PLOC_RADIUS: the search window (default 16).  One reference per triangle (no pre-splitting: the outputs hold n_tris entries).
static int env_algorithm() { const char* a = getenv("CTL_GPU_BUILDER"); const std::string alg = a ? a : "ploc"; return alg == "lbvh" ? 0 : alg == "auto" ? 2 : 1; }
int ctl_bvh_build_gpu(int device, const float* verts9, uint32_t n_tris, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms) {
    const char* r = getenv("CTL_PLOC_RADIUS");
    return ctl_bvh_build_gpu_ex(device, verts9, n_tris, env_algorithm(), r ? atoi(r) : 0, nodes_out, n_nodes_out, woop_out, index_out, build_ms);
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
    // CTL_GPU_SPLIT: reference budget of the triangle pre-splitting as growth over the triangle count (default 3 = at most four times the triangles of a mesh -- only
    // meshes of slivers get near it; 0 = off); CTL_GPU_SPLIT_SCALE: multiplier of the pieces per triangle (default 1)
    const char* ge = getenv("CTL_GPU_SPLIT"); const char* re = getenv("CTL_PLOC_RADIUS");
    const float growth = ge ? std::min(8.0f, std::max(0.0f, (float)atof(ge))) : 3.0f; const int radius = re ? atoi(re) : 0;
    for (size_t mi = 0; mi < S.meshes.size(); mi++) {
        const uint32_t nt = (uint32_t)(S.mesh_verts9[mi].size() / 9);
        const uint32_t cap = nt + (uint32_t)((double)nt * growth) + 1;
        std::vector<ctl_bvh_node> nodes(cap); std::vector<ctl_woop_tri> woop(cap); std::vector<uint32_t> index(cap);
        uint32_t nn = 0, ns = 0; float ms = 0;
        if (ctl_bvh_build_gpu_split(device, S.mesh_verts9[mi].data(), nt, env_algorithm(), radius, growth, cap, nodes.data(), &nn, woop.data(), index.data(), &ns, &ms)) return 1;
        woop.resize(ns); index.resize(ns);
        total += ms;
        if (getenv("CTL_LBVH_OPTIMIZE")) { nodes.resize(nn); ctlb::optimize_bvh(nodes); }   // experiment for round 2: the mesh trees' host post-pass (re-insertion + rotations) on the LBVH; node count unchanged
        meshes[mi].bvh_node_offset = (uint32_t)all_nodes.size() * 4;
        meshes[mi].bvh_tri_offset = (uint32_t)all_woop.size() * 3;
        meshes[mi].bvh_idx_offset = (uint32_t)all_index.size();
        slot_of_tri[mi].assign(nt, 0);
        for (uint32_t k = 0; k < ns; k++) slot_of_tri[mi][index[k] >> 1] = meshes[mi].bvh_idx_offset + k;   // any slot of a pre-split triangle: they hold the same Woop record
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

} // extern "C"
