// ctl_bvh_gpu.cu -- GPU BVH build entry points of the C ABI: the agglomerative builder of csrc/bvh_ploc.cuh (default) and the LBVH of
// csrc/bvh_build.cuh (algorithm 0: fastest build, 1.1-1.3x slower traversal), both in the reference layout.
#include "ctl_internal.h"
#include "bvh_ploc.cuh"
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

int ctl_bvh_build_gpu_ex(int device, const float* verts9, uint32_t n_tris, int algorithm, int radius, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms) {
    using namespace ctlbvh;
    if (!verts9 || !n_tris || !nodes_out || !n_nodes_out || !woop_out || !index_out) return set_err("null / empty argument");
    if (n_tris > 0x3fffffffu) return set_err("too many triangles");
    if (algorithm < 0 || algorithm > 1) return set_err("algorithm must be 0 (LBVH) or 1 (agglomerative, PLOC)");
    if (radius <= 0) radius = 16;
    if (radius > PLOC_MAX_RADIUS) radius = PLOC_MAX_RADIUS;
    CK(cudaSetDevice(device));
    const int n = (int)n_tris;
    const int nb_sort = (n + SORT_TILE - 1) / SORT_TILE;
    DevBuf<float> d_verts; DevBuf<float4> d_boxes, d_nbox; DevBuf<unsigned> d_sbox, d_counts, d_flags, d_emit; DevBuf<uint32_t> d_k0, d_k1, d_v0, d_v1, d_index;
    DevBuf<int> d_left, d_right, d_pint, d_pleaf, d_first, d_last; DevBuf<ctl_bvh_node> d_nodes; DevBuf<ctl_woop_tri> d_woop; DevBuf<unsigned char> d_lastflag, d_collapse; DevBuf<float> d_cost;
    DevBuf<int> d_cid0, d_cid1, d_nn, d_count, d_ecount, d_slot, d_pleaf2; DevBuf<float4> d_cb0, d_cb1; DevBuf<unsigned long long> d_scan; DevBuf<uint32_t> d_vals2;   // agglomerative builder
    auto free_all = [&]() { d_cid0.release(); d_cid1.release(); d_nn.release(); d_count.release(); d_ecount.release(); d_slot.release(); d_pleaf2.release(); d_cb0.release(); d_cb1.release(); d_scan.release(); d_vals2.release(); d_verts.release(); d_boxes.release(); d_nbox.release(); d_sbox.release(); d_counts.release(); d_flags.release(); d_emit.release(); d_k0.release(); d_k1.release(); d_v0.release(); d_v1.release();
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
    if (n > MAX_LEAF && algorithm == 1) {
        CKF(d_cid0.ensure(n)); CKF(d_cid1.ensure(n)); CKF(d_nn.ensure(n)); CKF(d_count.ensure(n)); CKF(d_ecount.ensure(n)); CKF(d_slot.ensure(n)); CKF(d_pleaf2.ensure(n));
        CKF(d_cb0.ensure((size_t)n * 2)); CKF(d_cb1.ensure((size_t)n * 2)); CKF(d_scan.ensure((size_t)n + 1)); CKF(d_vals2.ensure(n));
        k_ploc_init<<<g, 256, 0, st>>>(n, vin, d_boxes.p, d_cid0.p, d_cb0.p);
        int *cin = d_cid0.p, *cout = d_cid1.p; float4 *bin = d_cb0.p, *bout = d_cb1.p;
        int nc = n, next_id = n - 2, rounds = 0;
        while (nc > 1) {   // one round: nearest partner in the window, mutual pairs merge, survivors compact (order kept)
            const int gc = (nc + 255) / 256;
            k_ploc_nn<<<gc, 256, 0, st>>>(nc, radius, bin, d_nn.p);
            k_ploc_flags<<<(nc + 256) / 256, 256, 0, st>>>(nc, d_nn.p, d_scan.p);
            k_scan_exclusive64<<<1, 1024, 0, st>>>(d_scan.p, (uint32_t)nc + 1u);
            k_ploc_merge<<<gc, 256, 0, st>>>(nc, d_nn.p, d_scan.p, next_id, cin, bin, cout, bout, d_left.p, d_right.p, d_pint.p, d_pleaf.p);
            unsigned long long tot = 0;
            CKF(cudaMemcpyAsync(&tot, d_scan.p + nc, sizeof(tot), cudaMemcpyDeviceToHost, st));
            CKF(cudaStreamSynchronize(st));
            const int merges = (int)(uint32_t)(tot >> 32), left_over = (int)(uint32_t)tot;
            if (merges <= 0 || left_over != nc - merges) { free_all(); return set_err("agglomerative build made no progress (internal error)"); }
            nc = left_over; next_id -= merges; rounds++;
            std::swap(cin, cout); std::swap(bin, bout);
        }
        (void)rounds;
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
    if (algorithm == 1 && tree_depth(nodes_out, n_nodes) > 56) {   // pathological input (merge chains): the LBVH's depth is bounded by the key length
        float ms2 = 0;
        const int rc = ctl_bvh_build_gpu_ex(device, verts9, n_tris, 0, 0, nodes_out, n_nodes_out, woop_out, index_out, &ms2);
        if (build_ms) *build_ms = ms + ms2;
        return rc;
    }
    return 0;
}

// Default builder: the agglomerative one; CTL_GPU_BUILDER=lbvh selects the LBVH, CTL_PLOC_RADIUS the search window (default 16).
int ctl_bvh_build_gpu(int device, const float* verts9, uint32_t n_tris, ctl_bvh_node* nodes_out, uint32_t* n_nodes_out, ctl_woop_tri* woop_out, uint32_t* index_out, float* build_ms) {
    const char* a = getenv("CTL_GPU_BUILDER"); const char* r = getenv("CTL_PLOC_RADIUS");
    return ctl_bvh_build_gpu_ex(device, verts9, n_tris, (a && std::string(a) == "lbvh") ? 0 : 1, r ? atoi(r) : 0, nodes_out, n_nodes_out, woop_out, index_out, build_ms);
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

} // extern "C"
