// staging.cpp -- derived device-side records of the staged traversal kernel (device/traverse_staged.cuh), built from a reference-layout view.
//
// Nothing here changes the data surface: ctl_scene_view keeps the reference's arrays (KernelDynamicScene, Engine/KernelDynamicScene.h:28-57);
// ctl_upload_scene derives three more arrays from it
//   tri64   : per leaf slot, the Woop record (Engine/TriIntersectorData.h:30-40) and its leaf word (:8-28) in ONE 64-byte record
//   inst    : per node, the inverse-transform rows and the KernelMesh offsets (Engine/Mesh.h:12-19) an instance entry reads (Kernel/TraceHelper.cu:91-99)
//   treelet : the inner nodes with the largest world-space boxes -- the top of the scene-level tree and of the mesh trees -- as a
//             shared-memory image; children inside the image are re-addressed (slot * 4 | 1), children that leave it keep their address.
#include "staging.h"
#include <algorithm>
#include <queue>
#include <unordered_map>
#include <cstring>
#include <cmath>

namespace ctlb {

namespace {
inline bool is_inner(int c) { return c >= 0 && c != CTL_SENTINEL; }
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct B3 { float lo[3], hi[3]; };
inline float area_of(const B3& b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.0f;
    const float a = 2.0f * (dx * dy + dy * dz + dz * dx);
    return a < 3.0e38f ? a : 0.0f;
}
inline B3 child_box(const ctl_bvh_node& n, int k) {
    B3 b;
    if (k == 0) { b.lo[0] = n.a[0]; b.hi[0] = n.a[1]; b.lo[1] = n.a[2]; b.hi[1] = n.a[3]; b.lo[2] = n.c[0]; b.hi[2] = n.c[1]; }
    else { b.lo[0] = n.b[0]; b.hi[0] = n.b[1]; b.lo[1] = n.b[2]; b.hi[1] = n.b[3]; b.lo[2] = n.c[2]; b.hi[2] = n.c[3]; }
    return b;
}
inline B3 box_union(const B3& a, const B3& b) {
    B3 r;
    for (int k = 0; k < 3; k++) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}
// world-space box of a local box under a row-major 4x4 (the 8 corners, as the scene builder does)
inline B3 xf_box(const float* m, const B3& b) {
    B3 r; for (int k = 0; k < 3; k++) { r.lo[k] = 3.0e38f; r.hi[k] = -3.0e38f; }
    for (int c = 0; c < 8; c++) {
        const float p[3] = {c & 1 ? b.hi[0] : b.lo[0], c & 2 ? b.hi[1] : b.lo[1], c & 4 ? b.hi[2] : b.lo[2]};
        float q[4];
        for (int i = 0; i < 4; i++) q[i] = m[4 * i] * p[0] + m[4 * i + 1] * p[1] + m[4 * i + 2] * p[2] + m[4 * i + 3];
        for (int k = 0; k < 3; k++) { const float v = q[k] / q[3]; r.lo[k] = fminf(r.lo[k], v); r.hi[k] = fmaxf(r.hi[k], v); }
    }
    return r;
}
} // namespace

void build_staging_tris(const ctl_scene_view& v, StagedHost& out) {
    out.usable = false; out.why.clear(); out.tri64.clear();
    if (v.n_woop != v.n_tri_index) { out.why = "Woop and leaf-word arrays differ in length"; return; }
    for (uint32_t m = 0; m < v.n_meshes; m++)
        if (v.meshes[m].bvh_tri_offset != 3u * v.meshes[m].bvh_idx_offset) { out.why = "a mesh's Woop offset is not 3 x its leaf-word offset"; return; }
    out.tri64.resize((size_t)v.n_woop * 16);
    for (uint32_t s = 0; s < v.n_woop; s++) {
        float* d = out.tri64.data() + (size_t)s * 16;
        memcpy(d, &v.woop[s], 48);
        d[12] = bitsf(v.tri_index[s]); d[13] = d[14] = d[15] = 0.0f;
    }
    // material class of every leaf slot (for the per-class shade launches): slot -> owning mesh (largest leaf-word offset <= slot) -> triangle -> material
    out.class_mask = 0;
    std::vector<std::pair<uint32_t, uint32_t>> owners; // (leaf-word offset, mesh); re-braided views hold many mesh records over one range, all with the same material block
    for (uint32_t m = 0; m < v.n_meshes; m++) owners.emplace_back(v.meshes[m].bvh_idx_offset, m);
    std::sort(owners.begin(), owners.end());
    size_t oi = 0;
    for (uint32_t s = 0; s < v.n_woop && !owners.empty(); s++) {
        while (oi + 1 < owners.size() && owners[oi + 1].first <= s) oi++;
        const ctl_mesh& M = v.meshes[owners[oi].second];
        const uint64_t tri = (uint64_t)(v.tri_index[s] >> 1) + M.tri_offset;
        uint32_t cls = 0;
        if (tri < v.n_tri_data) {
            const uint64_t mi = (uint64_t)((v.tri_data[tri].w[1] >> 16) & 0xffu) + M.mat_offset;
            if (mi < v.n_materials) {
                const ctl_material& mat = v.materials[mi];
                cls = mat.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR ? 1u + (mat.distr_type & 1u) : (mat.bsdf_type == CTL_BSDF_DIELECTRIC ? 3u : 0u);
            }
        }
        out.class_mask |= 1u << cls;
        out.tri64[(size_t)s * 16 + 13] = bitsf(cls);
    }
    out.usable = true;
}

void build_staging_nodes(const ctl_scene_view& v, int treelet_budget, StagedHost& out) {
    out.inst.assign((size_t)v.n_nodes * 16, 0.0f); out.treelet.clear(); out.tl_nodes = 0; out.scene_root = v.scene_start_node;
    const uint32_t n_bvh = v.n_bvh_nodes;
    // ---- treelet selection: best-first by world-space surface area over the scene tree and every referenced mesh tree
    struct Item { float key; int tree; uint32_t idx; }; // tree: -1 = scene level, else 0 (mesh level; idx = global node index)
    auto cmp = [](const Item& a, const Item& b) { return a.key != b.key ? a.key < b.key : (a.tree != b.tree ? a.tree > b.tree : a.idx > b.idx); };
    std::priority_queue<Item, std::vector<Item>, decltype(cmp)> pq(cmp);
    std::unordered_map<uint64_t, int> slot_of;  // (tree, idx) -> treelet slot
    auto key_of = [](int tree, uint32_t idx) { return ((uint64_t)(tree < 0 ? 1 : 0) << 40) | idx; };
    std::unordered_map<uint32_t, std::vector<uint32_t>> users; // mesh tree base (node units) -> nodes that instance it
    for (uint32_t i = 0; i < v.n_nodes; i++) {
        const uint32_t m = v.nodes[i].mesh_index;
        if (m < v.n_meshes) users[v.meshes[m].bvh_node_offset / 4].push_back(i);
    }
    // mesh-level node `g` (global index) belongs to the tree whose base is the largest base <= g: resolved through the root it was reached from
    struct Sel { int tree; uint32_t idx; uint32_t base; };
    std::vector<Sel> sel;
    std::unordered_map<uint64_t, uint32_t> base_of; // queued mesh-level node -> its tree base
    auto world_area = [&](uint32_t base, const B3& local) {
        float a = 0.0f;
        auto it = users.find(base);
        if (it != users.end()) for (uint32_t ni : it->second) a += area_of(xf_box(v.node_xf + (size_t)ni * 16, local));
        return a;
    };
    if (treelet_budget > 0) {
        if (v.scene_start_node >= 0 && v.n_scene_bvh_nodes && v.n_nodes > 1) {
            const ctl_bvh_node& r = v.scene_bvh_nodes[v.scene_start_node / 4];
            pq.push({area_of(box_union(child_box(r, 0), child_box(r, 1))), -1, (uint32_t)v.scene_start_node / 4});
        }
        for (auto& u : users) {
            if (u.first >= n_bvh) continue;
            const ctl_bvh_node& r = v.bvh_nodes[u.first];
            base_of[u.first] = u.first;
            pq.push({world_area(u.first, box_union(child_box(r, 0), child_box(r, 1))), 0, u.first});
        }
        while (!pq.empty() && (int)sel.size() < treelet_budget) {
            const Item it = pq.top(); pq.pop();
            const uint32_t base = it.tree < 0 ? 0u : base_of[it.idx];
            slot_of[key_of(it.tree, it.idx)] = (int)sel.size();
            sel.push_back({it.tree, it.idx, base});
            const ctl_bvh_node& n = it.tree < 0 ? v.scene_bvh_nodes[it.idx] : v.bvh_nodes[it.idx];
            const int ch[2] = {n.child0, n.child1};
            for (int k = 0; k < 2; k++) {
                if (!is_inner(ch[k])) continue;
                const uint32_t ci = (uint32_t)ch[k] / 4 + base;
                if (it.tree < 0) { if (ci < v.n_scene_bvh_nodes) pq.push({area_of(child_box(n, k)), -1, ci}); }
                else if (ci < n_bvh) { base_of[ci] = base; pq.push({world_area(base, child_box(n, k)), 0, ci}); }
            }
        }
    }
    out.tl_nodes = (int)sel.size();
    out.treelet.assign((size_t)out.tl_nodes * 16, 0.0f);
    for (int s = 0; s < out.tl_nodes; s++) {
        ctl_bvh_node n = sel[s].tree < 0 ? v.scene_bvh_nodes[sel[s].idx] : v.bvh_nodes[sel[s].idx];
        int* ch[2] = {&n.child0, &n.child1};
        for (int k = 0; k < 2; k++) {
            if (!is_inner(*ch[k])) continue;
            auto f = slot_of.find(key_of(sel[s].tree, (uint32_t)*ch[k] / 4 + sel[s].base));
            if (f != slot_of.end()) *ch[k] = f->second * 4 + 1;
        }
        const float* src = (const float*)&n;
        for (int j = 0; j < 4; j++) memcpy(out.treelet.data() + (size_t)(s * 4 + (j ^ ((s >> 1) & 3))) * 4, src + 4 * j, 16);
    }
    if (v.scene_start_node >= 0) { auto f = slot_of.find(key_of(-1, (uint32_t)v.scene_start_node / 4)); if (f != slot_of.end()) out.scene_root = f->second * 4 + 1; }
    // ---- instance records
    out.class_ok = true;
    for (uint32_t i = 0; i < v.n_nodes; i++) if (v.nodes[i].mesh_index < v.n_meshes && v.nodes[i].material_offset != v.meshes[v.nodes[i].mesh_index].mat_offset) out.class_ok = false;
    for (uint32_t i = 0; i < v.n_nodes; i++) {
        float* d = out.inst.data() + (size_t)i * 16;
        const float* inv = v.node_inv_xf + (size_t)i * 16;
        memcpy(d, inv, 48);
        const uint32_t m = v.nodes[i].mesh_index;
        if (m >= v.n_meshes) continue;
        const ctl_mesh& M = v.meshes[m];
        uint32_t root = 0;
        auto f = slot_of.find(key_of(0, M.bvh_node_offset / 4));
        if (f != slot_of.end()) root = (uint32_t)f->second * 4u + 1u;
        if (!(inv[12] == 0.0f && inv[13] == 0.0f && inv[14] == 0.0f && inv[15] == 1.0f)) root |= 2u;
        d[12] = bitsf(M.bvh_node_offset); d[13] = bitsf(M.bvh_idx_offset); d[14] = bitsf(M.tri_offset); d[15] = bitsf(root);
    }
}

} // namespace ctlb
