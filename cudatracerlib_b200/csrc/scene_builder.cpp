// scene_builder.cpp -- see scene_builder.h for the reference citations.
#include "scene_builder.h"
#include <cstdlib>
#include <string>
#include <algorithm>
#include <cstdio>
#include <stdexcept>
#include <atomic>
#include <exception>
#include <thread>

namespace ctlb {

// ------------------------------------------------------------------ encoders

void decode_woop(const ctl_woop_tri& w, V3& v0, V3& v1, V3& v2) { // TriIntersectorData.cu:20-32
    M4 m = M4::identity();
    for (int k = 0; k < 4; k++) { m(0, k) = w.b[k]; m(1, k) = w.c[k]; m(2, k) = w.a[k]; }
    m(2, 3) *= -1.0f;
    M4 i = m.inverse();
    V3 e02(i(0, 0), i(1, 0), i(2, 0)), e12(i(0, 1), i(1, 1), i(2, 1));
    v2 = V3(i(0, 3), i(1, 3), i(2, 3));
    v0 = v2 + e02;
    v1 = v2 + e12;
}

void encode_tri_data(const V3 p[3], const V3 n[3], const float uv[6], uint32_t mat, ctl_tri_data* out) {
    uint16_t h[6];
    for (int i = 0; i < 6; i++) h[i] = float_to_half(uv[i]);
    out->w[5] = h[0] | ((uint32_t)h[1] << 16);
    out->w[6] = h[2] | ((uint32_t)h[3] << 16);
    out->w[7] = h[4] | ((uint32_t)h[5] << 16);
    // setData reads the uvs back through the half encoding (TriangleData.cu:37-40)
    float t0x = half_to_float(h[0]), t0y = half_to_float(h[1]), t1x = half_to_float(h[2]), t1y = half_to_float(h[3]);
    float t2x = half_to_float(h[4]), t2y = half_to_float(h[5]);
    V3 dP1 = p[1] - p[0], dP2 = p[2] - p[0];
    float du1 = t1x - t0x, dv1 = t1y - t0y, du2 = t2x - t0x, dv2 = t2y - t0y;
    float det = du1 * dv2 - dv1 * du2;
    V3 dpdu, dpdv;
    if (det == 0) {
        V3 nn = normalize(cross(dP1, dP2));
        coordinate_system(nn, dpdu, dpdv);
    } else {
        float inv = 1.0f / det;
        dpdu = (dv2 * dP1 - dv1 * dP2) * inv;
        dpdv = (-du2 * dP1 + du1 * dP2) * inv;
    }
    uint16_t a[3] = {float_to_half(dpdu.x), float_to_half(dpdu.y), float_to_half(dpdu.z)};
    uint16_t b[3] = {float_to_half(dpdv.x), float_to_half(dpdv.y), float_to_half(dpdv.z)};
    out->w[0] = encode_normal(n[0]) | ((uint32_t)encode_normal(n[1]) << 16);
    out->w[1] = encode_normal(n[2]) | ((mat & 0xffu) << 16);
    out->w[2] = a[0] | ((uint32_t)a[1] << 16);
    out->w[3] = a[2] | ((uint32_t)b[0] << 16);
    out->w[4] = b[1] | ((uint32_t)b[2] << 16);
}

void compute_vertex_normals(const std::vector<V3>& verts, const std::vector<uint32_t>& idx, std::vector<V3>& normals) {
    // "sphere inscribed polytope" weighting, Engine/Mesh.cpp:151-190
    normals.assign(verts.size(), V3(0.0f));
    auto nor = [](V3 base, V3 n1, V3 n2) { return cross(n1 - base, n2 - base) / (len_sqr(n1 - base) * len_sqr(n2 - base)); };
    for (size_t f = 0; f < idx.size() / 3; f++) {
        uint32_t i1 = idx[f * 3], i2 = idx[f * 3 + 1], i3 = idx[f * 3 + 2];
        V3 v1 = verts[i1], v2 = verts[i2], v3 = verts[i3];
        normals[i1] = normals[i1] + nor(v1, v3, v2);
        normals[i2] = normals[i2] + nor(v2, v1, v3);
        normals[i3] = normals[i3] + nor(v3, v2, v1);
    }
    for (auto& n : normals) n = normalize(n);
}

V3 shading_normal_at(const ctl_tri_data& td, const M4& l2w, float u, float v) { // TriangleData.cu:75-90
    V3 na = decode_normal(td.w[0] & 0xffff), nb = decode_normal(td.w[0] >> 16), nc = decode_normal(td.w[1] & 0xffff);
    float w = 1.0f - u - v;
    V3 n = normalize(u * na + v * nb + w * nc);
    V3 dpdu(half_to_float(td.w[2] & 0xffff), half_to_float(td.w[2] >> 16), half_to_float(td.w[3] & 0xffff));
    V3 s = dpdu - n * dot(n, dpdu);
    V3 t = cross(s, n);
    s = l2w.transform_dir(s);
    t = l2w.transform_dir(t);
    return normalize(cross(t, s));
}

// ------------------------------------------------------------------ BVH build (binned SAH)

namespace {
struct BuildCtx {
    const std::vector<Box>& boxes;
    std::vector<V3> centroids;
    std::vector<uint32_t> prims;
    int max_leaf;
    std::vector<ctl_bvh_node>& nodes;
    std::vector<uint32_t>& ordered;
    std::vector<uint8_t>& last;
    BuildCtx(const std::vector<Box>& b, int ml, std::vector<ctl_bvh_node>& n, std::vector<uint32_t>& o, std::vector<uint8_t>& l)
        : boxes(b), max_leaf(ml), nodes(n), ordered(o), last(l) {}
};

static void set_child_box(ctl_bvh_node& n, int which, const Box& b) {
    if (which == 0) { n.a[0] = b.lo.x; n.a[1] = b.hi.x; n.a[2] = b.lo.y; n.a[3] = b.hi.y; n.c[0] = b.lo.z; n.c[1] = b.hi.z; }
    else { n.b[0] = b.lo.x; n.b[1] = b.hi.x; n.b[2] = b.lo.y; n.b[3] = b.hi.y; n.c[2] = b.lo.z; n.c[3] = b.hi.z; }
}

static int emit_leaf(BuildCtx& c, uint32_t first, uint32_t count) {
    uint32_t slot = (uint32_t)c.ordered.size();
    for (uint32_t i = 0; i < count; i++) { c.ordered.push_back(c.prims[first + i]); c.last.push_back(i == count - 1); }
    return ~(int)slot;
}

// returns the child reference; bounds of the range in `bounds`
static int build_range(BuildCtx& c, uint32_t first, uint32_t count, const Box& bounds, uint32_t parent, bool is_root) {
    const int NB = 16;
    bool want_leaf = false;
    int best_axis = -1, best_bin = -1;
    float best_cost = 3.0e38f;
    Box cb;
    for (uint32_t i = 0; i < count; i++) cb.grow(c.centroids[c.prims[first + i]]);
    if (count > 1) {
        for (int axis = 0; axis < 3; axis++) {
            float lo = cb.lo[axis], ext = cb.hi[axis] - lo;
            if (!(ext > 0)) continue;
            Box bb[NB]; uint32_t bc[NB] = {0};
            float scale = NB / ext;
            for (uint32_t i = 0; i < count; i++) {
                uint32_t p = c.prims[first + i];
                int b = (int)((c.centroids[p][axis] - lo) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                bb[b].grow(c.boxes[p]); bc[b]++;
            }
            float ra[NB]; Box acc; uint32_t n = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); n += bc[b]; ra[b] = acc.area() * n; }
            acc = Box(); n = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]); n += bc[b];
                if (n == 0 || n == count) continue;
                float cost = acc.area() * n + ra[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
    }
    float area = bounds.area();
    float split_cost = (best_axis >= 0 && area > 0) ? 1.0f + best_cost / area : 3.0e38f;
    if ((int)count <= c.max_leaf && (float)count <= split_cost) want_leaf = true;
    if (count == 1) want_leaf = true;
    if (is_root && count >= 2) want_leaf = false; // a root always gets two real children; only 1-primitive BVHs use the sentinel form
    if (want_leaf && !is_root) return emit_leaf(c, first, count);

    uint32_t node_idx = (uint32_t)c.nodes.size();
    c.nodes.push_back(ctl_bvh_node());
    memset(&c.nodes[node_idx], 0, sizeof(ctl_bvh_node));
    if (want_leaf) { // root-is-leaf case (single primitive), SplitBVHBuilder.cpp:176-189: right child = sentinel, degenerate box at the origin
        int leaf = emit_leaf(c, first, count);
        ctl_bvh_node& n = c.nodes[node_idx];
        n.child0 = leaf; n.child1 = CTL_SENTINEL; n.parent = 0xffffffffu;
        set_child_box(n, 0, bounds);
        set_child_box(n, 1, Box(V3(0.0f), V3(0.0f)));
        return (int)(node_idx * 4);
    }
    uint32_t mid;
    if (best_axis >= 0) {
        float lo = cb.lo[best_axis], scale = NB / (cb.hi[best_axis] - lo);
        auto it = std::partition(c.prims.begin() + first, c.prims.begin() + first + count, [&](uint32_t p) {
            int b = (int)((c.centroids[p][best_axis] - lo) * scale);
            b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
            return b <= best_bin;
        });
        mid = (uint32_t)(it - c.prims.begin());
    } else mid = first + count / 2;
    if (mid == first || mid == first + count) mid = first + count / 2;
    Box lb, rb;
    for (uint32_t i = first; i < mid; i++) lb.grow(c.boxes[c.prims[i]]);
    for (uint32_t i = mid; i < first + count; i++) rb.grow(c.boxes[c.prims[i]]);
    int a = build_range(c, first, mid - first, lb, node_idx * 4, false);
    int b = build_range(c, mid, first + count - mid, rb, node_idx * 4, false);
    ctl_bvh_node& n = c.nodes[node_idx];
    n.child0 = a; n.child1 = b; n.parent = is_root ? 0xffffffffu : parent;
    set_child_box(n, 0, lb); set_child_box(n, 1, rb);
    return (int)(node_idx * 4);
}
} // namespace

void build_bvh(const std::vector<Box>& prim_boxes, int max_leaf, std::vector<ctl_bvh_node>& nodes_out,
               std::vector<uint32_t>& ordered, std::vector<uint8_t>& last) {
    nodes_out.clear(); ordered.clear(); last.clear();
    BuildCtx c(prim_boxes, max_leaf, nodes_out, ordered, last);
    c.centroids.resize(prim_boxes.size());
    c.prims.resize(prim_boxes.size());
    Box all;
    for (size_t i = 0; i < prim_boxes.size(); i++) { c.centroids[i] = prim_boxes[i].center(); c.prims[i] = (uint32_t)i; all.grow(prim_boxes[i]); }
    if (prim_boxes.empty()) return;
    build_range(c, 0, (uint32_t)prim_boxes.size(), all, 0xffffffffu, true);
}

// ------------------------------------------------------------------ camera

void make_camera(V3 pos, V3 target, V3 up, float fov_deg, int w, int h, ctl_camera* cam) {
    float aspect = (float)w / (float)h;
    float fov = fov_deg * (kPi / 180.0f);
    M4 c2s = M4::scale(V3(-0.5f, -0.5f * aspect, 1.0f)).mul(M4::translate(V3(-1.0f, -1.0f / aspect, 0.0f))).mul(M4::perspective(fov, 1.0f, 100000.0f));
    M4 s2c = c2s.inverse();
    M4 tw = M4::look_at(pos, target, up);
    memcpy(cam->sample_to_camera, s2c.m, 64);
    memcpy(cam->to_world, tw.m, 64);
    cam->resolution[0] = (float)w; cam->resolution[1] = (float)h;
    cam->inv_resolution[0] = 1.0f / (float)w; cam->inv_resolution[1] = 1.0f / (float)h;
}

// ------------------------------------------------------------------ assembly

void assemble_scene(const std::vector<MeshInput>& meshes, const std::vector<NodeInput>& nodes, V3 cam_pos, V3 cam_target,
                    V3 cam_up, float fov_deg, int width, int height, SceneStorage& S) {
    S = SceneStorage();
    std::vector<Box> mesh_box(meshes.size());
    // Meshes that need compiling (TriangleData, BVH, Woop records) are independent: they are built concurrently into per-mesh arrays and appended in
    // mesh order afterwards, so the scene arrays do not depend on the schedule.
    struct Built { std::vector<ctl_tri_data> tri_data; std::vector<float> verts9; std::vector<ctl_bvh_node> nodes; std::vector<ctl_woop_tri> woop; std::vector<uint32_t> index; Box box; std::exception_ptr err; };
    std::vector<Built> built(meshes.size());
    auto build_one = [&](size_t mi) {
        const MeshInput& M = meshes[mi]; Built& B = built[mi];
        const uint32_t nt = (uint32_t)M.indices.size() / 3;
        if (M.materials.size() > 255) throw std::runtime_error("more than 255 materials in one mesh (8-bit index, TriangleData.h:24)");
        std::vector<V3> vn;
        compute_vertex_normals(M.verts, M.indices, vn);
        std::vector<Box> pb(nt);
        B.verts9.reserve((size_t)nt * 9); B.tri_data.reserve(nt);
        for (uint32_t t = 0; t < nt; t++) {
            V3 p[3], n[3];
            float uv[6] = {0, 0, 0, 0, 0, 0};
            for (int k = 0; k < 3; k++) {
                const uint32_t vi = M.indices[t * 3 + k];
                p[k] = M.verts[vi]; n[k] = M.normals.empty() ? vn[vi] : normalize(M.normals[vi]); pb[t].grow(p[k]);
                if (!M.uvs.empty()) { uv[2 * k] = M.uvs[2 * vi]; uv[2 * k + 1] = M.uvs[2 * vi + 1]; }
                B.verts9.insert(B.verts9.end(), {p[k].x, p[k].y, p[k].z});
            }
            ctl_tri_data td;
            encode_tri_data(p, n, uv, M.mat_index[t], &td);
            B.tri_data.push_back(td);
            B.box.grow(pb[t]);
        }
        std::vector<uint32_t> ord; std::vector<uint8_t> last;
        const char* which = getenv("CTL_BVH_BUILDER");
        if (which && std::string(which) == "sah") build_bvh(pb, 8, B.nodes, ord, last);
        else build_sbvh(B.verts9.data(), nt, getenv("CTL_SBVH_MAXLEAF") ? atoi(getenv("CTL_SBVH_MAXLEAF")) : 8, B.nodes, ord, last); // maxLeafSize 8: BVHBuilderHelper.cpp:119 (env: experiments)
        B.woop.resize(ord.size()); B.index.resize(ord.size());
        for (size_t s = 0; s < ord.size(); s++) {
            const uint32_t t = ord[s];
            encode_woop(M.verts[M.indices[t * 3]], M.verts[M.indices[t * 3 + 1]], M.verts[M.indices[t * 3 + 2]], &B.woop[s]);
            B.index[s] = (t << 1) | (last[s] ? 1u : 0u);
        }
    };
    {
        std::vector<size_t> todo;
        for (size_t mi = 0; mi < meshes.size(); mi++) if (meshes[mi].pre_tri_data.empty()) todo.push_back(mi);
        std::atomic<size_t> next(0);
        auto worker = [&]() { for (size_t k; (k = next.fetch_add(1)) < todo.size();) { try { build_one(todo[k]); } catch (...) { built[todo[k]].err = std::current_exception(); } } };
        const size_t n_workers = std::min<size_t>(todo.size(), std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (size_t t = 1; t < n_workers; t++) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
        for (size_t mi : todo) if (built[mi].err) std::rethrow_exception(built[mi].err);
    }
    for (size_t mi = 0; mi < meshes.size(); mi++) {
        const MeshInput& M = meshes[mi];
        ctl_mesh km;
        km.tri_offset = (uint32_t)S.tri_data.size(); km.bvh_node_offset = (uint32_t)S.bvh_nodes.size() * 4; km.bvh_tri_offset = (uint32_t)S.woop.size() * 3;
        km.bvh_idx_offset = (uint32_t)S.tri_index.size(); km.mat_offset = (uint32_t)S.materials.size();
        for (auto m : M.materials) { m.node_light_index = 0xffffffffu; S.materials.push_back(m); }
        if (!M.pre_tri_data.empty()) { // pre-compiled mesh (.xmsh): reference-layout arrays appended as they are
            S.tri_data.insert(S.tri_data.end(), M.pre_tri_data.begin(), M.pre_tri_data.end());
            S.bvh_nodes.insert(S.bvh_nodes.end(), M.pre_nodes.begin(), M.pre_nodes.end());
            S.woop.insert(S.woop.end(), M.pre_woop.begin(), M.pre_woop.end());
            S.tri_index.insert(S.tri_index.end(), M.pre_index.begin(), M.pre_index.end());
            S.mesh_verts9.emplace_back(); // no source vertices: GPU BVH rebuilds are not available for imported meshes
            mesh_box[mi] = M.pre_box;
        } else {
            Built& B = built[mi];
            S.tri_data.insert(S.tri_data.end(), B.tri_data.begin(), B.tri_data.end());
            S.bvh_nodes.insert(S.bvh_nodes.end(), B.nodes.begin(), B.nodes.end());
            S.woop.insert(S.woop.end(), B.woop.begin(), B.woop.end());
            S.tri_index.insert(S.tri_index.end(), B.index.begin(), B.index.end());
            S.mesh_verts9.emplace_back(std::move(B.verts9));
            mesh_box[mi] = B.box;
            B = Built();
        }
        S.meshes.push_back(km);
    }
    S.mesh_boxes = mesh_box;
    S.node_inputs = nodes;
    for (const MeshInput& M : meshes) { S.mesh_n_materials.push_back((uint32_t)M.materials.size()); S.mesh_emissive.push_back(M.emissive); }
    S.cam_pos = cam_pos; S.cam_target = cam_target; S.cam_up = cam_up; S.cam_fov = fov_deg; S.cam_w = width; S.cam_h = height;
    assemble_nodes(S);
}

void rebraid(SceneStorage& S, const std::vector<Box>& node_box);

void assemble_nodes(SceneStorage& S) {
    const std::vector<NodeInput>& nodes = S.node_inputs;
    const std::vector<Box>& mesh_box = S.mesh_boxes;
    S.nodes.clear(); S.node_xf.clear(); S.node_inv_xf.clear(); S.scene_bvh.clear(); S.lights.clear(); S.light_tris.clear(); S.light_cdf_data.clear();
    S.box = Box();
    for (ctl_material& m : S.materials) m.node_light_index = 0xffffffffu;
    // nodes
    std::vector<Box> node_box(nodes.size());
    for (size_t ni = 0; ni < nodes.size(); ni++) {
        const NodeInput& N = nodes[ni];
        ctl_node kn;
        kn.mesh_index = N.mesh;
        kn.material_offset = N.material_override >= 0 ? (uint32_t)N.material_override : S.meshes[N.mesh].mat_offset;
        kn.instanciated_material = 0;
        kn.lights[0] = kn.lights[1] = 0xffffffffu; kn.n_lights = 0;
        S.nodes.push_back(kn);
        M4 inv = N.xf.inverse();
        S.node_xf.insert(S.node_xf.end(), N.xf.m, N.xf.m + 16);
        S.node_inv_xf.insert(S.node_inv_xf.end(), inv.m, inv.m + 16);
        const Box& lb = mesh_box[N.mesh];
        for (int c = 0; c < 8; c++)
            node_box[ni].grow(N.xf.transform_point(V3(c & 1 ? lb.hi.x : lb.lo.x, c & 2 ? lb.hi.y : lb.lo.y, c & 4 ? lb.hi.z : lb.lo.z)));
        S.box.grow(node_box[ni]);
    }
    // scene-level BVH: one node per leaf, leaf = ~nodeIdx (BVHRebuilder.cpp:172-177)
    {
        std::vector<uint32_t> ord; std::vector<uint8_t> last;
        build_bvh(node_box, 1, S.scene_bvh, ord, last);
        for (auto& n : S.scene_bvh) {
            if (n.child0 < 0) n.child0 = ~(int)ord[~n.child0];
            if (n.child1 < 0) n.child1 = ~(int)ord[~n.child1];
        }
        S.scene_start = 0;
        if (nodes.size() == 1) S.scene_start = ~0; // single node: start at the leaf (TraceHelper.cu:417, BVHTraversal.h:11-12)
    }
    // area lights: one DiffuseLight per (node, emissive material), DynamicScene.cpp:689-711
    std::vector<std::vector<uint32_t>> first_slot(S.meshes.size());   // per mesh with emitters: first leaf slot referencing each triangle (built on demand, one pass over the leaf words)
    auto slot_table = [&](uint32_t mesh_id) -> const std::vector<uint32_t>& {
        std::vector<uint32_t>& tab = first_slot[mesh_id];
        if (!tab.empty()) return tab;
        const ctl_mesh& km = S.meshes[mesh_id];
        const uint32_t nt = (mesh_id + 1 < S.meshes.size() ? S.meshes[mesh_id + 1].tri_offset : (uint32_t)S.tri_data.size()) - km.tri_offset;
        const uint32_t slot_end = (mesh_id + 1 < S.meshes.size()) ? S.meshes[mesh_id + 1].bvh_idx_offset : (uint32_t)S.tri_index.size();
        tab.assign((size_t)nt + 1, 0xffffffffu);
        for (uint32_t slot = km.bvh_idx_offset; slot < slot_end; slot++) { const uint32_t t = S.tri_index[slot] >> 1; if (t < nt && tab[t] == 0xffffffffu) tab[t] = slot; }
        return tab;
    };
    for (size_t ni = 0; ni < nodes.size(); ni++) {
        const uint32_t mesh_id = nodes[ni].mesh;
        const ctl_mesh& km = S.meshes[mesh_id];
        const std::vector<V3>& emissive = S.mesh_emissive[mesh_id];
        for (size_t m = 0; m < S.mesh_n_materials[mesh_id]; m++) {
            V3 L = m < emissive.size() ? emissive[m] : V3(0.0f);
            if (L.x == 0 && L.y == 0 && L.z == 0) continue;
            ctl_node& kn = S.nodes[ni];
            if (kn.n_lights >= 2) throw std::runtime_error("Node already has maximum number of area lights!");
            if (S.lights.size() >= CTL_MAX_NUM_LIGHTS) throw std::runtime_error("too many lights");
            ctl_light lt;
            lt.radiance[0] = L.x; lt.radiance[1] = L.y; lt.radiance[2] = L.z;
            lt.tri_offset = (uint32_t)S.light_tris.size();
            lt.cdf_offset = (uint32_t)S.light_cdf_data.size();
            lt.node_idx = (uint32_t)ni;
            lt.count = 0;
            const uint32_t nt = (mesh_id + 1 < S.meshes.size() ? S.meshes[mesh_id + 1].tri_offset : (uint32_t)S.tri_data.size()) - km.tri_offset;
            M4 xf = nodes[ni].xf;
            for (uint32_t t = 0; t < nt; t++) {
                const uint32_t tm = (S.tri_data[km.tri_offset + t].w[1] >> 16) & 0xffu; // TriangleData::getMatIndex, TriangleData.h:40-44
                if (tm != m) continue;
                ctl_light_tri lt3; memset(&lt3, 0, sizeof(lt3));
                // first woop slot referencing this triangle
                const uint32_t slot = slot_table(mesh_id)[t];
                if (slot == 0xffffffffu || slot >= S.woop.size()) throw std::runtime_error("emissive triangle " + std::to_string(t) + " of mesh " + std::to_string(mesh_id) + " is not referenced by the BVH");
                lt3.i_dat = slot; lt3.t_dat = km.tri_offset + t;
                V3 p0, p1, p2;
                decode_woop(S.woop[slot], p0, p1, p2);
                V3 n = shading_normal_at(S.tri_data[lt3.t_dat], xf, 1.0f / 3.0f, 1.0f / 3.0f);
                p0 = xf.transform_point(p0); p1 = xf.transform_point(p1); p2 = xf.transform_point(p2);
                V3 pp[3] = {p0, p1, p2};
                for (int k = 0; k < 3; k++) { lt3.p[k][0] = pp[k].x; lt3.p[k][1] = pp[k].y; lt3.p[k][2] = pp[k].z; }
                lt3.n[0] = n.x; lt3.n[1] = n.y; lt3.n[2] = n.z;
                lt3.area = 0.5f * length(cross(p2 - p0, p1 - p0));
                S.light_tris.push_back(lt3);
                lt.count++;
            }
            if (lt.count == 0) continue;
            float sum = 0; std::vector<float> cdf(lt.count + 1); cdf[0] = 0.0f;
            for (uint32_t i = 0; i < lt.count; i++) { float a = S.light_tris[lt.tri_offset + i].area; sum += a; cdf[i + 1] = cdf[i] + a; }
            for (auto& c : cdf) c = c / sum;
            lt.sum_area = sum;
            S.light_cdf_data.insert(S.light_cdf_data.end(), cdf.begin(), cdf.end());
            uint32_t light_idx = (uint32_t)S.lights.size();
            S.lights.push_back(lt);
            S.materials[kn.material_offset + m].node_light_index = kn.n_lights;
            kn.lights[kn.n_lights++] = light_idx;
        }
    }
    // light selection CDF, all weights 1 (DynamicScene.cpp:173-196)
    S.num_lights = (uint32_t)std::min<size_t>(CTL_MAX_NUM_LIGHTS, S.lights.size());
    for (int i = 0; i < CTL_MAX_NUM_LIGHTS; i++) { S.light_indices[i] = 0; S.light_cdf[i] = 0; }
    float accum = 0; for (uint32_t i = 0; i < S.lights.size(); i++) accum += 1.0f;
    for (uint32_t i = 0; i < S.num_lights; i++) {
        S.light_indices[i] = i;
        float pdf = 1.0f / accum;
        S.light_cdf[i] = (i > 0 ? S.light_cdf[i - 1] : 0.0f) + pdf;
    }
    make_camera(S.cam_pos, S.cam_target, S.cam_up, S.cam_fov, S.cam_w, S.cam_h, &S.camera);
    S.ray_eps = 1e-4f * length(S.box.hi - S.box.lo); // DynamicScene.cpp:587
    rebraid(S, node_box);
    // the traversal kernels keep 64 stack entries per ray (BVHTraversal.h: traversalStack[64]): trees built here are checked against that, a re-braided
    // level that does not fit is dropped, a plain view that does not fit is refused
    ctl_scene_view v; S.fill_view(&v);
    if (view_stack_depth(v) > 64) {
        if (S.rb_active) { S.rb_active = false; S.fill_view(&v); }
        if (view_stack_depth(v) > 64) throw std::runtime_error("scene trees are deeper than the 64-entry traversal stack (scene level + mesh level)");
    }
}

// Partial re-braiding (Benthin, Woop, Wald, Afra: "Improved two-level BVHs using partial re-braiding", HPG 2017) inside the reference's data
// layout.  The reference's scene level has one leaf per instance (BVHRebuilder); where instance boxes overlap, a ray descends every overlapping
// mesh tree from its root (config 4: 5.7 instance entries and 95 inner nodes per ray).  Here the instances with the largest world-space boxes are
// opened, top-down, into entries (instance, sub-tree) until the budget of scene-level leaves is reached, and the scene-level tree is built over the
// entries.  No kernel knows: an entry is an ordinary node record (a copy of the instance's: same transform, materials, lights) whose mesh record
// points at a re-based copy of the sub-tree (same Woop / index / TriangleData offsets).  Hits are the same triangles at the same distances; the
// node index a hit reports is the pseudo-node's, node_alias gives the instance.
void rebraid(SceneStorage& S, const std::vector<Box>& node_box) {
    S.rb_active = false;
    S.rb_bvh_nodes.clear(); S.rb_scene_bvh.clear(); S.rb_meshes.clear(); S.rb_nodes.clear(); S.rb_node_xf.clear(); S.rb_node_inv_xf.clear(); S.rb_node_alias.clear();
    size_t budget = S.rebraid_entries;
    if (S.rebraid_entries == kRebraidAuto) {
        // default: large multi-instance scenes get 1 024 scene-level leaves (measured on the 1 M-triangle configs, profiles/r02a_rebraid_summary.log: +11 % / +22 %
        // Mrays/s at 1 024, nothing more at 4 096); small scenes keep the reference's one leaf per instance.  CTL_REBRAID=<n> overrides (0 = off).
        if (const char* e = getenv("CTL_REBRAID")) budget = (size_t)atoll(e);
        else budget = S.bvh_nodes.size() >= 32768 ? 1024 : 0;
    }
    const size_t n_real = S.nodes.size();
    if (budget <= n_real || n_real < 2) return;
    struct Entry { uint32_t node; int ref; Box world; float area; uint32_t out_node; };
    auto mesh_node = [&](uint32_t ni, int ref) -> const ctl_bvh_node& { return S.bvh_nodes[S.meshes[S.nodes[ni].mesh_index].bvh_node_offset / 4 + (uint32_t)ref / 4]; };
    auto inner = [](int c) { return c >= 0 && c != CTL_SENTINEL; };
    auto openable = [&](const Entry& e) { const ctl_bvh_node& n = mesh_node(e.node, e.ref); return inner(n.child0) && inner(n.child1); };
    auto world_of = [&](uint32_t ni, const Box& lb) {
        Box b; const M4& xf = S.node_inputs[ni].xf;
        for (int c = 0; c < 8; c++) b.grow(xf.transform_point(V3(c & 1 ? lb.hi.x : lb.lo.x, c & 2 ? lb.hi.y : lb.lo.y, c & 4 ? lb.hi.z : lb.lo.z)));
        return b;
    };
    auto cmp = [](const Entry& a, const Entry& b) { return a.area != b.area ? a.area < b.area : (a.node != b.node ? a.node > b.node : a.ref > b.ref); }; // max-heap on area, ties by index
    std::vector<Entry> heap, done;
    // which entry to open next: the one whose box shrinks most when replaced by its two children (area - mean child area).  CTL_REBRAID_MODE=0 selects the
    // plain "largest box first" of the paper for A/B; on config 4 at 1 024 entries the oracle counts 5 468 algorithmic bytes per ray against 5 682.
    const int mode = getenv("CTL_REBRAID_MODE") ? atoi(getenv("CTL_REBRAID_MODE")) : 1;
    auto child_boxes = [&](const Entry& e, Box& b0, Box& b1) {
        const ctl_bvh_node& n = mesh_node(e.node, e.ref);
        b0 = world_of(e.node, Box(V3(n.a[0], n.a[2], n.c[0]), V3(n.a[1], n.a[3], n.c[1]))); b1 = world_of(e.node, Box(V3(n.b[0], n.b[2], n.c[2]), V3(n.b[1], n.b[3], n.c[3])));
    };
    auto place = [&](Entry e) {
        e.area = e.world.area();
        if (!openable(e)) { done.push_back(e); return; }
        if (mode == 1) { Box b0, b1; child_boxes(e, b0, b1); e.area = e.world.area() - 0.5f * (b0.area() + b1.area()); }
        if (!(e.area > -3.0e38f && e.area < 3.0e38f)) e.area = 0.0f;   // non-finite boxes must not poison the heap order
        heap.push_back(e); std::push_heap(heap.begin(), heap.end(), cmp);
    };
    for (uint32_t ni = 0; ni < n_real; ni++) place(Entry{ni, 0, node_box[ni], 0.0f, ni});
    while (!heap.empty() && heap.size() + done.size() < budget) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        const Entry e = heap.back(); heap.pop_back();
        const ctl_bvh_node& n = mesh_node(e.node, e.ref);
        Box b0, b1; child_boxes(e, b0, b1);
        place(Entry{e.node, n.child0, b0, 0.0f, 0});
        place(Entry{e.node, n.child1, b1, 0.0f, 0});
    }
    done.insert(done.end(), heap.begin(), heap.end());
    std::sort(done.begin(), done.end(), [](const Entry& a, const Entry& b) { return a.node != b.node ? a.node < b.node : a.ref < b.ref; });
    bool any_opened = false;
    for (const Entry& e : done) any_opened |= e.ref != 0;
    if (!any_opened) return;
    S.rb_bvh_nodes = S.bvh_nodes; S.rb_meshes = S.meshes; S.rb_nodes = S.nodes; S.rb_node_xf = S.node_xf; S.rb_node_inv_xf = S.node_inv_xf;
    S.rb_node_alias.resize(n_real); for (uint32_t i = 0; i < n_real; i++) S.rb_node_alias[i] = i;
    std::vector<Box> entry_box;
    for (Entry& e : done) {
        entry_box.push_back(e.world);
        if (e.ref == 0) { e.out_node = e.node; continue; }   // unopened instance: its own node record
        // re-based copy of the sub-tree, pre-order; the copy's node 0 is the sub-tree root (traversal enters a mesh at its node 0)
        const uint32_t base = S.meshes[S.nodes[e.node].mesh_index].bvh_node_offset / 4, first = (uint32_t)S.rb_bvh_nodes.size();
        std::vector<std::pair<uint32_t, uint32_t>> todo;   // (old local index, new local index)
        S.rb_bvh_nodes.push_back(S.bvh_nodes[base + (uint32_t)e.ref / 4]); S.rb_bvh_nodes.back().parent = 0xffffffffu;
        todo.emplace_back((uint32_t)e.ref / 4, 0u);
        while (!todo.empty()) {
            const auto cur = todo.back(); todo.pop_back();
            int ch[2] = {S.bvh_nodes[base + cur.first].child0, S.bvh_nodes[base + cur.first].child1};
            for (int k = 1; k >= 0; k--) {
                if (!inner(ch[k])) continue;
                const uint32_t nu = (uint32_t)S.rb_bvh_nodes.size() - first;
                S.rb_bvh_nodes.push_back(S.bvh_nodes[base + (uint32_t)ch[k] / 4]); S.rb_bvh_nodes.back().parent = cur.second * 4;
                todo.emplace_back((uint32_t)ch[k] / 4, nu);
                ch[k] = (int)(nu * 4);
            }
            S.rb_bvh_nodes[first + cur.second].child0 = ch[0]; S.rb_bvh_nodes[first + cur.second].child1 = ch[1];
        }
        ctl_mesh pm = S.meshes[S.nodes[e.node].mesh_index]; pm.bvh_node_offset = first * 4;
        S.rb_meshes.push_back(pm);
        ctl_node pn = S.nodes[e.node]; pn.mesh_index = (uint32_t)S.rb_meshes.size() - 1;
        e.out_node = (uint32_t)S.rb_nodes.size();
        S.rb_nodes.push_back(pn);
        S.rb_node_xf.insert(S.rb_node_xf.end(), S.node_xf.begin() + 16 * e.node, S.node_xf.begin() + 16 * e.node + 16);
        S.rb_node_inv_xf.insert(S.rb_node_inv_xf.end(), S.node_inv_xf.begin() + 16 * e.node, S.node_inv_xf.begin() + 16 * e.node + 16);
        S.rb_node_alias.push_back(e.node);
    }
    std::vector<uint32_t> ord; std::vector<uint8_t> last;
    build_bvh(entry_box, 1, S.rb_scene_bvh, ord, last);
    for (auto& n : S.rb_scene_bvh) {
        if (n.child0 < 0) n.child0 = ~(int)done[ord[~n.child0]].out_node;
        if (n.child1 < 0) n.child1 = ~(int)done[ord[~n.child1]].out_node;
    }
    // (the mesh trees' post-pass, optimize_bvh, was tried on this tree: -0.4 % algorithmic bytes at 1 024 entries, +7 % at 16: left out)
    S.rb_active = true;
}

void SceneStorage::fill_view(ctl_scene_view* v) const {
    memset(v, 0, sizeof(*v));
    const bool rb = rb_active;
    v->bvh_nodes = rb ? rb_bvh_nodes.data() : bvh_nodes.data(); v->n_bvh_nodes = (uint32_t)(rb ? rb_bvh_nodes.size() : bvh_nodes.size());
    v->woop = woop.data(); v->n_woop = (uint32_t)woop.size();
    v->tri_index = tri_index.data(); v->n_tri_index = (uint32_t)tri_index.size();
    v->tri_data = tri_data.data(); v->n_tri_data = (uint32_t)tri_data.size();
    v->meshes = rb ? rb_meshes.data() : meshes.data(); v->n_meshes = (uint32_t)(rb ? rb_meshes.size() : meshes.size());
    v->nodes = rb ? rb_nodes.data() : nodes.data(); v->n_nodes = (uint32_t)(rb ? rb_nodes.size() : nodes.size());
    v->node_xf = rb ? rb_node_xf.data() : node_xf.data(); v->node_inv_xf = rb ? rb_node_inv_xf.data() : node_inv_xf.data();
    v->scene_bvh_nodes = rb ? rb_scene_bvh.data() : scene_bvh.data(); v->n_scene_bvh_nodes = (uint32_t)(rb ? rb_scene_bvh.size() : scene_bvh.size());
    v->scene_start_node = rb ? 0 : scene_start;
    v->node_alias = rb ? rb_node_alias.data() : nullptr;
    v->materials = materials.data(); v->n_materials = (uint32_t)materials.size();
    v->lights = lights.data(); v->n_lights_buf = (uint32_t)lights.size();
    v->light_tris = light_tris.data(); v->n_light_tris = (uint32_t)light_tris.size();
    v->light_cdf_data = light_cdf_data.data(); v->n_light_cdf_data = (uint32_t)light_cdf_data.size();
    v->num_lights = num_lights;
    memcpy(v->light_indices, light_indices, sizeof(light_indices));
    memcpy(v->light_cdf, light_cdf, sizeof(light_cdf));
    v->camera = camera;
    v->box_min[0] = box.lo.x; v->box_min[1] = box.lo.y; v->box_min[2] = box.lo.z;
    v->box_max[0] = box.hi.x; v->box_max[1] = box.hi.y; v->box_max[2] = box.hi.z;
    v->ray_eps = ray_eps;
}

// ------------------------------------------------------------------ synthetic scenes

namespace {
struct Rng { // splitmix-style 64-bit generator; scene generation only
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    float uni() { return (float)((next() >> 40) * (1.0 / 16777216.0)); }
    float uni(float a, float b) { return a + (b - a) * uni(); }
    float gauss() { float u1 = fmaxf(uni(), 1e-7f), u2 = uni(); return sqrtf(-2.0f * logf(u1)) * cosf(2.0f * kPi * u2); }
};

static ctl_material mat_diffuse(float r, float g, float b) {
    ctl_material m; memset(&m, 0, sizeof(m));
    m.bsdf_type = CTL_BSDF_DIFFUSE; m.flags = CTL_MAT_TWO_SIDED; m.node_light_index = 0xffffffffu;
    m.reflectance[0] = r; m.reflectance[1] = g; m.reflectance[2] = b;
    return m;
}
static ctl_material mat_conductor(int distr, float alpha) { // Cu eta/k, SURVEY §8d C3
    ctl_material m; memset(&m, 0, sizeof(m));
    m.bsdf_type = CTL_BSDF_ROUGHCONDUCTOR; m.flags = CTL_MAT_TWO_SIDED; m.node_light_index = 0xffffffffu; m.distr_type = distr;
    m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = 1.0f;
    m.alpha_u = m.alpha_v = alpha;
    m.eta[0] = 0.200f; m.eta[1] = 0.924f; m.eta[2] = 1.102f;
    m.k[0] = 3.912f; m.k[1] = 2.452f; m.k[2] = 2.142f;
    return m;
}
static ctl_material mat_dielectric(float eta) {
    ctl_material m; memset(&m, 0, sizeof(m));
    m.bsdf_type = CTL_BSDF_DIELECTRIC; m.flags = 0; m.node_light_index = 0xffffffffu;
    m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = 1.0f;
    m.eta[0] = eta; m.transmittance = 1.0f;
    return m;
}

// Triangles are wound so that the reference's vertex-normal rule (Mesh.cpp:181-184,
// n ~ cross(v3-v1, v2-v1)) yields the normal `outward`.
static void add_tri(MeshInput& M, V3 a, V3 b, V3 c, V3 outward, uint8_t mat, bool shared = false, uint32_t ia = 0, uint32_t ib = 0, uint32_t ic = 0) {
    bool flip = dot(cross(c - a, b - a), outward) < 0;
    if (!shared) {
        uint32_t base = (uint32_t)M.verts.size();
        M.verts.push_back(a); M.verts.push_back(b); M.verts.push_back(c);
        ia = base; ib = base + 1; ic = base + 2;
    }
    if (flip) std::swap(ib, ic);
    M.indices.push_back(ia); M.indices.push_back(ib); M.indices.push_back(ic);
    M.mat_index.push_back(mat);
}
static void add_quad(MeshInput& M, V3 a, V3 b, V3 c, V3 d, V3 outward, uint8_t mat) {
    add_tri(M, a, b, c, outward, mat);
    add_tri(M, a, c, d, outward, mat);
}
// axis-aligned unit cube [0,1]^3, 5 faces (no bottom) = 10 triangles, outward normals
static void add_open_cube(MeshInput& M, const M4& xf, uint8_t mat) {
    V3 p[8];
    for (int i = 0; i < 8; i++) p[i] = xf.transform_point(V3((float)(i & 1), (float)((i >> 1) & 1), (float)((i >> 2) & 1)));
    V3 ctr = xf.transform_point(V3(0.5f, 0.5f, 0.5f));
    auto face = [&](int a, int b, int c, int d) {
        V3 fc = (p[a] + p[b] + p[c] + p[d]) * 0.25f;
        add_quad(M, p[a], p[b], p[c], p[d], fc - ctr, mat);
    };
    face(2, 3, 7, 6); // top (y=1)
    face(0, 1, 3, 2); // z=0
    face(4, 5, 7, 6); // z=1
    face(0, 2, 6, 4); // x=0
    face(1, 3, 7, 5); // x=1
}
static void add_icosphere(MeshInput& M, V3 center, float radius, int subdiv, uint8_t mat) {
    const float t = (1.0f + sqrtf(5.0f)) / 2.0f;
    std::vector<V3> v = {V3(-1, t, 0), V3(1, t, 0), V3(-1, -t, 0), V3(1, -t, 0), V3(0, -1, t), V3(0, 1, t),
                         V3(0, -1, -t), V3(0, 1, -t), V3(t, 0, -1), V3(t, 0, 1), V3(-t, 0, -1), V3(-t, 0, 1)};
    for (auto& p : v) p = normalize(p);
    std::vector<uint32_t> f = {0, 11, 5, 0, 5, 1, 0, 1, 7, 0, 7, 10, 0, 10, 11, 1, 5, 9, 5, 11, 4, 11, 10, 2, 10, 7, 6, 7, 1, 8,
                               3, 9, 4, 3, 4, 2, 3, 2, 6, 3, 6, 8, 3, 8, 9, 4, 9, 5, 2, 4, 11, 6, 2, 10, 8, 6, 7, 9, 8, 1};
    for (int s = 0; s < subdiv; s++) {
        std::vector<uint32_t> nf;
        std::vector<std::pair<uint64_t, uint32_t>> cache;
        auto midpoint = [&](uint32_t a, uint32_t b) {
            uint64_t key = a < b ? ((uint64_t)a << 32 | b) : ((uint64_t)b << 32 | a);
            for (auto& e : cache) if (e.first == key) return e.second;
            v.push_back(normalize((v[a] + v[b]) * 0.5f));
            cache.push_back({key, (uint32_t)v.size() - 1});
            return (uint32_t)v.size() - 1;
        };
        // hash lookups dominate for subdiv 3 otherwise; sort-free linear cache is fine for <= 2k edges
        for (size_t i = 0; i < f.size(); i += 3) {
            uint32_t a = f[i], b = f[i + 1], c = f[i + 2];
            uint32_t ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            uint32_t add[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
            nf.insert(nf.end(), add, add + 12);
        }
        f.swap(nf);
    }
    uint32_t base = (uint32_t)M.verts.size();
    for (auto& p : v) M.verts.push_back(center + p * radius);
    for (size_t i = 0; i < f.size(); i += 3) {
        V3 a = M.verts[base + f[i]], b = M.verts[base + f[i + 1]], c = M.verts[base + f[i + 2]];
        add_tri(M, a, b, c, (a + b + c) * (1.0f / 3.0f) - center, mat, true, base + f[i], base + f[i + 1], base + f[i + 2]);
    }
}

// Cornell box data (classic measurements / 555 so the box is ~1 unit)
static void cornell(int kind, int w, int h, SceneStorage& S) {
    const float k = 1.0f / 555.0f;
    ctl_material white = mat_diffuse(0.725f, 0.71f, 0.68f), red = mat_diffuse(0.63f, 0.065f, 0.05f), green = mat_diffuse(0.14f, 0.45f, 0.091f);
    ctl_material lightm = mat_diffuse(0.78f, 0.78f, 0.78f);
    V3 Le(17.0f, 12.0f, 4.0f);
    V3 ctr(0.5f, 0.5f, 0.5f);
    auto quad = [&](MeshInput& M, V3 a, V3 b, V3 c, V3 d, uint8_t mat, V3 toward) { add_quad(M, a, b, c, d, toward, mat); };
    float X = 1.0f, Y = 548.8f * k, Z = 559.2f * k;
    V3 f00(0, 0, 0), f10(X, 0, 0), f11(X, 0, Z), f01(0, 0, Z), c00(0, Y, 0), c10(X, Y, 0), c11(X, Y, Z), c01(0, Y, Z);
    // light quad just below the ceiling, emitting downwards
    V3 l0(213 * k, Y - 0.001f, 227 * k), l1(343 * k, Y - 0.001f, 227 * k), l2(343 * k, Y - 0.001f, 332 * k), l3(213 * k, Y - 0.001f, 332 * k);
    M4 short_xf = M4::translate(V3(130 * k, 0, 65 * k)).mul(M4::rotate_y(-0.29f)).mul(M4::scale(V3(165 * k, 165 * k, 165 * k)));
    M4 tall_xf = M4::translate(V3(265 * k, 0, 296 * k)).mul(M4::rotate_y(0.30f)).mul(M4::scale(V3(165 * k, 330 * k, 165 * k)));
    std::vector<MeshInput> meshes; std::vector<NodeInput> nodes;
    if (kind == 0) {
        MeshInput M;
        M.materials = {white, red, green, lightm};
        M.emissive = {V3(0.0f), V3(0.0f), V3(0.0f), Le};
        quad(M, f00, f10, f11, f01, 0, V3(0, 1, 0));
        quad(M, c00, c10, c11, c01, 0, V3(0, -1, 0));
        quad(M, f01, f11, c11, c01, 0, V3(0, 0, -1));
        quad(M, f10, f11, c11, c10, 1, V3(-1, 0, 0)); // x = 1 wall (left as seen from the camera looking +z with x to the left)
        quad(M, f00, f01, c01, c00, 2, V3(1, 0, 0));
        add_open_cube(M, short_xf, 0);
        add_open_cube(M, tall_xf, 0);
        quad(M, l0, l1, l2, l3, 3, V3(0, -1, 0));
        meshes.push_back(M);
        nodes.push_back({0, M4::identity(), -1});
    } else {
        // 7 nodes, 6 meshes: floor | ceiling+back | left | right | unit cube (instanced twice) | light
        MeshInput m_floor, m_cb, m_left, m_right, m_cube, m_light;
        m_floor.materials = {white}; quad(m_floor, f00, f10, f11, f01, 0, V3(0, 1, 0));
        m_cb.materials = {white}; quad(m_cb, c00, c10, c11, c01, 0, V3(0, -1, 0)); quad(m_cb, f01, f11, c11, c01, 0, V3(0, 0, -1));
        m_left.materials = {red}; quad(m_left, f10, f11, c11, c10, 0, V3(-1, 0, 0));
        m_right.materials = {green}; quad(m_right, f00, f01, c01, c00, 0, V3(1, 0, 0));
        m_cube.materials = {white}; add_open_cube(m_cube, M4::identity(), 0);
        m_light.materials = {lightm}; m_light.emissive = {Le}; quad(m_light, l0, l1, l2, l3, 0, V3(0, -1, 0));
        meshes = {m_floor, m_cb, m_left, m_right, m_cube, m_light};
        nodes = {{0, M4::identity(), -1}, {1, M4::identity(), -1}, {2, M4::identity(), -1}, {3, M4::identity(), -1},
                 {4, short_xf, -1}, {4, tall_xf, -1}, {5, M4::identity(), -1}};
    }
    (void)ctr;
    assemble_scene(meshes, nodes, V3(278 * k, 273 * k, -800 * k), V3(278 * k, 273 * k, 0), V3(0, 1, 0), 39.3f, w, h, S);
}

static ctl_material c3_material(int i, Rng& rng) {
    switch (i % 6) {
    case 0: return mat_diffuse(rng.uni(0.2f, 0.9f), rng.uni(0.2f, 0.9f), rng.uni(0.2f, 0.9f));
    case 1: return mat_conductor(CTL_DISTR_BECKMANN, 0.05f);
    case 2: return mat_conductor(CTL_DISTR_BECKMANN, 0.1f);
    case 3: return mat_conductor(CTL_DISTR_BECKMANN, 0.3f);
    case 4: return mat_conductor(CTL_DISTR_GGX, 0.2f);
    default: return mat_dielectric(1.5f);
    }
}

static void add_room(MeshInput& M, float R, uint8_t wall_mat, uint8_t light_mat, float light_half) {
    V3 p[8];
    for (int i = 0; i < 8; i++) p[i] = V3(i & 1 ? R : -R, i & 2 ? R : -R, i & 4 ? R : -R);
    auto face = [&](int a, int b, int c, int d) { V3 fc = (p[a] + p[b] + p[c] + p[d]) * 0.25f; add_quad(M, p[a], p[b], p[c], p[d], -fc, wall_mat); };
    face(0, 1, 3, 2); face(4, 5, 7, 6); face(0, 2, 6, 4); face(1, 3, 7, 5); face(0, 1, 5, 4); face(2, 3, 7, 6);
    float y = R - 0.01f, s = light_half;
    add_quad(M, V3(-s, y, -s), V3(s, y, -s), V3(s, y, s), V3(-s, y, s), V3(0, -1, 0), light_mat);
}

// C2 / C3: room + 78 icospheres (99 854 triangles), one mesh / one node
static void scene_100k(bool microfacet, int w, int h, uint32_t seed, SceneStorage& S) {
    Rng rng(seed);
    MeshInput M;
    M.materials.push_back(mat_diffuse(0.7f, 0.7f, 0.7f)); M.emissive.push_back(V3(0.0f));
    M.materials.push_back(mat_diffuse(0.78f, 0.78f, 0.78f)); M.emissive.push_back(V3(25.0f));
    add_room(M, 10.0f, 0, 1, 2.0f);
    for (int i = 0; i < 78; i++) {
        V3 c(rng.uni(-8, 8), rng.uni(-8, 8), rng.uni(-8, 8));
        float r = rng.uni(0.3f, 1.2f);
        ctl_material m = mat_diffuse(rng.uni(0.2f, 0.9f), rng.uni(0.2f, 0.9f), rng.uni(0.2f, 0.9f));
        if (microfacet) m = c3_material(i, rng);
        M.materials.push_back(m); M.emissive.push_back(V3(0.0f));
        add_icosphere(M, c, r, 3, (uint8_t)(M.materials.size() - 1));
    }
    std::vector<MeshInput> meshes = {M};
    std::vector<NodeInput> nodes = {{0, M4::identity(), -1}};
    assemble_scene(meshes, nodes, V3(0, 0, -9.5f), V3(0, 0, 0), V3(0, 1, 0), 60.0f, w, h, S);
}

// C4 / C5: 703 clustered icospheres + 100 000 thin foliage triangles + room/light = 999 854 triangles, 8 meshes / 8 nodes
static void scene_1m(bool microfacet, int w, int h, uint32_t seed, int n_spheres, int n_foliage, SceneStorage& S) {
    Rng rng(seed);
    const int NM = 8;
    std::vector<MeshInput> meshes(NM);
    std::vector<M4> xf(NM), inv(NM);
    for (int k = 0; k < NM; k++) {
        xf[k] = k == NM - 1 ? M4::identity() : M4::translate(V3(0.3f * (k - 3), 0.1f * k, 0)).mul(M4::rotate_y(0.3f * k));
        inv[k] = xf[k].inverse();
    }
    std::vector<ctl_material> palette;
    for (int i = 0; i < 64; i++) palette.push_back(microfacet ? c3_material(i, rng) : mat_diffuse(rng.uni(0.2f, 0.9f), rng.uni(0.2f, 0.9f), rng.uni(0.2f, 0.9f)));
    for (int k = 0; k < NM; k++) { meshes[k].materials = palette; meshes[k].emissive.assign(64, V3(0.0f)); }
    V3 cl[16];
    for (int i = 0; i < 16; i++) cl[i] = V3(rng.uni(-6, 6), rng.uni(-6, 6), rng.uni(-6, 6));
    for (int i = 0; i < n_spheres; i++) {
        V3 c = cl[i % 16] + V3(rng.gauss(), rng.gauss(), rng.gauss()) * 1.5f;
        c = vmax(V3(-9.0f), vmin(V3(9.0f), c));
        float r = rng.uni(0.15f, 0.6f);
        int k = i % (NM - 1);
        add_icosphere(meshes[k], inv[k].transform_point(c), r, 3, (uint8_t)(i % 64));
    }
    MeshInput& F = meshes[NM - 1];
    F.materials.push_back(mat_diffuse(0.7f, 0.7f, 0.7f)); F.emissive.push_back(V3(0.0f));
    F.materials.push_back(mat_diffuse(0.78f, 0.78f, 0.78f)); F.emissive.push_back(V3(25.0f));
    add_room(F, 10.0f, 64, 65, 2.0f);
    for (int i = 0; i < n_foliage; i++) { // long thin triangles, aspect 50:1, random orientation
        V3 c(rng.uni(-9, 9), rng.uni(-9, 9), rng.uni(-9, 9));
        V3 d = normalize(V3(rng.gauss(), rng.gauss(), rng.gauss()));
        V3 s, t; coordinate_system(d, s, t);
        float L = rng.uni(0.5f, 1.5f), W = L / 50.0f;
        V3 a = c - d * (0.5f * L), b = c + d * (0.5f * L), e = c + s * W;
        add_tri(F, a, b, e, t, (uint8_t)(i % 64));
    }
    std::vector<NodeInput> nodes;
    for (int k = 0; k < NM; k++) nodes.push_back({(uint32_t)k, xf[k], -1});
    assemble_scene(meshes, nodes, V3(0, 0, -9.5f), V3(0, 0, 0), V3(0, 1, 0), 60.0f, w, h, S);
}

// small random scene for tests: n triangles of mixed materials in a lit room, 3 meshes / 4 nodes (one instanced)
static void scene_soup(int w, int h, uint32_t seed, int n, SceneStorage& S) {
    Rng rng(seed);
    MeshInput room, soup, ball;
    room.materials = {mat_diffuse(0.7f, 0.7f, 0.7f), mat_diffuse(0.78f, 0.78f, 0.78f)};
    room.emissive = {V3(0.0f), V3(10.0f, 9.0f, 8.0f)};
    add_room(room, 4.0f, 0, 1, 1.0f);
    for (int i = 0; i < 6; i++) { soup.materials.push_back(c3_material(i, rng)); soup.emissive.push_back(V3(0.0f)); }
    for (int i = 0; i < n; i++) {
        V3 c(rng.uni(-3, 3), rng.uni(-3, 3), rng.uni(-3, 3));
        V3 a = c + V3(rng.uni(-0.5f, 0.5f), rng.uni(-0.5f, 0.5f), rng.uni(-0.5f, 0.5f));
        V3 b = c + V3(rng.uni(-0.5f, 0.5f), rng.uni(-0.5f, 0.5f), rng.uni(-0.5f, 0.5f));
        add_tri(soup, c, a, b, V3(0, 1, 0), (uint8_t)(i % 3 == 0 ? 0 : (i % 6)));
    }
    ball.materials = {mat_dielectric(1.5f), mat_conductor(CTL_DISTR_BECKMANN, 0.1f)}; ball.emissive = {V3(0.0f), V3(0.0f)};
    add_icosphere(ball, V3(0, 0, 0), 1.0f, 2, 0);
    std::vector<MeshInput> meshes = {room, soup, ball};
    std::vector<NodeInput> nodes = {{0, M4::identity(), -1}, {1, M4::identity(), -1},
                                    {2, M4::translate(V3(1.5f, -2.0f, 0.5f)).mul(M4::scale(V3(0.8f, 0.8f, 0.8f))), -1},
                                    {2, M4::translate(V3(-1.5f, -1.0f, 1.0f)).mul(M4::rotate_y(0.7f)).mul(M4::scale(V3(0.6f, 1.1f, 0.6f))), -1}};
    assemble_scene(meshes, nodes, V3(0, 0, -3.8f), V3(0, 0, 0), V3(0, 1, 0), 60.0f, w, h, S);
}
} // namespace

void make_scene(int kind, int w, int h, uint32_t seed, int n_hint, SceneStorage& S) {
    switch (kind) {
    case 0: case 1: cornell(kind, w, h, S); break;
    case 2: scene_100k(false, w, h, seed, S); break;
    case 3: scene_100k(true, w, h, seed, S); break;
    case 4: scene_1m(false, w, h, seed, n_hint > 0 ? n_hint : 703, n_hint > 0 ? n_hint * 142 : 100000, S); break;
    case 5: scene_1m(true, w, h, seed, n_hint > 0 ? n_hint : 703, n_hint > 0 ? n_hint * 142 : 100000, S); break;
    case 6: scene_soup(w, h, seed, n_hint > 0 ? n_hint : 500, S); break;
    default: throw std::runtime_error("unknown scene kind");
    }
}

} // namespace ctlb
