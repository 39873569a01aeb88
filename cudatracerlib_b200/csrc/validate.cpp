// validate.cpp -- structural check of a scene view before it reaches the GPU.
//
// The reference trusts its own builders and file writers: a damaged .xmsh (a child index past the node array, a leaf run without its end flag, a
// cycle) makes its traversal (Kernel/TraceHelper.cu:88-172, Kernel/BVHTraversal.h:122-232) read out of bounds or spin.  On a GPU that is a hung
// device, so everything that comes from outside this library (files, caller-built views) can be checked here first.  What is checked is exactly what
// the traversal kernels rely on (csrc/device/traverse_persistent.cuh):
//   * every inner child reference is a multiple of 4 (float4 units) inside the mesh's / the scene's node array, every node is reached at most once
//     (the structure is a tree, so the walk terminates), no node is its own ancestor;
//   * every leaf reference of a mesh tree points inside the mesh's reference array and its run ends (bit 0 of a leaf word) inside that array;
//     every leaf word names a triangle of the mesh; every leaf of the scene-level tree names an existing instance;
//   * depth(scene tree) + 1 + depth(mesh tree) + 1 fits the 64-entry traversal stack for every instance;
//   * mesh / material / light indices of nodes, triangles and lights are in range.
// Box planes are not examined: NaN or inverted boxes only make rays miss.
#include <algorithm>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#include "scene_builder.h"

namespace ctlb {

namespace {
struct TreeInfo { int depth = 0; uint32_t nodes_reached = 0; };

// Walks one tree of `n_nodes` nodes; leaf(ref) validates a leaf reference (already complemented).  Iterative: damaged files may be deep.
template <typename LEAF> TreeInfo walk_tree(const ctl_bvh_node* nodes, uint32_t n_nodes, int start, const std::string& what, LEAF leaf) {
    TreeInfo info;
    if (start < 0) { leaf((uint32_t)~start); return info; }
    if (start == CTL_SENTINEL) return info;
    std::vector<unsigned char> seen(n_nodes, 0);
    std::vector<std::pair<uint32_t, int>> todo;
    auto enter = [&](int ref, int depth) {
        if (ref < 0) { leaf((uint32_t)~ref); return; }
        if (ref == CTL_SENTINEL) return;   // single-primitive root: the second child is the sentinel (SplitBVHBuilder.cpp:176-189)
        if ((ref & 3) || (uint32_t)ref / 4 >= n_nodes) throw std::runtime_error(what + ": child reference " + std::to_string(ref) + " outside the node array (" + std::to_string(n_nodes) + " nodes)");
        const uint32_t i = (uint32_t)ref / 4;
        if (seen[i]) throw std::runtime_error(what + ": node " + std::to_string(i) + " is referenced twice (not a tree)");
        seen[i] = 1; info.nodes_reached++;
        todo.emplace_back(i, depth);
    };
    enter(start, 1);
    while (!todo.empty()) {
        const auto cur = todo.back(); todo.pop_back();
        if (cur.second > info.depth) info.depth = cur.second;
        enter(nodes[cur.first].child0, cur.second + 1);
        enter(nodes[cur.first].child1, cur.second + 1);
    }
    return info;
}
} // namespace

// Mesh level of one mesh: nodes [0, n_nodes), leaf words [0, n_refs) with triangle ids < n_tris.  Returns the tree depth.
static int validate_mesh_tree(const ctl_bvh_node* nodes, uint32_t n_nodes, const uint32_t* index, uint32_t n_refs, uint32_t n_tris, const std::string& what, bool check_words) {
    if (!n_nodes || !n_refs) throw std::runtime_error(what + ": empty BVH");
    // end_after[i]: a run starting at i meets an end flag before the array ends  <=>  some word in [i, n_refs) has bit 0; true for all i iff the last has
    if (!(index[n_refs - 1] & 1u)) throw std::runtime_error(what + ": the last leaf run has no end flag");
    if (check_words) for (uint32_t i = 0; i < n_refs; i++) if ((index[i] >> 1) >= n_tris) throw std::runtime_error(what + ": leaf word " + std::to_string(i) + " references triangle " + std::to_string(index[i] >> 1) + " of " + std::to_string(n_tris));
    const TreeInfo t = walk_tree(nodes, n_nodes, 0, what, [&](uint32_t ref) {
        if (ref >= n_refs) throw std::runtime_error(what + ": leaf reference " + std::to_string(ref) + " outside the reference array (" + std::to_string(n_refs) + ")");
    });
    return t.depth;
}

int validate_mesh_bvh(const ctl_bvh_node* nodes, uint32_t n_nodes, const uint32_t* index, uint32_t n_refs, uint32_t n_tris, const std::string& what) {
    return validate_mesh_tree(nodes, n_nodes, index, n_refs, n_tris, what, true);
}

void validate_view(const ctl_scene_view& v) {
    if (!v.n_nodes) return; // empty scene: every ray misses (TraceHelper.cu:92)
    if (!v.bvh_nodes || !v.woop || !v.tri_index || !v.tri_data || !v.meshes || !v.nodes || !v.node_xf || !v.node_inv_xf || !v.materials)
        throw std::runtime_error("scene view: null array");
    if (v.n_woop != v.n_tri_index) throw std::runtime_error("scene view: woop / index arrays differ in length");
    std::vector<int> mesh_depth(v.n_meshes, -1);
    std::set<uint32_t> words_checked;
    // arrays of a mesh end where the next larger offset of any mesh begins (re-braided views hold many mesh records over the same triangle / reference ranges)
    std::vector<uint32_t> node_offs, ref_offs, tri_offs;
    for (uint32_t o = 0; o < v.n_meshes; o++) { node_offs.push_back(v.meshes[o].bvh_node_offset / 4); ref_offs.push_back(v.meshes[o].bvh_idx_offset); tri_offs.push_back(v.meshes[o].tri_offset); }
    std::sort(node_offs.begin(), node_offs.end()); std::sort(ref_offs.begin(), ref_offs.end()); std::sort(tri_offs.begin(), tri_offs.end());
    auto next_after = [](const std::vector<uint32_t>& sorted, uint32_t x, uint32_t end) { auto it = std::upper_bound(sorted.begin(), sorted.end(), x); return it == sorted.end() || *it > end ? end : *it; };
    auto mesh_extent = [&](uint32_t m, uint32_t& node0, uint32_t& n_nodes, uint32_t& ref0, uint32_t& n_refs, uint32_t& n_tris) {
        const ctl_mesh& K = v.meshes[m];
        if ((K.bvh_node_offset & 3) || K.bvh_node_offset / 4 >= v.n_bvh_nodes || K.bvh_idx_offset >= v.n_tri_index || K.tri_offset >= v.n_tri_data)
            throw std::runtime_error("scene view: mesh " + std::to_string(m) + " offsets outside the arrays");
        if ((uint64_t)K.bvh_tri_offset != (uint64_t)K.bvh_idx_offset * 3) throw std::runtime_error("scene view: mesh " + std::to_string(m) + " woop offset is not 3 x its index offset");
        node0 = K.bvh_node_offset / 4; n_nodes = next_after(node_offs, node0, v.n_bvh_nodes) - node0;
        ref0 = K.bvh_idx_offset; n_refs = next_after(ref_offs, ref0, v.n_tri_index) - ref0;
        n_tris = next_after(tri_offs, K.tri_offset, v.n_tri_data) - K.tri_offset;
    };
    for (uint32_t m = 0; m < v.n_meshes; m++) {
        uint32_t node0, n_nodes, ref0, n_refs, n_tris;
        mesh_extent(m, node0, n_nodes, ref0, n_refs, n_tris);
        const bool first_use = words_checked.insert(ref0).second;   // re-braided views: thousands of mesh records share one reference range -- its words are scanned once
        mesh_depth[m] = validate_mesh_tree(v.bvh_nodes + node0, n_nodes, v.tri_index + ref0, n_refs, n_tris, "mesh " + std::to_string(m), first_use);
        const ctl_mesh& K = v.meshes[m];
        if (K.mat_offset > v.n_materials) throw std::runtime_error("scene view: mesh " + std::to_string(m) + " material offset out of range");
    }
    std::vector<unsigned char> node_in_tree(v.n_nodes, 0);
    const TreeInfo top = walk_tree(v.scene_bvh_nodes, v.n_scene_bvh_nodes, v.scene_start_node, "scene-level BVH", [&](uint32_t ref) {
        if (ref >= v.n_nodes) throw std::runtime_error("scene-level BVH: leaf names instance " + std::to_string(ref) + " of " + std::to_string(v.n_nodes));
        node_in_tree[ref] = 1;
    });
    for (uint32_t n = 0; n < v.n_nodes; n++) {
        const ctl_node& N = v.nodes[n];
        if (N.mesh_index >= v.n_meshes) throw std::runtime_error("scene view: node " + std::to_string(n) + " names mesh " + std::to_string(N.mesh_index) + " of " + std::to_string(v.n_meshes));
        if (N.material_offset > v.n_materials) throw std::runtime_error("scene view: node " + std::to_string(n) + " material offset out of range");
        if (node_in_tree[n] && top.depth + 1 + mesh_depth[N.mesh_index] + 1 > 64)
            throw std::runtime_error("scene view: node " + std::to_string(n) + ": tree depths " + std::to_string(top.depth) + " + " + std::to_string(mesh_depth[N.mesh_index]) + " exceed the 64-entry traversal stack");
    }
    // every triangle's material byte (TriangleData::getMatIndex, Engine/TriangleData.h:40-44) + the node's material offset names an existing material
    {
        std::vector<int> mesh_max_mat(v.n_meshes, -1);
        for (uint32_t n = 0; n < v.n_nodes; n++) {
            const uint32_t m = v.nodes[n].mesh_index;
            if (mesh_max_mat[m] < 0) {
                uint32_t node0, n_nodes, ref0, n_refs, n_tris;
                mesh_extent(m, node0, n_nodes, ref0, n_refs, n_tris);
                int mx = 0;
                for (uint32_t t = 0; t < n_tris; t++) mx = std::max(mx, (int)((v.tri_data[v.meshes[m].tri_offset + t].w[1] >> 16) & 0xffu));
                mesh_max_mat[m] = mx;
            }
            if ((uint64_t)v.nodes[n].material_offset + (uint64_t)mesh_max_mat[m] >= v.n_materials)
                throw std::runtime_error("scene view: node " + std::to_string(n) + ": a triangle names material " + std::to_string(mesh_max_mat[m]) + " + offset " + std::to_string(v.nodes[n].material_offset) + " of " + std::to_string(v.n_materials));
        }
    }
    if (v.num_lights > CTL_MAX_NUM_LIGHTS) throw std::runtime_error("scene view: too many lights");
    if (v.num_lights && (!v.lights || !v.light_tris || !v.light_cdf_data)) throw std::runtime_error("scene view: null light array");
    for (uint32_t i = 0; i < v.num_lights; i++) {
        if (v.light_indices[i] >= v.n_lights_buf) throw std::runtime_error("scene view: light index out of range");
        const ctl_light& L = v.lights[v.light_indices[i]];
        if ((uint64_t)L.tri_offset + L.count > v.n_light_tris || (uint64_t)L.cdf_offset + L.count + 1 > v.n_light_cdf_data || L.node_idx >= v.n_nodes)
            throw std::runtime_error("scene view: light " + std::to_string(i) + " ranges outside the light arrays");
        for (uint32_t k = 0; k < L.count; k++) {
            const ctl_light_tri& T = v.light_tris[L.tri_offset + k];
            if (T.i_dat >= v.n_woop || T.t_dat >= v.n_tri_data) throw std::runtime_error("scene view: light " + std::to_string(i) + " triangle " + std::to_string(k) + " points outside the triangle arrays");
        }
    }
}

// Deepest traversal-stack use of a view built in this library (trees are trees): depth(scene level) + 1 + depth(mesh) + 1 over the instances.
int view_stack_depth(const ctl_scene_view& v) {
    auto depth_of = [](const ctl_bvh_node* nodes, uint32_t n_nodes, int start) {
        if (start < 0 || start == CTL_SENTINEL || !n_nodes) return 0;
        int deepest = 0; uint64_t visited = 0;
        std::vector<std::pair<uint32_t, int>> todo; todo.emplace_back((uint32_t)start / 4, 1);
        while (!todo.empty()) {
            const auto cur = todo.back(); todo.pop_back();
            if (cur.first >= n_nodes || ++visited > n_nodes) return 1 << 20;   // not a tree of this array
            deepest = std::max(deepest, cur.second);
            for (int c : {nodes[cur.first].child0, nodes[cur.first].child1}) if (c >= 0 && c != CTL_SENTINEL) todo.emplace_back((uint32_t)c / 4, cur.second + 1);
        }
        return deepest;
    };
    if (!v.n_nodes) return 0;
    const int top = depth_of(v.scene_bvh_nodes, v.n_scene_bvh_nodes, v.scene_start_node);
    int worst = 0;
    std::vector<int> mesh_depth(v.n_meshes, -1);
    for (uint32_t n = 0; n < v.n_nodes; n++) {
        const uint32_t m = v.nodes[n].mesh_index;
        if (m >= v.n_meshes) continue;
        if (mesh_depth[m] < 0) { const uint32_t node0 = v.meshes[m].bvh_node_offset / 4; mesh_depth[m] = node0 < v.n_bvh_nodes ? depth_of(v.bvh_nodes + node0, v.n_bvh_nodes - node0, 0) : 0; }
        worst = std::max(worst, top + 1 + mesh_depth[m] + 1);
    }
    return worst;
}

} // namespace ctlb
