// scene_builder.h -- host-side construction of the flat scene view.
//
// Replaces, for synthetic scenes, the reference's DynamicScene / Mesh::CompileMesh /
// SplitBVHBuilder / SceneBVH / ShapeSet host machinery with a
// small builder that EMITS EXACTLY the reference encodings:
//   BVHNodeData   Engine/TriIntersectorData.h:42-117, child/leaf rules SplitBVHBuilder.cpp:163-203
//   Woop tris     Engine/TriIntersectorData.cu:5-18
//   leaf words    Engine/TriIntersectorData.h:8-28, BVHBuilderHelper.cpp:51-62
//   TriangleData  Engine/TriangleData.cu:8-65
//   vertex normals Engine/Mesh.cpp:151-190
//   ShapeSet tris Engine/ShapeSet.cu:11-22, ShapeSet.cpp:37-59
//   light CDF     Engine/DynamicScene.cpp:173-196, rayEps :587
//   camera        SceneTypes/Sensor.cu:76-96
#pragma once
#include <vector>
#include <cstdint>
#include "../../include/ctl_b200.h"
#include <string>
#include "host_math.h"

namespace ctlb {

struct MeshInput {
    std::vector<V3> verts;
    std::vector<uint32_t> indices;   // 3 per triangle
    std::vector<uint8_t> mat_index;  // per triangle, local to the mesh's material block
    std::vector<ctl_material> materials;
    std::vector<V3> emissive;        // per material; non-zero => area light
    std::vector<V3> normals;         // optional, per vertex: used (normalised) instead of computed vertex normals (Mesh::CompileMesh a_normals)
    std::vector<float> uvs;          // optional, 2 per vertex: UV set 0 (drives dpdu / dpdv, TriangleData.cu:37-57)
    // pre-compiled mesh (.xmsh import, xmsh.cpp): when pre_tri_data is non-empty the arrays below are taken as they are (reference layouts)
    // instead of being built from verts / indices; mat_index is then derived from the TriangleData words
    std::vector<ctl_tri_data> pre_tri_data;
    std::vector<ctl_bvh_node> pre_nodes;
    std::vector<ctl_woop_tri> pre_woop;
    std::vector<uint32_t> pre_index;
    Box pre_box;
};

// .xmsh (the reference's compiled-mesh format: Engine/Mesh.cpp:46-98 reader, :199-290 writer, Engine/MeshLoader/BVHBuilderHelper.cpp:129-147)
void read_xmsh(const char* path, MeshInput& out);                                  // throws std::runtime_error
void read_ply(const char* path, MeshInput& out);                                   // PLY (obj_import.cpp), == the reference's compileply front end
void read_obj(const char* path, MeshInput& out);                                   // Wavefront OBJ + MTL (obj_import.cpp), == the reference's compileobj front end
void write_xmsh(const char* path, const struct SceneStorage& S, uint32_t mesh);    // mesh `mesh` of an assembled scene

struct NodeInput {
    uint32_t mesh;
    M4 xf;
    int material_override; // -1: use the mesh's material block; else global material offset
};

constexpr uint32_t kRebraidAuto = 0xffffffffu;
struct SceneStorage {
    std::vector<ctl_bvh_node> bvh_nodes;
    std::vector<ctl_woop_tri> woop;
    std::vector<uint32_t> tri_index;
    std::vector<ctl_tri_data> tri_data;
    std::vector<ctl_mesh> meshes;
    std::vector<Box> mesh_boxes;                   // per mesh: local AABB (Mesh::m_sLocalBox)
    // what the node level is (re)assembled from (assemble_nodes): instances, per-mesh material counts / emission, camera
    std::vector<NodeInput> node_inputs;
    std::vector<uint32_t> mesh_n_materials;
    std::vector<std::vector<V3>> mesh_emissive;
    V3 cam_pos, cam_target, cam_up; float cam_fov = 60.0f; int cam_w = 0, cam_h = 0;
    std::vector<std::vector<float>> mesh_verts9;   // per mesh: 9 floats per triangle (for BVH rebuilds, e.g. on the GPU)
    std::vector<ctl_node> nodes;
    std::vector<float> node_xf, node_inv_xf;
    std::vector<ctl_bvh_node> scene_bvh;
    int32_t scene_start = 0;
    std::vector<ctl_material> materials;
    std::vector<ctl_light> lights;
    std::vector<ctl_light_tri> light_tris;
    std::vector<float> light_cdf_data;
    uint32_t num_lights = 0;
    uint32_t light_indices[CTL_MAX_NUM_LIGHTS];
    float light_cdf[CTL_MAX_NUM_LIGHTS];
    ctl_camera camera;
    Box box;
    float ray_eps = 0;
    // Partial re-braiding of the scene level (opt-in, rebraid()): instances whose boxes overlap are opened into sub-tree entries.  The combined arrays
    // (real + pseudo nodes / meshes / BVH nodes) live beside the real ones, which stay what every other function reads; fill_view hands out the
    // combined set while rb_active.
    uint32_t rebraid_entries = 0xffffffffu;        // budget of scene-level leaves; 0 = off; kRebraidAuto = by scene size (or CTL_REBRAID in the environment)
    bool rb_active = false;
    std::vector<ctl_bvh_node> rb_bvh_nodes, rb_scene_bvh;
    std::vector<ctl_mesh> rb_meshes;
    std::vector<ctl_node> rb_nodes;
    std::vector<float> rb_node_xf, rb_node_inv_xf;
    std::vector<uint32_t> rb_node_alias;           // real node of every (pseudo-)node
    void fill_view(ctl_scene_view* v) const;
};

// BVH over boxes; emits reference node layout. leaf_cb(first, count order) appends leaf payload
// and returns the first slot. max_leaf = 8 for meshes (BVHBuilderHelper.cpp:119), 1 for the scene level.
struct BvhBuildResult { std::vector<ctl_bvh_node> nodes; std::vector<uint32_t> leaf_order; std::vector<uint32_t> leaf_first; };
void build_bvh(const std::vector<Box>& prim_boxes, int max_leaf, std::vector<ctl_bvh_node>& nodes_out,
               std::vector<uint32_t>& ordered_prims_out, std::vector<uint8_t>& last_in_leaf_out);

// Mesh-level builder: split BVH (SAH object splits + spatial splits of triangle references), sbvh_builder.cpp.  `ordered` may name a triangle
// more than once (one Woop record + leaf word per reference).  CTL_BVH_BUILDER=sah in the environment selects build_bvh instead (A/B).
void build_sbvh(const float* verts9, uint32_t n_tris, int max_leaf, std::vector<ctl_bvh_node>& nodes_out,
                std::vector<uint32_t>& ordered_prims_out, std::vector<uint8_t>& last_in_leaf_out);

// Post-build optimisation of a finished tree in the reference node layout (sub-tree re-insertion + rotations, sbvh_builder.cpp); leaves are untouched
void optimize_bvh(std::vector<ctl_bvh_node>& nodes);

// Woop unit-triangle transform (Engine/TriIntersectorData.cu:5-18); host + device (GPU BVH builder), same expressions
CTLB_HD inline void encode_woop(V3 v0, V3 v1, V3 v2, ctl_woop_tri* out) {
    // M = [v0-v2 | v1-v2 | (v0-v2)x(v1-v2) | v2] (columns), inverted; store row2 (w negated), row0, row1.
    V3 e0 = v0 - v2, e1 = v1 - v2, nn = cross(e0, e1);
    M4 m;
    m(0, 0) = e0.x; m(1, 0) = e0.y; m(2, 0) = e0.z; m(3, 0) = 0;
    m(0, 1) = e1.x; m(1, 1) = e1.y; m(2, 1) = e1.z; m(3, 1) = 0;
    m(0, 2) = nn.x; m(1, 2) = nn.y; m(2, 2) = nn.z; m(3, 2) = 0;
    m(0, 3) = v2.x; m(1, 3) = v2.y; m(2, 3) = v2.z; m(3, 3) = 1;
    M4 i = m.inverse();
    out->a[0] = i(2, 0); out->a[1] = i(2, 1); out->a[2] = i(2, 2); out->a[3] = -i(2, 3);
    for (int k = 0; k < 4; k++) { out->b[k] = i(0, k); out->c[k] = i(1, k); }
}
void decode_woop(const ctl_woop_tri& w, V3& v0, V3& v1, V3& v2);
void encode_tri_data(const V3 p[3], const V3 n[3], const float uv[6], uint32_t mat, ctl_tri_data* out);
void compute_vertex_normals(const std::vector<V3>& verts, const std::vector<uint32_t>& idx, std::vector<V3>& normals);
// host fillDG restricted to what ShapeSet needs (sys.n at a barycentric position)
V3 shading_normal_at(const ctl_tri_data& td, const M4& local_to_world, float u, float v);

void assemble_scene(const std::vector<MeshInput>& meshes, const std::vector<NodeInput>& nodes, V3 cam_pos, V3 cam_target,
                    V3 cam_up, float fov_deg, int width, int height, SceneStorage& out);
void make_camera(V3 pos, V3 target, V3 up, float fov_deg, int w, int h, ctl_camera* cam);
// Node level of a scene whose meshes are assembled: nodes, transforms, scene-level BVH, area lights (world-space ShapeSets), light CDF, scene box,
// ray epsilon, camera -- everything that changes when an instance moves (DynamicScene::SetNodeTransform + BVHRebuilder, Engine/DynamicScene.cpp:433-443).
// Re-runnable: ctl_scene_set_node_transform edits S.node_inputs and calls it again.
void assemble_nodes(SceneStorage& S);

// csrc/validate.cpp: structural checks of externally supplied BVHs / views (index ranges, tree shape, leaf-run end flags, stack depth); throw std::runtime_error
int validate_mesh_bvh(const ctl_bvh_node* nodes, uint32_t n_nodes, const uint32_t* index, uint32_t n_refs, uint32_t n_tris, const std::string& what);
void validate_view(const ctl_scene_view& v);
int view_stack_depth(const ctl_scene_view& v);   // depth(scene level) + 1 + depth(mesh) + 1, worst instance: must fit the kernels' 64-entry stack

// synthetic scenes
void make_scene(int kind, int width, int height, uint32_t seed, int n_hint, SceneStorage& out);

} // namespace ctlb
