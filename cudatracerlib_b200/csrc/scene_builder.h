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
#include "host_math.h"

namespace ctlb {

struct MeshInput {
    std::vector<V3> verts;
    std::vector<uint32_t> indices;   // 3 per triangle
    std::vector<uint8_t> mat_index;  // per triangle, local to the mesh's material block
    std::vector<ctl_material> materials;
    std::vector<V3> emissive;        // per material; non-zero => area light
};

struct NodeInput {
    uint32_t mesh;
    M4 xf;
    int material_override; // -1: use the mesh's material block; else global material offset
};

struct SceneStorage {
    std::vector<ctl_bvh_node> bvh_nodes;
    std::vector<ctl_woop_tri> woop;
    std::vector<uint32_t> tri_index;
    std::vector<ctl_tri_data> tri_data;
    std::vector<ctl_mesh> meshes;
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
    void fill_view(ctl_scene_view* v) const;
};

// BVH over boxes; emits reference node layout. leaf_cb(first, count order) appends leaf payload
// and returns the first slot. max_leaf = 8 for meshes (BVHBuilderHelper.cpp:119), 1 for the scene level.
struct BvhBuildResult { std::vector<ctl_bvh_node> nodes; std::vector<uint32_t> leaf_order; std::vector<uint32_t> leaf_first; };
void build_bvh(const std::vector<Box>& prim_boxes, int max_leaf, std::vector<ctl_bvh_node>& nodes_out,
               std::vector<uint32_t>& ordered_prims_out, std::vector<uint8_t>& last_in_leaf_out);

void encode_woop(V3 v0, V3 v1, V3 v2, ctl_woop_tri* out);
void decode_woop(const ctl_woop_tri& w, V3& v0, V3& v1, V3& v2);
void encode_tri_data(const V3 p[3], const V3 n[3], const float uv[6], uint32_t mat, ctl_tri_data* out);
void compute_vertex_normals(const std::vector<V3>& verts, const std::vector<uint32_t>& idx, std::vector<V3>& normals);
// host fillDG restricted to what ShapeSet needs (sys.n at a barycentric position)
V3 shading_normal_at(const ctl_tri_data& td, const M4& local_to_world, float u, float v);

void assemble_scene(const std::vector<MeshInput>& meshes, const std::vector<NodeInput>& nodes, V3 cam_pos, V3 cam_target,
                    V3 cam_up, float fov_deg, int width, int height, SceneStorage& out);
void make_camera(V3 pos, V3 target, V3 up, float fov_deg, int w, int h, ctl_camera* cam);

// synthetic scenes
void make_scene(int kind, int width, int height, uint32_t seed, int n_hint, SceneStorage& out);

} // namespace ctlb
