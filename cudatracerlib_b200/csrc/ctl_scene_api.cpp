// ctl_scene_api.cpp -- host-side scene entry points of the C ABI (include/ctl_b200.h) and the per-thread error text.  No device code.
#include <string>
#include <vector>
#include <cstring>
#include <stdexcept>
#include <memory>
#include "../../include/ctl_b200.h"
#include "scene_builder.h"
#include "sampler_tables.h"

struct ctl_scene { ctlb::SceneStorage S; };

static thread_local std::string g_err;
int ctl_set_err(const std::string& s) { g_err = s; return 1; }
#define set_err ctl_set_err

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
// Passes first .. first+n-1 into n consecutive table sets, the passes produced concurrently (ctlb::generate_passes: start states by jump-ahead); what
// the context does for the frames it renders with host-generated tables.  Bit-identical to n ctl_generate_sample_tables calls (tests/test_golden_cpu.py).
int ctl_generate_sample_tables_n(uint32_t first, int n, float* d1, float* d2) {
    if (!d1 || !d2 || n < 1) return 1;
    ctlb::SamplerTableGenerator g;
    const size_t T1 = (size_t)ctlb::kNumSeq * ctlb::kSeqLen;
    for (uint32_t p = 0; p < first; p++) { ctlb::pass_jump().apply(g.v, g.v); g.d += 362437u * (uint32_t)ctlb::kDrawsPerPass; }   // skip whole passes by jump-ahead
    ctlb::generate_passes_threaded(g, n, d1, d2, T1, 2 * T1);
    return 0;
}


} // extern "C"
