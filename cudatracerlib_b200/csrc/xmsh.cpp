// xmsh.cpp -- the reference's compiled-mesh format ".xmsh" (SURVEY 8 f4): the on-disk form that feeds real scenes into the scene view.
//
// Reader == the static-mesh branch of DynamicScene::CreateNode + Mesh::Mesh(IInStream&) (Engine/DynamicScene.cpp:313-319, Engine/Mesh.cpp:46-98);
// writer == Mesh::CompileMesh's output sequence (Engine/Mesh.cpp:278-289) + ConstructBVH (Engine/MeshLoader/BVHBuilderHelper.cpp:129-147).
// Little-endian, no padding between records:
//
//   u32  MeshCompileType (0 = Static; 1 = Animated is not on this path)            Engine/MeshLoader/MeshCompiler.cpp:94
//   AABB local box (6 floats)                                                      Mesh.cpp:278
//   u32  numLights;  MeshPartLight[numLights]  = { FixedString<32> MatName (u32 length incl. NUL + 32 chars), Spectrum L }   48 B, Mesh.h:21-34
//   u32  numTriangles;  TriangleData[numTriangles]                                 32 B each
//   u32  numMaterials;  Material[numMaterials]                                     3344 B each (tagged unions, see below)
//   u64  numNodes;  BVHNodeData[numNodes]                                          64 B each
//   u64  numRefs;   TriIntersectorData[numRefs]                                    48 B each (Woop), leaf order
//   u64  numRefs;   TriIntersectorData2[numRefs]                                   4 B each (tri << 1 | last-in-leaf)
//
// Material (Engine/Material.h:38-61) is a 3344-byte struct of tagged unions (CudaVirtualAggregate: u32 type tag, payload at +16, the payload
// starts with a vtable pointer the reference re-patches after loading, Mesh.cpp:67-75).  The path reads: Name (+0, FixedString<64>),
// NodeLightIndex (+68), bsdf (+512: tag; payload +528 = BSDF base {vptr, m_combinedType +8, ..., m_enableTwoSided +52} followed by the BSDF's
// fields).  Supported BSDFs are the three of the hot path -- diffuse (tag 1), dielectric (3), roughconductor (7) -- with constant textures
// (Texture tag 2, value at payload +8); anything else is rejected with a message naming the material.  All offsets are static_assert-ed against
// the reference headers in oracle/ref_driver.cpp.
#include "scene_builder.h"
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace ctlb {
namespace {

constexpr size_t MATERIAL_SIZE = 3344, MAT_NAME = 0, MAT_NODE_LIGHT = 68, MAT_BSDF_TAG = 512, MAT_BSDF = 528;
constexpr size_t BSDF_COMBINED = 8, BSDF_TWO_SIDED = 52, TEX_SIZE = 208, TEX_PAYLOAD = 16, CONST_TEX_VAL = 8;
constexpr size_t DIFFUSE_REFL = 64;
constexpr size_t RC_REFL = 64, RC_ALPHA_U = 272, RC_ALPHA_V = 480, RC_ETA = 688, RC_K = 700, RC_TYPE = 716;
constexpr size_t DI_DISPERSION = 64, DI_TRANS = 128, DI_REFL = 336, DISP_PAYLOAD = 16, CAUCHY_B = 8, CAUCHY_C = 12;
constexpr uint32_t TAG_DIFFUSE = 1, TAG_DIELECTRIC = 3, TAG_ROUGHCONDUCTOR = 7, TAG_CONST_TEX = 2, TAG_CAUCHY = 1;
constexpr size_t LIGHT_SIZE = 48, LIGHT_L = 36;
constexpr size_t MAT_USED_BSSRDF = 88, MAT_NORMAL_MAP_USED = 2656, MAT_HEIGHT_MAP_USED = 2880, MAT_ALPHA_STATE = 3104; // the other things the path reads

struct File {
    FILE* f; std::string path;
    File(const char* p, const char* mode) : f(fopen(p, mode)), path(p) { if (!f) throw std::runtime_error("Could not open file: " + path); }
    ~File() { if (f) fclose(f); }
    void read(void* dst, size_t n) { if (n && fread(dst, 1, n, f) != n) throw std::runtime_error("Passed end of file: " + path); }
    // refuse to allocate for a record count the rest of the file cannot hold (corrupt / truncated headers)
    void need(uint64_t count, size_t record) {
        const long pos = ftell(f); fseek(f, 0, SEEK_END); const long end = ftell(f); fseek(f, pos, SEEK_SET);
        if (pos < 0 || end < pos || count > (uint64_t)(end - pos) / record) throw std::runtime_error("Passed end of file: " + path);
    }
    void write(const void* src, size_t n) { if (n && fwrite(src, 1, n, f) != n) throw std::runtime_error("Could not write to file: " + path); }
    template <typename T> T get() { T v; read(&v, sizeof(T)); return v; }
    template <typename T> void put(const T& v) { write(&v, sizeof(T)); }
};

std::string fixed_string(const unsigned char* p, size_t cap) { // FixedString<N>: u32 length (counts the NUL), chars
    uint32_t len; memcpy(&len, p, 4);
    if (len > cap) len = (uint32_t)cap;
    std::string s((const char*)p + 4, len);
    const size_t z = s.find('\0'); if (z != std::string::npos) s.resize(z);
    return s;
}
void put_fixed_string(unsigned char* p, size_t cap, const std::string& s) {
    const uint32_t len = (uint32_t)std::min(cap, s.size() + 1);
    memcpy(p, &len, 4); memcpy(p + 4, s.c_str(), len - 1); p[4 + len - 1] = 0;
}
uint32_t u32_at(const unsigned char* p, size_t off) { uint32_t v; memcpy(&v, p + off, 4); return v; }
float f32_at(const unsigned char* p, size_t off) { float v; memcpy(&v, p + off, 4); return v; }
void put_u32(unsigned char* p, size_t off, uint32_t v) { memcpy(p + off, &v, 4); }
void put_f32(unsigned char* p, size_t off, float v) { memcpy(p + off, &v, 4); }

void const_texture(const unsigned char* tex, const std::string& mat, const char* what, float rgb[3]) {
    if (u32_at(tex, 0) != TAG_CONST_TEX) throw std::runtime_error("material '" + mat + "': texture '" + what + "' is not a ConstantTexture (textures are outside the B200 path)");
    for (int k = 0; k < 3; k++) rgb[k] = f32_at(tex, TEX_PAYLOAD + CONST_TEX_VAL + 4 * k);
}
void put_const_texture(unsigned char* tex, const float rgb[3]) {
    put_u32(tex, 0, TAG_CONST_TEX);
    for (int k = 0; k < 3; k++) put_f32(tex, TEX_PAYLOAD + CONST_TEX_VAL + 4 * k, rgb[k]);
}

ctl_material decode_material(const unsigned char* m, std::string& name) {
    name = fixed_string(m + MAT_NAME, 64);
    ctl_material out; memset(&out, 0, sizeof(out));
    out.node_light_index = 0xffffffffu;
    const unsigned char* b = m + MAT_BSDF;
    out.flags = b[BSDF_TWO_SIDED] ? CTL_MAT_TWO_SIDED : 0u;
    // features of the reference's getBsdfSample / traceRay that the B200 path does not have must not be dropped silently
    if (u32_at(m, MAT_NORMAL_MAP_USED)) throw std::runtime_error("material '" + name + "': normal maps are outside the B200 path");
    if (u32_at(m, MAT_HEIGHT_MAP_USED)) throw std::runtime_error("material '" + name + "': height maps are outside the B200 path");
    if (u32_at(m, MAT_ALPHA_STATE)) throw std::runtime_error("material '" + name + "': alpha maps are outside the B200 path");
    if (u32_at(m, MAT_USED_BSSRDF)) throw std::runtime_error("material '" + name + "': subsurface scattering (bssrdf) is outside the B200 path");
    const uint32_t tag = u32_at(m, MAT_BSDF_TAG);
    out.transmittance = 1.0f; out.alpha_u = out.alpha_v = 0.1f; out.eta[0] = out.eta[1] = out.eta[2] = 1.5f;
    if (tag == TAG_DIFFUSE) {
        out.bsdf_type = CTL_BSDF_DIFFUSE;
        const_texture(b + DIFFUSE_REFL, name, "m_reflectance", out.reflectance);
    } else if (tag == TAG_ROUGHCONDUCTOR) {
        out.bsdf_type = CTL_BSDF_ROUGHCONDUCTOR;
        const_texture(b + RC_REFL, name, "m_specularReflectance", out.reflectance);
        float a[3]; const_texture(b + RC_ALPHA_U, name, "m_alphaU", a); out.alpha_u = a[0]; const_texture(b + RC_ALPHA_V, name, "m_alphaV", a); out.alpha_v = a[0];
        for (int k = 0; k < 3; k++) { out.eta[k] = f32_at(b, RC_ETA + 4 * k); out.k[k] = f32_at(b, RC_K + 4 * k); }
        const uint32_t distr = u32_at(b, RC_TYPE);
        if (distr > 1) throw std::runtime_error("material '" + name + "': microfacet distribution " + std::to_string(distr) + " not supported (Beckmann, GGX)");
        out.distr_type = distr;
    } else if (tag == TAG_DIELECTRIC) {
        out.bsdf_type = CTL_BSDF_DIELECTRIC;
        const unsigned char* d = b + DI_DISPERSION;
        if (u32_at(d, 0) != TAG_CAUCHY) throw std::runtime_error("material '" + name + "': only Cauchy dispersion is supported");
        const float B = f32_at(d, DISP_PAYLOAD + CAUCHY_B), C = f32_at(d, DISP_PAYLOAD + CAUCHY_C);
        if (C != 0.0f) throw std::runtime_error("material '" + name + "': dispersive dielectrics are outside the B200 path");
        out.eta[0] = out.eta[1] = out.eta[2] = B; // DispersionCauchy::calc_eta with C = 0 (SceneTypes/Dispersion.h:26-29)
        float t[3]; const_texture(b + DI_TRANS, name, "m_specularTransmittance", t); out.transmittance = t[0];
        const_texture(b + DI_REFL, name, "m_specularReflectance", out.reflectance);
    } else throw std::runtime_error("material '" + name + "': BSDF type " + std::to_string(tag) + " is not supported by the B200 path (diffuse, roughconductor, dielectric)");
    return out;
}

void encode_material(const ctl_material& in, const std::string& name, unsigned char* m) {
    memset(m, 0, MATERIAL_SIZE);
    put_fixed_string(m + MAT_NAME, 64, name);
    put_u32(m, MAT_NODE_LIGHT, 0xffffffffu);
    put_f32(m, MAT_NODE_LIGHT + 4, 1.0f); // HeightScale (Material.cpp ctor default)
    unsigned char* b = m + MAT_BSDF;
    b[BSDF_TWO_SIDED] = (in.flags & CTL_MAT_TWO_SIDED) ? 1 : 0;
    const float one[3] = {1.0f, 1.0f, 1.0f};
    if (in.bsdf_type == CTL_BSDF_DIFFUSE) {
        put_u32(m, MAT_BSDF_TAG, TAG_DIFFUSE); put_u32(b, BSDF_COMBINED, 0x2u);
        put_const_texture(b + DIFFUSE_REFL, in.reflectance);
    } else if (in.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) {
        put_u32(m, MAT_BSDF_TAG, TAG_ROUGHCONDUCTOR); put_u32(b, BSDF_COMBINED, 0x8u);
        put_const_texture(b + RC_REFL, in.reflectance);
        const float au[3] = {in.alpha_u, in.alpha_u, in.alpha_u}, av[3] = {in.alpha_v, in.alpha_v, in.alpha_v};
        put_const_texture(b + RC_ALPHA_U, au); put_const_texture(b + RC_ALPHA_V, av);
        for (int k = 0; k < 3; k++) { put_f32(b, RC_ETA + 4 * k, in.eta[k]); put_f32(b, RC_K + 4 * k, in.k[k]); }
        put_u32(b, RC_TYPE, in.distr_type);
    } else {
        put_u32(m, MAT_BSDF_TAG, TAG_DIELECTRIC); put_u32(b, BSDF_COMBINED, 0x20u | 0x40u);
        unsigned char* d = b + DI_DISPERSION;
        put_u32(d, 0, TAG_CAUCHY); put_f32(d, DISP_PAYLOAD + CAUCHY_B, in.eta[0]); put_f32(d, DISP_PAYLOAD + CAUCHY_C, 0.0f);
        const float t[3] = {in.transmittance, in.transmittance, in.transmittance};
        put_const_texture(b + DI_TRANS, t); put_const_texture(b + DI_REFL, in.reflectance);
        (void)one;
    }
}

} // namespace

void read_xmsh(const char* path, MeshInput& M) {
    File in(path, "rb");
    M = MeshInput();
    const uint32_t token = in.get<uint32_t>();
    if (token == 1) throw std::runtime_error(std::string("animated meshes (MeshCompileType::Animated) are not on the B200 path: ") + path);
    if (token != 0) throw std::runtime_error(std::string("Mesh file parser error. ") + path); // DynamicScene.cpp:319
    float box[6]; in.read(box, sizeof(box));
    M.pre_box.lo = V3(box[0], box[1], box[2]); M.pre_box.hi = V3(box[3], box[4], box[5]);
    const uint32_t n_lights = in.get<uint32_t>();
    if (n_lights > 4096) throw std::runtime_error(std::string("Mesh file parser error (light count). ") + path);
    std::vector<std::pair<std::string, V3>> lights;
    for (uint32_t i = 0; i < n_lights; i++) {
        unsigned char rec[LIGHT_SIZE]; in.read(rec, LIGHT_SIZE);
        lights.emplace_back(fixed_string(rec, 32), V3(f32_at(rec, LIGHT_L), f32_at(rec, LIGHT_L + 4), f32_at(rec, LIGHT_L + 8)));
    }
    const uint32_t n_tris = in.get<uint32_t>();
    if (n_tris == 0 || n_tris > 0x3fffffffu) throw std::runtime_error(std::string("Mesh file parser error (triangle count). ") + path);
    in.need(n_tris, sizeof(ctl_tri_data));
    M.pre_tri_data.resize(n_tris); in.read(M.pre_tri_data.data(), (size_t)n_tris * sizeof(ctl_tri_data));
    const uint32_t n_mats = in.get<uint32_t>();
    if (n_mats == 0 || n_mats > 256) throw std::runtime_error(std::string("Mesh file parser error (material count). ") + path);
    std::vector<unsigned char> blob(MATERIAL_SIZE);
    std::vector<std::string> names(n_mats);
    for (uint32_t i = 0; i < n_mats; i++) { in.read(blob.data(), MATERIAL_SIZE); M.materials.push_back(decode_material(blob.data(), names[i])); }
    M.emissive.assign(n_mats, V3(0.0f));
    for (auto& l : lights) { // DynamicScene::CreateLight(node, MatName, L): the light belongs to the material of that name (DynamicScene.cpp:340-341, 689-711)
        bool found = false;
        for (uint32_t i = 0; i < n_mats; i++) if (names[i] == l.first) { M.emissive[i] = l.second; found = true; break; }
        if (!found) throw std::runtime_error("area light refers to unknown material '" + l.first + "': " + path);
    }
    const uint64_t n_nodes = in.get<uint64_t>();
    if (n_nodes == 0 || n_nodes > 0x7fffffffull / 4) throw std::runtime_error(std::string("Mesh file parser error (node count). ") + path);
    in.need(n_nodes, sizeof(ctl_bvh_node));
    M.pre_nodes.resize((size_t)n_nodes); in.read(M.pre_nodes.data(), (size_t)n_nodes * sizeof(ctl_bvh_node));
    const uint64_t n_refs = in.get<uint64_t>();
    if (n_refs == 0 || n_refs > 0x7fffffffull) throw std::runtime_error(std::string("Mesh file parser error (triangle reference count). ") + path);
    in.need(n_refs, sizeof(ctl_woop_tri));
    M.pre_woop.resize((size_t)n_refs); in.read(M.pre_woop.data(), (size_t)n_refs * sizeof(ctl_woop_tri));
    const uint64_t n_idx = in.get<uint64_t>();
    if (n_idx != n_refs) throw std::runtime_error(std::string("Mesh file parser error (index count != reference count). ") + path);
    in.need(n_idx, 4);
    M.pre_index.resize((size_t)n_idx); in.read(M.pre_index.data(), (size_t)n_idx * 4);
    for (uint32_t w : M.pre_index) if ((w >> 1) >= n_tris) throw std::runtime_error(std::string("Mesh file parser error (leaf references a triangle out of range). ") + path);
    validate_mesh_bvh(M.pre_nodes.data(), (uint32_t)n_nodes, M.pre_index.data(), (uint32_t)n_refs, n_tris, std::string("Mesh file parser error: ") + path); // a damaged tree must not reach the GPU
    for (const ctl_tri_data& t : M.pre_tri_data) if (((t.w[1] >> 16) & 0xffu) >= n_mats) throw std::runtime_error(std::string("Mesh file parser error (triangle references a material out of range). ") + path);
}

void write_xmsh(const char* path, const SceneStorage& S, uint32_t mesh) {
    if (mesh >= S.meshes.size()) throw std::runtime_error("no such mesh");
    const ctl_mesh& km = S.meshes[mesh];
    const bool last = mesh + 1 == S.meshes.size();
    const uint32_t tri_end = last ? (uint32_t)S.tri_data.size() : S.meshes[mesh + 1].tri_offset;
    const uint32_t node_end = last ? (uint32_t)S.bvh_nodes.size() : S.meshes[mesh + 1].bvh_node_offset / 4;
    const uint32_t ref_end = last ? (uint32_t)S.tri_index.size() : S.meshes[mesh + 1].bvh_idx_offset;
    const uint32_t n_tris = tri_end - km.tri_offset, n_nodes = node_end - km.bvh_node_offset / 4, n_refs = ref_end - km.bvh_idx_offset;
    uint32_t mat_end = last ? (uint32_t)S.materials.size() : S.meshes[mesh + 1].mat_offset;
    const uint32_t n_mats = mat_end - km.mat_offset;
    const Box box = S.mesh_boxes.at(mesh);
    // emissive materials: the lights of the first node that instantiates this mesh
    std::vector<V3> emissive(n_mats, V3(0.0f));
    for (size_t ni = 0; ni < S.nodes.size(); ni++) {
        if (S.nodes[ni].mesh_index != mesh) continue;
        for (uint32_t m = 0; m < n_mats; m++) {
            const ctl_material& cm = S.materials[S.nodes[ni].material_offset + m];
            if (cm.node_light_index < 2 && cm.node_light_index < S.nodes[ni].n_lights) { const ctl_light& L = S.lights[S.nodes[ni].lights[cm.node_light_index]]; emissive[m] = V3(L.radiance[0], L.radiance[1], L.radiance[2]); }
        }
        break;
    }
    File out(path, "wb");
    out.put<uint32_t>(0);
    const float b6[6] = {box.lo.x, box.lo.y, box.lo.z, box.hi.x, box.hi.y, box.hi.z}; out.write(b6, sizeof(b6));
    uint32_t n_lights = 0; for (auto& e : emissive) if (e.x != 0 || e.y != 0 || e.z != 0) n_lights++;
    out.put<uint32_t>(n_lights);
    for (uint32_t m = 0; m < n_mats; m++) {
        const V3 e = emissive[m]; if (e.x == 0 && e.y == 0 && e.z == 0) continue;
        unsigned char rec[LIGHT_SIZE]; memset(rec, 0, sizeof(rec));
        put_fixed_string(rec, 32, "material_" + std::to_string(m)); put_f32(rec, LIGHT_L, e.x); put_f32(rec, LIGHT_L + 4, e.y); put_f32(rec, LIGHT_L + 8, e.z);
        out.write(rec, LIGHT_SIZE);
    }
    out.put<uint32_t>(n_tris); out.write(S.tri_data.data() + km.tri_offset, (size_t)n_tris * sizeof(ctl_tri_data));
    out.put<uint32_t>(n_mats);
    std::vector<unsigned char> blob(MATERIAL_SIZE);
    for (uint32_t m = 0; m < n_mats; m++) { encode_material(S.materials[km.mat_offset + m], "material_" + std::to_string(m), blob.data()); out.write(blob.data(), MATERIAL_SIZE); }
    out.put<uint64_t>(n_nodes); out.write(S.bvh_nodes.data() + km.bvh_node_offset / 4, (size_t)n_nodes * sizeof(ctl_bvh_node));
    out.put<uint64_t>(n_refs); out.write(S.woop.data() + km.bvh_tri_offset / 3, (size_t)n_refs * sizeof(ctl_woop_tri));
    // leaf words are mesh-local triangle indices already
    out.put<uint64_t>(n_refs); out.write(S.tri_index.data() + km.bvh_idx_offset, (size_t)n_refs * 4);
}

} // namespace ctlb
