// obj_import.cpp -- Wavefront OBJ / MTL import ("then OBJ -> xmsh"): the front end of the reference's mesh pipeline
// (Engine/MeshLoader/ObjParser.cpp compileobj -> Mesh::CompileMesh) for the subset of materials the hot path has.  Produces a MeshInput that
// assemble_scene turns into exactly what the reference's compiler writes into an .xmsh for the same file (tests compare the two bit for bit):
//
//   * statements: v, vt (stored as (u, 1 - v)), vn, f (polygons are fanned from their first vertex), usemtl, mtllib; every other statement
//     is skipped like the reference does (ObjParser.cpp:595-782);
//   * vertices are the distinct (v, vt, vn) index triples in order of first use (ObjParser.cpp:644-690); absent or out-of-range slots = none;
//   * one sub-mesh (= material slot) per material in order of its first `usemtl`; faces before any `usemtl` form a default sub-mesh;
//     triangles are emitted sub-mesh by sub-mesh with REVERSED winding (ObjParser.cpp:857-862);
//   * numbers are read by digit accumulation in single precision (value = value * 10 + digit; fraction digits scaled by repeated * 0.1f), the
//     reference's reader -- not correctly rounded, and vertex positions must match it to the bit for the Woop / TriangleData records to match;
//   * MTL: Kd, Ks, Ke, Ns, Ni, Tf, d, illum; illum 2 with Ks = 0 -> diffuse(Kd); illum 7 -> dielectric(Ni, reflectance Ks, transmittance Tf);
//     illum 9 -> dielectric(Ni, reflectance 0, transmittance Tf); Ke != 0 -> area light.  illum 2 with specular (phong), illum 5 (smooth
//     conductor) and texture maps are outside the B200 path and are rejected with a message naming the material.
#include "scene_builder.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>

namespace ctlb {
namespace {

struct Cursor {
    const char* p;
    void space() { while (*p == ' ' || *p == '\t') p++; }
    bool lit(const char* s) { const char* t = p; while (*s && *t == *s) { t++; s++; } if (*s) return false; p = t; return true; }
    bool integer(int& out) {
        const char* t = p; bool neg = false;
        if (*t == '+') t++; else if (*t == '-') { neg = true; t++; }
        if (*t < '0' || *t > '9') return false;
        int v = 0; while (*t >= '0' && *t <= '9') v = v * 10 + (*t++ - '0');
        out = neg ? -v : v; p = t; return true;
    }
    // single-precision digit accumulation (see header): integer digits, fraction digits, optional exponent
    bool real(float& out) {
        const char* t = p; bool neg = false;
        if (*t == '+') t++; else if (*t == '-') { neg = true; t++; }
        float v = 0.0f; int digits = 0;
        for (; *t >= '0' && *t <= '9'; t++, digits++) v = v * 10.0f + (float)(*t - '0');
        if (*t == '.') { t++; float scale = 1.0f; for (; *t >= '0' && *t <= '9'; t++, digits++) { scale *= 0.1f; v += scale * (float)(*t - '0'); } }
        if (!digits) return false;
        p = t;
        if (*t == 'e' || *t == 'E') { Cursor c{t + 1}; int e = 0; if (c.integer(e)) { p = c.p; if (e) v *= powf(10.0f, (float)e); } }
        out = neg ? -v : v; return true;
    }
    bool reals(float* out, int n) { Cursor c{p}; for (int i = 0; i < n; i++) { if (i) c.space(); if (!c.real(out[i])) return false; } p = c.p; return true; }
};

std::string trimmed(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

struct ObjMat { // defaults of ObjParser.cpp:232-243
    std::string name; int illum = 2; float kd[3] = {0.75f, 0.75f, 0.75f}, ks[3] = {0.5f, 0.5f, 0.5f}, ke[3] = {0, 0, 0}, tf[3] = {1, 1, 1}, ni = 1.0f, ns = 32.0f;
    bool has_map = false; int submesh = -1;
};

bool read_lines(const std::string& path, std::vector<std::string>& lines) {
    FILE* f = fopen(path.c_str(), "rb"); if (!f) return false;
    std::string cur; int c;
    while ((c = fgetc(f)) != EOF) { if (c == '\n') { lines.push_back(cur); cur.clear(); } else cur.push_back((char)c); }
    if (!cur.empty()) lines.push_back(cur);
    fclose(f); return true;
}

void load_mtl(const std::string& path, std::vector<ObjMat>& mats) {
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) throw std::runtime_error("Could not open file: " + path);
    ObjMat cur; bool have = false;
    auto flush = [&]() { if (!have) return; for (auto& m : mats) if (m.name == cur.name) return; mats.push_back(cur); };
    for (const std::string& raw : lines) {
        const std::string line = trimmed(raw);
        Cursor c{line.c_str()}; c.space();
        if (!*c.p || c.lit("#")) continue;
        if (c.lit("newmtl ")) { c.space(); if (*c.p) { flush(); cur = ObjMat(); cur.name = c.p; have = true; } }
        else if (c.lit("Kd ")) { c.space(); c.reals(cur.kd, 3); }
        else if (c.lit("Ks ")) { c.space(); c.reals(cur.ks, 3); }
        else if (c.lit("Ke ")) { c.space(); c.reals(cur.ke, 3); }
        else if (c.lit("Tf ")) { c.space(); c.reals(cur.tf, 3); }
        else if (c.lit("Ni ")) { c.space(); c.real(cur.ni); }
        else if (c.lit("Ns ")) { c.space(); c.real(cur.ns); if (cur.ns <= 0.0f) { cur.ns = 1.0f; cur.ks[0] = cur.ks[1] = cur.ks[2] = 0.0f; } } // ObjParser.cpp:488-492
        else if (c.lit("illum ")) { c.space(); c.integer(cur.illum); }
        else if (c.lit("map_") || c.lit("disp ") || c.lit("bump ") || c.lit("refl ")) cur.has_map = true;
    }
    flush();
}

} // namespace

void read_obj(const char* path_c, MeshInput& M) {
    const std::string path(path_c);
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) throw std::runtime_error("Could not open file: " + path);
    const size_t slash = path.find_last_of("/\\");
    const std::string dir = slash == std::string::npos ? std::string(".") : path.substr(0, slash);

    std::vector<V3> positions, normals; std::vector<float> texcoords; // 2 per vt
    std::vector<ObjMat> mats;
    struct Sub { int material; std::vector<uint32_t> tris; };          // material: index into mats, -1 = default sub-mesh
    std::vector<Sub> subs;
    std::map<std::tuple<int, int, int>, uint32_t> vertex_of;           // (v, vt, vn) -> vertex
    std::vector<V3> vp, vn; std::vector<float> vt;                     // per vertex
    std::vector<uint32_t> pending, face;                               // triangles since the last material switch; current polygon
    int submesh = -1, default_submesh = -1;
    bool any_vn = false;

    for (const std::string& raw : lines) {
        const std::string line = trimmed(raw);
        Cursor c{line.c_str()}; c.space();
        if (!*c.p || c.lit("#")) continue;
        if (c.lit("v ")) { c.space(); float v[3]; if (c.reals(v, 3)) { c.space(); if (!*c.p) positions.push_back(V3(v[0], v[1], v[2])); } }
        else if (c.lit("vt ")) { c.space(); float v[2]; if (c.reals(v, 2)) { texcoords.push_back(v[0]); texcoords.push_back(1.0f - v[1]); } }
        else if (c.lit("vn ")) { c.space(); float v[3]; if (c.reals(v, 3)) { c.space(); if (!*c.p) { normals.push_back(V3(v[0], v[1], v[2])); any_vn = true; } } }
        else if (c.lit("f ")) {
            c.space(); face.clear();
            while (*c.p) {
                int idx[3] = {0, 0, 0};
                if (!c.integer(idx[0])) break;
                for (int i = 1; i < 4 && c.lit("/"); i++) { int tmp = 0; c.integer(tmp); if (i < 3) idx[i] = tmp; }
                c.space();
                const int size[3] = {(int)positions.size(), (int)texcoords.size() / 2, (int)normals.size()};
                for (int i = 0; i < 3; i++) { if (idx[i] < 0) idx[i] += size[i]; else idx[i]--; if (idx[i] < 0 || idx[i] >= size[i]) idx[i] = -1; }
                const auto key = std::make_tuple(idx[0], idx[1], idx[2]);
                auto it = vertex_of.find(key);
                if (it == vertex_of.end()) {
                    it = vertex_of.emplace(key, (uint32_t)vp.size()).first;
                    vp.push_back(idx[0] < 0 ? V3(0.0f) : positions[idx[0]]);
                    vt.push_back(idx[1] < 0 ? 0.0f : texcoords[2 * idx[1]]); vt.push_back(idx[1] < 0 ? 0.0f : texcoords[2 * idx[1] + 1]);
                    vn.push_back(idx[2] < 0 ? V3(0.0f) : normals[idx[2]]);
                }
                face.push_back(it->second);
            }
            if (!*c.p) {
                if (submesh == -1) { if (default_submesh == -1) { default_submesh = (int)subs.size(); subs.push_back({-1, {}}); } submesh = default_submesh; }
                for (size_t i = 2; i < face.size(); i++) { pending.push_back(face[0]); pending.push_back(face[i - 1]); pending.push_back(face[i]); }
            }
        }
        else if (c.lit("usemtl ")) {
            c.space();
            const std::string name(c.p);
            if (submesh != -1) { subs[submesh].tris.insert(subs[submesh].tris.end(), pending.begin(), pending.end()); pending.clear(); submesh = -1; }
            for (size_t m = 0; m < mats.size(); m++) if (mats[m].name == name) {
                if (mats[m].submesh == -1) { mats[m].submesh = (int)subs.size(); subs.push_back({(int)m, {}}); }
                submesh = mats[m].submesh; pending.clear(); break;
            }
        }
        else if (c.lit("mtllib ")) { c.space(); if (*c.p) load_mtl(dir + "/" + trimmed(c.p), mats); }
    }
    if (submesh != -1) subs[submesh].tris.insert(subs[submesh].tris.end(), pending.begin(), pending.end());
    if (subs.empty()) throw std::runtime_error("Invalid obj file, did not find submeshes!");
    if (subs.size() > 255) throw std::runtime_error("more than 255 materials in one mesh (8-bit index, TriangleData.h:24): " + path);

    M = MeshInput();
    M.verts = vp; M.uvs = vt;
    if (any_vn) M.normals = vn;   // Mesh::CompileMesh takes the file's normals only when the file has vn statements (ObjParser.cpp:870)
    for (size_t s = 0; s < subs.size(); s++) {
        ObjMat om; if (subs[s].material >= 0) om = mats[subs[s].material]; else om.name = "default";
        ctl_material cm; memset(&cm, 0, sizeof(cm));
        cm.node_light_index = 0xffffffffu; cm.alpha_u = cm.alpha_v = 0.1f; cm.eta[0] = cm.eta[1] = cm.eta[2] = 1.5f; cm.transmittance = 1.0f;
        if (om.has_map) throw std::runtime_error("material '" + om.name + "': texture maps are outside the B200 path");
        if (om.illum == 2) {
            if (om.ks[0] != 0 || om.ks[1] != 0 || om.ks[2] != 0) throw std::runtime_error("material '" + om.name + "': illum 2 with a specular colour compiles to a phong BSDF, which is not supported by the B200 path (set Ks 0 0 0 for diffuse)");
            cm.bsdf_type = CTL_BSDF_DIFFUSE; memcpy(cm.reflectance, om.kd, 12);
        } else if (om.illum == 7 || om.illum == 9) {
            if (om.tf[0] != om.tf[1] || om.tf[0] != om.tf[2]) throw std::runtime_error("material '" + om.name + "': coloured transmittance (Tf) is outside the B200 path");
            cm.bsdf_type = CTL_BSDF_DIELECTRIC; cm.eta[0] = cm.eta[1] = cm.eta[2] = om.ni; cm.transmittance = om.tf[0];
            if (om.illum == 7) memcpy(cm.reflectance, om.ks, 12);
        } else throw std::runtime_error("material '" + om.name + "': illum " + std::to_string(om.illum) + " is not supported by the B200 path (2 diffuse, 7 / 9 dielectric)");
        M.materials.push_back(cm);
        M.emissive.push_back(V3(om.ke[0], om.ke[1], om.ke[2]));
        for (size_t t = 0; t + 2 < subs[s].tris.size(); t += 3) {
            M.indices.push_back(subs[s].tris[t + 2]); M.indices.push_back(subs[s].tris[t + 1]); M.indices.push_back(subs[s].tris[t]); // reversed winding
            M.mat_index.push_back((uint8_t)s);
        }
    }
    if (M.indices.empty()) throw std::runtime_error("Invalid obj file, no faces: " + path);
}

// ---- PLY (Engine/MeshLoader/PlyParser.cpp compileply): ascii and binary (little / big endian) files with float x y z vertices and triangle or
// quad faces; one red diffuse default material (PlyParser.cpp:364-368), vertex normals computed.  Kept quirks of the reference reader, needed for
// identical output: triangles are emitted with reversed winding but quads as (2,3,0),(0,1,2); a `u` property is used for both texture
// coordinates (:283); the binary branch reads the vertex block as consecutive x y z floats, i.e. it supports position-only vertices;
// binary indices beyond the vertex count become 0.
void read_ply(const char* path_c, MeshInput& M) {
    const std::string path(path_c);
    FILE* f = fopen(path.c_str(), "rb"); if (!f) throw std::runtime_error("Could not open file: " + path);
    std::vector<unsigned char> data; { unsigned char buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n); }
    fclose(f);
    size_t pos = 0;
    auto getline = [&](std::string& line) { if (pos >= data.size()) return false; line.clear(); while (pos < data.size() && data[pos] != '\n') line.push_back((char)data[pos++]); if (pos < data.size()) pos++; if (!line.empty() && line.back() == '\r') line.pop_back(); return true; };
    auto words = [](const std::string& line) { std::vector<std::string> w; size_t i = 0; while (i < line.size()) { while (i < line.size() && isspace((unsigned char)line[i])) i++; size_t j = i; while (j < line.size() && !isspace((unsigned char)line[j])) j++; if (j > i) w.push_back(line.substr(i, j - i)); i = j; } return w; };
    std::string line;
    if (!getline(line) || line.substr(0, 3) != "ply") throw std::runtime_error("not a ply file: " + path);
    int format = -1, vertex_count = -1, face_count = -1, n_props = 0, pos_start = -1, uv_start = -1, has_pos = 0; bool has_uv = false, in_vertex = false;
    int pos_col[3] = {-1, -1, -1}; bool pos_is_float32 = true;
    while (getline(line)) {
        const std::vector<std::string> w = words(line);
        if (w.empty()) continue;
        if (w[0] == "format" && w.size() >= 2) format = w[1] == "ascii" ? 2 : (w[1] == "binary_big_endian" ? 1 : (w[1] == "binary_little_endian" ? 0 : -1));
        else if (w[0] == "element" && w.size() >= 3) { in_vertex = false; if (w[1] == "vertex") { vertex_count = atoi(w[2].c_str()); in_vertex = true; n_props = 0; } else if (w[1] == "face") face_count = atoi(w[2].c_str()); }
        else if (w[0] == "property" && in_vertex) {
            if (w.size() < 3 || w[1] == "list") throw std::runtime_error("compileply: unsupported vertex property: " + path);
            const std::string &type = w[1], &name = w[2];
            const bool real = type == "float" || type == "double";
            if (name.size() == 1 && name[0] >= 'x' && name[0] <= 'z' && real) { has_pos |= 1 << (name[0] - 'x'); pos_col[name[0] - 'x'] = n_props; if (type != "float") pos_is_float32 = false; if (name[0] == 'x') pos_start = n_props; }
            else if (name[0] == 'u' || (name[0] == 'v' && name.size() == 1 && real)) { has_uv = true; if (name[0] == 'u') uv_start = n_props; }
            else if (name.find("material") != std::string::npos) throw std::runtime_error("compileply: per-vertex materials are not supported: " + path);
            n_props++;
        }
        else if (w[0] == "end_header") break;
    }
    if (has_pos != 7 || vertex_count <= 0 || face_count <= 0 || format < 0) throw std::runtime_error("compileply: header without float x y z vertices and faces: " + path);
    // the reference reader takes x, y, z from three consecutive columns starting at x (PlyParser.cpp:262-283); anything else would read the wrong columns
    if (pos_col[1] != pos_col[0] + 1 || pos_col[2] != pos_col[0] + 2) throw std::runtime_error("compileply: unsupported vertex layout (x y z must be consecutive properties): " + path);
    if (format != 2 && (n_props != 3 || pos_start != 0 || !pos_is_float32)) throw std::runtime_error("compileply: unsupported vertex layout (binary files: exactly the float properties x y z): " + path);
    if (has_uv && (uv_start < 0 || uv_start >= n_props)) has_uv = false;
    M = MeshInput();
    M.verts.resize(vertex_count);
    if (has_uv) M.uvs.assign((size_t)vertex_count * 2, 0.0f);
    auto emit = [&](const uint32_t* v, int n) {
        if (n == 3) { M.indices.push_back(v[2]); M.indices.push_back(v[1]); M.indices.push_back(v[0]); }
        else if (n == 4) { for (int i = 2; i < 5; i++) M.indices.push_back(v[i % 4]); for (int i = 0; i < 3; i++) M.indices.push_back(v[i]); }
        else throw std::runtime_error("compileply: faces must be triangles or quads: " + path);
    };
    if (format == 2) {
        for (int v = 0; v < vertex_count; v++) {
            if (!getline(line)) throw std::runtime_error("Passed end of file: " + path);
            const std::vector<std::string> w = words(line);
            if ((int)w.size() < n_props) throw std::runtime_error("compileply: short vertex line: " + path);
            auto val = [&](int i) { return (float)atof(w[i].c_str()); };
            M.verts[v] = V3(val(pos_start), val(pos_start + 1), val(pos_start + 2));
            if (has_uv && uv_start >= 0) { M.uvs[2 * v] = val(uv_start); M.uvs[2 * v + 1] = val(uv_start); }
        }
        for (int fc = 0; fc < face_count; fc++) {
            if (!getline(line)) throw std::runtime_error("Passed end of file: " + path);
            const std::vector<std::string> w = words(line);
            const int n = w.empty() ? 0 : atoi(w[0].c_str());
            if ((int)w.size() < n + 1 || n > 4) throw std::runtime_error("compileply: faces must be triangles or quads: " + path);
            uint32_t idx[4] = {0, 0, 0, 0}; for (int i = 0; i < n; i++) idx[i] = (uint32_t)atoi(w[1 + i].c_str());
            emit(idx, n);
        }
    } else {
        auto swap32 = [](uint32_t i) { return (i << 24) | ((i << 8) & 0xff0000u) | ((i >> 8) & 0xff00u) | (i >> 24); };
        if (pos + (size_t)vertex_count * 12 > data.size()) throw std::runtime_error("Passed end of file: " + path);
        for (int v = 0; v < vertex_count; v++) {
            uint32_t b[3]; memcpy(b, &data[pos + (size_t)v * 12], 12);
            if (format == 1) for (int k = 0; k < 3; k++) b[k] = swap32(b[k]);
            float xyz[3]; memcpy(xyz, b, 12); M.verts[v] = V3(xyz[0], xyz[1], xyz[2]);
        }
        pos += (size_t)vertex_count * 12;
        for (int fc = 0; fc < face_count; fc++) {
            if (pos >= data.size()) throw std::runtime_error("Passed end of file: " + path);
            const int n = data[pos];
            if ((n != 3 && n != 4) || pos + 1 + 4 * (size_t)n > data.size()) throw std::runtime_error("compileply: faces must be triangles or quads: " + path);
            uint32_t idx[4] = {0, 0, 0, 0};
            for (int i = 0; i < n; i++) { memcpy(&idx[i], &data[pos + 1 + 4 * i], 4); if (format == 1) idx[i] = swap32(idx[i]); }
            const size_t first = M.indices.size();
            emit(idx, n);
            for (size_t i = M.indices.size() - n; i < M.indices.size(); i++) if (M.indices[i] > (uint32_t)vertex_count) M.indices[i] = 0;   // PlyParser.cpp:349-350 (the last n indices)
            (void)first;
            pos += 4 * (size_t)n + 1;
        }
    }
    for (uint32_t i : M.indices) if (i >= (uint32_t)vertex_count) throw std::runtime_error("compileply: vertex index out of range: " + path);
    M.mat_index.assign(M.indices.size() / 3, 0);
    ctl_material cm; memset(&cm, 0, sizeof(cm));
    cm.bsdf_type = CTL_BSDF_DIFFUSE; cm.node_light_index = 0xffffffffu; cm.reflectance[0] = 1.0f; cm.alpha_u = cm.alpha_v = 0.1f; cm.eta[0] = cm.eta[1] = cm.eta[2] = 1.5f; cm.transmittance = 1.0f;
    M.materials.push_back(cm); M.emissive.push_back(V3(0.0f));
}

} // namespace ctlb
