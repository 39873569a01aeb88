// staging.h -- host side of the staged traversal kernel's derived records (staging.cpp, device/traverse_staged.cuh)
#pragma once
#include <vector>
#include <string>
#include <cstdint>
#include "../../include/ctl_b200.h"

namespace ctlb {

struct StagedHost {
    std::vector<float> tri64;    // 16 floats per leaf slot
    std::vector<float> inst;     // 16 floats per (pseudo-)node
    std::vector<float> treelet;  // 16 floats per treelet node: the shared-memory image (swizzled chunks)
    int tl_nodes = 0;
    int scene_root = 0;          // node address the scene-level walk starts at
    uint32_t class_mask = 0;     // material classes present among the leaf slots (bit c = class c: 0 diffuse, 1 conductor / Beckmann, 2 conductor / GGX, 3 dielectric)
    bool class_ok = false;       // the per-slot classes hold for every instance (no node overrides its mesh's material block)
    bool usable = false;         // false: the view's leaf arrays do not have the one-slot-one-Woop-record shape (why says so); the persistent kernel is used
    std::string why;
};

// Mesh-level half (leaf triangles): changes only when meshes change.
void build_staging_tris(const ctl_scene_view& v, StagedHost& out);
// Node-level half (instance records, treelet of at most `treelet_budget` nodes): changes when instances move or the scene level is re-braided.
void build_staging_nodes(const ctl_scene_view& v, int treelet_budget, StagedHost& out);

} // namespace ctlb
