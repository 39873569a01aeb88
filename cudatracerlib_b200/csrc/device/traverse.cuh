// traverse.cuh -- two-level BVH traversal + Woop triangle test on the reference data surface.
//
// Replaces Kernel/TraceHelper.cu:88-180 (traceRay / __traceRay_internal__), :326-734 (intersectKernel)
// and Engine/SpatialStructures/BVH/BVHTraversal.h:8-232 (TracerayTemplate).  Same node layout
// (64 B, children in float4 units, ~leaf, 0x76543210 sentinel), same child-order rule
// (swp = c1min < c0min, far child pushed), same Woop test; per-thread "leaf as soon as met" order,
// i.e. the order of the reference's host branch, so visit counts equal the CPU oracle's.
#pragma once
#include "dmath.cuh"
#include "../../../include/ctl_b200.h"

namespace ctld {

// Device view of the scene (device pointers). Passed to kernels by value (__grid_constant__).
struct DScene {
    const float4* scene_nodes;   // KernelSceneBVH::m_pNodes
    const float4* bvh_nodes;     // m_sBVHNodeData
    const float4* woop;          // m_sBVHIntData
    const uint32_t* tri_index;   // m_sBVHIndexData
    const uint4* tri_data;       // m_sTriData (2 x uint4 per triangle)
    const ctl_mesh* meshes;
    const ctl_node* nodes;
    const float4* node_xf;
    const float4* node_inv_xf;
    const ctl_material* materials;
    const ctl_light* lights;
    const ctl_light_tri* light_tris;
    const float* light_cdf_data;
    const float* normal_lut;     // [0..255] sin(theta) [256..] cos(theta) [512..] sin(phi) [768..] cos(phi)
    const float* d1;             // SequenceSamplerData 1-D table
    const float2* d2;            // 2-D table
    uint32_t num_lights;
    uint32_t light_indices[CTL_MAX_NUM_LIGHTS];
    float light_cdf[CTL_MAX_NUM_LIGHTS];
    ctl_camera camera;
    float ray_eps;
    float box_min[3], box_inv_extent[3]; // scene box (m_sBox) for the ray-sort keys: cell = (o - min) * inv_extent
    int scene_start;
    uint32_t n_nodes;
    int img_w, img_h;
};

struct Hit { float dist, u, v; uint32_t tri, node; };

CTL_DEV float guard_inv(float d) { // BVHTraversal.h:16-19
    const float ooeps = 0x1p-80f;
    return 1.0f / (fabsf(d) > ooeps ? d : copysignf(ooeps, d));
}

constexpr int SENT = CTL_SENTINEL;
constexpr int TRI_CLS_SHIFT = 29;                 // wavefront hit records written by the staged kernel carry the hit triangle's material class in the top 3 bits of the triangle word (7 = miss)
constexpr uint32_t TRI_IDX_MASK = 0x1fffffffu;
constexpr int STACK_N = 64;

template <bool COUNT> struct VisitCounters { };
template <> struct VisitCounters<true> { unsigned inner = 0, tris = 0, inst = 0; };

// One BVH level.  LEAF(int leaf_payload) -> void; `stop` lets any-hit abort.
template <bool COUNT, typename LEAF>
CTL_DEV void traverse_level(const float4* __restrict__ nodes, int start, V3 o, V3 d, float box_lo, const float& rayT,
                            const bool& stop, VisitCounters<COUNT>& cnt, LEAF leaf) {
    if (start < 0) { leaf(~start); return; }
    int stack[STACK_N];
    int sp = 0;
    stack[0] = SENT;
    const float idx = guard_inv(d.x), idy = guard_inv(d.y), idz = guard_inv(d.z);
    const float oodx = o.x * idx, oody = o.y * idy, oodz = o.z * idz;
    int nodeAddr = start;
    while (nodeAddr != SENT) {
        int leafAddr = 0;
        while ((unsigned)nodeAddr < (unsigned)SENT) {
            const F8 nA = ldg256(nodes + nodeAddr), nB = ldg256(nodes + nodeAddr + 2);
                const float4 n0xy = nA.lo, n1xy = nA.hi, nz = nB.lo, cn = nB.hi;
            if (COUNT) ((VisitCounters<true>&)cnt).inner++;
            int c0 = __float_as_int(cn.x), c1 = __float_as_int(cn.y);
            const float c0lox = fmaf(n0xy.x, idx, -oodx), c0hix = fmaf(n0xy.y, idx, -oodx);
            const float c0loy = fmaf(n0xy.z, idy, -oody), c0hiy = fmaf(n0xy.w, idy, -oody);
            const float c0loz = fmaf(nz.x, idz, -oodz), c0hiz = fmaf(nz.y, idz, -oodz);
            const float c1loz = fmaf(nz.z, idz, -oodz), c1hiz = fmaf(nz.w, idz, -oodz);
            const float c1lox = fmaf(n1xy.x, idx, -oodx), c1hix = fmaf(n1xy.y, idx, -oodx);
            const float c1loy = fmaf(n1xy.z, idy, -oody), c1hiy = fmaf(n1xy.w, idy, -oody);
            // spanBegin/EndKepler (MathFunc.h:443-444) == float min/max chains for t >= 0
            const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), box_lo));
            const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rayT));
            const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), box_lo));
            const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rayT));
            const bool swp = (c1min < c0min), t0 = (c0max >= c0min), t1 = (c1max >= c1min);
            if (!t0 && !t1) { nodeAddr = stack[sp]; sp--; }
            else {
                nodeAddr = t0 ? c0 : c1;
                if (t0 && t1) {
                    if (swp) { int tmp = nodeAddr; nodeAddr = c1; c1 = tmp; }
                    sp++; stack[sp] = c1;
                }
            }
            if (nodeAddr < 0) { leafAddr = nodeAddr; nodeAddr = stack[sp]; sp--; break; }
        }
        while (leafAddr < 0) {
            leaf(~leafAddr);
            if (stop) return;
            leafAddr = nodeAddr;
            if (nodeAddr < 0) { nodeAddr = stack[sp]; sp--; }
        }
    }
}

// Woop unit-triangle test (TraceHelper.cu:118-134 / 646-672) with explicit FMAs.
CTL_DEV bool woop_test(const float4 a, const float4 b, const float4 c, V3 o, V3 d, float tlo, float thi, float& t, float& u, float& v) {
    const float Oz = fmaf(-o.z, a.z, fmaf(-o.y, a.y, fmaf(-o.x, a.x, a.w)));
    const float invDz = 1.0f / fmaf(d.z, a.z, fmaf(d.y, a.y, d.x * a.x));
    t = Oz * invDz;
    if (t > tlo && t < thi) {
        const float Ox = fmaf(o.z, b.z, fmaf(o.y, b.y, fmaf(o.x, b.x, b.w)));
        const float Dx = fmaf(d.z, b.z, fmaf(d.y, b.y, d.x * b.x));
        u = fmaf(t, Dx, Ox);
        if (u >= 0.0f) {
            const float Oy = fmaf(o.z, c.z, fmaf(o.y, c.y, fmaf(o.x, c.x, c.w)));
            const float Dy = fmaf(d.z, c.z, fmaf(d.y, c.y, d.x * c.x));
            v = fmaf(t, Dy, Oy);
            if (v >= 0.0f && u + v <= 1.0f) return true;
        }
    }
    return false;
}

// Two-level query.  hit.dist must hold the upper bound on entry (FLT_MAX or ray.tmax); hit.tri = UINT_MAX.
// tri_lo: lower t bound for triangles; box_lo: lower bound for box entry.
template <bool ANY_HIT, bool COUNT>
CTL_DEV void trace_ray(const DScene& S, V3 ori, V3 dir, float tri_lo, float box_lo, Hit& hit, VisitCounters<COUNT>& cnt) {
    if (!S.n_nodes) return;
    bool stop = false;
    traverse_level<COUNT>(S.scene_nodes, S.scene_start, ori, dir, box_lo, hit.dist, stop, cnt, [&](int nodeIdx) {
        if (COUNT) ((VisitCounters<true>&)cnt).inst++;
        const ctl_node* N = S.nodes + nodeIdx;
        const uint32_t mesh_index = __ldg(&N->mesh_index);
        const ctl_mesh* M = S.meshes + mesh_index;
        const uint32_t node_off = __ldg(&M->bvh_node_offset), tri_off4 = __ldg(&M->bvh_tri_offset), idx_off = __ldg(&M->bvh_idx_offset), tri_base = __ldg(&M->tri_offset);
        const float4* inv = S.node_inv_xf + (size_t)nodeIdx * 4;
        const V3 d = xf_dir(inv, dir), o = xf_point(inv, ori);
        const float4* woop = S.woop + tri_off4;
        const uint32_t* tidx = S.tri_index + idx_off;
        traverse_level<COUNT>(S.bvh_nodes + node_off, 0, o, d, box_lo, hit.dist, stop, cnt, [&](int triIdx) {
            for (int triAddr = triIdx;; triAddr++) {
                const float4 v00 = __ldg(woop + triAddr * 3 + 0);
                const float4 v11 = __ldg(woop + triAddr * 3 + 1);
                const float4 v22 = __ldg(woop + triAddr * 3 + 2);
                const uint32_t index = __ldg(tidx + triAddr);
                if (COUNT) ((VisitCounters<true>&)cnt).tris++;
                float t, u, v;
                if (woop_test(v00, v11, v22, o, d, tri_lo, hit.dist, t, u, v)) {
                    hit.node = (uint32_t)nodeIdx; hit.tri = (index >> 1) + tri_base; hit.u = u; hit.v = v; hit.dist = t;
                    if (ANY_HIT) { stop = true; break; }
                }
                if (index & 1) break;
            }
        });
    });
}

} // namespace ctld
