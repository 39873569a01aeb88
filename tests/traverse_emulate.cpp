// tests/traverse_emulate.cpp -- TEST INFRASTRUCTURE (method: tests/nlm_emulate.cpp).  The device traversal source (cudatracerlib_b200/csrc/device/
// traverse.cuh: traverse_level, woop_test, trace_ray -- the arithmetic both traversal kernels share) compiled for the HOST and run ray by ray over a host
// scene view, with the argument set-up and result packing of k_intersect_simple's MODE 2 (== intersectKernel) and MODE 3 (== traceRay)
// (csrc/wavefront.cuh:150-199).  Hits, barycentrics and visit counts must equal the oracle's bit for bit: the slab and Woop FMAs are explicit in both.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cfloat>
#include <cmath>
#include <cstring>
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
#include "../cudatracerlib_b200/csrc/device/traverse.cuh"

using namespace ctld;

static DScene host_scene(const ctl_scene_view* v) {
    DScene S; memset(&S, 0, sizeof(S));
    S.scene_nodes = (const float4*)v->scene_bvh_nodes; S.bvh_nodes = (const float4*)v->bvh_nodes; S.woop = (const float4*)v->woop; S.tri_index = v->tri_index;
    S.tri_data = (const uint4*)v->tri_data; S.meshes = v->meshes; S.nodes = v->nodes; S.node_xf = (const float4*)v->node_xf; S.node_inv_xf = (const float4*)v->node_inv_xf;
    S.ray_eps = v->ray_eps; S.scene_start = v->scene_start_node; S.n_nodes = v->n_nodes;
    return S;
}

// MODE 3: ctl_trace_rays_host
extern "C" void emu_trace_rays(const ctl_scene_view* v, int n, const ctl_traversal_ray* rays, ctl_trace_result* out, unsigned long long counts[3]) {
    const DScene S = host_scene(v);
    VisitCounters<true> cnt;
    for (int i = 0; i < n; i++) {
        Hit hit; hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu; hit.dist = FLT_MAX;
        trace_ray<false, true>(S, mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]), S.ray_eps, 0.0f, hit, cnt);
        float* o = (float*)out + (size_t)i * 5;
        o[0] = hit.dist; o[1] = hit.u; o[2] = hit.v; memcpy(o + 3, &hit.tri, 4); memcpy(o + 4, &hit.node, 4);
    }
    counts[0] = cnt.inner; counts[1] = cnt.tris; counts[2] = cnt.inst;
}

// MODE 2: ctl_intersect / ctl_intersect_host
extern "C" void emu_intersect(const ctl_scene_view* v, int n, const ctl_traversal_ray* rays, ctl_traversal_result* out, int any_hit) {
    const DScene S = host_scene(v);
    VisitCounters<false> cnt;
    for (int i = 0; i < n; i++) {
        Hit hit; hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu; hit.dist = rays[i].tmax;
        const V3 o = mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        if (any_hit) trace_ray<true, false>(S, o, d, rays[i].tmin, rays[i].tmin, hit, cnt); else trace_ray<false, false>(S, o, d, rays[i].tmin, rays[i].tmin, hit, cnt);
        uint4 res = make_uint4(__float_as_uint(hit.dist), 0xffffffffu, 0xffffffffu, 0u);
        if (hit.tri != 0xffffffffu) {
            res.y = hit.node; res.z = hit.tri;
            const unsigned short xd = (unsigned short)(hit.u * 65535), yd = (unsigned short)(hit.v * 65535);
            res.w = ((uint32_t)yd << 16) | (uint32_t)xd;
        }
        memcpy(&out[i], &res, 16);
    }
}
