// tests/staged_emulate.cpp -- TEST INFRASTRUCTURE (method: tests/traverse_emulate.cpp).  One lane of the staged traversal kernel
// (cudatracerlib_b200/csrc/device/traverse_staged.cuh: trace_staged) walked on the host over the derived records that
// csrc/staging.cpp builds from a scene view: 64-byte leaf-triangle records, 64-byte instance records, the swizzled treelet image with
// re-addressed children, the stack with its top in a register, `stack_rows` rows in "shared memory" and the local-memory overflow.
// The shared arithmetic (slab test constants, woop_test, guard_inv, dot4) is the device source itself, compiled for the host.
// Hits, barycentrics and visit counts must equal the oracle's bit for bit, for every treelet budget and stack split.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
#include "../cudatracerlib_b200/csrc/device/traverse.cuh"
#include "../cudatracerlib_b200/csrc/staging.h"

using namespace ctld;

namespace {
constexpr int kTL = 1, kNeedsW = 2, kStack = 64;
inline int tl_chunk(int n, int j) { return n * 4 + (j ^ ((n >> 1) & 3)); }

struct Walker {
    const ctl_scene_view* v; const ctlb::StagedHost* H; int SD;
    unsigned long long inner = 0, tris = 0, insts = 0;
    int max_sp = 0;

    // MODE 2 (tmin / tmax from the ray) when api2, else MODE 3 (rayEps, FLT_MAX)
    void trace(const ctl_traversal_ray& R, bool api2, bool any_hit, Hit& hit) {
        const float4* tl = (const float4*)H->treelet.data();
        const float4* tri64 = (const float4*)H->tri64.data();
        const float4* instr = (const float4*)H->inst.data();
        std::vector<int> rows(SD + 1, 0); int ovf[kStack];
        int sp = 0, tos = SENT;
        auto push = [&](int x) { sp++; if (sp > max_sp) max_sp = sp; if (sp <= SD) rows[sp] = tos; else ovf[sp - SD - 1] = tos; tos = x; };
        auto pop = [&]() { const int r = tos; tos = sp <= SD ? rows[sp] : ovf[sp - SD - 1]; sp--; return r; };
        float ox = R.o[0], oy = R.o[1], oz = R.o[2], dx = R.d[0], dy = R.d[1], dz = R.d[2];
        float tri_lo, box_lo;
        hit.u = hit.v = 0.0f; hit.tri = hit.node = 0xffffffffu;
        if (api2) { tri_lo = R.tmin; box_lo = R.tmin; hit.dist = R.tmax; } else { tri_lo = v->ray_eps; box_lo = 0.0f; hit.dist = FLT_MAX; }
        int inst = -1, triAddr = 0; uint32_t tri_slot_base = 0, tri_base = 0;
        const float4* nbase = (const float4*)v->scene_bvh_nodes;
        int nodeAddr = v->n_nodes ? H->scene_root : SENT;
        float idx = 0, idy = 0, idz = 0, oodx = 0, oody = 0, oodz = 0;
        if (nodeAddr >= 0) { idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz); oodx = ox * idx; oody = oy * idy; oodz = oz * idz; }
        for (;;) {
            const int state = ((unsigned)nodeAddr < (unsigned)SENT) ? 0 : (nodeAddr < 0 ? (inst >= 0 ? 1 : 2) : (inst >= 0 ? 2 : 3));
            if (state == 3) return;
            if (state == 2) {
                if (nodeAddr < 0) {
                    const int nodeIdx = ~nodeAddr; insts++;
                    const float4* I = instr + (size_t)nodeIdx * 4;
                    const float4 r0 = I[0], r1 = I[1], r2 = I[2], meta = I[3];
                    const uint32_t root = __float_as_uint(meta.w);
                    const float ddx = dot4(r0, dx, dy, dz, 0.0f), ddy = dot4(r1, dx, dy, dz, 0.0f), ddz = dot4(r2, dx, dy, dz, 0.0f);
                    float px = dot4(r0, ox, oy, oz, 1.0f), py = dot4(r1, ox, oy, oz, 1.0f), pz = dot4(r2, ox, oy, oz, 1.0f);
                    if (root & kNeedsW) { const float4 r3 = ((const float4*)v->node_inv_xf)[(size_t)nodeIdx * 4 + 3]; const float w = dot4(r3, ox, oy, oz, 1.0f); px = px / w; py = py / w; pz = pz / w; }
                    ox = px; oy = py; oz = pz; dx = ddx; dy = ddy; dz = ddz;
                    nbase = (const float4*)v->bvh_nodes + __float_as_uint(meta.x); tri_slot_base = __float_as_uint(meta.y); tri_base = __float_as_uint(meta.z);
                    inst = nodeIdx;
                    push(SENT);
                    nodeAddr = (int)(root & ~(uint32_t)kNeedsW);
                } else {
                    ox = R.o[0]; oy = R.o[1]; oz = R.o[2]; dx = R.d[0]; dy = R.d[1]; dz = R.d[2];
                    nbase = (const float4*)v->scene_bvh_nodes; inst = -1;
                    nodeAddr = pop();
                }
                idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz); oodx = ox * idx; oody = oy * idy; oodz = oz * idz;
            } else if (state == 0) {
                float4 n0xy, n1xy, nz, cn;
                if (nodeAddr & kTL) { const int t = nodeAddr >> 2; n0xy = tl[tl_chunk(t, 0)]; n1xy = tl[tl_chunk(t, 1)]; nz = tl[tl_chunk(t, 2)]; cn = tl[tl_chunk(t, 3)]; }
                else { n0xy = nbase[nodeAddr]; n1xy = nbase[nodeAddr + 1]; nz = nbase[nodeAddr + 2]; cn = nbase[nodeAddr + 3]; }
                inner++;
                int c0 = __float_as_int(cn.x), c1 = __float_as_int(cn.y);
                const float c0lox = fmaf(n0xy.x, idx, -oodx), c0hix = fmaf(n0xy.y, idx, -oodx);
                const float c0loy = fmaf(n0xy.z, idy, -oody), c0hiy = fmaf(n0xy.w, idy, -oody);
                const float c0loz = fmaf(nz.x, idz, -oodz), c0hiz = fmaf(nz.y, idz, -oodz);
                const float c1loz = fmaf(nz.z, idz, -oodz), c1hiz = fmaf(nz.w, idz, -oodz);
                const float c1lox = fmaf(n1xy.x, idx, -oodx), c1hix = fmaf(n1xy.y, idx, -oodx);
                const float c1loy = fmaf(n1xy.z, idy, -oody), c1hiy = fmaf(n1xy.w, idy, -oody);
                const float rayT = hit.dist;
                const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), box_lo));
                const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rayT));
                const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), box_lo));
                const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rayT));
                const bool swp = (c1min < c0min), t0 = (c0max >= c0min), t1 = (c1max >= c1min);
                if (!t0 && !t1) nodeAddr = pop();
                else {
                    nodeAddr = t0 ? c0 : c1;
                    if (t0 && t1) { if (swp) { const int tmp = nodeAddr; nodeAddr = c1; c1 = tmp; } push(c1); }
                }
                if (nodeAddr < 0) triAddr = (int)tri_slot_base + ~nodeAddr;
            } else {
                const float4* T = tri64 + (size_t)triAddr * 4;
                const uint32_t index = __float_as_uint(T[3].x);
                tris++;
                float t, u, w; bool done = false;
                if (woop_test(T[0], T[1], T[2], mk(ox, oy, oz), mk(dx, dy, dz), tri_lo, hit.dist, t, u, w)) {
                    hit.node = (uint32_t)inst; hit.tri = (index >> 1) + tri_base; hit.u = u; hit.v = w; hit.dist = t;
                    if (any_hit) { done = true; nodeAddr = SENT; inst = -1; }
                }
                if (!done) {
                    if (index & 1) { nodeAddr = pop(); if (nodeAddr < 0) triAddr = (int)tri_slot_base + ~nodeAddr; }
                    else triAddr++;
                }
            }
        }
    }
};
} // namespace

// info_out: [0] usable [1] treelet nodes [2] scene root [3] deepest stack entry seen
extern "C" int emu_staged_trace_rays(const ctl_scene_view* v, int treelet_budget, int stack_rows, int n, const ctl_traversal_ray* rays, ctl_trace_result* out, unsigned long long counts[3], int info_out[4]) {
    ctlb::StagedHost H;
    ctlb::build_staging_tris(*v, H);
    info_out[0] = H.usable; if (!H.usable) return 1;
    ctlb::build_staging_nodes(*v, treelet_budget, H);
    info_out[1] = H.tl_nodes; info_out[2] = H.scene_root;
    Walker W{v, &H, stack_rows};
    for (int i = 0; i < n; i++) {
        Hit hit; W.trace(rays[i], false, false, hit);
        float* o = (float*)out + (size_t)i * 5;
        o[0] = hit.dist; o[1] = hit.u; o[2] = hit.v; memcpy(o + 3, &hit.tri, 4); memcpy(o + 4, &hit.node, 4);
    }
    counts[0] = W.inner; counts[1] = W.tris; counts[2] = W.insts; info_out[3] = W.max_sp;
    return 0;
}

extern "C" int emu_staged_intersect(const ctl_scene_view* v, int treelet_budget, int stack_rows, int n, const ctl_traversal_ray* rays, ctl_traversal_result* out, int any_hit) {
    ctlb::StagedHost H;
    ctlb::build_staging_tris(*v, H);
    if (!H.usable) return 1;
    ctlb::build_staging_nodes(*v, treelet_budget, H);
    Walker W{v, &H, stack_rows};
    for (int i = 0; i < n; i++) {
        Hit hit; W.trace(rays[i], true, any_hit != 0, hit);
        uint4 res = make_uint4(__float_as_uint(hit.dist), 0xffffffffu, 0xffffffffu, 0u);
        if (hit.tri != 0xffffffffu) {
            res.y = hit.node; res.z = hit.tri;
            const unsigned short xd = (unsigned short)(hit.u * 65535), yd = (unsigned short)(hit.v * 65535);
            res.w = ((uint32_t)yd << 16) | (uint32_t)xd;
        }
        memcpy(&out[i], &res, 16);
    }
    return 0;
}
