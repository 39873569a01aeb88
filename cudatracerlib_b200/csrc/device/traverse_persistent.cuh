// traverse_persistent.cuh -- persistent-warp, block-scheduled two-level BVH traversal (the production kernel).
//
// Replaces intersectKernel<ANY_HIT> (Kernel/TraceHelper.cu:326-734): the same per-ray algorithm as traverse.cuh
// (identical node/triangle visit ORDER per ray, identical arithmetic => identical results and visit counts), but the
// warp is scheduled B200-style instead of ray-batch style:
//
//   * every lane owns one ray; a finished lane is refilled from a warp-private chunk of the ray queue
//     (one global atomic per TP_CHUNK rays instead of one per 32) -- no lane idles while the queue has work;
//   * every loop iteration ALL lanes that want an inner-node step (N) take one, then all lanes that want a triangle
//     test (T) take one (gated by a small lane threshold so the T block is not run for 1-2 lanes); the rare, expensive
//     blocks -- instance enter/exit (L) and finish+fetch (F) -- run only when enough lanes (th_l, th_f) ask for
//     them or nothing else is runnable.  Lane states are voted with two __ballot_sync per iteration;
//   * one traversal stack per lane for both BVH levels (instance entry pushes a sentinel marker);
//   * 64-byte nodes are fetched with two 256-bit loads (LDG.E.256, sm_100a).
#pragma once
#include "traverse.cuh"
#include <cfloat>

namespace ctld {

constexpr int TP_CHUNK = 128;    // rays a warp claims per global atomic (large queues; small queues use smaller chunks, see chunk below)
constexpr int TP_STACK = 64;     // BVHTraversal.h: int traversalStack[64]

struct TravOut { // where results go (MODE-dependent, see k_intersect)
    float4* hit_a; uint32_t* hit_node; const float4* sh_payload; float4* cl; void* api_out;
    // MODE 4 (fused launch): work items [0, n_ext) are extension rays (closest hit, results as MODE 0) taken from `rays`,
    // items [n_ext, n) are shadow rays (any hit, results as MODE 1) taken from `sh_rays`
    const float4* sh_rays; int n_ext;
    // MODE 5 (fused API launch, WavefrontPathTracer): items [0, n_ext) are closest-hit queries from `rays` -> api_out, items [n_ext, n) are
    // any-hit queries from `sh_rays` -> api_out2; both with intersectKernel semantics (MODE 2: tmin / tmax from the ray, 16-byte results)
    void* api_out2;
    unsigned* cls_hist;   // staged kernel, MODE 0 / 4: 8 counters of this bounce's hit records by material class (null = not wanted)
};

struct TravTune { int th_t, th_l, th_f, th_n_exit, t_steps, chunk, drain_prefetch; }; // lane thresholds of the T / L / F blocks; th_n_exit = node steps per iteration; t_steps = triangle tests per iteration (staged kernel); chunk = rays per claim (0 = by queue size); drain_prefetch = 1: once the queue is exhausted, every node step prefetches both children (staged kernel)

template <int MODE, bool ANY_HIT, bool COUNT>
__device__ __forceinline__ void trace_persistent(const DScene& S, const float4* __restrict__ rays, int n, unsigned* work_ctr, const TravOut& out,
                                                 const TravTune& tune, VisitCounters<COUNT>& cnt) {
    const int TH_T = tune.th_t, TH_L = tune.th_l, TH_F = tune.th_f, N_STEPS = tune.th_n_exit > 0 ? tune.th_n_exit : 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int stack[TP_STACK];

    // per-lane ray state
    int ray_i = -1;                 // queue index, -1 = idle
    int nodeAddr = SENT;            // >= 0 inner node (float4 units, level-relative), < 0 leaf, SENT = level exhausted
    int sp = 0;
    int inst = -1;                  // instance (Node) index while inside a mesh BVH, -1 at scene level
    int triAddr = 0;                // current slot inside a mesh leaf
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, idx = 0, idy = 0, idz = 0, oodx = 0, oody = 0, oodz = 0;
    float tri_lo = 0, box_lo = 0;
    Hit hit; hit.dist = 0; hit.u = hit.v = 0; hit.tri = hit.node = 0xffffffffu;
    bool lane_any = ANY_HIT;        // MODE 4: per-lane any-hit flag (shadow rays)
    const float4* my_rays = rays;   // MODE 4: queue this lane's ray came from (for the world-space reload at instance exit)
    const float4* nbase = S.scene_nodes;
    const float4* wbase = S.woop; const uint32_t* ibase = S.tri_index; uint32_t tri_base = 0;

    // warp-uniform pool of claimed rays.  Chunk size: ~1/8 of a warp's fair share, 32..TP_CHUNK, so that small queues
    // (late bounces, image split over many GPUs) still balance across the grid's warps.
    const int n_warps = (int)(gridDim.x * (blockDim.x >> 5));
    const int chunk = max(32, min(TP_CHUNK, (n / (n_warps * 8)) & ~31));
    int pool_next = 0, pool_end = 0;
    bool exhausted = (n <= 0);

    // lane state: 0 = N (inner node step), 1 = T (triangle test), 2 = L (instance enter / exit), 3 = F (finished or idle)
    int state = 3;
    auto classify = [&]() { // after nodeAddr / inst changed
        state = ((unsigned)nodeAddr < (unsigned)SENT) ? 0 : (nodeAddr < 0 ? (inst >= 0 ? 1 : 2) : (inst >= 0 ? 2 : 3));
    };

    for (;;) {
        // two votes give every lane's state; everything below is warp-uniform arithmetic on the masks
        const unsigned b0 = __ballot_sync(0xffffffffu, state & 1), b1 = __ballot_sync(0xffffffffu, state & 2);
        const unsigned mNT = ~b1;              // lanes in N or T
        const unsigned mF = b0 & b1, mL = b1 & ~b0;
        bool runF = false, runL = false;
        if (b1) { // some lane is finished / idle / at a level switch: decide whether the rare blocks are worth running now
            const int nF = exhausted ? __popc(mF & __ballot_sync(0xffffffffu, ray_i >= 0)) : __popc(mF);
            if (mNT == 0u && mL == 0u && nF == 0) break; // queue exhausted and every lane idle
            runF = nF >= TH_F || (mNT == 0u && mL == 0u);
            runL = mL != 0u && (__popc(mL) >= TH_L || mNT == 0u);
        }

        // ---- F: write finished results, fetch new rays (rare, expensive: gated)
        if (runF) {
            if (state == 3 && ray_i >= 0) {
                const int i = ray_i;
                if (MODE == 0 || (MODE == 4 && !lane_any)) {
                    out.hit_a[i] = make_float4(hit.dist, hit.u, hit.v, __uint_as_float(hit.tri));
                    out.hit_node[i] = hit.node;
                } else if (MODE == 1 || MODE == 4) {
                    if (hit.tri == 0xffffffffu) { // unoccluded: add the pending NEE term (each path has <= 1 shadow ray per bounce)
                        const float4 pl = ldg_stream(out.sh_payload + i);
                        const uint32_t p = __float_as_uint(pl.w);
                        float4 c = out.cl[p];
                        c.x = c.x + pl.x; c.y = c.y + pl.y; c.z = c.z + pl.z;
                        out.cl[p] = c;
                    }
                } else if (MODE == 2 || MODE == 5) {
                    uint4 res = make_uint4(__float_as_uint(hit.dist), 0xffffffffu, 0xffffffffu, 0u);
                    if (hit.tri != 0xffffffffu) {
                        res.y = hit.node; res.z = hit.tri;
                        const unsigned short xd = (unsigned short)(hit.u * 65535), yd = (unsigned short)(hit.v * 65535); // TraceHelper.cu:726-727
                        res.w = ((uint32_t)yd << 16) | (uint32_t)xd;
                    }
                    ((uint4*)((MODE == 5 && lane_any) ? out.api_out2 : out.api_out))[i] = res;
                } else {
                    float* o5 = (float*)out.api_out + (size_t)i * 5;
                    o5[0] = hit.dist; o5[1] = hit.u; o5[2] = hit.v; o5[3] = __uint_as_float(hit.tri); o5[4] = __uint_as_float(hit.node);
                }
                ray_i = -1;
            }
            if (!exhausted) {
                const unsigned mFree = mF;
                int need = __popc(mFree);
                const int my_rank = __popc(mFree & lt_mask);
                const bool is_free = (mFree >> lane) & 1u;
                int got_before = 0; // rays handed out in earlier rounds of this block
                while (need > 0) {
                    if (pool_next >= pool_end) {
                        unsigned base = 0;
                        if (lane == 0) base = atomicAdd(work_ctr, (unsigned)chunk);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if ((int)base >= n) { exhausted = true; break; }
                        pool_next = (int)base; pool_end = min((int)base + chunk, n);
                    }
                    const int take = min(need, pool_end - pool_next);
                    const int r = my_rank - got_before;
                    if (is_free && r >= 0 && r < take) {
                        int i = pool_next + r;
                        if (MODE == 4 || MODE == 5) { lane_any = i >= out.n_ext; my_rays = lane_any ? out.sh_rays : rays; if (lane_any) i -= out.n_ext; }
                        const float4 ro = ldg_stream(my_rays + 2 * i), rd = ldg_stream(my_rays + 2 * i + 1);
                        ray_i = i;
                        ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                        hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu;
                        if (MODE == 3) { tri_lo = S.ray_eps; box_lo = 0.0f; hit.dist = FLT_MAX; }
                        else if (MODE == 2 || MODE == 5) { tri_lo = ro.w; box_lo = ro.w; hit.dist = rd.w; }
                        else { tri_lo = ro.w; box_lo = 0.0f; hit.dist = rd.w; }
                        sp = 0; stack[0] = SENT;
                        inst = -1; nbase = S.scene_nodes;
                        nodeAddr = S.n_nodes ? S.scene_start : SENT; // start < 0: the single instance leaf (BVHTraversal.h:11-12)
                        if (nodeAddr >= 0) { // scene-level traversal needs the world-space slab constants; a leaf start goes straight to L
                            idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz);
                            oodx = ox * idx; oody = oy * idy; oodz = oz * idz;
                        }
                        classify();
                    }
                    pool_next += take; need -= take; got_before += take;
                }
            }
        }

        // ---- L: instance enter / exit (gated; always runs right after a fetch, whose new rays usually start at a leaf)
        if (runL || runF) {
            if (state == 2) {
                if (nodeAddr < 0) { // enter instance ~nodeAddr (TraceHelper.cu:91-99)
                    const int nodeIdx = ~nodeAddr;
                    if (COUNT) ((VisitCounters<true>&)cnt).inst++;
                    const ctl_node* N = S.nodes + nodeIdx;
                    const ctl_mesh* M = S.meshes + __ldg(&N->mesh_index);
                    const uint32_t node_off = __ldg(&M->bvh_node_offset), tri_off4 = __ldg(&M->bvh_tri_offset), idx_off = __ldg(&M->bvh_idx_offset);
                    tri_base = __ldg(&M->tri_offset);
                    const float4* inv = S.node_inv_xf + (size_t)nodeIdx * 4;
                    const V3 d = xf_dir(inv, mk(dx, dy, dz)), o = xf_point(inv, mk(ox, oy, oz));
                    ox = o.x; oy = o.y; oz = o.z; dx = d.x; dy = d.y; dz = d.z;
                    nbase = S.bvh_nodes + node_off; wbase = S.woop + tri_off4; ibase = S.tri_index + idx_off;
                    inst = nodeIdx;
                    sp++; stack[sp] = SENT; // marker: popping it ends the mesh level
                    nodeAddr = 0;
                } else { // mesh level exhausted: back to the scene level with the world-space ray
                    const float4 ro = __ldg(my_rays + 2 * ray_i), rd = __ldg(my_rays + 2 * ray_i + 1);
                    ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                    nbase = S.scene_nodes;
                    inst = -1;
                    nodeAddr = stack[sp]; sp--;
                }
                idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz);
                oodx = ox * idx; oody = oy * idy; oodz = oz * idz;
                classify();
            }
        }

        // ---- N: inner-node step(s) for every lane that has one
        for (int ns = 0; ns < N_STEPS; ns++)
        if (state == 0) {
            const F8 nA = ldg256(nbase + nodeAddr), nB = ldg256(nbase + nodeAddr + 2);
            const float4 n0xy = nA.lo, n1xy = nA.hi, nz = nB.lo, cn = nB.hi;
            if (COUNT) ((VisitCounters<true>&)cnt).inner++;
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
            if (!t0 && !t1) { nodeAddr = stack[sp]; sp--; }
            else {
                nodeAddr = t0 ? c0 : c1;
                if (t0 && t1) {
                    if (swp) { const int tmp = nodeAddr; nodeAddr = c1; c1 = tmp; }
                    sp++; stack[sp] = c1;
                }
            }
            if (nodeAddr < 0) triAddr = ~nodeAddr;
            classify();
        }

        // ---- T: one triangle test for every lane inside a mesh leaf (skipped while only a few lanes want it)
        const unsigned mT2 = __ballot_sync(0xffffffffu, state == 1);
        if (mT2 && (__popc(mT2) >= TH_T || (mT2 | b1) == 0xffffffffu)) { // enough T lanes, or no lane can take a node step
            if (state == 1) {
                const float4 v00 = ldg_stream(wbase + triAddr * 3 + 0);
                const float4 v11 = ldg_stream(wbase + triAddr * 3 + 1);
                const float4 v22 = ldg_stream(wbase + triAddr * 3 + 2);
                const uint32_t index = ldg_stream(ibase + triAddr);
                if (COUNT) ((VisitCounters<true>&)cnt).tris++;
                float t, u, v;
                bool done = false;
                if (woop_test(v00, v11, v22, mk(ox, oy, oz), mk(dx, dy, dz), tri_lo, hit.dist, t, u, v)) {
                    hit.node = (uint32_t)inst; hit.tri = (index >> 1) + tri_base; hit.u = u; hit.v = v; hit.dist = t;
                    if ((MODE == 4 || MODE == 5) ? lane_any : ANY_HIT) { done = true; nodeAddr = SENT; inst = -1; } // first hit terminates the ray (TraceHelper.cu:675-679)
                }
                if (!done) {
                    if (index & 1) { nodeAddr = stack[sp]; sp--; if (nodeAddr < 0) triAddr = ~nodeAddr; }
                    else triAddr++;
                }
                classify();
            }
        }
    }
}

} // namespace ctld
