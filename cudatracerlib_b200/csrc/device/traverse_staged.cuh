// traverse_staged.cuh -- the persistent-warp traversal with its hot state staged in shared memory (sm_100a).
//
// Same per-ray algorithm, visit order and arithmetic as traverse_persistent.cuh / traverse.cuh (== TracerayTemplate,
// Engine/SpatialStructures/BVH/BVHTraversal.h:8-120, and the leaf callbacks of Kernel/TraceHelper.cu:88-180, :326-734), so hits and
// visit counts stay bit-identical to the oracle.  What changes is where the bytes come from.  ncu of the persistent kernel
// (profiles/r01m_ncu_traversal_final.md, r01v_ncu_c4_trav.md) shows DRAM at 1-3 % and L1/TEX at 73-78 % of peak: the kernel is bound by
// L1 tag look-ups of divergent 32-byte sectors, of which a path ray of config 2 issues about 100 (50 for its 25 nodes, ~20 for stack
// pushes / pops in local memory, ~23 for its 6 triangle tests).  This kernel removes the look-ups that are not node fetches:
//
//   * per-lane traversal stack in SHARED memory (column layout [row][thread]: bank = lane, conflict-free for any mix of depths), the
//     top of the stack in a register so that a pop never waits for memory; rows beyond `stack_rows` spill to local memory;
//   * leaf triangles from a derived 64-byte record (Woop rows + leaf word): 2 x LDG.256 instead of 3 x LDG.128 + LDG.32 over two arrays;
//   * instance entry from a derived 64-byte record (inverse-transform rows + mesh offsets): 2 x LDG.256 instead of nine scattered loads;
//   * the top of the scene-level / mesh trees (the nodes with the largest world-space boxes) as a TREELET in shared memory, filled once
//     per CTA by a TMA bulk copy (cp.async.bulk, mbarrier complete_tx) -- node addresses inside the treelet carry bit 0, children that
//     leave it hold their ordinary global address, so the walk itself is unchanged.
//
// The derived arrays are built by ctl_upload_scene (csrc/staging.cpp) from the reference-layout view; the view itself is untouched.
#pragma once
#include "traverse_persistent.cuh"

namespace ctld {

struct StagedScene {
    const float4* tri64;     // 4 x float4 per leaf slot: Woop rows a, b, c, (leaf word bits, material class, 0, 0)
    const float4* inst;      // 4 x float4 per (pseudo-)node: inverse transform rows 0..2, (bvh node base [float4 units], tri64 slot base, TriangleData base, root word)
    const float4* treelet;   // shared-memory image of the treelet: tl_nodes x 64 bytes, 16-byte chunks swizzled (see tl_chunk)
    int tl_nodes;            // 0 = no treelet
    int scene_root;          // node address the scene-level walk starts at: the view's start node, or a treelet address
    int stack_rows;          // stack entries per lane kept in shared memory
    int ray_tma;             // 1: ray-queue chunks reach the warp through TMA bulk copies into a per-warp shared-memory buffer (k_intersect_staged<.., RT = true>)
};
constexpr int ST_TL = 1;       // node address bit 0: treelet node, slot = address >> 2 (ordinary addresses are float4 units, multiples of 4)
constexpr int ST_NEEDS_W = 2;  // root word bit 1: the instance's inverse transform has a projective last row -> divide by w like xf_point

// 16-byte chunk j (0..3) of treelet node n lives at float4 index n * 4 + (j ^ ((n >> 1) & 3)): the eight lanes of an LDS.128 phase that read
// chunk j of eight different nodes then spread over all eight 16-byte bank groups instead of two.
CTL_DEV int tl_chunk(int n, int j) { return n * 4 + (j ^ ((n >> 1) & 3)); }

CTL_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
CTL_DEV void prefetch_l1(const void* p) {
#ifdef __CUDACC__
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#endif
}

// One CTA-wide TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS).  Called by every thread of the CTA.
CTL_DEV void tma_fill(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar_smem) {
#ifdef __CUDACC__
    const uint32_t bar = smem_u32(bar_smem), dst = smem_u32(dst_smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        for (uint32_t off = 0; off < bytes; off += 16384u) {
            const uint32_t len = bytes - off < 16384u ? bytes - off : 16384u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst + off), "l"((const char*)src_gmem + off), "r"(len), "r"(bar) : "memory");
        }
    }
    __syncthreads(); // the initialised barrier is visible to every waiter
    uint32_t done = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    } while (!done);
#endif
}

// Per-warp mbarrier helpers of the ray-queue staging (RT): the barrier is initialised with one arrival; every refill adds its bytes with expect_tx,
// issues the bulk copies and arrives once, so a phase completes exactly when all bytes of that refill have landed.
CTL_DEV void mbar_init1(uint32_t bar) {
#ifdef __CUDACC__
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
CTL_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
#ifdef __CUDACC__
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
#endif
}
CTL_DEV void mbar_arrive(uint32_t bar) {
#ifdef __CUDACC__
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
#endif
}
CTL_DEV bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done = 1;
#ifdef __CUDACC__
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
    return done != 0;
}
CTL_DEV void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
#ifdef __CUDACC__
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
#endif
}

// RT: ray-queue staging.  A refill claims consecutive queue entries, so the rays of the lanes it feeds are one contiguous block (two when the claim
// crosses the extension / shadow queue boundary of a fused launch): lane 0 sends it with cp.async.bulk into the warp's 1 KB buffer and the lanes pick
// their ray up on the first iteration after the bytes have landed -- the DRAM latency of the queue read is hidden behind the node steps of the
// warp's other lanes instead of stalling all 32 at the first use.
template <int MODE, bool ANY_HIT, bool COUNT, bool RT>
__device__ __forceinline__ void trace_staged(const DScene& S, const StagedScene& SS, const float4* __restrict__ rays, int n, unsigned* work_ctr, const TravOut& out,
                                             const TravTune& tune, VisitCounters<COUNT>& cnt, const float4* __restrict__ tl, int* __restrict__ ss, const int NT,
                                             const float4* wbuf /* RT: this warp's 32-ray buffer */, const uint32_t wbar /* RT: this warp's mbarrier (shared address) */) {
    const int TH_T = tune.th_t, TH_L = tune.th_l, TH_F = tune.th_f, N_STEPS = tune.th_n_exit > 0 ? tune.th_n_exit : 1, T_STEPS = tune.t_steps > 0 ? tune.t_steps : 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int SD = SS.stack_rows;
    int ovf[TP_STACK]; // rows beyond SD (rare: deep trees)

    // per-lane ray state
    int ray_i = -1;
    int nodeAddr = SENT;
    int sp = 0, tos = SENT;         // entries 0..sp of the reference's stack; entry sp lives in `tos`, entry k < sp in row k + 1
    int inst = -1;
    int triAddr = 0;                // current slot of the tri64 array (absolute)
    uint32_t tri_slot_base = 0, tri_base = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, idx = 0, idy = 0, idz = 0, oodx = 0, oody = 0, oodz = 0;
    float tri_lo = 0, box_lo = 0;
    Hit hit; hit.dist = 0; hit.u = hit.v = 0; hit.tri = hit.node = 0xffffffffu;
    bool lane_any = ANY_HIT;
    const float4* nbase = S.scene_nodes;

    auto push = [&](int v) { sp++; if (sp <= SD) ss[sp * NT] = tos; else ovf[sp - SD - 1] = tos; tos = v; };
    auto pop = [&]() { const int r = tos; tos = sp <= SD ? ss[sp * NT] : ovf[sp - SD - 1]; sp--; return r; }; // row 0 is never written: read (and ignored) by the last pop of a ray

    const int n_warps = (int)(gridDim.x * (blockDim.x >> 5));
    const int chunk = tune.chunk > 0 ? tune.chunk : max(32, min(TP_CHUNK, (n / (n_warps * 8)) & ~31));
    int pool_next = 0, pool_end = 0;
    bool exhausted = (n <= 0);

    int pend_slot = -1;             // RT: >= 0 while this lane's ray is on its way into the warp buffer (the lane idles in state 3)
    uint32_t wphase = 0;            // RT: parity of the warp barrier's current phase (warp-uniform)

#if defined(CTL_EXP_OLD_RAY_BOOST) && defined(__CUDACC__)
    int x_age = 0;
#endif
    auto start_ray = [&](const float4 ro, const float4 rd) { // a fetched ray enters the scene level
#if defined(CTL_EXP_OLD_RAY_BOOST) && defined(__CUDACC__)
        x_age = 0;
#endif
        ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
        hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu;
        if (MODE == 3) { tri_lo = S.ray_eps; box_lo = 0.0f; hit.dist = FLT_MAX; }
        else if (MODE == 2 || MODE == 5) { tri_lo = ro.w; box_lo = ro.w; hit.dist = rd.w; }
        else { tri_lo = ro.w; box_lo = 0.0f; hit.dist = rd.w; }
        sp = 0; tos = SENT;
        inst = -1; nbase = S.scene_nodes;
        nodeAddr = S.n_nodes ? SS.scene_root : SENT;
        if (nodeAddr >= 0) {
            idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz);
            oodx = ox * idx; oody = oy * idy; oodz = oz * idz;
        }
    };

    int state = 3;
    auto classify = [&]() { state = ((unsigned)nodeAddr < (unsigned)SENT) ? 0 : (nodeAddr < 0 ? (inst >= 0 ? 1 : 2) : (inst >= 0 ? 2 : 3)); };

#if defined(CTL_EXP_DROP_STRAGGLERS) && defined(__CUDACC__)
    int x_drain_iters = 0;   // experiment (build variant only, WRONG results): upper bound of what deferring a draining launch's last rays could gain
#endif
    for (;;) {
#if defined(CTL_EXP_DROP_STRAGGLERS) && defined(__CUDACC__)
        if (exhausted && (MODE == 0 || MODE == 4)) {
            const unsigned x_live = __ballot_sync(0xffffffffu, state != 3);
            if (x_live && __popc(x_live) <= CTL_EXP_DROP_LANES && ++x_drain_iters > CTL_EXP_DROP_STRAGGLERS) {
                if (state != 3) { atomicAdd(work_ctr + 1, 1u); hit.tri = 0xffffffffu; nodeAddr = SENT; inst = -1; sp = 0; tos = SENT; state = 3; }   // abandoned: written out as a miss
            }
        }
#endif
        unsigned mPend = 0;
        if (RT) { // rays on their way: have the bytes landed?  (one test per iteration, warp-uniform)
            mPend = __ballot_sync(0xffffffffu, pend_slot >= 0);
            if (mPend && mbar_test(wbar, wphase)) {
                wphase ^= 1u; mPend = 0u;
                if (pend_slot >= 0) { start_ray(wbuf[2 * pend_slot], wbuf[2 * pend_slot + 1]); classify(); pend_slot = -1; }
            }
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, state & 1), b1 = __ballot_sync(0xffffffffu, state & 2);
        const unsigned mNT = ~b1;
        const unsigned mF = b0 & b1 & ~mPend, mL = b1 & ~b0;
        bool runF = false, runL = false;
        if (b1) {
            const int nF = exhausted ? __popc(mF & __ballot_sync(0xffffffffu, ray_i >= 0)) : __popc(mF);
            if (mNT == 0u && mL == 0u && nF == 0 && mPend == 0u) break;
            runF = nF >= TH_F || (mNT == 0u && mL == 0u);
            runL = mL != 0u && (__popc(mL) >= TH_L || mNT == 0u);
        }

        // ---- F: write finished results, fetch new rays
        if (runF) {
            if ((MODE == 0 || MODE == 4) && out.cls_hist) { // class histogram of this bounce's hit records (the shade stage runs one launch per material class)
                const bool wr = state == 3 && ray_i >= 0 && !(MODE == 4 && lane_any) && (!RT || pend_slot < 0);
                const unsigned mw = __ballot_sync(0xffffffffu, wr);
                if (wr) {
                    const unsigned cls = hit.tri >> TRI_CLS_SHIFT; // 7 = miss
                    const unsigned peers = __match_any_sync(mw, cls);
                    if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(out.cls_hist + cls, (unsigned)__popc(peers));
                }
            }
            if (state == 3 && ray_i >= 0 && (!RT || pend_slot < 0)) {
                const int i = ray_i;
                if (MODE == 0 || (MODE == 4 && !lane_any)) {
                    out.hit_a[i] = make_float4(hit.dist, hit.u, hit.v, __uint_as_float(hit.tri)); // tri word keeps the class bits: k_shade strips them
                    out.hit_node[i] = hit.node;
                } else if (MODE == 1 || MODE == 4) {
                    if (hit.tri == 0xffffffffu) {
                        const float4 pl = ldg_stream(out.sh_payload + i);
                        const uint32_t p = __float_as_uint(pl.w);
                        float4 c = out.cl[p];
                        c.x = c.x + pl.x; c.y = c.y + pl.y; c.z = c.z + pl.z;
                        out.cl[p] = c;
                    }
                } else if (MODE == 2 || MODE == 5) {
                    uint4 res = make_uint4(__float_as_uint(hit.dist), 0xffffffffu, 0xffffffffu, 0u);
                    if (hit.tri != 0xffffffffu) {
                        res.y = hit.node; res.z = hit.tri & TRI_IDX_MASK;
                        const unsigned short xd = (unsigned short)(hit.u * 65535), yd = (unsigned short)(hit.v * 65535); // TraceHelper.cu:726-727
                        res.w = ((uint32_t)yd << 16) | (uint32_t)xd;
                    }
                    ((uint4*)((MODE == 5 && lane_any) ? out.api_out2 : out.api_out))[i] = res;
                } else {
                    float* o5 = (float*)out.api_out + (size_t)i * 5;
                    o5[0] = hit.dist; o5[1] = hit.u; o5[2] = hit.v; o5[3] = __uint_as_float(hit.tri == 0xffffffffu ? hit.tri : (hit.tri & TRI_IDX_MASK)); o5[4] = __uint_as_float(hit.node);
                }
                ray_i = -1;
            }
            if (!exhausted && (!RT || mPend == 0u)) { // RT: one refill in flight per warp (single buffer)
                const unsigned mFree = mF;
                int need = __popc(mFree);
                const int my_rank = __popc(mFree & lt_mask);
                const bool is_free = (mFree >> lane) & 1u;
                int got_before = 0;
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
                    if (RT && lane == 0) { // the block [pool_next, pool_next + take) of the queue(s) -> warp buffer slots [got_before, got_before + take)
                        int a = pool_next, cnt_a = take, cnt_b = 0;
                        if (MODE == 4 || MODE == 5) { if (a >= out.n_ext) { cnt_b = take; cnt_a = 0; } else if (a + take > out.n_ext) { cnt_a = out.n_ext - a; cnt_b = take - cnt_a; } }
                        const uint32_t dst = smem_u32(wbuf) + (uint32_t)got_before * 32u;
                        mbar_expect_tx(wbar, (uint32_t)take * 32u);
                        if (cnt_a) bulk_g2s(dst, rays + 2 * (size_t)a, (uint32_t)cnt_a * 32u, wbar);
                        if (cnt_b) bulk_g2s(dst + (uint32_t)cnt_a * 32u, out.sh_rays + 2 * (size_t)(a + cnt_a - out.n_ext), (uint32_t)cnt_b * 32u, wbar);
                    }
                    if (is_free && r >= 0 && r < take) {
                        int i = pool_next + r;
                        const float4* q = rays;
                        if (MODE == 4 || MODE == 5) { lane_any = i >= out.n_ext; if (lane_any) { i -= out.n_ext; q = out.sh_rays; } }
                        ray_i = i;
                        if (RT) pend_slot = my_rank; // == got_before + r: picked up when the refill's bytes have landed; the lane idles in state 3 until then
                        else { start_ray(ldg_stream(q + 2 * i), ldg_stream(q + 2 * i + 1)); classify(); }
                    }
                    pool_next += take; need -= take; got_before += take;
                }
                if (RT && got_before > 0 && lane == 0) mbar_arrive(wbar); // closes the refill: the phase completes when its bytes are in
            }
        }

        // ---- L: instance enter / exit
        if (runL || runF) {
            if (state == 2) {
                if (nodeAddr < 0) { // enter instance ~nodeAddr (TraceHelper.cu:91-99): one 64-byte record
                    const int nodeIdx = ~nodeAddr;
                    if (COUNT) ((VisitCounters<true>&)cnt).inst++;
                    const float4* I = SS.inst + (size_t)nodeIdx * 4;
                    const F8 iA = ldg256(I), iB = ldg256(I + 2);
                    const uint32_t root = __float_as_uint(iB.hi.w);
                    // xf_dir / xf_point (dmath.cuh) on rows 0..2; w = 1 exactly for an affine inverse, so the division is skipped unless flagged
                    const float ddx = dot4(iA.lo, dx, dy, dz, 0.0f), ddy = dot4(iA.hi, dx, dy, dz, 0.0f), ddz = dot4(iB.lo, dx, dy, dz, 0.0f);
                    float px = dot4(iA.lo, ox, oy, oz, 1.0f), py = dot4(iA.hi, ox, oy, oz, 1.0f), pz = dot4(iB.lo, ox, oy, oz, 1.0f);
                    if (root & ST_NEEDS_W) { const float w = dot4(__ldg(S.node_inv_xf + (size_t)nodeIdx * 4 + 3), ox, oy, oz, 1.0f); px = px / w; py = py / w; pz = pz / w; }
                    ox = px; oy = py; oz = pz; dx = ddx; dy = ddy; dz = ddz;
                    nbase = S.bvh_nodes + __float_as_uint(iB.hi.x); tri_slot_base = __float_as_uint(iB.hi.y); tri_base = __float_as_uint(iB.hi.z);
                    inst = nodeIdx;
                    push(SENT); // marker: popping it ends the mesh level
                    nodeAddr = (int)(root & ~(uint32_t)ST_NEEDS_W);
                } else { // mesh level exhausted: back to the scene level with the world-space ray
                    const float4* q = ((MODE == 4 || MODE == 5) && lane_any) ? out.sh_rays : rays;
                    const float4 ro = __ldg(q + 2 * ray_i), rd = __ldg(q + 2 * ray_i + 1);
                    ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                    nbase = S.scene_nodes;
                    inst = -1;
                    nodeAddr = pop();
                }
                idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz);
                oodx = ox * idx; oody = oy * idy; oodz = oz * idz;
                classify();
            }
        }

        const bool drain = exhausted && tune.drain_prefetch;
        // ---- N: up to N_STEPS inner-node steps for every lane that has one.  Inside the block a lane only asks "still an inner node?"; its
        // state is classified once, on leaving
        if (state == 0) {
            int ns = N_STEPS;
#if defined(CTL_EXP_OLD_RAY_BOOST) && defined(__CUDACC__)   // experiment (build variant): a ray that has been in its lane for many iterations takes more node steps per iteration
            if (++x_age > CTL_EXP_OLD_RAY_AGE) ns *= CTL_EXP_OLD_RAY_BOOST;
#endif
            do {
                F8 nA, nB;
                if (nodeAddr & ST_TL) { // treelet node: four conflict-spread LDS.128
                    const int t = nodeAddr >> 2;
                    nA.lo = tl[tl_chunk(t, 0)]; nA.hi = tl[tl_chunk(t, 1)]; nB.lo = tl[tl_chunk(t, 2)]; nB.hi = tl[tl_chunk(t, 3)];
                } else { nA = ldg256(nbase + nodeAddr); nB = ldg256(nbase + nodeAddr + 2); }
#if defined(CTL_EXP_EXTRA_NODE_LOADS) && defined(__CUDACC__)   // experiment (build variant only): what one more L1 wavefront per lane and node step costs -- the result is not used
                bool x_never = false;   // (a box coordinate never has this NaN payload: the load cannot be eliminated, the walk is unchanged)
                for (int xk = 0; xk < CTL_EXP_EXTRA_NODE_LOADS; xk++) { float xd; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(xd) : "l"(nbase + nodeAddr + (xk & 3))); x_never |= __float_as_uint(xd) == 0x7fc12345u; }
#endif
                const float4 n0xy = nA.lo, n1xy = nA.hi, nz = nB.lo, cn = nB.hi;
                if (COUNT) ((VisitCounters<true>&)cnt).inner++;
                int c0 = __float_as_int(cn.x), c1 = __float_as_int(cn.y);
#if defined(CTL_EXP_EXTRA_NODE_LOADS) && defined(__CUDACC__)
                if (x_never) c0 = c1;
#endif
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
                if (drain) { // the launch is draining: what is left are its longest rays, each alone in its warp and paced by the node-fetch latency -- bring the
                             // children that will be visited into L1 while the slab tests run (a loss in the throughput regime, profiles/r01k_*: hence only here)
                    if (t0 && c0 >= 0 && !(c0 & ST_TL)) { prefetch_l1(nbase + c0); prefetch_l1(nbase + c0 + 2); }
                    if (t1 && c1 >= 0 && !(c1 & ST_TL)) { prefetch_l1(nbase + c1); prefetch_l1(nbase + c1 + 2); }
                }
                if (!t0 && !t1) nodeAddr = pop();
                else {
                    nodeAddr = t0 ? c0 : c1;
                    if (t0 && t1) {
                        if (swp) { const int tmp = nodeAddr; nodeAddr = c1; c1 = tmp; }
                        push(c1);
                    }
                }
            } while (--ns > 0 && (unsigned)nodeAddr < (unsigned)SENT);
            if (nodeAddr < 0) triAddr = (int)tri_slot_base + ~nodeAddr;
            classify();
        }

        // ---- T: up to T_STEPS triangle tests for every lane inside a mesh leaf
        const unsigned mT2 = __ballot_sync(0xffffffffu, state == 1);
        if (mT2 && (__popc(mT2) >= TH_T || (mT2 | b1) == 0xffffffffu)) {
            if (state == 1) {
                int ts = T_STEPS;
                bool more;
                do {
                    const float4* T = SS.tri64 + (size_t)triAddr * 4;
                    const F8 tA = ldg256(T), tB = ldg256(T + 2);
                    const uint32_t index = __float_as_uint(tB.hi.x);
                    if (COUNT) ((VisitCounters<true>&)cnt).tris++;
                    float t, u, v;
                    more = true;
                    if (woop_test(tA.lo, tA.hi, tB.lo, mk(ox, oy, oz), mk(dx, dy, dz), tri_lo, hit.dist, t, u, v)) {
                        hit.node = (uint32_t)inst; hit.tri = ((index >> 1) + tri_base) | (__float_as_uint(tB.hi.y) << TRI_CLS_SHIFT); hit.u = u; hit.v = v; hit.dist = t; // + the material class of the slot (staging.cpp)
                        if ((MODE == 4 || MODE == 5) ? lane_any : ANY_HIT) { more = false; nodeAddr = SENT; inst = -1; } // first hit terminates the ray (TraceHelper.cu:675-679)
                    }
                    if (more) {
                        if (index & 1) { nodeAddr = pop(); more = nodeAddr < 0; if (more) triAddr = (int)tri_slot_base + ~nodeAddr; } // end of the leaf: the next entry may be another leaf
                        else triAddr++;
                    }
                } while (--ts > 0 && more);
                classify();
            }
        }
    }
}

// Shared-memory layout of one CTA: [treelet image: tl_nodes * 64 B][mbarrier: 16 B][stack: (stack_rows + 1) rows x blockDim ints]
//                                  [RT: ray buffers, 1 KB per warp][RT: mbarriers, 8 B per warp]

#ifndef CTL_STAGED_MAX_THREADS
#define CTL_STAGED_MAX_THREADS 1024 // <= 64 registers per thread, so that any block size up to 1024 keeps 1024 threads per SM resident
#endif
#ifndef CTL_STAGED_MIN_BLOCKS
#define CTL_STAGED_MIN_BLOCKS 1
#endif

// MODEs as k_intersect (wavefront.cuh).  MODE 4 / 5 (fused launches) take their second queue through `out`.
template <int MODE, bool ANY_HIT, bool COUNT, bool RT>
__global__ void __launch_bounds__(CTL_STAGED_MAX_THREADS, CTL_STAGED_MIN_BLOCKS) k_intersect_staged(const __grid_constant__ DScene S, const __grid_constant__ StagedScene SS, const __grid_constant__ TravTune tune,
        const float4* __restrict__ rays, const unsigned* __restrict__ n_ptr, const unsigned* __restrict__ n2_ptr, int n_fixed, unsigned* work_ctr,
        const __grid_constant__ TravOut out_in, unsigned long long* visit_out) {
    extern __shared__ __align__(128) unsigned char staged_smem[];
    const int NT = (int)blockDim.x;
    const float4* tl = (const float4*)staged_smem;
    unsigned char* bar = staged_smem + (size_t)SS.tl_nodes * 64;
    int* ss = (int*)(bar + 16) + threadIdx.x;
    // RT: after the stack rows, one 1 KB ray buffer per warp, then one 8-byte mbarrier per warp
    const float4* wbuf = nullptr; uint32_t wbar = 0;
    if (RT) {
        unsigned char* rb = bar + 16 + (size_t)(SS.stack_rows + 1) * NT * 4;
        const int warp = (int)(threadIdx.x >> 5);
        wbuf = (const float4*)(rb + (size_t)warp * 1024);
        wbar = smem_u32(rb + (size_t)(NT >> 5) * 1024 + (size_t)warp * 8);
        if ((threadIdx.x & 31) == 0) mbar_init1(wbar);
        __syncwarp();
    }
    if (SS.tl_nodes) tma_fill(staged_smem, SS.treelet, (uint32_t)SS.tl_nodes * 64u, bar);
    TravOut out = out_in;
    int n = n_ptr ? (int)*n_ptr : n_fixed;
    if (MODE == 4 || MODE == 5) { out.n_ext = n; n += (int)*n2_ptr; }
    VisitCounters<COUNT> cnt;
    trace_staged<MODE, ANY_HIT, COUNT, RT>(S, SS, rays, n, work_ctr, out, tune, cnt, tl, ss, NT, wbuf, wbar);
    if (COUNT) {
        VisitCounters<true>& c = (VisitCounters<true>&)cnt;
        unsigned a = c.inner, b = c.tris, e = c.inst;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(visit_out, (unsigned long long)a); atomicAdd(visit_out + 1, (unsigned long long)b); atomicAdd(visit_out + 2, (unsigned long long)e); }
    }
}

} // namespace ctld
