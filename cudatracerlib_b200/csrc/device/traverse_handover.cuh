// traverse_handover.cuh -- the staged traversal kernel for frames rendered as two interleaved half-wavefronts: a launch that has run out of queue
// HANDS ITS UNFINISHED RAYS OVER to the next launch instead of draining.
//
// Why: a persistent launch ends when its last ray does.  151 K rays are in flight per GPU; once the queue is empty they finish over ~0.45 ms during
// which the SMs run emptier and emptier -- nine times per wavefront, 12 % of the frame at 1/8 of the image per GPU (DESIGN.md section 5;
// profiles/r03b_drop_probe.log: abandoning what is still in flight 16 iterations after the queue ran dry takes the frame from 37.9 to 33.1 ms = the ideal
// 1/8, and touches 4 % of a launch's rays).  The rays cannot be abandoned, and a kernel boundary is the only place the shade stage can start -- but the NEXT
// traversal launch can finish them if it belongs to another wavefront: the frame is cut into two half-wavefronts A and B (half the passes each) whose
// launches alternate on one stream,
//     T(A,0)  T(B,0)  S(A,0)  T(A,1)  S(B,0)  T(B,1)  S(A,1)  ...          T = traversal launch, S = shade launches of a bounce
// every T first resumes the rays its predecessor (the other half) suspended, writing their results into the other half's records, and S(A,b) is simply
// enqueued after T(B,b).  No path ever lags a bounce, nothing is traced twice, results are the same rays' same hits.
//
// Same per-ray algorithm, visit order and arithmetic as traverse_staged.cuh (fused MODE 4 semantics: extension rays closest hit, shadow rays any hit).
// A suspended ray is 18 words of lane state + its stack entries (HO_WORDS per record, one record per lane at most).
#pragma once
#include "traverse_staged.cuh"

namespace ctld {

constexpr int HO_WORDS = 96;   // 18 + TP_STACK (64) entries, padded

struct HandOver {
    const uint32_t* resume; const unsigned* n_resume;   // records suspended by the previous launch (the other half's): resumed first; n_resume may be null (none)
    TravOut alt; const float4* alt_rays;                 // the other half's result / queue arrays of that launch: where a resumed ray's answer goes, where its ray lives
    uint32_t* suspend; unsigned* n_suspend;              // records this launch suspends (capacity = its resident lanes); null = this launch must finish everything
    int drain_iters;                                     // loop iterations a warp keeps going after it found the queue empty, before it suspends its rays
};

template <bool DUMMY = true>
__global__ void __launch_bounds__(CTL_STAGED_MAX_THREADS, CTL_STAGED_MIN_BLOCKS) k_intersect_handover(const __grid_constant__ DScene S, const __grid_constant__ StagedScene SS, const __grid_constant__ TravTune tune,
        const float4* __restrict__ rays, const unsigned* __restrict__ n_ext_ptr, const unsigned* __restrict__ n_sh_ptr, unsigned* work_ctr, const __grid_constant__ TravOut out, const __grid_constant__ HandOver H) {
    extern __shared__ __align__(128) unsigned char ho_smem[];
    const int NT = (int)blockDim.x;
    int* ss = (int*)(ho_smem + 16) + threadIdx.x;
    const int TH_T = tune.th_t, TH_L = tune.th_l, TH_F = tune.th_f, N_STEPS = tune.th_n_exit > 0 ? tune.th_n_exit : 1, T_STEPS = tune.t_steps > 0 ? tune.t_steps : 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int SD = SS.stack_rows;
    int ovf[TP_STACK];

    const int n_res = H.n_resume ? (int)min(*H.n_resume, 0x3fffffffu) : 0;
    const int n_ext = n_ext_ptr ? (int)*n_ext_ptr : 0;
    const int n = n_res + n_ext + (n_sh_ptr ? (int)*n_sh_ptr : 0);

    int ray_i = -1;
    int nodeAddr = SENT;
    int sp = 0, tos = SENT;
    int inst = -1;
    int triAddr = 0;
    uint32_t tri_slot_base = 0, tri_base = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, idx = 0, idy = 0, idz = 0, oodx = 0, oody = 0, oodz = 0;
    float tri_lo = 0;
    Hit hit; hit.dist = 0; hit.u = hit.v = 0; hit.tri = hit.node = 0xffffffffu;
    bool lane_any = false, lane_alt = false;
    const float4* nbase = S.scene_nodes;

    auto push = [&](int v) { sp++; if (sp <= SD) ss[sp * NT] = tos; else ovf[sp - SD - 1] = tos; tos = v; };
    auto pop = [&]() { const int r = tos; tos = sp <= SD ? ss[sp * NT] : ovf[sp - SD - 1]; sp--; return r; };

    const int chunk = tune.chunk > 0 ? tune.chunk : 32;
    int pool_next = 0, pool_end = 0;
    bool exhausted = (n <= 0);
    int drain = 0;   // iterations since this warp found the queue empty (warp-uniform)

    int state = 3;
    auto classify = [&]() { state = ((unsigned)nodeAddr < (unsigned)SENT) ? 0 : (nodeAddr < 0 ? (inst >= 0 ? 1 : 2) : (inst >= 0 ? 2 : 3)); };
    auto derive = [&]() { idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz); oodx = ox * idx; oody = oy * idy; oodz = oz * idz; };

    for (;;) {
        // ---- hand-over: the queue is empty and this warp has kept going for `drain_iters` iterations -> its own unfinished rays go to the next launch
        if (exhausted && H.suspend && drain >= 0 && ++drain > H.drain_iters) {
            drain = -1;   // once
            if (state != 3 && !lane_alt) {
                const unsigned slot = atomicAdd(H.n_suspend, 1u);
                uint32_t* R = H.suspend + (size_t)slot * HO_WORDS;
                R[0] = (uint32_t)ray_i | (lane_any ? 0x80000000u : 0u); R[1] = (uint32_t)nodeAddr; R[2] = (uint32_t)inst; R[3] = (uint32_t)triAddr; R[4] = (uint32_t)sp; R[5] = (uint32_t)tos;
                R[6] = __float_as_uint(ox); R[7] = __float_as_uint(oy); R[8] = __float_as_uint(oz); R[9] = __float_as_uint(dx); R[10] = __float_as_uint(dy); R[11] = __float_as_uint(dz);
                R[12] = __float_as_uint(hit.dist); R[13] = __float_as_uint(hit.u); R[14] = __float_as_uint(hit.v); R[15] = hit.tri; R[16] = hit.node; R[17] = __float_as_uint(tri_lo);
                for (int k = 1; k <= sp; k++) R[17 + k] = (uint32_t)(k <= SD ? ss[k * NT] : ovf[k - SD - 1]);
                ray_i = -1; nodeAddr = SENT; inst = -1; sp = 0; tos = SENT; state = 3;
            }
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, state & 1), b1 = __ballot_sync(0xffffffffu, state & 2);
        const unsigned mNT = ~b1;
        const unsigned mF = b0 & b1, mL = b1 & ~b0;
        bool runF = false, runL = false;
        if (b1) {
            const int nF = exhausted ? __popc(mF & __ballot_sync(0xffffffffu, ray_i >= 0)) : __popc(mF);
            if (mNT == 0u && mL == 0u && nF == 0) break;
            runF = nF >= TH_F || (mNT == 0u && mL == 0u);
            runL = mL != 0u && (__popc(mL) >= TH_L || mNT == 0u);
        }

        // ---- F: write finished results (to this half's records, or the other half's for a resumed ray), fetch new work
        if (runF) {
            const TravOut& O = lane_alt ? H.alt : out;
            {   // class histogram of the hit records (one shade launch per material class)
                const bool wr = state == 3 && ray_i >= 0 && !lane_any && O.cls_hist != nullptr;
                const unsigned mw = __ballot_sync(0xffffffffu, wr);
                if (wr) {
                    const unsigned cls = (hit.tri >> TRI_CLS_SHIFT) | (lane_alt ? 8u : 0u);
                    const unsigned peers = __match_any_sync(mw, cls);
                    if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(O.cls_hist + (cls & 7u), (unsigned)__popc(peers));
                }
            }
            if (state == 3 && ray_i >= 0) {
                const int i = ray_i;
                if (!lane_any) {
                    O.hit_a[i] = make_float4(hit.dist, hit.u, hit.v, __uint_as_float(hit.tri));
                    O.hit_node[i] = hit.node;
                } else if (hit.tri == 0xffffffffu) {
                    const float4 pl = ldg_stream(O.sh_payload + i);
                    const uint32_t p = __float_as_uint(pl.w);
                    float4 c = O.cl[p];
                    c.x = c.x + pl.x; c.y = c.y + pl.y; c.z = c.z + pl.z;
                    O.cl[p] = c;
                }
                ray_i = -1; lane_alt = false;
            }
            if (!exhausted) {
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
                    if (is_free && r >= 0 && r < take) {
                        int i = pool_next + r;
                        if (i < n_res) {   // a ray the previous launch suspended: restore the lane state, results go to the other half
                            const uint32_t* R = H.resume + (size_t)i * HO_WORDS;
                            const uint32_t w0 = R[0];
                            ray_i = (int)(w0 & 0x7fffffffu); lane_any = (w0 >> 31) != 0u; lane_alt = true;
                            nodeAddr = (int)R[1]; inst = (int)R[2]; triAddr = (int)R[3]; sp = (int)R[4]; tos = (int)R[5];
                            ox = __uint_as_float(R[6]); oy = __uint_as_float(R[7]); oz = __uint_as_float(R[8]); dx = __uint_as_float(R[9]); dy = __uint_as_float(R[10]); dz = __uint_as_float(R[11]);
                            hit.dist = __uint_as_float(R[12]); hit.u = __uint_as_float(R[13]); hit.v = __uint_as_float(R[14]); hit.tri = R[15]; hit.node = R[16]; tri_lo = __uint_as_float(R[17]);
                            for (int k = 1; k <= sp; k++) { const int v = (int)R[17 + k]; if (k <= SD) ss[k * NT] = v; else ovf[k - SD - 1] = v; }
                            if (inst >= 0) {
                                const F8 iB = ldg256(SS.inst + (size_t)inst * 4 + 2);
                                nbase = S.bvh_nodes + __float_as_uint(iB.hi.x); tri_slot_base = __float_as_uint(iB.hi.y); tri_base = __float_as_uint(iB.hi.z);
                            } else nbase = S.scene_nodes;
                        } else {
                            i -= n_res;
                            const float4* q = rays;
                            lane_any = i >= n_ext; lane_alt = false;
                            if (lane_any) { i -= n_ext; q = out.sh_rays; }
                            ray_i = i;
                            const float4 ro = ldg_stream(q + 2 * i), rd = ldg_stream(q + 2 * i + 1);
                            ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                            hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu;
                            tri_lo = ro.w; hit.dist = rd.w;
                            sp = 0; tos = SENT;
                            inst = -1; nbase = S.scene_nodes;
                            nodeAddr = S.n_nodes ? SS.scene_root : SENT;
                        }
                        derive(); classify();
                    }
                    pool_next += take; need -= take; got_before += take;
                }
            }
        }

        // ---- L: instance enter / exit
        if (runL || runF) {
            if (state == 2) {
                if (nodeAddr < 0) {
                    const int nodeIdx = ~nodeAddr;
                    const float4* I = SS.inst + (size_t)nodeIdx * 4;
                    const F8 iA = ldg256(I), iB = ldg256(I + 2);
                    const uint32_t root = __float_as_uint(iB.hi.w);
                    const float ddx = dot4(iA.lo, dx, dy, dz, 0.0f), ddy = dot4(iA.hi, dx, dy, dz, 0.0f), ddz = dot4(iB.lo, dx, dy, dz, 0.0f);
                    float px = dot4(iA.lo, ox, oy, oz, 1.0f), py = dot4(iA.hi, ox, oy, oz, 1.0f), pz = dot4(iB.lo, ox, oy, oz, 1.0f);
                    if (root & ST_NEEDS_W) { const float w = dot4(__ldg(S.node_inv_xf + (size_t)nodeIdx * 4 + 3), ox, oy, oz, 1.0f); px = px / w; py = py / w; pz = pz / w; }
                    ox = px; oy = py; oz = pz; dx = ddx; dy = ddy; dz = ddz;
                    nbase = S.bvh_nodes + __float_as_uint(iB.hi.x); tri_slot_base = __float_as_uint(iB.hi.y); tri_base = __float_as_uint(iB.hi.z);
                    inst = nodeIdx;
                    push(SENT);
                    nodeAddr = (int)(root & ~(uint32_t)ST_NEEDS_W);
                } else {   // back to the scene level with the world-space ray (this half's queues, or the other half's for a resumed ray)
                    const float4* q = lane_alt ? (lane_any ? H.alt.sh_rays : H.alt_rays) : (lane_any ? out.sh_rays : rays);
                    const float4 ro = __ldg(q + 2 * ray_i), rd = __ldg(q + 2 * ray_i + 1);
                    ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                    nbase = S.scene_nodes;
                    inst = -1;
                    nodeAddr = pop();
                }
                derive(); classify();
            }
        }

        // ---- N: inner-node steps
        if (state == 0) {
            int ns = N_STEPS;
            do {
                const F8 nA = ldg256(nbase + nodeAddr), nB = ldg256(nbase + nodeAddr + 2);
                const float4 n0xy = nA.lo, n1xy = nA.hi, nz = nB.lo, cn = nB.hi;
                int c0 = __float_as_int(cn.x), c1 = __float_as_int(cn.y);
                const float c0lox = fmaf(n0xy.x, idx, -oodx), c0hix = fmaf(n0xy.y, idx, -oodx);
                const float c0loy = fmaf(n0xy.z, idy, -oody), c0hiy = fmaf(n0xy.w, idy, -oody);
                const float c0loz = fmaf(nz.x, idz, -oodz), c0hiz = fmaf(nz.y, idz, -oodz);
                const float c1loz = fmaf(nz.z, idz, -oodz), c1hiz = fmaf(nz.w, idz, -oodz);
                const float c1lox = fmaf(n1xy.x, idx, -oodx), c1hix = fmaf(n1xy.y, idx, -oodx);
                const float c1loy = fmaf(n1xy.z, idy, -oody), c1hiy = fmaf(n1xy.w, idy, -oody);
                const float rayT = hit.dist;
                const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
                const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rayT));
                const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
                const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rayT));
                const bool swp = (c1min < c0min), t0 = (c0max >= c0min), t1 = (c1max >= c1min);
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

        // ---- T: triangle tests
        const unsigned mT2 = __ballot_sync(0xffffffffu, state == 1);
        if (mT2 && (__popc(mT2) >= TH_T || (mT2 | b1) == 0xffffffffu)) {
            if (state == 1) {
                int ts = T_STEPS;
                bool more;
                do {
                    const float4* T = SS.tri64 + (size_t)triAddr * 4;
                    const F8 tA = ldg256(T), tB = ldg256(T + 2);
                    const uint32_t index = __float_as_uint(tB.hi.x);
                    float t, u, v;
                    more = true;
                    if (woop_test(tA.lo, tA.hi, tB.lo, mk(ox, oy, oz), mk(dx, dy, dz), tri_lo, hit.dist, t, u, v)) {
                        hit.node = (uint32_t)inst; hit.tri = ((index >> 1) + tri_base) | (__float_as_uint(tB.hi.y) << TRI_CLS_SHIFT); hit.u = u; hit.v = v; hit.dist = t;
                        if (lane_any) { more = false; nodeAddr = SENT; inst = -1; }
                    }
                    if (more) {
                        if (index & 1) { nodeAddr = pop(); more = nodeAddr < 0; if (more) triAddr = (int)tri_slot_base + ~nodeAddr; }
                        else triAddr++;
                    }
                } while (--ts > 0 && more);
                classify();
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------------------------------------
// Deferral inside ONE wavefront ("DeferStragglers"): the rays a launch has not finished a few iterations after its queue ran dry move to the wavefront's
// NEXT traversal launch -- appended to the front of the next bounce's queues, resumed there from a record of their lane state -- and their paths run one
// bounce behind the others from then on (at most `max_lag` bounces; the path id carries the lag in its top two bits).  The shade launches of the bounce
// skip the deferred hit records (TRI_DEFERRED); the host runs max_lag extra iterations at the end of the wavefront for the paths that lag.
constexpr uint32_t TRI_DEFERRED = (6u << TRI_CLS_SHIFT) | TRI_IDX_MASK;   // class 6: no shade launch covers it
constexpr uint32_t PATH_ID_MASK = 0x3fffffffu;

struct Defer {
    const uint32_t* resume; const unsigned* n_resume;        // records of the rays the previous launch deferred (= k_ext + k_sh of them)
    const unsigned* k_ext; const unsigned* k_sh;              // how many entries at the front of this launch's extension / shadow queue they are
    uint32_t* suspend; unsigned* n_suspend;                  // records this launch writes; null = this launch finishes everything
    const uint32_t* paths_in;                                // path ids of this launch's extension queue
    float4* next_rays; uint32_t* next_paths; unsigned* next_ext_ctr;            // the next bounce's extension queue
    float4* next_sh_rays; float4* next_sh_payload; unsigned* next_sh_ctr;       // the shadow queue the next launch traces
    int drain_iters, max_lag;
};

template <bool DUMMY2 = true>
__global__ void __launch_bounds__(CTL_STAGED_MAX_THREADS, CTL_STAGED_MIN_BLOCKS) k_intersect_defer(const __grid_constant__ DScene S, const __grid_constant__ StagedScene SS, const __grid_constant__ TravTune tune,
        const float4* __restrict__ rays, const unsigned* __restrict__ n_ext_ptr, const unsigned* __restrict__ n_sh_ptr, unsigned* work_ctr, const __grid_constant__ TravOut out, const __grid_constant__ Defer H) {
    extern __shared__ __align__(128) unsigned char df_smem[];
    const int NT = (int)blockDim.x;
    int* ss = (int*)(df_smem + 16) + threadIdx.x;
    const int TH_T = tune.th_t, TH_L = tune.th_l, TH_F = tune.th_f, N_STEPS = tune.th_n_exit > 0 ? tune.th_n_exit : 1, T_STEPS = tune.t_steps > 0 ? tune.t_steps : 1;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int SD = SS.stack_rows;
    int ovf[TP_STACK];

    // the first k_ext entries of the extension queue and the first k_sh of the shadow queue are the rays the previous launch deferred: they are resumed from
    // their records (work items [0, n_res)), never started afresh
    const int n_res = H.n_resume ? (int)min(*H.n_resume, 0x3fffffffu) : 0;
    const int k_ext = H.k_ext ? (int)*H.k_ext : 0, k_sh = H.k_sh ? (int)*H.k_sh : 0;
    const int n_ext = (n_ext_ptr ? (int)*n_ext_ptr : 0) - k_ext;
    const int n = n_res + n_ext + (n_sh_ptr ? (int)*n_sh_ptr - k_sh : 0);

    int ray_i = -1;
    int nodeAddr = SENT;
    int sp = 0, tos = SENT;
    int inst = -1;
    int triAddr = 0;
    uint32_t tri_slot_base = 0, tri_base = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, idx = 0, idy = 0, idz = 0, oodx = 0, oody = 0, oodz = 0;
    float tri_lo = 0;
    Hit hit; hit.dist = 0; hit.u = hit.v = 0; hit.tri = hit.node = 0xffffffffu;
    bool lane_any = false;
    const float4* nbase = S.scene_nodes;

    auto push = [&](int v) { sp++; if (sp <= SD) ss[sp * NT] = tos; else ovf[sp - SD - 1] = tos; tos = v; };
    auto pop = [&]() { const int r = tos; tos = sp <= SD ? ss[sp * NT] : ovf[sp - SD - 1]; sp--; return r; };

    const int chunk = tune.chunk > 0 ? tune.chunk : 32;
    int pool_next = 0, pool_end = 0;
    bool exhausted = (n <= 0);
    int drain = 0;   // iterations since this warp found the queue empty (warp-uniform)

    int state = 3;
    auto classify = [&]() { state = ((unsigned)nodeAddr < (unsigned)SENT) ? 0 : (nodeAddr < 0 ? (inst >= 0 ? 1 : 2) : (inst >= 0 ? 2 : 3)); };
    auto derive = [&]() { idx = guard_inv(dx); idy = guard_inv(dy); idz = guard_inv(dz); oodx = ox * idx; oody = oy * idy; oodz = oz * idz; };

    for (;;) {
        // ---- deferral: the queue is empty and this warp has kept going for `drain_iters` iterations -> its unfinished rays move to the NEXT launch of this
        // wavefront: the ray is appended to the next bounce's queue (its path runs one bounce behind from here on: lag counter in the top bits of the path id),
        // the lane state goes into a record, the hit record of an extension ray is marked "deferred" for this bounce's shade launches
        if (exhausted && H.suspend && drain >= 0 && ++drain > H.drain_iters) {
            drain = -1;   // once
            bool go = state != 3;
            uint32_t pid = 0;
            if (go && !lane_any) { pid = H.paths_in[ray_i]; go = (pid >> 30) < (uint32_t)H.max_lag; }
            if (go) {
                const float4* q = lane_any ? out.sh_rays : rays;
                const float4 ro = __ldg(q + 2 * ray_i), rd = __ldg(q + 2 * ray_i + 1);
                unsigned j;
                if (!lane_any) {
                    j = atomicAdd(H.next_ext_ctr, 1u);
                    H.next_rays[2 * j] = ro; H.next_rays[2 * j + 1] = rd; H.next_paths[j] = pid + 0x40000000u;
                    out.hit_a[ray_i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(TRI_DEFERRED));
                    if (out.cls_hist) atomicAdd(out.cls_hist + 6, 1u);
                } else {
                    j = atomicAdd(H.next_sh_ctr, 1u);
                    H.next_sh_rays[2 * j] = ro; H.next_sh_rays[2 * j + 1] = rd; H.next_sh_payload[j] = out.sh_payload[ray_i];
                }
                const unsigned slot = atomicAdd(H.n_suspend, 1u);
                uint32_t* R = H.suspend + (size_t)slot * HO_WORDS;
                R[0] = j | (lane_any ? 0x80000000u : 0u); R[1] = (uint32_t)nodeAddr; R[2] = (uint32_t)inst; R[3] = (uint32_t)triAddr; R[4] = (uint32_t)sp; R[5] = (uint32_t)tos;
                R[6] = __float_as_uint(ox); R[7] = __float_as_uint(oy); R[8] = __float_as_uint(oz); R[9] = __float_as_uint(dx); R[10] = __float_as_uint(dy); R[11] = __float_as_uint(dz);
                R[12] = __float_as_uint(hit.dist); R[13] = __float_as_uint(hit.u); R[14] = __float_as_uint(hit.v); R[15] = hit.tri; R[16] = hit.node; R[17] = __float_as_uint(tri_lo);
                for (int k = 1; k <= sp; k++) R[17 + k] = (uint32_t)(k <= SD ? ss[k * NT] : ovf[k - SD - 1]);
                ray_i = -1; nodeAddr = SENT; inst = -1; sp = 0; tos = SENT; state = 3;
            }
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, state & 1), b1 = __ballot_sync(0xffffffffu, state & 2);
        const unsigned mNT = ~b1;
        const unsigned mF = b0 & b1, mL = b1 & ~b0;
        bool runF = false, runL = false;
        if (b1) {
            const int nF = exhausted ? __popc(mF & __ballot_sync(0xffffffffu, ray_i >= 0)) : __popc(mF);
            if (mNT == 0u && mL == 0u && nF == 0) break;
            runF = nF >= TH_F || (mNT == 0u && mL == 0u);
            runL = mL != 0u && (__popc(mL) >= TH_L || mNT == 0u);
        }

        // ---- F: write finished results (to this half's records, or the other half's for a resumed ray), fetch new work
        if (runF) {
            const TravOut& O = out;
            {   // class histogram of the hit records (one shade launch per material class)
                const bool wr = state == 3 && ray_i >= 0 && !lane_any && O.cls_hist != nullptr;
                const unsigned mw = __ballot_sync(0xffffffffu, wr);
                if (wr) {
                    const unsigned cls = hit.tri >> TRI_CLS_SHIFT;
                    const unsigned peers = __match_any_sync(mw, cls);
                    if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(O.cls_hist + (cls & 7u), (unsigned)__popc(peers));
                }
            }
            if (state == 3 && ray_i >= 0) {
                const int i = ray_i;
                if (!lane_any) {
                    O.hit_a[i] = make_float4(hit.dist, hit.u, hit.v, __uint_as_float(hit.tri));
                    O.hit_node[i] = hit.node;
                } else if (hit.tri == 0xffffffffu) {   // (atomics: a path that lags can have the shadow rays of two of its vertices in one launch)
                    const float4 pl = ldg_stream(O.sh_payload + i);
                    float* c = (float*)(O.cl + __float_as_uint(pl.w));
                    atomicAdd(c, pl.x); atomicAdd(c + 1, pl.y); atomicAdd(c + 2, pl.z);
                }
                ray_i = -1;
            }
            if (!exhausted) {
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
                    if (is_free && r >= 0 && r < take) {
                        int i = pool_next + r;
                        if (i < n_res) {   // a ray the previous launch deferred: restore the lane state; its queue entry is ray_i of this launch's queues
                            const uint32_t* R = H.resume + (size_t)i * HO_WORDS;
                            const uint32_t w0 = R[0];
                            ray_i = (int)(w0 & 0x7fffffffu); lane_any = (w0 >> 31) != 0u;
                            nodeAddr = (int)R[1]; inst = (int)R[2]; triAddr = (int)R[3]; sp = (int)R[4]; tos = (int)R[5];
                            ox = __uint_as_float(R[6]); oy = __uint_as_float(R[7]); oz = __uint_as_float(R[8]); dx = __uint_as_float(R[9]); dy = __uint_as_float(R[10]); dz = __uint_as_float(R[11]);
                            hit.dist = __uint_as_float(R[12]); hit.u = __uint_as_float(R[13]); hit.v = __uint_as_float(R[14]); hit.tri = R[15]; hit.node = R[16]; tri_lo = __uint_as_float(R[17]);
                            for (int k = 1; k <= sp; k++) { const int v = (int)R[17 + k]; if (k <= SD) ss[k * NT] = v; else ovf[k - SD - 1] = v; }
                            if (inst >= 0) {
                                const F8 iB = ldg256(SS.inst + (size_t)inst * 4 + 2);
                                nbase = S.bvh_nodes + __float_as_uint(iB.hi.x); tri_slot_base = __float_as_uint(iB.hi.y); tri_base = __float_as_uint(iB.hi.z);
                            } else nbase = S.scene_nodes;
                        } else {
                            i -= n_res;
                            const float4* q = rays;
                            lane_any = i >= n_ext;
                            if (lane_any) { i = i - n_ext + k_sh; q = out.sh_rays; } else i += k_ext;
                            ray_i = i;
                            const float4 ro = ldg_stream(q + 2 * i), rd = ldg_stream(q + 2 * i + 1);
                            ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                            hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu;
                            tri_lo = ro.w; hit.dist = rd.w;
                            sp = 0; tos = SENT;
                            inst = -1; nbase = S.scene_nodes;
                            nodeAddr = S.n_nodes ? SS.scene_root : SENT;
                        }
                        derive(); classify();
                    }
                    pool_next += take; need -= take; got_before += take;
                }
            }
        }

        // ---- L: instance enter / exit
        if (runL || runF) {
            if (state == 2) {
                if (nodeAddr < 0) {
                    const int nodeIdx = ~nodeAddr;
                    const float4* I = SS.inst + (size_t)nodeIdx * 4;
                    const F8 iA = ldg256(I), iB = ldg256(I + 2);
                    const uint32_t root = __float_as_uint(iB.hi.w);
                    const float ddx = dot4(iA.lo, dx, dy, dz, 0.0f), ddy = dot4(iA.hi, dx, dy, dz, 0.0f), ddz = dot4(iB.lo, dx, dy, dz, 0.0f);
                    float px = dot4(iA.lo, ox, oy, oz, 1.0f), py = dot4(iA.hi, ox, oy, oz, 1.0f), pz = dot4(iB.lo, ox, oy, oz, 1.0f);
                    if (root & ST_NEEDS_W) { const float w = dot4(__ldg(S.node_inv_xf + (size_t)nodeIdx * 4 + 3), ox, oy, oz, 1.0f); px = px / w; py = py / w; pz = pz / w; }
                    ox = px; oy = py; oz = pz; dx = ddx; dy = ddy; dz = ddz;
                    nbase = S.bvh_nodes + __float_as_uint(iB.hi.x); tri_slot_base = __float_as_uint(iB.hi.y); tri_base = __float_as_uint(iB.hi.z);
                    inst = nodeIdx;
                    push(SENT);
                    nodeAddr = (int)(root & ~(uint32_t)ST_NEEDS_W);
                } else {   // back to the scene level with the world-space ray (this half's queues, or the other half's for a resumed ray)
                    const float4* q = lane_any ? out.sh_rays : rays;
                    const float4 ro = __ldg(q + 2 * ray_i), rd = __ldg(q + 2 * ray_i + 1);
                    ox = ro.x; oy = ro.y; oz = ro.z; dx = rd.x; dy = rd.y; dz = rd.z;
                    nbase = S.scene_nodes;
                    inst = -1;
                    nodeAddr = pop();
                }
                derive(); classify();
            }
        }

        // ---- N: inner-node steps
        if (state == 0) {
            int ns = N_STEPS;
            do {
                const F8 nA = ldg256(nbase + nodeAddr), nB = ldg256(nbase + nodeAddr + 2);
                const float4 n0xy = nA.lo, n1xy = nA.hi, nz = nB.lo, cn = nB.hi;
                int c0 = __float_as_int(cn.x), c1 = __float_as_int(cn.y);
                const float c0lox = fmaf(n0xy.x, idx, -oodx), c0hix = fmaf(n0xy.y, idx, -oodx);
                const float c0loy = fmaf(n0xy.z, idy, -oody), c0hiy = fmaf(n0xy.w, idy, -oody);
                const float c0loz = fmaf(nz.x, idz, -oodz), c0hiz = fmaf(nz.y, idz, -oodz);
                const float c1loz = fmaf(nz.z, idz, -oodz), c1hiz = fmaf(nz.w, idz, -oodz);
                const float c1lox = fmaf(n1xy.x, idx, -oodx), c1hix = fmaf(n1xy.y, idx, -oodx);
                const float c1loy = fmaf(n1xy.z, idy, -oody), c1hiy = fmaf(n1xy.w, idy, -oody);
                const float rayT = hit.dist;
                const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
                const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rayT));
                const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
                const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rayT));
                const bool swp = (c1min < c0min), t0 = (c0max >= c0min), t1 = (c1max >= c1min);
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

        // ---- T: triangle tests
        const unsigned mT2 = __ballot_sync(0xffffffffu, state == 1);
        if (mT2 && (__popc(mT2) >= TH_T || (mT2 | b1) == 0xffffffffu)) {
            if (state == 1) {
                int ts = T_STEPS;
                bool more;
                do {
                    const float4* T = SS.tri64 + (size_t)triAddr * 4;
                    const F8 tA = ldg256(T), tB = ldg256(T + 2);
                    const uint32_t index = __float_as_uint(tB.hi.x);
                    float t, u, v;
                    more = true;
                    if (woop_test(tA.lo, tA.hi, tB.lo, mk(ox, oy, oz), mk(dx, dy, dz), tri_lo, hit.dist, t, u, v)) {
                        hit.node = (uint32_t)inst; hit.tri = ((index >> 1) + tri_base) | (__float_as_uint(tB.hi.y) << TRI_CLS_SHIFT); hit.u = u; hit.v = v; hit.dist = t;
                        if (lane_any) { more = false; nodeAddr = SENT; inst = -1; }
                    }
                    if (more) {
                        if (index & 1) { nodeAddr = pop(); more = nodeAddr < 0; if (more) triAddr = (int)tri_slot_base + ~nodeAddr; }
                        else triAddr++;
                    }
                } while (--ts > 0 && more);
                classify();
            }
        }
    }
}


} // namespace ctld
