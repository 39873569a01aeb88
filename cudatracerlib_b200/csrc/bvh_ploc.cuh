// bvh_ploc.cuh -- agglomerative GPU build of a mesh BVH in the reference's layout ("-> SAH-quality").
//
// The yard-stick is the quality of Engine/SpatialStructures/BVH/SplitBVHBuilder.cpp:163-203 (SAH partition per node).  The LBVH of bvh_build.cuh
// splits where a Morton bit flips, which costs 1.31x / 1.13x traversal time on configs 2 / 4.  This builder keeps the Morton order only as a
// search structure and decides every merge by surface area -- parallel locally-ordered clustering (Meister & Bittner 2018):
//   clusters = the triangles in Morton order; each round every cluster finds, among the `radius` clusters on either side, the partner with the
//   smallest union area; mutual choices merge into a new inner node; the survivors are compacted (order kept) and the round repeats.
// Pairs are compared under a total order (area, min index ^ 1, max index): the globally best pair is always mutual (progress), and among equal
// areas -- duplicated or degenerate geometry -- neighbours pair up (2k, 2k+1) so that a run of identical boxes still halves per round.
// Node ids are handed out from n-2 downwards by rank (scan, no atomics): the last merge is the root = id 0 and the build is deterministic.
// The generic back end then fits boxes bottom-up, takes the SAH leaf / split decision for subtrees of <= 8 triangles, lays the leaves out in
// tree order (every subtree owns a contiguous slot range) and the emitted nodes in pre-order (as the CPU builder's re-layout), and hands over to
// the emitters of bvh_build.cuh.
#pragma once
#include "bvh_build.cuh"

namespace ctlbvh {

constexpr int PLOC_MAX_RADIUS = 32;

__global__ void k_ploc_init(int n, const uint32_t* __restrict__ vals, const float4* __restrict__ tri_boxes, int* __restrict__ cid, float4* __restrict__ cbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = vals[i];
    cid[i] = ~i; cbox[2 * i] = tri_boxes[2 * p]; cbox[2 * i + 1] = tri_boxes[2 * p + 1];
}

__device__ __forceinline__ float union_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}
// total order on unordered pairs {i, j}: (area, min ^ 1, max)
__device__ __forceinline__ bool pair_less(float a0, int i0, int j0, float a1, int i1, int j1) {
    if (a0 != a1) return a0 < a1;
    const int m0 = min(i0, j0) ^ 1, m1 = min(i1, j1) ^ 1;
    if (m0 != m1) return m0 < m1;
    return max(i0, j0) < max(i1, j1);
}
// Round state on the device: st[0] = clusters left, st[1] = next node id, st[2] = rounds that merged something.  The host launches rounds in groups
// over grids sized for the last count it has seen (an upper bound) and looks at the state once per group.
__global__ void __launch_bounds__(256) k_ploc_nn(const int* __restrict__ st, int radius, const float4* __restrict__ cbox, int* __restrict__ nn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = st[0];
    if (i >= nc) return;
    const float4 lo = cbox[2 * i], hi = cbox[2 * i + 1];
    const int j0 = max(0, i - radius), j1 = min(nc - 1, i + radius);
    int best = -1; float best_a = 3.0e38f;
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        float a = union_area(lo, hi, cbox[2 * j], cbox[2 * j + 1]);
        if (!(a == a)) a = 3.0e38f;   // NaN boxes (non-finite vertices) still pair up
        if (best < 0 || pair_less(a, i, j, best_a, i, best)) { best = j; best_a = a; }
    }
    nn[i] = best;
}
// flags[i] = survives | merges << 32; slot nc = 0 (receives the totals after the exclusive scan)
__global__ void k_ploc_flags(const int* __restrict__ st, const int* __restrict__ nn, unsigned long long* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = st[0];
    if (i > nc) return;
    if (i == nc) { flags[i] = 0ull; return; }
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    flags[i] = (mutual && i > j) ? 0ull : (1ull | ((mutual && i < j) ? (1ull << 32) : 0ull));
}
__global__ void __launch_bounds__(1024) k_scan_exclusive64(unsigned long long* __restrict__ data, const int* __restrict__ st) {
    __shared__ unsigned long long warp_sums[32];
    const int t = threadIdx.x;
    const uint32_t m = (uint32_t)st[0] + 1u;
    const uint32_t per = (m + 1023) / 1024, lo = (uint32_t)t * per, hi = min(lo + per, m);
    unsigned long long local = 0;
    for (uint32_t i = lo; i < hi; i++) local += data[i];
    unsigned long long incl = local;
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
    if ((t & 31) == 31) warp_sums[t >> 5] = incl;
    __syncthreads();
    if (t < 32) { unsigned long long w = warp_sums[t], wi = w; for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += v; } warp_sums[t] = wi - w; }
    __syncthreads();
    unsigned long long run = warp_sums[t >> 5] + incl - local;
    for (uint32_t i = lo; i < hi; i++) { const unsigned long long v = data[i]; data[i] = run; run += v; }
}
// survivors keep their order; a merge at position i (partner j > i) becomes inner node next_id - rank
__global__ void k_ploc_merge(const int* __restrict__ st, const int* __restrict__ nn, const unsigned long long* __restrict__ scan, const int* __restrict__ cid_in, const float4* __restrict__ cbox_in,
                             int* __restrict__ cid_out, float4* __restrict__ cbox_out, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent_int, int* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = st[0], next_id = st[1];
    if (i >= nc) return;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    if (mutual && i > j) return;
    const unsigned long long s = scan[i];
    const int p = (int)(uint32_t)s;
    float4 lo = cbox_in[2 * i], hi = cbox_in[2 * i + 1];
    int id = cid_in[i];
    if (mutual) {
        const int a = id, b = cid_in[j];
        id = next_id - (int)(uint32_t)(s >> 32);
        left[id] = a; right[id] = b;
        if (a >= 0) parent_int[a] = id; else parent_leaf[~a] = id;
        if (b >= 0) parent_int[b] = id; else parent_leaf[~b] = id;
        const float4 blo = cbox_in[2 * j], bhi = cbox_in[2 * j + 1];
        lo.x = fminf(lo.x, blo.x); lo.y = fminf(lo.y, blo.y); lo.z = fminf(lo.z, blo.z);
        hi.x = fmaxf(hi.x, bhi.x); hi.y = fmaxf(hi.y, bhi.y); hi.z = fmaxf(hi.z, bhi.z);
        if (id == 0) parent_int[0] = -1;
    }
    cid_out[p] = id; cbox_out[2 * p] = lo; cbox_out[2 * p + 1] = hi;
}

__global__ void k_ploc_advance(int* __restrict__ st, const unsigned long long* __restrict__ scan) {   // <<<1, 1>>> after the merge
    const unsigned long long tot = scan[st[0]];
    const int merges = (int)(uint32_t)(tot >> 32);
    st[0] = (int)(uint32_t)tot; st[1] -= merges; st[2] += merges > 0 ? 1 : 0;
}

// ---- generic back end: any binary tree over the sorted triangles (children: >= 0 inner id, < 0 ~sorted position; root = inner 0) ------------------
// bottom-up: boxes, triangle counts, SAH cost, leaf decision (as k_fit_boxes), number of emitted inner nodes per subtree
__global__ void k_fit_counts(const float4* __restrict__ tri_boxes, const uint32_t* __restrict__ vals, int n, const int* __restrict__ left, const int* __restrict__ right,
                             const int* __restrict__ parent_int, const int* __restrict__ parent_leaf, unsigned* __restrict__ flags, float4* __restrict__ node_box, float* __restrict__ cost,
                             unsigned char* __restrict__ collapse, int* __restrict__ count, int* __restrict__ ecount) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int node = parent_leaf[s];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(flags + node, 1u) == 0u) return;
        float4 lo = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0), hi = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0);
        const int ch[2] = {left[node], right[node]};
        float child_cost = 0.0f; int cnt = 0, ecnt = 0;
        for (int k = 0; k < 2; k++) {
            float4 a, b;
            if (ch[k] < 0) { const uint32_t p = vals[~ch[k]]; a = tri_boxes[2 * p]; b = tri_boxes[2 * p + 1]; child_cost += box_area(a, b); cnt += 1; }
            else { a = __ldcg(node_box + 2 * ch[k]); b = __ldcg(node_box + 2 * ch[k] + 1); child_cost += __ldcg(cost + ch[k]); cnt += __ldcg(count + ch[k]); ecnt += __ldcg(ecount + ch[k]); }
            lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
            hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
        }
        const float area = box_area(lo, hi);
        const float split_cost = 1.2f * area + child_cost, leaf_cost = cnt <= MAX_LEAF ? (float)cnt * area : 3.0e38f;
        const bool leaf = leaf_cost <= split_cost;
        collapse[node] = leaf ? 1 : 0;
        __stcg(cost + node, fminf(leaf_cost, split_cost));
        __stcg(count + node, cnt); __stcg(ecount + node, leaf ? 0 : ecnt + 1);
        __stcg(node_box + 2 * node, lo); __stcg(node_box + 2 * node + 1, hi);
        node = parent_int[node];
    }
}
// tree order: thread t < n-1 = inner node t (first / last slot of its subtree, pre-order index among the emitted nodes); thread n-1+s = leaf s (its slot).
// Walks to the root: everything to the left of the path lies before.
__global__ void k_tree_order(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent_int, const int* __restrict__ parent_leaf,
                             const int* __restrict__ count, const int* __restrict__ ecount, int* __restrict__ first, int* __restrict__ last, unsigned* __restrict__ pre, int* __restrict__ slot) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n - 1) return;
    const bool is_leaf = t >= n - 1;
    int cur = is_leaf ? ~(t - (n - 1)) : t;
    int p = is_leaf ? parent_leaf[t - (n - 1)] : parent_int[t];
    int before = 0; unsigned pre_idx = 0;
    while (p >= 0) {
        pre_idx += 1u;
        if (right[p] == cur) { const int l = left[p]; before += l < 0 ? 1 : count[l]; pre_idx += l < 0 ? 0u : (unsigned)ecount[l]; }
        cur = p; p = parent_int[p];
    }
    if (is_leaf) slot[t - (n - 1)] = before;
    else { first[t] = before; last[t] = before + count[t] - 1; pre[t] = pre_idx; }
}
// leaves move to their tree-order slots
__global__ void k_leaf_remap(int n, const int* __restrict__ slot, const uint32_t* __restrict__ vals, const int* __restrict__ parent_leaf, uint32_t* __restrict__ vals_out, int* __restrict__ parent_leaf_out,
                             int* __restrict__ left, int* __restrict__ right) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) { const int s = slot[t]; vals_out[s] = vals[t]; parent_leaf_out[s] = parent_leaf[t]; }
    if (t < n - 1) { const int l = left[t], r = right[t]; if (l < 0) left[t] = ~slot[~l]; if (r < 0) right[t] = ~slot[~r]; }
}

} // namespace ctlbvh
