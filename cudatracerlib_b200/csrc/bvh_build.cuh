// bvh_build.cuh -- GPU construction of a mesh BVH in the reference's layout ("next" row).
//
// Replaces, for triangle meshes, the CPU pre-process Engine/SpatialStructures/BVH/SplitBVHBuilder.cpp +
// Engine/MeshLoader/BVHBuilderHelper.cpp (which emit BVHNodeData / TriIntersectorData / TriIntersectorData2 arrays,
// handleNode SplitBVHBuilder.cpp:163-203, maxLeafSize 8 BVHBuilderHelper.cpp:119) by an LBVH:
//   triangle boxes + scene box -> 30-bit Morton codes of the centroids -> LSD radix sort (4 x 8 bit, hand-written:
//   per-block digit histograms, one scan, stable scatter with warp match/ballot ranking) -> Karras 2012 radix tree
//   (one thread per internal node) -> bottom-up box fit + SAH leaf decision (atomic arrival flags) -> subtrees of <= 8 triangles whose leaf
//   cost beats their split cost become leaves -> emit 64-byte nodes (children in float4 units, ~first-slot leaves, parent links, 0x76543210 sentinel
//   for a single-leaf mesh), Woop triangles in Morton order and (triIdx << 1 | lastInLeaf) index words.
// No spatial splits (every triangle is referenced exactly once), so the tree is not the SBVH the reference builds, but any
// tree in this layout is traversed by the same kernels with identical hit results.  All device-side, one stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "scene_builder.h"

namespace ctlbvh {

constexpr int MAX_LEAF = 8;          // BVHBuilderHelper.cpp:119
constexpr int SORT_TILE = 2048;      // keys per block and pass
constexpr int SORT_THREADS = 256;

__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// boxes[2i] = (lo, -), boxes[2i+1] = (hi, -); scene_box: 6 ordered-uint min/max
__global__ void k_tri_boxes(const float* __restrict__ verts9, uint32_t n, float4* __restrict__ boxes, unsigned* __restrict__ scene_box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (i < n) {
        const float* v = verts9 + (size_t)i * 9;
        for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) { const float x = v[k * 3 + a]; lo[a] = fminf(lo[a], x); hi[a] = fmaxf(hi[a], x); }
        boxes[2 * i] = make_float4(lo[0], lo[1], lo[2], 0.0f); boxes[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
    for (int a = 0; a < 3; a++) {
        float l = lo[a], h = hi[a];
        for (int o = 16; o > 0; o >>= 1) { l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o)); h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o)); }
        if ((threadIdx.x & 31) == 0 && l <= h) { atomicMin(scene_box + a, f2ord(l)); atomicMax(scene_box + 3 + a, f2ord(h)); }
    }
}

__device__ __forceinline__ unsigned expand10(unsigned v) { v &= 1023u; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u; return v; }

__global__ void k_morton(const float4* __restrict__ boxes, uint32_t n, const unsigned* __restrict__ scene_box, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 lo = boxes[2 * i], hi = boxes[2 * i + 1];
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    unsigned q[3];
    for (int a = 0; a < 3; a++) {
        const float l = ord2f(scene_box[a]), h = ord2f(scene_box[3 + a]);
        const float e = h - l, f = e > 0.0f ? (c[a] - l) / e : 0.0f;
        q[a] = (unsigned)fminf(fmaxf(f * 1024.0f, 0.0f), 1023.0f);
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = i;
}

// ---- LSD radix sort, one 8-bit digit per pass --------------------------------------------------------------------
// counts[digit * n_blocks + block]
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, unsigned* __restrict__ counts, int n_blocks) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE;
    for (int r = 0; r < SORT_TILE / SORT_THREADS; r++) {
        const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan of m entries in place, one block of 1024 threads
__global__ void __launch_bounds__(1024) k_scan_exclusive(unsigned* __restrict__ data, uint32_t m) {
    __shared__ unsigned warp_sums[32];
    const int t = threadIdx.x;
    const uint32_t per = (m + 1023) / 1024, lo = (uint32_t)t * per, hi = min(lo + per, m);
    unsigned local = 0;
    for (uint32_t i = lo; i < hi; i++) local += data[i];
    unsigned incl = local;
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
    if ((t & 31) == 31) warp_sums[t >> 5] = incl;
    __syncthreads();
    if (t < 32) { unsigned w = warp_sums[t], wi = w; for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += v; } warp_sums[t] = wi - w; }
    __syncthreads();
    unsigned run = warp_sums[t >> 5] + incl - local;
    for (uint32_t i = lo; i < hi; i++) { const unsigned v = data[i]; data[i] = run; run += v; }
}
// stable scatter: rank of a key among equal digits = (keys of earlier rounds) + (earlier warps of this round) + (earlier lanes)
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t n, int shift,
                                                                const unsigned* __restrict__ offsets, int n_blocks, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ unsigned running[256];                       // next output slot per digit for this block
    __shared__ unsigned warp_cnt[SORT_THREADS / 32][256];
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    running[t] = offsets[t * n_blocks + blockIdx.x];
    const uint32_t base = blockIdx.x * SORT_TILE;
    for (int r = 0; r < SORT_TILE / SORT_THREADS; r++) {
        for (int k = 0; k < SORT_THREADS / 32; k++) warp_cnt[k][t] = 0;
        __syncthreads();
        const uint32_t i = base + r * SORT_THREADS + t;
        const bool valid = i < n;
        const uint32_t key = valid ? keys_in[i] : 0xffffffffu;
        const unsigned d = valid ? (key >> shift) & 255u : 256u;   // invalid lanes form their own group
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[w][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned before = 0;
            for (int k = 0; k < w; k++) before += warp_cnt[k][d];
            const unsigned pos = running[d] + before + rank;
            keys_out[pos] = key; vals_out[pos] = vals_in[i];
        }
        __syncthreads();
        unsigned tot = 0;
        for (int k = 0; k < SORT_THREADS / 32; k++) tot += warp_cnt[k][t];
        running[t] += tot;
        __syncthreads();
    }
}

// ---- Karras 2012 radix tree over the sorted keys -----------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint32_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    return a == b ? 32 + __clz((unsigned)i ^ (unsigned)j) : __clz(a ^ b);
}
// internal node i in [0, n-2]; child encoding: >= 0 internal node, < 0 : ~leaf (sorted position)
__global__ void k_radix_tree(const uint32_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent_int, int* __restrict__ parent_leaf,
                             int* __restrict__ first, int* __restrict__ last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int L = (lo == gamma) ? ~gamma : gamma, R = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = L; right[i] = R; first[i] = lo; last[i] = hi;
    if (L >= 0) parent_int[L] = i; else parent_leaf[~L] = i;
    if (R >= 0) parent_int[R] = i; else parent_leaf[~R] = i;
    if (i == 0) parent_int[0] = -1;
}

// bottom-up boxes: node_box[2i], node_box[2i+1] for internal nodes; leaves read the sorted triangle boxes
// Also decides, bottom-up, which subtrees become leaves: SAH with C_inner = 1.2, C_tri = 1 (leaf cost = N * area, split cost =
// 1.2 * area + cost of the children); a subtree of <= MAX_LEAF triangles collapses when the leaf is not more expensive.
__device__ __forceinline__ float box_area(float4 lo, float4 hi) { const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z; return 2.0f * (dx * dy + dy * dz + dz * dx); }
__global__ void k_fit_boxes(const float4* __restrict__ tri_boxes, const uint32_t* __restrict__ vals, int n, const int* __restrict__ left, const int* __restrict__ right,
                            const int* __restrict__ parent_int, const int* __restrict__ parent_leaf, const int* __restrict__ first, const int* __restrict__ last,
                            unsigned* __restrict__ flags, float4* __restrict__ node_box, float* __restrict__ cost, unsigned char* __restrict__ collapse) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int node = parent_leaf[s];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(flags + node, 1u) == 0u) return;   // first child to arrive: the sibling will finish this node
        float4 lo = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0), hi = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0);
        const int ch[2] = {left[node], right[node]};
        float child_cost = 0.0f;
        for (int k = 0; k < 2; k++) {
            float4 a, b;
            if (ch[k] < 0) { const uint32_t p = vals[~ch[k]]; a = tri_boxes[2 * p]; b = tri_boxes[2 * p + 1]; child_cost += box_area(a, b); }
            else { a = __ldcg(node_box + 2 * ch[k]); b = __ldcg(node_box + 2 * ch[k] + 1); child_cost += __ldcg(cost + ch[k]); }
            lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
            hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
        }
        const float area = box_area(lo, hi);
        const int count = last[node] - first[node] + 1;
        const float split_cost = 1.2f * area + child_cost, leaf_cost = count <= MAX_LEAF ? (float)count * area : 3.0e38f;
        collapse[node] = leaf_cost <= split_cost ? 1 : 0;
        __stcg(cost + node, fminf(leaf_cost, split_cost));
        __stcg(node_box + 2 * node, lo); __stcg(node_box + 2 * node + 1, hi);
        node = parent_int[node];
    }
}

// An internal node becomes a BVHNodeData iff neither it nor any ancestor collapsed into a leaf (only subtrees of <= MAX_LEAF
// triangles can collapse, so the upward walk is at most a few steps).
__device__ __forceinline__ bool node_emitted(int i, const int* __restrict__ parent_int, const int* __restrict__ first, const int* __restrict__ last, const unsigned char* __restrict__ collapse) {
    for (int a = i; a >= 0 && (last[a] - first[a] + 1) <= MAX_LEAF; a = parent_int[a]) if (collapse[a]) return false;
    return true;
}
__global__ void k_mark_emitted(int n, const int* __restrict__ parent_int, const int* __restrict__ first, const int* __restrict__ last, const unsigned char* __restrict__ collapse, unsigned* __restrict__ emitted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) emitted[i] = node_emitted(i, parent_int, first, last, collapse) ? 1u : 0u;
    else if (i == n - 1) emitted[i] = 0u;   // slot n-1 receives the total after the exclusive scan
}

// one thread per internal node: write its BVHNodeData if emitted; mark the last slot of every leaf it creates
__global__ void k_emit_nodes(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent_int, const int* __restrict__ first, const int* __restrict__ last,
                             const unsigned* __restrict__ new_index /* exclusive scan of emitted */, const float4* __restrict__ tri_boxes, const uint32_t* __restrict__ vals,
                             const float4* __restrict__ node_box, const unsigned char* __restrict__ collapse, ctl_bvh_node* __restrict__ nodes, unsigned char* __restrict__ last_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (!node_emitted(i, parent_int, first, last, collapse)) return;
    ctl_bvh_node out;
    const int ch[2] = {left[i], right[i]};
    int addr[2];
    for (int k = 0; k < 2; k++) {
        float4 lo, hi; int c = ch[k];
        if (c < 0) { const uint32_t p = vals[~c]; lo = tri_boxes[2 * p]; hi = tri_boxes[2 * p + 1]; addr[k] = c; last_flag[~c] = 1; }   // single triangle: leaf ~slot
        else {
            lo = node_box[2 * c]; hi = node_box[2 * c + 1];
            if (collapse[c]) { addr[k] = ~first[c]; last_flag[last[c]] = 1; }   // collapsed subtree = leaf over slots first..last
            else addr[k] = (int)(new_index[c] * 4u);
        }
        if (k == 0) { out.a[0] = lo.x; out.a[1] = hi.x; out.a[2] = lo.y; out.a[3] = hi.y; out.c[0] = lo.z; out.c[1] = hi.z; }
        else { out.b[0] = lo.x; out.b[1] = hi.x; out.b[2] = lo.y; out.b[3] = hi.y; out.c[2] = lo.z; out.c[3] = hi.z; }
    }
    out.child0 = addr[0]; out.child1 = addr[1];
    out.parent = parent_int[i] >= 0 ? new_index[parent_int[i]] * 4u : 0xffffffffu;
    out.pad = 0;
    nodes[new_index[i]] = out;
}

// one thread per sorted slot: Woop data + index word
// (ref_tri: triangle of every reference when the triangles were pre-split, bvh_presplit.cuh; nullptr = reference i is triangle i)
__global__ void k_emit_tris(const float* __restrict__ verts9, const uint32_t* __restrict__ vals, int n, const unsigned char* __restrict__ last_flag, ctl_woop_tri* __restrict__ woop, uint32_t* __restrict__ index,
                            const uint32_t* __restrict__ ref_tri = nullptr) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t p = ref_tri ? ref_tri[vals[s]] : vals[s];
    const float* v = verts9 + (size_t)p * 9;
    ctl_woop_tri w;
    ctlb::encode_woop(ctlb::V3(v[0], v[1], v[2]), ctlb::V3(v[3], v[4], v[5]), ctlb::V3(v[6], v[7], v[8]), &w);
    woop[s] = w;
    index[s] = (p << 1) | (last_flag[s] ? 1u : 0u);
}

// whole mesh fits one leaf (n <= MAX_LEAF): root = {child0 = ~0, child1 = sentinel, left box = bounds, right box = origin}
__global__ void k_single_leaf_root(const unsigned* __restrict__ scene_box, int n, ctl_bvh_node* __restrict__ nodes, unsigned char* __restrict__ last_flag) {
    if (threadIdx.x || blockIdx.x) return;
    ctl_bvh_node out; memset(&out, 0, sizeof(out));
    out.a[0] = ord2f(scene_box[0]); out.a[1] = ord2f(scene_box[3]); out.a[2] = ord2f(scene_box[1]); out.a[3] = ord2f(scene_box[4]); out.c[0] = ord2f(scene_box[2]); out.c[1] = ord2f(scene_box[5]);
    out.child0 = ~0; out.child1 = CTL_SENTINEL; out.parent = 0xffffffffu;
    nodes[0] = out;
    last_flag[n - 1] = 1;
}

} // namespace ctlbvh
