// bvh_presplit.cuh -- triangle pre-splitting for the GPU builders (the quality bar is SplitBVHBuilder.cpp:163-203, whose advantage on
// long thin triangles is the spatial split, SplitBVHBuilder.cpp:205-330).
//
// An object-partitioning builder can only put a triangle's whole box into one leaf; a sliver that runs diagonally through the scene drags a box hundreds
// of times its own size through every level above it.  The CPU split BVH cuts such references while it builds.  On the device the cut happens BEFORE the
// build: a triangle whose box is much larger than a flat triangle of its orientation needs is replaced by k references, each with the tight box of one
// piece of the triangle (the piece is found by clipping the triangle, recursively, against the mid-plane of the longest box axis), and the builders run on
// references.  Leaves then hold (triangle << 1 | last) words as always -- a triangle may simply appear in several leaves, exactly as in the trees the
// reference's own builder writes.  k = round(scale / 2 * sqrt(box area / ideal area)), ideal area = 2 * sum of the triangle's axis-projected areas: an
// axis-aligned right triangle has ratio 2 (k = 1, untouched), a 50:1 diagonal sliver ~67 (k = 4).  `scale` is lowered by the host when the reference
// budget would be exceeded.  Two passes (count, scan, emit): positions come from the scan, so the output is deterministic.
#pragma once
#include "bvh_build.cuh"

namespace ctlbvh {

constexpr int SPLIT_MAX_PIECES = 16;
constexpr int SPLIT_MAX_VERTS = 8;    // 3 + one per clip plane on the way down (depth <= 4) + 1

__device__ __forceinline__ int split_pieces(const float* __restrict__ v, float scale, int max_pieces) {
    const float e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2], e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
    const float ideal = fabsf(e1y * e2z - e1z * e2y) + fabsf(e1z * e2x - e1x * e2z) + fabsf(e1x * e2y - e1y * e2x);   // 2 * (projected areas)
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = fminf(v[a], fminf(v[3 + a], v[6 + a])); hi[a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a])); }
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    const float sa = 2.0f * (dx * dy + dy * dz + dz * dx);
    if (!(ideal > 0.0f) || !(sa > 0.0f) || !(sa < 3.0e38f)) return 1;   // degenerate or non-finite: left alone
    const float k = floorf(0.5f * scale * sqrtf(sa / ideal) + 0.5f);
    return k >= (float)max_pieces ? max_pieces : (k >= 1.0f ? (int)k : 1);
}

__global__ void k_split_count(const float* __restrict__ verts9, uint32_t n, float scale, int max_pieces, unsigned* __restrict__ counts /* n + 1 */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = (unsigned)split_pieces(verts9 + (size_t)i * 9, scale, max_pieces);
    else if (i == n) counts[i] = 0u;
}

struct SplitPoly { float p[SPLIT_MAX_VERTS][3]; int n; };

// both sides of the plane x[axis] = pos (Sutherland-Hodgman; vertices on the plane go to both sides, cut points are shared bit for bit)
__device__ void split_clip(const SplitPoly& in, int axis, float pos, SplitPoly& lo, SplitPoly& hi) {
    lo.n = hi.n = 0;
    for (int i = 0; i < in.n; i++) {
        const float* a = in.p[i]; const float* b = in.p[i + 1 == in.n ? 0 : i + 1];
        const float da = a[axis] - pos, db = b[axis] - pos;
        if (da <= 0.0f && lo.n < SPLIT_MAX_VERTS) { lo.p[lo.n][0] = a[0]; lo.p[lo.n][1] = a[1]; lo.p[lo.n][2] = a[2]; lo.n++; }
        if (da >= 0.0f && hi.n < SPLIT_MAX_VERTS) { hi.p[hi.n][0] = a[0]; hi.p[hi.n][1] = a[1]; hi.p[hi.n][2] = a[2]; hi.n++; }
        if ((da < 0.0f && db > 0.0f) || (da > 0.0f && db < 0.0f)) {
            const float t = da / (da - db);
            float q[3];
            for (int c = 0; c < 3; c++) { const float x = a[c] + t * (b[c] - a[c]); q[c] = fminf(fmaxf(x, fminf(a[c], b[c])), fmaxf(a[c], b[c])); }
            q[axis] = pos;
            if (lo.n < SPLIT_MAX_VERTS) { lo.p[lo.n][0] = q[0]; lo.p[lo.n][1] = q[1]; lo.p[lo.n][2] = q[2]; lo.n++; }
            if (hi.n < SPLIT_MAX_VERTS) { hi.p[hi.n][0] = q[0]; hi.p[hi.n][1] = q[1]; hi.p[hi.n][2] = q[2]; hi.n++; }
        }
    }
}
// two representable steps outward: the cut points are rounded, the piece's true outline may pass an ulp outside their hull
__device__ __forceinline__ float step_down(float x) { return x == 0.0f ? -1.0e-37f : (x > 0.0f ? __uint_as_float(__float_as_uint(x) - 2u) : __uint_as_float(__float_as_uint(x) + 2u)); }
__device__ __forceinline__ float step_up(float x) { return x == 0.0f ? 1.0e-37f : (x > 0.0f ? __uint_as_float(__float_as_uint(x) + 2u) : __uint_as_float(__float_as_uint(x) - 2u)); }

__global__ void __launch_bounds__(128) k_split_emit(const float* __restrict__ verts9, uint32_t n, float scale, int max_pieces, const unsigned* __restrict__ offsets /* exclusive scan of the counts */,
                                                    const float4* __restrict__ tri_boxes, float4* __restrict__ ref_boxes, uint32_t* __restrict__ ref_tri) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* v = verts9 + (size_t)t * 9;
    const int k = split_pieces(v, scale, max_pieces);
    unsigned out = offsets[t];
    const float4 tlo = tri_boxes[2 * t], thi = tri_boxes[2 * t + 1];
    if (k == 1) { ref_boxes[2 * out] = tlo; ref_boxes[2 * out + 1] = thi; ref_tri[out] = t; return; }
    SplitPoly stack[5]; int cnt[5]; int sp = 0;
    for (int i = 0; i < 3; i++) for (int c = 0; c < 3; c++) stack[0].p[i][c] = v[i * 3 + c];
    stack[0].n = 3; cnt[0] = k;
    while (sp >= 0) {
        const SplitPoly P = stack[sp]; const int c = cnt[sp]; sp--;
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (int i = 0; i < P.n; i++) for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], P.p[i][a]); hi[a] = fmaxf(hi[a], P.p[i][a]); }
        bool emit = c == 1 || sp + 2 > 4;
        SplitPoly A, B; int ca = 0;
        if (!emit) {
            const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
            const int axis = ex >= ey && ex >= ez ? 0 : (ey >= ez ? 1 : 2);
            ca = c / 2;
            const float pos = lo[axis] + (hi[axis] - lo[axis]) * ((float)ca / (float)c);
            split_clip(P, axis, pos, A, B);
            if (A.n < 3 || B.n < 3 || !(pos > lo[axis]) || !(pos < hi[axis])) emit = true;   // nothing to cut along this axis: the piece stays whole
        }
        if (emit) {   // c references of this piece (c > 1 only for pieces that could not be cut: equal boxes end up in one leaf)
            const float4 blo = make_float4(fmaxf(step_down(lo[0]), tlo.x), fmaxf(step_down(lo[1]), tlo.y), fmaxf(step_down(lo[2]), tlo.z), 0.0f);
            const float4 bhi = make_float4(fminf(step_up(hi[0]), thi.x), fminf(step_up(hi[1]), thi.y), fminf(step_up(hi[2]), thi.z), 0.0f);
            for (int r = 0; r < c; r++) { ref_boxes[2 * out] = blo; ref_boxes[2 * out + 1] = bhi; ref_tri[out] = t; out++; }
        } else {
            sp++; stack[sp] = B; cnt[sp] = c - ca;
            sp++; stack[sp] = A; cnt[sp] = ca;
        }
    }
}

} // namespace ctlbvh
