// sbvh_builder.cpp -- split BVH (object splits by SAH + spatial splits of triangle references, Stich, Friedrich, Dietrich 2009) for the mesh level.
//
// The reference compiles every mesh with SplitBVHBuilder (Engine/SpatialStructures/BVH/SplitBVHBuilder.cpp, called from
// Engine/MeshLoader/BVHBuilderHelper.cpp:116-147 with maxLeafSize 8); this is this repo's own implementation of the same published algorithm,
// emitting the reference node layout (Engine/TriIntersectorData.h:42-117; leaf = ~first reference slot; one Woop record + one leaf word PER
// REFERENCE, so a triangle cut by spatial splits appears in several leaves).  Compared with the plain binned-SAH builder it replaces for meshes,
// long thin triangles no longer inflate the boxes of their neighbours: on the 1 M-triangle config-4 scene the oracle pops 43 % fewer inner nodes
// per ray, on par with the reference's own builder (scripts/sbvh_compare.py, profiles/r01u_sbvh_builder.log).
//
//   per node:  object split  = SAH over the reference centroids (32 bins above 4096 references, exact sweep below)
//              spatial split = tried when the object split's child boxes overlap by more than alpha * root area: 64 bins per axis, references
//                              chopped at the bin planes (triangle clipped, not just its box), entry / exit counters, SAH per plane
//              references straddling the chosen plane are split, or moved whole to one side when that is cheaper ("unsplitting")
//   cost:      C_node = 1, C_tri = 1 per reference; leaves hold <= max_leaf references; depth capped for the 64-entry traversal stack
#include "scene_builder.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <thread>
#include <exception>

namespace ctlb {
namespace {

struct Ref { uint32_t tri; Box b; };

inline Box box_intersect(const Box& a, const Box& b) { return Box(vmax(a.lo, b.lo), vmin(a.hi, b.hi)); }
inline float half_area(const Box& b) { const V3 d = b.hi - b.lo; return (d.x < 0 || d.y < 0 || d.z < 0) ? 0.0f : d.x * d.y + d.y * d.z + d.z * d.x; }

struct Sbvh {
    const float* v9;   // 9 floats per triangle
    int max_leaf;
    std::vector<ctl_bvh_node>& nodes; std::vector<uint32_t>& ordered; std::vector<uint8_t>& last;
    float root_area = 0.0f;
    float CT = 1.0f;                            // cost of a triangle test relative to a node step (CTL_SBVH_CT overrides)
    float ALPHA = 1e-5f;                        // overlap threshold for trying a spatial split, relative to the root area
    static constexpr int OBJ_BINS = 32, SPATIAL_BINS = 64, MAX_DEPTH = 48, PARALLEL_ABOVE = 8192, PARALLEL_DEPTH = 4;
    int SWEEP_BELOW = 4096;   // exact SAH sweep below this many references, binned above (CTL_SBVH_SWEEP overrides, experiments)

    V3 vert(uint32_t t, int k) const { const float* p = v9 + (size_t)t * 9 + 3 * k; return V3(p[0], p[1], p[2]); }

    // bounds of the parts of reference r left / right of the plane axis = pos (triangle clipped against the plane, then against the reference box)
    void split_ref(const Ref& r, int axis, float pos, Ref& L, Ref& R) const {
        Box l, rr;
        V3 a = vert(r.tri, 2);
        for (int k = 0; k < 3; k++) {
            const V3 b = vert(r.tri, k);
            const float a1 = a[axis], b1 = b[axis];
            if (a1 <= pos) l.grow(a);
            if (a1 >= pos) rr.grow(a);
            if ((a1 < pos && b1 > pos) || (a1 > pos && b1 < pos)) {
                float t = (pos - a1) / (b1 - a1); t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
                V3 p = a + (b - a) * t; p[axis] = pos;
                const V3 e = V3(fabsf(p.x), fabsf(p.y), fabsf(p.z)) * 4e-7f; // the interpolated point is rounded: keep the boxes conservative
                l.grow(p - e); l.grow(p + e); rr.grow(p - e); rr.grow(p + e);
            }
            a = b;
        }
        l.hi[axis] = pos; rr.lo[axis] = pos;
        L.tri = R.tri = r.tri;
        L.b = box_intersect(l, r.b); R.b = box_intersect(rr, r.b);
    }

    struct ObjSplit { float cost = 3.0e38f; int axis = -1; float pos = 0; int bin = -1; uint32_t n_left = 0; Box lb, rb; bool binned = false; float lo = 0, scale = 0; };
    struct SpatialSplit { float cost = 3.0e38f; int axis = -1; float pos = 0; };

    static float centroid2(const Ref& r, int axis) { return r.b.lo[axis] + r.b.hi[axis]; }

    ObjSplit find_object_split(std::vector<Ref>& refs) const {
        ObjSplit best;
        const uint32_t n = (uint32_t)refs.size();
        if (n > (uint32_t)SWEEP_BELOW) { // binned
            Box cb; for (const Ref& r : refs) cb.grow(V3(centroid2(r, 0), centroid2(r, 1), centroid2(r, 2)));
            for (int axis = 0; axis < 3; axis++) {
                const float lo = cb.lo[axis], ext = cb.hi[axis] - lo;
                if (!(ext > 0)) continue;
                const float scale = OBJ_BINS / ext;
                Box bb[OBJ_BINS]; uint32_t bc[OBJ_BINS] = {0};
                for (const Ref& r : refs) { int b = (int)((centroid2(r, axis) - lo) * scale); b = b < 0 ? 0 : (b >= OBJ_BINS ? OBJ_BINS - 1 : b); bb[b].grow(r.b); bc[b]++; }
                float ra[OBJ_BINS]; Box rbx[OBJ_BINS]; Box acc; uint32_t k = 0;
                for (int b = OBJ_BINS - 1; b > 0; b--) { acc.grow(bb[b]); k += bc[b]; ra[b] = half_area(acc) * k; rbx[b] = acc; }
                acc = Box(); k = 0;
                for (int b = 0; b < OBJ_BINS - 1; b++) {
                    acc.grow(bb[b]); k += bc[b];
                    if (k == 0 || k == n) continue;
                    const float cost = half_area(acc) * k + ra[b + 1];
                    if (cost < best.cost) { best.cost = cost; best.axis = axis; best.bin = b; best.n_left = k; best.lb = acc; best.rb = rbx[b + 1]; best.binned = true; best.lo = lo; best.scale = scale; }
                }
            }
            return best;
        }
        std::vector<float> right_area(n);
        std::vector<Ref> sorted = refs;
        for (int axis = 0; axis < 3; axis++) {
            std::sort(sorted.begin(), sorted.end(), [axis](const Ref& a, const Ref& b) {
                const float ca = centroid2(a, axis), cb2 = centroid2(b, axis);
                return ca < cb2 || (ca == cb2 && a.tri < b.tri);
            });
            Box acc;
            for (uint32_t i = n - 1; i > 0; i--) { acc.grow(sorted[i].b); right_area[i] = half_area(acc); }
            acc = Box();
            for (uint32_t i = 1; i < n; i++) {
                acc.grow(sorted[i - 1].b);
                const float cost = half_area(acc) * i + right_area[i] * (n - i);
                if (cost < best.cost) { best.cost = cost; best.axis = axis; best.n_left = i; best.binned = false; }
            }
        }
        if (best.axis >= 0) { // leave refs sorted along the chosen axis and compute the child boxes
            const int axis = best.axis;
            std::sort(refs.begin(), refs.end(), [axis](const Ref& a, const Ref& b) {
                const float ca = centroid2(a, axis), cb2 = centroid2(b, axis);
                return ca < cb2 || (ca == cb2 && a.tri < b.tri);
            });
            best.lb = Box(); best.rb = Box();
            for (uint32_t i = 0; i < n; i++) (i < best.n_left ? best.lb : best.rb).grow(refs[i].b);
        }
        return best;
    }

    SpatialSplit find_spatial_split(const std::vector<Ref>& refs, const Box& bounds) const {
        SpatialSplit best;
        struct Bin { Box b; uint32_t enter = 0, exit = 0; };
        // small nodes: every reference spans most of the node, so chopping at 64 planes costs 64 clips per reference and buys nothing
        const int NB = refs.size() >= 1024 ? SPATIAL_BINS : (refs.size() >= 32 ? 32 : 16);
        Bin bins[SPATIAL_BINS];
        for (int axis = 0; axis < 3; axis++) {
            const float lo = bounds.lo[axis], ext = bounds.hi[axis] - lo;
            if (!(ext > 0)) continue;
            const float bin_w = ext / NB, inv = NB / ext;
            for (int b = 0; b < NB; b++) bins[b] = Bin();
            for (const Ref& r : refs) {
                int first = (int)((r.b.lo[axis] - lo) * inv), lastb = (int)((r.b.hi[axis] - lo) * inv);
                first = first < 0 ? 0 : (first >= NB ? NB - 1 : first);
                lastb = lastb < first ? first : (lastb >= NB ? NB - 1 : lastb);
                Ref cur = r;
                for (int b = first; b < lastb; b++) { // chop the reference at every bin plane it crosses
                    Ref l, rr; split_ref(cur, axis, lo + bin_w * (float)(b + 1), l, rr);
                    bins[b].b.grow(l.b); cur = rr;
                }
                bins[lastb].b.grow(cur.b);
                bins[first].enter++; bins[lastb].exit++;
            }
            float right_cost[SPATIAL_BINS]; uint32_t right_n[SPATIAL_BINS]; Box acc; uint32_t k = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bins[b].b); k += bins[b].exit; right_cost[b] = half_area(acc) * k; right_n[b] = k; }
            acc = Box(); k = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bins[b].b); k += bins[b].enter;
                if (k == 0 || right_n[b + 1] == 0) continue;
                const float cost = half_area(acc) * k + right_cost[b + 1];
                if (cost < best.cost) { best.cost = cost; best.axis = axis; best.pos = lo + bin_w * (float)(b + 1); }
            }
        }
        return best;
    }

    int emit_leaf(const std::vector<Ref>& refs) {
        const uint32_t slot = (uint32_t)ordered.size();
        for (size_t i = 0; i < refs.size(); i++) { ordered.push_back(refs[i].tri); last.push_back(i + 1 == refs.size()); }
        return ~(int)slot;
    }
    static void set_box(ctl_bvh_node& n, int which, const Box& b) {
        if (which == 0) { n.a[0] = b.lo.x; n.a[1] = b.hi.x; n.a[2] = b.lo.y; n.a[3] = b.hi.y; n.c[0] = b.lo.z; n.c[1] = b.hi.z; }
        else { n.b[0] = b.lo.x; n.b[1] = b.hi.x; n.b[2] = b.lo.y; n.b[3] = b.hi.y; n.c[2] = b.lo.z; n.c[3] = b.hi.z; }
    }

    // append a privately built subtree; returns its root reference in this context's numbering
    int append(const std::vector<ctl_bvh_node>& sn, const std::vector<uint32_t>& so, const std::vector<uint8_t>& sl, int sub_root, uint32_t parent_idx) {
        const uint32_t node_off = (uint32_t)nodes.size(), slot_off = (uint32_t)ordered.size();
        for (ctl_bvh_node nd : sn) {
            nd.child0 = nd.child0 >= 0 ? nd.child0 + (int)(node_off * 4) : ~(int)((uint32_t)~nd.child0 + slot_off);
            if (nd.child1 != CTL_SENTINEL) nd.child1 = nd.child1 >= 0 ? nd.child1 + (int)(node_off * 4) : ~(int)((uint32_t)~nd.child1 + slot_off);
            nd.parent = nd.parent == 0xfffffffeu ? parent_idx * 4 : nd.parent + node_off * 4;
            nodes.push_back(nd);
        }
        ordered.insert(ordered.end(), so.begin(), so.end()); last.insert(last.end(), sl.begin(), sl.end());
        return sub_root >= 0 ? sub_root + (int)(node_off * 4) : ~(int)((uint32_t)~sub_root + slot_off);
    }

    int build(std::vector<Ref>& refs, const Box& bounds, uint32_t parent, bool is_root, int depth) {
        const uint32_t n = (uint32_t)refs.size();
        if (n == 1 && !is_root) return emit_leaf(refs);
        if (is_root && n == 1) { // single-primitive mesh: root with a sentinel right child (SplitBVHBuilder.cpp:176-189)
            const uint32_t node_idx = (uint32_t)nodes.size(); nodes.push_back(ctl_bvh_node()); memset(&nodes[node_idx], 0, sizeof(ctl_bvh_node));
            const int leaf = emit_leaf(refs);
            ctl_bvh_node& nd = nodes[node_idx]; nd.child0 = leaf; nd.child1 = CTL_SENTINEL; nd.parent = 0xffffffffu;
            set_box(nd, 0, bounds); set_box(nd, 1, Box(V3(0.0f), V3(0.0f)));
            return (int)(node_idx * 4);
        }
        const float area = half_area(bounds);
        const float leaf_cost = CT * (float)n;
        ObjSplit os; SpatialSplit ss;
        if (depth < MAX_DEPTH) {
            os = find_object_split(refs);
            if (os.axis >= 0 && area > 0) {
                const float overlap = half_area(box_intersect(os.lb, os.rb));
                if (overlap >= ALPHA * root_area && n > (uint32_t)max_leaf) ss = find_spatial_split(refs, bounds); // leaf-sized nodes: object splits only
            } else if (area > 0 && n > (uint32_t)max_leaf) ss = find_spatial_split(refs, bounds);
        }
        const float obj_cost = (os.axis >= 0 && area > 0) ? 1.0f + CT * os.cost / area : 3.0e38f;
        const float spa_cost = (ss.axis >= 0 && area > 0) ? 1.0f + CT * ss.cost / area : 3.0e38f;
        const float split_cost = obj_cost < spa_cost ? obj_cost : spa_cost;
        if (!is_root && (int)n <= max_leaf && (leaf_cost <= split_cost || depth >= MAX_DEPTH)) return emit_leaf(refs);

        std::vector<Ref> left, right; Box lb, rb;
        bool done = false;
        if (spa_cost < obj_cost) { // spatial split with unsplitting
            const int axis = ss.axis; const float pos = ss.pos;
            std::vector<const Ref*> straddle;
            for (const Ref& r : refs) {
                if (r.b.hi[axis] <= pos) { left.push_back(r); lb.grow(r.b); }
                else if (r.b.lo[axis] >= pos) { right.push_back(r); rb.grow(r.b); }
                else straddle.push_back(&r);
            }
            for (const Ref* rp : straddle) {
                Ref l, r2; split_ref(*rp, axis, pos, l, r2);
                Box lub = lb, rub = rb, ldb = lb, rdb = rb;
                lub.grow(rp->b); rub.grow(rp->b); ldb.grow(l.b); rdb.grow(r2.b);
                const float nl = (float)left.size(), nr = (float)right.size();
                const float c_left = half_area(lub) * (nl + 1) + half_area(rb) * nr;       // whole reference to the left
                const float c_right = half_area(lb) * nl + half_area(rub) * (nr + 1);      // whole reference to the right
                const float c_split = half_area(ldb) * (nl + 1) + half_area(rdb) * (nr + 1);
                if (c_left <= c_right && c_left <= c_split) { left.push_back(*rp); lb = lub; }
                else if (c_right <= c_split) { right.push_back(*rp); rb = rub; }
                else { left.push_back(l); right.push_back(r2); lb = ldb; rb = rdb; }
            }
            done = !left.empty() && !right.empty() && left.size() < (size_t)n && right.size() < (size_t)n; // each side must shed at least one reference
            if (!done) { left.clear(); right.clear(); lb = Box(); rb = Box(); }
        }
        if (!done && os.axis >= 0) {
            if (os.binned) {
                for (const Ref& r : refs) {
                    int b = (int)((centroid2(r, os.axis) - os.lo) * os.scale); b = b < 0 ? 0 : (b >= OBJ_BINS ? OBJ_BINS - 1 : b);
                    if (b <= os.bin) { left.push_back(r); lb.grow(r.b); } else { right.push_back(r); rb.grow(r.b); }
                }
            } else {
                left.assign(refs.begin(), refs.begin() + os.n_left); right.assign(refs.begin() + os.n_left, refs.end());
                lb = os.lb; rb = os.rb;
            }
            done = !left.empty() && !right.empty();
            if (!done) { left.clear(); right.clear(); lb = Box(); rb = Box(); }
        }
        if (!done) { // identical centroids (or depth cap): split the list in the middle
            if ((int)n <= max_leaf && !is_root) return emit_leaf(refs);
            const uint32_t mid = n / 2;
            left.assign(refs.begin(), refs.begin() + mid); right.assign(refs.begin() + mid, refs.end());
            for (const Ref& r : left) lb.grow(r.b);
            for (const Ref& r : right) rb.grow(r.b);
        }
        std::vector<Ref>().swap(refs); // release before recursing
        const uint32_t node_idx = (uint32_t)nodes.size();
        nodes.push_back(ctl_bvh_node()); memset(&nodes[node_idx], 0, sizeof(ctl_bvh_node));
        int a, b;
        if (n >= (uint32_t)PARALLEL_ABOVE && depth < PARALLEL_DEPTH) {
            // big subtrees are built by two threads into private arrays and appended [left][right]: the numbering (nodes in pre-order, leaf
            // slots in depth-first order) is exactly the sequential one, so the tree does not depend on the thread count
            std::vector<ctl_bvh_node> ln, rn; std::vector<uint32_t> lo_, ro_; std::vector<uint8_t> ll, rl;
            Sbvh SL{v9, max_leaf, ln, lo_, ll}, SR{v9, max_leaf, rn, ro_, rl};
            SL.root_area = SR.root_area = root_area; SL.ALPHA = SR.ALPHA = ALPHA; SL.CT = SR.CT = CT; SL.SWEEP_BELOW = SR.SWEEP_BELOW = SWEEP_BELOW;
            int la = 0, ra = 0;
            std::exception_ptr err;   // an exception in the helper thread (e.g. bad_alloc) is re-thrown here instead of terminating the process
            std::thread th([&]() { try { la = SL.build(left, lb, 0xfffffffeu, false, depth + 1); } catch (...) { err = std::current_exception(); } });
            try { ra = SR.build(right, rb, 0xfffffffeu, false, depth + 1); } catch (...) { th.join(); throw; }
            th.join();
            if (err) std::rethrow_exception(err);
            a = append(ln, lo_, ll, la, node_idx);
            b = append(rn, ro_, rl, ra, node_idx);
        } else {
            a = build(left, lb, node_idx * 4, false, depth + 1);
            b = build(right, rb, node_idx * 4, false, depth + 1);
        }
        ctl_bvh_node& nd = nodes[node_idx];
        nd.child0 = a; nd.child1 = b; nd.parent = is_root ? 0xffffffffu : parent;
        set_box(nd, 0, lb); set_box(nd, 1, rb);
        return (int)(node_idx * 4);
    }
};

} // namespace

// ---- which tree serves path rays better?  Measured, not assumed: spatial splits pay off massively on meshes with long thin triangles (config 4:
// -32 % inner-node visits per path ray) but cost ~10 % on config 2, whose 14 room-sized triangles get chopped into 1 500 references that every
// path -- all of which end on a wall -- has to dig for (GPU A/B in profiles/r01u_builder_ab.log).  The SAH's uniform-ray model cannot see that,
// so the mesh builder builds both candidates and keeps the one that a sample of random-walk rays (the population a path tracer sends through the
// mesh) traverses with fewer steps.
namespace {
struct SampleRay { V3 o, d; };

inline bool ray_tri(const float* t, V3 o, V3 d, float tmin, float& tmax) { // Moeller-Trumbore; only used to shrink tmax like the real traversal does
    const V3 v0(t[0], t[1], t[2]), e1 = V3(t[3], t[4], t[5]) - v0, e2 = V3(t[6], t[7], t[8]) - v0;
    const V3 p = cross(d, e2); const float det = dot(e1, p);
    if (fabsf(det) < 1e-20f) return false;
    const float inv = 1.0f / det; const V3 s = o - v0; const float u = dot(s, p) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const V3 q = cross(s, e1); const float v = dot(d, q) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float tt = dot(e2, q) * inv;
    if (tt <= tmin || tt >= tmax) return false;
    tmax = tt; return true;
}

// closest-hit query through a reference-layout tree (same child-order rule as the device kernel); counts inner pops and triangle tests
struct Tree { const float* v9; const std::vector<ctl_bvh_node>& nodes; const std::vector<uint32_t>& ordered; const std::vector<uint8_t>& last; };
int trace_tree(const Tree& T, const SampleRay& r, float& t_hit, double& inner, double& tris) {
    const V3 id(1.0f / (fabsf(r.d.x) > 1e-20f ? r.d.x : 1e-20f), 1.0f / (fabsf(r.d.y) > 1e-20f ? r.d.y : 1e-20f), 1.0f / (fabsf(r.d.z) > 1e-20f ? r.d.z : 1e-20f));
    float tmax = 3.0e38f; const float tmin = 1e-4f;
    int hit = -1, stack[128], sp = 0, cur = 0;
    for (;;) {
        if (cur >= 0) {
            if (cur == CTL_SENTINEL) { if (!sp) break; cur = stack[--sp]; continue; }
            const ctl_bvh_node& n = T.nodes[(size_t)cur / 4]; inner += 1;
            auto slab = [&](float lx, float hx, float ly, float hy, float lz, float hz, float& t0) {
                const float ax = (lx - r.o.x) * id.x, bx = (hx - r.o.x) * id.x, ay = (ly - r.o.y) * id.y, by = (hy - r.o.y) * id.y, az = (lz - r.o.z) * id.z, bz = (hz - r.o.z) * id.z;
                t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
                const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
                return t1 >= t0;
            };
            float t0a = 0, t0b = 0;
            const bool ha = slab(n.a[0], n.a[1], n.a[2], n.a[3], n.c[0], n.c[1], t0a);
            const bool hb = n.child1 != CTL_SENTINEL && slab(n.b[0], n.b[1], n.b[2], n.b[3], n.c[2], n.c[3], t0b);
            if (!ha && !hb) { if (!sp) break; cur = stack[--sp]; continue; }
            if (ha && hb) { const bool swp = t0b < t0a; cur = swp ? n.child1 : n.child0; if (sp < 127) stack[sp++] = swp ? n.child0 : n.child1; }
            else cur = ha ? n.child0 : n.child1;
        } else {
            for (uint32_t slot = (uint32_t)~cur;; slot++) { tris += 1; if (ray_tri(T.v9 + (size_t)T.ordered[slot] * 9, r.o, r.d, tmin, tmax)) hit = (int)T.ordered[slot]; if (T.last[slot]) break; }
            if (!sp) break; cur = stack[--sp];
        }
    }
    t_hit = tmax;
    return hit;
}

// a small random walk through the mesh: origins inside its box, then up to 4 diffuse-like bounces off whatever is hit -- the population of
// extension and shadow rays a path tracer sends through this mesh (every walk ray is kept as a sample)
std::vector<SampleRay> walk_rays(const Tree& T, const Box& box, int n_walks) {
    uint64_t state = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { state = state * 6364136223846793005ull + 1442695040888963407ull; return (float)((state >> 40) * (1.0 / 16777216.0)); };
    auto dir = [&]() { const float z = 2.0f * rnd() - 1.0f, ph = 6.2831853f * rnd(), rr = sqrtf(fmaxf(0.0f, 1.0f - z * z)); return V3(rr * cosf(ph), rr * sinf(ph), z); };
    std::vector<SampleRay> rays;
    for (int i = 0; i < n_walks; i++) {
        SampleRay r{V3(box.lo.x + (box.hi.x - box.lo.x) * rnd(), box.lo.y + (box.hi.y - box.lo.y) * rnd(), box.lo.z + (box.hi.z - box.lo.z) * rnd()), dir()};
        for (int depth = 0; depth < 5; depth++) {
            rays.push_back(r);
            float t; double a = 0, b = 0;
            const int tri = trace_tree(T, r, t, a, b);
            if (tri < 0) break;
            const float* p = T.v9 + (size_t)tri * 9;
            V3 n = normalize(cross(V3(p[3], p[4], p[5]) - V3(p[0], p[1], p[2]), V3(p[6], p[7], p[8]) - V3(p[0], p[1], p[2])));
            if (dot(n, r.d) > 0) n = -n;
            V3 d = dir(); if (dot(d, n) < 0) d = -d;
            r = SampleRay{r.o + r.d * t + n * 1e-4f, d};
        }
    }
    return rays;
}
double traversal_cost(const Tree& T, const std::vector<SampleRay>& rays) {
    double inner = 0, tris = 0; float t;
    for (const SampleRay& r : rays) trace_tree(T, r, t, inner, tris);
    return inner + 0.8 * tris;
}
} // namespace

static void build_sbvh_alpha(const float* verts9, uint32_t n_tris, int max_leaf, float alpha, std::vector<ctl_bvh_node>& nodes_out, std::vector<uint32_t>& ordered, std::vector<uint8_t>& last);

// Post-pass: tree rotations (Kensler, "Tree rotations for improving bounding volume hierarchies",
// 2008) on the finished node array.  For a node with children L, R the four exchanges "child <-> grandchild on the other side" are tried and the one
// that shrinks the surface area of the rebuilt inner child most is kept; sweeps run children-first until nothing improves.  Leaves (index runs) are
// never touched, so the reference set and the hits are unchanged; the array is re-laid out in pre-order at the end.
namespace {
inline Box child_box(const ctl_bvh_node& n, int w) { return w == 0 ? Box(V3(n.a[0], n.a[2], n.c[0]), V3(n.a[1], n.a[3], n.c[1])) : Box(V3(n.b[0], n.b[2], n.c[2]), V3(n.b[1], n.b[3], n.c[3])); }
inline void put_box(ctl_bvh_node& n, int w, const Box& b) {
    if (w == 0) { n.a[0] = b.lo.x; n.a[1] = b.hi.x; n.a[2] = b.lo.y; n.a[3] = b.hi.y; n.c[0] = b.lo.z; n.c[1] = b.hi.z; }
    else { n.b[0] = b.lo.x; n.b[1] = b.hi.x; n.b[2] = b.lo.y; n.b[3] = b.hi.y; n.c[2] = b.lo.z; n.c[3] = b.hi.z; }
}
inline int& child_ref(ctl_bvh_node& n, int w) { return w == 0 ? n.child0 : n.child1; }
inline bool is_inner(int c) { return c >= 0 && c != CTL_SENTINEL; }

size_t rotate_tree(std::vector<ctl_bvh_node>& nodes, int max_sweeps) {
    if (nodes.size() < 3) return 0;
    size_t total = 0;
    // children-first order: reverse of a pre-order walk from the root (recomputed every sweep: rotations move sub-trees)
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        std::vector<uint32_t> order; order.reserve(nodes.size());
        std::vector<uint32_t> st = {0};
        while (!st.empty()) { const uint32_t i = st.back(); st.pop_back(); order.push_back(i); if (is_inner(nodes[i].child0)) st.push_back((uint32_t)nodes[i].child0 / 4); if (is_inner(nodes[i].child1)) st.push_back((uint32_t)nodes[i].child1 / 4); }
        size_t done = 0;
        for (size_t k = order.size(); k-- > 0;) {
            ctl_bvh_node& N = nodes[order[k]];
            if (N.child1 == CTL_SENTINEL) continue;
            float best = 0.0f; int best_side = -1, best_g = -1;   // exchange child (1 - side) with grandchild g of child `side`; best = most negative area change
            for (int side = 0; side < 2; side++) {
                const int c = child_ref(N, side);
                if (!is_inner(c)) continue;
                const ctl_bvh_node& C = nodes[(uint32_t)c / 4];
                if (C.child1 == CTL_SENTINEL) continue;
                const Box other = child_box(N, 1 - side);
                const float old_area = child_box(N, side).area();
                for (int g = 0; g < 2; g++) {   // g goes up, `other` takes its place next to grandchild 1 - g
                    Box nb = other; nb.grow(child_box(C, 1 - g));
                    const float delta = nb.area() - old_area;
                    if (delta < -1e-5f * old_area && delta < best) { best = delta; best_side = side; best_g = g; }
                }
            }
            // grandchild <-> grandchild across the two children (both inner): L keeps its other grandchild and gets R's g2, R likewise
            int best_g1 = -1, best_g2 = -1;
            if (is_inner(N.child0) && is_inner(N.child1)) {
                const ctl_bvh_node& L = nodes[(uint32_t)N.child0 / 4]; const ctl_bvh_node& R = nodes[(uint32_t)N.child1 / 4];
                if (L.child1 != CTL_SENTINEL && R.child1 != CTL_SENTINEL) {
                    const float old_area = child_box(N, 0).area() + child_box(N, 1).area();
                    for (int g1 = 0; g1 < 2; g1++) for (int g2 = 0; g2 < 2; g2++) {
                        Box lb = child_box(L, 1 - g1); lb.grow(child_box(R, g2));
                        Box rb = child_box(R, 1 - g2); rb.grow(child_box(L, g1));
                        const float delta = lb.area() + rb.area() - old_area;
                        if (delta < -1e-5f * old_area && delta < best) { best = delta; best_g1 = g1; best_g2 = g2; }
                    }
                }
            }
            if (best_g1 >= 0) {
                ctl_bvh_node& L = nodes[(uint32_t)N.child0 / 4]; ctl_bvh_node& R = nodes[(uint32_t)N.child1 / 4];
                const int a_ref = child_ref(L, best_g1), b_ref = child_ref(R, best_g2);
                const Box a_box = child_box(L, best_g1), b_box = child_box(R, best_g2);
                child_ref(L, best_g1) = b_ref; put_box(L, best_g1, b_box);
                child_ref(R, best_g2) = a_ref; put_box(R, best_g2, a_box);
                Box lb = child_box(L, 0); lb.grow(child_box(L, 1)); put_box(N, 0, lb);
                Box rb = child_box(R, 0); rb.grow(child_box(R, 1)); put_box(N, 1, rb);
                done++;
                continue;
            }
            if (best_side < 0) continue;
            const int side = best_side, g = best_g;
            ctl_bvh_node& C = nodes[(uint32_t)child_ref(N, side) / 4];
            const int up = child_ref(C, g), down = child_ref(N, 1 - side);
            const Box up_box = child_box(C, g), down_box = child_box(N, 1 - side);
            child_ref(C, g) = down; put_box(C, g, down_box);
            child_ref(N, 1 - side) = up; put_box(N, 1 - side, up_box);
            Box cb = child_box(C, 0); cb.grow(child_box(C, 1)); put_box(N, side, cb);
            done++;
        }
        total += done;
        if (!done) break;
    }
    // pre-order re-layout, parents re-linked
    std::vector<ctl_bvh_node> out; out.reserve(nodes.size());
    std::vector<std::pair<uint32_t, uint32_t>> st;   // (old index, new index)
    out.push_back(nodes[0]); out[0].parent = 0xffffffffu; st.emplace_back(0u, 0u);
    while (!st.empty()) {
        const auto cur = st.back(); st.pop_back();
        int ch[2] = {nodes[cur.first].child0, nodes[cur.first].child1};
        for (int k = 1; k >= 0; k--) {
            if (!is_inner(ch[k])) continue;
            const uint32_t nu = (uint32_t)out.size();
            out.push_back(nodes[(uint32_t)ch[k] / 4]); out.back().parent = cur.second * 4;
            st.emplace_back((uint32_t)ch[k] / 4, nu);
            ch[k] = (int)(nu * 4);
        }
        out[cur.second].child0 = ch[0]; out[cur.second].child1 = ch[1];
    }
    nodes.swap(out);
    return total;
}
} // namespace

// Insertion-based optimisation (Bittner, Hapala, Havran: "Fast insertion-based optimization of bounding volume hierarchies", CGF 2013), the variant that
// re-inserts whole sub-trees: per pass the nodes with the largest boxes are detached one at a time (their parent is spliced out) and re-attached where a
// branch-and-bound search over the tree finds the smallest total surface-area increase; the old place is among the candidates, so a step never makes
// the tree worse.  Works on a pointer form of the tree (inner nodes + one node per leaf reference), written back in pre-order.  Leaves are untouched.
static size_t reinsert_tree(std::vector<ctl_bvh_node>& nodes, int passes, float fraction) {
    if (nodes.size() < 4 || nodes[0].child1 == CTL_SENTINEL) return 0;
    struct T { Box box; int parent, left, right, leafref; float area; };
    std::vector<T> t; t.reserve(nodes.size() * 2 + 1);
    const int n_inner = (int)nodes.size();
    t.resize(n_inner);
    for (int i = 0; i < n_inner; i++) {
        int ch[2] = {nodes[i].child0, nodes[i].child1};
        for (int k = 0; k < 2; k++) {
            int id;
            if (is_inner(ch[k])) { id = ch[k] / 4; t[id].box = child_box(nodes[i], k); }
            else { id = (int)t.size(); t.push_back(T{child_box(nodes[i], k), i, -1, -1, ch[k], 0.0f}); }
            t[id].parent = i;
            (k == 0 ? t[i].left : t[i].right) = id;
        }
        t[i].leafref = 0;
    }
    int root = 0; t[0].parent = -1; t[0].box = t[t[0].left].box; t[0].box.grow(t[t[0].right].box);
    auto finite_area = [](const Box& b) { const float a = b.area(); return a >= 0.0f && a < 3.0e38f ? a : 0.0f; };   // NaN / inf vertices must not poison the orderings
    for (T& n : t) n.area = finite_area(n.box);
    auto refit_up = [&](int i) { for (; i >= 0; i = t[i].parent) { Box b = t[t[i].left].box; b.grow(t[t[i].right].box); const float a = finite_area(b); if (a == t[i].area && b.lo.x == t[i].box.lo.x && b.lo.y == t[i].box.lo.y && b.lo.z == t[i].box.lo.z && b.hi.x == t[i].box.hi.x && b.hi.y == t[i].box.hi.y && b.hi.z == t[i].box.hi.z) break; t[i].box = b; t[i].area = a; } };
    size_t moved = 0;
    std::vector<int> cand;
    std::vector<std::pair<float, int>> heap;   // (-induced cost, node)
    for (int pass = 0; pass < passes; pass++) {
        cand.clear();
        for (int i = 0; i < (int)t.size(); i++) if (i != root && t[i].parent != root) cand.push_back(i);
        const size_t take = std::max<size_t>(1, (size_t)(cand.size() * fraction));
        std::partial_sort(cand.begin(), cand.begin() + std::min(take, cand.size()), cand.end(), [&](int a, int b) { return t[a].area != t[b].area ? t[a].area > t[b].area : a < b; });
        cand.resize(std::min(take, cand.size()));
        size_t moved_pass = 0;
        for (int X : cand) {
            const int P = t[X].parent;
            if (X == root || P < 0 || P == root) continue;
            const int G = t[P].parent, S = t[P].left == X ? t[P].right : t[P].left;
            // detach X: S takes P's place under G
            (t[G].left == P ? t[G].left : t[G].right) = S; t[S].parent = G;
            refit_up(G);
            // branch-and-bound for the best sibling Y of X
            const float ax = t[X].area;
            float best_cost = 3.0e38f; int best = -1;
            heap.clear(); heap.emplace_back(-0.0f, root);
            while (!heap.empty()) {
                std::pop_heap(heap.begin(), heap.end());
                const float ci = -heap.back().first; const int Y = heap.back().second; heap.pop_back();
                if (ci + ax >= best_cost) break;
                Box u = t[Y].box; u.grow(t[X].box);
                const float cd = u.area(), total = ci + cd;
                if (total < best_cost) { best_cost = total; best = Y; }
                const float child_ci = total - t[Y].area;
                if (t[Y].left >= 0 && child_ci + ax < best_cost) {
                    heap.emplace_back(-child_ci, t[Y].left); std::push_heap(heap.begin(), heap.end());
                    heap.emplace_back(-child_ci, t[Y].right); std::push_heap(heap.begin(), heap.end());
                }
            }
            // attach: P becomes the parent of (best, X) where best was
            if (best < 0) best = S;   // only with non-finite boxes (every comparison false): back to where it was
            const int Y = best, YP = t[Y].parent;
            t[P].parent = YP; t[P].left = Y; t[P].right = X; t[Y].parent = P; t[X].parent = P;
            if (YP < 0) root = P; else (t[YP].left == Y ? t[YP].left : t[YP].right) = P;
            t[P].box = t[Y].box; t[P].box.grow(t[X].box); t[P].area = finite_area(t[P].box);
            refit_up(YP);
            if (Y != S) moved_pass++;
        }
        moved += moved_pass;
        if (moved_pass * 200 < cand.size()) break;   // fewer than 0.5 % of the candidates found a better place
    }
    // depth guard for the 64-entry traversal stack, then pre-order write-back
    {
        int max_depth = 0; std::vector<std::pair<int, int>> st = {{root, 1}};
        while (!st.empty()) { const auto c = st.back(); st.pop_back(); if (c.second > max_depth) max_depth = c.second; if (t[c.first].left >= 0) { st.emplace_back(t[c.first].left, c.second + 1); st.emplace_back(t[c.first].right, c.second + 1); } }
        if (max_depth > 50) return 0;   // keep the builder's tree
    }
    std::vector<ctl_bvh_node> out; out.reserve(nodes.size());
    std::vector<std::pair<int, uint32_t>> st;   // (tree node, array index)
    auto emit = [&](int id, uint32_t parent4) { ctl_bvh_node n; memset(&n, 0, sizeof(n)); n.parent = parent4; out.push_back(n); return (uint32_t)out.size() - 1; };
    st.emplace_back(root, emit(root, 0xffffffffu));
    while (!st.empty()) {
        const auto cur = st.back(); st.pop_back();
        const int kids[2] = {t[cur.first].left, t[cur.first].right};
        int refs[2];
        for (int k = 1; k >= 0; k--) {
            if (t[kids[k]].left >= 0) { const uint32_t nu = emit(kids[k], cur.second * 4); st.emplace_back(kids[k], nu); refs[k] = (int)(nu * 4); }
            else refs[k] = t[kids[k]].leafref;
        }
        ctl_bvh_node& n = out[cur.second];
        n.child0 = refs[0]; n.child1 = refs[1];
        put_box(n, 0, t[kids[0]].box); put_box(n, 1, t[kids[1]].box);
    }
    if (out.size() != nodes.size()) return 0;   // cannot happen: the number of inner nodes is invariant
    nodes.swap(out);
    return moved;
}

// Post-pass of every mesh tree: sub-tree re-insertion (2 passes over all nodes), then up to 8 sweeps of tree rotations (CTL_SBVH_ROTATE=<sweeps> overrides,
// 0 = neither).  Together they take the oracle's path rays on config 2 from 27.9 to 25.1 inner nodes per ray (-7.4 % algorithmic bytes), config 4 -4.4 %.  ~1 700 rotations on the 57 K-node tree of
// config 2 take the oracle's path rays from 27.9 to 26.2 inner nodes per ray (-4.8 % algorithmic bytes); hits, images and ray counts are unchanged.
static void finish_tree(std::vector<ctl_bvh_node>& nodes) {
    const char* r = getenv("CTL_SBVH_ROTATE");
    const int sweeps = r ? atoi(r) : 8;
    if (sweeps <= 0) return;   // CTL_SBVH_ROTATE=0: the builder's tree as it is (A/B)
    const char* q = getenv("CTL_SBVH_REINSERT");   // passes of sub-tree re-insertion before the rotations (0 = off); CTL_SBVH_REINSERT_FRAC = share of the nodes tried per pass
    // (by default only for trees up to 4 M nodes: the searches of a heavily overlapping 235 K-reference mesh already take 8 s; rotations always run)
    const int ins_passes = q ? atoi(q) : 2;   // measured: 2 passes reach the result of 4 within 0.3 % (config 2: 2 060 vs 2 065 bytes per path ray)
    const int rounds = getenv("CTL_SBVH_ROUNDS") ? atoi(getenv("CTL_SBVH_ROUNDS")) : 1;   // experiments: (re-insertion, rotations) repeated
    const std::vector<ctl_bvh_node> before = nodes;
    for (int round = 0; round < rounds; round++) {
    const size_t n_ins = ins_passes > 0 && (q || nodes.size() <= 4000000) ? reinsert_tree(nodes, ins_passes, getenv("CTL_SBVH_REINSERT_FRAC") ? (float)atof(getenv("CTL_SBVH_REINSERT_FRAC")) : 1.0f) : 0;
    const size_t n_rot = rotate_tree(nodes, sweeps);
    if (getenv("CTL_SBVH_VERBOSE")) fprintf(stderr, "  re-insertions: %zu, tree rotations: %zu\n", n_ins, n_rot);
    }
    // the traversal stack holds 64 entries for the scene level and the mesh level together: an optimised tree deeper than 52 is not worth it
    int max_depth = 0; std::vector<std::pair<uint32_t, int>> st = {{0u, 1}};
    while (!st.empty()) {
        const auto c = st.back(); st.pop_back();
        if (c.second > max_depth) max_depth = c.second;
        if (is_inner(nodes[c.first].child0)) st.emplace_back((uint32_t)nodes[c.first].child0 / 4, c.second + 1);
        if (is_inner(nodes[c.first].child1)) st.emplace_back((uint32_t)nodes[c.first].child1 / 4, c.second + 1);
    }
    if (max_depth > 52) nodes = before;
}

void optimize_bvh(std::vector<ctl_bvh_node>& nodes) { finish_tree(nodes); }

void build_sbvh(const float* verts9, uint32_t n_tris, int max_leaf, std::vector<ctl_bvh_node>& nodes_out, std::vector<uint32_t>& ordered, std::vector<uint8_t>& last) {
    if (getenv("CTL_SBVH_ALPHA")) { build_sbvh_alpha(verts9, n_tris, max_leaf, (float)atof(getenv("CTL_SBVH_ALPHA")), nodes_out, ordered, last); finish_tree(nodes_out); return; } // experiments: no selection
    build_sbvh_alpha(verts9, n_tris, max_leaf, 1e-5f, nodes_out, ordered, last);   // the published default (and the reference's BuildParams::splitAlpha)
    if (ordered.size() == n_tris || n_tris < 64) { finish_tree(nodes_out); return; } // no reference was split: nothing to choose
    std::vector<ctl_bvh_node> n2; std::vector<uint32_t> o2; std::vector<uint8_t> l2;
    build_sbvh_alpha(verts9, n_tris, max_leaf, 3.0e38f, n2, o2, l2);               // object splits only
    const Tree split{verts9, nodes_out, ordered, last}, plain{verts9, n2, o2, l2};
    Box box; for (size_t i = 0; i < (size_t)n_tris * 3; i++) box.grow(V3(verts9[3 * i], verts9[3 * i + 1], verts9[3 * i + 2]));
    const std::vector<SampleRay> rays = walk_rays(plain, box, 4096);
    const double c_split = traversal_cost(split, rays), c_plain = traversal_cost(plain, rays);
    if (getenv("CTL_SBVH_VERBOSE")) fprintf(stderr, "mesh %u tris: split tree %zu refs cost %.0f, plain tree cost %.0f over %zu walk rays -> %s\n", n_tris, ordered.size(), c_split, c_plain, rays.size(), c_plain < c_split ? "plain" : "split");
    if (c_plain < c_split) { nodes_out.swap(n2); ordered.swap(o2); last.swap(l2); }
    finish_tree(nodes_out);
}

static void build_sbvh_alpha(const float* verts9, uint32_t n_tris, int max_leaf, float alpha, std::vector<ctl_bvh_node>& nodes_out, std::vector<uint32_t>& ordered, std::vector<uint8_t>& last) {
    nodes_out.clear(); ordered.clear(); last.clear();
    if (!n_tris) return;
    Sbvh S{verts9, max_leaf, nodes_out, ordered, last};
    S.ALPHA = alpha;
    std::vector<Ref> refs(n_tris); Box all;
    for (uint32_t t = 0; t < n_tris; t++) {
        refs[t].tri = t;
        for (int k = 0; k < 3; k++) refs[t].b.grow(S.vert(t, k));
        all.grow(refs[t].b);
    }
    S.root_area = half_area(all);
    if (const char* a = getenv("CTL_SBVH_CT")) S.CT = (float)atof(a);
    if (const char* a = getenv("CTL_SBVH_SWEEP")) S.SWEEP_BELOW = atoi(a);
    S.build(refs, all, 0xffffffffu, true, 0);
}

} // namespace ctlb
