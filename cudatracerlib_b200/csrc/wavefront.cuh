// wavefront.cuh -- the wavefront stages of the path tracer (device kernels).
//
// Replaces the megakernel pathKernel2 -> PathTrace<DIRECT> (Integrators/PathTracer.cu:10-113,182-194)
// by stages  generate -> [ intersect(ext) -> shade (+NEE emit) -> intersect(shadow, any-hit, adds the
// pending NEE term) ] x bounces -> finish (Image::AddSample, Engine/Image.cu:22-44),
// with warp-ballot stream compaction of surviving paths between bounces.  Per-path arithmetic and
// random-number bookkeeping are those of PathTrace.
#pragma once
#include "device/shading.cuh"
#include "device/traverse_persistent.cuh"
#include <cfloat>

namespace ctld {

struct Window { // which pixels a pass renders
    int mode;   // 0 = rectangle [x0,x1) x [y0,y1); 1 = interleaved tiles
    int x0, y0, x1, y1;
    int tile_w, tile_h, part, n_parts, tiles_x, tiles_y;
    int n_slots;   // pixels of the window (paths per pass)
    int n_passes;  // passes rendered by this wavefront; path slot = pass * n_slots + pixel slot
    int warp_blocks; // mode 1: slots inside a tile run over 8x4 pixel blocks (tile_w % 8 == 0 and tile_h % 4 == 0) instead of rows
    int tab0;      // sample-table set of this wavefront's first pass (pass k of the wavefront draws from set tab0 + k)
};

// SoA path state, indexed by path id (= window slot)
struct PathState {
    float4* cf;   // throughput rgb, brdf_scattering_pdf
    float4* cl;   // radiance rgb, -
    float4* nor;  // last_nor xyz, packed (depth | specular<<8 | i1<<9 | i2<<19)
    float4* px;   // pX, pY, sampler index bits, (pass-in-batch + 1) as uint bits (0 = no path)
    float4* wo;   // StopZeroThroughput=0 only (else null): the last sampled local direction -- the reference's BSDFSamplingRecord lives across the
                  // loop, so a failed sample continues along the previous wo (PathTracer.cu:44, 80-93)
};

struct Queues {
    float4* rays_in;  uint32_t* path_in;   // extension rays of this bounce (2 x float4 each) + owning path
    float4* rays_out; uint32_t* path_out;  // next bounce
    float4* hit_a;    uint32_t* hit_node;  // (dist,u,v,tri) + node, by queue position
    float4* sh_rays;  float4* sh_payload;  // shadow rays + (pending radiance rgb, path id)
    uint32_t* keys_out; unsigned* hist;    // SortMode 1: sort key of every emitted extension ray + bucket histogram (else null)
    const uint32_t* order;                 // SortMode 2: queue positions grouped by material class (else null = identity)
};

CTL_DEV unsigned lane_id() { return threadIdx.x & 31; }

// all 32 lanes must call
CTL_DEV int warp_append(bool pred, unsigned* counter) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (!mask) return -1;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? (int)(base + __popc(mask & ((1u << lane_id()) - 1u))) : -1;
}

CTL_DEV uint32_t pack_ctl(int depth, bool spec, unsigned i1, unsigned i2) { return (uint32_t)depth | (spec ? 256u : 0u) | (i1 << 9) | (i2 << 19); }

CTL_DEV bool slot_to_pixel(const Window& W, int s, int img_w, int img_h, int& x, int& y) {
    if (W.mode == 0) {
        const int bw = W.x1 - W.x0;
        x = W.x0 + s % bw; y = W.y0 + s / bw;
    } else {
        const int per = W.tile_w * W.tile_h;
        const int t_local = s / per, r = s % per;
        const int tile = W.part + t_local * W.n_parts;
        const int tx = tile % W.tiles_x, ty = tile / W.tiles_x;
        if (W.warp_blocks) { // a warp = an 8x4 pixel block instead of 32 pixels of one row: tighter ray bundles at bounce 0 and, because compaction
                             // keeps the queue order, neighbouring origins at the later bounces (pixels, samples and results are unchanged)
            const int b = r >> 5, l = r & 31, bpr = W.tile_w >> 3;
            x = tx * W.tile_w + (b % bpr) * 8 + (l & 7); y = ty * W.tile_h + (b / bpr) * 4 + (l >> 3);
        } else { x = tx * W.tile_w + r % W.tile_w; y = ty * W.tile_h + r / W.tile_w; }
    }
    return x >= 0 && y >= 0 && x < img_w && y < img_h;
}

// ---- sample tables on the device: SamplingSequenceGeneratorHost<IndependantSamplingSequenceGenerator>::Compute
// (Kernel/Sampler.h:36-85) + CudaRNG host twin (Base/CudaRandom.h:108-127), bit-identical, one thread per sequence.
// states: kNumSeq x 6 words, the stream state right before this thread's slice of the NEXT pass to generate (updated);
// jump: 160 x 5 words = skip (4096*90 - 90) draws (csrc/sampler_tables.h).  Fills n_passes consecutive table sets.
__global__ void __launch_bounds__(128) k_gen_tables(uint32_t* __restrict__ states, const uint32_t* __restrict__ jump, int n_passes, float* __restrict__ d1, float2* __restrict__ d2) {
    __shared__ uint32_t sj[160 * 5];
    for (int i = threadIdx.x; i < 160 * 5; i += blockDim.x) sj[i] = jump[i];
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N_SEQ) return;
    uint32_t v0 = states[s * 6], v1 = states[s * 6 + 1], v2 = states[s * 6 + 2], v3 = states[s * 6 + 3], v4 = states[s * 6 + 4], d = states[s * 6 + 5];
    auto next_float = [&]() {
        const uint32_t t = v0 ^ (v0 >> 2);
        v0 = v1; v1 = v2; v2 = v3; v3 = v4;
        v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
        d += 362437u;
        const float f = __uint2float_rn(v4 + d) * 2.3283064e-10f + (2.3283064e-10f / 2.0f); // curand_uniform, no FMA contraction (-fmad=false)
        return f * (1 - 1e-5f);
    };
    for (int p = 0; p < n_passes; p++) {
        float* t1 = d1 + (size_t)p * N_SEQ * SEQ_LEN; float2* t2 = d2 + (size_t)p * N_SEQ * SEQ_LEN;
        for (int i = 0; i < SEQ_LEN; i++) t1[i * N_SEQ + s] = next_float();
        for (int i = 0; i < SEQ_LEN; i++) { const float y = next_float(), x = next_float(); t2[i * N_SEQ + s] = make_float2(x, y); } // y is drawn first (sampler_tables.h)
        // jump to this sequence's slice of the next pass
        uint32_t in[5] = {v0, v1, v2, v3, v4}, r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
        for (int w = 0; w < 5; w++)
            for (int b = 0; b < 32; b++)
                if ((in[w] >> b) & 1u) { const uint32_t* row = sj + (w * 32 + b) * 5; r0 ^= row[0]; r1 ^= row[1]; r2 ^= row[2]; r3 ^= row[3]; r4 ^= row[4]; }
        v0 = r0; v1 = r1; v2 = r2; v3 = r3; v4 = r4;
        d += 362437u * (uint32_t)(N_SEQ * 3 * SEQ_LEN - 3 * SEQ_LEN);
    }
    states[s * 6] = v0; states[s * 6 + 1] = v1; states[s * 6 + 2] = v2; states[s * 6 + 3] = v3; states[s * 6 + 4] = v4; states[s * 6 + 5] = d;
}

// ---- generate: pathKernel2 prologue (PathTracer.cu:184-190) -------------------
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ DScene S, const __grid_constant__ Window W, PathState st,
                                                   float4* rays, uint32_t* paths, unsigned* q_count) {
    const int n_total = W.n_slots * W.n_passes;
    const int n_round = (n_total + 31) & ~31;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_round; s += gridDim.x * blockDim.x) {
        int x = 0, y = 0;
        const bool in_range = s < n_total;
        const unsigned pass = (unsigned)(s / W.n_slots) + (unsigned)W.tab0;
        const bool valid = in_range && slot_to_pixel(W, s % W.n_slots, S.img_w, S.img_h, x, y);
        V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
        if (valid) {
            Sampler rng; rng.idx = (unsigned)(y * S.img_w + x); rng.i1 = 0; rng.i2 = 0; rng.tab = pass;
            const float2 j = rng.f2(S);
            const float pX = (float)x + j.x, pY = (float)y + j.y;
            rng.f2(S); // aperture sample: drawn, unused by the pinhole camera
            camera_ray(S, pX, pY, o, d);
            st.px[s] = make_float4(pX, pY, __uint_as_float(rng.idx), __uint_as_float(pass + 1u));
            st.cf[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            st.cl[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            st.nor[s] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(pack_ctl(0, false, rng.i1, rng.i2)));
        } else if (in_range) {
            st.px[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        const int jq = warp_append(valid, q_count);
        if (valid) {
            rays[2 * jq] = make_float4(o.x, o.y, o.z, S.ray_eps);
            rays[2 * jq + 1] = make_float4(d.x, d.y, d.z, FLT_MAX);
            paths[jq] = (uint32_t)s;
        }
    }
}

// ---- intersect ------------------------------------------------------------------
// MODE 0: wavefront extension (closest hit -> hit_a / hit_node)
// MODE 1: wavefront shadow (any hit; unoccluded => cl[path] += pending)
// MODE 2: API, 16-byte traversalResult, box/tri lower bound = ray.tmin  (== intersectKernel, TraceHelper.cu:326-734)
// MODE 3: API, ctl_trace_result, t in (rayEps, FLT_MAX)                  (== traceRay, TraceHelper.cu:174-180)
#ifndef CTL_SIMPLE_MIN_BLOCKS
#define CTL_SIMPLE_MIN_BLOCKS 10
#endif
template <int MODE, bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128, CTL_SIMPLE_MIN_BLOCKS) k_intersect_simple(const __grid_constant__ DScene S, const float4* __restrict__ rays, const unsigned* __restrict__ n_ptr, int n_fixed,
                                                    unsigned* work_ctr, float4* __restrict__ hit_a, uint32_t* __restrict__ hit_node,
                                                    const float4* __restrict__ sh_payload, float4* __restrict__ cl,
                                                    void* __restrict__ api_out, unsigned long long* visit_out) {
    const int n = n_ptr ? (int)*n_ptr : n_fixed;
    VisitCounters<COUNT> cnt;
    for (;;) {
        unsigned base = 0;
        if (lane_id() == 0) base = atomicAdd(work_ctr, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((int)base >= n) break;
        const int i = (int)base + (int)lane_id();
        if (i < n) {
            const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
            Hit hit; hit.u = hit.v = 0.0f; hit.tri = 0xffffffffu; hit.node = 0xffffffffu;
            float tri_lo, box_lo;
            if (MODE == 3) { tri_lo = S.ray_eps; box_lo = 0.0f; hit.dist = FLT_MAX; }
            else if (MODE == 2) { tri_lo = ro.w; box_lo = ro.w; hit.dist = rd.w; }
            else { tri_lo = ro.w; box_lo = 0.0f; hit.dist = rd.w; }
            trace_ray<ANY_HIT, COUNT>(S, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), tri_lo, box_lo, hit, cnt);
            if (MODE == 0) {
                hit_a[i] = make_float4(hit.dist, hit.u, hit.v, __uint_as_float(hit.tri));
                hit_node[i] = hit.node;
            } else if (MODE == 1) {
                if (hit.tri == 0xffffffffu) {
                    const float4 pl = __ldg(sh_payload + i);
                    const uint32_t p = __float_as_uint(pl.w);
                    float4 c = cl[p];
                    c.x = c.x + pl.x; c.y = c.y + pl.y; c.z = c.z + pl.z;
                    cl[p] = c;
                }
            } else if (MODE == 2) {
                uint4 res = make_uint4(__float_as_uint(hit.dist), 0xffffffffu, 0xffffffffu, 0u);
                if (hit.tri != 0xffffffffu) {
                    res.y = hit.node; res.z = hit.tri;
                    const unsigned short xd = (unsigned short)(hit.u * 65535), yd = (unsigned short)(hit.v * 65535); // TraceHelper.cu:726-727
                    res.w = ((uint32_t)yd << 16) | (uint32_t)xd;
                }
                ((uint4*)api_out)[i] = res;
            } else {
                float* out = (float*)api_out + (size_t)i * 5;
                out[0] = hit.dist; out[1] = hit.u; out[2] = hit.v; out[3] = __uint_as_float(hit.tri); out[4] = __uint_as_float(hit.node);
            }
        }
    }
    if (COUNT) {
        VisitCounters<true>& c = (VisitCounters<true>&)cnt;
        unsigned a = c.inner, b = c.tris, e = c.inst;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
        if (lane_id() == 0) { atomicAdd(visit_out, (unsigned long long)a); atomicAdd(visit_out + 1, (unsigned long long)b); atomicAdd(visit_out + 2, (unsigned long long)e); }
    }
}

// Production traversal kernel: persistent warps, phase-scheduled (device/traverse_persistent.cuh). Same MODEs as above.
#ifndef CTL_TRAV_MIN_BLOCKS
#define CTL_TRAV_MIN_BLOCKS 1
#endif
template <int MODE, bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128, CTL_TRAV_MIN_BLOCKS) k_intersect(const __grid_constant__ DScene S, const __grid_constant__ TravTune tune, const float4* __restrict__ rays, const unsigned* __restrict__ n_ptr, int n_fixed,
                                                    unsigned* work_ctr, float4* __restrict__ hit_a, uint32_t* __restrict__ hit_node,
                                                    const float4* __restrict__ sh_payload, float4* __restrict__ cl,
                                                    void* __restrict__ api_out, unsigned long long* visit_out) {
    const int n = n_ptr ? (int)*n_ptr : n_fixed;
    VisitCounters<COUNT> cnt;
    TravOut out = {hit_a, hit_node, sh_payload, cl, api_out, nullptr, 0, nullptr};
    trace_persistent<MODE, ANY_HIT, COUNT>(S, rays, n, work_ctr, out, tune, cnt);
    if (COUNT) {
        VisitCounters<true>& c = (VisitCounters<true>&)cnt;
        unsigned a = c.inner, b = c.tris, e = c.inst;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
        if (lane_id() == 0) { atomicAdd(visit_out, (unsigned long long)a); atomicAdd(visit_out + 1, (unsigned long long)b); atomicAdd(visit_out + 2, (unsigned long long)e); }
    }
}

// ---- ray sorting between bounces (the wavefront scheduler's coherence step) -----------------------------------------
// key = direction octant (3 bits, major) | 15-bit Morton code of the origin's cell in a 32^3 grid over the scene box:
// rays that start in the same cell and head into the same octant take near-identical top-of-tree paths, so a warp of
// the traversal kernel stays converged longer and its node fetches hit the same L1 lines.  Counting sort: histogram
// (atomics in k_shade) -> exclusive scan (one block) -> scatter of the 36-byte (ray, path id) records.
constexpr int SORT_BUCKETS = 1 << 18;
CTL_DEV uint32_t spread5(uint32_t x) { x &= 31u; x = (x | (x << 8)) & 0x100fu; x = (x | (x << 4)) & 0x10c3u; x = (x | (x << 2)) & 0x1249u; return x; }
CTL_DEV uint32_t ray_sort_key(const DScene& S, V3 o, V3 d) {
    const float fx = (o.x - S.box_min[0]) * S.box_inv_extent[0], fy = (o.y - S.box_min[1]) * S.box_inv_extent[1], fz = (o.z - S.box_min[2]) * S.box_inv_extent[2];
    const uint32_t cx = (uint32_t)fminf(fmaxf(fx * 32.0f, 0.0f), 31.0f), cy = (uint32_t)fminf(fmaxf(fy * 32.0f, 0.0f), 31.0f), cz = (uint32_t)fminf(fmaxf(fz * 32.0f, 0.0f), 31.0f);
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    return (oct << 15) | spread5(cx) | (spread5(cy) << 1) | (spread5(cz) << 2);
}
// exclusive scan of hist[SORT_BUCKETS] into offsets; hist is zeroed for the next bounce.  One block of 1024 threads.
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned* __restrict__ hist, unsigned* __restrict__ offsets) {
    __shared__ unsigned warp_sums[32];
    constexpr int PER = SORT_BUCKETS / 1024;
    const int t = threadIdx.x;
    unsigned local = 0;
    for (int i = 0; i < PER; i++) local += hist[t * PER + i];
    unsigned incl = local;
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += v; }
    if ((t & 31) == 31) warp_sums[t >> 5] = incl;
    __syncthreads();
    if (t < 32) { unsigned w = warp_sums[t], wi = w; for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += v; } warp_sums[t] = wi - w; }
    __syncthreads();
    unsigned run = warp_sums[t >> 5] + incl - local;
    for (int i = 0; i < PER; i++) { const unsigned h = hist[t * PER + i]; offsets[t * PER + i] = run; run += h; hist[t * PER + i] = 0; }
}
__global__ void __launch_bounds__(256) k_sort_scatter(const unsigned* __restrict__ n_ptr, const uint32_t* __restrict__ keys, const float4* __restrict__ rays_in, const uint32_t* __restrict__ path_in,
                                                       unsigned* __restrict__ offsets, float4* __restrict__ rays_out, uint32_t* __restrict__ path_out) {
    const int n = (int)*n_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned pos = atomicAdd(offsets + __ldg(keys + i), 1u);
        rays_out[2 * pos] = __ldg(rays_in + 2 * i); rays_out[2 * pos + 1] = __ldg(rays_in + 2 * i + 1);
        path_out[pos] = __ldg(path_in + i);
    }
}

// Fused traversal launch: the shadow rays of bounce b (any hit -> cl += pending) and the extension rays of bounce b+1
// (closest hit -> hit records) share one persistent kernel, so each bounce has ONE traversal tail instead of two and small
// queues (late bounces, 8-GPU tiles) fill the machine together.  Lane-level any-hit flag; same arithmetic as MODE 0 / 1.
__global__ void __launch_bounds__(128, 8) k_intersect_fused(const __grid_constant__ DScene S, const __grid_constant__ TravTune tune,
        const float4* __restrict__ ext_rays, const unsigned* __restrict__ n_ext_ptr, const float4* __restrict__ sh_rays, const unsigned* __restrict__ n_sh_ptr,
        unsigned* work_ctr, float4* __restrict__ hit_a, uint32_t* __restrict__ hit_node, const float4* __restrict__ sh_payload, float4* __restrict__ cl) {
    const int n_ext = (int)*n_ext_ptr, n_sh = (int)*n_sh_ptr;
    VisitCounters<false> cnt;
    TravOut out = {hit_a, hit_node, sh_payload, cl, nullptr, sh_rays, n_ext, nullptr};
    trace_persistent<4, false, false>(S, ext_rays, n_ext + n_sh, work_ctr, out, tune, cnt);
}

// Fused API launch (WavefrontPathTracer's FinishIteration): the primary queue (closest hit) and the secondary queue (any hit against the
// ray's own tmax) of one iteration share one persistent kernel; 16-byte traversalResult records into two result buffers.
__global__ void __launch_bounds__(128, 8) k_intersect_fused_api(const __grid_constant__ DScene S, const __grid_constant__ TravTune tune,
        const float4* __restrict__ rays, const unsigned* __restrict__ n_ptr, const float4* __restrict__ sec_rays, const unsigned* __restrict__ n_sec_ptr,
        unsigned* work_ctr, void* __restrict__ res, void* __restrict__ sec_res) {
    const int n_ext = (int)*n_ptr, n_sec = (int)*n_sec_ptr;
    VisitCounters<false> cnt;
    TravOut out = {nullptr, nullptr, nullptr, nullptr, res, sec_rays, n_ext, sec_res};
    trace_persistent<5, false, false>(S, rays, n_ext + n_sec, work_ctr, out, tune, cnt);
}

// ---- material sort before shading (SortMode 2) ----------------------------------------------------------------------
// The BSDF dispatch of the reference is a 15-way if-chain inside one thread (Base/VirtualFuncType.h:90-111); a warp that
// holds diffuse, rough-conductor and dielectric hits executes all three bodies.  Grouping the hit queue by material class
// (bsdf type, microfacet distribution; misses last) lets each warp of k_shade run one body.  8 buckets: warp-aggregated
// histogram -> 8-entry scan (done by every block of the scatter) -> index scatter; only 4-byte indices move.
constexpr int MAT_CLASSES = 8;
CTL_DEV unsigned material_class(const DScene& S, float4 ha, uint32_t node) {
    const uint32_t tri_word = __float_as_uint(ha.w);
    if (tri_word == 0xffffffffu) return MAT_CLASSES - 1;
    const uint32_t tri = tri_word & TRI_IDX_MASK;
    const uint32_t w1 = __ldg(&S.tri_data[(size_t)tri * 2].y);
    const ctl_material* m = S.materials + ((w1 >> 16) & 0xff) + __ldg(&S.nodes[node].material_offset);
    const uint32_t bt = __ldg(&m->bsdf_type);
    return bt == CTL_BSDF_ROUGHCONDUCTOR ? 1u + (__ldg(&m->distr_type) & 1u) : (bt == CTL_BSDF_DIELECTRIC ? 3u : 0u);
}
__global__ void __launch_bounds__(256) k_matsort_classify(const __grid_constant__ DScene S, const unsigned* __restrict__ n_ptr, const float4* __restrict__ hit_a, const uint32_t* __restrict__ hit_node,
                                                           unsigned char* __restrict__ cls, unsigned* __restrict__ hist /* MAT_CLASSES, zeroed */) {
    const int n = (int)*n_ptr;
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        unsigned c = MAT_CLASSES; // out of range lanes: no class
        if (i < n) { c = material_class(S, __ldg(hit_a + i), __ldg(hit_node + i)); cls[i] = (unsigned char)c; }
        for (unsigned b = 0; b < MAT_CLASSES; b++) {
            const unsigned m = __ballot_sync(0xffffffffu, c == b);
            if (m && lane_id() == 0) atomicAdd(hist + b, (unsigned)__popc(m));
        }
    }
}
__global__ void __launch_bounds__(256) k_matsort_scatter(const unsigned* __restrict__ n_ptr, const unsigned char* __restrict__ cls, const unsigned* __restrict__ hist, unsigned* __restrict__ cursor /* MAT_CLASSES, zeroed */,
                                                          uint32_t* __restrict__ order) {
    const int n = (int)*n_ptr;
    unsigned start[MAT_CLASSES]; unsigned run = 0;
    for (int b = 0; b < MAT_CLASSES; b++) { start[b] = run; run += hist[b]; }
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const unsigned c = i < n ? cls[i] : MAT_CLASSES;
        for (unsigned b = 0; b < MAT_CLASSES; b++) {
            const unsigned m = __ballot_sync(0xffffffffu, c == b);
            if (!m) continue;
            unsigned base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane_id() == leader) base = atomicAdd(cursor + b, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (c == b) order[start[b] + base + __popc(m & ((1u << lane_id()) - 1u))] = (uint32_t)i;
        }
    }
}

// Class-grouped order of a bounce's hit records for the per-class shade launches (ShadeMode 1).  The staged traversal kernel has already put the
// material class of every hit into the top bits of its triangle word and counted the classes (TravOut::cls_hist), so this pass only reads the
// 16-byte hit records and writes 4-byte indices; misses (class 7) need no shading and get no slot.
__global__ void __launch_bounds__(256) k_class_scatter(const unsigned* __restrict__ n_ptr, const float4* __restrict__ hit_a, const unsigned* __restrict__ hist, unsigned* __restrict__ cursor /* MAT_CLASSES, zeroed */,
                                                        uint32_t* __restrict__ order) {
    const int n = (int)*n_ptr;
    unsigned start[MAT_CLASSES]; unsigned run = 0;
    for (int b = 0; b < MAT_CLASSES; b++) { start[b] = run; run += hist[b]; }
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const unsigned c = i < n ? (__float_as_uint(__ldg(hit_a + i).w) >> TRI_CLS_SHIFT) : MAT_CLASSES;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c < MAT_CLASSES - 1) {
            const int leader = __ffs(peers) - 1;
            unsigned base = 0;
            if ((int)lane_id() == leader) base = atomicAdd(cursor + c, (unsigned)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            order[start[c] + base + __popc(peers & ((1u << lane_id()) - 1u))] = (uint32_t)i;
        }
    }
}

// ---- shade: one path vertex (PathTracer.cu:58-96) ----------------------------------
struct ShadeParams { int max_path_length, rr_start, direct, stop_zero; };   // (k_shade<CLS, REGU>: REGU = the reference's KEY_Regularization)

#ifndef CTL_SHADE_MIN_BLOCKS
#define CTL_SHADE_MIN_BLOCKS 8 // 64 registers: 8 resident blocks per SM; measured -18% (diffuse) / -27% (microfacet) shade time vs 116 registers
#endif
// CLS = -1: any material (the reference's run-time BSDF dispatch, Base/VirtualFuncType.h:90-111).  CLS = 0..3: every work item of the launch is known to
// hit that material class (0 diffuse, 1 rough conductor / Beckmann, 2 rough conductor / GGX, 3 dielectric), so one BSDF body is compiled in and a warp
// never waits for another class's code (the Beckmann visible-normal Newton loop in particular).  With `seg_hist` the launch covers only its class's
// segment of Q.order (k_class_scatter); without it the whole queue (single-class scenes).  Same per-path arithmetic either way.
CTL_DEV constexpr uint32_t cls_bsdf_type(int cls) { return cls == 0 ? CTL_BSDF_DIFFUSE : (cls == 3 ? CTL_BSDF_DIELECTRIC : CTL_BSDF_ROUGHCONDUCTOR); }
#ifndef CTL_SHADE_MICRO_MIN_BLOCKS
#define CTL_SHADE_MICRO_MIN_BLOCKS CTL_SHADE_MIN_BLOCKS // resident blocks per SM of the rough-conductor launches (their bodies spill ~150 B at 64 registers)
#endif
// REGU = true: one vertex of PathTraceRegularization<DIRECT> (Integrators/PathTracer.cu:115-170) instead of PathTrace<DIRECT>: emitted radiance only at
// depth 1 / after a specular bounce / without direct lighting and without MIS weight; every non-delta vertex samples ALL lights (UniformSampleAllLights ->
// EstimateDirect with light pdf 1: one shadow-queue entry per light, appended by the lane itself), a delta vertex draws sampleEmitterPosition's two numbers
// and adds nothing (this path has DiffuseLights only, cu:138); Russian roulette also after specular bounces; a path that survives its last vertex still
// emits its next ray (the reference traces before it tests the depth, cu:125) and the next shade launch leaves it alone.
template <int CLS, bool REGU = false>
__global__ void __launch_bounds__(128, (CLS == 1 || CLS == 2) ? CTL_SHADE_MICRO_MIN_BLOCKS : CTL_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ DScene S, const __grid_constant__ ShadeParams P, PathState st, Queues Q,
                                                const unsigned* __restrict__ n_in, unsigned* n_out, unsigned* n_shadow, const unsigned* __restrict__ seg_hist) {
    int n = (int)*n_in, seg_start = 0;
    if (CLS >= 0 && seg_hist) { for (int b = 0; b < CLS; b++) seg_start += (int)seg_hist[b]; n = (int)seg_hist[CLS]; }
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool alive = false, shadow = false;
        uint32_t p = 0, p_lag = 0;   // path id; its lag bits (frames with DeferStragglers: the path runs behind the wavefront's bounce, device/traverse_handover.cuh)
        V3 no = mk(0, 0, 0), nd = mk(0, 0, 1), sd = mk(0, 0, 1);
        float sh_tmax = 0.0f;
        Spec pending = sp(0.0f);
        if (i < n) {
            const int i_unsorted = i;
            const int i = Q.order ? (int)__ldg(Q.order + seg_start + i_unsorted) : i_unsorted; // material-sorted shading: same work items, grouped
            p = Q.path_in[i]; p_lag = p & ~PATH_ID_MASK; p &= PATH_ID_MASK;
            const float4 ha = Q.hit_a[i];
            const uint32_t tri_word = __float_as_uint(ha.w);
            const uint32_t tri = tri_word & TRI_IDX_MASK;   // the staged traversal kernel leaves the material class in the top bits
            if (tri_word != 0xffffffffu && tri_word != TRI_DEFERRED && !(REGU && (int)(__float_as_uint(st.nor[p].w) & 0xff) >= P.max_path_length)) {   // (REGU: the ray past the last vertex is traced, not shaded)
                const uint32_t node = Q.hit_node[i];
                const float4 r0 = Q.rays_in[2 * i], r1 = Q.rays_in[2 * i + 1];
                const V3 ro = mk(r0.x, r0.y, r0.z), rd = mk(r1.x, r1.y, r1.z);
                const float4 cf4 = st.cf[p], nor4 = st.nor[p];
                float4 cl4 = st.cl[p];
                Spec cf = mk_sp(cf4.x, cf4.y, cf4.z), cl = mk_sp(cl4.x, cl4.y, cl4.z);
                float brdf_pdf = cf4.w;
                V3 last_nor = mk(nor4.x, nor4.y, nor4.z);
                const uint32_t ctl = __float_as_uint(nor4.w);
                const int depth = (int)(ctl & 0xff) + 1; // depth++ at loop entry
                bool specularBounce = (ctl & 256u) != 0;
                const float4 px4 = st.px[p];
                Sampler rnd; rnd.idx = __float_as_uint(px4.z); rnd.tab = __float_as_uint(px4.w) - 1u; rnd.i1 = (ctl >> 9) & 0x3ff; rnd.i2 = ctl >> 19;
                // getBsdfSample (Kernel/TraceResult.cu:16-43)
                DG dg; uint32_t mat_local;
                dg.P = ro + rd * ha.x;
                fill_dg(S, ha.y, ha.z, tri, node, dg, mat_local);
                const ctl_node* N = S.nodes + node;
                ctl_material mat = S.materials[mat_local + __ldg(&N->material_offset)];
                if (CLS >= 0) { mat.bsdf_type = cls_bsdf_type(CLS); if (CLS == 1) mat.distr_type = CTL_DISTR_BECKMANN; if (CLS == 2) mat.distr_type = CTL_DISTR_GGX; } // known at compile time: the other bodies fold away
                BRec bRec; bRec.eta = 1.0f; bRec.sampledType = 0; bRec.typeMask = E_ALL; bRec.wo = mk(0, 0, 1);
                if (st.wo && depth > 1) { const float4 w4 = st.wo[p]; bRec.wo = mk(w4.x, w4.y, w4.z); }
                bRec.wi = to_local(dg.sys, -rd);
                if ((mat.flags & CTL_MAT_TWO_SIDED) && bRec.wi.z < 0) { dg.n = -dg.n; dg.sys.n = -dg.sys.n; bRec.wi.z *= -1.0f; }
                // emitter hit with MIS (PathTracer.cu:64-77)
                if (mat.node_light_index != 0xffffffffu && (!REGU || !P.direct || depth == 1 || specularBounce)) {
                    const unsigned li = mat.node_light_index == 0 ? __ldg(&N->lights[0]) : __ldg(&N->lights[1]);
                    const ctl_light L = S.lights[li];
                    float misWeight = 1.0f;
                    if (!REGU && !(!P.direct || depth == 1 || specularBounce)) {
                        DRec dRec; dRec.ref = ro; dRec.refN = last_nor; dRec.p = dg.P; dRec.n = dg.n; dRec.d = rd; dRec.dist = ha.x;
                        const float direct_pdf = light_pdf_direct(L, dRec) * pdf_emitter(S, li);
                        misWeight = power_heuristic(brdf_pdf, direct_pdf);
                    }
                    const Spec Le = dot(dg.sys.n, -rd) <= 0 ? sp(0.0f) : sp3(L.radiance);
                    cl = cl + (cf * misWeight) * Le;
                }
                const float2 bs = rnd.f2(S);
                const Spec f = bsdf_sample(mat, bRec, brdf_pdf, bs.x, bs.y);
                last_nor = dg.sys.n;
                if (REGU && P.direct) {
                    if (bsdf_combined_type(mat.bsdf_type) & E_DELTA) rnd.f2(S);   // sampleEmitterPosition's sample (cu:136-137)
                    else for (unsigned li = 0; li < S.num_lights; li++) {   // UniformSampleAllLights, one sample per light (TraceAlgorithms.cu:75-90)
                        const ctl_light L = S.lights[S.light_indices[li]];
                        DRec dRec; dRec.ref = dg.P; dRec.refN = dg.sys.n; dRec.p = dg.P; dRec.n = dg.sys.n; dRec.pdf = 0;
                        const float2 es = rnd.f2(S);
                        const Spec value = light_sample_direct(S, L, dRec, es.x, es.y);
                        if (!is_zero(value)) {
                            BRec b2 = bRec;
                            b2.wo = to_local(dg.sys, dRec.d);
                            b2.typeMask = E_ALL & ~E_DELTA;
                            const Spec bsdfVal = bsdf_f(mat, b2);
                            if (!is_zero(bsdfVal)) {
                                const float weight = power_heuristic(dRec.pdf * 1.0f, bsdf_pdf(mat, b2));
                                const Spec contrib = cf * ((value * bsdfVal) * weight);
                                const unsigned ks = atomicAdd(n_shadow, 1u);
                                Q.sh_rays[2 * ks] = make_float4(dg.P.x, dg.P.y, dg.P.z, S.ray_eps);
                                Q.sh_rays[2 * ks + 1] = make_float4(dRec.d.x, dRec.d.y, dRec.d.z, dRec.dist - S.ray_eps);
                                Q.sh_payload[ks] = make_float4(contrib.r, contrib.g, contrib.b, __uint_as_float(p));
                            }
                        }
                    }
                }
                // next-event estimation (TraceAlgorithms.cu:44-101)
                if (!REGU && P.direct && (bsdf_combined_type(mat.bsdf_type) & E_SMOOTH) && S.num_lights) {
                    const float2 ls = rnd.f2(S);
                    unsigned first = 0, count = S.num_lights; // STL_upper_bound, Base/STL.h:21-38
                    while (count > 0) { const unsigned c2 = count / 2, mid = first + c2; if (!(ls.x < S.light_cdf[mid])) { first = mid + 1; count -= c2 + 1; } else count = c2; }
                    unsigned idx = first; if (idx >= S.num_lights) idx = S.num_lights - 1;
                    const float fU = S.light_cdf[idx], fL = idx > 0 ? S.light_cdf[idx - 1] : 0.0f;
                    const float emPdf = fU - fL;
                    const ctl_light L = S.lights[S.light_indices[idx]];
                    DRec dRec; dRec.ref = dg.P; dRec.refN = dg.sys.n; dRec.p = dg.P; dRec.n = dg.sys.n; dRec.pdf = 0;
                    const float2 es = rnd.f2(S);
                    const Spec value = light_sample_direct(S, L, dRec, es.x, es.y);
                    if (!is_zero(value)) {
                        BRec b2 = bRec;
                        b2.wo = to_local(dg.sys, dRec.d);
                        b2.typeMask = E_ALL & ~E_DELTA;
                        const Spec bsdfVal = bsdf_f(mat, b2);
                        if (!is_zero(bsdfVal)) {
                            const float bsdfPdf = bsdf_pdf(mat, b2);
                            const float directPdf = dRec.pdf * emPdf;
                            const float weight = power_heuristic(directPdf, bsdfPdf);
                            const Spec retVal = (value * bsdfVal) * weight;
                            pending = cf * (retVal / emPdf);
                            shadow = true; sd = dRec.d; sh_tmax = dRec.dist - S.ray_eps;
                        }
                    }
                }
                specularBounce = (bRec.sampledType & E_DELTA) != 0;
                cf = cf * f;
                nd = to_world(dg.sys, bRec.wo); no = dg.P;
                alive = !P.stop_zero || !is_zero(cf);   // stop_zero: a path of zero throughput ends here (no image effect); off: it lives until Russian roulette, as in the reference
                if (alive && depth > P.rr_start && (REGU || !specularBounce)) {
                    const float q = smax(cf);
                    if (rnd.f1(S) >= q) alive = false;
                    else cf = cf / q;
                }
                if (!REGU && depth >= P.max_path_length) alive = false;
                cl4.x = cl.r; cl4.y = cl.g; cl4.z = cl.b;
                st.cl[p] = cl4;
                if (alive) {
                    st.cf[p] = make_float4(cf.r, cf.g, cf.b, brdf_pdf);
                    st.nor[p] = make_float4(last_nor.x, last_nor.y, last_nor.z, __uint_as_float(pack_ctl(depth, specularBounce, rnd.i1, rnd.i2)));
                    if (st.wo) st.wo[p] = make_float4(bRec.wo.x, bRec.wo.y, bRec.wo.z, 0.0f);
                }
            }
        }
        const int jq = warp_append(alive, n_out);
        if (alive) {
            Q.rays_out[2 * jq] = make_float4(no.x, no.y, no.z, S.ray_eps);
            Q.rays_out[2 * jq + 1] = make_float4(nd.x, nd.y, nd.z, FLT_MAX);
            Q.path_out[jq] = p | p_lag;
            if (Q.keys_out) { const uint32_t key = ray_sort_key(S, no, nd); Q.keys_out[jq] = key; atomicAdd(Q.hist + key, 1u); }
        }
        const int ks = warp_append(shadow, n_shadow);
        if (shadow) {
            Q.sh_rays[2 * ks] = make_float4(no.x, no.y, no.z, S.ray_eps);
            Q.sh_rays[2 * ks + 1] = make_float4(sd.x, sd.y, sd.z, sh_tmax);
            Q.sh_payload[ks] = make_float4(pending.r, pending.g, pending.b, __uint_as_float(p));
        }
    }
}

// ---- finish: img.AddSample(pX.x, pX.y, imp * L) (PathTracer.cu:191-192, Image.cu:22-44) ----
__global__ void __launch_bounds__(256) k_finish(int n_slots /* all passes */, PathState st, float* accum, int img_w, int img_h) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
        const float4 px = st.px[s];
        if (__float_as_uint(px.w) == 0u) continue;
        const float4 c = st.cl[s];
        const float r = fmaxf(0.0f, c.x), g = fmaxf(0.0f, c.y), b = fmaxf(0.0f, c.z);
        const int x = (int)floorf(px.x), y = (int)floorf(px.y);
        if (x < 0 || x >= img_w || y < 0 || y >= img_h || !isfinite(r) || !isfinite(g) || !isfinite(b)) continue;
        float* dst = accum + ((size_t)y * img_w + x) * 7;
        atomicAdd(dst + 0, r); atomicAdd(dst + 1, g); atomicAdd(dst + 2, b); atomicAdd(dst + 6, 1.0f);
    }
}

// rays of the pass = sum of extension + shadow queue sizes (every traceRay call counts, TraceHelper.cu:176)
// (deferred: 2 * n_bounces + 2 counters of queue entries that were deferred rays re-queued by the traversal launches -- counted once, where they were first queued)
__global__ void k_tally(const unsigned* q_count, const unsigned* sh_count, int n_bounces, unsigned long long* rays_last, unsigned long long* rays_total, int add_to_last = 0, const unsigned* deferred = nullptr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long s = 0;
        for (int b = 0; b < n_bounces; b++) s += (unsigned long long)q_count[b] + (unsigned long long)sh_count[b];
        if (deferred) for (int k = 0; k < 2 * n_bounces + 2; k++) s -= (unsigned long long)deferred[k];
        *rays_last = add_to_last ? *rays_last + s : s; atomicAdd(rays_total, s);   // two wavefronts of a frame may tally concurrently (OverlapWavefronts)
    }
}

} // namespace ctld
