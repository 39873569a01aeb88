// wavefront_pt.cuh -- drop-in for the reference's own wavefront integrator: WavefrontPathTracer over DoubleRayBuffer
// (Integrators/PseudoRealtime/WavefrontPathTracer.cu:17-191, Kernel/DoubleRayBuffer.h), SURVEY 8 f1 -- the second consumer of the
// intersect kernel (__internal__IntersectBuffers == k_intersect<2, ...>).
//
// Parity target: the reference's algorithm in the SERIAL order of its queue atomics (fetch slot i = 0, 1, 2, ...; the k-th insertion lands
// in slot k).  The reference's GPU order depends on the hardware scheduler and a path's random numbers are keyed by its queue slot
// (cu:59-60), so its output is not reproducible run to run; the serial order is the deterministic member of its possible outputs, and is what
// the reference's own kernel text produces on one host thread (oracle/_ref).  Here that order is produced IN PARALLEL: the iterate kernel
// compacts with a single-pass chained scan (decoupled look-back over 256-slot tiles, tile ids handed out by an atomic ticket so a tile only
// waits on tiles that already run), so slot ranks -- and with them every random number, the secondary-ray indices and the image -- equal the
// serial schedule's.  The queue is compacted in place like the reference's (one payload / ray / result buffer + two secondary buffers): a tile
// writes only into slots of tiles <= itself, all of which have been read before its look-back completes.
//
// Quirks of the reference kept on purpose (oracle/oracle.cpp orc_render_wavefront): sampler keyed by queue slot and re-skipped to
// dimension passesDone + 2 at every bounce; Russian roulette before the BSDF sample from pathDepth >= RRStartDepth; one 2-D sample re-used for
// light selection and position; 16-bit barycentrics (traversalResult);
// 16-bit spherical previous normal; half-precision un-jittered splat position; a failed BSDF sample still launches a ray (zero direction).
#pragma once
#include "wavefront.cuh"

namespace ctld {

// WavefrontPTRayData (WavefrontPathTracer.h:11-22) as SoA + the DoubleRayBuffer arrays (DoubleRayBuffer.h:16-36)
struct WptBuf {
    float4* thr;     // throughput rgb, bsdf_pdf
    float4* lxy;     // L rgb, (half x | half y << 16) bits
    float4* df;      // directF rgb, dDist
    uint2* misc;     // dIdx, prev_normal | specular_bounce << 16
    float4* ray;     // m_payload_ray_buffer (traversalRay = 2 x float4)
    uint4* res;      // m_payload_res_buffer (traversalResult)
    float4* sec_out; // m_secondary_buf2.m_ray_buffer: filled by this iteration
    const uint4* sec_res; // m_secondary_buf1.m_res_buffer: results of the previous iteration's secondary rays
};

struct WptParams { int pathDepth, iterationIdx, maxPathDepth, rrStart; };

constexpr int WPT_TILE = 256;
#ifndef WPT_MIN_BLOCKS
#define WPT_MIN_BLOCKS 4 // 256-slot tiles x 4 blocks per SM at 64 registers: measured +2 % (occupancy) and +1.3 % (half as many tiles to look back over) over
                         // 128 x 6 at 80 registers on C2, profiles/r01v_wpt_occupancy_ab.log
#endif

CTL_DEV unsigned short enc_normal_dev(V3 v) { // NormalizedFloat3ToUchar2_Spherical, Math/Compression.h:12-18
    const float theta = acosf(v.z) * (255.0f / PI_F);
    float phi = atan2f(v.y, v.x) * (255.0f / (2.0f * PI_F));
    phi = phi < 0 ? (phi + 255) : phi;
    return (unsigned short)(((unsigned short)theta << 8) | (unsigned short)phi);
}

// ---- pathCreateKernelWPT (cu:17-49): slot = pixel index, one sample per pixel (uniform block sampler)
__global__ void __launch_bounds__(256) k_wpt_create(const __grid_constant__ DScene S, WptBuf B, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = i % S.img_w, y = i / S.img_w;
        Sampler rng; rng.idx = (unsigned)i; rng.i1 = 0; rng.i2 = 0; rng.tab = 0;
        const float2 j = rng.f2(S);
        rng.f2(S); // aperture sample (pinhole: unused)
        V3 o, d; camera_ray(S, (float)x + j.x, (float)y + j.y, o, d);
        const unsigned hx = __half_as_ushort(__float2half_rn((float)x)), hy = __half_as_ushort(__float2half_rn((float)y));
        B.thr[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        B.lxy[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(hx | (hy << 16)));
        B.df[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        B.misc[i] = make_uint2(0xffffffffu, 1u << 16); // dIdx = UINT_MAX, specular_bounce = true
        B.ray[2 * i] = make_float4(o.x, o.y, o.z, S.ray_eps);
        B.ray[2 * i + 1] = make_float4(d.x, d.y, d.z, FLT_MAX);
    }
}

// ---- pathIterateKernel<NEXT_EVENT_EST> (cu:51-164) with order-preserving compaction
// desc: one 64-bit descriptor per tile: [63:62] status (0 = not ready, 1 = tile aggregate, 2 = inclusive prefix), [61:31] secondary count,
// [30:0] payload count; desc[n_tiles_max] is the tile ticket.  All zero at launch.
template <bool NEE>
__global__ void __launch_bounds__(WPT_TILE, WPT_MIN_BLOCKS) k_wpt_iterate(const __grid_constant__ DScene S, const __grid_constant__ WptParams P, WptBuf B, const unsigned* __restrict__ n_in,
                                                             unsigned* __restrict__ n_pay_out, unsigned* __restrict__ n_sec_out, unsigned long long* desc, int n_tiles_max, float* accum) {
    __shared__ unsigned s_tile;
    __shared__ unsigned s_warp_pay[WPT_TILE / 32], s_warp_sec[WPT_TILE / 32];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = (unsigned)atomicAdd(desc + n_tiles_max, 1ull);
    __syncthreads();
    const int tile = (int)s_tile;
    const int n = (int)*n_in;
    const int base = tile * WPT_TILE;
    if (base >= n) return;
    const int i = base + tid;

    bool alive = false, shadow = false, terminated = false;
    float4 thr4 = make_float4(0, 0, 0, 0), lxy4 = make_float4(0, 0, 0, 0), df4 = make_float4(0, 0, 0, 0);
    uint2 misc = make_uint2(0xffffffffu, 0u);
    V3 no = mk(0, 0, 0), nd = mk(0, 0, 0), sd = mk(0, 0, 0);
    if (i < n) {
        thr4 = B.thr[i]; lxy4 = B.lxy[i]; df4 = B.df[i]; misc = B.misc[i];
        const float4 r0 = B.ray[2 * i], r1 = B.ray[2 * i + 1];
        const uint4 res = B.res[i];
        const V3 ro = mk(r0.x, r0.y, r0.z), rd = mk(r1.x, r1.y, r1.z);
        Spec thr = mk_sp(thr4.x, thr4.y, thr4.z), L = mk_sp(lxy4.x, lxy4.y, lxy4.z), directF = mk_sp(df4.x, df4.y, df4.z);
        float pay_pdf = thr4.w, dDist = df4.w; // payload.bsdf_pdf
        unsigned dIdx = misc.x, prev_normal = misc.y & 0xffffu;
        bool specular = (misc.y >> 16) != 0;
        Sampler rng; rng.idx = (unsigned)i; rng.tab = 0; rng.i1 = rng.i2 = (unsigned)P.iterationIdx + 2u; // cu:59-60
        if (NEE && P.pathDepth > 0 && dIdx != 0xffffffffu) { // cu:62-73
            // reference: closest hit of the secondary ray, then `dist >= dDist * (1 - eps)`.  Here the ray carries tmax = dDist * (1 - eps) and is
            // traced as an any-hit query: "no hit below tmax" is the same predicate (same strict t < tmax test), at a fraction of the traversal
            if (B.sec_res[dIdx].z == 0xffffffffu) L = L + directF;
            dIdx = 0xffffffffu; directF = sp(0.0f);
        }
        terminated = (P.pathDepth + 1 == P.maxPathDepth);
        const uint32_t tri = res.z;
        if (tri != 0xffffffffu) {
            const uint32_t node = res.y;
            const float dist = __uint_as_float(res.x);
            const float bu = (float)(unsigned short)(res.w & 0xffffu) / 65535.0f, bv = (float)(unsigned short)(res.w >> 16) / 65535.0f; // toResult, TraceHelper.cu:44-51
            DG dg; uint32_t mat_local;
            dg.P = ro + rd * dist;
            fill_dg(S, bu, bv, tri, node, dg, mat_local);
            const ctl_node* N = S.nodes + node;
            const ctl_material mat = S.materials[mat_local + __ldg(&N->material_offset)];
            BRec bRec; bRec.eta = 1.0f; bRec.sampledType = 0; bRec.typeMask = E_ALL; bRec.wo = mk(0, 0, 0); // wo: unset in the reference, defined as zero
            bRec.wi = to_local(dg.sys, -rd);
            if ((mat.flags & CTL_MAT_TWO_SIDED) && bRec.wi.z < 0) { dg.n = -dg.n; dg.sys.n = -dg.sys.n; bRec.wi.z *= -1.0f; }
            if (mat.node_light_index != 0xffffffffu) { // emission with MIS (cu:84-99)
                const unsigned li = mat.node_light_index == 0 ? __ldg(&N->lights[0]) : __ldg(&N->lights[1]);
                const ctl_light Lt = S.lights[li];
                float misWeight = 1.0f;
                if (!(!NEE || P.pathDepth == 0 || specular)) {
                    DRec dRec; dRec.ref = ro; dRec.refN = dec_normal(S, prev_normal); dRec.p = dg.P; dRec.n = dg.n; dRec.d = rd; dRec.dist = dist;
                    const float direct_pdf = light_pdf_direct(Lt, dRec) * pdf_emitter(S, li);
                    misWeight = power_heuristic(pay_pdf, direct_pdf);
                }
                const Spec Le = dot(dg.sys.n, -rd) <= 0 ? sp(0.0f) : sp3(Lt.radiance);
                L = L + (Le * misWeight) * thr;
            }
            bool surviveRR = true;
            if (P.pathDepth >= P.rrStart) { // cu:102-109
                const float q = smax(thr);
                if (rng.f1(S) < q) thr = thr / q;
                else surviveRR = false;
            }
            if (P.pathDepth + 1 != P.maxPathDepth && surviveRR) {
                const float2 bs = rng.f2(S);
                const Spec f = bsdf_sample(mat, bRec, pay_pdf, bs.x, bs.y);
                specular = (bRec.sampledType & E_DELTA) != 0;
                nd = to_world(dg.sys, bRec.wo); no = dg.P;
                dIdx = 0xffffffffu;
                if (NEE && (bsdf_combined_type(mat.bsdf_type) & E_SMOOTH)) { // cu:118-135
                    DRec dRec; dRec.ref = dg.P; dRec.refN = dg.sys.n; dRec.p = dg.P; dRec.n = dg.sys.n; dRec.pdf = 0; dRec.d = mk(0, 0, 1); dRec.dist = 0;
                    float2 ls = rng.f2(S);
                    Spec value = sp(0.0f);
                    if (S.num_lights) { // sampleEmitterDirect with sample re-use (KernelDynamicScene.cu:25-40, 98-117)
                        unsigned first = 0, count = S.num_lights;
                        while (count > 0) { const unsigned c2 = count / 2, mid = first + c2; if (!(ls.x < S.light_cdf[mid])) { first = mid + 1; count -= c2 + 1; } else count = c2; }
                        unsigned idx = first; if (idx >= S.num_lights) idx = S.num_lights - 1;
                        const float fU = S.light_cdf[idx], fL = idx > 0 ? S.light_cdf[idx - 1] : 0.0f;
                        ls.x = (ls.x - fL) / (fU - fL);
                        const float emPdf = fU - fL;
                        value = light_sample_direct(S, S.lights[S.light_indices[idx]], dRec, ls.x, ls.y);
                        if (dRec.pdf != 0) { dRec.pdf *= emPdf; value = value / emPdf; } else value = sp(0.0f);
                    }
                    if (!is_zero(value)) {
                        bRec.typeMask = E_ALL & ~E_DELTA;
                        bRec.wo = to_local(dg.sys, dRec.d);
                        const Spec bsdfVal = bsdf_f(mat, bRec);
                        const float bsdfPdf = bsdf_pdf(mat, bRec);
                        const float weight = power_heuristic(dRec.pdf, bsdfPdf);
                        directF = ((thr * value) * bsdfVal) * weight;
                        dDist = dRec.dist;
                        shadow = true; sd = dRec.d;
                    }
                }
                prev_normal = enc_normal_dev(dg.sys.n);
                thr = thr * f;
                alive = true;
            } else terminated = true;
        } else {
            terminated = true;
            L = L + (thr * sp(0.0f)) * 1.0f; // no environment map (cu:143-156)
        }
        if (terminated) { // I.AddSample(x, y, L) (cu:159-162, Engine/Image.cu:22-44)
            const float r = fmaxf(0.0f, L.r), g = fmaxf(0.0f, L.g), b = fmaxf(0.0f, L.b);
            const unsigned xy = __float_as_uint(lxy4.w);
            const int x = (int)floorf(h2f(xy)), y = (int)floorf(h2f(xy >> 16));
            if (!(x < 0 || x >= S.img_w || y < 0 || y >= S.img_h || !isfinite(r) || !isfinite(g) || !isfinite(b))) {
                float* dst = accum + ((size_t)y * S.img_w + x) * 7;
                atomicAdd(dst + 0, r); atomicAdd(dst + 1, g); atomicAdd(dst + 2, b); atomicAdd(dst + 6, 1.0f);
            }
        }
        thr4 = make_float4(thr.r, thr.g, thr.b, pay_pdf);
        lxy4 = make_float4(L.r, L.g, L.b, lxy4.w);
        df4 = make_float4(directF.r, directF.g, directF.b, dDist);
        misc = make_uint2(dIdx, prev_normal | (specular ? 1u << 16 : 0u));
    }

    // ---- ranks inside the tile
    const unsigned lane = tid & 31u, warp = tid >> 5;
    const unsigned m_pay = __ballot_sync(0xffffffffu, alive), m_sec = __ballot_sync(0xffffffffu, shadow);
    if (lane == 0) { s_warp_pay[warp] = __popc(m_pay); s_warp_sec[warp] = __popc(m_sec); }
    __syncthreads(); // also: every thread of the tile has read its slot before the aggregate is published
    unsigned pre_pay = 0, pre_sec = 0, tot_pay = 0, tot_sec = 0;
#pragma unroll
    for (int w = 0; w < WPT_TILE / 32; w++) {
        if (w < (int)warp) { pre_pay += s_warp_pay[w]; pre_sec += s_warp_sec[w]; }
        tot_pay += s_warp_pay[w]; tot_sec += s_warp_sec[w];
    }
    const unsigned long long aggregate = (unsigned long long)tot_pay | ((unsigned long long)tot_sec << 31);

    // ---- chained scan across tiles (decoupled look-back, warp 0)
    if (warp == 0) {
        volatile unsigned long long* vdesc = desc;
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) { __threadfence(); vdesc[0] = (2ull << 62) | aggregate; }
        } else {
            if (lane == 0) { __threadfence(); vdesc[tile] = (1ull << 62) | aggregate; }
            int look = tile - 1;
            for (;;) {
                const int idx = look - (int)lane;
                unsigned long long d;
                do { d = idx >= 0 ? vdesc[idx] : (2ull << 62); } while (__any_sync(0xffffffffu, (d >> 62) == 0ull));
                const unsigned incl = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
                const int first = incl ? __ffs(incl) - 1 : 31;
                unsigned long long v = ((int)lane <= first) ? (d & 0x3fffffffffffffffull) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                excl += v;
                if (incl) break;
                look -= 32;
            }
            __threadfence(); // acquire side of the hand-over: the reads of the predecessor tiles (ordered before their descriptors by their barrier + fence)
                             // happen before this tile's writes into their slots
            if (lane == 0) vdesc[tile] = (2ull << 62) | (excl + aggregate);
        }
        if (lane == 0) {
            s_excl = excl;
            if (base + WPT_TILE >= n) { // last tile: queue sizes of the next iteration
                const unsigned long long total = excl + aggregate;
                *n_pay_out = (unsigned)(total & 0x7fffffffull); *n_sec_out = (unsigned)(total >> 31);
            }
        }
    }
    __syncthreads();
    const unsigned long long excl = s_excl;
    const unsigned excl_pay = (unsigned)(excl & 0x7fffffffull), excl_sec = (unsigned)(excl >> 31);
    const unsigned lt = (1u << lane) - 1u;
    if (shadow) { // insertSecondaryRay (DoubleRayBuffer.h:166-177)
        const unsigned k = excl_sec + pre_sec + __popc(m_sec & lt);
        B.sec_out[2 * k] = make_float4(no.x, no.y, no.z, S.ray_eps);
        B.sec_out[2 * k + 1] = make_float4(sd.x, sd.y, sd.z, df4.w * (1 - S.ray_eps)); // tmax = dDist * (1 - eps), see the occlusion test above
        misc.x = k;
    }
    if (alive) { // insertPayloadElement (DoubleRayBuffer.h:139-152)
        const unsigned j = excl_pay + pre_pay + __popc(m_pay & lt);
        B.thr[j] = thr4; B.lxy[j] = lxy4; B.df[j] = df4; B.misc[j] = misc;
        B.ray[2 * j] = make_float4(no.x, no.y, no.z, S.ray_eps);
        B.ray[2 * j + 1] = make_float4(nd.x, nd.y, nd.z, FLT_MAX);
    }
}

} // namespace ctld
