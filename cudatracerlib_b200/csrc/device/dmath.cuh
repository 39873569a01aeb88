// dmath.cuh -- device vector / spectrum helpers.
//
// The translation unit is compiled with -fmad=false: products and sums are rounded
// separately exactly as written (reference host order, Math/Vector.h), and the only fused
// multiply-adds are the explicit fmaf() calls in traverse.cuh.  That makes traversal results
// bit-identical to the CPU oracle and leaves libm-vs-libdevice transcendentals as the only
// source of GPU/CPU differences in shading.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define CTL_DEV __device__ __forceinline__

namespace ctld {

constexpr float PI_F = 3.14159265358979f;   // Math/MathFunc.h:12
constexpr float INV_PI_F = 1.0f / PI_F;
constexpr float DELTA_EPS = 1e-3f;
constexpr unsigned E_DIFFUSE_REFL = 0x2, E_GLOSSY_REFL = 0x8, E_DELTA_REFL = 0x20, E_DELTA_TRANS = 0x40; // SceneTypes/Samples.h:32-71
constexpr unsigned E_SMOOTH = 0x2 | 0x4 | 0x8 | 0x10, E_DELTA = 0x1 | 0x20 | 0x40, E_ALL = E_SMOOTH | E_DELTA | 0x80 | 0x100;

struct V3 { float x, y, z; };
CTL_DEV V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CTL_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
CTL_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
CTL_DEV V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
CTL_DEV V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
CTL_DEV V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
CTL_DEV float dot(V3 a, V3 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
CTL_DEV V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CTL_DEV float rcpf(float a) { return a != 0.0f ? 1.0f / a : 0.0f; } // MathFunc.h:399
CTL_DEV float length(V3 a) { return sqrtf(dot(a, a)); }
CTL_DEV V3 normalize(V3 a) { return a * rcpf(length(a)); }

struct Spec { float r, g, b; };
CTL_DEV Spec sp(float v) { Spec s; s.r = v; s.g = v; s.b = v; return s; }
CTL_DEV Spec sp3(const float* p) { Spec s; s.r = p[0]; s.g = p[1]; s.b = p[2]; return s; }
CTL_DEV Spec mk_sp(float r, float g, float b) { Spec s; s.r = r; s.g = g; s.b = b; return s; }
CTL_DEV Spec operator*(Spec a, Spec b) { return mk_sp(a.r * b.r, a.g * b.g, a.b * b.b); }
CTL_DEV Spec operator*(Spec a, float f) { return mk_sp(a.r * f, a.g * f, a.b * f); }
CTL_DEV Spec operator/(Spec a, float f) { const float recip = 1.0f / f; return mk_sp(a.r * recip, a.g * recip, a.b * recip); } // Math/Spectrum.h:122-128, 150-155: reciprocal multiply
CTL_DEV Spec operator/(Spec a, Spec b) { return mk_sp(a.r / b.r, a.g / b.g, a.b / b.b); }
CTL_DEV Spec operator+(Spec a, Spec b) { return mk_sp(a.r + b.r, a.g + b.g, a.b + b.b); }
CTL_DEV Spec operator-(Spec a, Spec b) { return mk_sp(a.r - b.r, a.g - b.g, a.b - b.b); }
CTL_DEV bool is_zero(Spec a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }
CTL_DEV float smax(Spec a) { float m = a.r; if (a.g > m) m = a.g; if (a.b > m) m = a.b; return m; }
CTL_DEV float savg(Spec a) { float s = 0.0f; s += a.r; s += a.g; s += a.b; return s * (1.0f / 3); }
CTL_DEV Spec safe_sqrt(Spec a) { return mk_sp(sqrtf(fmaxf(0.0f, a.r)), sqrtf(fmaxf(0.0f, a.g)), sqrtf(fmaxf(0.0f, a.b))); }

// row-major float4x4 rows as float4 (Math/float4x4.h:365-408)
CTL_DEV float dot4(float4 a, float b0, float b1, float b2, float b3) { float r = 0.0f; r += a.x * b0; r += a.y * b1; r += a.z * b2; r += a.w * b3; return r; }
CTL_DEV V3 xf_point(const float4* __restrict__ m, V3 p) {
    float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2), r3 = __ldg(m + 3);
    float x = dot4(r0, p.x, p.y, p.z, 1.0f), y = dot4(r1, p.x, p.y, p.z, 1.0f), z = dot4(r2, p.x, p.y, p.z, 1.0f), w = dot4(r3, p.x, p.y, p.z, 1.0f);
    return mk(x / w, y / w, z / w);
}
CTL_DEV V3 xf_dir(const float4* __restrict__ m, V3 d) {
    float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
    return mk(dot4(r0, d.x, d.y, d.z, 0.0f), dot4(r1, d.x, d.y, d.z, 0.0f), dot4(r2, d.x, d.y, d.z, 0.0f));
}

struct Frame { V3 s, t, n; };
CTL_DEV V3 to_local(const Frame& f, V3 v) { return mk(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }
CTL_DEV V3 to_world(const Frame& f, V3 v) { return f.s * v.x + f.t * v.y + f.n * v.z; }

// 256-bit read-only global load (sm_100a LDG.E.256): one 64-byte BVH node = 2 of these = 2 L1 wavefronts per lane
// instead of 4 with float4 loads.  `p` must be 32-byte aligned.
struct F8 { float4 lo, hi; };
#ifndef CTL_NODE_EVICT_LAST
#define CTL_NODE_EVICT_LAST 0 // experiment: L1::evict_last on the BVH node loads (keep the top of the tree resident against the ray / triangle streams)
#endif
CTL_DEV F8 ldg256(const void* p) {
    F8 r;
#ifndef __CUDACC__   // host build of the kernel source (tests/traverse_emulate.cpp): a plain 32-byte read
    r.lo = ((const float4*)p)[0]; r.hi = ((const float4*)p)[1];
    return r;
#endif
#if CTL_NODE_EVICT_LAST
    asm("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
#endif
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
}

// Streaming (touch-once) data must not evict the BVH nodes and per-lane stacks the traversal kernel keeps in L1:
// triangles and queue records are loaded with L1::no_allocate.  (Storing the results with .cs was measured and rejected:
// evict-first partial-sector writes cost +66 % extension-traversal time, profiles/r01i_cache_hint_variants.log.)
#ifndef CTL_STREAM_HINTS
#define CTL_STREAM_HINTS 0 // measured: L1::no_allocate on triangle/queue loads = -1 % on C4 but +66 % on C2 extension rays (coherent rays re-use triangles through L1)
#endif
CTL_DEV float4 ldg_stream(const float4* p) {
#if CTL_STREAM_HINTS == 2
    float4 r;
    asm("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#elif CTL_STREAM_HINTS
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}
CTL_DEV uint32_t ldg_stream(const uint32_t* p) {
#if CTL_STREAM_HINTS == 2
    uint32_t r;
    asm("ld.global.nc.L1::evict_first.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
#elif CTL_STREAM_HINTS
    uint32_t r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}
CTL_DEV float h2f(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)(h & 0xffff))); }

} // namespace ctld
