// oracle.cpp -- CPU restatement of CudaTracerLib's ray-traversal + path-tracing hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (cudatracerlib_b200/) never links, imports or calls it.
//
// Parity status: PINNED.  The reference ships no tests / golden vectors, so this
// restatement is pinned against outputs of the reference itself run here:
//  (a) oracle/_ref -- the reference's OWN host code for this path (BVH traversal template, PathTrace<DIRECT>,
//      BSDFs, light, sensor, sampler, XORWOW, Image::AddSample) compiled from /root/reference by oracle/build_ref.sh;
//      tests/test_golden_cpu.py::test_oracle_vs_live_reference compares the two live where _ref exists;
//  (b) tests/golden/reference_golden.npz -- vectors minted from (a) by tests/golden/make_golden.py (committed), checked by
//      tests/test_golden_cpu.py on every box: RNG / sample tables / Woop encoding bit-exact, traversal indices identical,
//      BSDF tables to 2e-5, same-seed images to 1e-3;
//  (c) the known-answer values of SURVEY Appendix C (tests/test_oracle_kat.py).
//
// Each function cites the reference file:line (relative to the CudaTracerLib tree) it follows.
// Independent code: nothing here includes product sources except the public C header for
// the POD layouts of the data surface it reads.
//
// Arithmetic: IEEE fp32, no implicit FMA contraction (-ffp-contract=off).  The reference's
// GPU build lets nvcc contract a*b+c; the product writes the FMAs that matter explicitly
// (slab test, Woop test) and this file restates exactly those with fmaf() so that traversal
// parity is bit-exact.  Everything else is evaluated in the reference's host order.
#include <cmath>
#include <cstdint>
// -DORC_NO_FMA builds liboracle_nofma.so: the same restatement with a*b+c in two roundings, i.e. the arithmetic of the reference's HOST build
// (g++, no contraction).  tests/test_golden_cpu.py requires that variant to be bit-identical to oracle/_ref, which pins every line of this file;
// the default build differs from it only in the explicit FMAs below (what nvcc emits for the reference's GPU build and what the product computes).
#ifdef ORC_NO_FMA
#define ORC_FMA(a, b, c) ((a) * (b) + (c))
#else
#define ORC_FMA(a, b, c) fmaf((a), (b), (c))
#endif
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <climits>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
// toolkit XORWOW jump matrices (third-party: CUDA 12.9 cuRAND headers, not vendored)
#ifndef __device__
#define __device__
#define ORC_UNDEF_DEVICE
#endif
#include <curand_globals.h>
#include <curand_precalc.h>
#ifdef ORC_UNDEF_DEVICE
#undef __device__
#endif
#include "../include/ctl_b200.h"

namespace {

const float PI_F = 3.14159265358979f;             // Math/MathFunc.h:12
const float INV_PI_F = 1.0f / PI_F;               // :13
const float DELTA_EPS = 1e-3f;                    // :26
const unsigned E_DIFFUSE_REFL = 0x2, E_GLOSSY_REFL = 0x8, E_DELTA_REFL = 0x20, E_DELTA_TRANS = 0x40; // SceneTypes/Samples.h:32-71
const unsigned E_SMOOTH = 0x2 | 0x4 | 0x8 | 0x10, E_DELTA = 0x1 | 0x20 | 0x40, E_ALL = E_SMOOTH | E_DELTA | 0x80 | 0x100; // :73-92

struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 divs(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; } // Math/Vector.h:97
inline V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } // :329
inline float rcpf(float a) { return a != 0.0f ? 1.0f / a : 0.0f; }                                             // MathFunc.h:399
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { return mul(a, rcpf(length(a))); } // Vector.h:369-372

struct Spec { float r, g, b; };
inline Spec sp(float v) { Spec s = {v, v, v}; return s; }
inline Spec sp3(const float* p) { Spec s = {p[0], p[1], p[2]}; return s; }
inline Spec smul(Spec a, Spec b) { Spec s = {a.r * b.r, a.g * b.g, a.b * b.b}; return s; }
inline Spec smulf(Spec a, float f) { Spec s = {a.r * f, a.g * f, a.b * f}; return s; }
inline Spec sdivf(Spec a, float f) { const float recip = 1.0f / f; Spec s = {a.r * recip, a.g * recip, a.b * recip}; return s; } // Spectrum.h:122-128,150-155: reciprocal multiply
inline Spec sadd(Spec a, Spec b) { Spec s = {a.r + b.r, a.g + b.g, a.b + b.b}; return s; }
inline Spec ssub(Spec a, Spec b) { Spec s = {a.r - b.r, a.g - b.g, a.b - b.b}; return s; }
inline Spec sdiv(Spec a, Spec b) { Spec s = {a.r / b.r, a.g / b.g, a.b / b.b}; return s; }
inline bool sis_zero(Spec a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }
inline float smax(Spec a) { float m = a.r; if (a.g > m) m = a.g; if (a.b > m) m = a.b; return m; } // Spectrum.h:247
inline float savg(Spec a) { float s = 0.0f; s += a.r; s += a.g; s += a.b; return s * (1.0f / 3); }  // :180-190
inline Spec ssafe_sqrt(Spec a) { Spec s = {sqrtf(fmaxf(0.0f, a.r)), sqrtf(fmaxf(0.0f, a.g)), sqrtf(fmaxf(0.0f, a.b))}; return s; }

// ---- row-major float4x4 (Math/float4x4.h) --------------------------------
inline float dot4(const float* a, float b0, float b1, float b2, float b3) { float r = 0.0f; r += a[0] * b0; r += a[1] * b1; r += a[2] * b2; r += a[3] * b3; return r; }
inline V3 xf_point(const float* m, V3 p) { // :398-402
    float x = dot4(m, p.x, p.y, p.z, 1.0f), y = dot4(m + 4, p.x, p.y, p.z, 1.0f), z = dot4(m + 8, p.x, p.y, p.z, 1.0f), w = dot4(m + 12, p.x, p.y, p.z, 1.0f);
    return mk(x / w, y / w, z / w);
}
inline V3 xf_dir(const float* m, V3 d) { // :404-408
    return mk(dot4(m, d.x, d.y, d.z, 0.0f), dot4(m + 4, d.x, d.y, d.z, 0.0f), dot4(m + 8, d.x, d.y, d.z, 0.0f));
}
void mat_inverse(const float* q, float* out) { // :132-193, cofactor expansion in the reference's evaluation order
#define Q(i, j) q[(i) * 4 + (j)]
    float m00 = Q(0, 0), m01 = Q(0, 1), m02 = Q(0, 2), m03 = Q(0, 3), m10 = Q(1, 0), m11 = Q(1, 1), m12 = Q(1, 2), m13 = Q(1, 3);
    float m20 = Q(2, 0), m21 = Q(2, 1), m22 = Q(2, 2), m23 = Q(2, 3), m30 = Q(3, 0), m31 = Q(3, 1), m32 = Q(3, 2), m33 = Q(3, 3);
#undef Q
    float v0 = m20 * m31 - m21 * m30, v1 = m20 * m32 - m22 * m30, v2 = m20 * m33 - m23 * m30, v3 = m21 * m32 - m22 * m31, v4 = m21 * m33 - m23 * m31, v5 = m22 * m33 - m23 * m32;
    float t00 = +(v5 * m11 - v4 * m12 + v3 * m13), t10 = -(v5 * m10 - v2 * m12 + v1 * m13), t20 = +(v4 * m10 - v2 * m11 + v0 * m13), t30 = -(v3 * m10 - v1 * m11 + v0 * m12);
    float id = 1 / (t00 * m00 + t10 * m01 + t20 * m02 + t30 * m03);
    float d00 = t00 * id, d10 = t10 * id, d20 = t20 * id, d30 = t30 * id;
    float d01 = -(v5 * m01 - v4 * m02 + v3 * m03) * id, d11 = +(v5 * m00 - v2 * m02 + v1 * m03) * id, d21 = -(v4 * m00 - v2 * m01 + v0 * m03) * id, d31 = +(v3 * m00 - v1 * m01 + v0 * m02) * id;
    v0 = m10 * m31 - m11 * m30; v1 = m10 * m32 - m12 * m30; v2 = m10 * m33 - m13 * m30; v3 = m11 * m32 - m12 * m31; v4 = m11 * m33 - m13 * m31; v5 = m12 * m33 - m13 * m32;
    float d02 = +(v5 * m01 - v4 * m02 + v3 * m03) * id, d12 = -(v5 * m00 - v2 * m02 + v1 * m03) * id, d22 = +(v4 * m00 - v2 * m01 + v0 * m03) * id, d32 = -(v3 * m00 - v1 * m01 + v0 * m02) * id;
    v0 = m21 * m10 - m20 * m11; v1 = m22 * m10 - m20 * m12; v2 = m23 * m10 - m20 * m13; v3 = m22 * m11 - m21 * m12; v4 = m23 * m11 - m21 * m13; v5 = m23 * m12 - m22 * m13;
    float d03 = -(v5 * m01 - v4 * m02 + v3 * m03) * id, d13 = +(v5 * m00 - v2 * m02 + v1 * m03) * id, d23 = -(v4 * m00 - v2 * m01 + v0 * m03) * id, d33 = +(v3 * m00 - v1 * m01 + v0 * m02) * id;
    float r[16] = {d00, d01, d02, d03, d10, d11, d12, d13, d20, d21, d22, d23, d30, d31, d32, d33};
    memcpy(out, r, 64);
}

// ---- half / normal codec (Math/half.h:20-82 IEEE form; Math/Compression.h:12-31) ---
uint16_t f2h(float f) {
    uint32_t ia; memcpy(&ia, &f, 4);
    uint16_t ir = (uint16_t)((ia >> 16) & 0x8000);
    uint32_t e = ia & 0x7f800000;
    if (e == 0x7f800000) { if ((ia & 0x7fffffff) == 0x7f800000) ir |= 0x7c00; else ir = 0x7fff; return ir; }
    if (e < 0x33000000) return ir;
    int shift = (int)((ia >> 23) & 0xff) - 127;
    if (shift > 15) return ir | 0x7c00;
    ia = (ia & 0x007fffff) | 0x00800000;
    if (shift < -14) { ir |= (uint16_t)(ia >> (-1 - shift)); ia <<= (32 - (-1 - shift)); }
    else { ir |= (uint16_t)(ia >> 13); ia <<= 19; ir = (uint16_t)(ir + ((14 + shift) << 10)); }
    if (ia > 0x80000000u || (ia == 0x80000000u && (ir & 1))) ir++;
    return ir;
}
float h2f(uint16_t h) { // IEEE decode = device __half2float
    int e = (h >> 10) & 31, m = h & 1023;
    float v;
    if (e == 0) v = ldexpf((float)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexpf((float)(m | 1024), e - 25);
    return (h & 0x8000) ? -v : v;
}
uint16_t enc_normal(V3 v) {
    float theta = (acosf(v.z) * (255.0f / PI_F));
    float phi = (atan2f(v.y, v.x) * (255.0f / (2.0f * PI_F)));
    phi = phi < 0 ? (phi + 255) : phi;
    return (uint16_t)(((unsigned short)theta << 8) | (unsigned short)phi);
}
V3 dec_normal(uint16_t c) {
    const float PI_4 = PI_F / 4.0f, PI_2 = PI_F / 2.0f;
    unsigned char x = c >> 8, y = c & 0xff;
    float theta = x == 63 ? PI_4 : (x == 127 ? PI_2 : (x == 191 ? 3 * PI_4 : float(x) * (1.0f / 255.0f) * PI_F));
    float phi = y == 63 ? PI_2 : (y == 127 ? PI_F : (y == 191 ? 3 * PI_2 : float(y) * (1.0f / 255.0f) * PI_F * 2.0f));
    float sp_ = sinf(phi), cp = cosf(phi), st = sinf(theta), ct = cosf(theta); // host sincos = sinf/cosf, MathFunc.h:70-74
    return mk(st * cp, st * sp_, ct);
}
void coordinate_system(V3 a, V3& s, V3& t) { // Math/Frame.h:9-22
    if (fabsf(a.x) > fabsf(a.y)) { float il = 1.0f / sqrtf(a.x * a.x + a.z * a.z); t = mk(a.z * il, 0.0f, -a.x * il); }
    else { float il = 1.0f / sqrtf(a.y * a.y + a.z * a.z); t = mk(0.0f, a.z * il, -a.y * il); }
    s = normalize(cross(t, a));
}

// ---- XORWOW (Base/CudaRandom.h:108-291, .cu:7-34; curand_kernel.h) ---------
struct Xorwow {
    uint32_t v[5], d;
    uint32_t next() { // CudaRandom.h:112-123
        uint32_t t = (v[0] ^ (v[0] >> 2));
        v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
        v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
        d += 362437;
        return v[4] + d;
    }
    float random_float() { // :124-127, .cu:7-16
        float f = next() * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
        return f * (1 - 1e-5f);
    }
};
void matvec(const uint32_t* vec, const uint32_t* mat, uint32_t* res, int n) {
    for (int i = 0; i < n; i++) res[i] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 32; j++)
            if (vec[i] & (1u << j))
                for (int k = 0; k < n; k++) res[k] ^= mat[n * (i * 32 + j) + k];
}
void matmat(uint32_t* A, const uint32_t* B, int n) {
    uint32_t res[8];
    for (int i = 0; i < n * 32; i++) { matvec(A + i * n, B, res, n); for (int j = 0; j < n; j++) A[i * n + j] = res[j]; }
}
// generic skip-ahead with the toolkit's precalculated matrices (CudaRandom.h:160-242)
void skipahead(unsigned long long x, Xorwow& st, unsigned int (*precalc)[800], bool is_offset) {
    const int n = 5;
    std::vector<uint32_t> matrix(n * n * 32), matrixA(n * n * 32);
    uint32_t vec[5], res[5];
    unsigned long long p = x;
    for (int i = 0; i < n; i++) vec[i] = st.v[i];
    int mn = 0;
    while (p && mn < PRECALC_NUM_MATRICES - 1) {
        for (unsigned t = 0; t < (p & PRECALC_BLOCK_MASK); t++) { matvec(vec, precalc[mn], res, n); memcpy(vec, res, sizeof(res)); }
        p >>= PRECALC_BLOCK_SIZE; mn++;
    }
    if (p) { memcpy(matrix.data(), precalc[PRECALC_NUM_MATRICES - 1], n * n * 32 * 4); matrixA = matrix; }
    while (p) {
        for (unsigned t = 0; t < (p & SKIPAHEAD_MASK); t++) { matvec(vec, matrixA.data(), res, n); memcpy(vec, res, sizeof(res)); }
        p >>= SKIPAHEAD_BLOCKSIZE;
        if (p) for (int i = 0; i < SKIPAHEAD_BLOCKSIZE; i++) { matmat(matrix.data(), matrixA.data(), n); matrixA = matrix; }
    }
    for (int i = 0; i < n; i++) st.v[i] = vec[i];
    if (is_offset) st.d += 362437 * (unsigned int)x;
}
void xorwow_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset, Xorwow& st) { // :243-262
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    st.d = 6615241 + t1 + t0;
    st.v[0] = 123456789u + t0; st.v[1] = 362436069u ^ t0; st.v[2] = 521288629u + t1; st.v[3] = 88675123u ^ t1; st.v[4] = 5783321u + t0;
    skipahead(subsequence, st, precalc_xorwow_matrix_host, false);
    skipahead(offset, st, precalc_xorwow_offset_matrix_host, true);
}

const int N_SEQ = 4096, SEQ_LEN = 30; // Kernel/TraceHelper.cu:257

// Kernel/Sampler.h:36-85: per pass, for every sequence: 30 1-D draws then 30 2-D draws.
// Vec2f(rng.randomFloat(), rng.randomFloat()) evaluates its arguments right-to-left with g++
// (unspecified order in C++; observed with g++ 13.3 and matches oracle/_ref): the FIRST draw is y.
void fill_tables(Xorwow& rng, float* d1, float* d2) {
    for (int s = 0; s < N_SEQ; s++) {
        for (int i = 0; i < SEQ_LEN; i++) d1[i * N_SEQ + s] = rng.random_float();
        for (int i = 0; i < SEQ_LEN; i++) {
            float y = rng.random_float(), x = rng.random_float();
            d2[(i * N_SEQ + s) * 2 + 0] = x; d2[(i * N_SEQ + s) * 2 + 1] = y;
        }
    }
}

// Kernel/Sampler_device.h:62-107
struct Sampler {
    const float *d1, *d2;
    unsigned idx, i1, i2;
    float f1() {
        unsigned a = idx % N_SEQ, b = (idx / N_SEQ) % N_SEQ, e = i1 % SEQ_LEN;
        float sum = 0.0f; sum += d1[e * N_SEQ + a]; sum += d1[e * N_SEQ + b];
        i1++;
        return sum - floorf(sum);
    }
    void f2(float& x, float& y) {
        unsigned a = idx % N_SEQ, b = (idx / N_SEQ) % N_SEQ, e = i2 % SEQ_LEN;
        float sx = 0.0f, sy = 0.0f;
        sx += d2[(e * N_SEQ + a) * 2]; sy += d2[(e * N_SEQ + a) * 2 + 1];
        sx += d2[(e * N_SEQ + b) * 2]; sy += d2[(e * N_SEQ + b) * 2 + 1];
        i2++;
        x = sx - floorf(sx); y = sy - floorf(sy);
    }
};

// ---- traversal -------------------------------------------------------------
struct Hit { float dist, u, v; uint32_t tri, node; };
struct Counters { uint64_t inner, tris, inst; std::vector<uint8_t>* ev = nullptr; }; // ev: optional per-ray event string (N/I/T) for scheduling studies

inline float guard_inv(float d) { // BVHTraversal.h:16-19
    const float ooeps = powf(2.0f, -80.0f);
    return 1.0f / (fabsf(d) > ooeps ? d : copysignf(ooeps, d));
}

// Generic stack traversal, host semantics of TracerayTemplate (BVHTraversal.h:122-232):
// both children tested, nearer first (swp = c1min < c0min), far pushed, a found leaf is processed
// as soon as it is met (host branch: mask = leafAddr >= 0), box entry clamped at tmin_box.
template <typename LEAF>
inline bool traverse(const float* nodes4, int node_off4, int start, V3 o, V3 d, float tmin_box, float& rayT, Counters* cnt, LEAF leaf, const bool* stop = nullptr) {
    const int SENT = CTL_SENTINEL;
    if (start < 0) return leaf(~start);
    bool found = false;
    int stack[64]; int sp_ = 0; stack[0] = SENT;
    float idx = guard_inv(d.x), idy = guard_inv(d.y), idz = guard_inv(d.z);
    float oodx = o.x * idx, oody = o.y * idy, oodz = o.z * idz;
    int leafAddr = 0, nodeAddr = start;
    while (nodeAddr != SENT && !(stop && *stop)) { // any-hit: the first hit ends the ray at once (TraceHelper.cu:675-679), no further node is popped
        while ((unsigned)nodeAddr < (unsigned)SENT) {
            const float* n = nodes4 + (size_t)(node_off4 + nodeAddr) * 4;
            if (cnt) { cnt->inner++; if (cnt->ev) cnt->ev->push_back('N'); }
            int c0, c1; memcpy(&c0, n + 12, 4); memcpy(&c1, n + 13, 4);
            float c0lox = ORC_FMA(n[0], idx, -oodx), c0hix = ORC_FMA(n[1], idx, -oodx), c0loy = ORC_FMA(n[2], idy, -oody), c0hiy = ORC_FMA(n[3], idy, -oody);
            float c0loz = ORC_FMA(n[8], idz, -oodz), c0hiz = ORC_FMA(n[9], idz, -oodz), c1loz = ORC_FMA(n[10], idz, -oodz), c1hiz = ORC_FMA(n[11], idz, -oodz);
            float c1lox = ORC_FMA(n[4], idx, -oodx), c1hix = ORC_FMA(n[5], idx, -oodx), c1loy = ORC_FMA(n[6], idy, -oody), c1hiy = ORC_FMA(n[7], idy, -oody);
            // spanBegin/EndKepler (MathFunc.h:443-444) == float min/max for t >= 0
            float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin_box));
            float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rayT));
            float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin_box));
            float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rayT));
            bool swp = (c1min < c0min), t0 = (c0max >= c0min), t1 = (c1max >= c1min);
            if (!t0 && !t1) { nodeAddr = stack[sp_]; sp_--; }
            else {
                nodeAddr = t0 ? c0 : c1;
                if (t0 && t1) { if (swp) { int tmp = nodeAddr; nodeAddr = c1; c1 = tmp; } sp_++; stack[sp_] = c1; }
            }
            if (nodeAddr < 0 && leafAddr >= 0) { leafAddr = nodeAddr; nodeAddr = stack[sp_]; sp_--; }
            if (leafAddr < 0) break;
        }
        while (leafAddr < 0) {
            found |= leaf(~leafAddr);
            if (stop && *stop) break;
            leafAddr = nodeAddr;
            if (nodeAddr < 0) { nodeAddr = stack[sp_]; sp_--; }
        }
    }
    return found;
}

// Woop test with the explicit FMA pattern shared with the device kernel (TraceHelper.cu:118-134)
inline bool woop_test(const float* w, V3 o, V3 d, float tlo, float thi, float& t, float& u, float& v) {
    float Oz = ORC_FMA(-o.z, w[2], ORC_FMA(-o.y, w[1], ORC_FMA(-o.x, w[0], w[3])));
    float invDz = 1.0f / ORC_FMA(d.z, w[2], ORC_FMA(d.y, w[1], d.x * w[0]));
    t = Oz * invDz;
    if (t > tlo && t < thi) {
        float Ox = ORC_FMA(o.z, w[6], ORC_FMA(o.y, w[5], ORC_FMA(o.x, w[4], w[7])));
        float Dx = ORC_FMA(d.z, w[6], ORC_FMA(d.y, w[5], d.x * w[4]));
        u = ORC_FMA(t, Dx, Ox);
        if (u >= 0.0f) {
            float Oy = ORC_FMA(o.z, w[10], ORC_FMA(o.y, w[9], ORC_FMA(o.x, w[8], w[11])));
            float Dy = ORC_FMA(d.z, w[10], ORC_FMA(d.y, w[9], d.x * w[8]));
            v = ORC_FMA(t, Dy, Oy);
            if (v >= 0.0f && u + v <= 1.0f) return true;
        }
    }
    return false;
}

// two-level closest hit: __traceRay_internal__<false> (TraceHelper.cu:88-172).
// tri_lo: lower t bound for triangles (rayEps for traceRay, ray.tmin for intersectKernel),
// box_lo: lower bound for boxes (0 for traceRay, ray.tmin for intersectKernel). any_hit: intersectKernel<true>.
bool trace(const ctl_scene_view& S, V3 ori, V3 dir, float tri_lo, float box_lo, Hit& hit, Counters* cnt, bool any_hit) {
    if (!S.n_nodes) return false;
    bool stop = false;
    return traverse((const float*)S.scene_bvh_nodes, 0, S.scene_start_node, ori, dir, box_lo, hit.dist, cnt, [&](int nodeIdx) {
        if (stop) return false;
        if (cnt) { cnt->inst++; if (cnt->ev) cnt->ev->push_back('I'); }
        const ctl_node& N = S.nodes[nodeIdx];
        const ctl_mesh& mesh = S.meshes[N.mesh_index];
        const float* inv = S.node_inv_xf + (size_t)nodeIdx * 16;
        V3 d = xf_dir(inv, dir), o = xf_point(inv, ori);
        return traverse((const float*)S.bvh_nodes, (int)mesh.bvh_node_offset, 0, o, d, box_lo, hit.dist, cnt, [&](int triIdx) {
            bool found = false;
            if (stop) return false;
            for (int triAddr = triIdx;; triAddr++) {
                const float* w = (const float*)S.woop + ((size_t)mesh.bvh_tri_offset + (size_t)triAddr * 3) * 4;
                uint32_t index = S.tri_index[mesh.bvh_idx_offset + triAddr];
                if (cnt) { cnt->tris++; if (cnt->ev) cnt->ev->push_back('T'); }
                float t, u, v;
                if (woop_test(w, o, d, tri_lo, hit.dist, t, u, v)) {
                    hit.node = (uint32_t)nodeIdx; hit.tri = (index >> 1) + mesh.tri_offset; hit.u = u; hit.v = v; hit.dist = t;
                    found = true;
                    if (any_hit) { stop = true; break; }
                }
                if (index & 1) break;
            }
            return found;
        }, &stop);
    }, &stop);
}

// ---- fillDG (TraceHelper.cu:274-307 -> TriangleData.cu:75-103) --------------
struct Frame { V3 s, t, n; };
struct DG { V3 P; Frame sys; V3 n; };
inline V3 to_local(const Frame& f, V3 v) { return mk(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }                  // Frame.h:38-40
inline V3 to_world(const Frame& f, V3 v) { return add(add(mul(f.s, v.x), mul(f.t, v.y)), mul(f.n, v.z)); }      // :41-43

void fill_dg(const ctl_scene_view& S, float bu, float bv, uint32_t tri, uint32_t node, DG& dg) {
    const float* l2w = S.node_xf + (size_t)node * 16;
    const uint32_t* w = S.tri_data[tri].w;
    V3 na = dec_normal(w[0] & 0xffff), nb = dec_normal(w[0] >> 16), nc = dec_normal(w[1] & 0xffff);
    float ww = 1.0f - bu - bv, u = bu, v = bv;
    V3 n = normalize(add(add(mul(na, u), mul(nb, v)), mul(nc, ww)));
    V3 dpdu = mk(h2f(w[2] & 0xffff), h2f(w[2] >> 16), h2f(w[3] & 0xffff));
    V3 dpdv = mk(h2f(w[3] >> 16), h2f(w[4] & 0xffff), h2f(w[4] >> 16));
    V3 s = sub(dpdu, mul(n, dot(n, dpdu)));
    V3 t = cross(s, n);
    s = xf_dir(l2w, s); t = xf_dir(l2w, t);
    dg.sys.s = normalize(s); dg.sys.t = normalize(t); dg.sys.n = normalize(cross(t, s));
    V3 wdpdu = xf_dir(l2w, dpdu), wdpdv = xf_dir(l2w, dpdv);
    dg.n = normalize(cross(wdpdu, wdpdv));
    if (dot(dg.n, dg.sys.n) < 0.0f) dg.n = neg(dg.n);
}

// ---- warps (Math/Warp.h:61-164) --------------------------------------------
void concentric_disk(float sx, float sy, float& px, float& py) {
    float r1 = 2.0f * sx - 1.0f, r2 = 2.0f * sy - 1.0f, phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (PI_F / 4.0f) * (r2 / r1); }
    else { r = r2; phi = (PI_F / 2.0f) - (r1 / r2) * (PI_F / 4.0f); }
    float cp = cosf(phi), sp_ = sinf(phi);
    px = r * cp; py = r * sp_;
}
V3 cosine_hemisphere(float sx, float sy) {
    float px, py; concentric_disk(sx, sy, px, py);
    float z = sqrtf(1.0f - px * px - py * py);
    return mk(px, py, z);
}

// ---- Fresnel (Math/FresnelHelper.h:27-58, 119-147) ---------------------------
float fresnel_dielectric_ext(float cosThetaI_, float& cosThetaT_, float eta) {
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    float scale = (cosThetaI_ > 0) ? 1.0f / eta : eta, cosThetaTSqr = 1.0f - (1.0f - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    float cosThetaI = fabsf(cosThetaI_), cosThetaT = sqrtf(fmaxf(0.0f, cosThetaTSqr));
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
Spec fresnel_conductor_exact(float cosThetaI, Spec eta, Spec k) {
    float c2 = cosThetaI * cosThetaI, s2 = 1 - c2, s4 = s2 * s2;
    Spec temp1 = ssub(ssub(smul(eta, eta), smul(k, k)), sp(s2));
    Spec a2pb2 = ssafe_sqrt(sadd(smul(temp1, temp1), smulf(smul(smul(smul(k, k), eta), eta), 4)));
    Spec a = ssafe_sqrt(smulf(sadd(a2pb2, temp1), 0.5f));
    Spec term1 = sadd(a2pb2, sp(c2)), term2 = smulf(a, (2 * cosThetaI));
    Spec Rs2 = sdiv(ssub(term1, term2), sadd(term1, term2));
    Spec term3 = sadd(smulf(a2pb2, c2), sp(s4)), term4 = smulf(term2, s2);
    Spec Rp2 = sdiv(smul(Rs2, ssub(term3, term4)), sadd(term3, term4));
    return smulf(sadd(Rp2, Rs2), 0.5f);
}

// ---- microfacet distribution (Engine/MicrofacetDistribution.h/.cu), Beckmann + GGX, visible normals
float m_erfinv(float x) { // MathFunc.h:343-373
    float w = -logf((1.0f - x) * (1.0f + x)), p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w; p = -4.39150654e-06f + p * w; p = 0.00021858087f + p * w;
        p = -0.00125372503f + p * w; p = -0.00417768164f + p * w; p = 0.246640727f + p * w; p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3;
        p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w; p = -0.00367342844f + p * w; p = 0.00573950773f + p * w;
        p = -0.0076224613f + p * w; p = 0.00943887047f + p * w; p = 1.00167406f + p * w; p = 2.83297682f + p * w;
    }
    return p * x;
}
float m_erf(float x) { // :375-393
    float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    float sign = copysignf(1.0f, x); x = fabsf(x);
    float t = 1.0f / (1.0f + p * x);
    float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * expf(-x * x);
    return sign * y;
}
float m_hypot2(float a, float b) { // :326-341
    float r;
    if (fabsf(a) > fabsf(b)) { r = b / a; r = fabsf(a) * sqrtf(1.0f + r * r); }
    else if (b != 0.0f) { r = a / b; r = fabsf(b) * sqrtf(1.0f + r * r); }
    else r = 0.0f;
    return r;
}
struct Distr {
    int type; float au, av;
    Distr(int t, float a, float b) : type(t), au(a > 1e-4f ? a : 1e-4f), av(b > 1e-4f ? b : 1e-4f) {} // .h:37-44
    float eval(V3 m) const { // .cu:6-42
        if (m.z <= 0) return 0.0f;
        float c2 = m.z * m.z;
        float be = ((m.x * m.x) / (au * au) + (m.y * m.y) / (av * av)) / c2;
        float result;
        if (type == CTL_DISTR_BECKMANN) result = expf(-be) / (PI_F * au * av * c2 * c2);
        else { float root = (1 + be) * c2; result = 1.0f / (PI_F * au * av * root * root); }
        if (result < 1e-20f) result = 0;
        return result;
    }
    float project_roughness(V3 v) const { // .h:126-136
        float invSinTheta2 = 1 / (1.0f - v.z * v.z);
        if (au == av || invSinTheta2 <= 0) return au;
        float cosPhi2 = v.x * v.x * invSinTheta2, sinPhi2 = v.y * v.y * invSinTheta2;
        return sqrtf(cosPhi2 * au * au + sinPhi2 * av * av);
    }
    float smith_g1(V3 v, V3 m) const { // .cu:306-340
        if (dot(v, m) * v.z <= 0) return 0.0f;
        float temp = 1 - v.z * v.z;
        float tanTheta = fabsf(temp <= 0.0f ? 0.0f : sqrtf(temp) / v.z); // Frame::tanTheta, Frame.h:90-95
        if (tanTheta == 0.0f) return 1.0f;
        float alpha = project_roughness(v);
        if (type == CTL_DISTR_BECKMANN) {
            float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        }
        float root = alpha * tanTheta;
        return 2.0f / (1.0f + m_hypot2(1.0f, root));
    }
    float pdf_visible(V3 wi, V3 m) const { // .h:114-119
        if (wi.z == 0) return 0.0f;
        return smith_g1(wi, m) * fabsf(dot(wi, m)) * eval(m) / fabsf(wi.z);
    }
    void sample_visible11(float thetaI, float sx, float sy, float& slx, float& sly) const { // .cu:188-304
        const float SQRT_PI_INV = 1 / sqrtf(PI_F);
        if (type == CTL_DISTR_BECKMANN) {
            if (thetaI < 1e-4f) {
                float r = sqrtf(-logf(1.0f - sx));
                float sinPhi = sinf(2 * PI_F * sy), cosPhi = cosf(2 * PI_F * sy);
                slx = r * cosPhi; sly = r * sinPhi; return;
            }
            float tanThetaI = tanf(thetaI), cotThetaI = 1 / tanThetaI;
            float a = -1, c = m_erf(cotThetaI);
            float sample_x = sx > 1e-6f ? sx : 1e-6f;
            float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1 + c) * powf(1 - sample_x, fit);
            float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * expf(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                float invErf = m_erfinv(b);
                float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * expf(-invErf * invErf)) - sample_x;
                float derivative = normalization * (1 - invErf * tanThetaI);
                if (fabsf(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slx = m_erfinv(b);
            sly = m_erfinv(2.0f * (sy > 1e-6f ? sy : 1e-6f) - 1.0f);
            return;
        }
        if (thetaI < 1e-4f) {
            float r = sqrtf(fmaxf(0.0f, sx / (1 - sx)));
            float sinPhi = sinf(2 * PI_F * sy), cosPhi = cosf(2 * PI_F * sy);
            slx = r * cosPhi; sly = r * sinPhi; return;
        }
        float tanThetaI = tanf(thetaI), a = 1 / tanThetaI;
        float G1 = 2.0f / (1.0f + sqrtf(fmaxf(0.0f, 1.0f + 1.0f / (a * a))));
        float A = 2.0f * sx / G1 - 1.0f;
        if (fabsf(A) == 1) A -= copysignf(1.0f, A) * 1e-7f;
        float tmp = 1.0f / (A * A - 1.0f), B = tanThetaI;
        float D = sqrtf(fmaxf(0.0f, B * B * tmp * tmp - (A * A - B * B) * tmp));
        float s1 = B * tmp - D, s2 = B * tmp + D;
        slx = (A < 0.0f || s2 > 1.0f / tanThetaI) ? s1 : s2;
        float Sg;
        if (sy > 0.5f) { Sg = 1.0f; sy = 2.0f * (sy - 0.5f); } else { Sg = -1.0f; sy = 2.0f * (0.5f - sy); }
        float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                  (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
        sly = Sg * z * sqrtf(1.0f + slx * slx);
    }
    V3 sample_visible(V3 _wi, float sx, float sy) const { // .cu:151-186
        V3 wi = normalize(mk(au * _wi.x, av * _wi.y, _wi.z));
        float theta = 0, phi = 0;
        if (wi.z < 0.99999f) { theta = acosf(wi.z); phi = atan2f(wi.y, wi.x); }
        float sinPhi = sinf(phi), cosPhi = cosf(phi);
        float slx, sly; sample_visible11(theta, sx, sy, slx, sly);
        float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
        rx *= au; ry *= av;
        float nrm = 1.0f / sqrtf(rx * rx + ry * ry + (float)1.0);
        return mk(-rx * nrm, -ry * nrm, nrm);
    }
};

// ---- BSDFs (SceneTypes/BSDF_Simple.cu:7-75 diffuse, 174-277 dielectric, 662-763 roughconductor) ---
struct BRec { DG dg; V3 wi, wo; float eta; unsigned typeMask, sampledType; };

unsigned bsdf_combined_type(const ctl_material& m) {
    return m.bsdf_type == CTL_BSDF_DIFFUSE ? E_DIFFUSE_REFL : (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR ? E_GLOSSY_REFL : (E_DELTA_REFL | E_DELTA_TRANS));
}
float mat_alpha(float a) { return savg(sp(a)); } // m_alphaU.Evaluate(dg).avg(), BSDF_Simple.cu:669-672

Spec bsdf_sample_inner(const ctl_material& m, BRec& b, float& pdf, float sx, float sy) {
    if (m.bsdf_type == CTL_BSDF_DIFFUSE) {
        if (!(b.typeMask & E_DIFFUSE_REFL) || b.wi.z <= 0) return sp(0.0f);
        b.sampledType = E_DIFFUSE_REFL;
        b.wo = cosine_hemisphere(sx, sy);
        b.eta = 1.0f;
        pdf = fabsf(INV_PI_F * b.wo.z) * 1;
        return smulf(sp3(m.reflectance), 1);
    }
    if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) {
        if (b.wi.z < 0 || !(b.typeMask & E_GLOSSY_REFL)) return sp(0.0f);
        Distr distr(m.distr_type, mat_alpha(m.alpha_u), mat_alpha(m.alpha_v));
        V3 mm = distr.sample_visible(b.wi, sx, sy);
        pdf = distr.pdf_visible(b.wi, mm);
        if (pdf == 0) return sp(0.0f);
        b.wo = normalize(sub(mul(mm, 2 * dot(b.wi, mm)), b.wi)); // FresnelHelper::reflect, FresnelHelper.h:144-147
        b.eta = 1.0f; b.sampledType = E_GLOSSY_REFL;
        if (b.wo.z <= 0) return sp(0.0f);
        Spec F = smul(fresnel_conductor_exact(dot(b.wi, mm), sp3(m.eta), sp3(m.k)), sp3(m.reflectance));
        float weight = distr.smith_g1(b.wo, mm);
        pdf /= 4.0f * dot(b.wo, mm);
        return smulf(F, weight);
    }
    // dielectric, no dispersion: eta_pdf = 1, f_o = 1 (SceneTypes/Dispersion.h:125-138)
    bool sR = (b.typeMask & E_DELTA_REFL) != 0, sT = (b.typeMask & E_DELTA_TRANS) != 0;
    float cosThetaT, eta = m.eta[0], invEta = 1.0f / eta;
    float F = fresnel_dielectric_ext(b.wi.z, cosThetaT, eta);
    auto refract = [&](V3 wi) { float scale = -(cosThetaT < 0 ? invEta : eta); return normalize(mk(scale * wi.x, scale * wi.y, cosThetaT)); }; // Frame.h:143-152
    if (sT && sR) {
        if (sx <= F) { b.sampledType = E_DELTA_REFL; b.wo = mk(-b.wi.x, -b.wi.y, b.wi.z); b.eta = 1.0f; pdf = F; return sp3(m.reflectance); }
        b.sampledType = E_DELTA_TRANS; b.wo = refract(b.wi); b.eta = cosThetaT < 0 ? eta : invEta; pdf = (1 - F) * 1.0f;
        float factor = cosThetaT < 0 ? invEta : eta;
        return smulf(smul(sp(1.0f), sp(m.transmittance)), (factor * factor));
    } else if (sR) { b.sampledType = E_DELTA_REFL; b.wo = mk(-b.wi.x, -b.wi.y, b.wi.z); b.eta = 1.0f; pdf = 1.0f; return sp3(m.reflectance); }
    else if (sT) {
        b.sampledType = E_DELTA_TRANS; b.wo = refract(b.wi); b.eta = cosThetaT < 0 ? eta : invEta; pdf = 1.0f;
        float factor = cosThetaT < 0 ? invEta : eta;
        return smulf(smul(sp(1.0f), sp(m.transmittance)), (factor * factor * (1 - F)));
    }
    return sp(0.0f);
}
Spec bsdf_f_inner(const ctl_material& m, const BRec& b) { // measure = ESolidAngle
    if (m.bsdf_type == CTL_BSDF_DIFFUSE) {
        if (!(b.typeMask & E_DIFFUSE_REFL)) return sp(0.0f);
        bool validRefl = b.wi.z > 0 && b.wo.z > 0;
        Spec s = smulf(sp3(m.reflectance), (INV_PI_F * fabsf(b.wo.z)));
        return validRefl ? s : sp(0.0f);
    }
    if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) {
        if (b.wi.z < 0 || b.wo.z < 0 || !(b.typeMask & E_GLOSSY_REFL)) return sp(0.0f);
        V3 H = normalize(add(b.wo, b.wi));
        Distr distr(m.distr_type, mat_alpha(m.alpha_u), mat_alpha(m.alpha_v));
        float D = distr.eval(H);
        if (D == 0) return sp(0.0f);
        Spec F = smul(fresnel_conductor_exact(dot(b.wi, H), sp3(m.eta), sp3(m.k)), sp3(m.reflectance));
        float G = distr.smith_g1(b.wi, H) * distr.smith_g1(b.wo, H);
        float value = D * G / (4.0f * b.wi.z);
        return smulf(F, value);
    }
    return sp(0.0f); // dielectric: delta lobes have no solid-angle density (BSDF_Simple.cu:227-252)
}
float bsdf_pdf_inner(const ctl_material& m, const BRec& b) {
    if (m.bsdf_type == CTL_BSDF_DIFFUSE) {
        if (!(b.typeMask & E_DIFFUSE_REFL)) return 0.0f;
        bool validRefl = b.wi.z > 0 && b.wo.z > 0;
        return validRefl ? fabsf(INV_PI_F * b.wo.z) : 0.0f;
    }
    if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) {
        if (b.wi.z < 0 || b.wo.z < 0 || !(b.typeMask & E_GLOSSY_REFL)) return 0.0f;
        V3 H = normalize(add(b.wo, b.wi));
        Distr distr(m.distr_type, mat_alpha(m.alpha_u), mat_alpha(m.alpha_v));
        return distr.eval(H) * distr.smith_g1(b.wi, H) / (4.0f * b.wi.z);
    }
    return 0.0f;
}
// BSDFALL two-sided wrapper (SceneTypes/BSDF.h:141-207)
Spec bsdf_sample(const ctl_material& m, BRec& b, float& pdf, float sx, float sy) {
    bool flip = b.wi.z < 0 && (m.flags & CTL_MAT_TWO_SIDED);
    if (flip) b.wi.z *= -1.0f;
    Spec r = bsdf_sample_inner(m, b, pdf, sx, sy);
    if (flip) { b.wi.z *= -1.0f; b.wo.z *= -1.0f; }
    return r;
}
Spec bsdf_f(const ctl_material& m, BRec& b) {
    bool flip = b.wi.z < 0 && (m.flags & CTL_MAT_TWO_SIDED);
    if (flip) b.wi.z *= -1.0f;
    Spec r = bsdf_f_inner(m, b);
    if (flip) { b.wi.z *= -1.0f; b.wo.z *= -1.0f; }
    return r;
}
float bsdf_pdf(const ctl_material& m, BRec& b) {
    bool flip = b.wi.z < 0 && (m.flags & CTL_MAT_TWO_SIDED);
    if (flip) b.wi.z *= -1.0f;
    float r = bsdf_pdf_inner(m, b);
    if (flip) { b.wi.z *= -1.0f; b.wo.z *= -1.0f; }
    return r;
}

// ---- area light (SceneTypes/Light.cu:67-155; Engine/ShapeSet.cu:51-69; Math/MonteCarlo.cu:7-14) ---
struct DRec { V3 p, n; float pdf; V3 ref, refN, d; float dist; };

const float* lower_bound_f(const float* first, const float* last, float val) { // Base/STL.h:40-57
    unsigned count = (unsigned)(last - first);
    for (; 0 < count;) { unsigned c2 = count / 2; const float* mid = first + c2; if (*mid < val) { first = ++mid; count -= c2 + 1; } else count = c2; }
    return first;
}
const float* upper_bound_f(const float* first, const float* last, float val) { // :21-38
    unsigned count = (unsigned)(last - first);
    for (; 0 < count;) { unsigned c2 = count / 2; const float* mid = first + c2; if (!(val < *mid)) { first = ++mid; count -= c2 + 1; } else count = c2; }
    return first;
}
Spec light_sample_direct(const ctl_scene_view& S, const ctl_light& L, DRec& dRec, float sx, float sy) {
    const float* cdf = S.light_cdf_data + L.cdf_offset;
    const float* entry = lower_bound_f(cdf, cdf + L.count + 1, sy);
    int ii = (int)(entry - cdf) - 1; if (ii < 0) ii = 0; if (ii > (int)L.count - 1) ii = (int)L.count - 1;
    unsigned index = (unsigned)ii;
    float pdf = cdf[index + 1] - cdf[index];
    sy = (sy - cdf[index]) / pdf;
    const ctl_light_tri& sn = S.light_tris[L.tri_offset + index];
    float a = sqrtf(1.0f - sx); float b0 = 1 - a, b1 = a * sy; // squareToUniformTriangle, Warp.h:160-164
    V3 p0 = mk(sn.p[0][0], sn.p[0][1], sn.p[0][2]), p1 = mk(sn.p[1][0], sn.p[1][1], sn.p[1][2]), p2 = mk(sn.p[2][0], sn.p[2][1], sn.p[2][2]);
    dRec.p = add(add(mul(p0, b0), mul(p1, b1)), mul(p2, (1.f - b0 - b1)));
    dRec.n = mk(sn.n[0], sn.n[1], sn.n[2]);
    dRec.pdf = 1.0f / L.sum_area;
    V3 dir = sub(dRec.p, dRec.ref);
    float distSquared = dot(dir, dir);
    dRec.dist = sqrtf(distSquared);
    dRec.d = divs(dir, dRec.dist);
    float dp = fabsf(dot(dRec.d, dRec.n));
    dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) return sdivf(sp3(L.radiance), dRec.pdf);
    dRec.pdf = 0.0f;
    return sp(0.0f);
}
float light_pdf_direct(const ctl_light& L, const DRec& dRec) { // Light.cu:137-155, measure ESolidAngle
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) {
        float pdfPos = 1.0f / L.sum_area;
        return pdfPos * (dRec.dist * dRec.dist) / fabsf(dot(dRec.d, dRec.n));
    }
    return 0.0f;
}
float pdf_emitter(const ctl_scene_view& S, unsigned light_buf_idx) { // KernelDynamicScene.cu:42-46
    return S.light_cdf[light_buf_idx] - (light_buf_idx == 0 ? 0.0f : S.light_cdf[light_buf_idx - 1]);
}
inline float power_heuristic(float fPdf, float gPdf) { float f = 1 * fPdf, g = 1 * gPdf; return (f * f) / (f * f + g * g); } // MonteCarlo.h:29-33

// ---- camera (SceneTypes/Sensor.cu:130-144) -----------------------------------
void camera_ray(const ctl_camera& C, float px, float py, V3& o, V3& d) {
    V3 nearP = xf_point(C.sample_to_camera, mk(px * C.inv_resolution[0], py * C.inv_resolution[1], 0.0f));
    V3 dn = normalize(nearP);
    o = mk(C.to_world[3], C.to_world[7], C.to_world[11]); // Translation()
    d = xf_dir(C.to_world, dn);
}

struct PTParams { int max_path_length, rr_start, direct; };
static int g_stop_zero = 1;   // the product's default (StopZeroThroughput=1); 0 reproduces the reference's ray count exactly

inline uint32_t mat_index_of(const ctl_scene_view& S, const Hit& h) { // TraceResult.cu:81-84
    return ((S.tri_data[h.tri].w[1] >> 16) & 0xff) + S.nodes[h.node].material_offset;
}

// PathTrace<DIRECT> (Integrators/PathTracer.cu:10-113), volumes absent, no environment map.
// Deviation (documented in DESIGN.md): a path whose throughput becomes exactly zero stops; the
// reference keeps tracing it with zero weight until Russian roulette ends it (no image effect).
Spec path_trace(const ctl_scene_view& S, V3 ro, V3 rd, Sampler& rnd, const PTParams& P, uint64_t& rays, Counters* cnt) {
    Spec cl = sp(0.0f), cf = sp(1.0f);
    int depth = 0; bool specularBounce = false; float brdf_pdf = 0; V3 last_nor = mk(0, 0, 0);
    BRec bRec; bRec.wo = mk(0, 0, 1);
    while (depth++ < P.max_path_length) {
        Hit r2; r2.dist = FLT_MAX; r2.tri = UINT_MAX; r2.node = UINT_MAX; r2.u = r2.v = 0;
        rays++;
        trace(S, ro, rd, S.ray_eps, 0.0f, r2, cnt, false);
        if (r2.tri == UINT_MAX) break;
        // getBsdfSample (Kernel/TraceResult.cu:16-43)
        const ctl_material& mat = S.materials[mat_index_of(S, r2)];
        bRec.eta = 1.0f; bRec.sampledType = 0; bRec.typeMask = E_ALL;
        bRec.dg.P = add(ro, mul(rd, r2.dist));
        fill_dg(S, r2.u, r2.v, r2.tri, r2.node, bRec.dg);
        bRec.wi = to_local(bRec.dg.sys, neg(rd));
        if ((mat.flags & CTL_MAT_TWO_SIDED) && bRec.wi.z < 0) { bRec.dg.n = neg(bRec.dg.n); bRec.dg.sys.n = neg(bRec.dg.sys.n); bRec.wi.z *= -1.0f; }
        // emitter hit (PathTracer.cu:64-77; TraceResult.cu:45-60)
        if (mat.node_light_index != UINT_MAX) {
            unsigned li = S.nodes[r2.node].lights[mat.node_light_index];
            const ctl_light& L = S.lights[li];
            float misWeight = 1.0f;
            if (!(!P.direct || depth == 1 || specularBounce)) {
                DRec dRec; dRec.ref = ro; dRec.refN = last_nor; dRec.p = bRec.dg.P; dRec.n = bRec.dg.n; dRec.d = rd; dRec.dist = r2.dist;
                float direct_pdf = light_pdf_direct(L, dRec) * pdf_emitter(S, li);
                misWeight = power_heuristic(brdf_pdf, direct_pdf);
            }
            Spec Le = dot(bRec.dg.sys.n, neg(rd)) <= 0 ? sp(0.0f) : sp3(L.radiance); // DiffuseLight::eval, Light.cu:67-70
            cl = sadd(cl, smul(smulf(cf, misWeight), Le));
        }
        float sx, sy; rnd.f2(sx, sy);
        Spec f = bsdf_sample(mat, bRec, brdf_pdf, sx, sy);
        last_nor = bRec.dg.sys.n;
        if (P.direct && (bsdf_combined_type(mat) & E_SMOOTH) && S.num_lights) { // UniformSampleOneLight, TraceAlgorithms.cu:92-101
            float lx, ly; rnd.f2(lx, ly);
            unsigned idx = (unsigned)(upper_bound_f(S.light_cdf, S.light_cdf + S.num_lights, lx) - S.light_cdf); // sampleEmitter, KernelDynamicScene.cu:25-40
            if (idx >= S.num_lights) idx = S.num_lights - 1;
            float fU = S.light_cdf[idx], fL = idx > 0 ? S.light_cdf[idx - 1] : 0.0f;
            float emPdf = fU - fL;
            const ctl_light& L = S.lights[S.light_indices[idx]];
            // EstimateDirect (TraceAlgorithms.cu:44-73)
            DRec dRec; dRec.ref = bRec.dg.P; dRec.refN = bRec.dg.sys.n; dRec.p = bRec.dg.P; dRec.n = bRec.dg.sys.n; dRec.pdf = 0;
            float ex, ey; rnd.f2(ex, ey);
            Spec value = light_sample_direct(S, L, dRec, ex, ey);
            Spec retVal = sp(0.0f);
            if (!sis_zero(value)) {
                BRec b2 = bRec;
                b2.wo = to_local(b2.dg.sys, dRec.d);
                b2.typeMask = E_ALL & ~E_DELTA;
                Spec bsdfVal = bsdf_f(mat, b2);
                if (!sis_zero(bsdfVal)) {
                    Hit sh; sh.dist = FLT_MAX; sh.tri = UINT_MAX; sh.node = UINT_MAX;
                    rays++;
                    trace(S, dRec.ref, dRec.d, S.ray_eps, 0.0f, sh, cnt, false); // Occluded = closest hit, KernelDynamicScene.cu:70-80
                    bool end = sh.dist < dRec.dist - S.ray_eps;
                    bool occluded = sh.dist > 0 + S.ray_eps && end;
                    if (!occluded) {
                        float bsdfPdf = bsdf_pdf(mat, b2);
                        float directPdf = dRec.pdf * emPdf;
                        float weight = power_heuristic(directPdf, bsdfPdf);
                        retVal = smulf(smul(value, bsdfVal), weight);
                    }
                }
            }
            cl = sadd(cl, smul(cf, sdivf(retVal, emPdf)));
        }
        specularBounce = (bRec.sampledType & E_DELTA) != 0;
        cf = smul(cf, f);
        rd = to_world(bRec.dg.sys, bRec.wo); ro = bRec.dg.P;
        if (g_stop_zero && sis_zero(cf)) break; // see deviation note above; orc_set_stop_zero_throughput(0) = the reference's own behaviour
        if (depth > P.rr_start && !specularBounce) {
            if (rnd.f1() >= smax(cf)) break;
            cf = sdivf(cf, smax(cf));
        }
    }
    return cl;
}

// PathTraceRegularization<DIRECT> (Integrators/PathTracer.cu:115-170; KEY_Regularization, off by default in the reference).  A different estimator, not a
// variant of the one above: the ray is traced BEFORE the depth test (a path that survives its last vertex traces one more ray whose hit is never
// shaded), emitted radiance is added only at depth 1 / after a specular bounce / without direct lighting (no MIS weight), every non-delta vertex samples
// ALL lights (UniformSampleAllLights -> EstimateDirect with light_pdf 1, TraceAlgorithms.cu:44-90), a delta vertex draws the two numbers of
// sampleEmitterPosition and adds nothing (the mollified connection is only taken for lights that are neither DiffuseLight nor InfiniteLight, cu:138; this
// path has DiffuseLights only), and Russian roulette applies after specular bounces too.  No environment map, no volumes.
Spec path_trace_regularized(const ctl_scene_view& S, V3 ro, V3 rd, Sampler& rnd, const PTParams& P, uint64_t& rays, Counters* cnt) {
    Spec cl = sp(0.0f), cf = sp(1.0f);
    int depth = 0; bool specularBounce = false;
    BRec bRec; bRec.wo = mk(0, 0, 1);
    for (;;) {
        Hit r2; r2.dist = FLT_MAX; r2.tri = UINT_MAX; r2.node = UINT_MAX; r2.u = r2.v = 0;
        rays++;
        trace(S, ro, rd, S.ray_eps, 0.0f, r2, cnt, false);
        if (r2.tri == UINT_MAX || !(depth++ < P.max_path_length)) break;   // while (traceRay(...) && depth++ < maxPathLength)
        const ctl_material& mat = S.materials[mat_index_of(S, r2)];
        bRec.eta = 1.0f; bRec.sampledType = 0; bRec.typeMask = E_ALL;
        bRec.dg.P = add(ro, mul(rd, r2.dist));
        fill_dg(S, r2.u, r2.v, r2.tri, r2.node, bRec.dg);
        bRec.wi = to_local(bRec.dg.sys, neg(rd));
        if ((mat.flags & CTL_MAT_TWO_SIDED) && bRec.wi.z < 0) { bRec.dg.n = neg(bRec.dg.n); bRec.dg.sys.n = neg(bRec.dg.sys.n); bRec.wi.z *= -1.0f; }
        if ((!P.direct || depth == 1 || specularBounce) && mat.node_light_index != UINT_MAX) {   // cl += cf * r2.Le(...), cu:131-132
            const ctl_light& L = S.lights[S.nodes[r2.node].lights[mat.node_light_index]];
            Spec Le = dot(bRec.dg.sys.n, neg(rd)) <= 0 ? sp(0.0f) : sp3(L.radiance);
            cl = sadd(cl, smul(cf, Le));
        }
        float sx, sy; rnd.f2(sx, sy);
        float unused_pdf = 0;
        Spec f = bsdf_sample(mat, bRec, unused_pdf, sx, sy);
        if (P.direct) {
            if (bsdf_combined_type(mat) & E_DELTA) { float ex, ey; rnd.f2(ex, ey); }   // sampleEmitterPosition's sample; DiffuseLights take no mollified connection
            else {
                Spec Lall = sp(0.0f);
                for (unsigned li = 0; li < S.num_lights; li++) {   // UniformSampleAllLights, nSamples 1
                    const ctl_light& L = S.lights[S.light_indices[li]];
                    DRec dRec; dRec.ref = bRec.dg.P; dRec.refN = bRec.dg.sys.n; dRec.p = bRec.dg.P; dRec.n = bRec.dg.sys.n; dRec.pdf = 0;
                    float ex, ey; rnd.f2(ex, ey);
                    Spec value = light_sample_direct(S, L, dRec, ex, ey);
                    Spec retVal = sp(0.0f);
                    if (!sis_zero(value)) {
                        BRec b2 = bRec;
                        b2.wo = to_local(b2.dg.sys, dRec.d);
                        b2.typeMask = E_ALL & ~E_DELTA;
                        Spec bsdfVal = bsdf_f(mat, b2);
                        if (!sis_zero(bsdfVal)) {
                            Hit sh; sh.dist = FLT_MAX; sh.tri = UINT_MAX; sh.node = UINT_MAX;
                            rays++;
                            trace(S, dRec.ref, dRec.d, S.ray_eps, 0.0f, sh, cnt, false);
                            bool end = sh.dist < dRec.dist - S.ray_eps;
                            bool occluded = sh.dist > 0 + S.ray_eps && end;
                            if (!occluded) {
                                float bsdfPdf = bsdf_pdf(mat, b2);
                                float directPdf = dRec.pdf * 1.0f;
                                retVal = smulf(smul(value, bsdfVal), power_heuristic(directPdf, bsdfPdf));
                            }
                        }
                    }
                    Lall = sadd(Lall, sdivf(retVal, 1.0f));   // L += Ld / float(nSamples)
                }
                cl = sadd(cl, smul(cf, Lall));
            }
        }
        specularBounce = (bRec.sampledType & E_DELTA) != 0;
        cf = smul(cf, f);
        if (depth > P.rr_start) {
            if (rnd.f1() < smax(cf)) cf = sdivf(cf, smax(cf));
            else break;
        }
        rd = to_world(bRec.dg.sys, bRec.wo); ro = bRec.dg.P;
        if (g_stop_zero && sis_zero(cf)) break;   // the product's StopZeroThroughput (see path_trace)
    }
    return cl;
}

struct PixSample { float sx, sy; Spec L; };

void add_sample(ctl_pixel_data* img, int w, int h, float sx, float sy, Spec L) { // Engine/Image.cu:22-44
    L.r = fmaxf(0.0f, L.r); L.g = fmaxf(0.0f, L.g); L.b = fmaxf(0.0f, L.b);
    int x = (int)floorf(sx), y = (int)floorf(sy);
    if (x < 0 || x >= w || y < 0 || y >= h || std::isnan(L.r) || std::isnan(L.g) || std::isnan(L.b) || !std::isfinite(L.r) || !std::isfinite(L.g) || !std::isfinite(L.b)) return;
    ctl_pixel_data& p = img[(size_t)y * w + x];
    p.rgb[0] += L.r; p.rgb[1] += L.g; p.rgb[2] += L.b; p.weight_sum += 1.0f;
}

} // namespace

// =========================================================================== C interface
extern "C" {

void orc_xorwow_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset, uint32_t state[6]) {
    Xorwow st; xorwow_init(seed, subsequence, offset, st);
    for (int i = 0; i < 5; i++) state[i] = st.v[i];
    state[5] = st.d;
}
void orc_xorwow_floats(uint32_t state[6], int n, float* out) {
    Xorwow st; for (int i = 0; i < 5; i++) st.v[i] = state[i]; st.d = state[5];
    for (int i = 0; i < n; i++) out[i] = st.random_float();
    for (int i = 0; i < 5; i++) state[i] = st.v[i]; state[5] = st.d;
}
// sample tables of pass `pass` of a fresh tracer (IndependantSamplingSequenceGenerator, rng(7539414))
void orc_sample_tables(uint32_t pass, float* d1, float* d2) {
    Xorwow st; xorwow_init(1234, 7539414, 0, st);
    for (uint32_t p = 0; p <= pass; p++) fill_tables(st, d1, d2);
}
void orc_sampler_draws(const float* d1, const float* d2, uint32_t idx, int n1, float* out1, int n2, float* out2) {
    Sampler s = {d1, d2, idx, 0, 0};
    for (int i = 0; i < n1; i++) out1[i] = s.f1();
    for (int i = 0; i < n2; i++) s.f2(out2[2 * i], out2[2 * i + 1]);
}
void orc_encode_woop(const float* v0, const float* v1, const float* v2, float* out12) { // TriIntersectorData.cu:5-18
    V3 a = mk(v0[0], v0[1], v0[2]), b = mk(v1[0], v1[1], v1[2]), c = mk(v2[0], v2[1], v2[2]);
    V3 e0 = sub(a, c), e1 = sub(b, c), n = cross(e0, e1);
    float m[16] = {e0.x, e1.x, n.x, c.x, e0.y, e1.y, n.y, c.y, e0.z, e1.z, n.z, c.z, 0, 0, 0, 1}, i[16];
    mat_inverse(m, i);
    float o[12] = {i[8], i[9], i[10], -i[11], i[0], i[1], i[2], i[3], i[4], i[5], i[6], i[7]};
    memcpy(out12, o, 48);
}
int orc_woop_intersect(const float* woop12, const float* o, const float* d, float tmax, float* tuv) { // TriIntersectorData.cu:34-61 (eps 1e-4)
    float t, u, v;
    bool h = woop_test(woop12, mk(o[0], o[1], o[2]), mk(d[0], d[1], d[2]), 0.0001f, tmax, t, u, v);
    if (h) { tuv[0] = t; tuv[1] = u; tuv[2] = v; }
    return h;
}
uint16_t orc_float_to_half(float f) { return f2h(f); }
float orc_half_to_float(uint16_t h) { return h2f(h); }
uint16_t orc_encode_normal(const float* n) { return enc_normal(mk(n[0], n[1], n[2])); }
void orc_decode_normal(uint16_t c, float* out) { V3 n = dec_normal(c); out[0] = n.x; out[1] = n.y; out[2] = n.z; }
void orc_encode_tri_data(const float* p9, const float* n9, const float* uv6, uint32_t mat, uint32_t* out8) { // TriangleData.cu:8-65
    uint16_t h[6]; for (int i = 0; i < 6; i++) h[i] = f2h(uv6[i]);
    out8[5] = h[0] | ((uint32_t)h[1] << 16); out8[6] = h[2] | ((uint32_t)h[3] << 16); out8[7] = h[4] | ((uint32_t)h[5] << 16);
    V3 v0 = mk(p9[0], p9[1], p9[2]), v1 = mk(p9[3], p9[4], p9[5]), v2 = mk(p9[6], p9[7], p9[8]);
    float t0x = h2f(h[0]), t0y = h2f(h[1]), t1x = h2f(h[2]), t1y = h2f(h[3]), t2x = h2f(h[4]), t2y = h2f(h[5]);
    V3 dP1 = sub(v1, v0), dP2 = sub(v2, v0);
    float dU1x = t1x - t0x, dU1y = t1y - t0y, dU2x = t2x - t0x, dU2y = t2y - t0y;
    float det = dU1x * dU2y - dU1y * dU2x;
    V3 dpdu, dpdv;
    if (det == 0) { V3 n = normalize(cross(dP1, dP2)); coordinate_system(n, dpdu, dpdv); }
    else { float id = 1.0f / det; dpdu = mul(sub(mul(dP1, dU2y), mul(dP2, dU1y)), id); dpdv = mul(add(mul(dP1, -dU2x), mul(dP2, dU1x)), id); }
    out8[0] = enc_normal(mk(n9[0], n9[1], n9[2])) | ((uint32_t)enc_normal(mk(n9[3], n9[4], n9[5])) << 16);
    out8[1] = enc_normal(mk(n9[6], n9[7], n9[8])) | ((mat & 0xff) << 16);
    out8[2] = f2h(dpdu.x) | ((uint32_t)f2h(dpdu.y) << 16);
    out8[3] = f2h(dpdu.z) | ((uint32_t)f2h(dpdv.x) << 16);
    out8[4] = f2h(dpdv.y) | ((uint32_t)f2h(dpdv.z) << 16);
}
void orc_warp(int which, float sx, float sy, float* out3) {
    if (which == 0) { V3 v = cosine_hemisphere(sx, sy); out3[0] = v.x; out3[1] = v.y; out3[2] = v.z; }
    else { float a = sqrtf(1.0f - sx); out3[0] = 1 - a; out3[1] = a * sy; out3[2] = 0; }
}
float orc_fresnel_dielectric_ext(float cosi, float eta, float* cost) { return fresnel_dielectric_ext(cosi, *cost, eta); }
void orc_fresnel_conductor_exact(float cosi, const float* eta, const float* k, float* out) { Spec s = fresnel_conductor_exact(cosi, sp3(eta), sp3(k)); out[0] = s.r; out[1] = s.g; out[2] = s.b; }
// microfacet probe: m = sample(wi, s), pdf, D(m), G1(wi, m)
void orc_microfacet_sample(int type, float alpha, const float* wi, float sx, float sy, float* out6) {
    Distr d(type, alpha, alpha); V3 w = mk(wi[0], wi[1], wi[2]);
    V3 m = d.sample_visible(w, sx, sy);
    out6[0] = m.x; out6[1] = m.y; out6[2] = m.z; out6[3] = d.pdf_visible(w, m); out6[4] = d.eval(m); out6[5] = d.smith_g1(w, m);
}
// BSDF probe with an identity shading frame: out = weight[3], pdf, wo[3], sampledType, eta ; f[3], pdf(wo)
void orc_bsdf_probe(const ctl_material* m, const float* wi, float sx, float sy, float* out9, float* f3, float* pdf1) {
    BRec b; memset(&b, 0, sizeof(b)); b.wi = mk(wi[0], wi[1], wi[2]); b.wo = mk(0, 0, 1); b.typeMask = E_ALL; b.eta = 1;
    float pdf = 0; Spec w = bsdf_sample(*m, b, pdf, sx, sy);
    out9[0] = w.r; out9[1] = w.g; out9[2] = w.b; out9[3] = pdf; out9[4] = b.wo.x; out9[5] = b.wo.y; out9[6] = b.wo.z; out9[7] = (float)b.sampledType; out9[8] = b.eta;
    BRec b2 = b; b2.typeMask = E_ALL & ~E_DELTA;
    Spec f = bsdf_f(*m, b2); f3[0] = f.r; f3[1] = f.g; f3[2] = f.b; *pdf1 = bsdf_pdf(*m, b2);
}
void orc_bsdf_eval(const ctl_material* m, const float* wi, const float* wo, float* f3, float* pdf1) {
    BRec b; memset(&b, 0, sizeof(b)); b.wi = mk(wi[0], wi[1], wi[2]); b.wo = mk(wo[0], wo[1], wo[2]); b.typeMask = E_ALL & ~E_DELTA; b.eta = 1;
    Spec f = bsdf_f(*m, b); f3[0] = f.r; f3[1] = f.g; f3[2] = f.b; *pdf1 = bsdf_pdf(*m, b);
}
// DiffuseLight::sampleDirect probe: out = value[3], pdf, p[3], d[3], dist
void orc_light_sample_direct(const ctl_scene_view* S, uint32_t light, const float* ref, const float* refN, float sx, float sy, float* out11) {
    DRec d; d.ref = mk(ref[0], ref[1], ref[2]); d.refN = mk(refN[0], refN[1], refN[2]); d.pdf = 0;
    Spec v = light_sample_direct(*S, S->lights[light], d, sx, sy);
    float o[11] = {v.r, v.g, v.b, d.pdf, d.p.x, d.p.y, d.p.z, d.d.x, d.d.y, d.d.z, d.dist};
    memcpy(out11, o, sizeof(o));
}
void orc_fill_dg(const ctl_scene_view* S, float u, float v, uint32_t tri, uint32_t node, float* out12) { // sys.s, sys.t, sys.n, n
    DG dg; fill_dg(*S, u, v, tri, node, dg);
    float o[12] = {dg.sys.s.x, dg.sys.s.y, dg.sys.s.z, dg.sys.t.x, dg.sys.t.y, dg.sys.t.z, dg.sys.n.x, dg.sys.n.y, dg.sys.n.z, dg.n.x, dg.n.y, dg.n.z};
    memcpy(out12, o, sizeof(o));
}
void orc_camera_ray(const ctl_scene_view* S, float px, float py, float* o3, float* d3) {
    V3 o, d; camera_ray(S->camera, px, py, o, d); o3[0] = o.x; o3[1] = o.y; o3[2] = o.z; d3[0] = d.x; d3[1] = d.y; d3[2] = d.z;
}

// traceRay batched (Kernel/TraceHelper.cu:174-180): ray tmin/tmax ignored; counts may be NULL
void orc_trace_rays(const ctl_scene_view* S, int n, const ctl_traversal_ray* rays, ctl_trace_result* res, uint64_t* counts) {
    Counters c = {0, 0, 0};
    for (int i = 0; i < n; i++) {
        Hit h; h.dist = FLT_MAX; h.tri = UINT_MAX; h.node = UINT_MAX; h.u = h.v = 0;
        trace(*S, mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]), S->ray_eps, 0.0f, h, counts ? &c : 0, false);
        res[i].dist = h.dist; res[i].u = h.u; res[i].v = h.v; res[i].tri_idx = h.tri; res[i].node_idx = h.node;
    }
    if (counts) { counts[0] = c.inner; counts[1] = c.tris; counts[2] = c.inst; }
}
// intersectKernel<ANY_HIT> semantics (Kernel/TraceHelper.cu:326-734): tmin/tmax from the ray, 16-byte packed result
void orc_intersect(const ctl_scene_view* S, int n, const ctl_traversal_ray* rays, ctl_traversal_result* res, int any_hit) {
    for (int i = 0; i < n; i++) {
        Hit h; h.dist = rays[i].tmax; h.tri = UINT_MAX; h.node = UINT_MAX; h.u = h.v = 0;
        trace(*S, mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmin, rays[i].tmin, h, 0, any_hit != 0);
        res[i].dist = h.dist; res[i].node_idx = -1; res[i].tri_idx = -1; res[i].bary = 0;
        if (h.tri != UINT_MAX) {
            res[i].node_idx = (int)h.node; res[i].tri_idx = (int)h.tri;
            uint16_t xd = (uint16_t)(h.u * 65535), yd = (uint16_t)(h.v * 65535); // :726-727
            res[i].bary = ((uint32_t)yd << 16) | (uint32_t)xd;
        }
    }
}

// Render passes [pass_first, pass_first + n_passes) of a fresh tracer over the pixel window into `img`
// (accumulated, not cleared): the loop of pathKernel2's body (PathTracer.cu:184-193) over pixels.
// Rows are distributed over n_threads; samples are splatted serially in pixel order so the result is
// independent of the thread count. rays_out: traceRay calls (extension + shadow).
// 1 (default) = the product's StopZeroThroughput=1; 0 = zero-throughput paths live until Russian roulette: the reference's own ray count (pinned against oracle/_ref)
void orc_set_stop_zero_throughput(int on) { g_stop_zero = on != 0; }
void orc_render(const ctl_scene_view* S, int w, int h, int x0, int y0, int x1, int y1, int pass_first, int n_passes,
                int max_path_length, int rr_start, int direct, ctl_pixel_data* img, uint64_t* rays_out, int n_threads, uint64_t* counts) {
    std::vector<float> d1((size_t)N_SEQ * SEQ_LEN), d2((size_t)N_SEQ * SEQ_LEN * 2);
    Xorwow st; xorwow_init(1234, 7539414, 0, st);
    for (int p = 0; p < pass_first; p++) fill_tables(st, d1.data(), d2.data());
    const bool regularization = (direct & 2) != 0;   // bit 1 of `direct`: PathTraceRegularization (KEY_Regularization)
    PTParams P = {max_path_length, rr_start, direct & 1};
    uint64_t total_rays = 0; Counters total_cnt = {0, 0, 0};
    if (n_threads < 1) n_threads = 1;
    int bw = x1 - x0, bh = y1 - y0;
    std::vector<PixSample> samples((size_t)bw * bh);
    for (int p = 0; p < n_passes; p++) {
        fill_tables(st, d1.data(), d2.data());
        std::atomic<int> next_row(0);
        std::vector<uint64_t> trays(n_threads, 0); std::vector<Counters> tcnt(n_threads, Counters{0, 0, 0});
        auto work = [&](int tid) {
            for (;;) {
                int ry = next_row.fetch_add(1); if (ry >= bh) break;
                int y = y0 + ry;
                for (int x = x0; x < x1; x++) {
                    Sampler rng = {d1.data(), d2.data(), (unsigned)(y * w + x), 0, 0};
                    float jx, jy; rng.f2(jx, jy);
                    float pX = (float)x + jx, pY = (float)y + jy;
                    float ax, ay; rng.f2(ax, ay); // aperture sample, unused by the pinhole
                    V3 o, d; camera_ray(S->camera, pX, pY, o, d);
                    Spec col = regularization ? path_trace_regularized(*S, o, d, rng, P, trays[tid], counts ? &tcnt[tid] : 0) : path_trace(*S, o, d, rng, P, trays[tid], counts ? &tcnt[tid] : 0);
                    PixSample ps = {pX, pY, col};
                    samples[(size_t)ry * bw + (x - x0)] = ps;
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
        for (auto& s : samples) add_sample(img, w, h, s.sx, s.sy, s.L);
        for (int t = 0; t < n_threads; t++) { total_rays += trays[t]; total_cnt.inner += tcnt[t].inner; total_cnt.tris += tcnt[t].tris; total_cnt.inst += tcnt[t].inst; }
    }
    if (rays_out) *rays_out = total_rays;
    if (counts) { counts[0] = total_cnt.inner; counts[1] = total_cnt.tris; counts[2] = total_cnt.inst; }
}

// WavefrontPathTracer::DoRender (Integrators/PseudoRealtime/WavefrontPathTracer.cu:17-191) over DoubleRayBuffer (Kernel/DoubleRayBuffer.h),
// restated in the SERIAL schedule of its queue atomics (fetch index i = 0, 1, 2, ...; the k-th insertion lands in slot k): the reference's GPU
// order depends on the hardware scheduler, and the random numbers of a path are a function of its queue slot (cu:59-60), so the serial order is
// the one deterministic member of the reference's possible outputs.  It is also what oracle/_ref executes (the reference's own kernel text on one
// host thread).  Quirks kept: the sampler is keyed by the queue slot and re-skipped to dimension passesDone + 2 at EVERY bounce
// (cu:59-60); Russian roulette before the BSDF sample, from pathDepth >= RRStartDepth (cu:102-109); one 2-D sample re-used for light selection
// and light position (KernelDynamicScene.cu:98-117); shadow rays are closest-hit queries compared with dDist * (1 - eps) (cu:68); hits arrive
// through the 16-byte traversalResult, i.e. with 16-bit barycentrics (TraceHelper.cu:44-60); the previous normal is the 16-bit spherical
// code (cu:94,137); the sample is splatted at the un-jittered half-precision pixel coordinate (cu:40-41,161).
// queue_sizes (may be NULL): [2*i] primary, [2*i+1] secondary rays intersected before iteration i of the last pass.
// visit_counts (may be NULL; summed over all passes): [0..2] inner nodes / triangle tests / instance leaves of the primary queries, [3] primary rays,
// [4..7] the same for the secondary rays traced the way the CUDA path traces them (any hit against tmax = dDist (1 - eps): the same occlusion
// predicate as the reference's closest-hit test, see csrc/wavefront_pt.cuh) -- the roofline's algorithmic bytes for this integrator.
void orc_render_wavefront_counted(const ctl_scene_view* Sp, int w, int h, int pass_first, int n_passes, int max_path_length, int rr_start, int direct,
                                  ctl_pixel_data* img, uint64_t* rays_out, uint32_t* queue_sizes, uint64_t* visit_counts);
void orc_render_wavefront(const ctl_scene_view* Sp, int w, int h, int pass_first, int n_passes, int max_path_length, int rr_start, int direct,
                          ctl_pixel_data* img, uint64_t* rays_out, uint32_t* queue_sizes) {
    orc_render_wavefront_counted(Sp, w, h, pass_first, n_passes, max_path_length, rr_start, direct, img, rays_out, queue_sizes, nullptr);
}
void orc_render_wavefront_counted(const ctl_scene_view* Sp, int w, int h, int pass_first, int n_passes, int max_path_length, int rr_start, int direct,
                                  ctl_pixel_data* img, uint64_t* rays_out, uint32_t* queue_sizes, uint64_t* visit_counts) {
    const ctl_scene_view& S = *Sp;
    struct Payload { Spec throughput; uint16_t x, y; Spec L, directF; float dDist; uint32_t dIdx; bool specular_bounce; float bsdf_pdf; uint32_t prev_normal; }; // WavefrontPathTracer.h:11-22
    const size_t N = (size_t)w * h;
    std::vector<Payload> pay(N); std::vector<ctl_traversal_ray> ray(N), sec_in(N), sec_out(N); std::vector<ctl_traversal_result> res(N), sec_res(N);
    std::vector<float> d1((size_t)N_SEQ * SEQ_LEN), d2((size_t)N_SEQ * SEQ_LEN * 2), sec_tmax(N);
    Xorwow st; xorwow_init(1234, 7539414, 0, st);
    for (int p = 0; p < pass_first; p++) fill_tables(st, d1.data(), d2.data());
    uint64_t rays = 0;
    auto mk_ray = [&](V3 o, V3 d) { ctl_traversal_ray r; r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.tmin = S.ray_eps; r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z; r.tmax = FLT_MAX; return r; }; // DoubleRayBuffer.h:234-237
    for (int p = 0; p < n_passes; p++) {
        fill_tables(st, d1.data(), d2.data());
        const unsigned iterationIdx = (unsigned)(pass_first + p + 1); // m_uPassesDone++ precedes DoRender (Kernel/Tracer.h:231-232)
        // pathCreateKernelWPT (cu:17-49), one sample per pixel
        uint32_t n_pay = 0, n_sec = 0;
        for (uint32_t rayidx = 0; rayidx < (uint32_t)N; rayidx++) {
            int x = (int)(rayidx % (uint32_t)w), y = (int)(rayidx / (uint32_t)w);
            Sampler rng = {d1.data(), d2.data(), rayidx, 0, 0};
            float jx, jy; rng.f2(jx, jy); float ax, ay; rng.f2(ax, ay);
            V3 o, d; camera_ray(S.camera, (float)x + jx, (float)y + jy, o, d);
            Payload dat; dat.x = f2h((float)x); dat.y = f2h((float)y); dat.throughput = sp(1.0f); dat.L = sp(0.0f); dat.directF = sp(0.0f); dat.dDist = 0; dat.dIdx = UINT_MAX;
            dat.specular_bounce = true; dat.bsdf_pdf = 0; dat.prev_normal = 0;
            pay[n_pay] = dat; ray[n_pay] = mk_ray(o, d); n_pay++;
        }
        int pathDepth = 0;
        do {
            // FinishIteration (DoubleRayBuffer.h:84-112): intersect the primaries and the new secondaries, swap the secondary buffers
            orc_intersect(Sp, (int)n_pay, ray.data(), res.data(), 0);
            if (n_sec) orc_intersect(Sp, (int)n_sec, sec_out.data(), sec_res.data(), 0);
            if (visit_counts) { // counting replay of the same queries (primaries as they are; secondaries in the CUDA path's any-hit form)
                Counters cp = {0, 0, 0}, cs = {0, 0, 0};
                for (uint32_t i = 0; i < n_pay; i++) { Hit hh; hh.dist = ray[i].tmax; hh.u = hh.v = 0; hh.tri = hh.node = UINT_MAX;
                    trace(S, mk(ray[i].o[0], ray[i].o[1], ray[i].o[2]), mk(ray[i].d[0], ray[i].d[1], ray[i].d[2]), ray[i].tmin, ray[i].tmin, hh, &cp, false); }
                for (uint32_t i = 0; i < n_sec; i++) { Hit hh; hh.dist = sec_tmax[i]; hh.u = hh.v = 0; hh.tri = hh.node = UINT_MAX;
                    trace(S, mk(sec_out[i].o[0], sec_out[i].o[1], sec_out[i].o[2]), mk(sec_out[i].d[0], sec_out[i].d[1], sec_out[i].d[2]), sec_out[i].tmin, sec_out[i].tmin, hh, &cs, true); }
                visit_counts[0] += cp.inner; visit_counts[1] += cp.tris; visit_counts[2] += cp.inst; visit_counts[3] += n_pay;
                visit_counts[4] += cs.inner; visit_counts[5] += cs.tris; visit_counts[6] += cs.inst; visit_counts[7] += n_sec;
            }
            rays += (uint64_t)n_pay + n_sec;
            if (queue_sizes && p == n_passes - 1) { queue_sizes[2 * pathDepth] = n_pay; queue_sizes[2 * pathDepth + 1] = n_sec; }
            const uint32_t n_fetch = n_pay; n_pay = 0; n_sec = 0;
            sec_in.swap(sec_out); // sec_in + sec_res: what accessSecondaryRay reads; sec_out: what insertSecondaryRay fills
            // pathIterateKernel<NEXT_EVENT_EST> (cu:51-164)
            for (uint32_t rayIdx = 0; rayIdx < n_fetch; rayIdx++) {
                Payload payload = pay[rayIdx];
                const V3 ro = mk(ray[rayIdx].o[0], ray[rayIdx].o[1], ray[rayIdx].o[2]), rd = mk(ray[rayIdx].d[0], ray[rayIdx].d[1], ray[rayIdx].d[2]);
                Hit r2; r2.dist = res[rayIdx].dist; r2.node = (uint32_t)res[rayIdx].node_idx; r2.tri = (uint32_t)res[rayIdx].tri_idx; // toResult, TraceHelper.cu:44-51
                { uint16_t xd = (uint16_t)(res[rayIdx].bary & 0xffff), yd = (uint16_t)(res[rayIdx].bary >> 16); r2.u = (float)xd / 65535.0f; r2.v = (float)yd / 65535.0f; }
                Sampler rng = {d1.data(), d2.data(), rayIdx, iterationIdx + 2, iterationIdx + 2};
                if (direct && pathDepth > 0 && payload.dIdx != UINT_MAX) {
                    if (sec_res[payload.dIdx].dist >= payload.dDist * (1 - S.ray_eps)) payload.L = sadd(payload.L, payload.directF);
                    payload.dIdx = UINT_MAX; payload.directF = sp(0.0f);
                }
                bool path_terminated = (pathDepth + 1 == max_path_length);
                if (r2.tri != UINT_MAX) {
                    const ctl_material& mat = S.materials[mat_index_of(S, r2)];
                    BRec bRec; bRec.wo = mk(0, 0, 0); // unset in the reference; a failed sample still launches a ray along it: defined as zero (oracle/build_ref.sh note 9)
                     bRec.eta = 1.0f; bRec.sampledType = 0; bRec.typeMask = E_ALL;
                    bRec.dg.P = add(ro, mul(rd, r2.dist));
                    fill_dg(S, r2.u, r2.v, r2.tri, r2.node, bRec.dg);
                    bRec.wi = to_local(bRec.dg.sys, neg(rd));
                    if ((mat.flags & CTL_MAT_TWO_SIDED) && bRec.wi.z < 0) { bRec.dg.n = neg(bRec.dg.n); bRec.dg.sys.n = neg(bRec.dg.sys.n); bRec.wi.z *= -1.0f; }
                    if (mat.node_light_index != UINT_MAX) { // cu:84-99
                        unsigned li = S.nodes[r2.node].lights[mat.node_light_index];
                        const ctl_light& L = S.lights[li];
                        float misWeight = 1.0f;
                        if (!(!direct || pathDepth == 0 || payload.specular_bounce)) {
                            DRec dRec; dRec.ref = ro; dRec.refN = dec_normal((uint16_t)payload.prev_normal); dRec.p = bRec.dg.P; dRec.n = bRec.dg.n; dRec.d = rd; dRec.dist = r2.dist;
                            float direct_pdf = light_pdf_direct(L, dRec) * pdf_emitter(S, li);
                            misWeight = power_heuristic(payload.bsdf_pdf, direct_pdf);
                        }
                        Spec Le = dot(bRec.dg.sys.n, neg(rd)) <= 0 ? sp(0.0f) : sp3(L.radiance);
                        payload.L = sadd(payload.L, smul(smulf(Le, misWeight), payload.throughput)); // misWeight * Le * throughput
                    }
                    bool surviveRR = true;
                    if (pathDepth >= rr_start) { // cu:102-109
                        if (rng.f1() < smax(payload.throughput)) payload.throughput = sdivf(payload.throughput, smax(payload.throughput));
                        else surviveRR = false;
                    }
                    if (pathDepth + 1 != max_path_length && surviveRR) {
                        float sx, sy; rng.f2(sx, sy);
                        Spec f = bsdf_sample(mat, bRec, payload.bsdf_pdf, sx, sy);
                        payload.specular_bounce = (bRec.sampledType & E_DELTA) != 0;
                        const V3 refl_d = to_world(bRec.dg.sys, bRec.wo);
                        payload.dIdx = UINT_MAX;
                        if (direct && (bsdf_combined_type(mat) & E_SMOOTH)) { // cu:118-135
                            DRec dRec; dRec.ref = bRec.dg.P; dRec.refN = bRec.dg.sys.n; dRec.p = bRec.dg.P; dRec.n = bRec.dg.sys.n; dRec.pdf = 0;
                            float lx, ly; rng.f2(lx, ly);
                            Spec value = sp(0.0f);
                            if (S.num_lights) { // sampleEmitterDirect / sampleEmitter with sample re-use (KernelDynamicScene.cu:25-40, 98-117)
                                unsigned idx = (unsigned)(upper_bound_f(S.light_cdf, S.light_cdf + S.num_lights, lx) - S.light_cdf);
                                if (idx >= S.num_lights) idx = S.num_lights - 1;
                                float fU = S.light_cdf[idx], fL = idx > 0 ? S.light_cdf[idx - 1] : 0.0f;
                                lx = (lx - fL) / (fU - fL);
                                float emPdf = fU - fL;
                                value = light_sample_direct(S, S.lights[S.light_indices[idx]], dRec, lx, ly);
                                if (dRec.pdf != 0) { dRec.pdf *= emPdf; value = sdivf(value, emPdf); } else value = sp(0.0f);
                            }
                            if (!sis_zero(value)) {
                                bRec.typeMask = E_ALL & ~E_DELTA;
                                bRec.wo = to_local(bRec.dg.sys, dRec.d);
                                Spec bsdfVal = bsdf_f(mat, bRec);
                                const float bsdfPdf = bsdf_pdf(mat, bRec);
                                const float directPdf = dRec.pdf; // DiffuseLight::sampleDirect reports ESolidAngle (Light.cu:84-135)
                                const float weight = power_heuristic(directPdf, bsdfPdf);
                                payload.directF = smulf(smul(smul(payload.throughput, value), bsdfVal), weight);
                                payload.dDist = dRec.dist;
                                if (n_sec < (uint32_t)N) { payload.dIdx = n_sec; sec_out[n_sec] = mk_ray(bRec.dg.P, dRec.d); sec_tmax[n_sec] = payload.dDist * (1 - S.ray_eps); n_sec++; } // insertSecondaryRay, DoubleRayBuffer.h:166-177
                            }
                        }
                        payload.prev_normal = enc_normal(bRec.dg.sys.n);
                        payload.throughput = smul(payload.throughput, f);
                        pay[n_pay] = payload; ray[n_pay] = mk_ray(bRec.dg.P, refl_d); n_pay++;
                    } else path_terminated = true;
                } else {
                    path_terminated = true; // no environment map: the miss adds misWeight * throughput * 0 (cu:143-156)
                    payload.L = sadd(payload.L, smulf(smul(payload.throughput, sp(0.0f)), 1.0f));
                }
                if (path_terminated) add_sample(img, w, h, h2f(payload.x), h2f(payload.y), payload.L);
            }
        } while (n_pay != 0 && ++pathDepth < max_path_length);
    }
    if (rays_out) *rays_out = rays;
}

// single path probe (Tracer::Debug / PathTracer::DebugInternal analogue): radiance of pixel (x,y) in pass `pass`
void orc_path_probe(const ctl_scene_view* S, int w, int x, int y, int pass, int max_path_length, int rr_start, int direct, float* rgb, uint64_t* rays) {
    std::vector<float> d1((size_t)N_SEQ * SEQ_LEN), d2((size_t)N_SEQ * SEQ_LEN * 2);
    Xorwow st; xorwow_init(1234, 7539414, 0, st);
    for (int p = 0; p <= pass; p++) fill_tables(st, d1.data(), d2.data());
    Sampler rng = {d1.data(), d2.data(), (unsigned)(y * w + x), 0, 0};
    float jx, jy; rng.f2(jx, jy); float ax, ay; rng.f2(ax, ay);
    V3 o, d; camera_ray(S->camera, (float)x + jx, (float)y + jy, o, d);
    PTParams P = {max_path_length, rr_start, direct};
    uint64_t r = 0; Spec c = path_trace(*S, o, d, rng, P, r, 0);
    rgb[0] = c.r; rgb[1] = c.g; rgb[2] = c.b; if (rays) *rays = r;
}

// Per-ray event strings ('N' inner node popped, 'I' instance leaf entered, 'T' triangle reference tested), in visit order,
// concatenated; offsets[i]..offsets[i+1] delimit ray i. Returns the total length (call with events == NULL to size).
// Used by scripts/sim_warp_schedule.py to study warp scheduling policies for the traversal kernel.
long long orc_trace_events(const ctl_scene_view* S, int n, const ctl_traversal_ray* rays, int any_hit, uint8_t* events, long long capacity, long long* offsets) {
    std::vector<uint8_t> ev; long long total = 0;
    for (int i = 0; i < n; i++) {
        ev.clear();
        Counters c = {0, 0, 0, &ev};
        Hit h; h.dist = rays[i].tmax; h.u = h.v = 0; h.tri = h.node = UINT_MAX;
        trace(*S, mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmin, 0.0f, h, &c, any_hit != 0);
        if (offsets) offsets[i] = total;
        if (events && total + (long long)ev.size() <= capacity) memcpy(events + total, ev.data(), ev.size());
        total += (long long)ev.size();
    }
    if (offsets) offsets[n] = total;
    return total;
}

// ---- image pipeline: applyImagePipeline (Kernel/ImagePipeline/ImagePipeline.cu:54-84) ---------------------------
// PixelData::toSpectrum (Engine/Image.h:20-27; Spectrum / float multiplies by the reciprocal, Math/Spectrum.h:122-128)
static inline void px_to_spectrum(const ctl_pixel_data& P, float splat_scale, float c[3]) {
    const float weight = P.weight_sum != 0 ? P.weight_sum : 1, recip = 1.0f / weight;
    for (int k = 0; k < 3; k++) c[k] = P.rgb[k] * recip + P.rgb_splat[k] * splat_scale;
}
// gammaCorrecture (ImagePipeline.cu:7-12): toSRGBComponent (Math/Spectrum.cu:229-235) -> Float3ToCOLORREF (Math/Spectrum.h:521-526)
static inline void gamma_to_rgba8(const float c[3], uint8_t* out) {
    for (int k = 0; k < 3; k++) {
        float v = c[k], s2 = v <= (float)0.0031308 ? (float)12.92 * v : (float)1.055 * powf(v, (float)(1.0 / 2.4)) - (float)0.055;
        float cl = s2 < 0.0f ? 0.0f : (s2 > 1.0f ? 1.0f : s2); // math::clamp01
        out[k] = (unsigned char)(cl * 255.0f);
    }
    out[3] = 255;
}
// Spectrum::toRGBE / fromRGBE (Math/Spectrum.h:534-565)
static inline void to_rgbe(const float c[3], uint8_t e4[4]) {
    float mx = std::max(c[0], std::max(c[1], c[2]));
    if (mx < 1e-32) { e4[0] = e4[1] = e4[2] = e4[3] = 0; return; }
    int e; float scale = (float)frexp((double)mx, &e) * 256.0f / mx;
    for (int k = 0; k < 3; k++) { const float v = c[k] * scale; e4[k] = (unsigned char)(unsigned int)(v > 0.0f ? v : 0.0f); } // device conversion (cvt.rzi.u32.f32): negative lobes -> 0
    e4[3] = (unsigned char)(e + 128);
}
static inline void from_rgbe(const uint8_t e4[4], float c[3]) {
    if (!e4[3]) { c[0] = c[1] = c[2] = 0; return; }
    float ex = ldexpf(1.0f, int(e4[3]) - (128 + 8));
    for (int k = 0; k < 3; k++) c[k] = e4[k] * ex;
}
// the five reconstruction filters of SceneTypes/Filter.h (Box :28-48, Gaussian :50-82, Mitchell :84-116, LanczosSinc :118-148, Triangle :151-171)
struct FilterEval {
    int type; float xw, yw, p0, p1, ix, iy, expx, expy;
    FilterEval(const ctl_image_pipeline& P) : type(P.filter_type), xw(P.x_width), yw(P.y_width), p0(P.param0), p1(P.param1), ix(1.f / P.x_width), iy(1.f / P.y_width),
        expx(expf(-P.param0 * P.x_width * P.x_width)), expy(expf(-P.param0 * P.y_width * P.y_width)) {}
    float mitchell1d(float x) const {
        const float B = p0, C = p1;
        x = fabsf(2.f * x);
        if (x > 1.f) return ((-B - 6 * C) * x * x * x + (6 * B + 30 * C) * x * x + (-12 * B - 48 * C) * x + (8 * B + 24 * C)) * (1.f / 6.f);
        return ((12 - 9 * B - 6 * C) * x * x * x + (-18 + 12 * B + 6 * C) * x * x + (6 - 2 * B)) * (1.f / 6.f);
    }
    float sinc1d(float x) const {
        x = fabsf(x);
        if (x < 1e-5) return 1.f;
        if (x > 1.) return 0.f;
        x *= PI_F;
        float sinc = sinf(x) / x, lanczos = sinf(x * p0) / (x * p0);
        return sinc * lanczos;
    }
    float operator()(float x, float y) const {
        switch (type) {
        case 0: return 1.0f;
        case 1: return fmaxf(0.f, float(expf(-p0 * x * x) - expx)) * fmaxf(0.f, float(expf(-p0 * y * y) - expy));
        case 2: return fmaxf(0.f, xw - fabsf(x)) * fmaxf(0.f, yw - fabsf(y));
        case 3: return mitchell1d(x * ix) * mitchell1d(y * iy);
        default: return sinc1d(x * ix) * sinc1d(y * iy);
        }
    }
};
// evalFilter (Kernel/ImagePipeline/Filter/CanonicalFilter.cu:6-27)
static void eval_filter(const FilterEval& F, const ctl_pixel_data* img, float splat_scale, int _x, int _y, int w, int h, float c[3]) {
    int x0 = std::max(0, (int)ceilf(_x - F.xw)), x1 = std::min(w - 1, (int)floorf(_x + F.xw));
    int y0 = std::max(0, (int)ceilf(_y - F.yw)), y1 = std::min(h - 1, (int)floorf(_y + F.yw));
    c[0] = c[1] = c[2] = 0;
    if ((x1 - x0) < 0 || (y1 - y0) < 0) return;
    float acc[3] = {0, 0, 0}, accw = 0;
    for (int y = y0; y <= y1; ++y) for (int x = x0; x <= x1; ++x) {
        float wt = F((float)abs(x - _x), (float)abs(y - _y));
        float pc[3]; px_to_spectrum(img[y * w + x], splat_scale, pc);
        for (int k = 0; k < 3; k++) acc[k] += pc[k] * wt;
        accw += wt;
    }
    const float recip = 1.0f / accw; // Spectrum / float
    for (int k = 0; k < 3; k++) c[k] = acc[k] * recip;
}
// Image::ComputeLuminanceInfo (Engine/Image.cu:88-173) over the RGBE stage.  The reference sums with float atomics (16x16 blocks into shared
// memory, then into globals): the order is the scheduler's.  Restated in the deterministic order block by block, row-major inside a block.
// lum: [0] min, [1] max, [2] avg, [3] exp(avg log(2.3e-5 + Y))
static void luminance_info(const uint8_t* rgbe, int w, int h, float lum[4]) {
    float mn = FLT_MAX, mx = 0.0f, sum = 0.0f, sumlog = 0.0f;
    for (int by = 0; by < h; by += 16) for (int bx = 0; bx < w; bx += 16) {
        float s = 0.0f, sl = 0.0f;
        for (int y = by; y < std::min(h, by + 16); y++) for (int x = bx; x < std::min(w, bx + 16); x++) {
            float c[3]; from_rgbe(rgbe + 4 * ((size_t)y * w + x), c);
            float Y = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f; // Spectrum::getLuminance, Math/Spectrum.cu:174-177
            mn = std::min(mn, Y); mx = std::max(mx, Y);
            s += Y; sl += logf(2.3e-5f + Y);
        }
        sum += s; sumlog += sl;
    }
    lum[0] = mn; lum[1] = mx; lum[2] = sum / (float)(w * h); lum[3] = expf(sumlog / (float)(w * h));
}
// Reinhard05Kernel (Kernel/ImagePipeline/PostProcess/ToneMapPostProcess.cu:6-25) with toYxy / fromYxy (Math/Spectrum.cu:286-302), then
// applyGammaCorrectureToOutput (ImagePipeline.cu:43-52): the processed RGBA8 value is read back, gamma-corrected and quantised again
static void reinhard_pixel(const uint8_t e4[4], float scale, float invWp2, uint8_t* out) {
    float c[3]; from_rgbe(e4, c);
    float X = c[0] * 0.412453f + c[1] * 0.357580f + c[2] * 0.180423f, Y = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f, Z = c[0] * 0.019334f + c[1] * 0.119193f + c[2] * 0.950227f;
    float sxyz = X + Y + Z; sxyz = sxyz < 0.001f ? 0.001f : (sxyz > 100000.0f ? 100000.0f : sxyz);
    float x = X / sxyz, y = Y / sxyz;
    float Lp = scale * Y;
    Y = Lp * (1.0f + Lp * invWp2) / (1.0f + Lp);
    float yc = y < 0.001f ? 0.001f : (y > 100000.0f ? 100000.0f : y);
    X = Y / yc * x; Z = Y / yc * (1 - x - y);
    float r[3] = {3.240479f * X + -1.537150f * Y + -0.498535f * Z, -0.969256f * X + 1.875991f * Y + 0.041556f * Z, 0.055648f * X + -0.204043f * Y + 1.057311f * Z};
    uint8_t q[3];
    for (int k = 0; k < 3; k++) { float cl = r[k] < 0.0f ? 0.0f : (r[k] > 1.0f ? 1.0f : r[k]); q[k] = (unsigned char)(cl * 255.0f); } // toRGBCOL
    float lin[3] = {float(q[0]) / 255.0f, float(q[1]) / 255.0f, float(q[2]) / 255.0f};                                              // fromRGBCOL
    gamma_to_rgba8(lin, out);
}

// What applyImagePipeline does after the filter has written the RGBE stage (ImagePipeline.cu:71-82): copyFilteredToOutput, or ToneMapPostProcess + gamma
void orc_pipeline_from_stage2(const uint8_t* rgbe, int w, int h, const ctl_image_pipeline* P, uint8_t* rgba, float* lum_out) {
    const size_t n = (size_t)w * h;
    if (!P->tonemap) { // copyFilteredToOutput
        for (size_t i = 0; i < n; i++) { float c[3]; from_rgbe(&rgbe[4 * i], c); gamma_to_rgba8(c, rgba + 4 * i); }
        return;
    }
    float lum[4]; luminance_info(rgbe, w, h, lum); // ToneMapPostProcess::Apply (ToneMapPostProcess.cu:27-39)
    const float scale = P->key / lum[3], Lwhite = lum[1] * scale;
    const float burn = std::min(1.0f, std::max(1e-8f, 1.0f - P->burn));
    const float invWp2 = 1 / (Lwhite * Lwhite * std::pow(burn, 4.0f));
    if (lum_out) { memcpy(lum_out, lum, sizeof(lum)); lum_out[4] = scale; lum_out[5] = invWp2; }
    for (size_t i = 0; i < n; i++) reinhard_pixel(&rgbe[4 * i], scale, invWp2, rgba + 4 * i);
}

// NonLocalMeansFilter::Apply (Kernel/ImagePipeline/Filter/NonLocalMeansFilter.cu:184-228) without its tile cache: R = 6 search window, F = 3 patch
// radius.  Stage 1 copyToCached (:150-158): PixelData -> RGBE.  Stage 2 computeWeights (:100-121) when asked: for every pixel p and every q of its
// 13x13 window inside the image, weight(p, q) (:91-98) = exp(-max(0, patchDistance)) cut to 0 below 0.05, patchDistance (:67-89) = mean over the
// 7x7 patch offsets that are inside the image around both p and q of
//     (avg((c_p - c_q)^2) - (var_p + min(var_p, var_q))) / (1e-10 + k^2 (var_p + var_q)),
// colours decoded from the RGBE cache, variances = PixelVarianceInfo::computeVariance() rounded to half precision (the shared-memory cache holds
// halves, :14,32) times sigma2Scale.  Stage 3 applyWeights (:123-148): weighted mean of the window's colours (NaN weights skipped), the pixel itself
// when the weights sum to <= 1e-4, written to the RGBE stage.  weights: w*h*169 floats, slot (yo + 6) * 13 + (xo + 6) of pixel y*w + x
// (NonLocalMeansFilter.h:13-31, 63-66), kept by the filter between calls: compute_weights = 0 applies them as they are.
void orc_nlm_filter(const ctl_pixel_data* img, const ctl_pixel_variance_info* var, int w, int h, float splat_scale, float k, float sigma2Scale,
                    float* weights, int compute_weights, uint8_t* rgbe_out) {
    const int R = 6, F = 3, NW = (2 * R + 1) * (2 * R + 1);
    const size_t n = (size_t)w * h;
    std::vector<uint8_t> cached(4 * n);
    std::vector<float> col(3 * n), varh(n);
    for (size_t i = 0; i < n; i++) {
        float c[3]; px_to_spectrum(img[i], splat_scale, c); to_rgbe(c, &cached[4 * i]); from_rgbe(&cached[4 * i], &col[3 * i]);
        const float invN = 1.0f / (float)var[i].num_samples_var;                                     // VarianceFromMoments, Math/VarAccumulator.h:7-11
        const float v = (var[i].sum_x2 - (var[i].sum_x * var[i].sum_x) * invN) * invN;
        varh[i] = (float)(_Float16)v;                                                                 // half(float): round to nearest even (Math/half.h:21-66)
    }
    if (compute_weights) {
        memset(weights, 0, n * NW * sizeof(float));
        const float eps = 1e-10f, alpha = 1.0f;
        auto work = [&](int y0, int y1) {
            for (int y = y0; y < y1; y++) for (int x = 0; x < w; x++) for (int xo = -R; xo <= R; xo++) for (int yo = -R; yo <= R; yo++) {
                const int qx = x + xo, qy = y + yo;
                if (qx < 0 || qx >= w || qy < 0 || qy >= h) continue;
                float d_range = 0, cnt = 0;
                for (int dx = -F; dx <= F; dx++) for (int dy = -F; dy <= F; dy++) {
                    if (x + dx < 0 || x + dx >= w || y + dy < 0 || y + dy >= h || qx + dx < 0 || qx + dx >= w || qy + dy < 0 || qy + dy >= h) continue;
                    const size_t ip = (size_t)(y + dy) * w + (x + dx), iq = (size_t)(qy + dy) * w + (qx + dx);
                    float var_p = varh[ip], var_q = varh[iq];
                    var_p *= sigma2Scale; var_q *= sigma2Scale;
                    float e[3]; for (int c = 0; c < 3; c++) { const float t = col[3 * ip + c] - col[3 * iq + c]; e[c] = t * t; }
                    const float u_diff = (((0.0f + e[0]) + e[1]) + e[2]) * (1.0f / 3);
                    const float d = (u_diff - alpha * (var_p + std::min(var_p, var_q))) / (eps + k * k * (var_p + var_q));
                    d_range += d; cnt++;
                }
                const float dist = cnt != 0 ? d_range / cnt : 0;
                const float we = expf(-std::max(0.0f, dist));
                weights[((size_t)y * w + x) * NW + (size_t)(yo + R) * (2 * R + 1) + (xo + R)] = we < 0.05f ? 0.0f : we;
            }
        };
        const int nt = std::max(1, std::min((int)std::thread::hardware_concurrency(), h));
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, (int)((long long)h * t / nt), (int)((long long)h * (t + 1) / nt));
        for (auto& t : th) t.join();
    }
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        float acc[3] = {0, 0, 0}, C_p = 0;
        for (int xo = -R; xo <= R; xo++) for (int yo = -R; yo <= R; yo++) {
            const int qx = x + xo, qy = y + yo;
            if (qx < 0 || qx >= w || qy < 0 || qy >= h) continue;
            const float we = weights[((size_t)y * w + x) * NW + (size_t)(yo + R) * (2 * R + 1) + (xo + R)];
            if (we != we) continue;
            const float* cq = &col[3 * ((size_t)qy * w + qx)];
            C_p += we;
            for (int c = 0; c < 3; c++) acc[c] += we * cq[c];
        }
        float out[3];
        if (C_p > 1e-4f) { const float r = 1.0f / C_p; for (int c = 0; c < 3; c++) out[c] = acc[c] * r; }   // Spectrum / float multiplies by the reciprocal
        else for (int c = 0; c < 3; c++) out[c] = col[3 * ((size_t)y * w + x) + c];
        to_rgbe(out, &rgbe_out[4 * ((size_t)y * w + x)]);
    }
}

void orc_apply_image_pipeline(const ctl_pixel_data* img, int w, int h, float splat_scale, const ctl_image_pipeline* P, uint8_t* rgba, float* lum_out) {
    const size_t n = (size_t)w * h;
    if (P->filter_type < 0 && !P->tonemap) { // copySamplesToOutput
        for (size_t i = 0; i < n; i++) { float c[3]; px_to_spectrum(img[i], splat_scale, c); gamma_to_rgba8(c, rgba + 4 * i); }
        return;
    }
    std::vector<uint8_t> rgbe(4 * n); // Stage 2 of the reference Image: the RGBE buffer
    if (P->filter_type >= 0) {
        FilterEval F(*P);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { float c[3]; eval_filter(F, img, splat_scale, x, y, w, h, c); to_rgbe(c, &rgbe[4 * ((size_t)y * w + x)]); }
    } else {
        for (size_t i = 0; i < n; i++) { float c[3]; px_to_spectrum(img[i], splat_scale, c); to_rgbe(c, &rgbe[4 * i]); } // copySamplesToFiltered
    }
    orc_pipeline_from_stage2(rgbe.data(), w, h, P, rgba, lum_out);
}
void orc_resolve_srgb8(const ctl_pixel_data* img, int n, float splat_scale, uint8_t* rgba) {
    ctl_image_pipeline P; memset(&P, 0, sizeof(P)); P.filter_type = -1;
    orc_apply_image_pipeline(img, n, 1, splat_scale, &P, rgba, nullptr);
}
void orc_resolve_filtered_srgb8(const ctl_pixel_data* img, int w, int h, float splat_scale, int type, float xw, float yw, float alpha, uint8_t* rgba) {
    ctl_image_pipeline P; memset(&P, 0, sizeof(P)); P.filter_type = type; P.x_width = xw; P.y_width = yw; P.param0 = alpha;
    orc_apply_image_pipeline(img, w, h, splat_scale, &P, rgba, nullptr);
}

// PixelVarianceInfo::updateMoments over the image = PixelVarianceBuffer::AddPass with the uniform block sampler (every block sampled once per
// pass: samplerPerformed = 1) (Kernel/PixelVarianceBuffer.h:19-42, .cu:10-36)
void orc_variance_add_pass(ctl_pixel_variance_info* var, const ctl_pixel_data* img, int n, float splat_scale) {
    for (int i = 0; i < n; i++) {
        ctl_pixel_variance_info& V = var[i]; const ctl_pixel_data& P = img[i];
        const float samplerPerformed = 1.0f, recip = 1.0f / samplerPerformed;
        float est[3];
        for (int k = 0; k < 3; k++) { float nps = P.rgb[k] + P.rgb_splat[k] * splat_scale; est[k] = (nps - V.prev_I[k]) * recip; V.prev_I[k] = nps; }
        V.weight = P.weight_sum;
        if (V.iterations_done++ % 2 == 1) for (int k = 0; k < 3; k++) V.half_buffer[k] += est[k];
        const float Y = est[0] * 0.212671f + est[1] * 0.715160f + est[2] * 0.072169f;
        V.sum_x += Y; V.sum_x2 += Y * Y; V.num_samples_var++;
    }
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
