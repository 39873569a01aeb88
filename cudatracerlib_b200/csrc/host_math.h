// host_math.h -- host-side vector / matrix / codec helpers used by the scene builder.
//
// Arithmetic follows the reference's operation order so that the encoded data
// surface is what CudaTracerLib's own encoders would produce:
//   Math/Vector.h (dot accumulates left to right, normalize = v * rcp(len), rcp(0)=0),
//   Math/float4x4.h:132-193 (cofactor inverse), 398-408 (TransformPoint divides by w),
//   Math/Compression.h:12-31 (16-bit spherical normal codec),
//   Math/half.h:20-82 (IEEE binary16 round-to-nearest-even; decode per IEEE).
// Host code; the vector / matrix-inverse helpers are also callable from device code (CTLB_HD) so that the GPU BVH builder
// encodes Woop triangles with the very same expressions.  Not used by oracle/ (which carries its own restatement).
#pragma once
#ifdef __CUDACC__
#define CTLB_HD __host__ __device__
#else
#define CTLB_HD
#endif
#include <cmath>
#include <cstdint>
#include <cstring>

namespace ctlb {

struct V3 {
    float x, y, z;
    CTLB_HD V3() : x(0), y(0), z(0) {}
    CTLB_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
    CTLB_HD explicit V3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
CTLB_HD inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
CTLB_HD inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
CTLB_HD inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
CTLB_HD inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
CTLB_HD inline V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
CTLB_HD inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
CTLB_HD inline float dot(V3 a, V3 b) { float r = 0.0f; r += a.x * b.x; r += a.y * b.y; r += a.z * b.z; return r; }
CTLB_HD inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float len_sqr(V3 a) { return dot(a, a); }
inline float length(V3 a) { return sqrtf(len_sqr(a)); }
inline float rcp(float a) { return a != 0.0f ? 1.0f / a : 0.0f; }
inline V3 normalize(V3 a) { return a * rcp(length(a)); }
CTLB_HD inline V3 vmin(V3 a, V3 b) { return V3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
CTLB_HD inline V3 vmax(V3 a, V3 b) { return V3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

struct Box {
    V3 lo, hi;
    Box() : lo(3.0e38f), hi(-3.0e38f) {}
    Box(V3 a, V3 b) : lo(a), hi(b) {}
    void grow(V3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
    float area() const {
        V3 d = hi - lo;
        if (d.x < 0 || d.y < 0 || d.z < 0) return 0.0f;
        return 2.0f * (d.x * d.y + d.y * d.z + d.z * d.x);
    }
    V3 center() const { return (lo + hi) * 0.5f; }
};

// Row-major 4x4, column-vector convention (M * v), Math/float4x4.h:12-18.
struct M4 {
    float m[16];
    CTLB_HD float operator()(int r, int c) const { return m[r * 4 + c]; }
    CTLB_HD float& operator()(int r, int c) { return m[r * 4 + c]; }
    static M4 identity() {
        M4 r; for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f; return r;
    }
    static M4 translate(V3 t) { M4 r = identity(); r(0, 3) = t.x; r(1, 3) = t.y; r(2, 3) = t.z; return r; }
    static M4 scale(V3 s) { M4 r = identity(); r(0, 0) = s.x; r(1, 1) = s.y; r(2, 2) = s.z; return r; }
    static M4 rotate_y(float a) {
        M4 r = identity(); float c = cosf(a), s = sinf(a);
        r(0, 0) = c; r(0, 2) = s; r(2, 0) = -s; r(2, 2) = c; return r;
    }
    // Math/float4x4.h:229-244
    static M4 perspective(float fov, float clip_near, float clip_far) {
        float recip = 1.0f / (clip_far - clip_near);
        float cot = 1.0f / tanf(fov / 2.0f);
        M4 r; memset(r.m, 0, sizeof(r.m));
        r(0, 0) = cot; r(1, 1) = cot; r(2, 2) = clip_far * recip; r(2, 3) = -clip_near * clip_far * recip; r(3, 2) = 1.0f;
        return r;
    }
    // Math/float4x4.h:612-623
    static M4 look_at(V3 p, V3 t, V3 up) {
        V3 dir = normalize(t - p), left = normalize(cross(up, dir)), new_up = cross(dir, left);
        M4 r = identity();
        r(0, 0) = left.x; r(1, 0) = left.y; r(2, 0) = left.z;
        r(0, 1) = new_up.x; r(1, 1) = new_up.y; r(2, 1) = new_up.z;
        r(0, 2) = dir.x; r(1, 2) = dir.y; r(2, 2) = dir.z;
        r(0, 3) = p.x; r(1, 3) = p.y; r(2, 3) = p.z;
        return r;
    }
    // 4-term dot, accumulated left to right from zero (Math/Vector.h:97)
    static float dot4(const float* a, float b0, float b1, float b2, float b3) {
        float r = 0.0f; r += a[0] * b0; r += a[1] * b1; r += a[2] * b2; r += a[3] * b3; return r;
    }
    M4 mul(const M4& o) const { // operator% , Math/float4x4.h:365-372
        M4 r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                r(i, j) = dot4(&m[i * 4], o(0, j), o(1, j), o(2, j), o(3, j));
        return r;
    }
    V3 transform_point(V3 p) const { // float4x4.h:398-402 (homogeneous divide)
        float x = dot4(&m[0], p.x, p.y, p.z, 1.0f), y = dot4(&m[4], p.x, p.y, p.z, 1.0f);
        float z = dot4(&m[8], p.x, p.y, p.z, 1.0f), w = dot4(&m[12], p.x, p.y, p.z, 1.0f);
        return V3(x / w, y / w, z / w);
    }
    V3 transform_dir(V3 d) const { // float4x4.h:404-408
        return V3(dot4(&m[0], d.x, d.y, d.z, 0.0f), dot4(&m[4], d.x, d.y, d.z, 0.0f), dot4(&m[8], d.x, d.y, d.z, 0.0f));
    }
    // Adjugate / determinant inverse, evaluation order of Math/float4x4.h:132-193.
    CTLB_HD M4 inverse() const {
        const M4& Q = *this;
        float a00 = Q(0, 0), a01 = Q(0, 1), a02 = Q(0, 2), a03 = Q(0, 3);
        float a10 = Q(1, 0), a11 = Q(1, 1), a12 = Q(1, 2), a13 = Q(1, 3);
        float a20 = Q(2, 0), a21 = Q(2, 1), a22 = Q(2, 2), a23 = Q(2, 3);
        float a30 = Q(3, 0), a31 = Q(3, 1), a32 = Q(3, 2), a33 = Q(3, 3);
        float s0 = a20 * a31 - a21 * a30, s1 = a20 * a32 - a22 * a30, s2 = a20 * a33 - a23 * a30;
        float s3 = a21 * a32 - a22 * a31, s4 = a21 * a33 - a23 * a31, s5 = a22 * a33 - a23 * a32;
        float c00 = +(s5 * a11 - s4 * a12 + s3 * a13);
        float c10 = -(s5 * a10 - s2 * a12 + s1 * a13);
        float c20 = +(s4 * a10 - s2 * a11 + s0 * a13);
        float c30 = -(s3 * a10 - s1 * a11 + s0 * a12);
        float inv_det = 1 / (c00 * a00 + c10 * a01 + c20 * a02 + c30 * a03);
        M4 r;
        r(0, 0) = c00 * inv_det; r(1, 0) = c10 * inv_det; r(2, 0) = c20 * inv_det; r(3, 0) = c30 * inv_det;
        r(0, 1) = -(s5 * a01 - s4 * a02 + s3 * a03) * inv_det;
        r(1, 1) = +(s5 * a00 - s2 * a02 + s1 * a03) * inv_det;
        r(2, 1) = -(s4 * a00 - s2 * a01 + s0 * a03) * inv_det;
        r(3, 1) = +(s3 * a00 - s1 * a01 + s0 * a02) * inv_det;
        s0 = a10 * a31 - a11 * a30; s1 = a10 * a32 - a12 * a30; s2 = a10 * a33 - a13 * a30;
        s3 = a11 * a32 - a12 * a31; s4 = a11 * a33 - a13 * a31; s5 = a12 * a33 - a13 * a32;
        r(0, 2) = +(s5 * a01 - s4 * a02 + s3 * a03) * inv_det;
        r(1, 2) = -(s5 * a00 - s2 * a02 + s1 * a03) * inv_det;
        r(2, 2) = +(s4 * a00 - s2 * a01 + s0 * a03) * inv_det;
        r(3, 2) = -(s3 * a00 - s1 * a01 + s0 * a02) * inv_det;
        s0 = a21 * a10 - a20 * a11; s1 = a22 * a10 - a20 * a12; s2 = a23 * a10 - a20 * a13;
        s3 = a22 * a11 - a21 * a12; s4 = a23 * a11 - a21 * a13; s5 = a23 * a12 - a22 * a13;
        r(0, 3) = -(s5 * a01 - s4 * a02 + s3 * a03) * inv_det;
        r(1, 3) = +(s5 * a00 - s2 * a02 + s1 * a03) * inv_det;
        r(2, 3) = -(s4 * a00 - s2 * a01 + s0 * a03) * inv_det;
        r(3, 3) = +(s3 * a00 - s1 * a01 + s0 * a02) * inv_det;
        return r;
    }
};

// ---- IEEE binary16 (Math/half.h:20-82) ---------------------------------
inline uint16_t float_to_half(float f) {
    uint32_t ia; memcpy(&ia, &f, 4);
    uint16_t ir = (ia >> 16) & 0x8000;
    if ((ia & 0x7f800000) == 0x7f800000) {
        if ((ia & 0x7fffffff) == 0x7f800000) ir |= 0x7c00; else ir = 0x7fff;
    } else if ((ia & 0x7f800000) >= 0x33000000) {
        int shift = (int)((ia >> 23) & 0xff) - 127;
        if (shift > 15) ir |= 0x7c00;
        else {
            ia = (ia & 0x007fffff) | 0x00800000;
            if (shift < -14) { ir |= ia >> (-1 - shift); ia = ia << (32 - (-1 - shift)); }
            else { ir |= ia >> (24 - 11); ia = ia << (32 - (24 - 11)); ir = ir + ((14 + shift) << 10); }
            if ((ia > 0x80000000) || ((ia == 0x80000000) && (ir & 1))) ir++;
        }
    }
    return ir;
}
// IEEE decode (what __half2float does on the device)
inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff, out;
    if (exp == 0) {
        if (man == 0) out = sign;
        else { // subnormal
            int e = -1; do { e++; man <<= 1; } while ((man & 0x400) == 0);
            out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) out = sign | 0x7f800000 | (man << 13);
    else out = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &out, 4); return f;
}

// ---- 16-bit spherical normal codec (Math/Compression.h:12-31) -----------
static const float kPi = 3.14159265358979323846f;
inline uint16_t encode_normal(V3 v) {
    float theta = (acosf(v.z) * (255.0f / kPi));
    float phi = (atan2f(v.y, v.x) * (255.0f / (2.0f * kPi)));
    phi = phi < 0 ? (phi + 255) : phi;
    return (uint16_t)(((uint16_t)theta << 8) | (uint16_t)phi);
}
inline void normal_code_angles(uint16_t code, float& theta, float& phi) {
    const float PI_4 = kPi / 4.0f, PI_2 = kPi / 2.0f;
    unsigned char x = code >> 8, y = code & 0xff;
    theta = x == 63 ? PI_4 : (x == 127 ? PI_2 : (x == 191 ? 3 * PI_4 : float(x) * (1.0f / 255.0f) * kPi));
    phi = y == 63 ? PI_2 : (y == 127 ? kPi : (y == 191 ? 3 * PI_2 : float(y) * (1.0f / 255.0f) * kPi * 2.0f));
}
inline V3 decode_normal(uint16_t code) {
    float theta, phi; normal_code_angles(code, theta, phi);
    float sp = sinf(phi), cp = cosf(phi), st = sinf(theta), ct = cosf(theta);
    return V3(st * cp, st * sp, ct);
}

// Math/Frame.h:9-22
inline void coordinate_system(V3 a, V3& s, V3& t) {
    if (fabsf(a.x) > fabsf(a.y)) {
        float inv_len = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        t = V3(a.z * inv_len, 0.0f, -a.x * inv_len);
    } else {
        float inv_len = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        t = V3(0.0f, a.z * inv_len, -a.y * inv_len);
    }
    s = normalize(cross(t, a));
}

} // namespace ctlb
