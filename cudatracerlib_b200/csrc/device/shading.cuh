// shading.cuh -- device shading functions of the radiance loop.
//
// Replaces (for the in-scope subset) Kernel/TraceHelper.cu:274-307 fillDG, Kernel/TraceResult.cu:16-43
// getBsdfSample, SceneTypes/BSDF_Simple.cu diffuse (7-75) / dielectric (174-277) / roughconductor (662-763),
// Engine/MicrofacetDistribution.{h,cu}, Math/FresnelHelper.h, Math/Warp.h, SceneTypes/Light.cu:67-155
// DiffuseLight, Engine/ShapeSet.cu:51-69, Kernel/Sampler_device.h:62-107, SceneTypes/Sensor.cu:130-144.
// Full-precision math (sinf/cosf/expf/logf...), evaluation order as in the reference host branches.
#pragma once
#include "traverse.cuh"

namespace ctld {

constexpr int N_SEQ = 4096, SEQ_LEN = 30;

struct Sampler { // SequenceSampler, Kernel/Sampler_device.h:62-107; `tab` selects the pass's table when several passes share a wavefront
    unsigned idx, i1, i2, tab;
    CTL_DEV float f1(const DScene& S) {
        const unsigned a = idx % N_SEQ, b = (idx / N_SEQ) % N_SEQ, e = i1 % SEQ_LEN + tab * SEQ_LEN;
        float sum = 0.0f; sum += __ldg(S.d1 + e * N_SEQ + a); sum += __ldg(S.d1 + e * N_SEQ + b);
        i1++;
        return sum - floorf(sum);
    }
    CTL_DEV float2 f2(const DScene& S) {
        const unsigned a = idx % N_SEQ, b = (idx / N_SEQ) % N_SEQ, e = i2 % SEQ_LEN + tab * SEQ_LEN;
        const float2 va = __ldg(S.d2 + e * N_SEQ + a), vb = __ldg(S.d2 + e * N_SEQ + b);
        float sx = 0.0f, sy = 0.0f; sx += va.x; sy += va.y; sx += vb.x; sy += vb.y;
        i2++;
        return make_float2(sx - floorf(sx), sy - floorf(sy));
    }
};

CTL_DEV V3 dec_normal(const DScene& S, uint32_t c) { // Math/Compression.h:20-31 via host-built sin/cos tables
    const unsigned x = (c >> 8) & 0xff, y = c & 0xff;
    const float st = __ldg(S.normal_lut + x), ct = __ldg(S.normal_lut + 256 + x), sp_ = __ldg(S.normal_lut + 512 + y), cp = __ldg(S.normal_lut + 768 + y);
    return mk(st * cp, st * sp_, ct);
}

struct DG { V3 P; Frame sys; V3 n; };

CTL_DEV void fill_dg(const DScene& S, float bu, float bv, uint32_t tri, uint32_t node, DG& dg, uint32_t& mat_local) {
    const float4* l2w = S.node_xf + (size_t)node * 4;
    const uint4 w0 = __ldg(S.tri_data + (size_t)tri * 2), w1 = __ldg(S.tri_data + (size_t)tri * 2 + 1);
    mat_local = (w0.y >> 16) & 0xff; // TriangleData::getMatIndex, TriangleData.h:40-44
    const V3 na = dec_normal(S, w0.x & 0xffff), nb = dec_normal(S, w0.x >> 16), nc = dec_normal(S, w0.y & 0xffff);
    const float ww = 1.0f - bu - bv, u = bu, v = bv;
    const V3 n = normalize(na * u + nb * v + nc * ww);
    const V3 dpdu = mk(h2f(w0.z), h2f(w0.z >> 16), h2f(w0.w));
    const V3 dpdv = mk(h2f(w0.w >> 16), h2f(w1.x), h2f(w1.x >> 16));
    V3 s = dpdu - n * dot(n, dpdu);
    V3 t = cross(s, n);
    s = xf_dir(l2w, s); t = xf_dir(l2w, t);
    dg.sys.s = normalize(s); dg.sys.t = normalize(t); dg.sys.n = normalize(cross(t, s));
    const V3 wdpdu = xf_dir(l2w, dpdu), wdpdv = xf_dir(l2w, dpdv);
    dg.n = normalize(cross(wdpdu, wdpdv));
    if (dot(dg.n, dg.sys.n) < 0.0f) dg.n = -dg.n;
}

CTL_DEV V3 cosine_hemisphere(float sx, float sy) { // Math/Warp.h:61-125
    const float r1 = 2.0f * sx - 1.0f, r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (PI_F / 4.0f) * (r2 / r1); }
    else { r = r2; phi = (PI_F / 2.0f) - (r1 / r2) * (PI_F / 4.0f); }
    const float cp = cosf(phi), sp_ = sinf(phi);
    const float px = r * cp, py = r * sp_;
    return mk(px, py, sqrtf(1.0f - px * px - py * py));
}

CTL_DEV float fresnel_dielectric_ext(float cosThetaI_, float& cosThetaT_, float eta) { // FresnelHelper.h:27-58
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    const float scale = (cosThetaI_ > 0) ? 1.0f / eta : eta, cosThetaTSqr = 1.0f - (1.0f - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    const float cosThetaI = fabsf(cosThetaI_), cosThetaT = sqrtf(fmaxf(0.0f, cosThetaTSqr));
    const float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    const float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
CTL_DEV Spec fresnel_conductor_exact(float cosThetaI, Spec eta, Spec k) { // FresnelHelper.h:119-142
    const float c2 = cosThetaI * cosThetaI, s2 = 1 - c2, s4 = s2 * s2;
    const Spec temp1 = eta * eta - k * k - sp(s2);
    const Spec a2pb2 = safe_sqrt(temp1 * temp1 + k * k * eta * eta * 4);
    const Spec a = safe_sqrt((a2pb2 + temp1) * 0.5f);
    const Spec term1 = a2pb2 + sp(c2), term2 = a * (2 * cosThetaI);
    const Spec Rs2 = (term1 - term2) / (term1 + term2);
    const Spec term3 = a2pb2 * c2 + sp(s4), term4 = term2 * s2;
    const Spec Rp2 = (Rs2 * (term3 - term4)) / (term3 + term4);
    return (Rp2 + Rs2) * 0.5f;
}

CTL_DEV float m_erfinv(float x) { // MathFunc.h:343-373 (Giles)
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
CTL_DEV float m_erf(float x) { // MathFunc.h:375-393 (A&S 7.1.26)
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    const float sign = copysignf(1.0f, x); x = fabsf(x);
    const float t = 1.0f / (1.0f + p * x);
    const float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * expf(-x * x);
    return sign * y;
}
CTL_DEV float m_hypot2(float a, float b) { // MathFunc.h:326-341
    float r;
    if (fabsf(a) > fabsf(b)) { r = b / a; r = fabsf(a) * sqrtf(1.0f + r * r); }
    else if (b != 0.0f) { r = a / b; r = fabsf(b) * sqrtf(1.0f + r * r); }
    else r = 0.0f;
    return r;
}

struct Distr { // Engine/MicrofacetDistribution.h:12-171, Beckmann + GGX, visible-normal sampling
    int type; float au, av;
    CTL_DEV Distr(int t, float a, float b) : type(t), au(a > 1e-4f ? a : 1e-4f), av(b > 1e-4f ? b : 1e-4f) {}
    CTL_DEV float eval(V3 m) const { // .cu:6-42
        if (m.z <= 0) return 0.0f;
        const float c2 = m.z * m.z;
        const float be = ((m.x * m.x) / (au * au) + (m.y * m.y) / (av * av)) / c2;
        float result;
        if (type == CTL_DISTR_BECKMANN) result = expf(-be) / (PI_F * au * av * c2 * c2);
        else { const float root = (1 + be) * c2; result = 1.0f / (PI_F * au * av * root * root); }
        if (result < 1e-20f) result = 0;
        return result;
    }
    CTL_DEV float project_roughness(V3 v) const {
        const float invSinTheta2 = 1 / (1.0f - v.z * v.z);
        if (au == av || invSinTheta2 <= 0) return au;
        const float cosPhi2 = v.x * v.x * invSinTheta2, sinPhi2 = v.y * v.y * invSinTheta2;
        return sqrtf(cosPhi2 * au * au + sinPhi2 * av * av);
    }
    CTL_DEV float smith_g1(V3 v, V3 m) const { // .cu:306-340
        if (dot(v, m) * v.z <= 0) return 0.0f;
        const float temp = 1 - v.z * v.z;
        const float tanTheta = fabsf(temp <= 0.0f ? 0.0f : sqrtf(temp) / v.z);
        if (tanTheta == 0.0f) return 1.0f;
        const float alpha = project_roughness(v);
        if (type == CTL_DISTR_BECKMANN) {
            const float a = 1.0f / (alpha * tanTheta);
            if (a >= 1.6f) return 1.0f;
            const float aSqr = a * a;
            return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
        }
        const float root = alpha * tanTheta;
        return 2.0f / (1.0f + m_hypot2(1.0f, root));
    }
    CTL_DEV float pdf_visible(V3 wi, V3 m) const {
        if (wi.z == 0) return 0.0f;
        return smith_g1(wi, m) * fabsf(dot(wi, m)) * eval(m) / fabsf(wi.z);
    }
    CTL_DEV void sample_visible11(float thetaI, float sx, float sy, float& slx, float& sly) const { // .cu:188-304
        const float SQRT_PI_INV = 1 / sqrtf(PI_F);
        if (type == CTL_DISTR_BECKMANN) {
            if (thetaI < 1e-4f) {
                const float r = sqrtf(-logf(1.0f - sx));
                const float sinPhi = sinf(2 * PI_F * sy), cosPhi = cosf(2 * PI_F * sy);
                slx = r * cosPhi; sly = r * sinPhi; return;
            }
            const float tanThetaI = tanf(thetaI), cotThetaI = 1 / tanThetaI;
            float a = -1, c = m_erf(cotThetaI);
            const float sample_x = sx > 1e-6f ? sx : 1e-6f;
            const float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1 + c) * powf(1 - sample_x, fit);
            const float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * expf(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                const float invErf = m_erfinv(b);
                const float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * expf(-invErf * invErf)) - sample_x;
                const float derivative = normalization * (1 - invErf * tanThetaI);
                if (fabsf(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slx = m_erfinv(b);
            sly = m_erfinv(2.0f * (sy > 1e-6f ? sy : 1e-6f) - 1.0f);
            return;
        }
        if (thetaI < 1e-4f) {
            const float r = sqrtf(fmaxf(0.0f, sx / (1 - sx)));
            const float sinPhi = sinf(2 * PI_F * sy), cosPhi = cosf(2 * PI_F * sy);
            slx = r * cosPhi; sly = r * sinPhi; return;
        }
        const float tanThetaI = tanf(thetaI), a = 1 / tanThetaI;
        const float G1 = 2.0f / (1.0f + sqrtf(fmaxf(0.0f, 1.0f + 1.0f / (a * a))));
        float A = 2.0f * sx / G1 - 1.0f;
        if (fabsf(A) == 1) A -= copysignf(1.0f, A) * 1e-7f;
        const float tmp = 1.0f / (A * A - 1.0f), B = tanThetaI;
        const float D = sqrtf(fmaxf(0.0f, B * B * tmp * tmp - (A * A - B * B) * tmp));
        const float s1 = B * tmp - D, s2 = B * tmp + D;
        slx = (A < 0.0f || s2 > 1.0f / tanThetaI) ? s1 : s2;
        float Sg;
        if (sy > 0.5f) { Sg = 1.0f; sy = 2.0f * (sy - 0.5f); } else { Sg = -1.0f; sy = 2.0f * (0.5f - sy); }
        const float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                        (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
        sly = Sg * z * sqrtf(1.0f + slx * slx);
    }
    CTL_DEV V3 sample_visible(V3 _wi, float sx, float sy) const { // .cu:151-186
        const V3 wi = normalize(mk(au * _wi.x, av * _wi.y, _wi.z));
        float theta = 0, phi = 0;
        if (wi.z < 0.99999f) { theta = acosf(wi.z); phi = atan2f(wi.y, wi.x); }
        const float sinPhi = sinf(phi), cosPhi = cosf(phi);
        float slx, sly; sample_visible11(theta, sx, sy, slx, sly);
        float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
        rx *= au; ry *= av;
        const float nrm = 1.0f / sqrtf(rx * rx + ry * ry + (float)1.0);
        return mk(-rx * nrm, -ry * nrm, nrm);
    }
};

struct BRec { V3 wi, wo; float eta; unsigned typeMask, sampledType; };

CTL_DEV unsigned bsdf_combined_type(uint32_t bsdf_type) {
    return bsdf_type == CTL_BSDF_DIFFUSE ? E_DIFFUSE_REFL : (bsdf_type == CTL_BSDF_ROUGHCONDUCTOR ? E_GLOSSY_REFL : (E_DELTA_REFL | E_DELTA_TRANS));
}
CTL_DEV float mat_alpha(float a) { return savg(sp(a)); }

CTL_DEV Spec bsdf_sample_inner(const ctl_material& m, BRec& b, float& pdf, float sx, float sy) {
    if (m.bsdf_type == CTL_BSDF_DIFFUSE) { // BSDF_Simple.cu:7-27
        if (!(b.typeMask & E_DIFFUSE_REFL) || b.wi.z <= 0) return sp(0.0f);
        b.sampledType = E_DIFFUSE_REFL;
        b.wo = cosine_hemisphere(sx, sy);
        b.eta = 1.0f;
        pdf = fabsf(INV_PI_F * b.wo.z) * 1;
        return sp3(m.reflectance) * 1;
    }
    if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) { // :662-705
        if (b.wi.z < 0 || !(b.typeMask & E_GLOSSY_REFL)) return sp(0.0f);
        const Distr distr(m.distr_type, mat_alpha(m.alpha_u), mat_alpha(m.alpha_v));
        const V3 mm = distr.sample_visible(b.wi, sx, sy);
        pdf = distr.pdf_visible(b.wi, mm);
        if (pdf == 0) return sp(0.0f);
        b.wo = normalize(mm * (2 * dot(b.wi, mm)) - b.wi);
        b.eta = 1.0f; b.sampledType = E_GLOSSY_REFL;
        if (b.wo.z <= 0) return sp(0.0f);
        const Spec F = fresnel_conductor_exact(dot(b.wi, mm), sp3(m.eta), sp3(m.k)) * sp3(m.reflectance);
        const float weight = distr.smith_g1(b.wo, mm);
        pdf /= 4.0f * dot(b.wo, mm);
        return F * weight;
    }
    // dielectric :174-225, no dispersion
    const bool sR = (b.typeMask & E_DELTA_REFL) != 0, sT = (b.typeMask & E_DELTA_TRANS) != 0;
    float cosThetaT;
    const float eta = m.eta[0], invEta = 1.0f / eta;
    const float F = fresnel_dielectric_ext(b.wi.z, cosThetaT, eta);
    const float rscale = -(cosThetaT < 0 ? invEta : eta);
    if (sT && sR) {
        if (sx <= F) { b.sampledType = E_DELTA_REFL; b.wo = mk(-b.wi.x, -b.wi.y, b.wi.z); b.eta = 1.0f; pdf = F; return sp3(m.reflectance); }
        b.sampledType = E_DELTA_TRANS; b.wo = normalize(mk(rscale * b.wi.x, rscale * b.wi.y, cosThetaT)); b.eta = cosThetaT < 0 ? eta : invEta; pdf = (1 - F) * 1.0f;
        const float factor = cosThetaT < 0 ? invEta : eta;
        return (sp(1.0f) * sp(m.transmittance)) * (factor * factor);
    } else if (sR) { b.sampledType = E_DELTA_REFL; b.wo = mk(-b.wi.x, -b.wi.y, b.wi.z); b.eta = 1.0f; pdf = 1.0f; return sp3(m.reflectance); }
    else if (sT) {
        b.sampledType = E_DELTA_TRANS; b.wo = normalize(mk(rscale * b.wi.x, rscale * b.wi.y, cosThetaT)); b.eta = cosThetaT < 0 ? eta : invEta; pdf = 1.0f;
        const float factor = cosThetaT < 0 ? invEta : eta;
        return (sp(1.0f) * sp(m.transmittance)) * (factor * factor * (1 - F));
    }
    return sp(0.0f);
}
CTL_DEV Spec bsdf_f_inner(const ctl_material& m, const BRec& b) { // measure ESolidAngle
    if (m.bsdf_type == CTL_BSDF_DIFFUSE) {
        if (!(b.typeMask & E_DIFFUSE_REFL)) return sp(0.0f);
        const bool validRefl = b.wi.z > 0 && b.wo.z > 0;
        const Spec s = sp3(m.reflectance) * (INV_PI_F * fabsf(b.wo.z));
        return validRefl ? s : sp(0.0f);
    }
    if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) {
        if (b.wi.z < 0 || b.wo.z < 0 || !(b.typeMask & E_GLOSSY_REFL)) return sp(0.0f);
        const V3 H = normalize(b.wo + b.wi);
        const Distr distr(m.distr_type, mat_alpha(m.alpha_u), mat_alpha(m.alpha_v));
        const float D = distr.eval(H);
        if (D == 0) return sp(0.0f);
        const Spec F = fresnel_conductor_exact(dot(b.wi, H), sp3(m.eta), sp3(m.k)) * sp3(m.reflectance);
        const float G = distr.smith_g1(b.wi, H) * distr.smith_g1(b.wo, H);
        const float value = D * G / (4.0f * b.wi.z);
        return F * value;
    }
    return sp(0.0f);
}
CTL_DEV float bsdf_pdf_inner(const ctl_material& m, const BRec& b) {
    if (m.bsdf_type == CTL_BSDF_DIFFUSE) {
        if (!(b.typeMask & E_DIFFUSE_REFL)) return 0.0f;
        const bool validRefl = b.wi.z > 0 && b.wo.z > 0;
        return validRefl ? fabsf(INV_PI_F * b.wo.z) : 0.0f;
    }
    if (m.bsdf_type == CTL_BSDF_ROUGHCONDUCTOR) {
        if (b.wi.z < 0 || b.wo.z < 0 || !(b.typeMask & E_GLOSSY_REFL)) return 0.0f;
        const V3 H = normalize(b.wo + b.wi);
        const Distr distr(m.distr_type, mat_alpha(m.alpha_u), mat_alpha(m.alpha_v));
        return distr.eval(H) * distr.smith_g1(b.wi, H) / (4.0f * b.wi.z);
    }
    return 0.0f;
}
// BSDFALL two-sided wrapper (SceneTypes/BSDF.h:141-207)
CTL_DEV Spec bsdf_sample(const ctl_material& m, BRec& b, float& pdf, float sx, float sy) {
    const bool flip = b.wi.z < 0 && (m.flags & CTL_MAT_TWO_SIDED);
    if (flip) b.wi.z *= -1.0f;
    const Spec r = bsdf_sample_inner(m, b, pdf, sx, sy);
    if (flip) { b.wi.z *= -1.0f; b.wo.z *= -1.0f; }
    return r;
}
CTL_DEV Spec bsdf_f(const ctl_material& m, BRec& b) {
    const bool flip = b.wi.z < 0 && (m.flags & CTL_MAT_TWO_SIDED);
    if (flip) b.wi.z *= -1.0f;
    const Spec r = bsdf_f_inner(m, b);
    if (flip) { b.wi.z *= -1.0f; b.wo.z *= -1.0f; }
    return r;
}
CTL_DEV float bsdf_pdf(const ctl_material& m, BRec& b) {
    const bool flip = b.wi.z < 0 && (m.flags & CTL_MAT_TWO_SIDED);
    if (flip) b.wi.z *= -1.0f;
    const float r = bsdf_pdf_inner(m, b);
    if (flip) { b.wi.z *= -1.0f; b.wo.z *= -1.0f; }
    return r;
}

struct DRec { V3 p, n; float pdf; V3 ref, refN, d; float dist; };

// DiffuseLight::sampleDirect (Light.cu:84-135) over ShapeSet::SamplePosition (ShapeSet.cu:51-69),
// MonteCarlo::sampleReuse (MonteCarlo.cu:7-14, STL_lower_bound Base/STL.h:40-57)
CTL_DEV Spec light_sample_direct(const DScene& S, const ctl_light& L, DRec& dRec, float sx, float sy) {
    const float* cdf = S.light_cdf_data + L.cdf_offset;
    unsigned first = 0, count = L.count + 1;
    while (count > 0) { const unsigned c2 = count / 2, mid = first + c2; if (__ldg(cdf + mid) < sy) { first = mid + 1; count -= c2 + 1; } else count = c2; }
    int ii = (int)first - 1; if (ii < 0) ii = 0; if (ii > (int)L.count - 1) ii = (int)L.count - 1;
    const unsigned index = (unsigned)ii;
    const float c_lo = __ldg(cdf + index), pdf = __ldg(cdf + index + 1) - c_lo;
    sy = (sy - c_lo) / pdf;
    const float* sn = (const float*)(S.light_tris + L.tri_offset + index);
    const float a = sqrtf(1.0f - sx); const float b0 = 1 - a, b1 = a * sy; // Warp::squareToUniformTriangle
    const V3 p0 = mk(__ldg(sn + 0), __ldg(sn + 1), __ldg(sn + 2)), p1 = mk(__ldg(sn + 3), __ldg(sn + 4), __ldg(sn + 5)), p2 = mk(__ldg(sn + 6), __ldg(sn + 7), __ldg(sn + 8));
    dRec.p = p0 * b0 + p1 * b1 + p2 * (1.f - b0 - b1);
    dRec.n = mk(__ldg(sn + 9), __ldg(sn + 10), __ldg(sn + 11));
    dRec.pdf = 1.0f / L.sum_area;
    const V3 dir = dRec.p - dRec.ref;
    const float distSquared = dot(dir, dir);
    dRec.dist = sqrtf(distSquared);
    dRec.d = dir / dRec.dist;
    const float dp = fabsf(dot(dRec.d, dRec.n));
    dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) return sp3(L.radiance) / dRec.pdf;
    dRec.pdf = 0.0f;
    return sp(0.0f);
}
CTL_DEV float light_pdf_direct(const ctl_light& L, const DRec& dRec) { // Light.cu:137-155
    if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) {
        const float pdfPos = 1.0f / L.sum_area;
        return pdfPos * (dRec.dist * dRec.dist) / fabsf(dot(dRec.d, dRec.n));
    }
    return 0.0f;
}
CTL_DEV float pdf_emitter(const DScene& S, unsigned li) { return S.light_cdf[li] - (li == 0 ? 0.0f : S.light_cdf[li - 1]); } // KernelDynamicScene.cu:42-46
CTL_DEV float power_heuristic(float fPdf, float gPdf) { const float f = 1 * fPdf, g = 1 * gPdf; return (f * f) / (f * f + g * g); } // MonteCarlo.h:29-33

CTL_DEV void camera_ray(const DScene& S, float px, float py, V3& o, V3& d) { // Sensor.cu:130-144
    const float* m = S.camera.sample_to_camera;
    const float qx = px * S.camera.inv_resolution[0], qy = py * S.camera.inv_resolution[1], qz = 0.0f;
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { float acc = 0.0f; acc += m[i * 4 + 0] * qx; acc += m[i * 4 + 1] * qy; acc += m[i * 4 + 2] * qz; acc += m[i * 4 + 3] * 1.0f; r[i] = acc; }
    const V3 nearP = mk(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
    const V3 dn = normalize(nearP);
    const float* tw = S.camera.to_world;
    o = mk(tw[3], tw[7], tw[11]);
    float dd[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { float acc = 0.0f; acc += tw[i * 4 + 0] * dn.x; acc += tw[i * 4 + 1] * dn.y; acc += tw[i * 4 + 2] * dn.z; acc += tw[i * 4 + 3] * 0.0f; dd[i] = acc; }
    d = mk(dd[0], dd[1], dd[2]);
}

} // namespace ctld
