// tests/shading_emulate.cpp -- TEST INFRASTRUCTURE (method: tests/nlm_emulate.cpp).  The device shading source (cudatracerlib_b200/csrc/device/shading.cuh:
// BSDF sample / f / pdf with the two-sided wrapper, microfacet distributions, Fresnel terms, DiffuseLight::sampleDirect) compiled for the HOST and probed
// like the oracle's orc_bsdf_probe / orc_light_sample_direct, so that it can be compared with the golden tables minted from the reference's own code.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cfloat>
#include <cmath>
#include <cstring>
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
using std::max; using std::min;
#include "../cudatracerlib_b200/csrc/device/shading.cuh"

using namespace ctld;

// out = weight[3], pdf, wo[3], sampledType, eta ; f[3], pdf(wo)   (identity shading frame), as orc_bsdf_probe
extern "C" void emu_bsdf_probe(const ctl_material* m, const float* wi, float sx, float sy, float* out9, float* f3, float* pdf1) {
    BRec b; memset(&b, 0, sizeof(b)); b.wi = mk(wi[0], wi[1], wi[2]); b.wo = mk(0, 0, 1); b.typeMask = E_ALL; b.eta = 1;
    float pdf = 0; Spec w = bsdf_sample(*m, b, pdf, sx, sy);
    out9[0] = w.r; out9[1] = w.g; out9[2] = w.b; out9[3] = pdf; out9[4] = b.wo.x; out9[5] = b.wo.y; out9[6] = b.wo.z; out9[7] = (float)b.sampledType; out9[8] = b.eta;
    BRec b2 = b; b2.typeMask = E_ALL & ~E_DELTA;
    Spec f = bsdf_f(*m, b2); f3[0] = f.r; f3[1] = f.g; f3[2] = f.b; *pdf1 = bsdf_pdf(*m, b2);
}
// out = value[3], pdf, p[3], d[3], dist, as orc_light_sample_direct
extern "C" void emu_light_sample_direct(const ctl_scene_view* v, uint32_t light, const float* ref, const float* refN, float sx, float sy, float* out11) {
    DScene S; memset(&S, 0, sizeof(S)); S.light_tris = v->light_tris; S.light_cdf_data = v->light_cdf_data; S.lights = v->lights;
    DRec d; d.ref = mk(ref[0], ref[1], ref[2]); d.refN = mk(refN[0], refN[1], refN[2]); d.pdf = 0;
    Spec val = light_sample_direct(S, v->lights[light], d, sx, sy);
    float o[11] = {val.r, val.g, val.b, d.pdf, d.p.x, d.p.y, d.p.z, d.d.x, d.d.y, d.d.z, d.dist};
    memcpy(out11, o, sizeof(o));
}
