// sampler_tables.h -- host generation of the per-pass SequenceSamplerData tables.
//
// Replaces SamplingSequenceGeneratorHost<IndependantSamplingSequenceGenerator>::Compute
// (Kernel/Sampler.h:22-85) and the host XORWOW twin (Base/CudaRandom.h:108-127, .cu:7-34).
// One stream per tracer: curand_init(1234, 7539414, 0); the state after that init is a constant
// (the test-suite re-derives it from the toolkit's jump matrices).
// Per pass and per sequence: 30 one-dimensional draws, then 30 two-dimensional ones whose FIRST
// draw is the y component (g++ evaluates Vec2f(rng.randomFloat(), rng.randomFloat()) right to left).
#pragma once
#include <cstdint>

namespace ctlb {

constexpr int kNumSeq = 4096, kSeqLen = 30; // Kernel/TraceHelper.cu:257

struct SamplerTableGenerator {
    uint32_t v[5], d;
    SamplerTableGenerator() { reset(); }
    void reset() {
        v[0] = 2779955570u; v[1] = 1996343557u; v[2] = 3815788579u; v[3] = 3068309824u; v[4] = 405030080u; d = 832094735u;
    }
    inline uint32_t next() {
        uint32_t t = (v[0] ^ (v[0] >> 2));
        v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
        v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
        d += 362437;
        return v[4] + d;
    }
    inline float random_float() {
        float f = next() * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
        return f * (1 - 1e-5f);
    }
    // element (seq, dim) at dim * kNumSeq + seq (Kernel/Sampler_device.h:20-24)
    void next_pass(float* d1, float* d2) {
        for (int s = 0; s < kNumSeq; s++) {
            for (int i = 0; i < kSeqLen; i++) d1[i * kNumSeq + s] = random_float();
            for (int i = 0; i < kSeqLen; i++) {
                float y = random_float(), x = random_float();
                d2[(i * kNumSeq + s) * 2 + 0] = x; d2[(i * kNumSeq + s) * 2 + 1] = y;
            }
        }
    }
};

} // namespace ctlb
