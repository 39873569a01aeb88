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
#include <cstddef>
#include <thread>
#include <vector>

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

// ---- data for the DEVICE-side generator (k_gen_tables in wavefront.cuh) ---------------------------------------
// The XORWOW state update is linear over GF(2) on the 160 bits of v[0..4] (the Weyl counter d just adds 362437 per
// draw), so "skip n draws" is a 160x160 bit matrix.  Sequence s of a pass starts 90*s draws into the pass and the next
// pass starts 4096*90 draws later; one thread per sequence therefore needs (a) its start state in pass 0 and (b) the
// jump matrix for 4096*90 - 90 draws to get from the end of its slice to its slice in the next pass.
struct XorwowJump {
    // row k = image of basis vector e_k (bit k of the 160-bit state, word k/32, bit k%32)
    uint32_t row[160][5];
    static void step_linear(uint32_t v[5]) {
        uint32_t t = (v[0] ^ (v[0] >> 2));
        v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
        v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
    }
    void set_one_step() {
        for (int k = 0; k < 160; k++) { uint32_t v[5] = {0, 0, 0, 0, 0}; v[k / 32] = 1u << (k % 32); step_linear(v); for (int w = 0; w < 5; w++) row[k][w] = v[w]; }
    }
    void apply(const uint32_t in[5], uint32_t out[5]) const {
        uint32_t r[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < 160; k++) if ((in[k / 32] >> (k % 32)) & 1u) for (int w = 0; w < 5; w++) r[w] ^= row[k][w];
        for (int w = 0; w < 5; w++) out[w] = r[w];
    }
    // this = this followed by other
    void then(const XorwowJump& other) { for (int k = 0; k < 160; k++) other.apply(row[k], row[k]); }
    static XorwowJump power(unsigned long long n) {
        XorwowJump result; for (int k = 0; k < 160; k++) for (int w = 0; w < 5; w++) result.row[k][w] = (w == k / 32) ? (1u << (k % 32)) : 0u; // identity
        XorwowJump base; base.set_one_step();
        while (n) { if (n & 1ull) result.then(base); XorwowJump sq = base; sq.then(base); base = sq; n >>= 1; }
        return result;
    }
};
constexpr int kDrawsPerSeq = 3 * kSeqLen;                 // 30 one-dimensional + 30 two-dimensional draws
constexpr int kDrawsPerPass = kNumSeq * kDrawsPerSeq;
// states0: kNumSeq x 6 words (v[5], d) = state of the stream right before sequence s of pass 0
inline void device_generator_data(uint32_t* states0, XorwowJump* jump_to_next_pass) {
    SamplerTableGenerator g;
    for (int s = 0; s < kNumSeq; s++) {
        for (int w = 0; w < 5; w++) states0[s * 6 + w] = g.v[w];
        states0[s * 6 + 5] = g.d;
        for (int i = 0; i < kDrawsPerSeq; i++) g.next();
    }
    *jump_to_next_pass = XorwowJump::power((unsigned long long)(kDrawsPerPass - kDrawsPerSeq));
}

// ---- several passes at once on the host ------------------------------------------------------------------------
// The passes of a frame are consecutive slices of ONE stream, but the start state of pass p+1 is the start state of pass p jumped kDrawsPerPass draws
// ahead (the same GF(2) matrix trick as the device generator, plus the Weyl counter's closed form), so the n passes of a frame can be produced by n
// threads, each running the plain sequential generator over its own pass: bit-identical tables, 1/n of the 1.3 ms per pass of host time per frame.
// `g` is the generator at the start of the first pass and is left at the start of the pass after the last one, as n next_pass calls would leave it.
inline const XorwowJump& pass_jump() { static const XorwowJump J = XorwowJump::power((unsigned long long)kDrawsPerPass); return J; }
template <typename Launch>   // Launch(fn): runs fn(k) for k = 0 .. n-1, possibly concurrently (std::thread in the library, a plain loop in tests)
inline void generate_passes(SamplerTableGenerator& g, int n, float* d1, float* d2, size_t stride1, size_t stride2, Launch&& launch) {
    if (n <= 0) return;
    SamplerTableGenerator* start = new SamplerTableGenerator[n + 1];
    start[0] = g;
    for (int p = 1; p <= n; p++) {
        start[p] = start[p - 1];
        pass_jump().apply(start[p - 1].v, start[p].v);
        start[p].d = start[p - 1].d + 362437u * (uint32_t)kDrawsPerPass;
    }
    launch([&](int k) { SamplerTableGenerator local = start[k]; local.next_pass(d1 + stride1 * k, d2 + stride2 * k); });
    g = start[n];
    delete[] start;
}

// The library's launcher: one thread per pass, at most 8 and at most the host's cores at a time (a frame of 8 passes: 10.3 ms -> ~1.5 ms of host time).
inline void generate_passes_threaded(SamplerTableGenerator& g, int n, float* d1, float* d2, size_t stride1, size_t stride2) {
    generate_passes(g, n, d1, d2, stride1, stride2, [n](auto&& fn) {
        unsigned hw = std::thread::hardware_concurrency(); if (hw == 0) hw = 1;
        const int width = n < 2 ? 1 : (int)(hw < 8u ? hw : 8u);
        if (width <= 1) { for (int k = 0; k < n; k++) fn(k); return; }
        for (int k0 = 0; k0 < n; k0 += width) {
            std::vector<std::thread> th;
            const int k1 = k0 + width < n ? k0 + width : n;
            for (int k = k0 + 1; k < k1; k++) th.emplace_back([&fn, k] { fn(k); });
            fn(k0);
            for (auto& t : th) t.join();
        }
    });
}

} // namespace ctlb
