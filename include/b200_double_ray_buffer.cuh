// b200_double_ray_buffer.cuh -- header-only CUDA/C++ counterpart of the reference's DoubleRayBuffer<T> (Kernel/DoubleRayBuffer.h) on the B200
// backend, for applications that drive their own payload kernels over it (WavefrontPathTracer, FastTracer / PT_Wave style integrators):
//
//     reference                                                     here
//     ---------                                                     ----
//     DoubleRayBuffer<T> buf(nPayload, nSecondary);                 ctlb200::DoubleRayBuffer<T> buf(nPayload, nSecondary);
//     buf.StartFrame(eps);                                          buf.StartFrame(eps);
//     CopyToSymbol(g_buf, buf); kernel<<<>>>(); CopyFromSymbol      kernel<<<..., stream>>>(buf.device(), ...);      // POD view by value, no symbol copies
//     g_buf->insertPayloadElement(p, ray) / tryFetch... (device)    same names on the view (traversalRay / traversalResult records)
//     buf.FinishIteration(skip_outer, any_hit_secondary);           buf.FinishIteration(ctx, any_hit_secondary, stream);  // -> ctl_intersect x2
//     buf.isEmpty(), buf.getNumPayloadElementsInQueue()             same
//
// What FinishIteration does is the reference's (DoubleRayBuffer.h:84-112): intersect the primary rays [0, insert index) and the secondary rays
// pushed during the iteration (closest hit, or any hit when asked), make the inserted payloads the next iteration's fetch range, swap the
// secondary buffers.  __internal__IntersectBuffers (Kernel/TraceHelper.cu:736-746) is ctl_intersect: same 32-byte ray / 16-byte result records,
// device pointers, no copies.  Differences, all deliberate: the payload / primary-ray arrays are ping-pong pairs (the reference inserts into the
// array it is still fetching from, which is only safe while no fetched slot is still being read); the three queue counters live on the device
// and are read back once per iteration (12 bytes) instead of copying the whole object to and from a device symbol around every kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include "ctl_b200.h"

namespace ctlb200 {

template <typename T> class DoubleRayBuffer {
public:
    // what kernels see (pass by value): the reference's device-side methods (DoubleRayBuffer.h:123-178)
    struct Device {
        const T* fetch_payload; const ctl_traversal_ray* fetch_ray; const ctl_traversal_result* fetch_res;
        T* insert_payload; ctl_traversal_ray* insert_ray;
        const ctl_traversal_ray* sec_prev_ray; const ctl_traversal_result* sec_prev_res;   // secondary rays of the previous iteration + results
        ctl_traversal_ray* sec_ray;                                                        // secondary rays pushed by this iteration
        unsigned* counters;   // [0] fetch index, [1] payload insert index, [2] secondary insert index
        unsigned num_payload_elements, payload_length, num_secondary_rays;
        float ray_eps;

        __device__ bool tryFetchPayloadElement(T& payload_el, ctl_traversal_ray& ray, ctl_traversal_result& res, unsigned* idx = nullptr) const {
            const unsigned i = atomicInc(&counters[0], 0xffffffffu);
            if (i >= num_payload_elements) return false;
            if (idx) *idx = i;
            payload_el = fetch_payload[i]; ray = fetch_ray[i]; res = fetch_res[i];
            return true;
        }
        __device__ bool insertPayloadElement(const T& payload_el, const ctl_traversal_ray& ray, unsigned* idx = nullptr) const {
            const unsigned i = atomicInc(&counters[1], 0xffffffffu);
            if (i >= payload_length) return false;
            if (idx) *idx = i;
            insert_payload[i] = payload_el; insert_ray[i] = ray;
            return true;
        }
        __device__ bool accessSecondaryRay(unsigned idx, ctl_traversal_ray& ray, ctl_traversal_result& res) const {
            if (idx >= num_secondary_rays) return false;
            ray = sec_prev_ray[idx]; res = sec_prev_res[idx];
            return true;
        }
        __device__ bool insertSecondaryRay(const ctl_traversal_ray& ray, unsigned& idx) const {
            idx = atomicInc(&counters[2], 0xffffffffu);
            if (idx >= num_secondary_rays) return false;
            sec_ray[idx] = ray;
            return true;
        }
        // convert(ray) of the reference (DoubleRayBuffer.h:234-237): origin + rayTraceEps, direction + FLT_MAX
        __device__ ctl_traversal_ray makeRay(float ox, float oy, float oz, float dx, float dy, float dz) const {
            ctl_traversal_ray r; r.o[0] = ox; r.o[1] = oy; r.o[2] = oz; r.tmin = ray_eps; r.d[0] = dx; r.d[1] = dy; r.d[2] = dz; r.tmax = 3.402823466e+38f; return r;
        }
    };

    DoubleRayBuffer(unsigned payload_length, unsigned secondary_length) : payload_length_(payload_length), num_secondary_(secondary_length) {
        for (int k = 0; k < 2; k++) {
            check(cudaMalloc((void**)&payload_[k], sizeof(T) * (size_t)payload_length_));
            check(cudaMalloc((void**)&ray_[k], sizeof(ctl_traversal_ray) * (size_t)payload_length_));
            check(cudaMalloc((void**)&sec_ray_[k], sizeof(ctl_traversal_ray) * (size_t)num_secondary_));
            check(cudaMalloc((void**)&sec_res_[k], sizeof(ctl_traversal_result) * (size_t)num_secondary_));
        }
        check(cudaMalloc((void**)&res_, sizeof(ctl_traversal_result) * (size_t)payload_length_));
        check(cudaMalloc((void**)&counters_, 3 * sizeof(unsigned)));
    }
    DoubleRayBuffer(const DoubleRayBuffer&) = delete; DoubleRayBuffer& operator=(const DoubleRayBuffer&) = delete;
    ~DoubleRayBuffer() { Free(); }
    void Free() {
        for (int k = 0; k < 2; k++) { cudaFree(payload_[k]); cudaFree(ray_[k]); cudaFree(sec_ray_[k]); cudaFree(sec_res_[k]); payload_[k] = nullptr; ray_[k] = nullptr; sec_ray_[k] = nullptr; sec_res_[k] = nullptr; }
        cudaFree(res_); cudaFree(counters_); res_ = nullptr; counters_ = nullptr;
    }

    void StartFrame(float rayTraceEps, cudaStream_t stream = nullptr) {
        eps_ = rayTraceEps; num_elements_ = 0; cur_ = 0; sec_cur_ = 0;
        check(cudaMemsetAsync(counters_, 0, 3 * sizeof(unsigned), stream));
    }
    // DoubleRayBuffer::FinishIteration<true> (DoubleRayBuffer.h:84-112); synchronises `stream` once to read the three counters
    void FinishIteration(ctl_ctx* ctx, bool any_hit_secondary = false, cudaStream_t stream = nullptr) {
        unsigned h[3];
        check(cudaMemcpyAsync(h, counters_, sizeof(h), cudaMemcpyDeviceToHost, stream));
        check(cudaStreamSynchronize(stream));
        if (h[1] > payload_length_) throw std::runtime_error("Storing too many primary rays in buffer!");
        if (h[2] > num_secondary_) throw std::runtime_error("Storing too many secondary rays in buffer!");
        const int ins = cur_ ^ 1, sec_ins = sec_cur_ ^ 1;   // arrays the iteration inserted into
        void* s = stream ? (void*)stream : (void*)cudaStreamLegacy; // NULL would mean "the context's own stream" to ctl_intersect: stay ordered with the caller's kernels
        if (h[1] && ctl_intersect(ctx, (int)h[1], ray_[ins], res_, 0, s)) throw std::runtime_error(ctl_last_error());
        if (h[2] && ctl_intersect(ctx, (int)h[2], sec_ray_[sec_ins], sec_res_[sec_ins], any_hit_secondary ? 1 : 0, s)) throw std::runtime_error(ctl_last_error());
        num_elements_ = h[1]; cur_ = ins; sec_cur_ = sec_ins;
        check(cudaMemsetAsync(counters_, 0, 3 * sizeof(unsigned), stream));
    }
    // "will the buffer be empty in the next iteration" (DoubleRayBuffer.h:115-118): no payload inserted since the last FinishIteration.  Reads the
    // device counter, i.e. waits for the kernels queued on `stream` -- what the reference's CopyFromSymbol after each kernel does
    bool isEmpty(cudaStream_t stream = nullptr) const {
        unsigned n = 0;
        check(cudaMemcpyAsync(&n, counters_ + 1, sizeof(n), cudaMemcpyDeviceToHost, stream));
        check(cudaStreamSynchronize(stream));
        return n == 0;
    }
    unsigned getNumPayloadElementsInQueue() const { return num_elements_; }

    Device device() const {
        Device d;
        d.fetch_payload = payload_[cur_]; d.fetch_ray = ray_[cur_]; d.fetch_res = res_;
        d.insert_payload = payload_[cur_ ^ 1]; d.insert_ray = ray_[cur_ ^ 1];
        d.sec_prev_ray = sec_ray_[sec_cur_]; d.sec_prev_res = sec_res_[sec_cur_]; d.sec_ray = sec_ray_[sec_cur_ ^ 1];
        d.counters = counters_; d.num_payload_elements = num_elements_; d.payload_length = payload_length_; d.num_secondary_rays = num_secondary_; d.ray_eps = eps_;
        return d;
    }

private:
    static void check(cudaError_t e) { if (e != cudaSuccess) throw std::runtime_error(std::string("In file b200_double_ray_buffer.cuh : ") + cudaGetErrorString(e)); }
    T* payload_[2] = {nullptr, nullptr}; ctl_traversal_ray* ray_[2] = {nullptr, nullptr}; ctl_traversal_result* res_ = nullptr;
    ctl_traversal_ray* sec_ray_[2] = {nullptr, nullptr}; ctl_traversal_result* sec_res_[2] = {nullptr, nullptr};
    unsigned* counters_ = nullptr;
    unsigned payload_length_, num_secondary_, num_elements_ = 0;
    int cur_ = 0, sec_cur_ = 0; float eps_ = 0.0f;
};

} // namespace ctlb200
