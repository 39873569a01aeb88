// ctl_comm.cu -- the multi-GPU step of the path behind the C ABI: image tiles per rank, ONE NCCL reduce of the PixelData accumulator per frame.
//
// The reference is single-GPU (Kernel/TraceHelper.cu:744 synchronises one device; no NCCL / MPI anywhere in its tree), so there is no reference
// interface to mirror: this is the B200-native scale-out of Tracer<true>::DoPass (Kernel/Tracer.h:209-248).  Random numbers are a pure function of
// (pass, pixel index, dimension) (Kernel/Sampler_device.h:91-107), so rendering disjoint tiles on different devices changes no path, and adding the
// zero-initialised accumulators of the other ranks is exact.  One 58 MB reduce per frame needs no fused compute + collective kernel: ncclReduce on
// the context's stream, ordered after the frame's last kernel, is the whole exchange.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) and only when a communicator is asked for: a process that never calls ctl_comm_* never loads it,
// and a Python process that has imported torch gets torch's own copy (same soname).  Two ways to form the communicator:
//   * one process (or thread) per GPU:  rank 0 calls ctl_comm_get_unique_id and hands the 128 bytes to the others; everybody ctl_comm_init_rank;
//   * one process driving all GPUs:     ctl_comm_init_all over the contexts, and the *_all forms of the collectives (they group the per-device calls).
#include "ctl_internal.h"
#include <dlfcn.h>
#include <mutex>

namespace {
struct NcclUniqueId { char internal[128]; };                // ncclUniqueId (NCCL_UNIQUE_ID_BYTES 128)
enum { NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_SUM = 0 };   // ncclDataType_t / ncclRedOp_t values (stable since NCCL 2.0)
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm**, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm**, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};
Nccl g_nccl; std::once_flag g_once;

const Nccl* nccl() {
    std::call_once(g_once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { g_nccl.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
        if (!g_nccl.h) { g_nccl.why = std::string("NCCL is not available: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found"); return; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name); if (!g_nccl.field) g_nccl.why = std::string("NCCL symbol missing: ") + name;
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommInitAll, "ncclCommInitAll") SYM(CommDestroy, "ncclCommDestroy")
        SYM(Reduce, "ncclReduce") SYM(AllReduce, "ncclAllReduce") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    });
    return g_nccl.why.empty() ? &g_nccl : nullptr;
}
int nccl_fail(const char* what, int rc) { return ctl_set_err(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")"); }
#define NK(what, call) do { const int rc_ = (call); if (rc_ != 0) return nccl_fail(what, rc_); } while (0)
#define NEED_NCCL() const Nccl* N = nccl(); if (!N) return ctl_set_err(g_nccl.why)
} // namespace

void ctl_comm_release(ctl_ctx* c) {
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);   // reduces of frames in flight
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    c->comm = nullptr; c->comm_rank = 0; c->comm_size = 1;
    if (c->comm_scratch) cudaFree(c->comm_scratch);
    c->comm_scratch = nullptr;
}

// The reduce of a frame in flight: after the frame's last kernel (F.done), on the communication stream -- a stream of its own (highest priority) so that
// neither the lane that rendered the frame nor the context's stream waits for the other ranks; F.ready marks the accumulator final.
static int reduce_enqueue(ctl_ctx* c, FrameSlot& F) {
    NEED_NCCL();
    CK(cudaSetDevice(c->device));
    if (!c->comm_stream) { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi)); }
    CK(cudaStreamWaitEvent(c->comm_stream, F.done, 0));
    NK("ncclReduce", N->Reduce(F.accum.p, F.accum.p, (size_t)c->w * c->h * 7, NCCL_FLOAT32, NCCL_SUM, c->frame_root, c->comm, c->comm_stream));
    return 0;
}
static int reduce_mark(ctl_ctx* c, FrameSlot& F) {
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(F.ready, c->comm_stream));
    F.reduced = true;
    return 0;
}
int ctl_comm_reduce_slot(ctl_ctx* c, FrameSlot& F) {
    if (!c->comm) { if (c->comm_size != 1) return ctl_set_err("no communicator: ctl_comm_init_rank / ctl_comm_init_all first"); return 0; }
    if (c->comm_defer) { c->comm_pending = &F; return 0; }   // ctl_comm_submit_frame_all: one process, several devices -> the reduces are grouped by the caller
    return reduce_enqueue(c, F) || reduce_mark(c, F);
}

extern "C" {

int ctl_comm_get_unique_id(void* id_out) {
    if (!id_out) return ctl_set_err("null argument");
    NEED_NCCL();
    NcclUniqueId id; NK("ncclGetUniqueId", N->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int ctl_comm_init_rank(ctl_ctx* c, const void* id, int rank, int n_ranks) {
    if (!c || !id) return ctl_set_err("null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return ctl_set_err("invalid rank / world size");
    NEED_NCCL();
    CK(cudaSetDevice(c->device));
    ctl_comm_release(c);
    NcclUniqueId uid; memcpy(&uid, id, sizeof(uid));
    NK("ncclCommInitRank", N->CommInitRank(&c->comm, n_ranks, uid, rank));
    c->comm_rank = rank; c->comm_size = n_ranks;
    return 0;
}

int ctl_comm_init_all(ctl_ctx* const* ctxs, int n) {
    if (!ctxs || n < 1) return ctl_set_err("null / empty argument");
    NEED_NCCL();
    std::vector<int> devs(n); std::vector<ncclComm*> comms(n, nullptr);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return ctl_set_err("null context");
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; j++) if (devs[j] == devs[i]) return ctl_set_err("ctl_comm_init_all: two contexts on one device (use one context per GPU)");
        ctl_comm_release(ctxs[i]);
    }
    NK("ncclCommInitAll", N->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->comm_rank = i; ctxs[i]->comm_size = n; }
    return 0;
}

int ctl_comm_rank(const ctl_ctx* c, int* rank, int* n_ranks) {
    if (!c) return ctl_set_err("null context");
    if (rank) *rank = c->comm_rank;
    if (n_ranks) *n_ranks = c->comm_size;
    return 0;
}

// The ONE collective of the path: sum of the per-rank PixelData accumulators (7 * w * h floats) to `root`, on the context's stream, i.e. ordered after
// the frame's kernels; asynchronous.  In place on the root; the other ranks' accumulators are left as they are.
int ctl_comm_reduce_accum(ctl_ctx* c, int root) {
    if (!c) return ctl_set_err("null context");
    if (!c->comm) return c->comm_size == 1 ? 0 : ctl_set_err("no communicator: ctl_comm_init_rank / ctl_comm_init_all first");
    if (root < 0 || root >= c->comm_size) return ctl_set_err("root out of range");
    if (c->f_submitted != c->f_acquired) return ctl_set_err("ctl_comm_reduce_accum: frames are in flight (their reduces run on the communication stream; ctl_acquire_frame them first)");
    NEED_NCCL();
    CK(cudaSetDevice(c->device));
    NK("ncclReduce", N->Reduce(c->accum, c->accum, (size_t)c->w * c->h * 7, NCCL_FLOAT32, NCCL_SUM, root, c->comm, c->stream));
    return 0;
}
int ctl_comm_reduce_accum_all(ctl_ctx* const* ctxs, int n, int root) {
    if (!ctxs || n < 1) return ctl_set_err("null / empty argument");
    if (n == 1) return 0;
    NEED_NCCL();
    NK("ncclGroupStart", N->GroupStart());
    int rc = 0;
    for (int i = 0; i < n && !rc; i++) rc = ctl_comm_reduce_accum(ctxs[i], root);
    const int rg = N->GroupEnd();
    if (rc) return rc;
    NK("ncclGroupEnd", rg);
    return 0;
}

// Sum of `count` host counters over the ranks (ray counts of a frame); synchronous.
int ctl_comm_allreduce_u64(ctl_ctx* c, uint64_t* host_inout, int count) {
    if (!c || !host_inout || count < 1 || count > 64) return ctl_set_err("invalid argument");
    if (c->comm_size == 1) return 0;
    if (!c->comm) return ctl_set_err("no communicator");
    NEED_NCCL();
    CK(cudaSetDevice(c->device));
    if (!c->comm_scratch) CK(cudaMalloc((void**)&c->comm_scratch, 64 * sizeof(unsigned long long)));
    CK(cudaMemcpyAsync(c->comm_scratch, host_inout, count * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    NK("ncclAllReduce", N->AllReduce(c->comm_scratch, c->comm_scratch, (size_t)count, NCCL_UINT64, NCCL_SUM, c->comm, c->stream));
    CK(cudaMemcpyAsync(host_inout, c->comm_scratch, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// One progressive frame of `spp` passes shared by the communicator's ranks: this rank's interleaved tiles (tile index % ranks == rank), `batch` passes
// fused per wavefront, then the reduce to `root`.  Asynchronous on the context's stream.  == what bench.py's step does at N GPUs.
int ctl_comm_render_frame(ctl_ctx* c, int spp, int batch, int tile, int root) {
    if (!c) return ctl_set_err("null context");
    if (tile <= 0) tile = 64;
    if (ctl_render_frame_tiled(c, spp, batch, tile, tile, c->comm_rank, c->comm_size)) return 1;
    return ctl_comm_reduce_accum(c, root);
}

// Frames in flight across the ranks: ctl_submit_frame_tiled on this rank's tiles; the reduce of every frame runs on the context's communication stream
// (ctl_comm_reduce_slot), in submission order on every rank, while the lanes render the next frames.  ctl_acquire_frame returns them in order; on the
// root the acquired accumulator holds the whole image.
int ctl_comm_submit_frame(ctl_ctx* c, int spp, int batch, int tile, int root) {
    if (!c) return ctl_set_err("null context");
    if (tile <= 0) tile = 64;
    if (root < 0 || root >= c->comm_size) return ctl_set_err("root out of range");
    c->frame_root = root;
    return ctl_submit_frame_tiled(c, spp, batch, tile, tile, c->comm_rank, c->comm_size);
}

int ctl_comm_submit_frame_all(ctl_ctx* const* ctxs, int n, int spp, int batch, int tile, int root) {   // the same for ctl_comm_init_all communicators
    if (!ctxs || n < 1) return ctl_set_err("null / empty argument");
    int rc = 0;
    for (int i = 0; i < n && !rc; i++) {
        if (!ctxs[i]) return ctl_set_err("null context");
        ctxs[i]->comm_defer = true; ctxs[i]->comm_pending = nullptr;
        rc = ctl_comm_submit_frame(ctxs[i], spp, batch, tile, root);
        ctxs[i]->comm_defer = false;
    }
    if (rc || n == 1 || !ctxs[0]->comm) return rc;
    NEED_NCCL();
    NK("ncclGroupStart", N->GroupStart());
    for (int i = 0; i < n && !rc; i++) rc = ctxs[i]->comm_pending ? reduce_enqueue(ctxs[i], *ctxs[i]->comm_pending) : ctl_set_err("ctl_comm_submit_frame_all: context without a communicator");
    const int rg = N->GroupEnd();
    if (rc) return rc;
    NK("ncclGroupEnd", rg);
    for (int i = 0; i < n && !rc; i++) { rc = reduce_mark(ctxs[i], *ctxs[i]->comm_pending); ctxs[i]->comm_pending = nullptr; }
    return rc;
}

int ctl_comm_destroy(ctl_ctx* c) {
    if (!c) return ctl_set_err("null context");
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    ctl_comm_release(c);
    return 0;
}

} // extern "C"
