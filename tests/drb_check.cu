// drb_check.cu -- a miniature DoubleRayBuffer application (compiled by tests/test_gpu_wavefront_pt.py with nvcc for sm_100a): camera-like primary
// rays per pixel, one "bounce" that pushes a new primary ray and a secondary (shadow-like) ray per hit, results written per pixel so that the
// nondeterministic queue order does not matter.  Prints per-pixel checksums that the test compares with ctl_intersect_host on the same rays.
#include "b200_double_ray_buffer.cuh"
#include <cstdio>
#include <vector>

struct Payload { int pixel; float throughput; unsigned sec_idx; };
using Buffer = ctlb200::DoubleRayBuffer<Payload>;

__global__ void create(Buffer::Device B, int w, int h, float ox, float oy, float oz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const int x = i % w, y = i / w;
    float dx = (x + 0.5f) / w - 0.5f, dy = (y + 0.5f) / h - 0.5f, dz = 1.0f;
    const float il = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz); dx *= il; dy *= il; dz *= il;   // compiled with -fmad=false: reproducible on the host
    Payload p; p.pixel = i; p.throughput = 1.0f; p.sec_idx = 0xffffffffu;
    B.insertPayloadElement(p, B.makeRay(ox, oy, oz, dx, dy, dz));
}

// out[pixel * 4 + k]: k = 0 first-hit distance, 1 first-hit triangle, 2 second-hit distance, 3 secondary-ray distance
__global__ void iterate(Buffer::Device B, int depth, float* out, float lx, float ly, float lz) {
    Payload p; ctl_traversal_ray ray; ctl_traversal_result res;
    while (B.tryFetchPayloadElement(p, ray, res)) {
        if (depth == 0) { out[p.pixel * 4 + 0] = res.dist; out[p.pixel * 4 + 1] = (float)res.tri_idx; }
        if (depth == 1) {
            out[p.pixel * 4 + 2] = res.dist;
            ctl_traversal_ray sr; ctl_traversal_result sres;
            if (p.sec_idx != 0xffffffffu && B.accessSecondaryRay(p.sec_idx, sr, sres)) out[p.pixel * 4 + 3] = sres.dist;
        }
        if (depth == 0 && res.tri_idx >= 0) {
            const float px = ray.o[0] + ray.d[0] * res.dist, py = ray.o[1] + ray.d[1] * res.dist, pz = ray.o[2] + ray.d[2] * res.dist;
            float sx = lx - px, sy = ly - py, sz = lz - pz; const float il = 1.0f / sqrtf(sx * sx + sy * sy + sz * sz); sx *= il; sy *= il; sz *= il;
            unsigned k = 0xffffffffu;
            if (!B.insertSecondaryRay(B.makeRay(px, py, pz, sx, sy, sz), k)) k = 0xffffffffu;
            p.sec_idx = k; p.throughput *= 0.5f;
            B.insertPayloadElement(p, B.makeRay(px, py, pz, -ray.d[0], ray.d[1], -ray.d[2]));   // some deterministic "bounce"
        }
    }
}

int main(int argc, char** argv) {
    const int w = 96, h = 64, kind = argc > 1 ? atoi(argv[1]) : 1;
    ctl_scene* scene = ctl_scene_create(kind, w, h, 1234, 0);
    ctl_scene_view view; ctl_scene_get_view(scene, &view);
    ctl_ctx* ctx = ctl_create(0, w, h);
    if (!ctx || ctl_upload_scene(ctx, &view)) { fprintf(stderr, "%s\n", ctl_last_error()); return 1; }
    const float cx = 0.5f * (view.box_min[0] + view.box_max[0]), cy = 0.5f * (view.box_min[1] + view.box_max[1]), cz = view.box_min[2] + 0.02f * (view.box_max[2] - view.box_min[2]);
    const float lx = cx, ly = view.box_max[1] - 0.05f * (view.box_max[1] - view.box_min[1]), lz = 0.5f * (view.box_min[2] + view.box_max[2]);
    Buffer buf(w * h, w * h);
    float* d_out; cudaMalloc(&d_out, sizeof(float) * w * h * 4); cudaMemset(d_out, 0, sizeof(float) * w * h * 4);
    buf.StartFrame(view.ray_eps);
    create<<<(w * h + 127) / 128, 128>>>(buf.device(), w, h, cx, cy, cz);
    int depth = 0; unsigned sizes[3] = {0, 0, 0};
    do {
        buf.FinishIteration(ctx);
        sizes[depth] = buf.getNumPayloadElementsInQueue();
        iterate<<<148, 128>>>(buf.device(), depth, d_out, lx, ly, lz);
    } while (++depth < 3 && !buf.isEmpty());
    buf.FinishIteration(ctx);
    std::vector<float> out((size_t)w * h * 4);
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    printf("%u %u %u %d\n", sizes[0], sizes[1], sizes[2], (int)buf.isEmpty());
    printf("%.9g %.9g %.9g %.9g %.9g %.9g\n", cx, cy, cz, lx, ly, lz);
    for (size_t i = 0; i < out.size(); i++) printf("%.9g%c", out[i], (i % 4 == 3) ? '\n' : ' ');
    cudaFree(d_out); ctl_destroy(ctx); ctl_scene_destroy(scene);
    return 0;
}
