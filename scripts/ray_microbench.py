"""Ray-level micro-benchmark of the traversal kernel: rays captured from bounce 0 (camera rays, coherent), bounce 2 and bounce 5
(incoherent) of a workload, up to 2^21 each, as 32-byte traversalRay records; ctl_intersect on device buffers, closest hit and any hit, timed with
CUDA events on the launch stream.  Algorithmic bytes per ray from the instrumented traversal of the same rays (ctl_trace_rays_host counts:
48 + 64 inner + 52 tris + 108 instances), reported against the HBM peak bench.py uses.
    python scripts/ray_microbench.py c2 > gpurun_out/ray_microbench_c2.json        (GPU; one JSON line per bounce)
Results: profiles/r02a_ray_microbench_*.json (round-2 start), profiles/r03a_ray_microbench_*.json (shipping kernel and trees)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cudatracerlib_b200 import Scene, PathTracer, api
from bench import WORKLOADS, hbm_peak

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
N_MAX = 1 << 21
kind, w, h, spp, depth, _ = WORKLOADS[wl]
s = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", max(depth, 8))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); t.setStream(st.cuda_stream)
peak, peak_src = hbm_peak()


def time_rays(d_rays, d_res, n, any_hit, reps=10):
    for _ in range(3):
        t.intersect_device(n, d_rays.data_ptr(), d_res.data_ptr(), any_hit, st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        t.intersect_device(n, d_rays.data_ptr(), d_res.data_ptr(), any_hit, st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for bounce in (0, 2, 5):
    t.setParameter("CaptureBounce", bounce + 1)             # 1-based
    t.DoPass(True); t.synchronize()
    rays = t.capturedRays(w * h)[:N_MAX]
    t.setParameter("CaptureBounce", 0)
    n = len(rays)
    if n == 0:
        continue
    _, counts = t.trace_rays(rays[: 1 << 18], counts=True)  # visit counts on a 2^18 sample of the same queue (instrumented kernel, slower)
    bytes_per_ray = api.traversal_bytes(counts, min(n, 1 << 18)) / min(n, 1 << 18)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_res = torch.zeros(n, 4, dtype=torch.int32, device="cuda")
    ms_closest, ms_any = time_rays(d_rays, d_res, n, False), time_rays(d_rays, d_res, n, True)
    gbs = bytes_per_ray * n / (ms_closest * 1e-3) / 1e9
    print(json.dumps({"workload": wl, "bounce": bounce, "rays": n, "inner_per_ray": counts[0] / min(n, 1 << 18), "tris_per_ray": counts[1] / min(n, 1 << 18),
                      "bytes_per_ray": bytes_per_ray, "closest_hit": {"ms": ms_closest, "mrays_s": n / ms_closest / 1e3, "algorithmic_gb_s": gbs, "frac_of_peak": gbs / peak},
                      "any_hit": {"ms": ms_any, "mrays_s": n / ms_any / 1e3}, "peak_gb_s": peak, "peak_source": peak_src,
                      "note": "tmin/tmax of the captured records are honoured (intersectKernel semantics); inputs (64 MB of rays at 2^21) exceed nothing: L2-resident, as in the render"}), flush=True)
t.close()
