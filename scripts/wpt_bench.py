"""WavefrontPathTracer timing on one GPU: Mrays/s per pass, stage split, vs the PathTracer wavefront on the same scene."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudatracerlib_b200 as ctl

kind = sys.argv[1] if len(sys.argv) > 1 else "c2"
w, h, mpl, spp = 1920, 1080, 8, 8
s = ctl.Scene(kind, w, h)
out = {"scene": kind, "w": w, "h": h, "max_path_length": mpl, "spp": spp}
t = ctl.WavefrontPathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", mpl)
for rep in range(3):   # warm-up frames
    for p in range(spp):
        t.DoPass(p == 0)
t.synchronize()
import torch
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.Stream(); t.setStream(st.cuda_stream)
frames = 5
with torch.cuda.stream(st):
    e0.record(st)
    for f in range(frames):
        for p in range(spp):
            t.DoPass(p == 0)
    e1.record(st)
st.synchronize()
ms = e0.elapsed_time(e1) / frames
rays = t.getTotalRays() // (frames + 3)
out["wavefront_pt"] = {"ms_per_frame": ms, "rays_per_frame": int(rays), "mrays_s": rays / ms / 1e3}
# the same frames through ctl_wavefront_frame: the passes of a frame on 1 .. 8 lanes (streams with their own queue buffers)
out["wavefront_pt_frame"] = {}
for lanes in (1, 2, 4, 8):
    t.setParameter("OverlapLanes", lanes)
    for rep in range(2): t.DoFrame(spp)
    t.synchronize(); r0 = t.getTotalRays()
    with torch.cuda.stream(st):
        e0.record(st)
        for f in range(frames): t.DoFrame(spp)
        e1.record(st)
    st.synchronize()
    msf = e0.elapsed_time(e1) / frames; rf = (t.getTotalRays() - r0) // frames
    out["wavefront_pt_frame"][str(lanes)] = {"ms_per_frame": msf, "mrays_s": rf / msf / 1e3}
t.setParameter("OverlapLanes", 4)
t.setParameter("StageTimers", 1); t.DoPass(True); t.synchronize()
sm, nl = t.stageTimes(); out["wavefront_pt"]["stage_ms_one_pass"] = dict(zip(["create", "primary_trav", "iterate", "secondary_trav", "tally"], [round(x, 3) for x in sm])); out["wavefront_pt"]["launches_per_pass"] = nl
e, sh = t.queueSizes(mpl); out["wavefront_pt"]["queues"] = [e.tolist(), sh.tolist()]
# roofline of the traversal launches of one pass: algorithmic bytes from an instrumented pass / CUDA-event time of the fused launches
trav_ms = sm[1] + sm[3]
t.setParameter("StageTimers", 0); t.setInstrumented(1); t.DoPass(True); t.synchronize()
e_cnt, s_cnt = t.visitCounts(); t.setInstrumented(0)
nbytes = ctl.traversal_bytes(e_cnt, e_cnt[3]) + ctl.traversal_bytes(s_cnt, s_cnt[3])
from bench import hbm_peak
peak, peak_src = hbm_peak()
out["wavefront_pt"]["roofline"] = {"bound": "hbm", "kernel": "k_intersect / k_intersect_fused_api (all traversal launches of a pass)", "achieved": nbytes / (trav_ms * 1e-3) / 1e9, "peak": peak,
                                   "peak_source": peak_src, "unit": "GB/s", "frac": nbytes / (trav_ms * 1e-3) / 1e9 / peak, "bytes_per_ray": nbytes / max(1, e_cnt[3] + s_cnt[3]), "traversal_ms_per_pass": trav_ms,
                                   "rays_per_pass": int(e_cnt[3] + s_cnt[3])}
t.close()
p = ctl.PathTracer(w, h); p.InitializeScene(s); p.setParameter("MaxPathLength", mpl); p.setStream(st.cuda_stream)
for rep in range(3):
    p.DoPasses(spp, new_trace=True)
p.synchronize()
with torch.cuda.stream(st):
    e0.record(st)
    for f in range(frames):
        p.DoPasses(spp, new_trace=True)
    e1.record(st)
st.synchronize()
ms = e0.elapsed_time(e1) / frames
out["path_tracer"] = {"ms_per_frame": ms, "rays_per_frame": int(p.getRaysInLastPass()), "mrays_s": p.getRaysInLastPass() / ms / 1e3}
p.close()
# the reference's own WavefrontPathTracer code on the host cores (oracle/_ref: its pathIterateKernel + DoubleRayBuffer on one thread, the
# intersections on all threads) on a bounded sample: the same scene at 1/16 of the pixels, one pass
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import time
    import ref_binding as rb
    if rb.available():
        sw, sh_ = w // 4, h // 4
        s_small = ctl.Scene(kind, sw, sh_)
        t0 = time.perf_counter(); _, crays, _ = rb.render_wavefront(s_small.view, sw, sh_, n_passes=4, max_path_length=mpl); dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": crays / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": f"{sw}x{sh_} frame of the same scene, 4 passes, depth {mpl}, {crays} rays in {dt:.2f} s (queue kernels on one thread, intersections on all)"}
except Exception as e:   # the baseline is optional here
    out["cpu_baseline"] = {"unavailable": str(e)}
print(json.dumps(out))
