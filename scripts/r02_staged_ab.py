"""A/B of the traversal kernels on full frames (1920x1080, 8 passes fused, depth 8): persistent (TraversalKernel=0) against the staged kernel
(TraversalKernel=2) over launch shapes (StagedThreads, StagedStackRows, StagedTreeletNodes).  One scene build per workload; CUDA-event frame times.
    python scripts/r02_staged_ab.py c2 c4 c4:1024 > gpurun_out/r02b_staged_ab.log"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudatracerlib_b200 as ctl

W, H, SPP, DEPTH = 1920, 1080, 8, 8
SHAPES = [(128, 16, 0), (256, 16, 0), (512, 16, 0), (1024, 16, 0), (512, 8, 0), (512, 24, 0), (512, 32, 0), (256, 16, 128), (512, 16, 256), (512, 16, 512), (512, 12, 512),
          (1024, 16, 1024), (1024, 12, 1024), (1024, 16, 2048)]
if os.environ.get("AB_SHAPES"):
    SHAPES = [tuple(int(x) for x in s.split(",")) for s in os.environ["AB_SHAPES"].split(";")]


def measure(t, n_warm=2, n_timed=4):
    for _ in range(n_warm):
        t.DoPasses(SPP, new_trace=True)
    t.synchronize()
    ms, rays = [], 0
    for _ in range(n_timed):
        t.DoPasses(SPP, new_trace=True)
        ms.append(1e3 * t.getLastTimeSpentRenderingSec()); rays = t.getRaysInLastPass()
    return float(np.median(ms)), rays


for spec in sys.argv[1:] or ["c2"]:
    wl, _, rb = spec.partition(":")
    s = ctl.Scene(wl, W, H)
    s.setRebraid(int(rb) if rb else 0)
    t = ctl.PathTracer(W, H); t.setParameter("MaxPathLength", DEPTH); t.setParameter("TraversalKernel", 0); t.InitializeScene(s)
    ms0, rays = measure(t)
    img0 = t.readAccumulator()["rgb"].mean()
    print(json.dumps({"workload": spec, "kernel": "persistent", "ms": round(ms0, 3), "mrays_s": round(rays / ms0 / 1e3, 1)}), flush=True)
    for th, rows, tl in SHAPES:
        t.setParameter("TraversalKernel", 2); t.setParameter("StagedThreads", th); t.setParameter("StagedStackRows", rows); t.setParameter("StagedTreeletNodes", tl)
        t.InitializeScene(s)
        try:
            ms, r2 = measure(t)
            ok = bool(r2 == rays and abs(t.readAccumulator()["rgb"].mean() - img0) <= 1e-6 * abs(img0))
            print(json.dumps({"workload": spec, "kernel": "staged", "threads": th, "rows": rows, "treelet": t.getParameter("StagedTreeletNodes"), "ms": round(ms, 3),
                              "mrays_s": round(r2 / ms / 1e3, 1), "speedup": round(ms0 / ms, 3), "same_result": ok}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"workload": spec, "kernel": "staged", "threads": th, "rows": rows, "treelet": tl, "error": str(e)}), flush=True)
    t.close()
