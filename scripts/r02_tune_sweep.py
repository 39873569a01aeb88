"""Scheduler-parameter sweep of the staged traversal kernel on full frames (1920x1080, 8 passes fused, depth 8).
    python scripts/r02_tune_sweep.py c2 c4:1024     (CTL_B200_LIB selects a build variant; SWEEP_RESIDENT / SWEEP_THREADS its launch shape)"""
import json, sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudatracerlib_b200 as ctl

W, H, SPP, DEPTH = 1920, 1080, 8, 8
RES = int(os.environ.get("SWEEP_RESIDENT", "1024")); THREADS = int(os.environ.get("SWEEP_THREADS", "512"))
BASE = {"TravThT": 2, "TravThL": 8, "TravThF": 8, "TravThNExit": 4, "TravTSteps": 1}
VARIANTS = [{}, {"TravTSteps": 2}, {"TravTSteps": 3}, {"TravThNExit": 3}, {"TravThNExit": 6}, {"TravThNExit": 8}, {"TravThNExit": 6, "TravTSteps": 2}, {"TravThNExit": 8, "TravTSteps": 2},
            {"TravThT": 4}, {"TravThT": 4, "TravTSteps": 2}, {"TravThT": 6, "TravTSteps": 3}, {"TravThL": 4}, {"TravThL": 12}, {"TravThF": 4}, {"TravThF": 12}, {"TravThL": 4, "TravThF": 4},
            {"TravThT": 1}, {"TravThNExit": 2, "TravThT": 1}]
if os.environ.get("SWEEP_QUICK"):
    VARIANTS = VARIANTS[:3]


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
    t = ctl.PathTracer(W, H); t.setParameter("MaxPathLength", DEPTH); t.setParameter("TraversalKernel", 2)
    t.setParameter("StagedResidentThreads", RES); t.setParameter("StagedThreads", THREADS)
    t.InitializeScene(s)
    for v in VARIANTS:
        p = dict(BASE); p.update(v)
        for k, x in p.items():
            t.setParameter(k, x)
        ms, rays = measure(t)
        print(json.dumps({"workload": spec, "lib": os.path.basename(os.environ.get("CTL_B200_LIB", "default")), "resident": RES, "threads": THREADS, **p, "ms": round(ms, 3), "mrays_s": round(rays / ms / 1e3, 1)}), flush=True)
    t.close()
