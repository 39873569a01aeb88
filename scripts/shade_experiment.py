import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
from bench import WORKLOADS
for wl in ("c2", "c3"):
    kind, w, h, spp, depth, _ = WORKLOADS[wl]
    s = Scene(kind, w, h)
    t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth); t.setParameter("StageTimers", 1)
    for bps in (4, 5, 6, 8, 12):
        t.setParameter("ShadeBlocksPerSM", bps)
        best = None
        for i in range(3):
            t.DoPasses(8, new_trace=True); t.synchronize(); ms, _ = t.stageTimes()
            if best is None or ms[2] < best[2]: best = ms
        print(os.environ.get("CTL_B200_LIB", "default"), wl, "shade blocks/SM", bps, "shade ms", round(best[2], 3), "ext", round(best[1], 2), "shadow", round(best[3], 2), flush=True)
    if "CTL_B200_LIB" not in os.environ:
        for co in (0, 25, 50, 75):
            t.setParameter("TravSmemCarveout", co)
            best = None
            for i in range(3):
                t.DoPasses(8, new_trace=True); t.synchronize(); ms, _ = t.stageTimes()
                if best is None or ms[1] < best[1]: best = ms
            print(wl, "traversal smem carveout %", co, "ext", round(best[1], 2), "shadow", round(best[3], 2), flush=True)
    t.close()
