import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
from bench import WORKLOADS
for wl in (sys.argv[1:] or ["c2"]):
    kind, w, h, spp, depth, _ = WORKLOADS[wl]
    s = Scene(kind, w, h)
    t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth); t.setParameter("StageTimers", 1)
    for bps in (8, 10, 12):
        t.setParameter("TraversalBlocksPerSM", bps)
        best = None
        for i in range(3):
            t.DoPasses(8 if wl != "c4" else 2, new_trace=True); t.synchronize(); ms, _ = t.stageTimes()
            if best is None or ms[1] + ms[3] < best[1] + best[3]: best = ms
        print(os.environ.get("CTL_B200_LIB", "default(hints)"), wl, "trav blocks/SM", bps, "ext", round(best[1], 2), "shadow", round(best[3], 2), "shade", round(best[2], 2), flush=True)
    t.close()
