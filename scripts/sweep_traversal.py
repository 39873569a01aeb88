"""Threshold sweep of the persistent traversal kernel on the GPU (c2 by default)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, w, h, spp, depth, _ = WORKLOADS[wl]
s = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth); t.setParameter("StageTimers", 1)
def run():
    best = None
    for i in range(3):
        t.DoPass(i == 0); t.synchronize()
        ms, _ = t.stageTimes()
        if best is None or ms[1] + ms[3] < best[0]: best = (ms[1] + ms[3], ms[1], ms[3])
    return best
t.setParameter("TraversalKernel", 1); print("simple", run())
t.setParameter("TraversalKernel", 0)
res = []
grid = ((1, 2, 4), (4, 8, 16), (4, 8, 16), (2, 4, 6, 8)) if len(sys.argv) > 2 and sys.argv[2] == "wide" else ((2, 4, 8), (4, 8), (8, 12, 16), (3, 4, 6, 8, 12))
for thT, thL, thF, ns in itertools.product(*grid):
    for k, v in (("TravThT", thT), ("TravThL", thL), ("TravThF", thF), ("TravThNExit", ns)): t.setParameter(k, v)
    r = run(); res.append((r, thT, thL, thF, ns))
res.sort()
for r in res[:12] + res[-3:]: print("v5", r)
