"""Small fixed target for ncu: N fused-pass wavefronts of one workload (default c2 @1080p, depth 8, 8 passes per wavefront).
Usage: profile_target.py [workload] [wavefronts] [passes_per_wavefront] [n_parts] [KEY=INT ...]   (n_parts > 1: the tiles of part 0 only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n_parts = int(sys.argv[4]) if len(sys.argv) > 4 else 1
kind, w, h, spp, depth, _ = WORKLOADS[wl]
s = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
for kv in sys.argv[5:]:
    k, v = kv.split("="); t.setParameter(k, int(v))
for i in range(n):
    t.DoPasses(batch, new_trace=(i == 0), part=0, n_parts=n_parts)
t.synchronize()
print(wl, "rays last wavefront", t.getRaysInLastPass(), "sec", t.getLastTimeSpentRenderingSec())
