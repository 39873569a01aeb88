"""Small fixed target for ncu: N passes of one workload (default c2 @1080p, depth 8). Usage: profile_target.py [workload] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind, w, h, spp, depth, _ = WORKLOADS[wl]
s = Scene(kind, w, h)
t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
for i in range(n):
    t.DoPass(i == 0)
t.synchronize()
print(wl, "rays last pass", t.getRaysInLastPass(), "sec", t.getLastTimeSpentRenderingSec())
