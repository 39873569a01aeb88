"""ncu target for the WavefrontPathTracer drop-in: two passes of one workload (default c2 @1080p, depth 8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, WavefrontPathTracer
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
kind, w, h, spp, depth, _ = WORKLOADS[wl]
s = Scene(kind, w, h)
t = WavefrontPathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
t.DoPass(True); t.DoPass(False); t.synchronize()
print(wl, "rays last pass", t.getRaysInLastPass(), "sec", t.getLastTimeSpentRenderingSec())
