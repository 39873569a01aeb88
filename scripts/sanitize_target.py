"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): all kernels incl. sort + resolve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudatracerlib_b200 import Scene, PathTracer, RAY_DTYPE
s = Scene("soup", 96, 64)
t = PathTracer(96, 64); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
t.DoPasses(3, new_trace=True)
t.setParameter("SortMode", 1); t.DoPasses(2); t.DoPassTiled(16, 16, 1, 3)
t.setParameter("TraversalKernel", 1); t.DoPass()
t.synchronize()
rays = np.zeros(777, RAY_DTYPE); rays["o"] = 0.1; rays["d"] = (0.3, 0.5, 0.8); rays["tmax"] = 1e30
t.intersect(rays); t.intersect(rays, True); t.trace_rays(rays, counts=True)
img = t.resolveSRGB8()
print("ok", t.getNumPassesDone(), img.mean())
t.close()
