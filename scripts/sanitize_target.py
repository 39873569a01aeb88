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
# round-1 additions: WavefrontPathTracer queue kernels (chained scan), fused API traversal, full image pipeline, variance buffer, SortMode 2, GPU BVH
from cudatracerlib_b200 import WavefrontPathTracer, ImagePipeline
w = WavefrontPathTracer(96, 64); w.InitializeScene(s); w.setParameter("MaxPathLength", 7); w.setParameter("RRStartDepth", 2); w.setParameter("PixelVarianceBuffer", 1)
w.DoPass(True); w.DoPass(False); w.setParameter("Direct", 0); w.DoPass(False); w.setParameter("FuseTraversal", 0); w.setParameter("Direct", 1); w.DoPass(True)
v = w.readVarianceBuffer()
for P in (ImagePipeline(3, 2, 2, 1 / 3, 1 / 3), ImagePipeline(4, 3, 2, 2.0, tonemap=1), ImagePipeline(-1, tonemap=1)):
    out = w.applyImagePipeline(P)
print("wavefront ok", w.getNumPassesDone(), int(v["iterations_done"].max()), out.mean())
w.close()
t = PathTracer(96, 64); t.InitializeScene(s); t.setParameter("MaxPathLength", 5); t.setParameter("SortMode", 2); t.DoPasses(2, new_trace=True); t.synchronize(); t.close()
s2 = Scene("soup", 96, 64); s2.rebuildBVHOnGPU()
print("all ok")
