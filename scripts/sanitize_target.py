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
# round-2 additions: staged traversal kernel variants (treelet through TMA, ray-queue TMA), per-class shade launches, frames on wavefront lanes,
# re-braided scene level, the agglomerative GPU builder with and without triangle pre-splitting, NonLocalMeansFilter
import ctypes as C
from cudatracerlib_b200 import lib
s3 = Scene("c3", 128, 72)
t = PathTracer(128, 72); t.InitializeScene(s3); t.setParameter("MaxPathLength", 6)
t.DoFrame(8, 2); t.setParameter("OverlapLanes", 8); t.DoFrame(8, 1); t.setParameter("OverlapWavefronts", 2); t.setParameter("OverlapLanes", 3); t.DoFrame(6, 6)
t.setParameter("DeviceSampleTables", 0); t.DoFrame(4, 1); t.setParameter("DeviceSampleTables", 1)
t.setParameter("StagedRayTMA", 1); t.DoPasses(2, new_trace=True); t.setParameter("StagedRayTMA", 0)
t.setParameter("StagedTreeletNodes", 256); t.InitializeScene(s3); t.DoPasses(2, new_trace=True); t.setParameter("StagedTreeletNodes", 0); t.InitializeScene(s3)
t.setParameter("ShadeMode", 0); t.DoPasses(2, new_trace=True); t.setParameter("TravDrainPrefetch", 1); t.DoPasses(1); t.setParameter("StopZeroThroughput", 0); t.DoPasses(1)
t.setParameter("PixelVarianceBuffer", 1); t.DoPass(True); t.DoPass(False)
t.synchronize(); out = t.applyImagePipeline(ImagePipeline(5, 1.0, 0.0, 0.45, 1.0)); print("round-2 frames ok", out.mean()); t.close()
L = lib()
L.ctl_bvh_build_gpu_split.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
rng = np.random.default_rng(3); n = 3000
a = rng.uniform(-1, 1, size=(n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); w3 = np.cross(d, rng.normal(size=(n, 3)))
verts = np.concatenate([a, a + 0.8 * d, a + 0.4 * d + 0.02 * w3], axis=1).astype(np.float32); verts[:300] = verts[0]
for alg, growth in ((1, 0.0), (1, 3.0), (0, 3.0), (2, 1.0)):
    cap = int(n * (1 + growth)) + 1
    nodes = np.zeros((cap, 16), np.float32); woop = np.zeros((cap, 12), np.float32); index = np.zeros(cap, np.uint32); nn = C.c_uint32(0); ns = C.c_uint32(0)
    assert L.ctl_bvh_build_gpu_split(0, verts.ctypes.data, n, alg, 0, growth, cap, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, C.byref(ns), None) == 0
s4 = Scene("soup", 96, 64); s4.setRebraid(64); s4.rebuildBVHOnGPU(); s4.validate()
t = PathTracer(96, 64); t.InitializeScene(s4); t.DoPasses(2, new_trace=True); t.synchronize(); t.close()
# ray hand-over between half-wavefronts, straggler deferral with lagging paths, WavefrontPathTracer frames on lanes, Regularization
t = PathTracer(128, 72); t.InitializeScene(s3); t.setParameter("MaxPathLength", 6)
t.setParameter("HandOver", 1); t.setParameter("HandOverDrain", 1); t.DoFrame(4, 4); t.setParameter("HandOver", 0)
t.setParameter("DeferStragglers", 1); t.DoFrame(4, 4); t.setParameter("DeferMaxLag", 1); t.DoFrame(8, 2); t.setParameter("DeferStragglers", 0)
t.setParameter("Regularization", 1); t.DoPasses(2, new_trace=True); t.setParameter("Regularization", 0); t.setParameter("ShadeConcurrent", 1); t.DoPasses(2, new_trace=True)
t.synchronize(); t.close()
w = WavefrontPathTracer(96, 64); w.InitializeScene(s); w.setParameter("MaxPathLength", 5); w.DoFrame(6); w.synchronize(); w.close()
print("round-2 all ok")
