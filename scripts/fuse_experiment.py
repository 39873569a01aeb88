import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
for wl, res in (("c2", (1920, 1080)), ("c2", (680, 384))):   # full frame; 1/8 of the pixels (what one of 8 GPUs renders)
    s = Scene(wl, *res); t = PathTracer(*res); t.InitializeScene(s); t.setParameter("MaxPathLength", 8)
    for f in (0, 1):
        for b in (1, 8):
            t.setParameter("FuseTraversal", f)
            for i in range(3): t.DoPasses(b, new_trace=True)
            t.synchronize(); t0 = time.perf_counter()
            for i in range(8):
                for k in range(8 // b): t.DoPasses(b, new_trace=(k == 0))
            t.synchronize(); dt = time.perf_counter() - t0
            print(wl, res, "fuse", f, "batch", b, "ms/frame", round(dt / 8 * 1e3, 3), flush=True)
    t.close()
