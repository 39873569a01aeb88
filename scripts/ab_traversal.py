"""A/B timing of the traversal kernels on the GPU: persistent phase-scheduled (0) vs simple ray-batch (1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
from bench import WORKLOADS

for wl in (sys.argv[1:] or ["c2"]):
    kind, w, h, spp, depth, _ = WORKLOADS[wl]
    t0 = time.time(); s = Scene(kind, w, h); tb = time.time() - t0
    t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth); t.setParameter("StageTimers", 1)
    print(f"{wl}: {s.n_triangles} tris, build {tb:.1f}s, nodes {s.view.n_nodes}")
    for kern in (1, 0):
        for bps in ((8,) if kern == 1 else (4, 6, 8, 10, 12)):
            t.setParameter("TraversalKernel", kern); t.setParameter("TraversalBlocksPerSM", bps)
            best = None
            for i in range(4):
                t.DoPass(i == 0); t.synchronize()
                ms, nl = t.stageTimes()
                r = t.getRaysInLastPass(); sec = t.getLastTimeSpentRenderingSec()
                if best is None or sec < best[1]:
                    best = (r, sec, ms)
            r, sec, ms = best
            print(f"  kernel {kern} blocks/SM {bps:2d}: {r/sec/1e6:8.1f} Mrays/s  pass {sec*1e3:7.3f} ms  ext {ms[1]:.3f} shade {ms[2]:.3f} shadow {ms[3]:.3f}")
    t.close()
