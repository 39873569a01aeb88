"""GPU builders (agglomerative PLOC = default, LBVH; with / without triangle pre-splitting) vs the CPU builder: build time and traversal cost of the resulting trees.
Usage: bvh_build_bench.py [radius ...]   (extra PLOC search windows to try besides the default)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cudatracerlib_b200 import Scene, PathTracer
for kind in ("c2", "c4"):
    t0 = time.time(); s = Scene(kind, 1920, 1080); t_cpu = time.time() - t0
    res = {}
    times = {}
    for tag in ["cpu", "gpu-lbvh", "gpu-ploc", "gpu-lbvh-split3", "gpu-ploc-split1", "gpu-ploc-split2", "gpu-ploc-split3", "gpu-ploc-split6-scale2"] + ["gpu-ploc-r" + r for r in sys.argv[1:]]:
        if tag != "cpu":
            os.environ["CTL_GPU_BUILDER"] = "lbvh" if "lbvh" in tag else "ploc"
            os.environ["CTL_PLOC_RADIUS"] = tag.split("-r")[1] if "-r" in tag else "0"
            os.environ["CTL_GPU_SPLIT"] = tag.split("-split")[1].split("-")[0] if "-split" in tag else "0"
            os.environ["CTL_GPU_SPLIT_SCALE"] = tag.split("-scale")[1] if "-scale" in tag else "1"
            t0 = time.time(); ms = s.rebuildBVHOnGPU(); wall = time.time() - t0; times[tag] = (ms, wall)
        t = PathTracer(1920, 1080); t.InitializeScene(s); t.setParameter("MaxPathLength", 8); t.setParameter("StageTimers", 1)
        best = None
        for i in range(3):
            t.DoPasses(2, new_trace=True); t.synchronize(); m, _ = t.stageTimes()
            if best is None or m[1] + m[3] < best: best = m[1] + m[3]
        t.setInstrumented(1); t.DoPass(True); t.synchronize(); e, sh = t.visitCounts(); t.setInstrumented(0)
        res[tag] = (best, e[0] / e[3], e[1] / e[3], (s.view.n_bvh_nodes, s.view.n_woop))
        t.close()
    print(kind, "tris", s.n_triangles, "CPU scene build (all meshes, split BVH + post-pass + encoders) %.2f s" % t_cpu, flush=True)
    for tag, (trav, ni, nt, nn) in res.items():
        bt = " build %.2f ms device, %.2f s wall incl. copies" % times[tag] if tag in times else ""
        print("   ", tag, "traversal ms / 2-pass wavefront %.2f (x%.3f of cpu)" % (trav, trav / res["cpu"][0]), "inner nodes/ray %.1f tris/ray %.1f" % (ni, nt), "nodes", nn[0], "slots", nn[1], bt, flush=True)
