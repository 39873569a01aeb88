import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cudatracerlib_b200 as ctl
import oracle_binding as ob
for kind, w, h, depth in (("c5", 64, 36, 32), ("c4", 96, 54, 8), ("c3", 160, 90, 8), ("c5", 128, 72, 32)):
    s = ctl.Scene(kind, w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
    t.DoPass(True); img = t.readAccumulator()
    ref, rr = ob.render(s.view, w, h, n_passes=1, max_path_length=depth)
    rel = np.linalg.norm(img["rgb"] - ref["rgb"], axis=-1) / (np.linalg.norm(ref["rgb"], axis=-1) + 1e-3)
    print(os.environ.get("CTL_B200_LIB", "default"), kind, w, h, depth, "frac<=1e-3", (rel <= 1e-3).mean(), "frac<=1e-5", (rel <= 1e-5).mean(), "rays", t.getRaysInLastPass(), rr, flush=True)
    t.close()
