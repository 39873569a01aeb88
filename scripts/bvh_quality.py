"""BVH quality probe (CPU): build time, nodes, references, and the oracle's visit counts on 20 000 random rays.  CTL_BVH_BUILDER=sah selects the plain binned-SAH builder."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np
import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
import oracle_binding as ob
kind = sys.argv[1]
t0=time.time(); s = ctl.Scene(kind, 256, 144); t1=time.time()
rng = np.random.default_rng(1)
lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max))
N=20000
rays = np.zeros(N, api.RAY_DTYPE); rays["o"] = rng.uniform(lo, hi, (N, 3)); d = rng.normal(size=(N, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
a, ca = ob.trace_rays(s.view, rays, counts=True)
print(os.environ.get("CTL_BVH_BUILDER","sbvh"), kind, "build %.1fs" % (t1-t0), "tris", s.n_triangles, "nodes", s.view.n_bvh_nodes, "refs", s.view.n_woop, "inner %.1f tris %.1f per ray -> %.0f B/ray" % (ca[0]/N, ca[1]/N, api.traversal_bytes(ca, N)/N), "hit checksum", int(a["tri_idx"].astype(np.uint64).sum()), float(a["dist"][a["tri_idx"]!=0xffffffff].astype(np.float64).sum()))
