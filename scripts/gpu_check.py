"""First-contact GPU check: traversal parity + Cornell image parity vs the oracle, plus rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from cudatracerlib_b200 import Scene, PathTracer, RAY_DTYPE, traversal_bytes
import oracle_binding as ob


def random_rays(scene, n, seed=1):
    rng = np.random.default_rng(seed)
    lo = np.array(list(scene.view.box_min)); hi = np.array(list(scene.view.box_max))
    r = np.zeros(n, RAY_DTYPE)
    r["o"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32); r["tmin"] = 0; r["tmax"] = 3.0e38
    return r


for kind, n in [("cornell", 4096), ("cornell7", 4096), ("soup", 4096), ("c2", 4096)]:
    s = Scene(kind, 256, 256)
    t = PathTracer(256, 256); t.InitializeScene(s)
    rays = random_rays(s, n)
    g, gc = t.trace_rays(rays, counts=True)
    o, oc = ob.trace_rays(s.view, rays, counts=True)
    same = (g["tri_idx"] == o["tri_idx"]) & (g["node_idx"] == o["node_idx"])
    bit = same & (g["dist"].view(np.uint32) == o["dist"].view(np.uint32)) & (g["u"].view(np.uint32) == o["u"].view(np.uint32)) & (g["v"].view(np.uint32) == o["v"].view(np.uint32))
    print(kind, "tris", s.n_triangles, "hit-rate", (o["tri_idx"] != 0xffffffff).mean(), "index-equal", same.mean(), "bit-exact", bit.mean(), "counts gpu", gc, "oracle", oc)
    g16 = t.intersect(rays); o16 = ob.intersect(s.view, rays)
    print("   intersect16 equal:", (g16 == o16).mean(), " anyhit occl equal:", ((t.intersect(rays, True)["tri_idx"] >= 0) == (ob.intersect(s.view, rays, True)["tri_idx"] >= 0)).mean())
    t.close()

for kind in ["cornell", "cornell7", "soup"]:
    s = Scene(kind, 256, 256)
    t = PathTracer(256, 256); t.InitializeScene(s)
    t.setParameter("MaxPathLength", 8)
    t.DoPass(True); t.synchronize()
    img = t.readAccumulator()
    ref, rays = ob.render(s.view, 256, 256, n_passes=1, max_path_length=8)
    a = img["rgb"]; b = ref["rgb"]
    diff = np.linalg.norm(a - b, axis=2) / (np.linalg.norm(b, axis=2) + 1e-3)
    print(kind, "rays gpu", t.getRaysInLastPass(), "oracle", rays, "pixels within 1e-3:", (diff <= 1e-3).mean(), "exact:", (a == b).all(axis=2).mean(),
          "rmse rel", np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()), "mean", a.mean(), b.mean(), "weights eq", (img["weight_sum"] == ref["weight_sum"]).mean())
    t.close()

# timing on C2 at 1080p
W, H = 1920, 1080
t0 = time.time(); s = Scene("c2", W, H); print("c2 build s", time.time() - t0, s.n_triangles, "nodes", s.view.n_bvh_nodes, "woop", s.view.n_woop)
t = PathTracer(W, H); t.InitializeScene(s); t.setParameter("MaxPathLength", 8); t.setParameter("StageTimers", 1)
for i in range(4):
    t.DoPass(i == 0); t.synchronize()
    print("pass", i, "rays", t.getRaysInLastPass(), "sec", t.getLastTimeSpentRenderingSec(), "Mrays/s", t.getRaysInLastPass() / t.getLastTimeSpentRenderingSec() / 1e6, t.stageTimes())
print("queues", t.queueSizes(8))
t.setInstrumented(1); t.DoPass(False); t.synchronize(); e, sh = t.visitCounts(); t.setInstrumented(0)
print("visit ext", e, "shadow", sh, "bytes/ray ext", traversal_bytes(e, e[3]) / e[3], "shadow", traversal_bytes(sh, sh[3]) / max(sh[3], 1))
