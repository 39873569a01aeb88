"""Reference SplitBVHBuilder (through oracle/_ref's .xmsh writer) vs this repo's binned-SAH builder on the same scene: nodes, references,
traversal visit counts of the oracle on random + camera rays.  CPU only (needs oracle/_ref, i.e. /root/reference at build time)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
import ref_binding as rb, oracle_binding as ob

kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
n_hint = int(sys.argv[2]) if len(sys.argv) > 2 else 0
s = ctl.Scene(kind, 256, 144, n_hint=n_hint) if n_hint else ctl.Scene(kind, 256, 144)
print(kind, "tris", s.n_triangles, "meshes", s.view.n_meshes, "nodes", s.view.n_bvh_nodes, "refs", s.view.n_woop)
tmp = tempfile.mkdtemp()
paths = []
tri_data = s.array("tri_data"); meshes = s.array("meshes")
import ctypes as C
mats_all = (api.Material * s.view.n_materials).from_address(C.addressof(s.view.materials.contents))
t0 = time.time()
for mi in range(s.view.n_meshes):
    T = s.mesh_triangles(mi)
    toff, moff = int(meshes[mi][0]), int(meshes[mi][4])
    mat_idx = ((tri_data[toff:toff + len(T), 1] >> 16) & 0xff).astype(np.int64)
    order = np.argsort(mat_idx, kind="stable")
    used = int(mat_idx.max()) + 1
    V = T[order].reshape(-1, 3); I = np.arange(len(V), dtype=np.uint32)
    p = os.path.join(tmp, f"m{mi}.xmsh")
    rb.write_xmsh(p, V, I, np.bincount(mat_idx, minlength=used), [mats_all[moff + k] for k in range(used)], None)
    paths.append(p)
print("reference SBVH build + write: %.1f s" % (time.time() - t0))
xf = s.array("node_xf").reshape(-1, 16)
node_mesh = s.array("nodes")[:, 0]
paths = [paths[int(m)] for m in node_mesh]   # one file per node (instanced meshes are simply read again)
cam = ((0, 0, -9.5), (0, 0, 0), (0, 1, 0), 60.0)
s2 = ctl.Scene.from_xmsh(paths, *cam, 256, 144, node_xforms=xf)
print("imported: nodes", s2.view.n_bvh_nodes, "refs", s2.view.n_woop, "(duplication %.3f)" % (s2.view.n_woop / s.n_triangles))
rng = np.random.default_rng(1)
lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max))
rays = np.zeros(20000, api.RAY_DTYPE); rays["o"] = rng.uniform(lo, hi, (20000, 3)); d = rng.normal(size=(20000, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
a, ca = ob.trace_rays(s.view, rays, counts=True); b, cb = ob.trace_rays(s2.view, rays, counts=True)
print("hits equal:", (a["dist"] == b["dist"]).mean())
print("own SAH   : inner %.1f tris %.1f inst %.2f per ray -> %.0f B/ray" % (ca[0] / 2e4, ca[1] / 2e4, ca[2] / 2e4, api.traversal_bytes(ca, 20000) / 2e4))
print("ref SBVH  : inner %.1f tris %.1f inst %.2f per ray -> %.0f B/ray" % (cb[0] / 2e4, cb[1] / 2e4, cb[2] / 2e4, api.traversal_bytes(cb, 20000) / 2e4))
