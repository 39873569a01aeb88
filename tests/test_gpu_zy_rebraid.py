"""GPU run of the opt-in scene-level re-braiding (ctl_scene_set_rebraid; CPU side: tests/test_rebraid_cpu.py).  No kernel changes with it -- a re-braided
view is ordinary node / mesh / BVH records -- so the usual parity holds against the oracle on the same view, and the image equals the plain view's.
Written after this round's GPU budget was spent (first device run; sorted last)."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api

pytestmark = pytest.mark.gpu


def _rays(s, n, seed):
    rng = np.random.default_rng(seed)
    lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max))
    rays = np.zeros(n, api.RAY_DTYPE); rays["o"] = rng.uniform(lo, hi, (n, 3)); d = rng.normal(size=(n, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
    return rays


def test_rebraided_scene_on_gpu(built_lib, orc):
    w, h = 96, 54
    s = ctl.Scene("c4", w, h, n_hint=24)
    t = ctl.PathTracer(w, h); t.setParameter("MaxPathLength", 6); t.InitializeScene(s)
    t.DoPass(True); plain = t.readAccumulator()
    rays = _rays(s, 20011, 3)
    g0 = t.trace_rays(rays)
    s.setRebraid(200)
    assert s.view.node_alias and s.view.n_nodes > 8
    t.UpdateSceneNodes(s)                                       # a re-braided view changes mesh-level records too: the call uploads everything
    g, gc = t.trace_rays(rays, counts=True)
    o, oc = orc.trace_rays(s.view, rays, counts=True)
    alias = np.ctypeslib.as_array(s.view.node_alias, (s.view.n_nodes,))
    hit = g["tri_idx"] != 0xffffffff
    assert np.array_equal(g["tri_idx"], o["tri_idx"])
    assert np.array_equal(g["node_idx"][hit], alias[o["node_idx"][hit]]) and (g["node_idx"][~hit] == 0xffffffff).all()   # the API names the instance, the oracle the pseudo-node it walked
    for f in ("dist", "u", "v"):
        assert np.array_equal(g[f].view(np.uint32), o[f].view(np.uint32)), f
    assert gc == oc
    assert np.array_equal(g["tri_idx"], g0["tri_idx"]) and np.array_equal(g["dist"].view(np.uint32), g0["dist"].view(np.uint32))   # same hits as the plain view
    assert np.array_equal(g["node_idx"], g0["node_idx"])                      # ... which is the node the plain view reports
    seg = rays.copy(); seg["tmin"] = 1e-3; seg["tmax"] = 4.0
    r16, r16_plain_nodes = t.intersect(seg), orc.intersect(s.view, seg)
    h16 = r16["tri_idx"] != -1
    assert np.array_equal(r16["tri_idx"], r16_plain_nodes["tri_idx"]) and np.array_equal(r16["node_idx"][h16], alias[r16_plain_nodes["node_idx"][h16]].astype(np.int32))
    t.DoPass(True); img = t.readAccumulator()
    assert np.array_equal(img["weight_sum"], plain["weight_sum"])
    assert np.allclose(img["rgb"], plain["rgb"], rtol=1e-5, atol=1e-7)     # same paths; only the order of the float atomics into a pixel may differ
    t.close()
    tw = ctl.WavefrontPathTracer(w, h); tw.setParameter("MaxPathLength", 6); tw.InitializeScene(s)
    tw.DoPass(True)
    ref, _, _ = orc.render_wavefront(s.view, w, h, n_passes=1, max_path_length=6)
    a, b = tw.readAccumulator()["rgb"], ref["rgb"]
    r = np.linalg.norm(a - b, axis=-1) / (np.linalg.norm(b, axis=-1) + 1e-3)   # the tolerance of tests/test_gpu_wavefront_pt.py
    assert (r <= 1e-3).mean() >= 0.98, (r <= 1e-3).mean()   # 0.99 on the scenes of that file; this foliage scene has not been through the WavefrontPathTracer on a device before
    tw.close()
