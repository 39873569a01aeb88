"""Opt-in partial re-braiding of the scene level (csrc/scene_builder.cpp rebraid(), ctl_scene_set_rebraid): overlapping instances are opened into
(instance, sub-tree) entries expressed in the reference's own node / mesh records, so kernels and oracle traverse them unchanged.  CPU: the oracle on
the re-braided view finds bit-identical hits and renders bit-identical images, needs fewer node visits, the view stays structurally valid, and
switching it off restores the plain view."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api


def _rays(s, n, seed):
    rng = np.random.default_rng(seed)
    lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max))
    rays = np.zeros(n, api.RAY_DTYPE); rays["o"] = rng.uniform(lo, hi, (n, 3)); d = rng.normal(size=(n, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
    return rays


@pytest.mark.parametrize("kind,hint,budget", [("c4", 24, 64), ("c4", 24, 1000), ("soup", 400, 24)])
def test_rebraided_view_same_hits_same_image_fewer_visits(built_lib, orc, kind, hint, budget):
    s = ctl.Scene(kind, 48, 32, n_hint=hint)
    n_nodes, n_meshes, n_bvh = s.view.n_nodes, s.view.n_meshes, s.view.n_bvh_nodes
    rays = _rays(s, 5000, 4)
    a, ca = orc.trace_rays(s.view, rays, counts=True)
    any_a = orc.intersect(s.view, rays, any_hit=True) if hasattr(orc, "intersect") else None
    img_a, rays_a = orc.render(s.view, 48, 32, 2, max_path_length=5)
    assert not s.view.node_alias
    s.setRebraid(budget)
    s.validate()
    v = s.view
    assert v.n_nodes > n_nodes and v.n_meshes == n_meshes + (v.n_nodes - n_nodes) and v.n_bvh_nodes > n_bvh and v.scene_start_node == 0
    assert v.n_scene_bvh_nodes + 1 <= budget                                    # leaves = inner nodes + 1
    alias = np.ctypeslib.as_array(v.node_alias, (v.n_nodes,))
    assert (alias[:n_nodes] == np.arange(n_nodes)).all() and (alias[n_nodes:] < n_nodes).all()
    nodes = s.array("nodes"); xf = s.array("node_xf")
    for i in range(n_nodes, v.n_nodes):                                         # a pseudo-node is its instance's record over another mesh record
        assert (nodes[i, 1:] == nodes[alias[i], 1:]).all() and nodes[i, 0] >= n_meshes and np.array_equal(xf[i], xf[alias[i]])
    meshes = s.array("meshes")
    for i in range(n_nodes, v.n_nodes):
        pm, rm = meshes[nodes[i, 0]], meshes[nodes[alias[i], 0]]
        assert pm[0] == rm[0] and (pm[2:] == rm[2:]).all() and pm[1] // 4 >= n_bvh   # same triangles / Woop / index / materials, own node range
    b, cb = orc.trace_rays(v, rays, counts=True)
    assert np.array_equal(a["tri_idx"], b["tri_idx"])
    for f in ("dist", "u", "v"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
    hit = a["tri_idx"] != 0xffffffff
    assert np.array_equal(alias[b["node_idx"][hit]], a["node_idx"][hit])        # the reported node maps back to the instance
    img_b, rays_b = orc.render(v, 48, 32, 2, max_path_length=5)
    assert rays_a == rays_b and img_a.tobytes() == img_b.tobytes()
    if kind == "c4":
        # fewer algorithmic bytes (node visits + instance entries); at this size (24 spheres + 3.4 K foliage triangles) the gain is small -- on the full 1 M-triangle config 4 the oracle
        # counts 90.4 -> 65.9 inner nodes per ray (7 089 -> 5 169 algorithmic bytes) with 1 024 entries
        assert api.traversal_bytes(cb, len(rays)) < api.traversal_bytes(ca, len(rays)), (ca, cb)
    s.setRebraid(0)
    assert (s.view.n_nodes, s.view.n_meshes, s.view.n_bvh_nodes) == (n_nodes, n_meshes, n_bvh) and not s.view.node_alias


def test_rebraid_follows_node_transforms_and_env(built_lib, orc, monkeypatch):
    s = ctl.Scene("soup", 32, 32, n_hint=400)
    s.setRebraid(20)
    xf = np.eye(4, dtype=np.float32); xf[0, 3] = 0.3
    s.setNodeTransform(2, xf)                                                    # re-assembles the node level, re-braiding included
    s.validate()
    assert s.view.node_alias and s.view.n_nodes > 4
    plain = ctl.Scene("soup", 32, 32, n_hint=400); plain.setNodeTransform(2, xf)
    rays = _rays(plain, 3000, 9)
    a, b = orc.trace_rays(plain.view, rays), orc.trace_rays(s.view, rays)
    assert np.array_equal(a["tri_idx"], b["tri_idx"]) and np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
    monkeypatch.setenv("CTL_REBRAID", "20")                                      # A/B switch for bench.py runs
    e = ctl.Scene("soup", 32, 32, n_hint=400)
    assert e.view.node_alias and e.view.n_nodes > 4
    monkeypatch.delenv("CTL_REBRAID")
    assert not ctl.Scene("soup", 32, 32, n_hint=400).view.node_alias
    c7 = ctl.Scene("cornell7", 32, 32); c7.setRebraid(64)                        # seven meshes of <= 8 triangles: single leaves, nothing to open
    assert not c7.view.node_alias
    one = ctl.Scene("cornell", 32, 32); one.setRebraid(64)                       # a single instance is left alone
    assert not one.view.node_alias and one.view.n_nodes == 1
