"""GPU side of the .xmsh import: a scene read from the file the reference's own writer produced (its SplitBVHBuilder tree,
its TriangleData, its Material blobs) renders through the CUDA path like the oracle and like the reference's own PathTrace."""
import os

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
from scene_fixtures import TWO_LIGHT_CAMERA

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
XMSH = os.path.join(HERE, "golden", "two_light_room.xmsh")
GOLD = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


def rel_l2(a, b):
    return np.linalg.norm(a - b, axis=-1) / (np.linalg.norm(b, axis=-1) + 1e-3)


def test_imported_scene_on_gpu(built_lib, orc):
    s = ctl.Scene.from_xmsh(XMSH, *TWO_LIGHT_CAMERA, 64, 64)
    t = ctl.PathTracer(64, 64); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    rng = np.random.default_rng(11)
    rays = np.zeros(8192, api.RAY_DTYPE); rays["o"] = rng.uniform(-0.99, 0.99, (8192, 3)); d = rng.normal(size=(8192, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
    g, gc = t.trace_rays(rays, counts=True); o, oc = orc.trace_rays(s.view, rays, counts=True)
    assert np.array_equal(g["tri_idx"], o["tri_idx"]) and np.array_equal(g["dist"].view(np.uint32), o["dist"].view(np.uint32)) and gc == oc   # the reference's SBVH, bit-exact
    t.DoPass(True); t.DoPass(False)
    img = t.readAccumulator()
    ref = np.ascontiguousarray(GOLD["xmsh_two_light_image_64x64_2spp"]).view(api.PIXEL_DTYPE).reshape(64, 64)   # the reference's PathTrace on its own file
    assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).mean() >= 0.99 and np.array_equal(img["weight_sum"], ref["weight_sum"])
    orc_img, _ = orc.render(s.view, 64, 64, n_passes=2, max_path_length=6)
    assert (rel_l2(img["rgb"], orc_img["rgb"]) <= 1e-3).mean() >= 0.99
    w = ctl.WavefrontPathTracer(64, 64); w.InitializeScene(s); w.setParameter("MaxPathLength", 6)
    w.DoPass(True)
    wref, wrays, _ = orc.render_wavefront(s.view, 64, 64, n_passes=1, max_path_length=6)
    assert (rel_l2(w.readAccumulator()["rgb"], wref["rgb"]) <= 1e-3).mean() >= 0.99 and w.getRaysInLastPass() == wrays
    t.close(); w.close()


def test_obj_scene_on_gpu(built_lib, orc):
    """The OBJ front end feeding the CUDA path: same image as the oracle, and as the reference's own PathTrace on the reference-compiled file."""
    s = ctl.Scene.from_files(os.path.join(HERE, "golden", "obj", "room.obj"), *TWO_LIGHT_CAMERA, 64, 64)
    t = ctl.PathTracer(64, 64); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    t.DoPass(True); t.DoPass(False)
    img = t.readAccumulator()
    ref = np.ascontiguousarray(GOLD["obj_room_image_64x64_2spp"]).view(api.PIXEL_DTYPE).reshape(64, 64)
    assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).mean() >= 0.99 and np.array_equal(img["weight_sum"], ref["weight_sum"])
    o, rays = orc.render(s.view, 64, 64, n_passes=2, max_path_length=6)
    assert (rel_l2(img["rgb"], o["rgb"]) <= 1e-3).mean() >= 0.99 and abs(t.getTotalRays() - rays) <= 2e-3 * rays
    t.close()


def test_node_transform_update_on_gpu(built_lib, orc):
    """Moving an instance and re-uploading only the node level (ctl_update_scene_nodes) renders like a full upload of the moved scene."""
    files = [XMSH, os.path.join(HERE, "golden", "obj", "room.obj")]
    cam = ((1.2, 0, -4.0), (1.2, 0, 0), (0, 1, 0), 60.0)
    x0 = np.stack([np.eye(4, dtype=np.float32)] * 2); x0[1, 0, 3] = 2.5
    x1 = x0.copy(); x1[1] = np.array([[0.8, 0, 0.6, 2.7], [0, 1, 0, 0.2], [-0.6, 0, 0.8, 0.3], [0, 0, 0, 1]], np.float32)
    moved = ctl.Scene.from_files(files, *cam, 64, 40, node_xforms=x1)
    s = ctl.Scene.from_files(files, *cam, 64, 40, node_xforms=x0)
    t = ctl.PathTracer(64, 40); t.InitializeScene(s); t.setParameter("MaxPathLength", 5)
    t.DoPass(True); before = t.readAccumulator().copy()
    s.setNodeTransform(1, x1[1]); t.UpdateSceneNodes(s)
    t.DoPass(True); after = t.readAccumulator().copy()
    full = ctl.PathTracer(64, 40); full.InitializeScene(moved); full.setParameter("MaxPathLength", 5)
    full.DoPass(True); ref = full.readAccumulator()
    assert np.array_equal(after["rgb"].view(np.uint32), ref["rgb"].view(np.uint32)) and not np.array_equal(before["rgb"], after["rgb"])
    o, _ = orc.render(moved.view, 64, 40, n_passes=1, max_path_length=5)
    assert (rel_l2(after["rgb"], o["rgb"]) <= 1e-3).mean() >= 0.99
    t.close(); full.close()
