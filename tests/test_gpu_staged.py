"""The staged traversal kernel (TraversalKernel=2, csrc/device/traverse_staged.cuh: shared-memory stack, 64-byte leaf / instance records, TMA-filled
treelet) against the oracle and against the persistent kernel: same hits, same visit counts, same images, bit for bit, for every launch shape."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
from test_gpu_parity import random_rays, make

pytestmark = pytest.mark.gpu

SHAPES = [(128, 16, 0), (512, 16, 512), (1024, 16, 1024), (256, 2, 64), (512, 0, 7), (64, 64, 2048)]   # (StagedThreads, StagedStackRows, StagedTreeletNodes)


def _staged(t, s, threads, rows, treelet):
    t.setParameter("TraversalKernel", 2); t.setParameter("StagedThreads", threads); t.setParameter("StagedStackRows", rows); t.setParameter("StagedTreeletNodes", treelet)
    t.InitializeScene(s)   # the treelet budget applies at upload
    assert t.getParameter("StagedUsable") == 1
    assert t.getParameter("StagedTreeletNodes") <= treelet


@pytest.mark.parametrize("kind,n,rebraid", [("cornell", 4096, 0), ("cornell7", 4096, 0), ("soup", 4096, 0), ("c2", 8192, 0), ("c4", 4096, 0), ("c4", 4096, 1024)])
def test_staged_trace_rays_bit_exact(built_lib, orc, kind, n, rebraid):
    s, t = make(kind)
    if rebraid:
        s.setRebraid(rebraid)
    for shape in SHAPES:
        _staged(t, s, *shape)
        for inside in (True, False):
            rays = random_rays(s, n, seed=5, inside=inside)
            g, gc = t.trace_rays(rays, counts=True)
            o, oc = orc.trace_rays(s.view, rays, counts=True)
            assert g.tobytes() == o.tobytes(), shape
            assert gc == oc, shape
    t.close()


@pytest.mark.parametrize("kind", ["cornell7", "soup", "c2"])
def test_staged_intersect_closest_and_any_hit(built_lib, orc, kind):
    s, t = make(kind)
    diag = float(np.linalg.norm(np.array(list(s.view.box_max)) - np.array(list(s.view.box_min))))
    rays = random_rays(s, 4099, seed=9, tmin=1e-3 * diag, tmax=0.4 * diag)
    o = orc.intersect(s.view, rays); oa = orc.intersect(s.view, rays, any_hit=True)
    for shape in SHAPES[:4]:
        _staged(t, s, *shape)
        assert np.array_equal(t.intersect(rays), o), shape
        ga = t.intersect(rays, any_hit=True)
        assert np.array_equal(ga["tri_idx"] >= 0, oa["tri_idx"] >= 0), shape
    for n in (1, 31, 33):
        r = random_rays(s, n, seed=n)
        assert np.array_equal(t.intersect(r), orc.intersect(s.view, r))
    t.close()


@pytest.mark.parametrize("kind,w,h,depth", [("cornell7", 96, 96, 8), ("soup", 128, 128, 8), ("c3", 160, 90, 8), ("c4", 96, 54, 8)])
def test_staged_render_equals_persistent_kernel(built_lib, kind, w, h, depth):
    """Same hits => the single-pass image, the ray count and the queue sizes are bit-identical between the two kernels, fused launches included."""
    s, t = make(kind, w, h, depth)
    t.setParameter("TraversalKernel", 0)
    t.DoPass(True); t.synchronize()
    ref = t.readAccumulator(); ref_rays = t.getRaysInLastPass(); ref_q = t.queueSizes(depth)
    for shape in SHAPES[:3]:
        _staged(t, s, *shape)
        for fuse in (1, 0):
            t.setParameter("FuseTraversal", fuse)
            t.DoPass(True); t.synchronize()
            img = t.readAccumulator()
            assert np.array_equal(img["weight_sum"], ref["weight_sum"]), (shape, fuse)
            assert np.allclose(img["rgb"], ref["rgb"], rtol=1e-6, atol=0), (shape, fuse)   # identical paths; only the order of the float atomics into a pixel may differ
            assert t.getRaysInLastPass() == ref_rays
            q = t.queueSizes(depth)
            assert np.array_equal(q[0], ref_q[0]) and np.array_equal(q[1], ref_q[1])
    t.close()


def test_staged_wavefront_path_tracer_and_instrumented_counts(built_lib, orc):
    w = h = 64
    s = ctl.Scene("cornell7", w, h)
    a = ctl.WavefrontPathTracer(w, h); a.InitializeScene(s); a.setParameter("MaxPathLength", 8); a.setParameter("TraversalKernel", 0)
    a.DoPass(True); a.synchronize(); ref = a.readAccumulator(); ref_q = a.queueSizes(8)
    _staged(a, s, 512, 16, 512)
    a.DoPass(True); a.synchronize()
    img = a.readAccumulator()
    assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and np.allclose(img["rgb"], ref["rgb"], rtol=1e-6, atol=0)
    q = a.queueSizes(8)
    assert np.array_equal(q[0], ref_q[0]) and np.array_equal(q[1], ref_q[1])
    a.close()
    # visit counts of an instrumented PathTracer pass: identical between the kernels (they feed the roofline of bench.py)
    s2, t = make("soup", 96, 96, 6)
    t.setParameter("TraversalKernel", 0); t.setInstrumented(1); t.DoPass(True); t.synchronize(); c0 = t.visitCounts()
    _staged(t, s2, 512, 16, 512)
    t.DoPass(True); t.synchronize(); c2 = t.visitCounts()
    assert c0 == c2
    t.close()


def test_staged_after_node_transform_update(built_lib, orc):
    """ctl_update_scene_nodes rebuilds the instance records and the treelet."""
    s, t = make("cornell7")
    _staged(t, s, 512, 16, 512)
    xf = np.eye(4, dtype=np.float32); xf[0, 3] = 0.31; xf[1, 3] = 0.05; xf[2, 3] = 0.4
    s.setNodeTransform(4, xf)
    t.UpdateSceneNodes(s)
    rays = random_rays(s, 4096, seed=3)
    g = t.trace_rays(rays); o = orc.trace_rays(s.view, rays)
    assert g.tobytes() == o.tobytes()
    t.close()


@pytest.mark.parametrize("kind,w,h,depth", [("cornell", 96, 96, 8), ("soup", 128, 128, 8), ("c3", 160, 90, 8), ("c5", 64, 36, 12)])
def test_per_class_shade_launches_equal_the_run_time_dispatch(built_lib, kind, w, h, depth):
    """ShadeMode 1 (default): one k_shade launch per material class present, each with a single BSDF body compiled in, over the class bits the staged
    kernel leaves in the hit records (single-class scenes: no sort pass).  Per-path arithmetic is unchanged, so images, ray counts and queue sizes
    equal ShadeMode 0 (the reference's run-time dispatch, Base/VirtualFuncType.h:90-111) exactly."""
    s, t = make(kind, w, h, depth)
    mask = t.getParameter("MaterialClassMask")
    assert mask != 0 and (bin(mask).count("1") > 1) == (kind != "cornell")
    t.setParameter("ShadeMode", 0)
    t.DoPass(True); t.synchronize()
    ref = t.readAccumulator(); ref_rays = t.getRaysInLastPass(); ref_q = t.queueSizes(depth)
    t.setParameter("ShadeMode", 1)
    for fuse in (1, 0):
        t.setParameter("FuseTraversal", fuse)
        t.DoPass(True); t.synchronize()
        img = t.readAccumulator()
        assert np.array_equal(img["weight_sum"], ref["weight_sum"])
        assert np.allclose(img["rgb"], ref["rgb"], rtol=1e-6, atol=0), fuse   # identical paths; only the order of the float atomics into a pixel may differ
        assert t.getRaysInLastPass() == ref_rays
        q = t.queueSizes(depth)
        assert np.array_equal(q[0], ref_q[0]) and np.array_equal(q[1], ref_q[1])
    # several passes fused into one wavefront (what bench.py runs)
    t.setParameter("FuseTraversal", 1)
    t.DoPasses(4, new_trace=True); t.synchronize(); a = t.readAccumulator(); ra = t.getRaysInLastPass()
    t.setParameter("ShadeMode", 0)
    t.DoPasses(4, new_trace=True); t.synchronize(); b = t.readAccumulator()
    assert ra == t.getRaysInLastPass() and np.array_equal(a["weight_sum"], b["weight_sum"]) and np.allclose(a["rgb"], b["rgb"], rtol=1e-5, atol=1e-7)
    t.close()


@pytest.mark.parametrize("kind,n", [("cornell", 3000), ("cornell7", 5000), ("soup", 5000), ("c2", 20000), ("c4", 6000)])
def test_ray_queue_staging_through_tma_is_exact(built_lib, orc, kind, n):
    """StagedRayTMA=1: refills reach the warp as cp.async.bulk copies into its shared-memory buffer (per-warp mbarrier).  Same rays, same results:
    API queries against the oracle (ragged sizes: the last refill is a partial block), rendered frames against the direct-load kernel (fused launches:
    a refill may cross the extension / shadow queue boundary)."""
    s, t = make(kind, 96, 64, 8)
    t.setParameter("StagedRayTMA", 1)
    assert t.getParameter("StagedRayTMA") == 1
    for threads in (512, 64, 1024):
        t.setParameter("StagedThreads", threads)
        for m in (n, 1, 33, 1000 + threads // 7):
            rays = random_rays(s, m, seed=m)
            g, gc = t.trace_rays(rays, counts=True); o, oc = orc.trace_rays(s.view, rays, counts=True)
            assert g.tobytes() == o.tobytes() and gc == oc, (threads, m)
        diag = float(np.linalg.norm(np.array(list(s.view.box_max)) - np.array(list(s.view.box_min))))
        seg = random_rays(s, 4099, seed=9, tmin=1e-3 * diag, tmax=0.4 * diag)
        assert np.array_equal(t.intersect(seg), orc.intersect(s.view, seg))
        assert np.array_equal(t.intersect(seg, any_hit=True)["tri_idx"] >= 0, orc.intersect(s.view, seg, any_hit=True)["tri_idx"] >= 0)
    t.setParameter("StagedThreads", 512)
    t.setParameter("StagedRayTMA", 0)
    t.DoPasses(3, new_trace=True); t.synchronize(); ref = t.readAccumulator(); ref_rays = t.getRaysInLastPass(); ref_q = t.queueSizes(8)
    t.setParameter("StagedRayTMA", 1)
    for fuse in (1, 0):
        t.setParameter("FuseTraversal", fuse)
        t.DoPasses(3, new_trace=True); t.synchronize(); img = t.readAccumulator()
        assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and np.allclose(img["rgb"], ref["rgb"], rtol=1e-5, atol=1e-7), fuse
        assert t.getRaysInLastPass() == ref_rays
        q = t.queueSizes(8)
        assert np.array_equal(q[0], ref_q[0]) and np.array_equal(q[1], ref_q[1])
    t.close()
    w = ctl.WavefrontPathTracer(64, 64); w.InitializeScene(ctl.Scene("cornell7", 64, 64)); w.setParameter("MaxPathLength", 6)
    w.DoPass(True); w.synchronize(); a = w.readAccumulator()
    w.setParameter("StagedRayTMA", 1); w.DoPass(True); w.synchronize(); b = w.readAccumulator()
    assert np.array_equal(a["weight_sum"], b["weight_sum"]) and np.allclose(a["rgb"], b["rgb"], rtol=1e-6, atol=0)
    w.close()
