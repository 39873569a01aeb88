"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI (ctypes on
libctl_b200.so); the CPU oracle is the checker.

Tolerances: traversal -- (node, tri) indices and t/u/v bit-exact (integer/index work and explicit-FMA
float work); radiance -- same seed, same pass: per-pixel relative L2 ||a-b|| / (||b|| + 1e-3) <= 1e-3 on >= 99 % of
the pixels, whole-image relative RMSE <= 1e-2 at 1 spp (<= 3e-3 at 8 spp), mean within 0.1 %."""
import ctypes as C

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api

pytestmark = pytest.mark.gpu


def random_rays(scene, n, seed=1, tmin=0.0, tmax=3.0e38, inside=True):
    rng = np.random.default_rng(seed)
    lo = np.array(list(scene.view.box_min)); hi = np.array(list(scene.view.box_max))
    if not inside:
        c, e = (lo + hi) / 2, (hi - lo)
        lo, hi = c - 1.5 * e, c + 1.5 * e
    r = np.zeros(n, api.RAY_DTYPE)
    r["o"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32); r["tmin"] = tmin; r["tmax"] = tmax
    return r


def make(kind, w=64, h=64, depth=8, **kw):
    s = ctl.Scene(kind, w, h, **kw)
    t = ctl.PathTracer(w, h)
    t.InitializeScene(s)
    t.setParameter("MaxPathLength", depth)
    return s, t


def rel_l2(a, b):
    return np.linalg.norm(a - b, axis=-1) / (np.linalg.norm(b, axis=-1) + 1e-3)


# ------------------------------------------------------------------ traversal
@pytest.mark.parametrize("kind,n", [("cornell", 4096), ("cornell7", 4096), ("soup", 4096), ("c2", 8192), ("c3", 2048), ("c4", 2048)])
def test_trace_rays_bit_exact(built_lib, orc, kind, n):
    s, t = make(kind)
    for inside in (True, False):
        rays = random_rays(s, n, seed=5, inside=inside)
        g, gc = t.trace_rays(rays, counts=True)
        o, oc = orc.trace_rays(s.view, rays, counts=True)
        for f in ("tri_idx", "node_idx"):
            assert np.array_equal(g[f], o[f]), f
        for f in ("dist", "u", "v"):
            assert np.array_equal(g[f].view(np.uint32), o[f].view(np.uint32)), f
        assert gc == oc  # inner nodes popped, triangle refs tested, instance leaves entered: the roofline's visit counts
    t.close()


@pytest.mark.parametrize("kind", ["cornell7", "soup", "c2"])
def test_intersect_buffers_closest_and_any_hit(built_lib, orc, kind):
    s, t = make(kind)
    diag = float(np.linalg.norm(np.array(list(s.view.box_max)) - np.array(list(s.view.box_min))))
    rays = random_rays(s, 4099, seed=9, tmin=1e-3 * diag, tmax=0.4 * diag)  # ragged size, finite segments
    g = t.intersect(rays); o = orc.intersect(s.view, rays)
    assert np.array_equal(g, o)  # 16-byte traversalResult incl. u16 barycentrics, bit for bit
    miss = o["tri_idx"] == -1
    assert miss.any() and (~miss).any()
    assert np.all(o["node_idx"][miss] == -1) and np.all(o["bary"][miss] == 0)
    ga = t.intersect(rays, any_hit=True); oa = orc.intersect(s.view, rays, any_hit=True)
    assert np.array_equal(ga["tri_idx"] >= 0, oa["tri_idx"] >= 0)      # occlusion boolean is order-independent
    assert np.array_equal(ga["tri_idx"] >= 0, o["tri_idx"] >= 0)        # and equals "closest hit exists"
    t.close()


def test_intersect_edge_sizes(built_lib, orc):
    s, t = make("cornell")
    assert len(t.intersect(np.zeros(0, api.RAY_DTYPE))) == 0           # empty input
    assert len(t.trace_rays(np.zeros(0, api.RAY_DTYPE))) == 0
    for n in (1, 31, 32, 33, 1000):
        rays = random_rays(s, n, seed=n)
        assert np.array_equal(t.intersect(rays), orc.intersect(s.view, rays))
    # degenerate rays: zero direction, axis-parallel directions (idir guard 2^-80, BVHTraversal.h:16-19), NaN origin
    rays = random_rays(s, 8, seed=3)
    rays["d"][0] = 0; rays["d"][1] = (1, 0, 0); rays["d"][2] = (0, -1, 0); rays["d"][3] = (0, 0, 1); rays["o"][4] = np.nan
    g = t.trace_rays(rays); o = orc.trace_rays(s.view, rays)
    assert np.array_equal(g["tri_idx"], o["tri_idx"])
    t.close()


def test_intersect_device_pointers_async(built_lib, orc):
    """ctl_intersect on device pointers (== __internal__IntersectBuffers) driven from torch tensors on a torch stream."""
    import torch
    s, t = make("soup")
    rays = random_rays(s, 10000, seed=2)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_res = torch.zeros(len(rays), 4, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        t.intersect_device(len(rays), d_rays.data_ptr(), d_res.data_ptr(), False, st.cuda_stream)
    st.synchronize()
    g = d_res.cpu().numpy().view(api.RESULT16_DTYPE).reshape(-1)
    assert np.array_equal(g, orc.intersect(s.view, rays))
    t.close()


# ------------------------------------------------------------------ radiance
@pytest.mark.parametrize("kind,w,h,depth", [("cornell", 256, 256, 8), ("cornell7", 128, 128, 8), ("soup", 128, 128, 8), ("c3", 160, 90, 8), ("cornell", 64, 64, 32),
                                              ("c4", 96, 54, 8), ("c5", 64, 36, 32)])
def test_render_pass_matches_oracle(built_lib, orc, kind, w, h, depth):
    s, t = make(kind, w, h, depth)
    t.DoPass(True); t.synchronize()
    img = t.readAccumulator()
    ref, ref_rays = orc.render(s.view, w, h, n_passes=1, max_path_length=depth)
    a, b = img["rgb"], ref["rgb"]
    frac = (rel_l2(a, b) <= 1e-3).mean()
    rmse = np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean())
    # c5 = 1 M triangles, glass + rough-conductor spheres, 32 bounces: specular chains amplify 1-ulp differences chaotically.  Measured on the
    # CPU alone: the oracle's FMA and no-FMA builds (which differ by <= 1 ulp in t) agree on only 90.8 % of these pixels at 1e-3 (96.3 % at
    # depth 8), the CUDA path agrees with the oracle on 96.3 % (libdevice vs libm) -> floor 0.93 for that case, 0.99 everywhere else.
    assert frac >= (0.93 if kind == "c5" else 0.99), frac
    # whole-image RMSE: 1e-2 at 1 spp; the tiny, dark 1M-triangle test images have a handful of paths whose discrete
    # decisions flip (libdevice vs libm) and each of them is a visible share of so few pixels -> 3e-2 there
    if kind == "c5":   # the few percent of chaotically diverged deep specular paths carry fireflies: RMSE / mean are not meaningful, the median is
        assert np.median(rel_l2(a, b)) <= 1e-5
    else:
        assert rmse <= (3e-2 if kind == "c4" else 1e-2), rmse
        assert abs(a.mean() - b.mean()) <= (2e-3 if kind == "c4" else 1e-3) * b.mean()
    assert np.array_equal(img["weight_sum"], ref["weight_sum"])
    assert np.all(img["rgb_splat"] == 0)
    assert abs(t.getRaysInLastPass() - ref_rays) <= (5e-3 if kind == "c5" else 2e-3) * ref_rays
    assert t.getNumPassesDone() == 1
    t.close()


def test_eight_passes_match_oracle(built_lib, orc):
    w = h = 96
    s, t = make("cornell7", w, h, 8)
    for p in range(8):
        t.DoPass(p == 0)
    t.synchronize()
    img = t.readAccumulator()
    ref, _ = orc.render(s.view, w, h, n_passes=8, max_path_length=8)
    a, b = img["rgb"], ref["rgb"]
    assert (rel_l2(a, b) <= 1e-3).mean() >= 0.99
    assert np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()) <= 3e-3
    assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and img["weight_sum"].sum() == pytest.approx(8 * w * h, abs=8)
    assert t.getNumPassesDone() == 8
    t.close()


def test_parameters_direct_rr(built_lib, orc):
    """Direct=false (pure BSDF sampling, PathTracer.cu:64-82) and RRStartDepth are honoured like the reference."""
    w = h = 64
    s, t = make("cornell", w, h, 6)
    for direct, rr in ((0, 5), (1, 1), (1, 0)):
        t.setParameter("Direct", direct); t.setParameter("RRStartDepth", rr)
        t.DoPass(True); t.synchronize()
        img = t.readAccumulator()
        ref, ref_rays = orc.render(s.view, w, h, n_passes=1, max_path_length=6, rr_start=rr, direct=direct)
        assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).mean() >= 0.985
        assert abs(t.getRaysInLastPass() - ref_rays) <= 5e-3 * ref_rays
    assert t.getParameter("MaxPathLength") == 6 and t.getParameter("RRStartDepth") == 0
    with pytest.raises(RuntimeError):
        t.setParameter("NoSuchKey", 1)
    with pytest.raises(RuntimeError):
        t.setParameter("MaxPathLength", 0)
    t.setParameter("Regularization", 1); assert t.getParameter("Regularization") == 1   # KEY_Regularization: tests/test_gpu_regularization.py
    t.setParameter("Regularization", 0)
    t.close()


def test_user_sample_tables(built_lib, orc):
    """ctl_upload_samples: caller-provided SequenceSamplerData tables (pass 3's) give the oracle's pass-3 image."""
    w = h = 64
    s, t = make("cornell", w, h, 8)
    d1, d2 = ctl.generate_sample_tables(3)
    t.uploadSamples(d1, d2)
    t.DoPass(True); t.synchronize()
    img = t.readAccumulator()
    ref, _ = orc.render(s.view, w, h, n_passes=1, pass_first=3, max_path_length=8)
    assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).mean() >= 0.99
    t.close()


# ------------------------------------------------------------------ windows, tiles, determinism, properties
def test_window_and_tiles_cover_image_exactly(built_lib):
    w, h = 200, 120
    s, t = make("cornell7", w, h, 8)
    t.DoPass(True); t.synchronize()
    whole = t.readAccumulator().copy(); rays_whole = t.getRaysInLastPass()
    # 3 interleaved parts accumulated into the same image == the whole image
    rays = 0
    for part in range(3):
        t.DoPassTiled(16, 16, part, 3, new_trace=(part == 0))
        if part > 0:
            pass
        t.synchronize(); rays += t.getRaysInLastPass()
        # passes_done advances per call; the sample stream must not: re-use pass-0 tables for the remaining parts
        if part < 2:
            t.uploadSamples(*ctl.generate_sample_tables(0))
    tiled = t.readAccumulator()
    assert rays == rays_whole
    assert np.array_equal(tiled["weight_sum"], whole["weight_sum"])
    assert np.allclose(tiled["rgb"], whole["rgb"], rtol=1e-6, atol=1e-7)
    # per-part ownership matches the host-side partition helper
    t.DoPassTiled(16, 16, 1, 3, new_trace=True); t.synchronize()
    part1 = t.readAccumulator()
    own = ctl.tile_owner(w, h, 16, 16, 3)
    inside = part1["weight_sum"] > 0
    # pixel jitter can spill a sample one pixel right/down (Appendix B #11): allow the 1-pixel fringe
    assert (inside & (own != 1)).sum() <= 0.02 * inside.sum()
    assert inside[own == 1].mean() > 0.97
    # rectangular window
    t.DoPass(True, window=(40, 20, 104, 84)); t.synchronize()
    win = t.readAccumulator()
    assert np.allclose(win["rgb"][24:80, 44:100], whole["rgb"][24:80, 44:100], rtol=1e-6, atol=1e-7)
    assert win["weight_sum"][:19].sum() == 0 and win["weight_sum"][:, :39].sum() == 0
    with pytest.raises(RuntimeError):
        t.DoPass(True, window=(0, 0, w + 1, h))
    t.close()


def test_determinism_and_progressive_linearity(built_lib):
    w = h = 128
    s, t = make("soup", w, h, 8)
    t.DoPass(True); t.synchronize(); a = t.readAccumulator().copy()
    t.DoPass(False); t.synchronize(); ab = t.readAccumulator().copy()
    t.DoPass(True); t.synchronize(); a2 = t.readAccumulator().copy()
    assert np.array_equal(a["weight_sum"], a2["weight_sum"])
    assert np.allclose(a["rgb"], a2["rgb"], rtol=1e-6, atol=1e-7)  # float atomics: only colliding (spilled) samples may reorder
    # pass 1 alone = (pass0 + pass1) - pass0
    t.uploadSamples(*ctl.generate_sample_tables(1))
    t.DoPass(True); t.synchronize(); b = t.readAccumulator().copy()
    assert np.allclose(ab["rgb"], a["rgb"] + b["rgb"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(ab["weight_sum"], a["weight_sum"] + b["weight_sum"])
    t.close()


def test_full_size_properties(built_lib):
    """BASELINE size (1920x1080, 100K triangles, depth 8): size-independent invariants instead of an oracle image."""
    w, h = 1920, 1080
    s, t = make("c2", w, h, 8)
    t.DoPass(True); t.synchronize()
    img = t.readAccumulator()
    assert img["weight_sum"].sum() == w * h                      # every path lands exactly once
    assert np.isfinite(img["rgb"]).all() and (img["rgb"] >= 0).all()
    ext, sh = t.queueSizes(8)
    assert ext[0] == w * h and np.all(np.diff(ext.astype(np.int64)) <= 0)  # queues only shrink (compaction)
    assert np.all(sh <= ext)                                    # <= 1 shadow ray per vertex
    assert t.getRaysInLastPass() == int(ext.sum()) + int(sh.sum())  # every traceRay-equivalent counts once
    t.setInstrumented(1); t.DoPass(True); t.synchronize()
    e, sc = t.visitCounts(); t.setInstrumented(0)
    assert e[3] == int(ext.sum()) and sc[3] == int(sh.sum())
    assert e[2] >= e[3] * 0.99 and e[0] > 10 * e[3]             # >= 1 instance leaf and >10 inner nodes per ray
    img2 = t.readAccumulator()
    assert np.allclose(img2["rgb"], img["rgb"], rtol=1e-6, atol=1e-6)   # instrumented build computes the same image
    t.close()


def test_error_paths(built_lib):
    t = ctl.PathTracer(32, 32)
    with pytest.raises(RuntimeError, match="no scene"):
        t.DoPass(True)
    with pytest.raises(RuntimeError, match="no scene"):
        t.intersect(np.zeros(4, api.RAY_DTYPE))
    t.close()
    with pytest.raises(RuntimeError):
        ctl.PathTracer(0, 32)
    with pytest.raises(RuntimeError, match="no such CUDA device"):
        ctl.PathTracer(32, 32, device=4096)


def test_resize_and_scene_swap(built_lib, orc):
    s, t = make("cornell", 64, 64, 8)
    t.DoPass(True); t.synchronize()
    t.Resize(48, 40)
    s2 = ctl.Scene("soup", 48, 40)
    t.InitializeScene(s2)
    t.DoPass(); t.synchronize()
    img = t.readAccumulator()
    ref, _ = orc.render(s2.view, 48, 40, n_passes=1, max_path_length=8)
    assert img.shape == (40, 48)
    assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).mean() >= 0.99
    t.close()


def test_cpp_adapter_renders(built_lib, tmp_path):
    import subprocess
    from test_abi_cpu import _build_adapter_check
    exe = _build_adapter_check(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "passes 1" in r.stdout and "weight 4096" in r.stdout


# ------------------------------------------------------------------ golden vectors minted from the reference's own host code
import os as _os
_GOLD = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "reference_golden.npz"))


@pytest.mark.parametrize("kind", ["cornell", "cornell7", "soup"])
def test_trace_rays_golden(built_lib, kind):
    """CUDA traversal vs results of the REFERENCE's traceRay (oracle/_ref, tests/golden/make_golden.py)."""
    s, t = make(kind)
    rays = np.ascontiguousarray(_GOLD[f"trace_{kind}_rays"]).view(api.RAY_DTYPE).reshape(-1)
    ref = np.ascontiguousarray(_GOLD[f"trace_{kind}_res"]).view(api.TRACE_RESULT_DTYPE).reshape(-1)
    got = t.trace_rays(rays)
    assert np.array_equal(got["tri_idx"], ref["tri_idx"]) and np.array_equal(got["node_idx"], ref["node_idx"])
    hit = ref["tri_idx"] != 0xffffffff
    assert np.all(np.abs(got["dist"][hit] - ref["dist"][hit]) <= 4e-6 * np.maximum(1, ref["dist"][hit]))  # explicit FMA vs the reference's uncontracted host build
    assert np.all(np.abs(got["u"][hit] - ref["u"][hit]) <= 2e-5) and np.all(np.abs(got["v"][hit] - ref["v"][hit]) <= 2e-5)
    t.close()


@pytest.mark.parametrize("key,kind,w,h,spp", [("image_cornell_128x128_1spp", "cornell", 128, 128, 1), ("image_cornell7_96x96_8spp", "cornell7", 96, 96, 8),
                                              ("image_soup_96x96_1spp", "soup", 96, 96, 1)])
def test_render_golden(built_lib, key, kind, w, h, spp):
    """CUDA path tracer vs PixelData images rendered by the REFERENCE's PathTrace<true> on its CPU path, same seed."""
    ref = np.ascontiguousarray(_GOLD[key]).view(api.PIXEL_DTYPE).reshape(h, w)
    s, t = make(kind, w, h, 8)
    for p in range(spp):
        t.DoPass(p == 0)
    t.synchronize()
    img = t.readAccumulator()
    a, b = img["rgb"], ref["rgb"]
    assert (rel_l2(a, b) <= 1e-3).mean() >= 0.99
    assert np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()) <= (1e-2 if spp == 1 else 3e-3)
    assert abs(a.mean() - b.mean()) <= 1e-3 * b.mean()
    assert np.array_equal(img["weight_sum"], ref["weight_sum"])
    t.close()


def test_config1_cornell_256_golden(built_lib):
    s, t = make("cornell", 256, 256, 8)
    t.DoPass(True); t.synchronize()
    img = t.readAccumulator()
    st = _GOLD["config1_cornell_256_1spp_stats"]
    assert abs(img["rgb"].astype(np.float64).mean() - st[0]) <= 1e-3 * st[0]
    assert img["weight_sum"].sum() == st[2] and abs(t.getRaysInLastPass() - int(st[3])) <= 2e-3 * st[3]
    assert np.allclose(img["rgb"].astype(np.float64).mean(axis=(1, 2)), _GOLD["config1_cornell_256_1spp_rowmeans"], rtol=2e-3, atol=1e-5)
    t.close()


# ------------------------------------------------------------------ device-side sample tables, fused passes
def test_device_sample_tables_bit_exact(built_lib):
    """k_gen_tables (XORWOW with GF(2) jump-ahead, one thread per sequence) == the host generator == the reference (golden)."""
    s, t = make("cornell", 32, 32, 4)
    t.DoPasses(3, new_trace=True); t.synchronize()
    for k in range(3):
        d1, d2 = t.readSampleTables(k)
        h1, h2 = ctl.generate_sample_tables(k)
        assert np.array_equal(d1.view(np.uint32), h1.view(np.uint32)) and np.array_equal(d2.view(np.uint32), h2.view(np.uint32))
    t.DoPasses(2); t.synchronize()            # continues the stream: passes 3, 4
    d1, d2 = t.readSampleTables(1)
    h1, h2 = ctl.generate_sample_tables(4)
    assert np.array_equal(d1.view(np.uint32), h1.view(np.uint32)) and np.array_equal(d2.view(np.uint32), h2.view(np.uint32))
    t.DoPass(False); t.synchronize()           # pass 5
    assert np.array_equal(t.readSampleTables(0)[0][:8192].view(np.uint32), _GOLD["tables_p5_d1_head"].view(np.uint32))
    # host-generated tables (reference UpdateKernel behaviour) give the same tables
    t.setParameter("DeviceSampleTables", 0)
    t.DoPasses(2, new_trace=True); t.synchronize()
    assert np.array_equal(t.readSampleTables(1)[1].view(np.uint32), ctl.generate_sample_tables(1)[1].view(np.uint32))
    t.close()


def test_fused_passes_equal_sequential_passes(built_lib, orc):
    w, h = 160, 96
    s, t = make("soup", w, h, 8)
    for p in range(6):
        t.DoPass(p == 0)
    t.synchronize(); seq = t.readAccumulator().copy(); rays_seq = t.getTotalRays()
    r0 = t.getTotalRays()
    t.DoPasses(6, new_trace=True); t.synchronize(); fused = t.readAccumulator().copy()
    assert t.getTotalRays() - r0 == rays_seq and t.getNumPassesDone() == 6
    assert np.array_equal(fused["weight_sum"], seq["weight_sum"])
    assert np.allclose(fused["rgb"], seq["rgb"], rtol=2e-6, atol=1e-6)        # same paths; only the float-atomic order differs
    # 2 + 4 split with host-generated tables (the reference's UpdateKernel behaviour): still the same image
    t.setParameter("DeviceSampleTables", 0)
    t.DoPasses(2, new_trace=True); t.DoPasses(4); t.synchronize(); split = t.readAccumulator()
    assert np.array_equal(split["weight_sum"], seq["weight_sum"]) and np.allclose(split["rgb"], seq["rgb"], rtol=2e-6, atol=1e-6)
    t.close()
    s2, t2 = make("cornell7", 96, 96, 8)
    t2.DoPasses(8, new_trace=True); t2.synchronize()
    img = t2.readAccumulator()
    ref = np.ascontiguousarray(_GOLD["image_cornell7_96x96_8spp"]).view(api.PIXEL_DTYPE).reshape(96, 96)
    assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).mean() >= 0.99 and np.array_equal(img["weight_sum"], ref["weight_sum"])
    t2.close()


def test_ray_sorting_preserves_results(built_lib):
    """SortMode=1 (counting sort of every bounce's extension queue by direction octant + origin cell) reorders work only."""
    w, h = 192, 108
    for kind in ("soup", "c3"):
        s, t = make(kind, w, h, 8)
        t.DoPasses(2, new_trace=True); t.synchronize(); a = t.readAccumulator().copy(); ra = t.getRaysInLastPass(); qa = t.queueSizes(8)
        t.setParameter("SortMode", 1)
        t.DoPasses(2, new_trace=True); t.synchronize(); b = t.readAccumulator().copy(); rb = t.getRaysInLastPass(); qb = t.queueSizes(8)
        assert ra == rb and np.array_equal(qa[0], qb[0]) and np.array_equal(qa[1], qb[1])
        assert np.array_equal(a["weight_sum"], b["weight_sum"])
        assert np.allclose(a["rgb"], b["rgb"], rtol=2e-6, atol=1e-6)
        # the sorted queue really is sorted: capture bounce 3 and check the keys are non-decreasing
        t.setParameter("CaptureBounce", 3)
        t.DoPass(True); t.synchronize()
        rays = t.capturedRays(w * h)
        t.setParameter("CaptureBounce", 0)
        lo = np.array(list(s.view.box_min), np.float32); hi = np.array(list(s.view.box_max), np.float32)
        inv = (np.float32(1) / (hi - lo)).astype(np.float32)
        c = np.clip(((rays["o"] - lo) * inv * np.float32(32)), 0, 31).astype(np.uint32)

        def spread(x):
            x = x & 31; x = (x | (x << 8)) & 0x100f; x = (x | (x << 4)) & 0x10c3; x = (x | (x << 2)) & 0x1249; return x
        octant = (rays["d"][:, 0] < 0).astype(np.uint32) | ((rays["d"][:, 1] < 0).astype(np.uint32) << 1) | ((rays["d"][:, 2] < 0).astype(np.uint32) << 2)
        key = (octant << 15) | spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)
        assert len(rays) > 1000 and np.all(np.diff(key.astype(np.int64)) >= 0)
        t.close()


def test_resolve_srgb8(built_lib, orc):
    """ctl_resolve_srgb8 (== applyImagePipeline without filter / post-process) vs the oracle on the SAME accumulator:
    identical bytes except where CUDA powf and libm powf straddle a 1/255 step (<= 1 LSB, < 0.5 % of the channels)."""
    w, h = 160, 120
    s, t = make("cornell7", w, h, 8)
    t.DoPasses(4, new_trace=True); t.synchronize()
    acc = t.readAccumulator()
    got = t.resolveSRGB8()
    ref = orc.resolve_srgb8(acc)
    assert got.shape == (h, w, 4) and np.all(got[:, :, 3] == 255)
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.005
    assert got[:, :, :3].mean() > 20          # a lit image, not black
    t.close()


def test_fused_traversal_launch_is_exact(built_lib):
    """FuseTraversal (shadow rays of bounce b + extension rays of bounce b+1 in one launch) changes scheduling only."""
    w, h = 200, 120
    for kind, depth in (("cornell7", 8), ("c3", 5), ("soup", 1)):
        s, t = make(kind, w, h, depth)
        t.setParameter("FuseTraversal", 0); t.DoPasses(2, new_trace=True); t.synchronize(); a = t.readAccumulator().copy(); ra = t.getRaysInLastPass()
        t.setParameter("FuseTraversal", 1); t.DoPasses(2, new_trace=True); t.synchronize(); b = t.readAccumulator().copy(); rb = t.getRaysInLastPass()
        assert ra == rb and np.array_equal(a["weight_sum"], b["weight_sum"])
        assert np.allclose(a["rgb"], b["rgb"], rtol=2e-6, atol=1e-6)
        assert (a["rgb"] == b["rgb"]).all(axis=2).mean() > 0.99   # per-path arithmetic is identical; only atomics into spill pixels reorder
        t.close()


def test_two_light_scene_from_mesh(built_lib, orc):
    """ctl_scene_create_from_mesh scene with two area lights / two-sided card / GGX block: CUDA vs oracle and vs the
    image rendered by the reference's own PathTrace<true> (golden)."""
    from scene_fixtures import two_light_room
    w = h = 96
    s = two_light_room(w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    t.DoPasses(4, new_trace=True); t.synchronize()
    img = t.readAccumulator()
    o, orays = orc.render(s.view, w, h, n_passes=4, max_path_length=6)
    ref = np.ascontiguousarray(_GOLD["image_two_light_96x96_4spp"]).view(api.PIXEL_DTYPE).reshape(h, w)
    for other in (o, ref):
        assert (rel_l2(img["rgb"], other["rgb"]) <= 1e-3).mean() >= 0.99
        assert np.array_equal(img["weight_sum"], other["weight_sum"])
    assert abs(t.getRaysInLastPass() - orays) <= 2e-3 * orays      # the 4-pass wavefront traced what the oracle traced
    rays = random_rays(s, 3000, seed=4)
    g = t.trace_rays(rays); oo = orc.trace_rays(s.view, rays)
    assert np.array_equal(g["tri_idx"], oo["tri_idx"]) and np.array_equal(g["dist"].view(np.uint32), oo["dist"].view(np.uint32))
    t.close()


# ------------------------------------------------------------------ GPU BVH build
def _walk_reference_bvh(nodes_f32, tri_index, n_slots):
    """Walk a reference-layout mesh BVH; returns (#inner nodes, leaf runs [(first, count)])."""
    nodes_u = nodes_f32.view(np.uint32).reshape(-1, 16)
    leaves, inner, stack = [], 0, [0]
    while stack:
        a = stack.pop(); assert a % 4 == 0
        inner += 1
        for ci in range(2):
            child = int(nodes_u[a // 4, 12 + ci]); child = child - (1 << 32) if child >= 0x80000000 else child
            if child == 0x76543210:
                continue
            if child < 0:
                first = ~child; k = first
                while not (tri_index[k] & 1):
                    k += 1
                leaves.append((first, k - first + 1))
            else:
                stack.append(child)
    return inner, leaves


@pytest.mark.parametrize("builder", ["ploc", "lbvh"])
@pytest.mark.parametrize("kind", ["cornell7", "soup", "c2"])
def test_gpu_bvh_build_matches_cpu_bvh_results(built_lib, orc, kind, builder, monkeypatch):
    """ctl_scene_rebuild_bvh_gpu: the agglomerative builder (default) and the LBVH, built on the device in the reference layout.  Different trees,
    same data surface: every triangle referenced exactly once, leaves <= 8, and closest hits / images identical to the CPU-built (SAH) tree."""
    monkeypatch.setenv("CTL_GPU_BUILDER", builder)
    w, h = 192, 108
    s_cpu = ctl.Scene(kind, w, h); s_gpu = ctl.Scene(kind, w, h)
    ms = s_gpu.rebuildBVHOnGPU()
    s_gpu.validate()                                                                  # child / leaf references, tree shape, depth within the traversal stack
    assert ms > 0 and s_gpu.view.n_woop == s_gpu.view.n_tri_index >= s_gpu.n_triangles and s_cpu.view.n_woop >= s_cpu.n_triangles   # split trees (CPU: spatial splits, GPU: pre-split slivers) may reference a triangle from several leaves
    assert s_gpu.view.n_woop <= 4 * s_gpu.n_triangles + 8                              # CTL_GPU_SPLIT budget (default: at most four times the triangles of a mesh)
    meshes = s_gpu.array("meshes"); tri_index = s_gpu.array("tri_index")[:, 0]; bvh = s_gpu.array("bvh_nodes")
    for mi, m in enumerate(meshes):
        node_off4, idx_off = int(m[1]), int(m[3])
        n_slots = (int(meshes[mi + 1][3]) if mi + 1 < len(meshes) else s_gpu.view.n_tri_index) - idx_off
        inner, leaves = _walk_reference_bvh(bvh[node_off4 // 4:], tri_index[idx_off:idx_off + n_slots], n_slots)
        covered = np.zeros(n_slots, np.int32)
        for first, cnt in leaves:
            assert 1 <= cnt <= 8
            covered[first:first + cnt] += 1
        assert np.all(covered == 1)                                                   # slots partitioned by the leaves
        tri_ids = np.unique(tri_index[idx_off:idx_off + n_slots] >> 1)
        n_mesh_tris = (int(meshes[mi + 1][0]) if mi + 1 < len(meshes) else s_gpu.n_triangles) - int(m[0])
        assert np.array_equal(tri_ids, np.arange(n_mesh_tris))                         # every triangle of the mesh is referenced (exactly once unless pre-split)
    # the oracle traverses the GPU-built tree and the CPU-built tree to the same hits
    rays = random_rays(s_cpu, 6000, seed=21)
    a = orc.trace_rays(s_cpu.view, rays); b = orc.trace_rays(s_gpu.view, rays)
    same = (a["tri_idx"] == b["tri_idx"]) & (a["dist"].view(np.uint32) == b["dist"].view(np.uint32))
    assert same.mean() >= 0.9995                                                      # exact ties may pick the other triangle
    # and so does the CUDA path, including the light sampling that points into the re-ordered Woop slots
    t1 = ctl.PathTracer(w, h); t1.InitializeScene(s_cpu); t1.setParameter("MaxPathLength", 6); t1.DoPasses(2, new_trace=True); t1.synchronize()
    t2 = ctl.PathTracer(w, h); t2.InitializeScene(s_gpu); t2.setParameter("MaxPathLength", 6); t2.DoPasses(2, new_trace=True); t2.synchronize()
    g = t2.trace_rays(rays)
    assert np.array_equal(g["tri_idx"], b["tri_idx"]) and np.array_equal(g["dist"].view(np.uint32), b["dist"].view(np.uint32))
    i1, i2 = t1.readAccumulator(), t2.readAccumulator()
    assert np.array_equal(i1["weight_sum"], i2["weight_sum"])
    assert (rel_l2(i2["rgb"], i1["rgb"]) <= 1e-4).mean() >= 0.999
    t1.close(); t2.close()


def _depth(nodes):
    d, st = 0, [(0, 1)]
    u = nodes.view(np.int32)
    while st:
        i, k = st.pop(); d = max(d, k)
        for c in (int(u[i, 12]), int(u[i, 13])):
            if c >= 0 and c != 0x76543210: st.append((c // 4, k + 1))
    return d


@pytest.mark.parametrize("algorithm", [1, 0, 2])
def test_gpu_bvh_build_api_edge_cases(built_lib, orc, algorithm):
    L = built_lib
    L.ctl_bvh_build_gpu_ex.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for n in (1, 2, 8, 9, 10, 17, 300, 5000):
        rng = np.random.default_rng(n)
        verts = rng.uniform(-1, 1, size=(n, 9)).astype(np.float32)
        if n == 5000:   # duplicated / degenerate geometry: 2 000 copies of one triangle, 1 000 zero-area triangles at one point, the rest random
            verts[:2000] = verts[0]; verts[2000:3000] = 0.25
        nodes = np.zeros((max(n, 1), 16), np.float32); woop = np.zeros((n, 12), np.float32); index = np.zeros(n, np.uint32)
        nn = C.c_uint32(0); ms = C.c_float(0)
        assert L.ctl_bvh_build_gpu_ex(0, verts.ctypes.data, n, algorithm, 0, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, C.byref(ms)) == 0
        assert _depth(nodes[:nn.value]) <= 56                                           # equal boxes pair up (2k, 2k+1): no merge chains
        inner, leaves = _walk_reference_bvh(nodes[:nn.value], index, n)
        assert inner == nn.value and sum(c for _, c in leaves) == n and all(c <= 8 for _, c in leaves)
        if n <= 8:   # single-leaf mesh: root = {~0, sentinel} (SplitBVHBuilder.cpp:176-189)
            assert nn.value == 1 and nodes.view(np.uint32)[0, 12] == 0xffffffff and nodes.view(np.uint32)[0, 13] == 0x76543210
        for s_ in range(n):  # Woop data of every slot == the host encoder on that triangle's vertices, bit for bit
            t = index[s_] >> 1
            ref = orc.encode_woop(verts[t, 0:3], verts[t, 3:6], verts[t, 6:9])
            assert np.array_equal(woop[s_].view(np.uint32), ref.view(np.uint32)) or (np.isnan(ref).all() and np.isnan(woop[s_]).all())   # zero-area triangle: NaN both (payload bits differ host / device)
    assert L.ctl_bvh_build_gpu(0, None, 0, None, None, None, None, None) != 0
    assert L.ctl_bvh_build_gpu_ex(0, verts.ctypes.data, 8, 3, 0, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, None) != 0   # unknown algorithm


@pytest.mark.parametrize("algorithm", [1, 0])
def test_gpu_bvh_presplit_slivers(built_lib, orc, algorithm):
    """ctl_bvh_build_gpu_split on a mesh of long thin diagonal triangles (what the foliage of configs[3] is made of): references multiply within the budget,
    every slot holds the Woop record of its triangle, every triangle is referenced, the tree is a valid reference-layout tree, and -- through a scene built
    from the same mesh -- split and unsplit trees give the same hits on the device; the split tree needs fewer node visits."""
    L = built_lib
    L.ctl_bvh_build_gpu_split.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(11); n = 4000
    a = rng.uniform(-1, 1, size=(n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    w = np.cross(d, rng.normal(size=(n, 3))); w /= np.linalg.norm(w, axis=1, keepdims=True)
    verts = np.concatenate([a, a + 0.8 * d, a + 0.4 * d + 0.02 * w], axis=1).astype(np.float32)     # 40:1 slivers
    verts[:50, 3:6] = verts[:50, 0:3] + np.float32(0.01); verts[:50, 6:9] = verts[:50, 0:3] + np.array([0.01, 0, 0], np.float32)   # and some small fat ones
    out = {}
    for growth in (0.0, 1.0, 3.0):
        cap = int(n * (1 + growth)) + 1
        nodes = np.zeros((cap, 16), np.float32); woop = np.zeros((cap, 12), np.float32); index = np.zeros(cap, np.uint32); nn = C.c_uint32(0); ns = C.c_uint32(0)
        assert L.ctl_bvh_build_gpu_split(0, verts.ctypes.data, n, algorithm, 0, growth, cap, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, C.byref(ns), None) == 0, L.ctl_last_error()
        assert n <= ns.value <= cap and (growth > 0) == (ns.value > n)
        inner, leaves = _walk_reference_bvh(nodes[:nn.value], index[:ns.value], ns.value)
        covered = np.zeros(ns.value, np.int32)
        for first, cnt in leaves:
            assert 1 <= cnt <= 8; covered[first:first + cnt] += 1
        assert inner == nn.value and np.all(covered == 1) and _depth(nodes[:nn.value]) <= 60
        assert np.array_equal(np.unique(index[:ns.value] >> 1), np.arange(n))
        for s_ in rng.integers(0, ns.value, 300):
            t = index[s_] >> 1
            assert np.array_equal(woop[s_].view(np.uint32), orc.encode_woop(verts[t, 0:3], verts[t, 3:6], verts[t, 6:9]).view(np.uint32))
        out[growth] = ns.value
    assert out[3.0] >= out[1.0] > out[0.0] == n
    assert L.ctl_bvh_build_gpu_split(0, verts.ctypes.data, n, algorithm, 0, 9.0, 10 * n, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, C.byref(ns), None) != 0   # growth out of range
    assert L.ctl_bvh_build_gpu_split(0, verts.ctypes.data, n, algorithm, 0, 1.0, n - 1, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, C.byref(ns), None) != 0    # capacity below n_tris


def test_gpu_bvh_presplit_scene_hits(built_lib, orc, monkeypatch):
    """The 1 M-triangle scene's foliage mesh through ctl_scene_rebuild_bvh_gpu with and without pre-splitting: identical closest hits (the same Woop tests
    decide them), fewer inner-node visits with the split tree."""
    w, h = 96, 54
    rays = None; res = {}
    for split in ("0", "1"):
        monkeypatch.setenv("CTL_GPU_SPLIT", "0" if split == "0" else "2")
        s = ctl.Scene("c4", w, h); s.setRebraid(0); s.rebuildBVHOnGPU(); s.validate()
        if rays is None: rays = random_rays(s, 20000, seed=5)
        t = ctl.PathTracer(w, h); t.InitializeScene(s)
        g, counts = t.trace_rays(rays, counts=True)
        o = orc.trace_rays(s.view, rays)
        assert np.array_equal(g["tri_idx"], o["tri_idx"]) and np.array_equal(g["dist"].view(np.uint32), o["dist"].view(np.uint32))
        res[split] = (g, counts, s.view.n_woop); t.close()
    same = (res["0"][0]["tri_idx"] == res["1"][0]["tri_idx"]) & (res["0"][0]["dist"].view(np.uint32) == res["1"][0]["dist"].view(np.uint32))
    print("pre-split: slots", res["0"][2], "->", res["1"][2], "inner-node visits", res["0"][1][0], "->", res["1"][1][0], "agreement", same.mean())
    assert same.mean() >= 0.9995 and res["1"][2] > res["0"][2]
    assert res["1"][1][0] < res["0"][1][0]


def test_gpu_bvh_build_is_deterministic(built_lib):
    """The agglomerative builder hands out node ids by rank (scan), not by atomics: two builds of one mesh are byte-identical."""
    L = built_lib
    rng = np.random.default_rng(5); n = 20000
    c = rng.uniform(-1, 1, size=(n, 1, 3)); verts = (c + rng.normal(scale=0.02, size=(n, 3, 3))).reshape(n, 9).astype(np.float32)
    outs = []
    for _ in range(2):
        nodes = np.zeros((n, 16), np.float32); woop = np.zeros((n, 12), np.float32); index = np.zeros(n, np.uint32); nn = C.c_uint32(0)
        assert L.ctl_bvh_build_gpu(0, verts.ctypes.data, n, nodes.ctypes.data, C.byref(nn), woop.ctypes.data, index.ctypes.data, None) == 0
        outs.append((nn.value, nodes[:nn.value].tobytes(), index.tobytes()))
    assert outs[0] == outs[1]


def test_material_sort_preserves_results(built_lib):
    """SortMode=2 (hit queue grouped by material class before shading) reorders work only."""
    w, h = 192, 108
    for kind in ("soup", "c3", "cornell"):
        s, t = make(kind, w, h, 8)
        t.DoPasses(2, new_trace=True); t.synchronize(); a = t.readAccumulator().copy(); ra = t.getRaysInLastPass(); qa = t.queueSizes(8)
        t.setParameter("SortMode", 2)
        t.DoPasses(2, new_trace=True); t.synchronize(); b = t.readAccumulator().copy(); rb = t.getRaysInLastPass(); qb = t.queueSizes(8)
        assert ra == rb and np.array_equal(qa[0], qb[0]) and np.array_equal(qa[1], qb[1])
        assert np.array_equal(a["weight_sum"], b["weight_sum"])
        assert np.allclose(a["rgb"], b["rgb"], rtol=2e-6, atol=1e-6)
        t.close()


def test_edge_cases_small_and_degenerate(built_lib, orc):
    """1x1 and odd-sized images, MaxPathLength 1, zero-area windows, a scene without lights, rays that miss everything."""
    # odd sizes, depth 1 (camera ray + one NEE only)
    for (w, h, depth) in ((1, 1, 8), (33, 7, 1), (65, 3, 2)):
        s, t = make("cornell", w, h, depth)
        t.DoPass(True); t.synchronize()
        img = t.readAccumulator()
        ref, rays = orc.render(s.view, w, h, n_passes=1, max_path_length=depth)
        assert np.array_equal(img["weight_sum"], ref["weight_sum"])
        assert np.allclose(img["rgb"], ref["rgb"], rtol=2e-3, atol=1e-5)
        assert t.getRaysInLastPass() == rays
        t.DoPass(False, window=(0, 0, 0, 0)); t.synchronize()          # zero-area window: nothing happens, no error
        assert np.array_equal(t.readAccumulator()["weight_sum"], ref["weight_sum"])
        t.close()
    # no lights: every path contributes exactly zero, weights still accumulate
    from scene_fixtures import _mat
    V = np.array([(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)], np.float32)
    s = ctl.Scene.from_mesh(V, np.array([0, 2, 1, 0, 3, 2], np.uint32), np.array([0, 0], np.uint8), [_mat()], np.zeros((1, 3), np.float32),
                            (0, 1.5, -2.0), (0, 0, 0), (0, 1, 0), 60.0, 40, 30)
    assert s.view.num_lights == 0
    t = ctl.PathTracer(40, 30); t.InitializeScene(s); t.setParameter("MaxPathLength", 4)
    t.DoPasses(2, new_trace=True); t.synchronize()
    img = t.readAccumulator()
    assert np.all(img["rgb"] == 0) and img["weight_sum"].sum() == 2 * 40 * 30
    ref, rays = orc.render(s.view, 40, 30, n_passes=2, max_path_length=4)
    assert t.getRaysInLastPass() == rays and np.array_equal(img["weight_sum"], ref["weight_sum"])
    # rays that start outside and point away: all miss, API returns the miss encoding
    r = np.zeros(64, api.RAY_DTYPE); r["o"] = (0, 50, 0); r["d"] = (0, 1, 0); r["tmax"] = 1e30
    assert np.all(t.intersect(r)["tri_idx"] == -1) and np.all(t.trace_rays(r)["tri_idx"] == 0xffffffff)
    t.close()


def test_filtered_resolve_golden(built_lib, orc):
    """ctl_resolve_filtered_srgb8 (applyImagePipeline with a reconstruction filter, the call of the reference's example main) on
    the reference-rendered accumulator: bytes equal to the reference's own pipeline except <= 1 LSB where CUDA powf/expf and libm
    straddle a quantisation step."""
    import torch
    acc = np.ascontiguousarray(_GOLD["pipeline_accum_cornell_80x64_4spp"])
    t = ctl.PathTracer(80, 64)
    d_acc = torch.from_numpy(acc.reshape(-1).copy()).cuda()
    t.setAccumDevicePtr(d_acc.data_ptr())
    got = t.resolveSRGB8()
    d = np.abs(got.astype(np.int32) - _GOLD["pipeline_resolve_default"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.005
    names = {0: "box", 1: "gaussian", 2: "triangle"}
    for k, (ft, xw, yw, a) in enumerate(_GOLD["pipeline_filter_cases"]):
        got = t.resolveFilteredSRGB8(names[int(ft)], float(xw), float(yw), float(a))
        ref = _GOLD[f"pipeline_resolve_filter{k}"]
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 2 and (d > 0).mean() < 0.01, (k, d.max(), (d > 0).mean())   # RGBE mantissa step +- sRGB step
    with pytest.raises(RuntimeError):
        t.resolveFilteredSRGB8("box", 0.0, 0.5)
    t.setAccumDevicePtr(0)
    t.close()


def test_cpp_example_driver(built_lib, tmp_path):
    """examples/main.cpp (the reference's example main on this backend) builds with plain g++ and renders."""
    import subprocess
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    libdir = _os.path.dirname(api.LIB_PATH)
    exe = str(tmp_path / "ctl_render"); out = str(tmp_path / "r.ppm")
    r = subprocess.run(["g++", "-std=c++17", "-O1", _os.path.join(root, "examples", "main.cpp"), "-I" + _os.path.join(root, "include"), "-L" + libdir, "-lctl_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    obj = _os.path.join(root, "tests", "golden", "obj", "room.obj")
    for args, tag in ((["cornell7", "4", "96x64", out], "PT, 4 passes"), ([obj, "3", "PT_Wave", "tonemap", "96x64", out], "PT_Wave, 3 passes")):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        data = open(out, "rb").read()
        assert data.startswith(b"P6\n96 64\n255\n") and len(data) == len(b"P6\n96 64\n255\n") + 96 * 64 * 3
        px = np.frombuffer(data[len(b"P6\n96 64\n255\n"):], np.uint8)
        assert px.mean() > 20 and tag in r.stdout, r.stdout
    assert subprocess.run([exe, "no_such_thing"], capture_output=True, text=True).returncode == 2
