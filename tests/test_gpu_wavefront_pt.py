"""GPU parity tests of the WavefrontPathTracer drop-in: ctl_wavefront_pass vs the CPU restatement
(oracle/oracle.cpp orc_render_wavefront, itself bit-identical to the reference's own pathIterateKernel + DoubleRayBuffer, see
tests/test_golden_cpu.py) and vs the goldens minted from the reference's own code (oracle/_ref).

What is exact and what is toleranced: the queue evolution (primary / secondary rays per iteration), the ray count and the
sample weights are integers -> equal.  Radiance: same seed, same pass, same queue slots -> per-pixel relative L2 <= 1e-3 on
>= 99 % of the pixels (libm vs libdevice transcendentals); these fixed small cases contain no flipped discrete decision (a flip
would shift every later queue slot and with it the slot-keyed random numbers -- the reference's own GPU build does not even
reproduce itself run to run for that reason)."""
import os

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))


def _scene(kind, w, h, **kw):
    if kind == "two_light":
        from scene_fixtures import two_light_room
        return two_light_room(w, h)
    return ctl.Scene(kind, w, h, **kw)


def _tracer(s, w, h, mpl, rr=5, direct=1):
    t = ctl.WavefrontPathTracer(w, h)
    t.InitializeScene(s)
    t.setParameter("MaxPathLength", mpl); t.setParameter("RRStartDepth", rr); t.setParameter("Direct", direct)
    return t


def _queues(t, mpl):
    """(mpl, 2) like the oracle: primary rays intersected before iteration i, secondary rays intersected before iteration i."""
    e, sh = t.queueSizes(mpl)
    q = np.zeros((mpl, 2), np.uint32); q[:, 0] = e; q[1:, 1] = sh[:-1]
    return q


def rel_l2(a, b):
    return np.linalg.norm(a - b, axis=-1) / (np.linalg.norm(b, axis=-1) + 1e-3)


@pytest.mark.parametrize("kind,w,h,spp,mpl,rr,direct", [("cornell7", 64, 64, 2, 8, 5, 1), ("soup", 96, 64, 1, 8, 5, 0), ("two_light", 64, 64, 2, 6, 3, 1),
                                                        ("cornell", 128, 128, 1, 12, 2, 1), ("c3", 160, 90, 1, 8, 5, 1), ("cornell", 33, 17, 3, 50, 5, 1)])
def test_wavefront_pass_matches_oracle(built_lib, orc, kind, w, h, spp, mpl, rr, direct):
    s = _scene(kind, w, h)
    t = _tracer(s, w, h, mpl, rr, direct)
    for p in range(spp):
        t.DoPass(p == 0)
    t.synchronize()
    img = t.readAccumulator()
    ref, ref_rays, q = orc.render_wavefront(s.view, w, h, n_passes=spp, max_path_length=mpl, rr_start=rr, direct=direct)
    assert np.array_equal(_queues(t, mpl), q), (_queues(t, mpl).tolist(), q.tolist())      # last pass: identical queue evolution
    assert t.getTotalRays() == ref_rays and t.getNumPassesDone() == spp
    assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and (img["weight_sum"] == spp).all()
    r = rel_l2(img["rgb"], ref["rgb"])
    assert (r <= 1e-3).mean() >= 0.99, (r <= 1e-3).mean()
    assert abs(img["rgb"].mean() - ref["rgb"].mean()) <= 1e-3 * ref["rgb"].mean()
    t.close()


def test_wavefront_pass_vs_reference_goldens(built_lib):
    """Against images produced by the reference's OWN pathIterateKernel + DoubleRayBuffer (oracle/_ref, host arithmetic)."""
    kinds = ["cornell7", "soup", "two_light", "cornell"]
    for k, kind in enumerate(kinds):
        w, h, spp, mpl, rr, direct = (int(v) for v in GOLD["wpt_cases"][k])
        ref = np.ascontiguousarray(GOLD[f"wpt_image_{k}_{kind}"]).view(api.PIXEL_DTYPE).reshape(h, w)
        s = _scene(kind, w, h)
        t = _tracer(s, w, h, mpl, rr, direct)
        for p in range(spp):
            t.DoPass(p == 0)
        img = t.readAccumulator()
        r = rel_l2(img["rgb"], ref["rgb"])
        assert (r <= 1e-3).mean() >= 0.99, (kind, (r <= 1e-3).mean())
        assert np.array_equal(img["weight_sum"], ref["weight_sum"])
        assert abs(t.getTotalRays() - int(GOLD[f"wpt_rays_{k}_{kind}"][0])) <= 0.01 * t.getTotalRays()
        t.close()


def test_wavefront_is_deterministic_and_order_preserving(built_lib, orc):
    """Full-size frame (config-2 shape): two runs are bit-identical (the reference's atomics make its own runs differ), every pixel
    receives exactly one sample per pass, and the queue shrinks monotonically.  Size-independent properties only: the oracle
    needs minutes at this size."""
    w, h, mpl = 1920, 1080, 8
    s = _scene("c2", w, h)
    t = _tracer(s, w, h, mpl)
    imgs = []
    for _ in range(2):
        t.DoPass(True); t.DoPass(False)
        imgs.append(t.readAccumulator().copy())
    assert np.array_equal(imgs[0]["rgb"].view(np.uint32), imgs[1]["rgb"].view(np.uint32))
    assert (imgs[0]["weight_sum"] == 2).all() and np.isfinite(imgs[0]["rgb"]).all() and (imgs[0]["rgb"] >= 0).all()
    e, sh = t.queueSizes(mpl)
    assert e[0] == w * h and (np.diff(e.astype(np.int64)) <= 0).all() and (sh <= e).all() and sh[-1] == 0
    assert t.getRaysInLastPass() == int(e.sum()) + int(sh.sum())
    # both integrators estimate the same image: means agree within Monte-Carlo noise of 2 spp at 2 M pixels
    p = ctl.PathTracer(w, h); p.InitializeScene(s); p.setParameter("MaxPathLength", mpl)
    p.DoPasses(2, new_trace=True); pm = p.readAccumulator()["rgb"].astype(np.float64).mean()
    assert abs(imgs[0]["rgb"].astype(np.float64).mean() - pm) <= 0.02 * pm
    t.close(); p.close()


def test_wavefront_progressive_passes_and_new_trace(built_lib, orc):
    s = _scene("cornell7", 48, 48)
    t = _tracer(s, 48, 48, 6)
    t.DoPass(True); a1 = t.readAccumulator().copy()
    t.DoPass(False); a2 = t.readAccumulator().copy()
    ref1, _, _ = orc.render_wavefront(s.view, 48, 48, n_passes=1, max_path_length=6)
    ref2only, _, _ = orc.render_wavefront(s.view, 48, 48, n_passes=1, pass_first=1, max_path_length=6)
    assert (rel_l2(a1["rgb"], ref1["rgb"]) <= 1e-3).mean() >= 0.99
    assert (rel_l2(a2["rgb"] - a1["rgb"], ref2only["rgb"]) <= 2e-3).mean() >= 0.99      # the second pass alone (iterationIdx = 2)
    t.DoPass(True); b1 = t.readAccumulator()
    assert np.array_equal(a1["rgb"].view(np.uint32), b1["rgb"].view(np.uint32))             # new trace restarts the sample stream
    with pytest.raises(ValueError):
        t.DoPass(False, window=(0, 0, 8, 8))
    t.close()


def test_wavefront_instrumented_pass(built_lib, orc):
    """The counting build of the traversal (roofline input) leaves image and queues unchanged and counts every intersected ray."""
    s = _scene("soup", 80, 48)
    t = _tracer(s, 80, 48, 6)
    t.DoPass(True); a = t.readAccumulator().copy(); qa = _queues(t, 6)
    t.setInstrumented(1); t.DoPass(True); b = t.readAccumulator().copy(); qb = _queues(t, 6)
    e_cnt, s_cnt = t.visitCounts(); t.setInstrumented(0)
    assert np.array_equal(a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)) and np.array_equal(qa, qb)
    assert e_cnt[3] == int(qa[:, 0].sum()) and s_cnt[3] == int(qa[:, 1].sum()) and e_cnt[0] > e_cnt[3] and e_cnt[2] >= e_cnt[3]   # >= 1 inner node, >= 1 instance per ray
    # and the counts are the oracle's for the same queues: primaries as closest-hit queries, secondaries in their any-hit form
    _, _, qo, cnt = orc.render_wavefront(s.view, 80, 48, n_passes=1, max_path_length=6, counts=True)
    assert np.array_equal(qo, qa) and list(e_cnt) == cnt[:4]                       # primaries: identical visit counts
    assert s_cnt[3] == cnt[7] and all(abs(a - b) <= 1e-3 * b for a, b in zip(s_cnt[:3], cnt[4:7])), (s_cnt, cnt[4:])   # secondaries: a handful of rays differ in the last ulp of tmax
    t.close()


def test_wavefront_edge_cases(built_lib, orc):
    for (w, h, mpl) in ((1, 1, 4), (127, 3, 1), (129, 2, 2), (255, 1, 3), (257, 1, 3)):     # single slot, around the 256-slot tile size; depth 1 = emission only
        s = _scene("cornell", w, h)
        t = _tracer(s, w, h, mpl)
        t.DoPass(True)
        img = t.readAccumulator()
        ref, rays, q = orc.render_wavefront(s.view, w, h, n_passes=1, max_path_length=mpl)
        assert np.array_equal(_queues(t, mpl), q) and t.getRaysInLastPass() == rays
        assert (rel_l2(img["rgb"], ref["rgb"]) <= 1e-3).all() and np.array_equal(img["weight_sum"], ref["weight_sum"])
        t.close()


def test_wavefront_degenerate_scenes(built_lib, orc):
    """No lights (every path contributes zero, queues still evolve like the oracle's) and a camera that sees nothing (every primary misses:
    one iteration, then an empty queue through all remaining launches)."""
    from scene_fixtures import _mat
    V = np.array([(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)], np.float32)
    I = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    for cam, sees in ((((0, 1.5, -2.0), (0, 0, 0), (0, 1, 0), 60.0), True), (((0, 1.5, -2.0), (0, 3.0, -4.0), (0, 1, 0), 30.0), False)):
        s = ctl.Scene.from_mesh(V, I, np.array([0, 0], np.uint8), [_mat()], np.zeros((1, 3), np.float32), *cam, 40, 30)
        assert s.view.num_lights == 0
        t = _tracer(s, 40, 30, 5, rr=1)
        t.DoPass(True); t.DoPass(False)
        img = t.readAccumulator()
        ref, rays, q = orc.render_wavefront(s.view, 40, 30, n_passes=2, max_path_length=5, rr_start=1)
        assert np.all(img["rgb"] == 0) and (img["weight_sum"] == 2).all() and np.array_equal(_queues(t, 5), q) and t.getTotalRays() == rays
        assert (q[1, 0] > 0) == sees and q[:, 1].sum() == 0
        t.close()


def test_wavefront_pass_stride_and_phase(built_lib, orc):
    """Multi-GPU by pass index on one device: contexts with PassStride = 2 and PassPhase = 0 / 1 render passes {0, 2} and {1, 3} of the frame that
    a third context renders alone; every strided pass is bit-identical to the same pass of the plain sequence, the summed accumulators equal the
    4-pass frame, ray counts add up (cudatracerlib_b200.DistributedPasses does this across ranks, tests/test_multirank_cpu.py)."""
    w, h, mpl = 72, 48, 6
    s = _scene("cornell7", w, h)
    a = _tracer(s, w, h, mpl)
    per_pass = []
    prev = np.zeros((h, w, 3), np.float32)
    for p in range(4):
        a.DoPass(p == 0); acc = a.readAccumulator()["rgb"].copy(); per_pass.append(acc - prev); prev = acc
    full = a.readAccumulator().copy(); rays_full = a.getTotalRays()
    parts, rays = [], 0
    for phase in (0, 1):
        t = _tracer(s, w, h, mpl); t.setParameter("PassStride", 2); t.setParameter("PassPhase", phase)
        t.DoPass(True); first = t.readAccumulator()["rgb"].copy()
        ref, _, _ = orc.render_wavefront(s.view, w, h, n_passes=1, pass_first=phase, max_path_length=mpl)
        assert (rel_l2(first, ref["rgb"]) <= 1e-3).mean() >= 0.99
        if phase == 0:
            assert np.array_equal(first.view(np.uint32), per_pass[0].view(np.uint32))       # pass 0 is pass 0
        t.DoPass(False)
        parts.append(t.readAccumulator().copy()); rays += t.getTotalRays()
        assert t.getNumPassesDone() == 2
        t.close()
    total = parts[0]["rgb"] + parts[1]["rgb"]
    assert np.allclose(total, full["rgb"], rtol=1e-5, atol=1e-6) and np.array_equal(parts[0]["weight_sum"] + parts[1]["weight_sum"], full["weight_sum"])
    assert rays == rays_full
    a.close()


def test_cpp_adapter_wavefront(built_lib, tmp_path):
    """ctlb200::WavefrontPathTracer (include/b200_path_tracer.hpp) renders through the same entry point."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "wpt.cpp"
    src.write_text('''#include "b200_path_tracer.hpp"
#include <cstdio>
int main() {
    ctlb200::Scene scene(1, 64, 64);
    ctlb200::WavefrontPathTracer tracer;
    tracer.Resize(64, 64); tracer.InitializeScene(scene.view()); tracer.setParameter("MaxPathLength", 8);
    std::vector<ctl_pixel_data> img(64 * 64);
    tracer.DoPass(img.data(), true); tracer.DoPass(img.data(), false);
    double s = 0; for (auto& p : img) s += p.rgb[0] + p.rgb[1] + p.rgb[2];
    std::printf("%u %llu %.6f %g\\n", tracer.getNumPassesDone(), tracer.getRaysInLastPass(), s, (double)img[0].weight_sum);
    return 0;
}''')
    exe = tmp_path / "wpt"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), api.LIB_PATH, "-Wl,-rpath," + os.path.dirname(api.LIB_PATH)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == 2 and int(out[1]) > 4096 and float(out[2]) > 0 and float(out[3]) == 2.0


def test_double_ray_buffer_header(built_lib, orc, tmp_path):
    """include/b200_double_ray_buffer.cuh: a user application over ctlb200::DoubleRayBuffer<T> (tests/drb_check.cu, compiled here with nvcc):
    payload kernels with the reference's device-side method names, FinishIteration -> ctl_intersect.  The per-pixel results of its two
    iterations (first hit, bounce hit, secondary-ray hit) equal the oracle's intersectKernel restatement on the same rays, bit for bit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "drb_check"
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-fmad=false", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "drb_check.cu"), "-o", str(exe), api.LIB_PATH, "-Xlinker", "-rpath=" + os.path.dirname(api.LIB_PATH)], check=True)
    f32 = np.float32

    def normalized(v):   # the application's float32 arithmetic (no FMA): 1 / sqrt((x*x + y*y) + z*z)
        s = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
        return v * (f32(1.0) / np.sqrt(s))[:, None]

    for kind_id, kind in ((1, "cornell7"), (6, "soup")):
        lines = subprocess.run([str(exe), str(kind_id)], check=True, capture_output=True, text=True).stdout.strip().split("\n")
        sizes = [int(v) for v in lines[0].split()]
        cam = np.array([f32(v) for v in lines[1].split()[:3]], f32); light = np.array([f32(v) for v in lines[1].split()[3:]], f32)
        out = np.array([[float(v) for v in l.split()] for l in lines[2:]], f32)
        w, h = 96, 64
        s = ctl.Scene(kind, w, h)
        n = w * h
        x = (np.arange(n) % w).astype(f32); y = (np.arange(n) // w).astype(f32)
        d = normalized(np.stack([(x + f32(0.5)) / f32(w) - f32(0.5), (y + f32(0.5)) / f32(h) - f32(0.5), np.ones(n, f32)], 1).astype(f32))
        rays = np.zeros(n, api.RAY_DTYPE); rays["o"] = cam; rays["d"] = d; rays["tmin"] = s.view.ray_eps; rays["tmax"] = f32(3.402823466e+38)
        first = orc.intersect(s.view, rays)
        hit = first["tri_idx"] >= 0
        assert sizes == [n, int(hit.sum()), 0, 1]                     # queue sizes per iteration; empty at the end
        assert np.array_equal(out[:, 0].view(np.uint32), first["dist"].view(np.uint32)) and np.array_equal(out[:, 1].astype(np.int64), first["tri_idx"])
        # iteration 1: the bounce ray and the secondary ray pushed per hit
        p = (cam[None, :] + d * first["dist"][:, None]).astype(f32)[hit]
        bounce = np.zeros(len(p), api.RAY_DTYPE); bounce["o"] = p; bounce["d"] = d[hit] * np.array([-1, 1, -1], f32); bounce["tmin"] = s.view.ray_eps; bounce["tmax"] = f32(3.402823466e+38)
        sec = bounce.copy(); sec["d"] = normalized((light[None, :] - p).astype(f32))
        b, sh = orc.intersect(s.view, bounce), orc.intersect(s.view, sec)
        assert np.array_equal(out[hit, 2].view(np.uint32), b["dist"].view(np.uint32)) and np.array_equal(out[hit, 3].view(np.uint32), sh["dist"].view(np.uint32))
        assert (out[~hit, 2:] == 0).all()


def test_wavefront_frame_equals_sequential_passes(built_lib):
    """ctl_wavefront_frame: the passes of a frame on several streams with their own queues -- same passes, same paths: weights and ray totals equal, radiance
    equal up to the order of the float atomics; also with host-generated tables and with one lane (= the plain loop)."""
    w, h, spp = 160, 120, 6
    s = ctl.Scene("soup", w, h)
    t = ctl.WavefrontPathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    r0 = t.getTotalRays()
    for p in range(spp): t.DoPass(p == 0)
    t.synchronize(); ref = t.readAccumulator(); ref_rays = t.getTotalRays() - r0
    for lanes, dev in ((4, 1), (8, 1), (3, 0), (1, 1)):
        t.setParameter("OverlapLanes", lanes); t.setParameter("DeviceSampleTables", dev)
        for _ in range(2):
            r0 = t.getTotalRays(); t.DoFrame(spp); t.synchronize()
            img = t.readAccumulator()
            assert t.getTotalRays() - r0 == ref_rays and t.getNumPassesDone() == spp
            assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and np.allclose(img["rgb"], ref["rgb"], rtol=2e-5, atol=1e-6)
    t.DoPass(False); t.synchronize(); assert t.getNumPassesDone() == spp + 1   # and single passes continue the trace
    t.close()
