"""Parity where the benchmark runs: 1920x1080, 8 passes fused into one wavefront through ctl_render_passes_tiled and through the frames-in-flight
pipeline (ctl_submit_frame_tiled / ctl_acquire_frame) -- the exact calls bench.py makes -- compared with the oracle on fixed windows of the frame: the bottom-right corner (sampler indices up to 2 073 599, far beyond the 65 536 the small test
images reach: (idx / 4096) % 4096 > 15, Kernel/Sampler_device.h:91-107) and a window straddling several 64x64 tile boundaries.

Tolerances: same seed, same passes; per-pixel relative L2 <= 1e-3 on >= 99 % of the pixels, image relative RMSE <= 3e-3 at 8 spp, mean
within 0.1 %, weights exact.  The 1 M-triangle foliage scene amplifies 1-ulp differences chaotically (a ray grazing one of 100 000 thin triangles
flips its hit): there the floor is MEASURED in the test -- the agreement of the oracle's own two arithmetic builds (explicit FMA = nvcc, no FMA =
the reference's g++ host build, bit-identical to oracle/_ref) on the same window -- and the CUDA path must agree with the oracle at least that well.
Ray counts are compared in the reference's definition (StopZeroThroughput=0; tests/test_ray_count_parity_cpu.py pins it to oracle/_ref) to <= 1e-3."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl

pytestmark = pytest.mark.gpu

W, H, SPP, DEPTH = 1920, 1080, 8, 8
WINDOWS = [(W - 256, H - 256, W, H), (900, 500, 1156, 756)]   # corner: idx <= 2 073 599; middle: tile edges at x = 960, 1024, 1088, 1152 and y = 512 .. 704


def _rel(a, b):
    return np.linalg.norm(a - b, axis=-1) / (np.linalg.norm(b, axis=-1) + 1e-3)


@pytest.mark.parametrize("kind", ["c2", "c3", "c4"])
def test_full_resolution_windows_match_oracle(built_lib, orc, kind):
    s = ctl.Scene(kind, W, H)
    t = ctl.PathTracer(W, H); t.InitializeScene(s); t.setParameter("MaxPathLength", DEPTH)
    t.DoPasses(SPP, new_trace=True); t.synchronize()             # == bench.py's frame at N = 1: ctl_render_passes_tiled(8 passes, 64x64 tiles, part 0 of 1)
    img = t.readAccumulator(); rays_stop = t.getRaysInLastPass()
    assert t.getNumPassesDone() == SPP
    landed = float(img["weight_sum"].astype(np.float64).sum())
    assert SPP * W * H - 64 <= landed <= SPP * W * H             # every path of every pass landed (a jittered sample may round into the next pixel; past the image edge it is dropped, Image.cu:22-44)
    # the same frame as bench.py renders it since the frames-in-flight pipeline: ctl_submit_frame_tiled / ctl_acquire_frame, 3 frames in flight, whole image
    t.setParameter("FramesInFlight", 3)
    r0 = t.getTotalRays()
    for _ in range(3):
        t.submitFrame(SPP, SPP)
    for _ in range(3):
        t.acquireFrame(); piped = t.readAccumulator()
        assert np.array_equal(piped["weight_sum"], img["weight_sum"])
        assert np.allclose(piped["rgb"], img["rgb"], rtol=2e-5, atol=1e-6)    # same paths; only the order of the float atomics into a pixel differs
    assert t.getTotalRays() - r0 == 3 * rays_stop
    # the same frame with the reference's ray definition
    t.setParameter("StopZeroThroughput", 0)
    t.DoPasses(SPP, new_trace=True); t.synchronize()
    img_ns = t.readAccumulator(); rays_ns = t.getRaysInLastPass()
    assert np.array_equal(img_ns["weight_sum"], img["weight_sum"])
    assert np.allclose(img_ns["rgb"], img["rgb"], rtol=1e-5, atol=1e-7)   # zero-throughput paths add nothing
    t.setParameter("StopZeroThroughput", 1)
    for win in WINDOWS:
        x0, y0, x1, y1 = win
        ref, ref_rays = orc.render(s.view, W, H, n_passes=SPP, max_path_length=DEPTH, window=win)
        inner = (slice(y0 + 1, y1 - 1), slice(x0 + 1, x1 - 1))   # a jittered sample can round into the next pixel: the window's rim may hold a neighbour's sample
        a, b = img["rgb"][inner], ref["rgb"][inner]
        assert np.array_equal(img["weight_sum"][inner], ref["weight_sum"][inner])
        r = _rel(a, b)
        frac = float((r <= 1e-3).mean()); rmse = float(np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean())); dmean = float(abs(a.mean() - b.mean()) / b.mean())
        if kind == "c4":
            with orc.host_arithmetic():
                ref2, _ = orc.render(s.view, W, H, n_passes=SPP, max_path_length=DEPTH, window=win)
            r2 = _rel(ref2["rgb"][inner], b)
            floor = float((r2 <= 1e-3).mean())
            print(f"{kind} {win}: frac {frac:.4f} (oracle FMA vs no-FMA floor {floor:.4f}) rmse {rmse:.2e} mean {dmean:.2e}")
            assert frac >= floor, (frac, floor)                   # at least as close to the oracle as the oracle's two arithmetic builds are to each other
            assert np.median(r) <= 1e-5 and dmean <= 2e-3
        else:
            print(f"{kind} {win}: frac {frac:.4f} rmse {rmse:.2e} mean {dmean:.2e}")
            assert frac >= 0.99, frac
            assert rmse <= 3e-3 or (kind == "c3" and rmse <= 3e-2), rmse   # c3: a handful of glass / rough-conductor paths flip a discrete decision (libdevice vs libm) and carry fireflies
            assert dmean <= 1e-3, dmean
    # ray counts in the reference's definition (StopZeroThroughput=0 on both sides): one pass of the corner window through ctl_render_pass
    win = WINDOWS[0]
    orc.set_stop_zero_throughput(0)
    try:
        t.setParameter("StopZeroThroughput", 0)
        t.DoPass(True, window=win); t.synchronize()
        g = t.getRaysInLastPass()
        _, o1 = orc.render(s.view, W, H, n_passes=1, max_path_length=DEPTH, window=win)
        assert abs(g - o1) <= 1e-3 * o1, (g, o1)
        assert rays_stop <= rays_ns <= 1.2 * rays_stop                       # the switch only adds the zero-weight tails
    finally:
        orc.set_stop_zero_throughput(1)
        t.setParameter("StopZeroThroughput", 1)
    t.close()


def test_full_resolution_window_vs_reference_live(built_lib):
    """The same corner window against the reference's OWN PathTrace (oracle/_ref, its g++ host arithmetic): c2, where the only differences are FMA
    contraction (<= 1 ulp in t) and libm vs libdevice."""
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref is built only where /root/reference is mounted (the prebuilt library travels with the snapshot)")
    s = ctl.Scene("c2", W, H)
    t = ctl.PathTracer(W, H); t.InitializeScene(s); t.setParameter("MaxPathLength", DEPTH)
    t.DoPasses(SPP, new_trace=True); t.synchronize()
    img = t.readAccumulator()
    win = WINDOWS[0]; x0, y0, x1, y1 = win
    ref, _ = rb.render(s.view, W, H, n_passes=SPP, max_path_length=DEPTH, window=win)
    inner = (slice(y0 + 1, y1 - 1), slice(x0 + 1, x1 - 1))
    a, b = img["rgb"][inner], ref["rgb"][inner]
    assert np.array_equal(img["weight_sum"][inner], ref["weight_sum"][inner])
    r = _rel(a, b)
    assert (r <= 1e-3).mean() >= 0.985, (r <= 1e-3).mean()      # measured floor between the oracle's two arithmetic builds on this window: 0.992
    assert abs(a.mean() - b.mean()) <= 1e-3 * b.mean()
    t.close()


def test_full_resolution_config5_frame_on_lanes(built_lib, orc):
    """configs[4] exactly as bench.py renders it -- 1920x1080, 64 spp, 32 bounces, ctl_render_frame_tiled with 16 passes per wavefront on four lanes --
    against the oracle on a 128x128 window.  Each pixel sums 64 paths of a glass / rough-conductor scene whose deep specular chains amplify 1-ulp
    differences chaotically, so almost every pixel holds a diverged path: the floor is MEASURED here as the agreement of the oracle's own two arithmetic
    builds (explicit FMA = what the device computes, no FMA = the reference's host build), and the CUDA path must be about as close to the oracle as
    those are to each other; weights exact, ray count within 5e-3, and the frame equals the one-wavefront-at-a-time frame."""
    spp, depth, batch = 64, 32, 16
    s = ctl.Scene("c5", W, H)
    t = ctl.PathTracer(W, H); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
    assert t.getParameter("OverlapWavefronts") == 1 and t.getParameter("OverlapLanes") == 4
    r0 = t.getTotalRays(); t.DoFrame(spp, batch); t.synchronize(); img = t.readAccumulator(); rays = t.getTotalRays() - r0
    assert t.getNumPassesDone() == spp
    t.setParameter("OverlapWavefronts", 0)
    r0 = t.getTotalRays(); t.DoFrame(spp, batch); t.synchronize(); img1 = t.readAccumulator(); rays1 = t.getTotalRays() - r0
    assert rays == rays1 and np.array_equal(img["weight_sum"], img1["weight_sum"]) and np.allclose(img["rgb"], img1["rgb"], rtol=5e-5, atol=1e-6)
    win = (1000, 600, 1128, 728); x0, y0, x1, y1 = win
    inner = (slice(y0 + 1, y1 - 1), slice(x0 + 1, x1 - 1))
    ref, ref_rays = orc.render(s.view, W, H, n_passes=spp, max_path_length=depth, window=win)
    with orc.host_arithmetic():
        ref2, _ = orc.render(s.view, W, H, n_passes=spp, max_path_length=depth, window=win)
    assert np.array_equal(img["weight_sum"][inner], ref["weight_sum"][inner])
    a, b, b2 = img["rgb"][inner], ref["rgb"][inner], ref2["rgb"][inner]
    r, r2 = _rel(a, b), _rel(b2, b)
    print(f"c5 {win}: median rel {np.median(r):.2e} (oracle FMA vs no-FMA: {np.median(r2):.2e}); within 1e-2: {(r <= 1e-2).mean():.4f} (floor {(r2 <= 1e-2).mean():.4f}); mean {abs(a.mean() - b.mean()) / b.mean():.2e}")
    assert np.median(r) <= 2.0 * np.median(r2) + 1e-5
    assert (r <= 1e-2).mean() >= (r2 <= 1e-2).mean() - 0.03
    assert abs(a.mean() - b.mean()) <= 5e-3 * b.mean()
    # ray count of the window: one pass through ctl_render_pass against the oracle's
    t.DoPass(True, window=win); t.synchronize()
    _, o1 = orc.render(s.view, W, H, n_passes=1, max_path_length=depth, window=win)
    assert abs(t.getRaysInLastPass() - o1) <= 5e-3 * o1
    t.close()
