"""ctl_render_frame_tiled: a frame whose wavefronts alternate between two streams ("OverlapWavefronts") traces the same paths as the sequential
wavefronts of ctl_render_passes_tiled -- same weights, same ray count, radiance equal up to the order of the float atomics into PixelData -- on the
whole image and on the tiles of one part of eight (what a rank of an 8-GPU frame renders), and agrees with the oracle like the plain path does."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl

pytestmark = pytest.mark.gpu


def _frame(t, spp, batch, overlap, part=0, n_parts=1):
    t.setParameter("OverlapWavefronts", 2 * overlap)
    r0 = t.getTotalRays()
    t.DoFrame(spp, batch, part=part, n_parts=n_parts); t.synchronize()
    return t.readAccumulator(), t.getTotalRays() - r0


@pytest.mark.parametrize("kind,w,h,spp,batch,parts", [("cornell7", 200, 136, 8, 8, 1), ("soup", 256, 192, 8, 2, 1), ("c3", 512, 288, 8, 8, 8), ("c3", 512, 288, 6, 6, 1), ("cornell7", 64, 64, 3, 3, 1)])
def test_overlapped_frame_equals_sequential_frame(built_lib, kind, w, h, spp, batch, parts):
    s = ctl.Scene(kind, w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 8)
    assert t.getParameter("OverlapWavefronts") == 1           # the default: lanes when the frame has several wavefronts anyway; 2 also cuts the batches
    for part in range(min(parts, 2)):
        a, rays_a = _frame(t, spp, batch, 1, part, parts)
        assert t.getNumPassesDone() == spp
        b, rays_b = _frame(t, spp, batch, 0, part, parts)
        # and the call bench.py's stage timers use: the same wavefronts through ctl_render_passes_tiled
        t.StartNewTrace()
        for p in range(0, spp, batch):
            t.DoPasses(batch, new_trace=(p == 0), part=part, n_parts=parts)
        t.synchronize(); c = t.readAccumulator()
        assert rays_a == rays_b and rays_a > 0
        assert np.array_equal(a["weight_sum"], b["weight_sum"]) and np.array_equal(b["weight_sum"], c["weight_sum"])
        assert np.allclose(a["rgb"], b["rgb"], rtol=2e-5, atol=1e-6) and np.allclose(b["rgb"], c["rgb"], rtol=2e-5, atol=1e-6)
    t.close()


def test_overlapped_frame_matches_oracle(built_lib, orc):
    w, h, spp = 160, 120, 4
    s = ctl.Scene("cornell7", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 8)
    img, rays = _frame(t, spp, 4, 1)
    ref, ref_rays = orc.render(s.view, w, h, n_passes=spp, max_path_length=8)
    rel = np.linalg.norm(img["rgb"] - ref["rgb"], axis=2) / (np.linalg.norm(ref["rgb"], axis=2) + 1e-3)
    assert np.array_equal(img["weight_sum"], ref["weight_sum"])
    assert (rel <= 1e-3).mean() >= 0.99
    assert abs(rays - ref_rays) <= 5e-3 * ref_rays              # StopZeroThroughput=1 (default) ends zero-weight paths early
    t.close()


def test_frame_with_host_generated_tables_and_lanes(built_lib):
    """DeviceSampleTables = 0 (tables from the host XORWOW twin, one set per pass of the frame, produced wavefront by wavefront on the table stream) renders the
    frame the device generator renders; 1 .. 8 lanes render the same frame."""
    w, h, spp, batch = 192, 128, 16, 2
    s = ctl.Scene("soup", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    ref = None
    for lanes, dev in [(1, 1), (2, 1), (4, 0), (8, 1), (8, 0), (3, 0)]:
        t.setParameter("OverlapLanes", lanes); t.setParameter("DeviceSampleTables", dev)
        for _ in range(2):                                          # twice: the second frame restarts the sample stream and reuses the pinned sets
            r0 = t.getTotalRays(); t.DoFrame(spp, batch); t.synchronize(); rays = t.getTotalRays() - r0
            img = t.readAccumulator()
            if ref is None: ref = (img, rays)
            assert rays == ref[1] and np.array_equal(img["weight_sum"], ref[0]["weight_sum"])
            assert np.allclose(img["rgb"], ref[0]["rgb"], rtol=2e-5, atol=1e-6)
    t.close()


def test_concurrent_class_shade_launches(built_lib):
    """ShadeConcurrent = 1: the per-class shade launches of a bounce run on their own streams (disjoint queue segments, atomic appends) -- same paths."""
    w, h = 320, 180
    s = ctl.Scene("c3", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 8)
    out = {}
    for conc in (0, 1, 0):
        t.setParameter("ShadeConcurrent", conc)
        r0 = t.getTotalRays(); t.DoPasses(4, new_trace=True); t.synchronize()
        out[conc] = (t.readAccumulator(), t.getTotalRays() - r0)
    assert out[0][1] == out[1][1] and np.array_equal(out[0][0]["weight_sum"], out[1][0]["weight_sum"])
    assert np.allclose(out[0][0]["rgb"], out[1][0]["rgb"], rtol=2e-5, atol=1e-6)
    t.setParameter("ShadeConcurrent", 1); t.setParameter("OverlapLanes", 3)
    t.DoFrame(8, 2); t.synchronize(); a = t.readAccumulator()
    t.setParameter("ShadeConcurrent", 0); t.DoFrame(8, 2); t.synchronize(); b = t.readAccumulator()
    assert np.array_equal(a["weight_sum"], b["weight_sum"]) and np.allclose(a["rgb"], b["rgb"], rtol=2e-5, atol=1e-6)
    t.close()


def test_frame_argument_errors(built_lib):
    t = ctl.PathTracer(64, 64)
    with pytest.raises(RuntimeError, match="multiple of batch"):
        t.DoFrame(8, 3)
    with pytest.raises(RuntimeError, match="no scene"):
        t.DoFrame(8, 4)
    t.close()
