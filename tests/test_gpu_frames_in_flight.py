"""Frames in flight (ctl_submit_frame_tiled / ctl_acquire_frame, "FramesInFlight"): a sequence of frames rendered as a pipeline -- every frame on a lane of
its own (stream, wavefront buffers, accumulator, sample-table sets) -- must hand back, in order, exactly the frames ctl_render_frame_tiled renders: same
weights, same ray counts, radiance equal up to the order of the float atomics into PixelData; with device- and host-generated sample tables, on the whole
image and on one part of several, and through the single-process communicator on >= 2 GPUs (reduce on the communication stream)."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl

pytestmark = pytest.mark.gpu


def _plain(t, spp, batch, part, parts):
    r0 = t.getTotalRays()
    t.DoFrame(spp, batch, part=part, n_parts=parts); t.synchronize()
    return t.readAccumulator(), t.getTotalRays() - r0


@pytest.mark.parametrize("kind,w,h,spp,batch,parts,fif,dev_tables", [
    ("cornell7", 200, 136, 8, 8, 1, 2, 1), ("soup", 256, 192, 8, 2, 1, 3, 1), ("c3", 512, 288, 8, 8, 8, 2, 1), ("c3", 320, 180, 4, 4, 2, 4, 0), ("cornell7", 64, 64, 3, 3, 1, 7, 0),
    ("cornell7", 96, 64, 2, 1, 1, 1, 1)])
def test_pipelined_frames_equal_plain_frames(built_lib, kind, w, h, spp, batch, parts, fif, dev_tables):
    s = ctl.Scene(kind, w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 8); t.setParameter("DeviceSampleTables", dev_tables)
    part = parts - 1
    ref, ref_rays = _plain(t, spp, batch, part, parts)
    assert t.getParameter("FramesInFlight") == 2                     # the default
    t.setParameter("FramesInFlight", fif)
    n_frames = 2 * fif + 1
    got = []
    r0 = t.getTotalRays()
    for i in range(n_frames):
        t.submitFrame(spp, batch, part=part, n_parts=parts)
        assert t.framesInFlight() == min(i + 1, fif)
        if i >= fif - 1:
            t.acquireFrame(); got.append(t.readAccumulator())        # (readAccumulator synchronises the context's stream: the acquired frame is complete, the others keep running)
            assert t.getNumPassesDone() == spp
    while t.framesInFlight():
        t.acquireFrame(); got.append(t.readAccumulator())
    t.synchronize()
    assert len(got) == n_frames
    assert t.getTotalRays() - r0 == n_frames * ref_rays and ref_rays > 0
    for img in got:
        assert np.array_equal(img["weight_sum"], ref["weight_sum"])
        assert np.allclose(img["rgb"], ref["rgb"], rtol=2e-5, atol=1e-6)
    assert t.getLastTimeSpentRenderingSec() > 0                     # ctl_stats: the acquired frame's begin -> done on its lane
    # and the plain call still renders the same frame afterwards (nothing in flight, the context's own sample stream re-synchronised)
    again, again_rays = _plain(t, spp, batch, part, parts)
    assert again_rays == ref_rays and np.array_equal(again["weight_sum"], ref["weight_sum"]) and np.allclose(again["rgb"], ref["rgb"], rtol=2e-5, atol=1e-6)
    t.close()


def test_pipeline_matches_oracle(built_lib, orc):
    w, h, spp = 160, 120, 4
    s = ctl.Scene("cornell7", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 8)
    ref, ref_rays = orc.render(s.view, w, h, n_passes=spp, max_path_length=8)
    t.setParameter("FramesInFlight", 3)
    for _ in range(3):
        t.submitFrame(spp, 4)
    for k in range(3):
        t.acquireFrame(); img = t.readAccumulator()
        rel = np.linalg.norm(img["rgb"] - ref["rgb"], axis=2) / (np.linalg.norm(ref["rgb"], axis=2) + 1e-3)
        assert np.array_equal(img["weight_sum"], ref["weight_sum"])
        assert (rel <= 1e-3).mean() >= 0.99
    t.close()


def test_pipeline_errors(built_lib):
    w, h = 64, 64
    t = ctl.PathTracer(w, h)
    with pytest.raises(RuntimeError, match="no scene"):
        t.submitFrame(4, 4)
    s = ctl.Scene("cornell7", w, h); t.InitializeScene(s)
    with pytest.raises(RuntimeError, match="no frame in flight"):
        t.acquireFrame()
    with pytest.raises(RuntimeError, match="multiple of batch"):
        t.submitFrame(8, 3)
    with pytest.raises(RuntimeError, match="out of range"):
        t.setParameter("FramesInFlight", 8)
    t.setParameter("FramesInFlight", 2)
    t.submitFrame(2, 2); t.submitFrame(2, 2)
    with pytest.raises(RuntimeError, match="outstanding"):
        t.submitFrame(2, 2)
    for call in (lambda: t.DoPass(), lambda: t.DoFrame(2, 2), lambda: t.DoPasses(2, new_trace=True), lambda: t.setParameter("FramesInFlight", 3), lambda: t.Resize(32, 32), lambda: t.InitializeScene(s)):
        with pytest.raises(RuntimeError, match="in flight"):
            call()
    t.acquireFrame(); t.acquireFrame(); t.synchronize()
    t.setParameter("StageTimers", 1)
    with pytest.raises(RuntimeError, match="StageTimers"):
        t.submitFrame(2, 2)
    t.setParameter("StageTimers", 0)
    t.Resize(48, 32)                                                # the slots are released with the image
    s2 = ctl.Scene("cornell7", 48, 32); t.InitializeScene(s2)
    t.submitFrame(2, 2); t.acquireFrame(); a = t.readAccumulator()
    t.DoFrame(2, 2); t.synchronize(); b = t.readAccumulator()
    assert a["rgb"].shape == (32, 48, 3) and np.array_equal(a["weight_sum"], b["weight_sum"]) and np.allclose(a["rgb"], b["rgb"], rtol=2e-5, atol=1e-6)
    t.close()


def test_pipeline_single_process_multi_gpu(built_lib):
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices")
    w, h, spp = 256, 144, 4
    s = ctl.Scene("soup", w, h)
    one = ctl.PathTracer(w, h, device=0); one.InitializeScene(s); one.setParameter("MaxPathLength", 6)
    one.DoFrame(spp, 4); one.synchronize(); ref = one.readAccumulator()
    ts = [ctl.PathTracer(w, h, device=d) for d in range(n)]
    for t in ts:
        t.InitializeScene(s); t.setParameter("MaxPathLength", 6); t.setParameter("FramesInFlight", 3)
    ctl.PathTracer.commInitAll(ts)
    for i in range(5):
        ctl.PathTracer.commSubmitFrameAll(ts, spp, 4, 32, 0)
        if i >= 2:
            for t in ts:
                t.acquireFrame()
            img = ts[0].readAccumulator()
            assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and np.allclose(img["rgb"], ref["rgb"], rtol=1e-5, atol=1e-7)
    while ts[0].framesInFlight():
        for t in ts:
            t.acquireFrame()
        img = ts[0].readAccumulator()
        assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and np.allclose(img["rgb"], ref["rgb"], rtol=1e-5, atol=1e-7)
    for t in ts + [one]:
        t.synchronize(); t.close()
