"""GPU tests of the full image pipeline and the PixelVarianceBuffer through the C ABI.

Tolerances: the reconstruction filters and the tone mapper use expf / sinf / logf / powf (libdevice vs libm), and every stage quantises
to 8 bits, so a last-ulp difference can flip a byte: <= 2 LSB on <= 1 % of the channels (filters), <= 2 LSB on <= 2 % (tone mapping, two
quantisations); the luminance min / max / average are sums of products in a fixed order -> bit-exact; log-average within 1e-6.
PixelVarianceBuffer: plain arithmetic -> bit-exact."""
import os

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api, ImagePipeline

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))


def _tracer_with_accum(acc):
    """A tracer whose accumulator holds `acc` (PixelData[h, w]) -- through ctl_set_accum_device_ptr with a torch tensor."""
    import torch
    h, w = acc.shape
    t = ctl.PathTracer(w, h)
    buf = torch.from_numpy(np.ascontiguousarray(acc).view(np.float32).reshape(-1).copy()).cuda()
    t.setAccumDevicePtr(buf.data_ptr())
    return t, buf


def _close(a, b, max_lsb, frac):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max() <= max_lsb and (d > 0).mean() <= frac


def test_full_pipeline_vs_reference_goldens(built_lib):
    acc = np.ascontiguousarray(GOLD["pipeline_accum_cornell_80x64_4spp"]).view(api.PIXEL_DTYPE).reshape(64, 80)
    t, buf = _tracer_with_accum(acc)
    for k, (ft, xw, yw, p0, p1, tm, key, burn) in enumerate(GOLD["pipeline_full_cases"]):
        P = ImagePipeline(int(ft), float(xw), float(yw), float(p0), float(p1), int(tm), float(key), float(burn))
        rgba, lum = t.applyImagePipeline(P, lum_info=True)
        ref = GOLD[f"pipeline_full_{k}"]
        assert _close(rgba, ref, 2, 0.02 if tm else 0.01), (k, np.abs(rgba.astype(int) - ref.astype(int)).max(), (rgba != ref).mean())
        assert (rgba[..., 3] == 255).all()
        if tm:
            rl = GOLD[f"pipeline_full_{k}_lum"]
            if ft < 0:   # no filter in front: the RGBE stage is pure arithmetic -> min / max / average bit-exact
                assert np.array_equal(lum[:3].view(np.uint32), rl[:3].view(np.uint32)), (lum, rl)
            assert np.allclose(lum, rl, rtol=2e-6), (lum, rl)
    t.close()


def test_full_pipeline_vs_oracle_on_rendered_frames(built_lib, orc):
    s = ctl.Scene("c3", 160, 90)
    t = ctl.PathTracer(160, 90); t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    t.DoPasses(4, new_trace=True)
    acc = t.readAccumulator()
    for P in (ImagePipeline(3, 2, 2, 1 / 3, 1 / 3), ImagePipeline(4, 3, 2, 2.0), ImagePipeline(-1, tonemap=1), ImagePipeline(2, 1.5, 1.5, tonemap=1, key=0.3, burn=0.2)):
        got, lum = t.applyImagePipeline(P, lum_info=True)
        ref, rl = orc.apply_image_pipeline(acc, P)
        assert _close(got, ref, 2, 0.02 if P.tonemap else 0.01), (P.filter_type, P.tonemap, (got != ref).mean())
        if P.tonemap:
            assert np.allclose(lum, rl, rtol=2e-6)
    # the old entry points are the same code path
    assert np.array_equal(t.resolveSRGB8(), t.applyImagePipeline(ImagePipeline(-1)))
    assert np.array_equal(t.resolveFilteredSRGB8("gaussian", 2, 2, 2.0), t.applyImagePipeline(ImagePipeline(1, 2, 2, 2.0)))
    t.close()


def test_pipeline_full_hd_properties(built_lib):
    """1920x1080: odd sizes vs 16x16 luminance blocks, constant image -> constant output, tone mapping is monotone in luminance."""
    import torch
    w, h = 1920, 1080
    t = ctl.PathTracer(w, h)
    acc = np.zeros((h, w), api.PIXEL_DTYPE); acc["weight_sum"] = 2.0
    ramp = np.linspace(0.01, 8.0, w, dtype=np.float32)
    acc["rgb"][:] = (2.0 * ramp)[None, :, None]
    buf = torch.from_numpy(acc.view(np.float32).reshape(-1).copy()).cuda(); t.setAccumDevicePtr(buf.data_ptr())
    out, lum = t.applyImagePipeline(ImagePipeline(-1, tonemap=1), lum_info=True)
    assert np.isclose(lum[0], 0.01, rtol=0.02) and np.isclose(lum[1], 8.0, rtol=0.02) and np.isclose(lum[2], ramp.mean(), rtol=0.01)
    assert (out[0] == out[-1]).all() and (np.diff(out[0, :, 0].astype(int)) >= 0).all() and out[0, -1, 0] >= 250 and out[0, 0, 0] < 120
    box = t.applyImagePipeline(ImagePipeline(0, 0.5, 0.5)); direct = t.applyImagePipeline(ImagePipeline(-1))
    assert np.abs(box.astype(int) - direct.astype(int)).max() <= 3   # a 1-pixel box filter only adds the RGBE quantisation
    with pytest.raises(RuntimeError):
        t.applyImagePipeline(ImagePipeline(7))
    with pytest.raises(RuntimeError):
        t.applyImagePipeline(ImagePipeline(0, 0.0, 1.0))
    t.close()


def test_pixel_variance_buffer(built_lib, orc):
    for cls in (ctl.PathTracer, ctl.WavefrontPathTracer):
        s = ctl.Scene("cornell7", 48, 40)
        t = cls(48, 40); t.InitializeScene(s); t.setParameter("MaxPathLength", 6); t.setParameter("PixelVarianceBuffer", 1)
        var = np.zeros(48 * 40, ctl.VARIANCE_DTYPE)
        for p in range(5):
            t.DoPass(p == 0)
            orc.variance_add_pass(var, t.readAccumulator())     # the restatement applied to the same accumulator snapshots
        got = t.readVarianceBuffer().reshape(-1)
        assert got.tobytes() == var.tobytes()
        assert (got["iterations_done"] == 5).all() and (got["sum_x2"] >= 0).all()
        t.DoPass(True)                                          # new trace clears the buffer (Tracer.h:222-226)
        assert (t.readVarianceBuffer()["iterations_done"] == 1).all()
        if cls is ctl.PathTracer:
            with pytest.raises(RuntimeError):
                t.DoPasses(4, new_trace=True)                   # fused passes have no per-pass image states
        t.close()
    t = ctl.PathTracer(16, 16)
    with pytest.raises(RuntimeError):
        t.readVarianceBuffer()
    t.close()
