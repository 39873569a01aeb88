"""GPU tests of NonLocalMeansFilter (csrc/nlm_filter.cuh) through the C ABI: ctl_apply_image_pipeline with filter_type 5 against the oracle,
which is pinned bit-identically to the reference's own kernels (tests/test_golden_cpu.py).  The kernels were written after this round's GPU budget was
spent: their source is verified on the host (tests/test_nlm_kernels_cpu.py), this file is their first run on a device -- it sorts last so that a problem
here cannot hide the rest of the suite.

Tolerances: weights are exp() of sums in a fixed order -- equal to the oracle's within 1e-6 except for the rare weight that falls on the other side of the
0.05 cut-off (<= 1e-4 of them); final bytes <= 2 LSB on <= 1 % of the channels (2 % behind the tone mapper), as for the other filters."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api, ImagePipeline

pytestmark = pytest.mark.gpu


def _close(a, b, max_lsb, frac):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max() <= max_lsb and (d > 0).mean() <= frac


def _weights_close(dev, ref):
    d = np.abs(dev - ref)
    flipped = (dev == 0) != (ref == 0)
    return flipped.mean() <= 1e-4 and d[~flipped].max() <= 1e-6


def _nlm(k, s2, period=25, tonemap=0):
    return ImagePipeline(5, float(period), 0.0, k, s2, tonemap)


@pytest.mark.parametrize("cls", ["PathTracer", "WavefrontPathTracer"])
def test_non_local_means_vs_oracle(built_lib, orc, cls):
    w, h = 72, 52                                                  # not multiples of the 16x16 blocks
    s = ctl.Scene("cornell", w, h)
    t = getattr(ctl, cls)(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 6); t.setParameter("PixelVarianceBuffer", 1)
    for p in range(6):
        t.DoPass(p == 0)
    acc, var = t.readAccumulator(), t.readVarianceBuffer().reshape(-1)
    for k, s2, tm in ((0.45, 1.0, 0), (1.0, 5.0, 0), (0.45, 1.0, 1), (0.45, 0.005, 0)):
        P = _nlm(k, s2, period=1, tonemap=tm)
        got = t.applyImagePipeline(P)
        stage, wts = orc.nlm_filter(acc, var, k, s2)
        ref, _ = orc.pipeline_from_stage2(stage, P)
        assert _weights_close(t.readNlmWeights(), wts), (k, s2)
        assert _close(got, ref, 2, 0.02 if tm else 0.01), (k, s2, tm, np.abs(got.astype(int) - ref.astype(int)).max(), (got != ref).mean())
        if (k, s2) == (0.45, 1.0) and not tm:
            assert (got != t.applyImagePipeline(ImagePipeline(-1))).any(axis=2).mean() > 0.5      # it does filter
        if s2 == 0.005:
            assert (wts > 0).mean() < 0.02                                                        # reference defaults at 6 spp: close to the identity
    t.close()


def test_non_local_means_weight_update_schedule(built_lib, orc):
    """NonLocalMeansFilter.cu:207-224: weights are recomputed when the pass count did not advance by exactly one since the last application, or is a
    multiple of UpdateWeightPeriodicity, or after a resize; otherwise the stored weights filter the new frame."""
    w, h = 48, 36
    s = ctl.Scene("cornell7", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 5); t.setParameter("PixelVarianceBuffer", 1)
    P = _nlm(0.45, 1.0, period=25)
    for p in range(3):
        t.DoPass(p == 0)
    t.applyImagePipeline(P)                                        # first use: computed on the 3-pass frame
    w3 = t.readNlmWeights()
    assert _weights_close(w3, orc.nlm_filter(t.readAccumulator(), t.readVarianceBuffer().reshape(-1), 0.45, 1.0)[1])
    t.DoPass(False)                                                # 4 passes: 3 + 1 == 4 and 4 % 25 != 0 -> stale weights
    got = t.applyImagePipeline(P)
    assert np.array_equal(t.readNlmWeights(), w3)
    acc4, var4 = t.readAccumulator(), t.readVarianceBuffer().reshape(-1)
    stage, _ = orc.nlm_filter(acc4, var4, 0.45, 1.0, weights=w3)
    assert _close(got, orc.pipeline_from_stage2(stage, P)[0], 2, 0.01)
    got_again = t.applyImagePipeline(P)                            # same pass count again: 4 + 1 != 4 -> recomputed
    w4 = t.readNlmWeights()
    assert not np.array_equal(w4, w3) and _weights_close(w4, orc.nlm_filter(acc4, var4, 0.45, 1.0)[1])
    t.DoPass(False)                                                # 5 passes with periodicity 5: multiple -> recomputed
    t.applyImagePipeline(_nlm(0.45, 1.0, period=5))
    assert _weights_close(t.readNlmWeights(), orc.nlm_filter(t.readAccumulator(), t.readVarianceBuffer().reshape(-1), 0.45, 1.0)[1])
    t.close()


def test_non_local_means_errors_and_full_hd(built_lib):
    t = ctl.PathTracer(32, 32)
    with pytest.raises(RuntimeError, match="PixelVarianceBuffer"):
        t.applyImagePipeline(_nlm(0.45, 1.0))                      # no variance buffer
    with pytest.raises(RuntimeError):
        t.readNlmWeights()
    t.close()
    w, h = 1920, 1080                                              # full size: 1.4 GB of weights; a flat image stays flat, a step edge stays an edge
    s = ctl.Scene("cornell", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 4); t.setParameter("PixelVarianceBuffer", 1)
    for p in range(3):
        t.DoPass(p == 0)
    with pytest.raises(RuntimeError, match="UpdateWeightPeriodicity"):
        t.applyImagePipeline(ImagePipeline(5, 0.0, 0.0, 0.45, 1.0))
    plain = t.applyImagePipeline(ImagePipeline(-1)).astype(np.float64)
    den = t.applyImagePipeline(_nlm(0.45, 1.0)).astype(np.float64)
    assert den.shape == (h, w, 4) and (den[..., 3] == 255).all()
    # denoising: less pixel-to-pixel variation on the lit walls, same mean brightness
    assert abs(den[..., :3].mean() - plain[..., :3].mean()) < 2.0
    assert np.abs(np.diff(den[..., :3], axis=1)).mean() < 0.8 * np.abs(np.diff(plain[..., :3], axis=1)).mean()
    t.close()
