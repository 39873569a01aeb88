"""KEY_Regularization on the device: ctl_set_param_i("Regularization", 1) renders PathTraceRegularization<DIRECT> (Integrators/PathTracer.cu:115-170) -- a
different estimator from PathTrace (all lights per vertex, no emitter MIS, roulette after specular bounces, one ray traced past the last vertex).
Compared with the oracle's restatement, which is bit-identical to the reference's own code (tests/test_regularization_cpu.py), and with the goldens
minted from that code: weights exact, >= 99 % of the pixels within 1e-3 (deep glass paths: >= 97 %), ray counts in the reference's definition
(StopZeroThroughput = 0) within 1e-3 of the reference's."""
import os
import sys

import numpy as np
import pytest

import cudatracerlib_b200 as ctl

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_regularization_golden import CASES  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regularization_golden.npz"))


def _rel(a, b):
    return np.linalg.norm(a - b, axis=-1) / (np.linalg.norm(b, axis=-1) + 1e-3)


@pytest.mark.parametrize("key,kind,w,h,spp,mpl,rr,direct", CASES)
def test_regularized_path_tracer_vs_oracle_and_reference_golden(built_lib, orc, key, kind, w, h, spp, mpl, rr, direct):
    s = ctl.Scene(kind, w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s)
    t.setParameter("MaxPathLength", mpl); t.setParameter("RRStartDepth", rr); t.setParameter("Direct", direct); t.setParameter("Regularization", 1)
    assert t.getParameter("Regularization") == 1
    imgs = {}
    for stop in (1, 0):
        t.setParameter("StopZeroThroughput", stop)
        r0 = t.getTotalRays()
        for p in range(spp): t.DoPass(p == 0)
        t.synchronize(); imgs[stop] = (t.readAccumulator(), t.getTotalRays() - r0)
    # fused passes and the frame call trace the same paths
    t.DoPasses(spp, new_trace=True); t.synchronize(); fused = t.readAccumulator()
    assert np.array_equal(fused["weight_sum"], imgs[0][0]["weight_sum"]) and np.allclose(fused["rgb"], imgs[0][0]["rgb"], rtol=2e-5, atol=1e-6)
    ref, ref_rays = orc.render(s.view, w, h, n_passes=spp, max_path_length=mpl, rr_start=rr, direct=direct | 2)
    floor = 0.97 if "c3" in key else 0.99
    for stop in (1, 0):
        img, rays = imgs[stop]
        assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and np.array_equal(img["weight_sum"], GOLD[key + "_weight"])
        r = _rel(img["rgb"], ref["rgb"]); g = _rel(img["rgb"], GOLD[key + "_rgb"])
        print(key, "stop", stop, "frac vs oracle", float((r <= 1e-3).mean()), "vs reference golden", float((g <= 1e-3).mean()), "rays", rays, "oracle", ref_rays, "reference", int(GOLD[key + "_rays"][0]))
        assert (r <= 1e-3).mean() >= floor and (g <= 1e-3).mean() >= floor
    assert abs(imgs[1][1] - ref_rays) <= 2e-3 * ref_rays                      # the oracle's default is the product's (zero-throughput paths stop)
    assert abs(imgs[0][1] - int(GOLD[key + "_rays"][0])) <= (5e-3 if "c3" in key else 1e-3) * imgs[0][1]   # the reference's own count
    t.close()


def test_regularization_differs_from_pathtrace_and_switches_back(built_lib):
    w, h = 96, 64
    s = ctl.Scene("soup", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 4)
    t.DoPasses(2, new_trace=True); t.synchronize(); a = t.readAccumulator(); ra = t.getRaysInLastPass()
    t.setParameter("Regularization", 1); t.DoPasses(2, new_trace=True); t.synchronize(); b = t.readAccumulator(); rb = t.getRaysInLastPass()
    t.setParameter("Regularization", 0); t.DoPasses(2, new_trace=True); t.synchronize(); c = t.readAccumulator(); rc = t.getRaysInLastPass()
    assert np.array_equal(a["weight_sum"], b["weight_sum"]) and not np.allclose(a["rgb"], b["rgb"], rtol=1e-3)
    assert ra == rc and np.allclose(a["rgb"], c["rgb"], rtol=2e-5, atol=1e-6) and rb != ra
    t.setParameter("Regularization", 1); t.setParameter("MaxPathLength", 256)
    with pytest.raises(RuntimeError, match="MaxPathLength <= 255"):
        t.DoPass(True)
    t.close()
