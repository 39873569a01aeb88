"""DeferStragglers = 1: the traversal launches of a frame's wavefronts move the rays they have not finished a few iterations after their queue ran dry into
the wavefront's next launch (device/traverse_handover.cuh, k_intersect_defer); the paths of those rays run up to DeferMaxLag bounces behind.  The rays,
their hits and the ray count are those of the plain frame: weights and ray totals equal, radiance equal up to the order of the float atomics -- also
with an aggressive deferral (one iteration), every lag limit, scenes with instances / several material classes / deep paths, one part of eight, and
several wavefronts on lanes; and the frame agrees with the oracle like the plain one."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl

pytestmark = pytest.mark.gpu


def _frame(t, spp, batch, part=0, n_parts=1):
    r0 = t.getTotalRays()
    t.DoFrame(spp, batch, part=part, n_parts=n_parts); t.synchronize()
    return t.readAccumulator(), t.getTotalRays() - r0, t.getRaysInLastPass()


@pytest.mark.parametrize("kind,w,h,spp,batch,mpl,parts", [("cornell7", 160, 120, 4, 4, 8, 1), ("soup", 256, 160, 8, 8, 8, 1), ("c3", 384, 216, 8, 8, 8, 8), ("c4", 320, 180, 4, 4, 8, 1),
                                                            ("c3", 200, 120, 2, 2, 32, 1), ("c3", 256, 144, 8, 2, 6, 1)])
def test_deferred_frame_equals_plain_frame(built_lib, kind, w, h, spp, batch, mpl, parts):
    s = ctl.Scene(kind, w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", mpl)
    for part in range(min(parts, 2)):
        t.setParameter("DeferStragglers", 0)
        a, rays_a, last_a = _frame(t, spp, batch, part, parts)
        for drain, lag in ((16, 3), (1, 3), (1, 1), (200, 2)):
            t.setParameter("DeferStragglers", 1); t.setParameter("HandOverDrain", drain); t.setParameter("DeferMaxLag", lag)
            b, rays_b, last_b = _frame(t, spp, batch, part, parts)
            assert t.getNumPassesDone() == spp
            assert rays_a == rays_b and rays_a > 0, (drain, lag, rays_a, rays_b)
            if spp == batch: assert last_b == rays_b
            assert np.array_equal(a["weight_sum"], b["weight_sum"]), (drain, lag)
            assert np.allclose(a["rgb"], b["rgb"], rtol=2e-5, atol=1e-6), (drain, lag, float(np.abs(a["rgb"] - b["rgb"]).max()))
    t.close()


def test_deferred_frame_matches_oracle(built_lib, orc):
    w, h, spp = 160, 120, 4
    s = ctl.Scene("cornell7", w, h)
    t = ctl.PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", 8); t.setParameter("DeferStragglers", 1); t.setParameter("HandOverDrain", 1)
    img, rays, _ = _frame(t, spp, spp)
    ref, ref_rays = orc.render(s.view, w, h, n_passes=spp, max_path_length=8)
    rel = np.linalg.norm(img["rgb"] - ref["rgb"], axis=2) / (np.linalg.norm(ref["rgb"], axis=2) + 1e-3)
    assert np.array_equal(img["weight_sum"], ref["weight_sum"]) and (rel <= 1e-3).mean() >= 0.99 and abs(rays - ref_rays) <= 5e-3 * ref_rays
    t.close()
