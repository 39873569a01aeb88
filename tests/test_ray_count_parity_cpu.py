"""Ray counts in the reference's own definition (every traceRay call counts, Kernel/TraceHelper.cu:176).  The product stops a path whose throughput is
exactly zero (StopZeroThroughput=1, no image effect); with the switch off the oracle's count must EQUAL the reference's own PathTrace run here
(oracle/_ref), image included -- that pins the switch the GPU tests then use at full resolution."""
import numpy as np
import pytest

import cudatracerlib_b200 as ctl
import oracle_binding as ob
import ref_binding as rb

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref is built only where /root/reference is mounted")


@pytest.mark.parametrize("kind,hint,w,h,depth,passes", [("cornell", 0, 64, 64, 8, 2), ("cornell7", 0, 48, 48, 8, 1), ("soup", 400, 64, 48, 8, 2), ("c4", 24, 48, 27, 8, 1), ("soup", 300, 32, 32, 32, 1)])
def test_oracle_ray_count_equals_reference_without_the_zero_throughput_stop(built_lib, kind, hint, w, h, depth, passes):
    s = ctl.Scene(kind, w, h, n_hint=hint)
    ref_img, ref_rays = rb.render(s.view, w, h, n_passes=passes, max_path_length=depth)
    with ob.host_arithmetic():
        try:
            ob.set_stop_zero_throughput(0)
            img, rays = ob.render(s.view, w, h, n_passes=passes, max_path_length=depth)
        finally:
            ob.set_stop_zero_throughput(1)
        img1, rays1 = ob.render(s.view, w, h, n_passes=passes, max_path_length=depth)
    assert rays == ref_rays                                   # the reference's count, exactly
    assert img.tobytes() == ref_img.tobytes()
    assert img1.tobytes() == ref_img.tobytes() and rays1 <= rays   # the stop changes no pixel, only drops zero-weight rays
