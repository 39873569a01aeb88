"""Golden vectors of NonLocalMeansFilter from the reference's OWN kernels (oracle/_ref: ref_nlm_filter runs lines 9-159 of
Kernel/ImagePipeline/Filter/NonLocalMeansFilter.cu on the host).  Inputs: Cornell frames accumulated by the oracle with their PixelVarianceBuffer.
Run here (needs /root/reference to build oracle/_ref); writes tests/golden/nlm_golden.npz.  Weights are stored as SHA-256 of their bytes."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
import oracle_binding as ob
import ref_binding as rb

CASES = {"a": ("cornell", 40, 28, 8, 0.45, 1.0), "b": ("cornell", 40, 28, 8, 1.0, 5.0), "wide": ("cornell7", 204, 10, 3, 0.45, 1.0), "default": ("cornell", 24, 20, 2, 0.45, 0.005)}


def frames(kind, w, h, n):
    s = ctl.Scene(kind, w, h)
    img = None; var = np.zeros(w * h, api.VARIANCE_DTYPE); snaps = []
    for p in range(n):
        img, _ = ob.render(s.view, w, h, 1, p, img=img)
        ob.variance_add_pass(var, img)
        snaps.append((img.copy(), var.copy()))
    return snaps


if __name__ == "__main__":
    out = {}
    for name, (kind, w, h, n, k, s2) in CASES.items():
        snaps = frames(kind, w, h, n)
        img, var = snaps[-1]
        rgbe, wts = rb.nlm_filter(img, var, k, s2)
        out[name + "_img"] = img; out[name + "_var"] = var; out[name + "_rgbe"] = rgbe
        out[name + "_weights_sha256"] = np.frombuffer(hashlib.sha256(wts.tobytes()).digest(), np.uint8)
        if name == "a":   # stale weights: computed on an earlier frame, applied to the last one (UpdateWeightPeriodicity > 1)
            img0, var0 = snaps[3]
            _, w0 = rb.nlm_filter(img0, var0, k, s2)
            out["a_img_early"] = img0; out["a_var_early"] = var0
            out["a_rgbe_stale"] = rb.nlm_filter(img, var, k, s2, weights=w0)[0]
    np.savez_compressed(os.path.join(HERE, "nlm_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
