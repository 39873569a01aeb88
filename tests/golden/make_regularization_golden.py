"""Mints tests/golden/regularization_golden.npz from oracle/_ref: the reference's OWN PathTraceRegularization<DIRECT> (Integrators/PathTracer.cu:115-170,
KEY_Regularization) compiled by oracle/build_ref.sh from /root/reference and driven like pathKernel2<DIRECT, true> (ref_render, bit 1 of `direct`).

    bash oracle/build_ref.sh && python tests/golden/make_regularization_golden.py

Pins the oracle's restatement (tests/test_regularization_cpu.py, everywhere) and the CUDA path (tests/test_gpu_regularization.py, GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cudatracerlib_b200 as ctl
import ref_binding as rb

CASES = [  # key, scene kind, w, h, passes, MaxPathLength, RRStartDepth, Direct
    ("cornell7", "cornell7", 64, 48, 2, 8, 5, 1),
    ("soup", "soup", 64, 48, 2, 8, 2, 1),
    ("c3", "c3", 96, 54, 2, 6, 2, 1),
    ("cornell7_nodirect", "cornell7", 48, 32, 1, 5, 5, 0),
    ("c3_deep", "c3", 64, 36, 1, 12, 3, 1),
]
if __name__ == "__main__":
    out = {}
    for key, kind, w, h, spp, mpl, rr, direct in CASES:
        s = ctl.Scene(kind, w, h)
        img, rays = rb.render(s.view, w, h, n_passes=spp, max_path_length=mpl, rr_start=rr, direct=direct | 2)
        out[key + "_rgb"] = img["rgb"].copy(); out[key + "_weight"] = img["weight_sum"].copy(); out[key + "_rays"] = np.array([rays], np.uint64)
        print(key, rays, float(img["rgb"].mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "regularization_golden.npz"), **out)
