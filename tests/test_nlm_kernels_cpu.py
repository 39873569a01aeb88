"""The NonLocalMeansFilter CUDA kernels (csrc/nlm_filter.cuh) checked where no GPU exists: tests/nlm_emulate.cpp compiles the SAME kernel source for the host
(g++; __shared__ -> static, thread / block indices set by a loop) and runs the launch sequence of ctl_apply_image_pipeline's filter_type 5 branch.  Against the
goldens minted from the reference's own kernels (tests/golden/make_nlm_golden.py): RGBE stage and final RGBA8 identical; weights identical except where the
kernels' exp (double, rounded) and the host's expf differ in the last bit.  The GPU run of the same kernels is tests/test_gpu_zz_nlm.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = np.load(os.path.join(HERE, "golden", "nlm_golden.npz"))
CASES = {"a": (0.45, 1.0), "b": (1.0, 5.0), "wide": (0.45, 1.0), "default": (0.45, 0.005)}


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("nlm_emu") / "libnlm_emu.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(HERE, "nlm_emulate.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.emu_nlm_filter.argtypes = [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_float] * 3 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]; L.emu_nlm_filter.restype = None
    return L


def _run(emu, img, var, k, s2, weights=None):
    h, w = img.shape[:2]
    compute = weights is None
    wts = np.zeros((169, w * h), np.float32) if compute else np.ascontiguousarray(weights)
    stage = np.zeros((h, w, 4), np.uint8); out = np.zeros((h, w, 4), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_nlm_filter(p(np.ascontiguousarray(img)), p(np.ascontiguousarray(var)), w, h, 0.0, k, s2, p(wts), int(compute), p(stage), p(out))
    return stage, out, wts


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernel_source_on_host_vs_reference_goldens(emu, orc, name):
    k, s2 = CASES[name]
    img, var = GOLD[name + "_img"], GOLD[name + "_var"]
    stage, out, wts = _run(emu, img, var, k, s2)
    assert np.array_equal(stage, GOLD[name + "_rgbe"])
    h, w = img.shape[:2]
    o_stage, o_w = orc.nlm_filter(np.ascontiguousarray(img).view(api.PIXEL_DTYPE).reshape(h, w), np.ascontiguousarray(var).view(ctl.VARIANCE_DTYPE).reshape(-1), k, s2)
    assert np.array_equal(out, orc.pipeline_from_stage2(o_stage, api.ImagePipeline(filter_type=5))[0])    # fused copyFilteredToOutput
    dev = wts.T                                                                                            # [169][pixel] -> the reference's [pixel][169]
    assert (dev.view(np.uint32) == o_w.view(np.uint32)).mean() > 0.9995 and np.abs(dev - o_w).max() <= 6e-8


def test_kernel_source_on_host_stale_weights(emu):
    img0, var0 = GOLD["a_img_early"], GOLD["a_var_early"]
    _, _, w0 = _run(emu, img0, var0, 0.45, 1.0)
    stage, _, w1 = _run(emu, GOLD["a_img"], GOLD["a_var"], 0.45, 1.0, weights=w0)
    assert np.array_equal(stage, GOLD["a_rgbe_stale"]) and np.array_equal(w0, w1)
