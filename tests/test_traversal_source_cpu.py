"""The device traversal source (csrc/device/traverse.cuh) checked where no GPU exists: tests/traverse_emulate.cpp compiles it for the host and walks a host
scene view ray by ray with the set-up and result packing of the API kernels.  Against the oracle (pinned to the reference's own traversal, tests/
test_golden_cpu.py): closest hits, t / u / v, the 16-byte traversalResult records, any-hit answers and the visit counts of the roofline, bit for bit --
on instanced scenes, a single-node scene, and a re-braided view."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("trav_emu") / "libtrav_emu.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-w", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(HERE, "traverse_emulate.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.emu_trace_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]; L.emu_trace_rays.restype = None
    L.emu_intersect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]; L.emu_intersect.restype = None
    return L


def _rays(s, n, seed, tmin=0.0, tmax=3e38):
    rng = np.random.default_rng(seed)
    lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max)); ext = hi - lo
    rays = np.zeros(n, api.RAY_DTYPE); rays["o"] = rng.uniform(lo - 0.2 * ext, hi + 0.2 * ext, (n, 3)); d = rng.normal(size=(n, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmin"] = tmin; rays["tmax"] = tmax
    return rays


@pytest.mark.parametrize("kind,hint,rebraid", [("cornell", 0, 0), ("cornell7", 0, 0), ("soup", 500, 0), ("c4", 24, 0), ("c4", 24, 300), ("c2", 3000, 0)])
def test_device_traversal_source_on_host_vs_oracle(built_lib, orc, emu, kind, hint, rebraid):
    s = ctl.Scene(kind, 32, 32, n_hint=hint)
    if rebraid:
        s.setRebraid(rebraid)
    n = 6000
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rays = _rays(s, n, 21)
    out = np.zeros(n, api.TRACE_RESULT_DTYPE); cnt = np.zeros(3, np.uint64)
    emu.emu_trace_rays(C.byref(s.view), n, p(rays), p(out), p(cnt))
    o, oc = orc.trace_rays(s.view, rays, counts=True, pseudo_nodes=True)
    assert out.tobytes() == o.tobytes()
    assert [int(x) for x in cnt] == oc
    assert 0.05 < (o["tri_idx"] != 0xffffffff).mean()
    diag = float(np.linalg.norm(np.array(list(s.view.box_max)) - np.array(list(s.view.box_min))))
    seg = _rays(s, n, 22, tmin=1e-3 * diag, tmax=0.35 * diag)                    # finite segments: intersectKernel honours tmin / tmax
    for any_hit in (0, 1):
        res = np.zeros(n, api.RESULT16_DTYPE)
        emu.emu_intersect(C.byref(s.view), n, p(seg), p(res), any_hit)
        ref = orc.intersect(s.view, seg, any_hit=bool(any_hit), pseudo_nodes=True)
        assert res.tobytes() == ref.tobytes(), any_hit
        assert (ref["tri_idx"] == -1).any() and (ref["tri_idx"] != -1).any()
