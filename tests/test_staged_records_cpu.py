"""The derived records of the staged traversal kernel (csrc/staging.cpp: 64-byte leaf triangles, 64-byte instance records, the shared-memory treelet image
with re-addressed children) checked where no GPU exists: tests/staged_emulate.cpp walks one lane of the kernel's state machine over them on the host
(register top-of-stack, `stack_rows` shared rows, local overflow).  Hits, t / u / v, the 16-byte traversalResult records, any-hit answers and visit
counts must equal the oracle's bit for bit for every treelet budget and stack split -- on instanced, single-node and re-braided scenes."""
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
    so = str(tmp_path_factory.mktemp("staged_emu") / "libstaged_emu.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-w", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(HERE, "staged_emulate.cpp"), os.path.join(ROOT, "cudatracerlib_b200", "csrc", "staging.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.emu_staged_trace_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]; L.emu_staged_trace_rays.restype = C.c_int
    L.emu_staged_intersect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]; L.emu_staged_intersect.restype = C.c_int
    return L


def _rays(s, n, seed, tmin=0.0, tmax=3e38):
    rng = np.random.default_rng(seed)
    lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max)); ext = hi - lo
    rays = np.zeros(n, api.RAY_DTYPE); rays["o"] = rng.uniform(lo - 0.2 * ext, hi + 0.2 * ext, (n, 3)); d = rng.normal(size=(n, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmin"] = tmin; rays["tmax"] = tmax
    return rays


@pytest.mark.parametrize("kind,hint,rebraid", [("cornell", 0, 0), ("cornell7", 0, 0), ("soup", 500, 0), ("c4", 24, 0), ("c4", 24, 300), ("c2", 3000, 0)])
def test_staged_records_walk_equals_oracle(built_lib, orc, emu, kind, hint, rebraid):
    s = ctl.Scene(kind, 32, 32, n_hint=hint)
    if rebraid:
        s.setRebraid(rebraid)
    n = 5000
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rays = _rays(s, n, 31)
    o, oc = orc.trace_rays(s.view, rays, counts=True, pseudo_nodes=True)
    diag = float(np.linalg.norm(np.array(list(s.view.box_max)) - np.array(list(s.view.box_min))))
    seg = _rays(s, n, 32, tmin=1e-3 * diag, tmax=0.35 * diag)
    refs = [orc.intersect(s.view, seg, any_hit=bool(a), pseudo_nodes=True) for a in (0, 1)]
    seen_tl = set()
    for budget, rows in ((0, 64), (0, 0), (5, 2), (64, 3), (512, 16), (2048, 1)):
        out = np.zeros(n, api.TRACE_RESULT_DTYPE); cnt = np.zeros(3, np.uint64); info = np.zeros(4, np.int32)
        assert emu.emu_staged_trace_rays(C.byref(s.view), budget, rows, n, p(rays), p(out), p(cnt), p(info)) == 0
        assert out.tobytes() == o.tobytes(), (budget, rows)
        assert [int(x) for x in cnt] == oc, (budget, rows)
        assert info[1] <= budget
        seen_tl.add(int(info[1]))
        if rows <= 2 and kind not in ("cornell",):
            assert info[3] > rows + 1, "the overflow rows were not exercised"
        for any_hit in (0, 1):
            res = np.zeros(n, api.RESULT16_DTYPE)
            assert emu.emu_staged_intersect(C.byref(s.view), budget, rows, n, p(seg), p(res), any_hit) == 0
            assert res.tobytes() == refs[any_hit].tobytes(), (budget, rows, any_hit)
    assert max(seen_tl) > 0, "no treelet was ever built"


def test_treelet_holds_the_largest_boxes_first(built_lib, emu):
    """Budget k keeps a prefix of budget k+1's selection (best-first by area), the scene root is re-addressed into the image for multi-node scenes."""
    s = ctl.Scene("c4", 32, 32, n_hint=24)
    rays = _rays(s, 8, 1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    out = np.zeros(8, api.TRACE_RESULT_DTYPE); cnt = np.zeros(3, np.uint64)
    roots = []
    for budget in (1, 16, 200):
        info = np.zeros(4, np.int32)
        assert emu.emu_staged_trace_rays(C.byref(s.view), budget, 16, 8, p(rays), p(out), p(cnt), p(info)) == 0
        assert info[1] == budget
        roots.append(int(info[2]))
    assert all(r & 1 for r in roots[1:]), roots   # with room for more than one node the scene root (largest box) is in the image
