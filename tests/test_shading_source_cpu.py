"""The device shading source (csrc/device/shading.cuh) checked where no GPU exists: tests/shading_emulate.cpp compiles it for the host.  BSDF sampling /
evaluation / pdf of the three hot-path BSDFs against the golden tables minted from the reference's OWN BSDF code (oracle/_ref, tests/golden/make_golden.py)
at the oracle's tolerance (2e-5: same formulas, libm), and against the oracle itself, which shares every expression: identical bits.  DiffuseLight::sampleDirect
against the oracle on the two-light scene."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
from test_golden_cpu import GOLD, MATS, _mat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shade_emu") / "libshade_emu.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-w", "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                        os.path.join(HERE, "shading_emulate.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.emu_bsdf_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]; L.emu_bsdf_probe.restype = None
    L.emu_light_sample_direct.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]; L.emu_light_sample_direct.restype = None
    return L


def _probe(emu, m, wi, sx, sy):
    w = np.ascontiguousarray(wi, np.float32); out = np.zeros(9, np.float32); f = np.zeros(3, np.float32); pdf = np.zeros(1, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_bsdf_probe(C.byref(m), p(w), sx, sy, p(out), p(f), p(pdf))
    return np.concatenate([out, f, pdf])


@pytest.mark.parametrize("name", sorted(MATS))
def test_device_bsdf_source_on_host_vs_reference_tables(emu, orc, name):
    m = _mat(**MATS[name]); ref = GOLD[f"bsdf_{name}"]; k = 0
    for two_sided in (0, 1):
        m.flags = two_sided
        for wi in GOLD["bsdf_wi"]:
            for (sx, sy) in GOLD["bsdf_samples"]:
                got = _probe(emu, m, wi, float(sx), float(sy))
                o9, f3, pdf = orc.bsdf_probe(m, wi, float(sx), float(sy))
                assert got.tobytes() == np.concatenate([o9, f3, [pdf]]).astype(np.float32).tobytes(), (name, wi, sx, sy)   # same expressions, same libm
                if two_sided:
                    continue
                r = ref[k]; k += 1
                if not np.any(r[:3]):   # failed sample: the reference leaves wo / sampledType / eta unset
                    assert not np.any(got[:3]); sel = [9, 10, 11, 12]
                else:
                    assert int(got[7]) == int(r[7]); sel = list(range(13))
                assert np.allclose(got[sel], r[sel], rtol=2e-5, atol=1e-7), (name, wi, sx, sy, got, r)


def test_device_light_sampling_source_on_host_vs_oracle(built_lib, emu, orc):
    from scene_fixtures import two_light_room
    s = two_light_room(32, 32)
    rng = np.random.default_rng(5)
    assert s.view.num_lights == 2
    n_ok = 0
    for light in range(2):
        for _ in range(200):
            ref = rng.uniform(-0.9, 0.9, 3).astype(np.float32); nrm = rng.normal(size=3); nrm = (nrm / np.linalg.norm(nrm)).astype(np.float32)
            sx, sy = (float(v) for v in rng.uniform(0, 1, 2).astype(np.float32))
            out = np.zeros(11, np.float32)
            emu.emu_light_sample_direct(C.byref(s.view), light, ref.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), sx, sy, out.ctypes.data_as(C.c_void_p))
            o = orc.light_sample_direct(s.view, light, ref, nrm, sx, sy)
            assert out.tobytes() == o.tobytes()
            n_ok += out[3] > 0
    assert n_ok > 50
