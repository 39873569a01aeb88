"""Several GPUs of one node through the C ABI's communicator (csrc/ctl_comm.cu: NCCL loaded at run time): tiles per rank + one reduce of the PixelData
accumulators must give the single-GPU image (same paths; only the order of float additions into a pixel may differ).  Needs >= 2 devices
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess

import numpy as np
import pytest

import cudatracerlib_b200 as ctl

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("kind,w,h", [("cornell7", 200, 136), ("soup", 256, 144)])
def test_single_process_multi_gpu_equals_single_gpu(built_lib, kind, w, h):
    n = min(_n_devices(), 4)
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices")
    s = ctl.Scene(kind, w, h)
    one = ctl.PathTracer(w, h, device=0); one.InitializeScene(s); one.setParameter("MaxPathLength", 6)
    one.DoPasses(4, new_trace=True); one.synchronize()
    ref = one.readAccumulator(); ref_rays = one.getRaysInLastPass()
    ts = [ctl.PathTracer(w, h, device=d) for d in range(n)]
    for t in ts:
        t.InitializeScene(s); t.setParameter("MaxPathLength", 6)
    ctl.PathTracer.commInitAll(ts)
    for frame in range(2):                                      # twice: the reduce leaves the non-root accumulators alone, new_trace clears them
        for d, t in enumerate(ts):
            t.DoPasses(4, new_trace=True, tile=(32, 32), part=d, n_parts=n)
        ctl.PathTracer.commReduceAccumAll(ts, 0)
        for t in ts:
            t.synchronize()
        img = ts[0].readAccumulator()
        assert np.array_equal(img["weight_sum"], ref["weight_sum"])
        assert np.allclose(img["rgb"], ref["rgb"], rtol=1e-5, atol=1e-7)
        assert sum(t.getRaysInLastPass() for t in ts) == ref_rays
    for t in ts + [one]:
        t.close()


def test_cpp_multi_gpu_example(built_lib, tmp_path):
    """examples/multi_gpu.cpp: C++ host code, one process, N GPUs, NCCL reduce -- no Python on the path; its `check` mode compares with one GPU."""
    n = _n_devices()
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices")
    libdir = os.path.join(ROOT, "cudatracerlib_b200"); exe = str(tmp_path / "ctl_multi_gpu")
    r = subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "examples", "multi_gpu.cpp"), "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-lctl_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "soup", f"gpus={min(n, 8)}", "frames=2", "spp=8", "640x360", "check"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert '"weights_differ": 0' in r.stdout
