"""In-tree build of the C-ABI shared library (sm_100a only) and its C++ adapter smoke binary.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot.  `python -m cudatracerlib_b200.build` rebuilds when sources are newer.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libctl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # no implicit FMA contraction: the only fused ops are the explicit fmaf() calls (see device/dmath.cuh)
    "-fmad=false",
    "-shared", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
]
SOURCES = ["ctl_api.cu", "scene_builder.cpp", "sbvh_builder.cpp", "xmsh.cpp", "obj_import.cpp", "validate.cpp", "staging.cpp"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(HERE, "..", "include", "ctl_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
