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
    "-shared", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-pthread",
]
SOURCES = ["ctl_api.cu", "ctl_pipeline.cu", "ctl_bvh_gpu.cu", "ctl_comm.cu", "ctl_scene_api.cpp", "scene_builder.cpp", "sbvh_builder.cpp", "xmsh.cpp", "obj_import.cpp", "validate.cpp",
           "staging.cpp"]
OBJ_DIR = os.path.join(HERE, "..", "build", "obj")


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


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every translation unit (in parallel) and link the shared library.  extra_flags / out: A/B build variants (build_variants/)."""
    out = out or LIB
    if not force and out == LIB and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = OBJ_DIR if out == LIB else OBJ_DIR + "_" + os.path.basename(out)
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])

    def one(src):
        obj = os.path.join(obj_dir, src + ".o")
        r = subprocess.run([nvcc] + compile_flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, r
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(one, SOURCES))
    log = ""
    for src, obj, r in results:
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on " + src + ":\n" + r.stdout + r.stderr)
        log += r.stderr
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [o for _, o, _ in results] + ["-ldl", "-o", out], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
