"""Randomised parity of the OBJ / PLY front ends against the reference's OWN compilers (oracle/_ref: compileobj / compileply -> Mesh::CompileMesh),
and robustness of the .xmsh reader against damaged files.  CPU only; the parity part needs oracle/_ref (skipped where it is not built)."""
import os
import struct

import numpy as np
import pytest

import cudatracerlib_b200 as ctl

CAM = ((0, 0.5, -4.0), (0, 0, 0), (0, 1, 0), 60.0)
HERE = os.path.dirname(os.path.abspath(__file__))


def _fmt(rng, x):
    """one of the number spellings OBJ files use"""
    k = rng.integers(0, 6)
    if k == 0:
        return "%.6f" % x
    if k == 1:
        return "%g" % x
    if k == 2:
        return "%.3f" % x
    if k == 3:
        return ("%.4e" % x)
    if k == 4:
        return ("+%.5f" % x) if x >= 0 else "%.5f" % x
    return "%d" % int(round(x * 4))


def _random_obj(rng, path, with_vt, with_vn):
    nv = int(rng.integers(8, 40))
    lines = ["# fuzz", "mtllib fuzz.mtl"]
    for _ in range(nv):
        lines.append("v " + " ".join(_fmt(rng, v) for v in rng.normal(size=3) * 1.5))
    nt = int(rng.integers(3, 12)) if with_vt else 0
    for _ in range(nt):
        lines.append("vt " + " ".join(_fmt(rng, v) for v in rng.uniform(0, 1, 2)))
    nn = int(rng.integers(3, 12)) if with_vn else 0
    for _ in range(nn):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        lines.append("vn " + " ".join(_fmt(rng, v) for v in n))
    mats = ["m%d" % i for i in range(int(rng.integers(1, 5)))]
    nf = int(rng.integers(6, 30))
    for f in range(nf):
        if f == 0 or rng.random() < 0.25:
            lines.append("usemtl " + mats[int(rng.integers(0, len(mats)))])
        k = int(rng.choice([3, 3, 4, 5]))
        vs = rng.choice(nv, size=k, replace=False)
        neg = rng.random() < 0.2
        toks = []
        for v in vs:
            vi = (int(v) - nv) if neg else int(v) + 1
            t = ""
            if with_vt:
                ti = int(rng.integers(0, nt)); t = str(ti - nt if neg else ti + 1)
            if with_vn:
                ni = int(rng.integers(0, nn)); toks.append("%d/%s/%d" % (vi, t, ni - nn if neg else ni + 1))
            elif with_vt:
                toks.append("%d/%s" % (vi, t))
            else:
                toks.append(str(vi))
        lines.append("f " + " ".join(toks))
        if rng.random() < 0.1:
            lines.append("g group%d" % f)
        if rng.random() < 0.1:
            lines.append("s %d" % int(rng.integers(0, 3)))
    open(path, "w").write("\n".join(lines) + "\n")
    mtl = []
    for i, m in enumerate(mats):
        kd = rng.uniform(0.05, 0.95, 3)
        mtl += ["newmtl " + m]
        if i % 3 == 2:
            mtl += ["Ks " + " ".join(_fmt(rng, v) for v in rng.uniform(0.5, 1, 3)), "Tf 0.75 0.75 0.75", "Ni " + _fmt(rng, rng.uniform(1.1, 1.9)), "illum 7"]
        else:
            mtl += ["Kd " + " ".join(_fmt(rng, v) for v in kd), "Ks 0 0 0", "Ni 1.0", "Tf 1 1 1", "illum 2"]
            if rng.random() < 0.4:
                mtl += ["Ke " + " ".join(_fmt(rng, v) for v in rng.uniform(1, 20, 3))]
    open(os.path.join(os.path.dirname(path), "fuzz.mtl"), "w").write("\n".join(mtl) + "\n")


def _same_scene(a, b):
    assert a.n_triangles == b.n_triangles and a.view.n_materials == b.view.n_materials and a.view.num_lights == b.view.num_lights
    assert np.array_equal(a.array("tri_data"), b.array("tri_data"))
    assert list(a.view.box_min) == list(b.view.box_min) and list(a.view.box_max) == list(b.view.box_max)
    la, lb = a.array("light_tris"), b.array("light_tris")
    assert la.shape == lb.shape and np.allclose(la[:, :13], lb[:, :13], rtol=1e-5, atol=1e-6)
    import ctypes as C
    ma = bytes((C.c_char * (64 * a.view.n_materials)).from_address(C.addressof(a.view.materials.contents)))
    mb = bytes((C.c_char * (64 * b.view.n_materials)).from_address(C.addressof(b.view.materials.contents)))
    for i in range(a.view.n_materials):
        x, y = struct.unpack("4I12f", ma[64 * i:64 * i + 64]), struct.unpack("4I12f", mb[64 * i:64 * i + 64])
        assert x[:3] == y[:3] and x[4:7] == y[4:7] and (x[0] != 2 or (x[8] == y[8] and x[15] == y[15])), (i, x, y)


@pytest.mark.parametrize("seed", range(12))
def test_obj_front_end_vs_reference_compiler_fuzz(built_lib, tmp_path, seed):
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1000 + seed)
    obj = str(tmp_path / "fuzz.obj")
    _random_obj(rng, obj, with_vt=seed % 3 == 1, with_vn=seed % 2 == 0)
    txt = open(obj).read()
    if seed % 4 == 1:
        txt = txt.replace("\n", "\r\n")                         # CRLF line ends
    if seed % 4 == 3:
        txt = txt.replace(" ", "  ").replace("\n", " \n\n")      # runs of blanks, trailing blanks, empty lines
    open(obj, "w").write(txt)
    xm = str(tmp_path / "fuzz_ref.xmsh")
    rb.compile_mesh(obj, xm)
    _same_scene(ctl.Scene.from_xmsh(xm, *CAM, 16, 16), ctl.Scene.from_files(obj, *CAM, 16, 16))


def test_obj_tab_separated_is_rejected_like_the_reference(built_lib, tmp_path):
    """The reference's tokenizer only knows blanks: a tab-separated file yields no submesh and is refused; same here, same message."""
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    obj = str(tmp_path / "fuzz.obj")
    _random_obj(np.random.default_rng(5), obj, False, True)
    open(obj, "w").write(open(obj).read().replace(" ", "\t"))
    with pytest.raises(RuntimeError):
        rb.compile_mesh(obj, str(tmp_path / "r.xmsh"))
    with pytest.raises(RuntimeError, match="did not find submeshes"):
        ctl.Scene.from_files(obj, *CAM, 16, 16)


@pytest.mark.parametrize("seed", range(6))
def test_ply_front_end_vs_reference_compiler_fuzz(built_lib, tmp_path, seed):
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2000 + seed)
    nv = int(rng.integers(6, 40)); nf = int(rng.integers(4, 30))
    V = (rng.normal(size=(nv, 3)) * 1.3).astype(np.float32)
    F = [rng.choice(nv, size=int(rng.choice([3, 3, 4])), replace=False) for _ in range(nf)]
    ply = str(tmp_path / "fuzz.ply")
    fmt = ["ascii", "binary_little_endian", "binary_big_endian"][seed % 3]
    hdr = "ply\nformat %s 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n" % (fmt, nv, nf)
    if fmt == "ascii":
        body = "".join("%s %s %s\n" % tuple(repr(float(c)) for c in v) for v in V) + "".join(" ".join([str(len(f))] + [str(int(i)) for i in f]) + "\n" for f in F)
        open(ply, "w").write(hdr + body)
    else:
        e = "<" if fmt == "binary_little_endian" else ">"
        b = hdr.encode() + b"".join(struct.pack(e + "3f", *v) for v in V) + b"".join(struct.pack(e + "B%di" % len(f), len(f), *[int(i) for i in f]) for f in F)
        open(ply, "wb").write(b + b"\n")
    xm = str(tmp_path / "fuzz_ref.xmsh")
    rb.compile_mesh(ply, xm)
    _same_scene(ctl.Scene.from_xmsh(xm, *CAM, 16, 16), ctl.Scene.from_files(ply, *CAM, 16, 16))


def test_xmsh_reader_survives_damaged_files(built_lib, tmp_path):
    """Truncations and bit flips of a valid file either still parse into a consistent scene or fail with RuntimeError -- never crash."""
    good = open(os.path.join(HERE, "golden", "obj", "room_ref.xmsh"), "rb").read()
    rng = np.random.default_rng(7)
    p = tmp_path / "d.xmsh"
    outcomes = {"ok": 0, "error": 0}
    cuts = list(range(0, 200, 7)) + [int(c) for c in rng.integers(200, len(good), 40)]
    for c in cuts:
        p.write_bytes(good[:c])
        with pytest.raises(RuntimeError):
            ctl.Scene.from_xmsh(p, *CAM, 8, 8)
    for _ in range(150):
        bad = bytearray(good)
        for _k in range(int(rng.integers(1, 4))):
            i = int(rng.integers(0, len(bad))); bad[i] ^= 1 << int(rng.integers(0, 8))
        p.write_bytes(bytes(bad))
        try:
            s = ctl.Scene.from_xmsh(p, *CAM, 8, 8)
            assert s.n_triangles == 25
            outcomes["ok"] += 1
        except RuntimeError:
            outcomes["error"] += 1
    assert outcomes["ok"] > 0 and outcomes["error"] > 0


def test_ply_vertex_layouts_outside_the_reference_reader(built_lib, tmp_path):
    """x / y / z must be three consecutive columns starting at x (what the reference reader assumes); other layouts are refused by name instead of
    indexing past the line (a header listing x last crashed the process), binary files must be position-only."""
    face = "3 0 1 2\n"
    def write(name, props, rows, fmt="ascii"):
        p = str(tmp_path / name)
        hdr = "ply\nformat %s 1.0\nelement vertex 3\n%selement face 1\nproperty list uchar int vertex_indices\nend_header\n" % (fmt, "".join("property float %s\n" % q for q in props))
        open(p, "w").write(hdr + rows + face)
        return p
    for props, rows in ((["y", "z", "x"], "0 0 0\n1 0 0\n0 1 0\n"), (["x", "nx", "y", "z"], "0 9 0 0\n1 9 0 0\n0 9 1 0\n"), (["z", "y", "x"], "0 0 0\n1 0 0\n0 1 0\n")):
        with pytest.raises(RuntimeError, match="unsupported vertex layout"):
            ctl.Scene.from_files(write("bad.ply", props, rows), *CAM, 16, 16)
    with pytest.raises(RuntimeError, match="unsupported vertex layout"):
        ctl.Scene.from_files(write("bin.ply", ["x", "y", "z", "nx"], "", fmt="binary_little_endian"), *CAM, 16, 16)
    # extra columns around a consecutive x y z block are fine in ascii files
    s = ctl.Scene.from_files(write("ok.ply", ["quality", "x", "y", "z", "nx"], "5 0 0 0 1\n5 1 0 0 1\n5 0 1 0 1\n"), *CAM, 16, 16)
    assert s.n_triangles == 1
