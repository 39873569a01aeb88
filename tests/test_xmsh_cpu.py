"""`.xmsh` import / export, CPU side.  tests/golden/two_light_room.xmsh was written by the reference's OWN writer
(Mesh::CompileMesh + SplitBVHBuilder, compiled into oracle/_ref; tests/golden/make_golden.py); the product's reader must turn it into a
scene view that renders exactly like the scene built from the same triangles here."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
from scene_fixtures import two_light_room, two_light_room_arrays, TWO_LIGHT_CAMERA

HERE = os.path.dirname(os.path.abspath(__file__))
XMSH = os.path.join(HERE, "golden", "two_light_room.xmsh")
GOLD = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


def _materials(s):
    raw = bytes((C.c_char * (64 * s.view.n_materials)).from_address(C.addressof(s.view.materials.contents)))
    return [struct.unpack("4I12f", raw[64 * i:64 * i + 64]) for i in range(s.view.n_materials)]


def _relevant(m):
    """fields of ctl_material the BSDF of that type reads"""
    t = m[0]
    if t == 0:
        return (t, m[1], m[2], m[4:7])
    if t == 1:
        return (t, m[1], m[2], m[3], m[4:8], m[8:11], m[11], m[12:15])
    return (t, m[1], m[2], m[4:7], m[8], m[15])


def _grouped_scene(w, h):
    V, I, M, mats, emissive = two_light_room_arrays()
    order = np.argsort(M, kind="stable")
    return ctl.Scene.from_mesh(V, I.reshape(-1, 3)[order].ravel(), M[order], mats, emissive, *TWO_LIGHT_CAMERA, w, h)


def test_reader_reproduces_the_scene(built_lib, orc):
    s1 = ctl.Scene.from_xmsh(XMSH, *TWO_LIGHT_CAMERA, 64, 64)
    s2 = _grouped_scene(64, 64)
    assert s1.n_triangles == s2.n_triangles == 26 and s1.view.n_materials == 7 and s1.view.num_lights == 2 and s1.view.n_nodes == 1
    # TriangleData written by the reference's Mesh::CompileMesh (vertex normals, dpdu/dpdv, material index) == this repo's encoder, bit for bit
    assert np.array_equal(s1.array("tri_data"), s2.array("tri_data"))
    assert [_relevant(m) for m in _materials(s1)] == [_relevant(m) for m in _materials(s2)]
    assert list(s1.view.box_min) == list(s2.view.box_min) and list(s1.view.box_max) == list(s2.view.box_max) and s1.view.ray_eps == s2.view.ray_eps
    for name in ("light_tris", "light_cdf_data"):   # ShapeSets of the two area lights: same triangles; Woop slots differ with the BVH
        a, b = s1.array(name), s2.array(name)
        assert a.shape == b.shape and np.allclose(a[:, :13], b[:, :13], rtol=1e-5, atol=1e-6) if name == "light_tris" else np.array_equal(a, b)
    # the reference's SBVH: every triangle referenced, leaves of <= 8, sentinel-terminated leaves
    idx = s1.array("tri_index").ravel()
    assert set((idx >> 1).tolist()) == set(range(26)) and s1.view.n_woop == len(idx) >= 26
    run = 0
    for wd in idx:
        run += 1
        if wd & 1:
            assert run <= 8; run = 0
    assert run == 0
    # same closest hits and the same images with the reference's tree and with ours
    rng = np.random.default_rng(3)
    rays = np.zeros(4000, api.RAY_DTYPE); rays["o"] = rng.uniform(-0.99, 0.99, (4000, 3)); d = rng.normal(size=(4000, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
    h1, h2 = orc.trace_rays(s1.view, rays), orc.trace_rays(s2.view, rays)
    assert np.array_equal(h1["tri_idx"], h2["tri_idx"]) and np.array_equal(h1["dist"].view(np.uint32), h2["dist"].view(np.uint32))
    a, ra = orc.render(s1.view, 64, 64, n_passes=2, max_path_length=6)
    b, rb_ = orc.render(s2.view, 64, 64, n_passes=2, max_path_length=6)
    assert np.array_equal(a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)) and ra == rb_


def test_imported_scene_vs_reference_image(built_lib, orc):
    """The reference's own PathTrace on the scene imported from its own file (golden) == the host-arithmetic oracle, bit for bit."""
    s = ctl.Scene.from_xmsh(XMSH, *TWO_LIGHT_CAMERA, 64, 64)
    ref = np.ascontiguousarray(GOLD["xmsh_two_light_image_64x64_2spp"]).view(api.PIXEL_DTYPE).reshape(64, 64)
    with orc.host_arithmetic():
        img, rays = orc.render(s.view, 64, 64, n_passes=2, max_path_length=6)
    assert np.array_equal(img["rgb"].view(np.uint32), ref["rgb"].view(np.uint32)) and rays <= int(GOLD["xmsh_two_light_rays"][0])


def test_writer_round_trip_and_multi_file_import(built_lib, orc, tmp_path):
    s2 = two_light_room(32, 32)
    p = tmp_path / "rt.xmsh"
    s2.write_xmsh(p)
    s3 = ctl.Scene.from_xmsh(p, *TWO_LIGHT_CAMERA, 32, 32)
    for name in ("bvh_nodes", "woop", "tri_index", "tri_data", "light_tris", "light_cdf_data", "meshes", "nodes"):
        assert np.array_equal(s3.array(name).view(np.uint32), s2.array(name).view(np.uint32)), name
    assert [_relevant(m) for m in _materials(s3)] == [_relevant(m) for m in _materials(s2)]
    # same header layout as the reference-written file
    a, b = open(XMSH, "rb").read(64), open(p, "rb").read(64)
    assert a[:4] == b[:4] == bytes(4) and a[4:28] == b[4:28] and a[28:32] == b[28:32] == struct.pack("I", 2)
    # two instances of the file with different transforms: two nodes, a real scene-level BVH, lights of both
    xf = np.stack([np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)]); xf[1, 0, 3] = 2.5; xf[1, 1, 1] = 0.5
    s4 = ctl.Scene.from_xmsh([XMSH, p], (1.2, 0, -4.0), (1.2, 0, 0), (0, 1, 0), 60.0, 48, 32, node_xforms=xf)
    assert s4.view.n_nodes == 2 and s4.view.n_meshes == 2 and s4.view.num_lights == 4 and s4.view.scene_start_node == 0
    img, rays = orc.render(s4.view, 48, 32, n_passes=1, max_path_length=4)
    assert (img["weight_sum"] == 1).all() and img["rgb"].mean() > 0 and rays > 48 * 32


def test_reader_error_paths(built_lib, tmp_path):
    good = open(XMSH, "rb").read()
    def expect(data, text):
        p = tmp_path / "bad.xmsh"; p.write_bytes(data)
        with pytest.raises(RuntimeError, match=text):
            ctl.Scene.from_xmsh(p, *TWO_LIGHT_CAMERA, 16, 16)
    with pytest.raises(RuntimeError, match="Could not open file"):
        ctl.Scene.from_xmsh(tmp_path / "missing.xmsh", *TWO_LIGHT_CAMERA, 16, 16)
    expect(good[:1000], "Passed end of file")
    expect(struct.pack("I", 7) + good[4:], "Mesh file parser error")
    expect(struct.pack("I", 1) + good[4:], "animated meshes")
    expect(b"", "Passed end of file")
    bad = bytearray(good); bad[4 + 24 + 4 + 2 * 48:4 + 24 + 4 + 2 * 48 + 4] = struct.pack("I", 0x30000000)   # absurd triangle count: refused before allocating
    expect(bytes(bad), "Passed end of file")
    # first material blob starts after: token, box, light count, 2 lights, triangle count, 26 triangles, material count
    m0 = 4 + 24 + 4 + 2 * 48 + 4 + 26 * 32 + 4
    bad = bytearray(good); bad[m0 + 512:m0 + 516] = struct.pack("I", 9)          # roughplastic
    expect(bytes(bad), "BSDF type 9 is not supported")
    bad = bytearray(good); bad[m0 + 528 + 64:m0 + 528 + 68] = struct.pack("I", 3)   # ImageTexture as the diffuse reflectance
    expect(bytes(bad), "not a ConstantTexture")
    for off, text in ((2656, "normal maps"), (2880, "height maps"), (3104, "alpha maps"), (88, "bssrdf")):
        bad = bytearray(good); bad[m0 + off:m0 + off + 4] = struct.pack("I", 1)
        expect(bytes(bad), text)
    bad = bytearray(good); bad[4 + 24 + 4 + 4:4 + 24 + 4 + 4 + 10] = b"nosuchmat" + bytes(1)
    expect(bytes(bad), "unknown material")


# ---- OBJ front end ---------------------------------------------------------------------------------------------------------------
OBJ = os.path.join(HERE, "golden", "obj", "room.obj")
OBJ_REF_XMSH = os.path.join(HERE, "golden", "obj", "room_ref.xmsh")    # the same file compiled by the reference's own compileobj


def test_obj_import_equals_reference_compiler_output(built_lib, orc):
    """ctl_scene_create_from_files on an .obj == reading the .xmsh that the reference's compileobj -> Mesh::CompileMesh wrote for the same
    file: vertex de-duplication order, fan triangulation, reversed winding, the single-precision number reader (Kd 0.065 becomes
    0.06500000507), (u, 1 - v), the file's normals, UV-driven dpdu / dpdv, material mapping, the area light -- TriangleData bit-identical,
    images bit-identical (the trees differ: the reference's SplitBVHBuilder vs this repo's)."""
    a = ctl.Scene.from_xmsh(OBJ_REF_XMSH, *TWO_LIGHT_CAMERA, 64, 64)
    b = ctl.Scene.from_files(OBJ, *TWO_LIGHT_CAMERA, 64, 64)
    assert a.n_triangles == b.n_triangles == 25 and a.view.n_materials == b.view.n_materials == 4 and a.view.num_lights == b.view.num_lights == 1
    assert np.array_equal(a.array("tri_data"), b.array("tri_data"))
    assert (a.array("tri_data")[:2, 5:] != 0).any() and (a.array("tri_data")[2:, 5:] == 0).all()      # only the floor carries texture coordinates
    assert [_relevant(m) for m in _materials(a)] == [_relevant(m) for m in _materials(b)]
    assert _materials(b)[1][4:7] == pytest.approx((0.63, 0.065, 0.05)) and _materials(b)[1][5] != np.float32(0.065)   # the reference's digit-accumulating reader, reproduced
    assert _materials(b)[3][0] == 2 and _materials(b)[3][8] == pytest.approx(1.45) and _materials(b)[3][15] == pytest.approx(0.8)
    assert list(a.view.box_min) == list(b.view.box_min) and list(a.view.box_max) == list(b.view.box_max)
    assert np.allclose(a.array("light_tris")[:, :13], b.array("light_tris")[:, :13]) and np.array_equal(a.array("light_cdf_data"), b.array("light_cdf_data"))
    ia, ra = orc.render(a.view, 64, 64, n_passes=2, max_path_length=6)
    ib, rb_ = orc.render(b.view, 64, 64, n_passes=2, max_path_length=6)
    assert np.array_equal(ia["rgb"].view(np.uint32), ib["rgb"].view(np.uint32)) and ra == rb_ and ia["rgb"].mean() > 0.1
    ref = np.ascontiguousarray(GOLD["obj_room_image_64x64_2spp"]).view(api.PIXEL_DTYPE).reshape(64, 64)   # the reference's own PathTrace
    with orc.host_arithmetic():
        ih, _ = orc.render(b.view, 64, 64, n_passes=2, max_path_length=6)
    assert np.array_equal(ih["rgb"].view(np.uint32), ref["rgb"].view(np.uint32))


def test_obj_import_error_paths(built_lib, tmp_path):
    def scene(obj_text, mtl_text=None):
        (tmp_path / "t.obj").write_text(obj_text)
        if mtl_text is not None:
            (tmp_path / "t.mtl").write_text(mtl_text)
        return ctl.Scene.from_files(tmp_path / "t.obj", *TWO_LIGHT_CAMERA, 16, 16)
    tri = "v 0 0 0\nv 1 0 0\nv 0 1 0\n"
    diffuse = "newmtl m\nKd 0.5 0.5 0.5\nKs 0 0 0\nillum 2\n"
    s = scene("mtllib t.mtl\n" + tri + "usemtl m\nf 1 2 3\nf -3 -2 -1\n", diffuse)
    assert s.n_triangles == 2 and s.view.n_materials == 1 and s.view.num_lights == 0
    with pytest.raises(RuntimeError, match="phong"):               # the reference's default material has Ks = 0.5 -> phong, not on the hot path
        scene(tri + "f 1 2 3\n")
    with pytest.raises(RuntimeError, match="phong"):
        scene("mtllib t.mtl\n" + tri + "usemtl m\nf 1 2 3\n", "newmtl m\nKd 0.5 0.5 0.5\nKs 0.2 0.2 0.2\nillum 2\n")
    with pytest.raises(RuntimeError, match="illum 5"):
        scene("mtllib t.mtl\n" + tri + "usemtl m\nf 1 2 3\n", "newmtl m\nillum 5\n")
    with pytest.raises(RuntimeError, match="texture maps"):
        scene("mtllib t.mtl\n" + tri + "usemtl m\nf 1 2 3\n", diffuse + "map_Kd wood.png\n")
    with pytest.raises(RuntimeError, match="did not find submeshes"):
        scene(tri)
    with pytest.raises(RuntimeError, match="Could not open file"):
        scene("mtllib missing.mtl\n" + tri + "f 1 2 3\n")
    with pytest.raises(RuntimeError, match="Could not open file"):
        ctl.Scene.from_files(tmp_path / "nope.obj", *TWO_LIGHT_CAMERA, 16, 16)


@pytest.mark.parametrize("name", ["terrain_ascii", "octa_le", "octa_be"])
def test_ply_import_equals_reference_compiler_output(built_lib, orc, name):
    """PLY front end (ascii with u / v properties and quads, binary little / big endian) == the .xmsh the reference's own compileply wrote for
    the same file: TriangleData bit-identical (incl. its quad winding and u-for-both-coordinates quirks), red default material, same hits."""
    cam = ((0, 0.3, -2.5), (0, -0.2, 0), (0, 1, 0), 50.0)
    a = ctl.Scene.from_xmsh(os.path.join(HERE, "golden", "obj", name + "_ref.xmsh"), *cam, 32, 32)
    b = ctl.Scene.from_files(os.path.join(HERE, "golden", "obj", name + ".ply"), *cam, 32, 32)
    assert a.n_triangles == b.n_triangles > 0 and np.array_equal(a.array("tri_data"), b.array("tri_data"))
    assert [_relevant(m) for m in _materials(a)] == [_relevant(m) for m in _materials(b)] == [(0, 0, 0xffffffff, (1.0, 0.0, 0.0))]
    assert list(a.view.box_min) == list(b.view.box_min) and list(a.view.box_max) == list(b.view.box_max) and b.view.num_lights == 0
    rng = np.random.default_rng(2)
    rays = np.zeros(3000, api.RAY_DTYPE); rays["o"] = rng.uniform(-1, 1, (3000, 3)) + np.array([0, 1.5, 0]); d = rng.normal(size=(3000, 3)); d[:, 1] = -abs(d[:, 1])
    rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
    ha, hb = orc.trace_rays(a.view, rays), orc.trace_rays(b.view, rays)
    assert np.array_equal(ha["tri_idx"], hb["tri_idx"]) and np.array_equal(ha["dist"].view(np.uint32), hb["dist"].view(np.uint32)) and (ha["tri_idx"] != 0xffffffff).mean() > 0.05


def test_ply_import_error_paths(built_lib, tmp_path):
    def expect(data, text):
        p = tmp_path / "bad.ply"; p.write_bytes(data)
        with pytest.raises(RuntimeError, match=text):
            ctl.Scene.from_files(p, (0, 0, -3), (0, 0, 0), (0, 1, 0), 50.0, 8, 8)
    good = open(os.path.join(HERE, "golden", "obj", "octa_le.ply"), "rb").read()
    expect(b"plx\n", "not a ply file")
    expect(good[:200], "Passed end of file|triangles or quads")
    expect(good.replace(b"property float z\n", b"property int z\n"), "float x y z")
    expect(b"ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n5 0 1 2 0 1\n", "triangles or quads")


def test_set_node_transform_equals_building_with_that_transform(built_lib, orc, tmp_path):
    """ctl_scene_set_node_transform (DynamicScene::SetNodeTransform): moving an instance re-assembles the node level -- scene-level BVH, inverse
    matrices, the node's area-light ShapeSets, scene box, ray epsilon -- exactly as if the scene had been imported with that transform; the mesh
    level is untouched."""
    cam = ((1.2, 0, -4.0), (1.2, 0, 0), (0, 1, 0), 60.0)
    x0 = np.stack([np.eye(4, dtype=np.float32)] * 3); x0[1, 0, 3] = 2.5; x0[2, 1, 3] = 2.2
    x1 = x0.copy()
    a = 0.6
    x1[1] = np.array([[np.cos(a), 0, np.sin(a), 2.8], [0, 0.7, 0, -0.3], [-np.sin(a), 0, np.cos(a), 0.4], [0, 0, 0, 1]], np.float32)
    files = [XMSH, OBJ, XMSH]
    A = ctl.Scene.from_files(files, *cam, 48, 32, node_xforms=x1)
    B = ctl.Scene.from_files(files, *cam, 48, 32, node_xforms=x0)
    before = {n: B.array(n).copy() for n in ("bvh_nodes", "woop", "tri_index", "tri_data", "meshes")}
    assert not np.array_equal(A.array("node_xf"), B.array("node_xf"))
    B.setNodeTransform(1, x1[1])
    for n in ("scene_bvh_nodes", "nodes", "node_xf", "node_inv_xf", "light_tris", "light_cdf_data", "bvh_nodes", "woop", "tri_index", "tri_data", "meshes"):
        assert np.array_equal(A.array(n).view(np.uint32), B.array(n).view(np.uint32)), n
    for n, v in before.items():
        assert np.array_equal(B.array(n).view(np.uint32), v.view(np.uint32)), n                     # mesh level untouched
    assert list(A.view.box_min) == list(B.view.box_min) and list(A.view.box_max) == list(B.view.box_max) and A.view.ray_eps == B.view.ray_eps
    assert A.view.scene_start_node == B.view.scene_start_node and A.view.num_lights == B.view.num_lights == 5
    ia, ra = orc.render(A.view, 48, 32, n_passes=1, max_path_length=4); ib, rb_ = orc.render(B.view, 48, 32, n_passes=1, max_path_length=4)
    assert np.array_equal(ia["rgb"].view(np.uint32), ib["rgb"].view(np.uint32)) and ra == rb_
    with pytest.raises(RuntimeError, match="no such node"):
        B.setNodeTransform(7, np.eye(4))


def test_structural_validation_of_imported_trees(built_lib, tmp_path):
    """A tree that would make the traversal kernel read out of bounds or spin never reaches the GPU: the .xmsh reader checks child / leaf references,
    tree shape and leaf-run end flags (csrc/validate.cpp), and ctl_validate_scene_view does the same for whole views."""
    import struct
    for kind in ("cornell", "cornell7", "soup", "c4"):
        ctl.Scene(kind, 32, 32, n_hint=40).validate()
    src = ctl.Scene("soup", 32, 32, n_hint=200)
    good_path = tmp_path / "good.xmsh"
    src.write_xmsh(good_path)
    good = good_path.read_bytes()
    ctl.Scene.from_xmsh(good_path, *TWO_LIGHT_CAMERA, 16, 16).validate()
    nodes = src.array("bvh_nodes")
    at = good.find(nodes.tobytes()[:64])
    assert at > 0
    n_nodes = struct.unpack_from("<Q", good, at - 8)[0]         # mesh 0 of the scene
    n_refs = struct.unpack_from("<Q", good, at + 64 * n_nodes)[0]
    idx_at = len(good) - 4 * n_refs
    assert 0 < n_nodes <= src.view.n_bvh_nodes and good[idx_at - 8:idx_at] == struct.pack("<Q", n_refs)
    n_tris = len(src.mesh_triangles(0))

    def damaged(offset, word, match):
        p = tmp_path / "bad.xmsh"
        p.write_bytes(good[:offset] + struct.pack("<i", word) + good[offset + 4:])
        with pytest.raises(RuntimeError, match=match):
            ctl.Scene.from_xmsh(p, *TWO_LIGHT_CAMERA, 16, 16)

    child0 = at + 48                                            # BVHNodeData: three float4 of planes, then (child0, child1, parent, pad)
    damaged(child0, 4 * n_nodes, "outside the node array")      # one past the last node
    damaged(child0, 6, "outside the node array")                # not a node boundary
    damaged(child0, 0, "referenced twice")                      # the root as its own child: a cycle
    damaged(child0, ~int(n_refs), "outside the reference array")
    first, second = struct.unpack_from("<ii", good, at + 48)
    assert first >= 0 or second >= 0
    if second >= 0:
        damaged(child0, second, "referenced twice")             # a DAG: both children the same subtree
    else:
        damaged(child0 + 4, first, "referenced twice")
    last_word = struct.unpack_from("<I", good, len(good) - 4)[0]
    assert last_word & 1
    damaged(len(good) - 4, last_word & ~1, "no end flag")
    damaged(idx_at, (n_tris << 1) | 1, "out of range|references triangle")
