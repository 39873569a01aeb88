"""CPU-side checks of the boundary: the C-ABI library loads (no GPU needed to dlopen it), exports every symbol
include/ctl_b200.h declares, the POD layouts have the reference's sizes (SURVEY §8 header / Appendix A), and the
host scene builder emits the reference encodings.  No compute call is made on the device here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "ctl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(ctl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(built_lib, n)]
    assert not missing, missing


def test_pod_sizes_match_reference():
    # sizes probed from the reference headers: BVHNodeData 64, TriIntersectorData 48, TriangleData 32,
    # KernelMesh 20, Node 24, traversalRay 32, traversalResult 16, PixelData 28, ShapeSet::triData 64
    assert C.sizeof(api.BvhNode) == 64 and C.sizeof(api.WoopTri) == 48 and C.sizeof(api.TriData) == 32
    assert C.sizeof(api.Mesh) == 20 and C.sizeof(api.Node) == 24 and C.sizeof(api.LightTri) == 64
    assert api.RAY_DTYPE.itemsize == 32 and api.RESULT16_DTYPE.itemsize == 16 and api.PIXEL_DTYPE.itemsize == 28
    assert api.TRACE_RESULT_DTYPE.itemsize == 20  # TraceResult fields (24 B in the reference incl. padding-free 5 words + vptr-less)
    assert C.sizeof(api.Material) == 64 and C.sizeof(api.Light) == 32


def test_no_device_fails_loudly(built_lib):
    """No CPU fallback: creating a tracer without a CUDA device raises with the reference's error format."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError) as e:
        ctl.PathTracer(16, 16)
    assert "In file" in str(e.value) or "no such CUDA device" in str(e.value)


def test_sample_tables_match_oracle(built_lib, orc):
    for p in (0, 1, 3):
        d1, d2 = ctl.generate_sample_tables(p)
        o1, o2 = orc.sample_tables(p)
        assert np.array_equal(d1.view(np.uint32), o1.view(np.uint32)) and np.array_equal(d2.view(np.uint32), o2.view(np.uint32))


def test_encoders_match_oracle(built_lib, orc):
    rng = np.random.default_rng(3)
    for _ in range(50):
        v = rng.normal(size=(3, 3)).astype(np.float32) * 3
        out = (C.c_float * 12)()
        built_lib.ctl_encode_woop(v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, out)
        assert np.array_equal(np.frombuffer(out, np.float32).view(np.uint32), orc.encode_woop(v[0], v[1], v[2]).view(np.uint32))
        n = rng.normal(size=(3, 3)).astype(np.float32); n /= np.linalg.norm(n, axis=1, keepdims=True)
        uv = rng.uniform(size=6).astype(np.float32)
        td = (C.c_uint32 * 8)()
        pp = np.ascontiguousarray(v.ravel()); nn = np.ascontiguousarray(n.ravel().astype(np.float32))
        built_lib.ctl_encode_tri_data(pp.ctypes.data, nn.ctypes.data, uv.ctypes.data, 5, td)
        assert list(td) == [int(x) for x in orc.encode_tri_data(v, n, uv, 5)]


def _check_bvh(nodes_f, n_leaf_slots_or_nodes, leaf_is_node_index, prim_boxes_of_leaf):
    """Walk a reference-layout BVH: children in float4 units, ~leaf, 0x76543210 sentinel; every leaf reachable once;
    child boxes contain the primitives below them."""
    nodes = nodes_f.view(np.uint32).reshape(-1, 16)
    seen = []
    stack = [0]
    visited_inner = 0
    while stack:
        a = stack.pop()
        assert a % 4 == 0 and a // 4 < len(nodes)
        visited_inner += 1
        nf = nodes_f[a // 4]
        for ci in range(2):
            child = int(nodes[a // 4, 12 + ci].astype(np.int64))
            if child >= 0x80000000:
                child -= 1 << 32
            if ci == 0:
                lo = np.array([nf[0], nf[2], nf[8]]); hi = np.array([nf[1], nf[3], nf[9]])
            else:
                lo = np.array([nf[4], nf[6], nf[10]]); hi = np.array([nf[5], nf[7], nf[11]])
            if child == 0x76543210:
                continue
            if child < 0:
                seen.append(~child)
                blo, bhi = prim_boxes_of_leaf(~child)
                if leaf_is_node_index:   # scene level: the box contains the whole object
                    assert np.all(lo <= blo + 1e-4 * (1 + np.abs(blo))) and np.all(hi >= bhi - 1e-4 * (1 + np.abs(bhi)))
                else:                    # mesh level (split BVH): a leaf box holds the part of each referenced triangle inside it -> it must overlap them
                    assert np.all(lo <= bhi + 1e-4 * (1 + np.abs(bhi))) and np.all(hi >= blo - 1e-4 * (1 + np.abs(blo)))
            else:
                stack.append(child)
    return seen, visited_inner


@pytest.mark.parametrize("kind", ["cornell", "cornell7", "soup"])
def test_scene_builder_emits_reference_layout(built_lib, orc, kind):
    s = ctl.Scene(kind, 64, 64, n_hint=300)
    v = s.view
    meshes = s.array("meshes"); nodes = s.array("nodes"); woop = s.array("woop"); tri_index = s.array("tri_index")[:, 0]
    bvh = s.array("bvh_nodes")
    assert v.n_woop == v.n_tri_index
    total_refs = 0
    for m in meshes:
        tri_off, node_off4, tri_off4, idx_off, _ = [int(x) for x in m]
        assert node_off4 % 4 == 0 and tri_off4 % 3 == 0 and tri_off4 // 3 == idx_off
        sub = bvh[node_off4 // 4:]

        def leaf_box(first, idx_off=idx_off):
            lo = np.full(3, np.inf); hi = np.full(3, -np.inf); a = first
            while True:
                w = woop[idx_off + a].reshape(3, 4).astype(np.float64)
                M = np.array([[w[1][0], w[1][1], w[1][2], w[1][3]], [w[2][0], w[2][1], w[2][2], w[2][3]], [w[0][0], w[0][1], w[0][2], -w[0][3]], [0, 0, 0, 1]])
                Mi = np.linalg.inv(M)  # columns: v0-v2, v1-v2, n, v2
                v2 = Mi[:3, 3]; v0 = Mi[:3, 0] + v2; v1 = Mi[:3, 1] + v2
                for p in (v0, v1, v2):
                    lo = np.minimum(lo, p); hi = np.maximum(hi, p)
                if tri_index[idx_off + a] & 1:
                    break
                a += 1
            return lo, hi
        seen, _ = _check_bvh(sub, None, False, leaf_box)
        # every leaf run ends with the last-in-leaf flag; runs partition the mesh's slots
        n_refs = 0
        for first in seen:
            a = first
            while not (tri_index[idx_off + a] & 1):
                a += 1
            n_refs += a - first + 1
            assert a - first + 1 <= 8  # maxLeafSize 8 (BVHBuilderHelper.cpp:119)
        total_refs += n_refs
    assert total_refs == v.n_woop >= v.n_tri_data
    # every mesh triangle is referenced at least once (spatial splits may reference one several times)
    for k, m in enumerate(meshes):
        tri_off, _, _, idx_off, _ = [int(x) for x in m]
        idx_end = int(meshes[k + 1][3]) if k + 1 < len(meshes) else v.n_tri_index
        tri_end = int(meshes[k + 1][0]) if k + 1 < len(meshes) else v.n_tri_data
        assert set((tri_index[idx_off:idx_end] >> 1).tolist()) == set(range(tri_end - tri_off))
    # scene level: one object per leaf, leaf = ~nodeIdx
    sb = s.array("scene_bvh_nodes")
    if v.scene_start_node >= 0:
        seen, _ = _check_bvh(sb, None, True, lambda i: (np.array(list(v.box_max)), np.array(list(v.box_min))))
        assert sorted(seen) == list(range(v.n_nodes))
    else:
        assert ~v.scene_start_node < v.n_nodes
    assert v.ray_eps == pytest.approx(1e-4 * np.linalg.norm(np.array(list(v.box_max)) - np.array(list(v.box_min))), rel=1e-5)


def test_oracle_traversal_vs_brute_force(built_lib, orc):
    """The oracle's two-level BVH traversal finds the same closest triangle as an O(N) Woop scan of all references."""
    s = ctl.Scene("soup", 64, 64, n_hint=300)
    v = s.view
    rng = np.random.default_rng(11)
    n = 400
    rays = np.zeros(n, api.RAY_DTYPE)
    lo = np.array(list(v.box_min)); hi = np.array(list(v.box_max))
    rays["o"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"] = d.astype(np.float32); rays["tmax"] = 3e38
    res = orc.trace_rays(v, rays)
    woop = s.array("woop").astype(np.float64).reshape(-1, 3, 4)
    tri_index = s.array("tri_index")[:, 0]
    meshes = s.array("meshes"); nodes = s.array("nodes"); inv = s.array("node_inv_xf").astype(np.float64).reshape(-1, 4, 4)
    for i in range(n):
        best = (np.inf, -1)
        for ni, nd in enumerate(nodes):
            m = meshes[int(nd[0])]
            o4 = inv[ni] @ np.append(rays["o"][i].astype(np.float64), 1.0); o = o4[:3] / o4[3]
            dd = inv[ni][:3, :3] @ rays["d"][i].astype(np.float64)
            n_slots = (int(meshes[int(nd[0]) + 1][3]) if int(nd[0]) + 1 < len(meshes) else v.n_woop) - int(m[3])
            W = woop[int(m[3]):int(m[3]) + n_slots]
            Oz = W[:, 0, 3] - W[:, 0, :3] @ o; iDz = 1.0 / (W[:, 0, :3] @ dd)
            t = Oz * iDz
            u = W[:, 1, 3] + W[:, 1, :3] @ o + t * (W[:, 1, :3] @ dd)
            vv = W[:, 2, 3] + W[:, 2, :3] @ o + t * (W[:, 2, :3] @ dd)
            ok = (t > v.ray_eps) & (u >= 0) & (vv >= 0) & (u + vv <= 1) & np.isfinite(t)
            if ok.any():
                k = np.argmin(np.where(ok, t, np.inf))
                if t[k] < best[0]:
                    best = (t[k], (int(tri_index[int(m[3]) + k]) >> 1) + int(m[0]))
        if best[1] < 0:
            assert res["tri_idx"][i] == 0xffffffff
        else:
            assert res["tri_idx"][i] != 0xffffffff
            assert abs(res["dist"][i] - best[0]) <= 1e-4 * max(1.0, best[0])


def test_oracle_render_is_partition_invariant(orc):
    """RNG is a pure function of (pass, pixel index, dimension): rendering windows separately equals the whole image."""
    s = ctl.Scene("cornell", 48, 48)
    whole, rays = orc.render(s.view, 48, 48, n_passes=1, max_path_length=6, n_threads=2)
    parts = np.zeros((48, 48), api.PIXEL_DTYPE)
    r2 = 0
    for win in ((0, 0, 24, 48), (24, 0, 48, 24), (24, 24, 48, 48)):
        _, r = orc.render(s.view, 48, 48, n_passes=1, max_path_length=6, window=win, n_threads=2, img=parts)
        r2 += r
    assert r2 == rays
    assert np.allclose(parts["rgb"], whole["rgb"], rtol=1e-6, atol=1e-7) and np.array_equal(parts["weight_sum"], whole["weight_sum"])


def test_tile_owner_partition():
    own = ctl.tile_owner(100, 70, 16, 16, 3)
    assert own.shape == (70, 100) and set(np.unique(own)) == {0, 1, 2}
    assert own[0, 0] == 0 and own[0, 16] == 1 and own[0, 32] == 2 and own[16, 0] == (7 % 3)


def _build_adapter_check(tmp_path):
    import subprocess
    exe = str(tmp_path / "adapter_check")
    libdir = os.path.dirname(api.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "adapter_check.cpp"), "-o", exe, "-L" + libdir, "-lctl_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_adapter_compiles_and_fails_loudly_without_gpu(built_lib, tmp_path):
    """include/b200_path_tracer.hpp (the Tracer<true>-shaped C++ adapter) builds against the C ABI with plain g++."""
    import subprocess, torch
    exe = _build_adapter_check(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "no device" in r.stdout


def test_split_bvh_builder_vs_plain_sah(built_lib, orc, monkeypatch):
    """The mesh builder is a split BVH (csrc/sbvh_builder.cpp; the reference's meshes come from SplitBVHBuilder) that keeps, per mesh, the
    better of the tree with spatial splits and the tree without, judged by random-walk sample rays.  Against the plain binned-SAH builder kept
    for A/B (CTL_BVH_BUILDER=sah): identical closest hits, bit for bit, on a scene with long thin triangles (the config-4 generator at a small
    size); the spatially split tree (forced with CTL_SBVH_ALPHA) duplicates references moderately and needs fewer triangle tests and node pops."""
    from cudatracerlib_b200 import api
    s1 = ctl.Scene("c4", 64, 64, n_hint=24)
    s1b = ctl.Scene("c4", 64, 64, n_hint=24)
    monkeypatch.setenv("CTL_BVH_BUILDER", "sah")
    s0 = ctl.Scene("c4", 64, 64, n_hint=24)
    monkeypatch.delenv("CTL_BVH_BUILDER")
    monkeypatch.setenv("CTL_SBVH_ALPHA", "1e-5")
    s2 = ctl.Scene("c4", 64, 64, n_hint=24)
    monkeypatch.delenv("CTL_SBVH_ALPHA")
    assert s0.view.n_woop == s0.n_triangles == s1.n_triangles and s1.n_triangles <= s1.view.n_woop <= s2.view.n_woop
    assert s2.n_triangles < s2.view.n_woop <= 1.35 * s2.n_triangles
    for name in ("bvh_nodes", "woop", "tri_index"):
        assert np.array_equal(s1.array(name).view(np.uint32), s1b.array(name).view(np.uint32))       # deterministic (threads only change the schedule)
    assert np.array_equal(s0.array("tri_data"), s1.array("tri_data"))
    rng = np.random.default_rng(8)
    lo = np.array(list(s1.view.box_min)); hi = np.array(list(s1.view.box_max))
    rays = np.zeros(6000, api.RAY_DTYPE); rays["o"] = rng.uniform(lo, hi, (6000, 3)); d = rng.normal(size=(6000, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmax"] = 3e38
    a, ca = orc.trace_rays(s0.view, rays, counts=True)
    for s in (s1, s2):
        b, cb = orc.trace_rays(s.view, rays, counts=True)
        assert np.array_equal(a["tri_idx"], b["tri_idx"]) and np.array_equal(a["node_idx"], b["node_idx"])
        for f in ("dist", "u", "v"):
            assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
        assert cb[0] <= ca[0]
    # spatial splits: fewer node pops and triangle tests -- modest at this size (24 spheres + 3.4 K foliage triangles); on the full 1 M-triangle
    # config 4 the oracle measures 95 vs 140 inner nodes and 12 vs 25 triangle tests per ray (profiles/r01u_sbvh_builder.log)
    assert cb[0] < 0.95 * ca[0] and cb[1] < 0.85 * ca[1], (ca, cb)
    assert api.traversal_bytes(cb, len(rays)) < 0.95 * api.traversal_bytes(ca, len(rays))


def test_double_ray_buffer_header_compiles_for_sm100a(built_lib, tmp_path):
    """include/b200_double_ray_buffer.cuh + its test application cross-compile for sm_100a without a GPU (the run is in tests/test_gpu_wavefront_pt.py)."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    obj = tmp_path / "drb_check.o"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-fmad=false", "-I", os.path.join(root, "include"), "-c",
                    os.path.join(root, "tests", "drb_check.cu"), "-o", str(obj)], check=True)
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass and "ATOM" in sass.upper()   # the queue counters are device atomics


def test_scene_from_mesh_validates_and_survives_degenerate_geometry(built_lib, orc):
    """ctl_scene_create_from_mesh refuses indices outside the vertex / material arrays (the reference's Mesh::CompileMesh trusts its compilers;
    a C ABI cannot), and the mesh builder terminates with a valid tree on geometry SAH cannot separate: coincident, collinear and point
    triangles, one far outlier, sticks through one point."""
    mat = api.Material()                                               # zero-filled record = diffuse, black
    cam = ((0, 0, -5.0), (0, 0, 0), (0, 1, 0), 60.0, 16, 16)
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    z1 = np.zeros((1, 3), np.float32)
    for idx, mi, msg in (([0, 1, 5], [0], "vertex index 5 out of range"), ([0, 1, 2], [3], "material index 3 out of range")):
        with pytest.raises(RuntimeError, match=msg):
            ctl.Scene.from_mesh(tri, np.array(idx, np.uint32), np.array(mi, np.uint8), [mat], z1, *cam)
    with pytest.raises(RuntimeError, match="empty mesh"):
        ctl.Scene.from_mesh(tri, np.zeros(0, np.uint32), np.zeros(0, np.uint8), [mat], z1, *cam)
    rng = np.random.default_rng(0)
    sticks = []
    for _ in range(600):
        a = rng.normal(size=3) * 10
        sticks += [a, -a + rng.normal(size=3) * 0.01, a + rng.normal(size=3) * 0.001]
    outlier = rng.normal(size=(900, 3)); outlier[:3] *= 1e18
    bad_nan = rng.normal(size=(300, 3)); bad_nan[7, 1] = np.nan
    bad_inf = rng.normal(size=(300, 3)); bad_inf[11, 2] = np.inf
    cases = {"nan vertex": bad_nan, "inf vertex": bad_inf, "identical": np.tile(tri, (300, 1)), "collinear": np.tile(np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]], np.float32), (60, 1)),
             "points": np.zeros((180, 3), np.float32), "outlier": outlier, "sticks": np.array(sticks)}
    for name, V in cases.items():
        nt = len(V) // 3
        s = ctl.Scene.from_mesh(V.astype(np.float32), np.arange(3 * nt, dtype=np.uint32), np.zeros(nt, np.uint8), [mat], z1, *cam)
        v = s.view
        assert s.n_triangles == nt and v.n_woop >= nt and 1 <= v.n_bvh_nodes <= 2 * v.n_woop, name
        s.validate()                                                        # tree shape, leaf runs, depth within the traversal stack
        refs = s.array("tri_index")[:, 0]
        assert set((refs >> 1).tolist()) == set(range(nt)), name            # every triangle is referenced
        assert np.bincount(refs & 1)[1] >= 1                                # leaf terminators present
        # a ray through the middle of every (non-degenerate) triangle finds something at or before it
        if name in ("identical", "sticks"):
            P = V.reshape(-1, 3, 3).astype(np.float64); c = P.mean(axis=1); n = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
            ok = np.linalg.norm(n, axis=1) > 1e-12
            n = n[ok] / np.linalg.norm(n[ok], axis=1, keepdims=True); c = c[ok]
            rays = np.zeros(len(c), api.RAY_DTYPE); rays["o"] = c + n * 0.5; rays["d"] = -n; rays["tmax"] = 3e38
            r = orc.trace_rays(v, rays)
            hit = r["tri_idx"] != 0xffffffff
            assert hit.mean() > 0.98 and (r["dist"][hit] <= 0.5 * (1 + 1e-3)).all(), (name, hit.mean())


def test_validate_scene_view_names_the_first_problem(built_lib):
    """ctl_validate_scene_view on views a caller filled by hand: each damaged array is reported, the untouched view passes."""
    import ctypes as C
    s = ctl.Scene("cornell7", 32, 32)
    L = api.lib()

    def check(field, arr, ctype, expect):
        v = api.SceneView.from_buffer_copy(s.view)
        keep = np.ascontiguousarray(arr)
        setattr(v, field, C.cast(keep.ctypes.data, C.POINTER(ctype)))
        rc = L.ctl_validate_scene_view(C.byref(v))
        if expect is None:
            assert rc == 0, L.ctl_last_error()
        else:
            assert rc != 0 and expect in L.ctl_last_error().decode(), L.ctl_last_error()

    nodes = s.array("nodes")
    check("nodes", nodes, api.Node, None)
    bad = nodes.copy(); bad[3, 0] = 99
    check("nodes", bad, api.Node, "node 3 names mesh 99")
    top = s.array("scene_bvh_nodes").view(np.int32)
    bad = top.copy(); bad[0, 12] = ~np.int32(7)                 # 7 instances: 0..6
    check("scene_bvh_nodes", bad, api.BvhNode, "leaf names instance 7")
    bad = top.copy(); bad[0, 13] = 0
    check("scene_bvh_nodes", bad, api.BvhNode, "referenced twice")
    meshes = s.array("meshes")
    bad = meshes.copy(); bad[1, 1] += 2
    check("meshes", bad, api.Mesh, "offsets outside the arrays")
    bad = meshes.copy(); bad[1, 2] += 3
    check("meshes", bad, api.Mesh, "woop offset")
    idx = s.array("tri_index")
    bad = idx.copy(); bad[-1, 0] &= ~np.uint32(1)
    check("tri_index", bad, C.c_uint32, "no end flag")
    # a chain deeper than the traversal stack: 70 inner nodes, each with a leaf and the next node
    deep = np.zeros((70, 16), np.int32)
    for i in range(70):
        deep[i, 12] = ~0; deep[i, 13] = 4 * (i + 1) if i < 69 else ~0
    v = api.SceneView.from_buffer_copy(s.view)
    keep = np.ascontiguousarray(deep)
    v.scene_bvh_nodes = C.cast(keep.ctypes.data, C.POINTER(api.BvhNode)); v.n_scene_bvh_nodes = 70; v.scene_start_node = 0
    assert L.ctl_validate_scene_view(C.byref(v)) != 0 and "traversal stack" in L.ctl_last_error().decode()
    assert L.ctl_validate_scene_view(None) != 0
