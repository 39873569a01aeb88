"""ctypes binding of include/ctl_b200.h and a Tracer-shaped Python mirror of the reference plugin API.

`PathTracer` mirrors `CudaTracerLib::PathTracer : Tracer<true>` (Integrators/PathTracer.h:7-24,
Kernel/Tracer.h:100-160): Resize / InitializeScene / DoPass / getRaysInLastPass / ... keep the
reference's names and argument meaning; errors surface as RuntimeError carrying the reference's
"In file ... at line ... : msg" text (Defines.cpp:15-29).

There is NO CPU fallback: if the CUDA library is missing or no device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTL_B200_LIB") or os.path.join(_HERE, "libctl_b200.so")  # env override: A/B builds only

MAX_NUM_LIGHTS = 16


class BvhNode(C.Structure):
    _fields_ = [("a", C.c_float * 4), ("b", C.c_float * 4), ("c", C.c_float * 4), ("child0", C.c_int32), ("child1", C.c_int32),
                ("parent", C.c_uint32), ("pad", C.c_uint32)]


class WoopTri(C.Structure):
    _fields_ = [("a", C.c_float * 4), ("b", C.c_float * 4), ("c", C.c_float * 4)]


class TriData(C.Structure):
    _fields_ = [("w", C.c_uint32 * 8)]


class Mesh(C.Structure):
    _fields_ = [("tri_offset", C.c_uint32), ("bvh_node_offset", C.c_uint32), ("bvh_tri_offset", C.c_uint32), ("bvh_idx_offset", C.c_uint32),
                ("mat_offset", C.c_uint32)]


class Node(C.Structure):
    _fields_ = [("mesh_index", C.c_uint32), ("material_offset", C.c_uint32), ("instanciated_material", C.c_uint32), ("n_lights", C.c_uint32),
                ("lights", C.c_uint32 * 2)]


class Material(C.Structure):
    _fields_ = [("bsdf_type", C.c_uint32), ("flags", C.c_uint32), ("node_light_index", C.c_uint32), ("distr_type", C.c_uint32),
                ("reflectance", C.c_float * 3), ("alpha_u", C.c_float), ("eta", C.c_float * 3), ("alpha_v", C.c_float), ("k", C.c_float * 3),
                ("transmittance", C.c_float)]


class Light(C.Structure):
    _fields_ = [("radiance", C.c_float * 3), ("sum_area", C.c_float), ("tri_offset", C.c_uint32), ("cdf_offset", C.c_uint32), ("count", C.c_uint32),
                ("node_idx", C.c_uint32)]


class LightTri(C.Structure):
    _fields_ = [("p", (C.c_float * 3) * 3), ("n", C.c_float * 3), ("area", C.c_float), ("i_dat", C.c_uint32), ("t_dat", C.c_uint32), ("pad", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("sample_to_camera", C.c_float * 16), ("to_world", C.c_float * 16), ("inv_resolution", C.c_float * 2), ("resolution", C.c_float * 2)]


class ImagePipeline(C.Structure):
    """ctl_image_pipeline: filter_type -1 none / 0 box / 1 gaussian / 2 triangle / 3 mitchell / 4 lanczos / 5 non-local means (x_width = UpdateWeightPeriodicity,
    param0 = k, param1 = sigma2Scale; needs PixelVarianceBuffer=1); tonemap 0 none / 1 Reinhard05."""
    _fields_ = [("filter_type", C.c_int32), ("x_width", C.c_float), ("y_width", C.c_float), ("param0", C.c_float), ("param1", C.c_float),
                ("tonemap", C.c_int32), ("key", C.c_float), ("burn", C.c_float)]

    def __init__(self, filter_type=-1, x_width=0.5, y_width=0.5, param0=0.0, param1=0.0, tonemap=0, key=0.18, burn=0.0):
        super().__init__(filter_type, x_width, y_width, param0, param1, tonemap, key, burn)


VARIANCE_DTYPE = np.dtype([("prev_I", np.float32, 3), ("half_buffer", np.float32, 3), ("iterations_done", np.int32), ("weight", np.float32),
                           ("sum_x", np.float32), ("sum_x2", np.float32), ("num_samples_var", np.int32)])


class SceneView(C.Structure):
    _fields_ = [
        ("bvh_nodes", C.POINTER(BvhNode)), ("n_bvh_nodes", C.c_uint32),
        ("woop", C.POINTER(WoopTri)), ("n_woop", C.c_uint32),
        ("tri_index", C.POINTER(C.c_uint32)), ("n_tri_index", C.c_uint32),
        ("tri_data", C.POINTER(TriData)), ("n_tri_data", C.c_uint32),
        ("meshes", C.POINTER(Mesh)), ("n_meshes", C.c_uint32),
        ("nodes", C.POINTER(Node)), ("n_nodes", C.c_uint32),
        ("node_xf", C.POINTER(C.c_float)), ("node_inv_xf", C.POINTER(C.c_float)),
        ("scene_bvh_nodes", C.POINTER(BvhNode)), ("n_scene_bvh_nodes", C.c_uint32),
        ("scene_start_node", C.c_int32),
        ("materials", C.POINTER(Material)), ("n_materials", C.c_uint32),
        ("lights", C.POINTER(Light)), ("n_lights_buf", C.c_uint32),
        ("light_tris", C.POINTER(LightTri)), ("n_light_tris", C.c_uint32),
        ("light_cdf_data", C.POINTER(C.c_float)), ("n_light_cdf_data", C.c_uint32),
        ("num_lights", C.c_uint32),
        ("light_indices", C.c_uint32 * MAX_NUM_LIGHTS), ("light_cdf", C.c_float * MAX_NUM_LIGHTS),
        ("camera", Camera), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3), ("ray_eps", C.c_float),
        ("node_alias", C.POINTER(C.c_uint32)),
    ]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
RESULT16_DTYPE = np.dtype([("dist", np.float32), ("node_idx", np.int32), ("tri_idx", np.int32), ("bary", np.uint32)])
TRACE_RESULT_DTYPE = np.dtype([("dist", np.float32), ("u", np.float32), ("v", np.float32), ("tri_idx", np.uint32), ("node_idx", np.uint32)])
PIXEL_DTYPE = np.dtype([("rgb", np.float32, 3), ("rgb_splat", np.float32, 3), ("weight_sum", np.float32)])

_lib = None


def lib():
    """Load the C-ABI library; fail loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m cudatracerlib_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64p, fp = C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_float)
    L.ctl_last_error.restype = C.c_char_p
    L.ctl_scene_create.restype = vp; L.ctl_scene_create.argtypes = [i32, i32, i32, u32, i32]
    L.ctl_scene_create_from_mesh.restype = vp
    L.ctl_scene_create_from_mesh.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, vp, vp, vp, C.c_float, i32, i32]
    L.ctl_scene_get_view.argtypes = [vp, C.POINTER(SceneView)]
    L.ctl_scene_create_from_xmsh.restype = vp
    L.ctl_scene_create_from_xmsh.argtypes = [C.POINTER(C.c_char_p), u32, vp, vp, vp, vp, C.c_float, i32, i32]
    L.ctl_scene_write_xmsh.argtypes = [vp, u32, C.c_char_p]
    L.ctl_scene_set_node_transform.argtypes = [vp, u32, vp]
    L.ctl_update_scene_nodes.argtypes = [vp, C.POINTER(SceneView)]
    L.ctl_scene_create_from_files.restype = vp
    L.ctl_scene_create_from_files.argtypes = [C.POINTER(C.c_char_p), u32, vp, vp, vp, vp, C.c_float, i32, i32]
    L.ctl_scene_get_mesh_triangles.argtypes = [vp, u32, vp, C.POINTER(C.c_uint32)]
    L.ctl_scene_destroy.argtypes = [vp]; L.ctl_scene_destroy.restype = None
    L.ctl_scene_set_rebraid.argtypes = [vp, C.c_uint32]; L.ctl_scene_set_rebraid.restype = C.c_int
    L.ctl_validate_scene_view.argtypes = [C.POINTER(SceneView)]; L.ctl_validate_scene_view.restype = C.c_int
    L.ctl_bvh_build_gpu.argtypes = [i32, vp, u32, vp, vp, vp, vp, vp]
    L.ctl_scene_rebuild_bvh_gpu.argtypes = [vp, i32, vp]
    L.ctl_encode_woop.argtypes = [vp, vp, vp, vp]; L.ctl_encode_woop.restype = None
    L.ctl_encode_tri_data.argtypes = [vp, vp, vp, u32, vp]; L.ctl_encode_tri_data.restype = None
    L.ctl_generate_sample_tables.argtypes = [u32, vp, vp]
    L.ctl_generate_sample_tables_n.argtypes = [u32, i32, vp, vp]
    L.ctl_create.restype = vp; L.ctl_create.argtypes = [i32, i32, i32]
    L.ctl_destroy.argtypes = [vp]; L.ctl_destroy.restype = None
    L.ctl_resize.argtypes = [vp, i32, i32]
    L.ctl_set_param_i.argtypes = [vp, C.c_char_p, i32]
    L.ctl_get_param_i.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
    L.ctl_upload_scene.argtypes = [vp, C.POINTER(SceneView)]
    L.ctl_upload_samples.argtypes = [vp, vp, vp]
    L.ctl_intersect.argtypes = [vp, i32, vp, vp, i32, vp]
    L.ctl_intersect_host.argtypes = [vp, i32, vp, vp, i32]
    L.ctl_trace_rays_host.argtypes = [vp, i32, vp, vp, vp]
    L.ctl_render_pass.argtypes = [vp, i32, i32, i32, i32, i32]
    L.ctl_render_pass_tiled.argtypes = [vp, i32, i32, i32, i32, i32]
    L.ctl_render_passes_tiled.argtypes = [vp, i32, i32, i32, i32, i32, i32]
    L.ctl_render_frame_tiled.argtypes = [vp, i32, i32, i32, i32, i32, i32]
    L.ctl_submit_frame_tiled.argtypes = [vp, i32, i32, i32, i32, i32, i32]; L.ctl_acquire_frame.argtypes = [vp]; L.ctl_frames_in_flight.argtypes = [vp]
    L.ctl_comm_submit_frame.argtypes = [vp, i32, i32, i32, i32]; L.ctl_comm_submit_frame_all.argtypes = [vp, i32, i32, i32, i32, i32]
    L.ctl_wavefront_pass.argtypes = [vp, i32]
    L.ctl_wavefront_frame.argtypes = [vp, i32]
    L.ctl_read_sample_tables.argtypes = [vp, i32, vp, vp]
    L.ctl_synchronize.argtypes = [vp]
    L.ctl_read_accum.argtypes = [vp, vp]
    L.ctl_accum_device_ptr.argtypes = [vp]; L.ctl_accum_device_ptr.restype = vp
    L.ctl_resolve_srgb8.argtypes = [vp, C.c_float, vp, vp]
    L.ctl_resolve_filtered_srgb8.argtypes = [vp, C.c_float, i32, C.c_float, C.c_float, C.c_float, vp, vp]
    L.ctl_set_accum_device_ptr.argtypes = [vp, vp]
    L.ctl_apply_image_pipeline.argtypes = [vp, C.c_float, C.POINTER(ImagePipeline), vp, vp, vp]
    L.ctl_read_variance.argtypes = [vp, vp]
    L.ctl_read_nlm_weights.argtypes = [vp, vp]
    L.ctl_stream.argtypes = [vp]; L.ctl_stream.restype = vp
    L.ctl_set_stream.argtypes = [vp, vp]
    L.ctl_stats.argtypes = [vp, u64p, fp, u64p, C.POINTER(C.c_uint32)]
    L.ctl_stage_times.argtypes = [vp, fp, C.POINTER(C.c_uint32)]
    L.ctl_set_instrumented.argtypes = [vp, i32]
    L.ctl_get_visit_counts.argtypes = [vp, u64p, u64p]
    L.ctl_get_queue_sizes.argtypes = [vp, vp, vp, i32]
    L.ctl_get_captured_rays.argtypes = [vp, vp, i32]
    L.ctl_comm_get_unique_id.argtypes = [vp]; L.ctl_comm_init_rank.argtypes = [vp, vp, i32, i32]; L.ctl_comm_init_all.argtypes = [vp, i32]
    L.ctl_comm_rank.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]; L.ctl_comm_reduce_accum.argtypes = [vp, i32]; L.ctl_comm_reduce_accum_all.argtypes = [vp, i32, i32]
    L.ctl_comm_allreduce_u64.argtypes = [vp, vp, i32]; L.ctl_comm_render_frame.argtypes = [vp, i32, i32, i32, i32]; L.ctl_comm_destroy.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().ctl_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """Host scene (replaces DynamicScene for synthetic scenes); owns the flat arrays behind a SceneView."""

    KINDS = {"cornell": 0, "cornell7": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5, "soup": 6}

    def __init__(self, kind, width, height, seed=1234, n_hint=0):
        k = self.KINDS[kind] if isinstance(kind, str) else int(kind)
        self._h = lib().ctl_scene_create(k, width, height, seed, n_hint)
        if not self._h:
            raise RuntimeError(lib().ctl_last_error().decode())
        self.width, self.height = width, height
        self.view = SceneView()
        _check(lib().ctl_scene_get_view(self._h, C.byref(self.view)))

    @classmethod
    def from_mesh(cls, verts, indices, mat_index, materials, emissive, cam_pos, cam_target, cam_up, fov_deg, width, height):
        self = cls.__new__(cls)
        v = np.ascontiguousarray(verts, np.float32); ix = np.ascontiguousarray(indices, np.uint32); mi = np.ascontiguousarray(mat_index, np.uint8)
        mats = (Material * len(materials))(*materials)
        em = np.ascontiguousarray(emissive, np.float32)
        cp, ct, cu = (np.ascontiguousarray(x, np.float32) for x in (cam_pos, cam_target, cam_up))
        self._h = lib().ctl_scene_create_from_mesh(_ptr(v), len(v), _ptr(ix), len(ix) // 3 if ix.ndim == 1 else len(ix), _ptr(mi), C.cast(mats, C.c_void_p),
                                                   len(materials), _ptr(em), _ptr(cp), _ptr(ct), _ptr(cu), fov_deg, width, height)
        if not self._h:
            raise RuntimeError(lib().ctl_last_error().decode())
        self.width, self.height = width, height
        self.view = SceneView()
        _check(lib().ctl_scene_get_view(self._h, C.byref(self.view)))
        return self

    @classmethod
    def from_xmsh(cls, paths, cam_pos, cam_target, cam_up, fov_deg, width, height, node_xforms=None):
        """Flat scene import from the reference's compiled-mesh files (ctl_scene_create_from_xmsh): one mesh + node per file."""
        self = cls.__new__(cls)
        paths = [paths] if isinstance(paths, (str, bytes, os.PathLike)) else list(paths)
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        xf = np.ascontiguousarray(node_xforms, np.float32).reshape(len(paths), 16) if node_xforms is not None else None
        cp, ct, cu = (np.ascontiguousarray(x, np.float32) for x in (cam_pos, cam_target, cam_up))
        self._h = lib().ctl_scene_create_from_xmsh(arr, len(paths), _ptr(xf) if xf is not None else None, _ptr(cp), _ptr(ct), _ptr(cu), fov_deg, width, height)
        if not self._h:
            raise RuntimeError(lib().ctl_last_error().decode())
        self.width, self.height = width, height
        self.view = SceneView()
        _check(lib().ctl_scene_get_view(self._h, C.byref(self.view)))
        return self

    def setNodeTransform(self, node, xf):
        """DynamicScene::SetNodeTransform: new row-major 4x4 local-to-world matrix of instance `node`; refreshes `self.view`."""
        m = np.ascontiguousarray(xf, np.float32).reshape(16)
        _check(lib().ctl_scene_set_node_transform(self._h, int(node), _ptr(m)))
        _check(lib().ctl_scene_get_view(self._h, C.byref(self.view)))

    def mesh_triangles(self, mesh=0):
        """(nt, 3, 3) float32 source triangles of one mesh, in TriangleData order."""
        n = C.c_uint32(0)
        _check(lib().ctl_scene_get_mesh_triangles(self._h, mesh, None, C.byref(n)))
        out = np.zeros((n.value, 3, 3), np.float32)
        _check(lib().ctl_scene_get_mesh_triangles(self._h, mesh, _ptr(out), C.byref(n)))
        return out

    from_files = from_xmsh   # .xmsh and .obj files, chosen by extension (ctl_scene_create_from_files)

    def write_xmsh(self, path, mesh=0):
        """Mesh `mesh` as an .xmsh file (the output sequence of the reference's Mesh::CompileMesh)."""
        _check(lib().ctl_scene_write_xmsh(self._h, mesh, os.fsencode(path)))

    def setRebraid(self, max_entries):
        """ctl_scene_set_rebraid: open overlapping instances into up to max_entries scene-level leaves (0 = off); refreshes self.view."""
        _check(lib().ctl_scene_set_rebraid(self._h, int(max_entries)))
        _check(lib().ctl_scene_get_view(self._h, C.byref(self.view)))

    def validate(self):
        """ctl_validate_scene_view: raises RuntimeError naming the first structural problem of the view (index ranges, tree shape, stack depth)."""
        _check(lib().ctl_validate_scene_view(C.byref(self.view)))

    def rebuildBVHOnGPU(self, device=0):
        """Replace every mesh BVH by one built on the GPU (ctl_scene_rebuild_bvh_gpu); returns the device build time in ms."""
        ms = C.c_float(0)
        _check(lib().ctl_scene_rebuild_bvh_gpu(self._h, device, C.byref(ms)))
        _check(lib().ctl_scene_get_view(self._h, C.byref(self.view)))
        return ms.value

    @property
    def n_triangles(self):
        return int(self.view.n_tri_data)

    def array(self, name):
        """numpy copy of one of the flat arrays (for tests)."""
        v = self.view
        table = {
            "bvh_nodes": (v.bvh_nodes, v.n_bvh_nodes, 16, np.float32), "woop": (v.woop, v.n_woop, 12, np.float32),
            "tri_index": (v.tri_index, v.n_tri_index, 1, np.uint32), "tri_data": (v.tri_data, v.n_tri_data, 8, np.uint32),
            "scene_bvh_nodes": (v.scene_bvh_nodes, v.n_scene_bvh_nodes, 16, np.float32),
            "meshes": (v.meshes, v.n_meshes, 5, np.uint32), "nodes": (v.nodes, v.n_nodes, 6, np.uint32),
            "node_xf": (v.node_xf, v.n_nodes, 16, np.float32), "node_inv_xf": (v.node_inv_xf, v.n_nodes, 16, np.float32),
            "light_tris": (v.light_tris, v.n_light_tris, 16, np.float32), "light_cdf_data": (v.light_cdf_data, v.n_light_cdf_data, 1, np.float32),
        }
        p, n, k, dt = table[name]
        if n == 0:
            return np.zeros((0, k), dt)
        buf = C.cast(p, C.POINTER(C.c_uint32 * (n * k))).contents
        return np.frombuffer(buf, dtype=dt).reshape(n, k).copy()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.ctl_scene_destroy(h)
            self._h = None


class PathTracer:
    """Mirror of CudaTracerLib::PathTracer (Tracer<true>): the drop-in surface of the hot path."""

    def __init__(self, width, height, device=0):
        self._ctx = lib().ctl_create(device, width, height)
        if not self._ctx:
            raise RuntimeError(lib().ctl_last_error().decode())
        self.w, self.h, self.device = width, height, device
        self._scene = None
        self._new_trace = True

    # -- TracerBase surface (Kernel/Tracer.h:100-160)
    def Resize(self, w, h):
        _check(lib().ctl_resize(self._ctx, w, h)); self.w, self.h = w, h; self._new_trace = True

    def InitializeScene(self, scene):
        self._scene = scene
        _check(lib().ctl_upload_scene(self._ctx, C.byref(scene.view))); self._new_trace = True

    def UpdateSceneNodes(self, scene):
        """After Scene.setNodeTransform: re-upload only the node level (ctl_update_scene_nodes)."""
        self._scene = scene
        _check(lib().ctl_update_scene_nodes(self._ctx, C.byref(scene.view)))

    def setParameter(self, key, value):
        _check(lib().ctl_set_param_i(self._ctx, key.encode(), int(value)))

    def getParameter(self, key):
        v = C.c_int(0); _check(lib().ctl_get_param_i(self._ctx, key.encode(), C.byref(v))); return v.value

    def DoPass(self, new_trace=None, window=None):
        """One progressive pass = one path per pixel (Kernel/Tracer.h:209-248). Asynchronous."""
        nt = self._new_trace if new_trace is None else bool(new_trace)
        x0, y0, x1, y1 = window if window is not None else (0, 0, self.w, self.h)
        _check(lib().ctl_render_pass(self._ctx, int(nt), x0, y0, x1, y1)); self._new_trace = False

    def DoPassTiled(self, tile_w, tile_h, part, n_parts, new_trace=None):
        nt = self._new_trace if new_trace is None else bool(new_trace)
        _check(lib().ctl_render_pass_tiled(self._ctx, int(nt), tile_w, tile_h, part, n_parts)); self._new_trace = False

    def DoPasses(self, n_passes, new_trace=None, tile=(64, 64), part=0, n_parts=1):
        """n_passes DoPass calls fused into one wavefront (ctl_render_passes_tiled); part/n_parts select interleaved tiles."""
        nt = self._new_trace if new_trace is None else bool(new_trace)
        _check(lib().ctl_render_passes_tiled(self._ctx, int(nt), int(n_passes), tile[0], tile[1], part, n_parts)); self._new_trace = False

    def DoFrame(self, spp, batch=8, tile=(64, 64), part=0, n_parts=1):
        """ctl_render_frame_tiled: a new trace of spp passes, `batch` per wavefront, the wavefronts overlapped on two streams ("OverlapWavefronts")."""
        _check(lib().ctl_render_frame_tiled(self._ctx, spp, batch, tile[0], tile[1], part, n_parts)); self._new_trace = False

    def readSampleTables(self, table_set=0):
        d1 = np.zeros(4096 * 30, np.float32); d2 = np.zeros(4096 * 30 * 2, np.float32)
        _check(lib().ctl_read_sample_tables(self._ctx, table_set, _ptr(d1), _ptr(d2)))
        return d1, d2

    def StartNewTrace(self):
        self._new_trace = True

    def synchronize(self):
        _check(lib().ctl_synchronize(self._ctx))

    def _stats(self):
        r, t, s, p = C.c_uint64(0), C.c_uint64(0), C.c_float(0), C.c_uint32(0)
        _check(lib().ctl_stats(self._ctx, C.byref(r), C.byref(s), C.byref(t), C.byref(p)))
        return r.value, s.value, t.value, p.value

    def getRaysInLastPass(self):
        return self._stats()[0]

    def getLastTimeSpentRenderingSec(self):
        return self._stats()[1]

    def getTotalRays(self):
        return self._stats()[2]

    def getNumPassesDone(self):
        return self._stats()[3]

    # -- Image surface (Engine/Image.h): PixelData accumulator
    def readAccumulator(self):
        out = np.zeros(self.w * self.h, PIXEL_DTYPE)
        _check(lib().ctl_read_accum(self._ctx, _ptr(out)))
        return out.reshape(self.h, self.w)

    def resolveSRGB8(self, splat_scale=0.0):
        """applyImagePipeline(tracer, img) with no filter / post-process: (h, w, 4) uint8 sRGB image."""
        out = np.zeros((self.h, self.w, 4), np.uint8)
        _check(lib().ctl_resolve_srgb8(self._ctx, float(splat_scale), None, _ptr(out)))
        return out

    FILTERS = {"box": 0, "gaussian": 1, "triangle": 2}

    def resolveFilteredSRGB8(self, filter="box", x_width=0.5, y_width=0.5, alpha=2.0, splat_scale=0.0):
        """applyImagePipeline(tracer, img, filter): CanonicalFilter reconstruction -> RGBE stage -> sRGB RGBA8."""
        out = np.zeros((self.h, self.w, 4), np.uint8)
        _check(lib().ctl_resolve_filtered_srgb8(self._ctx, float(splat_scale), self.FILTERS[filter], float(x_width), float(y_width), float(alpha), None, _ptr(out)))
        return out

    def applyImagePipeline(self, pipeline, splat_scale=0.0, lum_info=False):
        """applyImagePipeline(tracer, img, filter, process) (Kernel/ImagePipeline/ImagePipeline.cu:54-84): (h, w, 4) uint8 image
        [+ (min, max, avg, log-avg luminance, scale, invWp2) of the tone-mapping stage]."""
        out = np.zeros((self.h, self.w, 4), np.uint8); lum = np.zeros(6, np.float32)
        _check(lib().ctl_apply_image_pipeline(self._ctx, float(splat_scale), C.byref(pipeline), None, _ptr(out), _ptr(lum) if lum_info else None))
        return (out, lum) if lum_info else out

    def readNlmWeights(self):
        """Weights of the last NonLocalMeansFilter application, reference layout (w*h, 169)."""
        out = np.zeros((self.w * self.h, 169), np.float32)
        _check(lib().ctl_read_nlm_weights(self._ctx, _ptr(out)))
        return out

    def readVarianceBuffer(self):
        """PixelVarianceBuffer contents (needs setParameter("PixelVarianceBuffer", 1) before the passes)."""
        out = np.zeros(self.w * self.h, VARIANCE_DTYPE)
        _check(lib().ctl_read_variance(self._ctx, _ptr(out)))
        return out.reshape(self.h, self.w)

    def resolveSRGB8Device(self, d_rgba8, splat_scale=0.0):
        """Same, into caller-owned device memory (w*h*4 bytes), asynchronous on the tracer's stream."""
        _check(lib().ctl_resolve_srgb8(self._ctx, float(splat_scale), C.c_void_p(d_rgba8), None))

    def accumDevicePtr(self):
        return lib().ctl_accum_device_ptr(self._ctx)

    def setAccumDevicePtr(self, ptr):
        _check(lib().ctl_set_accum_device_ptr(self._ctx, C.c_void_p(ptr)))

    def stream(self):
        return lib().ctl_stream(self._ctx)

    def setStream(self, cuda_stream):
        """Run on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None = own stream."""
        _check(lib().ctl_set_stream(self._ctx, C.c_void_p(cuda_stream) if cuda_stream else None))

    # -- __internal__IntersectBuffers / traceRay surfaces
    def intersect(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(len(rays), RESULT16_DTYPE)
        _check(lib().ctl_intersect_host(self._ctx, len(rays), _ptr(rays), _ptr(out), int(any_hit)))
        return out

    def intersect_device(self, n, d_rays, d_results, any_hit=False, stream=None):
        _check(lib().ctl_intersect(self._ctx, n, C.c_void_p(d_rays), C.c_void_p(d_results), int(any_hit), C.c_void_p(stream) if stream else None))

    def trace_rays(self, rays, counts=False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(len(rays), TRACE_RESULT_DTYPE)
        cnt = (C.c_uint64 * 3)()
        _check(lib().ctl_trace_rays_host(self._ctx, len(rays), _ptr(rays), _ptr(out), C.cast(cnt, C.c_void_p) if counts else None))
        return (out, [int(x) for x in cnt]) if counts else out

    def uploadSamples(self, d1, d2):
        d1 = np.ascontiguousarray(d1, np.float32); d2 = np.ascontiguousarray(d2, np.float32)
        _check(lib().ctl_upload_samples(self._ctx, _ptr(d1), _ptr(d2)))

    # -- instrumentation
    def stageTimes(self):
        ms = (C.c_float * 5)(); n = C.c_uint32(0)
        _check(lib().ctl_stage_times(self._ctx, ms, C.byref(n)))
        return list(ms), n.value

    def setInstrumented(self, on):
        _check(lib().ctl_set_instrumented(self._ctx, int(on)))

    def visitCounts(self):
        e = (C.c_uint64 * 4)(); s = (C.c_uint64 * 4)()
        _check(lib().ctl_get_visit_counts(self._ctx, e, s))
        return [int(x) for x in e], [int(x) for x in s]

    def queueSizes(self, n):
        e = np.zeros(n, np.uint32); s = np.zeros(n, np.uint32)
        _check(lib().ctl_get_queue_sizes(self._ctx, _ptr(e), _ptr(s), n))
        return e, s

    def capturedRays(self, capacity):
        out = np.zeros(capacity, RAY_DTYPE)
        n = lib().ctl_get_captured_rays(self._ctx, _ptr(out), capacity)
        if n < 0:
            raise RuntimeError(lib().ctl_last_error().decode())
        return out[:n]

    # -- multi-GPU (csrc/ctl_comm.cu): NCCL behind the C ABI
    @staticmethod
    def commUniqueId():
        """ctl_comm_get_unique_id: 128 bytes for rank 0 to hand to the other ranks."""
        buf = C.create_string_buffer(128); _check(lib().ctl_comm_get_unique_id(buf)); return buf.raw

    def commInitRank(self, unique_id, rank, n_ranks):
        _check(lib().ctl_comm_init_rank(self._ctx, C.c_char_p(unique_id), rank, n_ranks))

    @staticmethod
    def commInitAll(tracers):
        """One process driving several GPUs: tracers[i] becomes rank i."""
        arr = (C.c_void_p * len(tracers))(*[t._ctx for t in tracers]); _check(lib().ctl_comm_init_all(arr, len(tracers)))

    def commReduceAccum(self, root=0):
        _check(lib().ctl_comm_reduce_accum(self._ctx, root))

    @staticmethod
    def commReduceAccumAll(tracers, root=0):
        arr = (C.c_void_p * len(tracers))(*[t._ctx for t in tracers]); _check(lib().ctl_comm_reduce_accum_all(arr, len(tracers), root))

    def commRenderFrame(self, spp, batch=8, tile=64, root=0):
        """ctl_comm_render_frame: this rank's tiles of a frame + the reduce to `root` (asynchronous)."""
        _check(lib().ctl_comm_render_frame(self._ctx, spp, batch, tile, root))

    # -- frames in flight (ctl_submit_frame_tiled / ctl_acquire_frame): a sequence of frames as a pipeline
    def submitFrame(self, spp, batch=8, tile=(64, 64), part=0, n_parts=1):
        """One more frame on a lane of its own (asynchronous); up to "FramesInFlight" may be outstanding."""
        _check(lib().ctl_submit_frame_tiled(self._ctx, spp, batch, tile[0], tile[1], part, n_parts))

    def acquireFrame(self):
        """The context's stream waits for the oldest outstanding frame; its accumulator becomes the context's (resolve / read-back calls see it)."""
        _check(lib().ctl_acquire_frame(self._ctx))

    def framesInFlight(self):
        return lib().ctl_frames_in_flight(self._ctx)

    def commSubmitFrame(self, spp, batch=8, tile=64, root=0):
        """ctl_comm_submit_frame: this rank's tiles of one more frame; its reduce to `root` runs on the communication stream."""
        _check(lib().ctl_comm_submit_frame(self._ctx, spp, batch, tile, root))

    @staticmethod
    def commSubmitFrameAll(tracers, spp, batch=8, tile=64, root=0):
        arr = (C.c_void_p * len(tracers))(*[t._ctx for t in tracers]); _check(lib().ctl_comm_submit_frame_all(arr, len(tracers), spp, batch, tile, root))

    def commAllReduce(self, values):
        a = np.ascontiguousarray(values, np.uint64); _check(lib().ctl_comm_allreduce_u64(self._ctx, _ptr(a), len(a))); return a

    def close(self):
        if self._ctx and _lib is not None:
            _lib.ctl_destroy(self._ctx)
        self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WavefrontPathTracer(PathTracer):
    """Mirror of CudaTracerLib::WavefrontPathTracer : Tracer<true> (Integrators/PseudoRealtime/WavefrontPathTracer.h:28-66): the
    reference's own wavefront integrator over DoubleRayBuffer, SURVEY 8 f1.  Same parameter keys ("Direct", "MaxPathLength",
    "RRStartDepth"); DoPass renders one path per pixel; results equal the reference's algorithm run in the serial order of its
    queue atomics (csrc/wavefront_pt.cuh)."""

    def __init__(self, width, height, device=0):
        super().__init__(width, height, device)
        self.setParameter("MaxPathLength", 50)   # WavefrontPathTracer.h:38-42 defaults
        self.setParameter("RRStartDepth", 5)
        self.setParameter("Direct", 1)

    def DoPass(self, new_trace=None, window=None):
        if window is not None:
            raise ValueError("WavefrontPathTracer renders whole frames (its queue slots are pixel indices)")
        nt = self._new_trace if new_trace is None else bool(new_trace)
        _check(lib().ctl_wavefront_pass(self._ctx, int(nt))); self._new_trace = False

    def DoPassTiled(self, *a, **k):
        raise NotImplementedError("WavefrontPathTracer renders whole frames")

    def DoFrame(self, spp, *a, **k):
        """ctl_wavefront_frame: a new trace of spp passes, the passes overlapped on up to "OverlapLanes" streams."""
        _check(lib().ctl_wavefront_frame(self._ctx, int(spp))); self._new_trace = False

    DoPasses = DoPassTiled


def generate_sample_tables(pass_index):
    d1 = np.zeros(4096 * 30, np.float32); d2 = np.zeros(4096 * 30 * 2, np.float32)
    _check(lib().ctl_generate_sample_tables(pass_index, _ptr(d1), _ptr(d2)))
    return d1, d2


def generate_sample_tables_n(first_pass, n):
    """ctl_generate_sample_tables_n: passes first_pass .. first_pass+n-1 produced concurrently (jump-ahead start states); returns [n, 4096*30] and [n, 4096*60]."""
    d1 = np.zeros((n, 4096 * 30), np.float32); d2 = np.zeros((n, 4096 * 30 * 2), np.float32)
    _check(lib().ctl_generate_sample_tables_n(first_pass, n, _ptr(d1), _ptr(d2)))
    return d1, d2


def traversal_bytes(counts, n_rays):
    """Algorithmic bytes of the traversal kernel: 32 B ray + 16 B result per ray,
    64 B per inner node popped, 52 B per triangle reference tested, 108 B per instance leaf entered."""
    return 48 * n_rays + 64 * counts[0] + 52 * counts[1] + 108 * counts[2]


def tile_owner(w, h, tile_w, tile_h, n_parts):
    """(h, w) int array: which part renders each pixel under ctl_render_pass_tiled's interleaved partition
    (tile index = ty * tiles_x + tx; owner = tile index % n_parts)."""
    tiles_x = (w + tile_w - 1) // tile_w
    ty, tx = np.meshgrid(np.arange(h) // tile_h, np.arange(w) // tile_w, indexing="ij")
    return (ty * tiles_x + tx) % n_parts
