"""ctypes binding of oracle/_ref/libctl_ref.so -- the REFERENCE's own host code (built by oracle/build_ref.sh where
/root/reference is mounted; the prebuilt .so travels to the GPU box).  TEST INFRASTRUCTURE, same rules as oracle/."""
import ctypes as C
import os

import numpy as np

from cudatracerlib_b200.api import RAY_DTYPE, TRACE_RESULT_DTYPE, PIXEL_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libctl_ref.so")
_lib = None


def available():
    return os.path.exists(REF_LIB)


def ref():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB)
        L.ref_render.restype = C.c_ulonglong
        L.ref_render.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_void_p, C.c_int]
        L.ref_trace_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_xorwow_floats.argtypes = [C.c_uint, C.c_int, C.c_void_p]
        L.ref_woop_setdata.argtypes = [C.c_void_p] * 4
        L.ref_sample_tables.argtypes = [C.c_uint, C.c_void_p, C.c_void_p]
        L.ref_bsdf_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def render(view, w, h, n_passes=1, pass_first=0, max_path_length=8, rr_start=5, direct=1, window=None, n_threads=0, img=None):
    if img is None:
        img = np.zeros((h, w), PIXEL_DTYPE)
    x0, y0, x1, y1 = window if window else (0, 0, w, h)
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    rays = ref().ref_render(C.byref(view), w, h, x0, y0, x1, y1, pass_first, n_passes, max_path_length, rr_start, direct, _p(img), n_threads)
    return img, int(rays)


def trace_rays(view, rays):
    rays = np.ascontiguousarray(rays, RAY_DTYPE); out = np.zeros(len(rays), TRACE_RESULT_DTYPE)
    ref().ref_trace_rays(C.byref(view), len(rays), _p(rays), _p(out)); return out


def xorwow_floats(subsequence, n):
    out = np.zeros(n, np.float32); ref().ref_xorwow_floats(subsequence, n, _p(out)); return out


def woop_setdata(v0, v1, v2):
    a, b, c = (np.ascontiguousarray(x, np.float32) for x in (v0, v1, v2)); out = np.zeros(12, np.float32)
    ref().ref_woop_setdata(_p(a), _p(b), _p(c), _p(out)); return out


def sample_tables(pass_index):
    d1 = np.zeros(4096 * 30, np.float32); d2 = np.zeros(4096 * 30 * 2, np.float32)
    ref().ref_sample_tables(pass_index, _p(d1), _p(d2)); return d1, d2


def resolve_srgb8(img, splat_scale=0.0):
    img = np.ascontiguousarray(img); out = np.zeros(img.shape + (4,), np.uint8)
    ref().ref_resolve_srgb8.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    ref().ref_resolve_srgb8(_p(img), img.size, splat_scale, _p(out)); return out


def bsdf_probe(mat, wi, sx, sy):
    w = np.ascontiguousarray(wi, np.float32); out = np.zeros(9, np.float32); f = np.zeros(3, np.float32); pdf = np.zeros(1, np.float32)
    ref().ref_bsdf_probe(C.byref(mat), _p(w), sx, sy, _p(out), _p(f), _p(pdf)); return out, f, pdf[0]


def resolve_filtered_srgb8(img, filter_type=0, xw=0.5, yw=0.5, alpha=2.0, splat_scale=0.0):
    img = np.ascontiguousarray(img); h, w = img.shape; out = np.zeros((h, w, 4), np.uint8)
    ref().ref_resolve_filtered_srgb8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    ref().ref_resolve_filtered_srgb8(_p(img), w, h, splat_scale, filter_type, xw, yw, alpha, _p(out)); return out


def render_wavefront(view, w, h, n_passes=1, pass_first=0, max_path_length=8, rr_start=5, direct=1, n_threads=0, img=None):
    """The reference's WavefrontPathTracer (pathIterateKernel + DoubleRayBuffer, serial queue order).  Returns (image, rays, queue sizes
    [max_path_length, 2] of the last pass: primary / secondary rays intersected per iteration)."""
    if img is None:
        img = np.zeros((h, w), PIXEL_DTYPE)
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    q = np.zeros((max_path_length, 2), np.uint32)
    f = ref().ref_render_wavefront
    f.restype = C.c_ulonglong
    f.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_int, C.c_void_p]
    rays = f(C.byref(view), w, h, pass_first, n_passes, max_path_length, rr_start, direct, _p(img), n_threads, _p(q))
    return img, int(rays), q


def apply_image_pipeline(img, pipeline, splat_scale=0.0):
    img = np.ascontiguousarray(img); h, w = img.shape; out = np.zeros((h, w, 4), np.uint8); lum = np.zeros(6, np.float32)
    f = ref().ref_apply_image_pipeline; f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(img), w, h, splat_scale, C.byref(pipeline), _p(out), _p(lum)); return out, lum


def variance_add_pass(var, img, splat_scale=0.0):
    f = ref().ref_variance_add_pass; f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
    f(_p(var), _p(np.ascontiguousarray(img)), img.size, splat_scale); return var


def nlm_filter(img, var, k=0.45, sigma2_scale=0.005, splat_scale=0.0, weights=None):
    """NonLocalMeansFilter::Apply with the reference's own kernels run on the host.  weights=None: computed (and returned); else applied as given.
    Returns (RGBE stage (h, w, 4) u8, weights (h*w, 169) f32)."""
    img = np.ascontiguousarray(img); h, w = img.shape; out = np.zeros((h, w, 4), np.uint8)
    compute = weights is None
    wts = np.zeros((h * w, 169), np.float32) if compute else np.ascontiguousarray(weights, np.float32)
    f = ref().ref_nlm_filter; f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]; f.restype = None
    f(_p(img), _p(np.ascontiguousarray(var)), w, h, splat_scale, k, sigma2_scale, _p(wts), int(compute), _p(out)); return out, wts


def write_xmsh(path, verts, indices, sub_tris, mats, emissive=None):
    """The reference's own .xmsh writer (Mesh::CompileMesh): verts (nv, 3) f32, indices (3 * nt) u32 with the triangles of sub-mesh k
    consecutive, sub_tris[k] triangles each, mats = list of Material (one per sub-mesh), emissive (n_sub, 3) or None."""
    from cudatracerlib_b200 import Material
    v = np.ascontiguousarray(verts, np.float32); i = np.ascontiguousarray(indices, np.uint32); st = np.ascontiguousarray(sub_tris, np.uint32)
    arr = (Material * len(mats))(*mats)
    e = np.ascontiguousarray(emissive, np.float32) if emissive is not None else None
    f = ref().ref_write_xmsh; f.argtypes = [C.c_char_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    rc = f(str(path).encode(), _p(v), len(v), _p(i), len(i), _p(st), len(st), C.cast(arr, C.c_void_p), _p(e) if e is not None else None)
    if rc:
        raise RuntimeError("ref_write_xmsh failed")


def compile_mesh(in_path, xmsh_path):
    """OBJ / PLY -> .xmsh with the reference's own compilers (compileobj / compileply)."""
    f = ref().ref_compile_mesh; f.argtypes = [C.c_char_p, C.c_char_p]
    if f(os.fsencode(str(in_path)), os.fsencode(str(xmsh_path))):
        raise RuntimeError("ref_compile_mesh failed")
