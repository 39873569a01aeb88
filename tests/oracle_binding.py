"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

from cudatracerlib_b200.api import SceneView, Material, RAY_DTYPE, RESULT16_DTYPE, TRACE_RESULT_DTYPE, PIXEL_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
ORACLE_LIB_NOFMA = os.path.join(ORACLE_DIR, "liboracle_nofma.so")   # host arithmetic (no FMA): the variant pinned bit-for-bit against oracle/_ref
_libs = {}
_variant = "fma"


class host_arithmetic:
    """with host_arithmetic(): every oracle call runs the -DORC_NO_FMA build (the reference's host-build arithmetic)."""
    def __enter__(self):
        global _variant
        self.prev = _variant; _variant = "nofma"
    def __exit__(self, *a):
        global _variant
        _variant = self.prev


def build_oracle():
    src = os.path.join(ORACLE_DIR, "oracle.cpp")
    if any(not os.path.exists(l) or os.path.getmtime(src) > os.path.getmtime(l) for l in (ORACLE_LIB, ORACLE_LIB_NOFMA)):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return ORACLE_LIB


def oracle():
    if _variant not in _libs:
        build_oracle()
        L = C.CDLL(ORACLE_LIB if _variant == "fma" else ORACLE_LIB_NOFMA)
        L.orc_half_to_float.restype = C.c_float; L.orc_half_to_float.argtypes = [C.c_uint16]
        L.orc_float_to_half.restype = C.c_uint16; L.orc_float_to_half.argtypes = [C.c_float]
        L.orc_encode_normal.restype = C.c_uint16
        L.orc_fresnel_dielectric_ext.restype = C.c_float; L.orc_fresnel_dielectric_ext.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.orc_warp.argtypes = [C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.orc_microfacet_sample.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.orc_bsdf_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bsdf_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fresnel_conductor_exact.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_light_sample_direct.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.orc_fill_dg.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_camera_ray.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_woop_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_trace_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_intersect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_render.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_path_probe.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]
        L.orc_sample_tables.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_sampler_draws.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_xorwow_init.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_void_p]
        L.orc_xorwow_floats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_render_wavefront.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p]
        _libs[_variant] = L
    return _libs[_variant]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def xorwow_state(seed, subsequence, offset=0):
    st = np.zeros(6, np.uint32); oracle().orc_xorwow_init(seed, subsequence, offset, _p(st)); return st


def xorwow_floats(state, n):
    out = np.zeros(n, np.float32); oracle().orc_xorwow_floats(_p(state), n, _p(out)); return out


def sample_tables(pass_index):
    d1 = np.zeros(4096 * 30, np.float32); d2 = np.zeros(4096 * 30 * 2, np.float32)
    oracle().orc_sample_tables(pass_index, _p(d1), _p(d2)); return d1, d2


def sampler_draws(d1, d2, idx, n1, n2):
    o1 = np.zeros(max(n1, 1), np.float32); o2 = np.zeros(max(n2, 1) * 2, np.float32)
    oracle().orc_sampler_draws(_p(d1), _p(d2), idx, n1, _p(o1), n2, _p(o2)); return o1[:n1], o2[:2 * n2].reshape(-1, 2)


def encode_woop(v0, v1, v2):
    a, b, c = (np.ascontiguousarray(x, np.float32) for x in (v0, v1, v2))
    out = np.zeros(12, np.float32); oracle().orc_encode_woop(_p(a), _p(b), _p(c), _p(out)); return out


def woop_intersect(woop12, o, d, tmax=3.4e38):
    w = np.ascontiguousarray(woop12, np.float32); o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
    tuv = np.zeros(3, np.float32)
    hit = oracle().orc_woop_intersect(_p(w), _p(o), _p(d), tmax, _p(tuv))
    return bool(hit), tuv


def encode_tri_data(p9, n9, uv6, mat):
    p, n, uv = (np.ascontiguousarray(x, np.float32).ravel() for x in (p9, n9, uv6))
    out = np.zeros(8, np.uint32); oracle().orc_encode_tri_data(_p(p), _p(n), _p(uv), mat, _p(out)); return out


def decode_normal(code):
    out = np.zeros(3, np.float32); oracle().orc_decode_normal(C.c_uint16(code), _p(out)); return out


def encode_normal(n):
    n = np.ascontiguousarray(n, np.float32); return oracle().orc_encode_normal(_p(n))


def warp(which, sx, sy):
    out = np.zeros(3, np.float32); oracle().orc_warp(which, sx, sy, _p(out)); return out


def fresnel_dielectric_ext(cosi, eta):
    ct = C.c_float(0); F = oracle().orc_fresnel_dielectric_ext(cosi, eta, C.byref(ct)); return F, ct.value


def fresnel_conductor_exact(cosi, eta, k):
    e = np.ascontiguousarray(eta, np.float32); kk = np.ascontiguousarray(k, np.float32); out = np.zeros(3, np.float32)
    oracle().orc_fresnel_conductor_exact(cosi, _p(e), _p(kk), _p(out)); return out


def microfacet_sample(type_, alpha, wi, sx, sy):
    w = np.ascontiguousarray(wi, np.float32); out = np.zeros(6, np.float32)
    oracle().orc_microfacet_sample(type_, alpha, _p(w), sx, sy, _p(out)); return out


def bsdf_probe(mat, wi, sx, sy):
    w = np.ascontiguousarray(wi, np.float32); out = np.zeros(9, np.float32); f = np.zeros(3, np.float32); pdf = np.zeros(1, np.float32)
    oracle().orc_bsdf_probe(C.byref(mat), _p(w), sx, sy, _p(out), _p(f), _p(pdf)); return out, f, pdf[0]


def light_sample_direct(view, light, ref, refN, sx, sy):
    r = np.ascontiguousarray(ref, np.float32); n = np.ascontiguousarray(refN, np.float32); out = np.zeros(11, np.float32)
    oracle().orc_light_sample_direct(C.byref(view), light, _p(r), _p(n), sx, sy, _p(out)); return out


def fill_dg(view, u, v, tri, node):
    out = np.zeros(12, np.float32); oracle().orc_fill_dg(C.byref(view), u, v, tri, node, _p(out)); return out.reshape(4, 3)


def camera_ray(view, px, py):
    o = np.zeros(3, np.float32); d = np.zeros(3, np.float32); oracle().orc_camera_ray(C.byref(view), px, py, _p(o), _p(d)); return o, d


def _alias(view, node_idx, miss):
    """Re-braided views (ctl_scene_set_rebraid; the default for large multi-instance scenes): the oracle walks the view and reports the pseudo-node it
    hit; the API names the instance that pseudo-node stands for (view.node_alias), as the reference's own scene would."""
    if not view.node_alias:
        return
    alias = np.ctypeslib.as_array(view.node_alias, shape=(view.n_nodes,))
    hit = node_idx != miss
    node_idx[hit] = alias[node_idx[hit]].astype(node_idx.dtype)


def trace_rays(view, rays, counts=False, pseudo_nodes=False):
    rays = np.ascontiguousarray(rays, RAY_DTYPE); out = np.zeros(len(rays), TRACE_RESULT_DTYPE); cnt = np.zeros(3, np.uint64)
    oracle().orc_trace_rays(C.byref(view), len(rays), _p(rays), _p(out), _p(cnt) if counts else None)
    if not pseudo_nodes:
        _alias(view, out["node_idx"], 0xffffffff)
    return (out, [int(x) for x in cnt]) if counts else out


def intersect(view, rays, any_hit=False, pseudo_nodes=False):
    rays = np.ascontiguousarray(rays, RAY_DTYPE); out = np.zeros(len(rays), RESULT16_DTYPE)
    oracle().orc_intersect(C.byref(view), len(rays), _p(rays), _p(out), int(any_hit))
    if not pseudo_nodes:
        _alias(view, out["node_idx"], -1)
    return out


def set_stop_zero_throughput(on):
    """orc_set_stop_zero_throughput: 1 = the product default (a path of exactly zero throughput ends), 0 = the reference's behaviour (it lives until
    Russian roulette; identical images, more rays).  Applies to both oracle builds when loaded."""
    oracle().orc_set_stop_zero_throughput(int(on))


def render(view, w, h, n_passes=1, pass_first=0, max_path_length=8, rr_start=5, direct=1, window=None, n_threads=0, counts=False, img=None):
    if img is None:
        img = np.zeros((h, w), PIXEL_DTYPE)
    x0, y0, x1, y1 = window if window else (0, 0, w, h)
    rays = C.c_uint64(0); cnt = np.zeros(3, np.uint64)
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    oracle().orc_render(C.byref(view), w, h, x0, y0, x1, y1, pass_first, n_passes, max_path_length, rr_start, direct, _p(img), C.byref(rays), n_threads,
                        _p(cnt) if counts else None)
    return (img, rays.value, [int(x) for x in cnt]) if counts else (img, rays.value)


def resolve_srgb8(img, splat_scale=0.0):
    img = np.ascontiguousarray(img); out = np.zeros(img.shape + (4,), np.uint8)
    oracle().orc_resolve_srgb8.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    oracle().orc_resolve_srgb8(_p(img), img.size, splat_scale, _p(out)); return out


def resolve_filtered_srgb8(img, filter_type=0, xw=0.5, yw=0.5, alpha=2.0, splat_scale=0.0):
    img = np.ascontiguousarray(img); h, w = img.shape; out = np.zeros((h, w, 4), np.uint8)
    oracle().orc_resolve_filtered_srgb8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    oracle().orc_resolve_filtered_srgb8(_p(img), w, h, splat_scale, filter_type, xw, yw, alpha, _p(out)); return out


def path_probe(view, w, x, y, pass_index=0, max_path_length=8, rr_start=5, direct=1):
    rgb = np.zeros(3, np.float32); rays = C.c_uint64(0)
    oracle().orc_path_probe(C.byref(view), w, x, y, pass_index, max_path_length, rr_start, direct, _p(rgb), C.byref(rays)); return rgb, rays.value


def render_wavefront(view, w, h, n_passes=1, pass_first=0, max_path_length=8, rr_start=5, direct=1, img=None, counts=False):
    """WavefrontPathTracer restatement (serial queue order).  Returns (image, rays, queue sizes [max_path_length, 2] of the last pass
    [, visit counts: primary inner / tris / inst / rays, secondary (any-hit form) inner / tris / inst / rays])."""
    if img is None:
        img = np.zeros((h, w), PIXEL_DTYPE)
    rays = np.zeros(1, np.uint64); q = np.zeros((max_path_length, 2), np.uint32); cnt = np.zeros(8, np.uint64)
    f = oracle().orc_render_wavefront_counted
    f.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(C.byref(view), w, h, pass_first, n_passes, max_path_length, rr_start, direct, _p(img), _p(rays), _p(q), _p(cnt) if counts else None)
    return (img, int(rays[0]), q, [int(x) for x in cnt]) if counts else (img, int(rays[0]), q)


def apply_image_pipeline(img, pipeline, splat_scale=0.0):
    """applyImagePipeline restatement; returns (rgba8 image, lum info[6])."""
    img = np.ascontiguousarray(img); h, w = img.shape; out = np.zeros((h, w, 4), np.uint8); lum = np.zeros(6, np.float32)
    f = oracle().orc_apply_image_pipeline; f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(img), w, h, splat_scale, C.byref(pipeline), _p(out), _p(lum)); return out, lum


def nlm_filter(img, var, k=0.45, sigma2_scale=0.005, splat_scale=0.0, weights=None):
    """NonLocalMeansFilter restatement; same contract as ref_binding.nlm_filter."""
    img = np.ascontiguousarray(img); h, w = img.shape; out = np.zeros((h, w, 4), np.uint8)
    compute = weights is None
    wts = np.zeros((h * w, 169), np.float32) if compute else np.ascontiguousarray(weights, np.float32)
    f = oracle().orc_nlm_filter; f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]; f.restype = None
    f(_p(img), _p(np.ascontiguousarray(var)), w, h, splat_scale, k, sigma2_scale, _p(wts), int(compute), _p(out)); return out, wts


def pipeline_from_stage2(rgbe, pipeline):
    """Tail of applyImagePipeline after a filter wrote the RGBE stage: (rgba8, lum info[6])."""
    rgbe = np.ascontiguousarray(rgbe, np.uint8); h, w = rgbe.shape[:2]; out = np.zeros((h, w, 4), np.uint8); lum = np.zeros(6, np.float32)
    f = oracle().orc_pipeline_from_stage2; f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]; f.restype = None
    f(_p(rgbe), w, h, C.byref(pipeline), _p(out), _p(lum)); return out, lum


def variance_add_pass(var, img, splat_scale=0.0):
    f = oracle().orc_variance_add_pass; f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
    f(_p(var), _p(np.ascontiguousarray(img)), img.size, splat_scale); return var
