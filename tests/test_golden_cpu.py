"""Pins the CPU oracle against golden vectors minted from the REFERENCE's OWN host code (oracle/_ref, see
tests/golden/make_golden.py).  Where oracle/_ref is present (this container; the prebuilt .so also travels to the GPU
box) the oracle is additionally compared with it live.  CPU only.

Tolerances: RNG / sample tables / Woop encoding bit-exact.  Traversal: indices identical, t within 4e-6 relative (the
reference host build has no FMA contraction, the oracle writes the slab/Woop FMAs of the CUDA kernel explicitly).
BSDF tables: 2e-5 relative (same formulas, libm).  Images (same seed, same pass): per-pixel rel. L2 <= 1e-3 on >= 99 % of
pixels, image rel. RMSE <= 1e-3, identical weights and ray counts (ray counts up to the documented zero-throughput stop)."""
import os

import numpy as np
import pytest

import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api, Material

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))


def test_xorwow_and_tables(orc):
    assert np.array_equal(orc.xorwow_floats(orc.xorwow_state(1234, 7539414, 0), 64).view(np.uint32), GOLD["xorwow_7539414_first64"].view(np.uint32))
    assert np.array_equal(orc.xorwow_floats(orc.xorwow_state(1234, 0, 0), 8).view(np.uint32), GOLD["xorwow_0_first8"].view(np.uint32))
    for p in (0, 1, 5):
        d1, d2 = orc.sample_tables(p)
        assert np.array_equal(d1[:8192].view(np.uint32), GOLD[f"tables_p{p}_d1_head"].view(np.uint32))
        assert np.array_equal(d2[:16384].view(np.uint32), GOLD[f"tables_p{p}_d2_head"].view(np.uint32))
        s = GOLD[f"tables_p{p}_sums"]
        assert d1.astype(np.float64).sum() == s[0] and d2.astype(np.float64).sum() == s[1] and d1[-1] == np.float32(s[2]) and d2[-1] == np.float32(s[3])


def test_product_table_generator_matches_reference(built_lib):
    for p in (0, 5):
        d1, d2 = ctl.generate_sample_tables(p)
        assert np.array_equal(d1[:8192].view(np.uint32), GOLD[f"tables_p{p}_d1_head"].view(np.uint32))
        assert d2.astype(np.float64).sum() == GOLD[f"tables_p{p}_sums"][1]


def test_concurrent_table_generator_matches_sequential_and_reference(built_lib):
    """ctl_generate_sample_tables_n (what the context runs for frames with host-generated tables): the passes of a frame produced by concurrent host threads from
    jump-ahead start states are the tables of the sequential generator, bit for bit -- and therefore the reference generator's (goldens of passes 0 and 5)."""
    d1, d2 = ctl.generate_sample_tables_n(0, 8)
    for p in (0, 5):
        assert np.array_equal(d1[p][:8192].view(np.uint32), GOLD[f"tables_p{p}_d1_head"].view(np.uint32))
        assert np.array_equal(d2[p][:16384].view(np.uint32), GOLD[f"tables_p{p}_d2_head"].view(np.uint32))
        s = GOLD[f"tables_p{p}_sums"]
        assert d1[p].astype(np.float64).sum() == s[0] and d2[p].astype(np.float64).sum() == s[1] and d1[p][-1] == np.float32(s[2]) and d2[p][-1] == np.float32(s[3])
    for first, n in ((0, 8), (3, 5), (6, 1), (2, 19)):
        e1, e2 = ctl.generate_sample_tables_n(first, n)
        for k in (0, n // 2, n - 1):
            a, b = ctl.generate_sample_tables(first + k)
            assert np.array_equal(a.view(np.uint32), e1[k].view(np.uint32)) and np.array_equal(b.view(np.uint32), e2[k].view(np.uint32))


def test_woop_encoding(orc, built_lib):
    for t, w in zip(GOLD["woop_tris"], GOLD["woop_data"]):
        assert np.array_equal(orc.encode_woop(t[0], t[1], t[2]).view(np.uint32), w.view(np.uint32))


@pytest.mark.parametrize("kind", ["cornell", "cornell7", "soup"])
def test_trace_rays(orc, kind):
    s = ctl.Scene(kind, 64, 64)
    rays = np.ascontiguousarray(GOLD[f"trace_{kind}_rays"]).view(api.RAY_DTYPE).reshape(-1)
    ref = np.ascontiguousarray(GOLD[f"trace_{kind}_res"]).view(api.TRACE_RESULT_DTYPE).reshape(-1)
    got = orc.trace_rays(s.view, rays)
    assert np.array_equal(got["tri_idx"], ref["tri_idx"]) and np.array_equal(got["node_idx"], ref["node_idx"])
    hit = ref["tri_idx"] != 0xffffffff
    assert hit.mean() > 0.5
    assert np.all(np.abs(got["dist"][hit] - ref["dist"][hit]) <= 4e-6 * np.maximum(1, ref["dist"][hit]))
    assert np.all(np.abs(got["u"][hit] - ref["u"][hit]) <= 2e-5) and np.all(np.abs(got["v"][hit] - ref["v"][hit]) <= 2e-5)
    assert np.all(got["dist"][~hit] == ref["dist"][~hit])


MATS = {
    "diffuse": dict(bsdf=0, refl=(0.5, 0.6, 0.7)), "rc_beck_0.1": dict(bsdf=1, distr=0, alpha=0.1), "rc_beck_0.3": dict(bsdf=1, distr=0, alpha=0.3),
    "rc_ggx_0.2": dict(bsdf=1, distr=1, alpha=0.2), "dielectric_1.5": dict(bsdf=2),
}


def _mat(bsdf, distr=0, refl=(1, 1, 1), alpha=0.1):
    m = Material(); m.bsdf_type = bsdf; m.flags = 0; m.node_light_index = 0xffffffff; m.distr_type = distr
    m.reflectance[:] = refl; m.alpha_u = m.alpha_v = alpha
    if bsdf == 1:
        m.eta[:] = (0.2, 0.924, 1.102); m.k[:] = (3.912, 2.452, 2.142)
    else:
        m.eta[:] = (1.5, 1.5, 1.5); m.k[:] = (0, 0, 0)
    m.transmittance = 1.0
    return m


@pytest.mark.parametrize("name", sorted(MATS))
def test_bsdf_tables(orc, name):
    m = _mat(**MATS[name]); ref = GOLD[f"bsdf_{name}"]; k = 0
    for wi in GOLD["bsdf_wi"]:
        for (sx, sy) in GOLD["bsdf_samples"]:
            o9, f3, pdf = orc.bsdf_probe(m, wi, float(sx), float(sy))
            got = np.concatenate([o9, f3, [pdf]]); r = ref[k]; k += 1
            if not np.any(r[:3]):   # failed sample: the reference leaves wo / sampledType / eta unset
                assert not np.any(got[:3])
                sel = [9, 10, 11, 12]
            else:
                assert int(got[7]) == int(r[7]), (name, wi, sx, sy)     # sampled component
                sel = list(range(13))
            assert np.allclose(got[sel], r[sel], rtol=2e-5, atol=1e-7), (name, wi, sx, sy, got, r)


@pytest.mark.parametrize("key,kind,w,h,spp", [("image_cornell_128x128_1spp", "cornell", 128, 128, 1), ("image_cornell7_96x96_8spp", "cornell7", 96, 96, 8),
                                              ("image_soup_96x96_1spp", "soup", 96, 96, 1)])
def test_images(orc, key, kind, w, h, spp):
    ref = np.ascontiguousarray(GOLD[key]).view(api.PIXEL_DTYPE).reshape(h, w)
    s = ctl.Scene(kind, w, h)
    img, rays = orc.render(s.view, w, h, n_passes=spp, max_path_length=8)
    a, b = img["rgb"], ref["rgb"]
    rel = np.linalg.norm(a - b, axis=2) / (np.linalg.norm(b, axis=2) + 1e-3)
    assert (rel <= 1e-3).mean() >= 0.99
    assert np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()) <= 1e-3
    assert np.array_equal(img["weight_sum"], ref["weight_sum"])
    ref_rays = int(GOLD[key + "_rays"][0])
    assert abs(rays - ref_rays) <= 0.01 * ref_rays   # fp-contraction flips a few discrete decisions; the oracle also stops zero-throughput paths


def test_two_light_scene(orc):
    """Hand-made scene through ctl_scene_create_from_mesh: two area lights (light-selection CDF, pdfEmitter), a two-sided
    card, a GGX block, back-facing one-sided walls (zero-throughput paths) -- vs the reference's own PathTrace."""
    from scene_fixtures import two_light_room
    s = two_light_room(96, 96)
    assert s.view.num_lights == 2 and s.view.light_cdf[0] == 0.5 and s.view.light_cdf[1] == 1.0
    ref = np.ascontiguousarray(GOLD["image_two_light_96x96_4spp"]).view(api.PIXEL_DTYPE).reshape(96, 96)
    img, rays = orc.render(s.view, 96, 96, n_passes=4, max_path_length=6)
    rel = np.linalg.norm(img["rgb"] - ref["rgb"], axis=2) / (np.linalg.norm(ref["rgb"], axis=2) + 1e-3)
    assert (rel <= 1e-3).mean() >= 0.99 and np.array_equal(img["weight_sum"], ref["weight_sum"])
    assert abs(img["rgb"].mean() - ref["rgb"].mean()) <= 1e-4 * ref["rgb"].mean()
    # the oracle stops zero-throughput paths (here: hits on the back of one-sided walls), the reference traces them on
    assert rays < int(GOLD["image_two_light_96x96_4spp_rays"][0])


def test_config1_cornell_256(orc):
    """BASELINE config 1 (Cornell-32, 256x256, 1 spp) against the reference's own CPU path."""
    s = ctl.Scene("cornell", 256, 256)
    img, rays = orc.render(s.view, 256, 256, n_passes=1, max_path_length=8)
    st = GOLD["config1_cornell_256_1spp_stats"]
    assert abs(img["rgb"].astype(np.float64).mean() - st[0]) <= 1e-5 * st[0]
    assert img["weight_sum"].sum() == st[2] and abs(rays - int(st[3])) <= 1e-3 * st[3]
    assert np.allclose(img["rgb"].astype(np.float64).mean(axis=(1, 2)), GOLD["config1_cornell_256_1spp_rowmeans"], rtol=2e-4, atol=1e-6)


def test_oracle_vs_live_reference(orc):
    """Where oracle/_ref exists: fresh inputs (not the minted ones) through both."""
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; the golden vectors above cover this box)")
    s = ctl.Scene("soup", 80, 60, seed=99, n_hint=700)
    rng = np.random.default_rng(123)
    rays = np.zeros(3000, api.RAY_DTYPE)
    lo = np.array(list(s.view.box_min)); hi = np.array(list(s.view.box_max))
    rays["o"] = rng.uniform(lo, hi, (3000, 3)); d = rng.normal(size=(3000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); rays["d"] = d; rays["tmax"] = 3e38
    a, b = rb.trace_rays(s.view, rays), orc.trace_rays(s.view, rays)
    assert (a["tri_idx"] == b["tri_idx"]).mean() >= 0.9999
    ia, ra = rb.render(s.view, 80, 60, n_passes=2, max_path_length=12, rr_start=3)
    ib, rbb = orc.render(s.view, 80, 60, n_passes=2, max_path_length=12, rr_start=3)
    rel = np.linalg.norm(ia["rgb"] - ib["rgb"], axis=2) / (np.linalg.norm(ib["rgb"], axis=2) + 1e-3)
    assert (rel <= 1e-3).mean() >= 0.99 and np.array_equal(ia["weight_sum"], ib["weight_sum"])
    assert abs(rbb - ra) <= 0.02 * ra


def test_resolve_srgb8_matches_reference(orc):
    """Default image pipeline (copySamplesToOutput): the oracle's restatement is byte-identical to the reference's own
    PixelData::toSpectrum -> toSRGB -> toRGBCOL where oracle/_ref is available; known answers otherwise."""
    px = np.zeros(6, api.PIXEL_DTYPE)
    px["rgb"] = [(0, 0, 0), (0.002, 0.5, 1.0), (4.0, 8.0, 2.0), (0.2, 0.2, 0.2), (1e-4, 5.0, -1.0), (0.18, 0.18, 0.18)]
    px["weight_sum"] = [0, 1, 8, 1, 1, 1]
    px["rgb_splat"][3] = (0.5, 0.0, 0.25)
    got = orc.resolve_srgb8(px, splat_scale=0.5)
    assert got[0].tolist() == [0, 0, 0, 255] and got[1].tolist() == [6, 187, 254, 255] and got[2].tolist() == [187, 254, 136, 255]  # 1.0 -> 254: 1.055f*1 - 0.055f < 1 in fp32
    assert got[3].tolist() == [178, 123, 154, 255] and got[4].tolist() == [0, 255, 0, 255]
    assert got[5].tolist() == [117, 117, 117, 255]      # 18 % grey -> sRGB 0.4614 -> 117
    import ref_binding as rb
    if rb.available():
        assert np.array_equal(got, rb.resolve_srgb8(px, splat_scale=0.5))
        s = ctl.Scene("cornell", 64, 64)
        img, _ = orc.render(s.view, 64, 64, n_passes=3, max_path_length=6)
        assert np.array_equal(orc.resolve_srgb8(img), rb.resolve_srgb8(img))


def test_image_pipeline_goldens(orc):
    """Default resolve and CanonicalFilter (box / Gaussian / triangle) + RGBE stage: oracle == bytes produced by the
    reference's own evalFilter / toRGBE / fromRGBE / toSRGB / toRGBCOL (minted from oracle/_ref)."""
    acc = np.ascontiguousarray(GOLD["pipeline_accum_cornell_80x64_4spp"]).view(api.PIXEL_DTYPE).reshape(64, 80)
    assert np.array_equal(orc.resolve_srgb8(acc), GOLD["pipeline_resolve_default"])
    for k, (t, xw, yw, a) in enumerate(GOLD["pipeline_filter_cases"]):
        assert np.array_equal(orc.resolve_filtered_srgb8(acc, int(t), float(xw), float(yw), float(a)), GOLD[f"pipeline_resolve_filter{k}"])


# ---- host-arithmetic variant: every line of the restatement pinned bit-for-bit ------------------------------------------------
PT_GOLD_CASES = [("image_cornell_128x128_1spp", "cornell", 128, 128, 1, 8), ("image_cornell7_96x96_8spp", "cornell7", 96, 96, 8, 8),
                 ("image_soup_96x96_1spp", "soup", 96, 96, 1, 8), ("image_two_light_96x96_4spp", "two_light", 96, 96, 4, 6)]


def _scene(kind, w, h):
    if kind == "two_light":
        from scene_fixtures import two_light_room
        return two_light_room(w, h)
    return ctl.Scene(kind, w, h)


@pytest.mark.parametrize("key,kind,w,h,spp,mpl", PT_GOLD_CASES)
def test_pathtracer_host_arithmetic_is_bit_identical_to_reference(orc, key, kind, w, h, spp, mpl):
    """oracle.cpp built with -DORC_NO_FMA (a*b+c in two roundings = the reference's g++ host build) reproduces the reference's own
    PathTrace images BIT FOR BIT; the default build differs from that variant only by the explicit FMAs of the slab / Woop tests."""
    ref = np.ascontiguousarray(GOLD[key]).view(api.PIXEL_DTYPE).reshape(h, w)
    s = _scene(kind, w, h)   # keep the scene object alive: the view holds pointers into it
    with orc.host_arithmetic():
        img, rays = orc.render(s.view, w, h, n_passes=spp, max_path_length=mpl)
    assert np.array_equal(img["rgb"].view(np.uint32), ref["rgb"].view(np.uint32))
    assert np.array_equal(img["weight_sum"], ref["weight_sum"])
    assert rays <= int(GOLD[key + "_rays"][0])   # only the zero-throughput early stop separates the counts


def _wpt_cases():
    kinds = ["cornell7", "soup", "two_light", "cornell"]
    return [(k, kinds[k]) + tuple(int(v) for v in GOLD["wpt_cases"][k]) for k in range(len(kinds))]


@pytest.mark.parametrize("k,kind,w,h,spp,mpl,rr,direct", _wpt_cases())
def test_wavefront_restatement_vs_reference(orc, k, kind, w, h, spp, mpl, rr, direct):
    """WavefrontPathTracer: goldens are the reference's own pathIterateKernel + DoubleRayBuffer run in the serial
    queue order.  Host-arithmetic oracle: image, ray count and per-iteration queue sizes bit-identical.  Default (FMA) oracle:
    same queue evolution on these cases, pixels within 1e-3 (an FMA-induced flip of one discrete decision would shift every
    later queue slot and with it the slot-keyed random numbers, so this is checked on small fixed cases)."""
    ref = np.ascontiguousarray(GOLD[f"wpt_image_{k}_{kind}"]).view(api.PIXEL_DTYPE).reshape(h, w)
    s = _scene(kind, w, h)
    with orc.host_arithmetic():
        img, rays, q = orc.render_wavefront(s.view, w, h, n_passes=spp, max_path_length=mpl, rr_start=rr, direct=direct)
    assert np.array_equal(img["rgb"].view(np.uint32), ref["rgb"].view(np.uint32)) and np.array_equal(img["weight_sum"], ref["weight_sum"])
    assert rays == int(GOLD[f"wpt_rays_{k}_{kind}"][0]) and np.array_equal(q, GOLD[f"wpt_queues_{k}_{kind}"])
    assert (ref["weight_sum"] == spp).all()   # exactly one sample per pixel and pass, splatted at the un-jittered pixel
    img, rays, q = orc.render_wavefront(s.view, w, h, n_passes=spp, max_path_length=mpl, rr_start=rr, direct=direct)
    rel = np.linalg.norm(img["rgb"] - ref["rgb"], axis=2) / (np.linalg.norm(ref["rgb"], axis=2) + 1e-3)
    assert (rel <= 1e-3).mean() >= 0.99, (rel <= 1e-3).mean()
    assert abs(rays - int(GOLD[f"wpt_rays_{k}_{kind}"][0])) <= 0.01 * rays


def test_wavefront_live_reference_fresh_inputs(orc):
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; the golden vectors above cover this box)")
    s = ctl.Scene("soup", 72, 40, seed=5, n_hint=500)
    a, ra, qa = rb.render_wavefront(s.view, 72, 40, n_passes=2, pass_first=3, max_path_length=10, rr_start=2)
    with orc.host_arithmetic():
        b, rbb, qb = orc.render_wavefront(s.view, 72, 40, n_passes=2, pass_first=3, max_path_length=10, rr_start=2)
    assert np.array_equal(a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)) and ra == rbb and np.array_equal(qa, qb)


def test_wavefront_converges_to_pathtracer(orc):
    """Both integrators estimate the same integral: 64-pass means agree within Monte-Carlo noise."""
    s = ctl.Scene("cornell", 32, 32)
    a, _, _ = orc.render_wavefront(s.view, 32, 32, n_passes=64, max_path_length=8)
    b, _ = orc.render(s.view, 32, 32, n_passes=64, max_path_length=8)
    ma, mb = a["rgb"].astype(np.float64).mean(axis=(0, 1)) / 64, b["rgb"].astype(np.float64).mean(axis=(0, 1)) / 64
    assert np.allclose(ma, mb, rtol=0.03), (ma, mb)


# ---- full image pipeline + variance buffer ----------------------------------------------------------------------
def test_full_image_pipeline_goldens(orc):
    """applyImagePipeline with Mitchell / Lanczos-sinc filters, ToneMapPostProcess (Reinhard05) and both: the restatement's bytes and
    luminance statistics equal those of the reference's own evalFilter / toRGBE / getLuminance / toYxy / fromYxy / toRGBCOL / toSRGB."""
    from cudatracerlib_b200 import ImagePipeline
    acc = np.ascontiguousarray(GOLD["pipeline_accum_cornell_80x64_4spp"]).view(api.PIXEL_DTYPE).reshape(64, 80)
    for k, (ft, xw, yw, p0, p1, tm, key, burn) in enumerate(GOLD["pipeline_full_cases"]):
        rgba, lum = orc.apply_image_pipeline(acc, ImagePipeline(int(ft), float(xw), float(yw), float(p0), float(p1), int(tm), float(key), float(burn)))
        assert np.array_equal(rgba, GOLD[f"pipeline_full_{k}"]), k
        assert np.array_equal(lum.view(np.uint32), GOLD[f"pipeline_full_{k}_lum"].view(np.uint32)), k
    # the tone mapper must actually change the image, and the burn parameter must matter
    assert not np.array_equal(GOLD["pipeline_full_3"], GOLD["pipeline_resolve_default"]) and not np.array_equal(GOLD["pipeline_full_3"], GOLD["pipeline_full_4"])


def test_pixel_variance_buffer_golden(orc):
    """PixelVarianceBuffer::AddPass after each of 4 passes: the restatement of updateMoments is bit-identical to the reference's own."""
    accs = GOLD["variance_accums_cornell_32x24"]
    var = np.zeros(32 * 24, ctl.VARIANCE_DTYPE)
    for a in accs:
        orc.variance_add_pass(var, np.ascontiguousarray(a).view(api.PIXEL_DTYPE).reshape(24, 32))
    assert np.array_equal(var.view(np.uint32).reshape(-1, 11), GOLD["variance_info_cornell_32x24"])
    assert (var["iterations_done"] == 4).all() and (var["num_samples_var"] == 4).all() and (var["weight"] == 4).all()
    assert np.allclose(var["sum_x"], (accs[-1][..., :3].reshape(-1, 3) * [0.212671, 0.715160, 0.072169]).sum(axis=1), rtol=1e-4, atol=1e-6)  # telescoping sum of the per-pass luminances


NLM_GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nlm_golden.npz"))
NLM_CASES = {"a": (0.45, 1.0), "b": (1.0, 5.0), "wide": (0.45, 1.0), "default": (0.45, 0.005)}


def _nlm_inputs(name, suffix=""):
    img = np.ascontiguousarray(NLM_GOLD[name + "_img" + suffix]).view(api.PIXEL_DTYPE).reshape(NLM_GOLD[name + "_img" + suffix].shape[:2])
    return img, np.ascontiguousarray(NLM_GOLD[name + "_var" + suffix]).view(ctl.VARIANCE_DTYPE).reshape(-1)


@pytest.mark.parametrize("name", sorted(NLM_CASES))
def test_non_local_means_filter_golden(orc, name):
    """NonLocalMeansFilter restatement against the reference's OWN kernels (computeWeights / applyWeights run on the host, tests/golden/make_nlm_golden.py):
    RGBE stage and all 169 weights per pixel bit-identical.  `wide` crosses the 200-pixel super-block boundary of the reference's launch loop;
    `default` (sigma2Scale 0.005 after 2 passes) is the regime where only the self-weight survives and the filter is the identity on the RGBE stage."""
    import hashlib
    k, s2 = NLM_CASES[name]
    img, var = _nlm_inputs(name)
    rgbe, wts = orc.nlm_filter(img, var, k, s2)
    assert np.array_equal(rgbe, NLM_GOLD[name + "_rgbe"])
    assert hashlib.sha256(wts.tobytes()).digest() == NLM_GOLD[name + "_weights_sha256"].tobytes()
    frac = float((wts > 0).mean())
    unfiltered = orc.nlm_filter(img, var, 0.45, 0.0)[0]                      # zero variance scale: distances explode, only w(p, p) = 1 is left
    if name == "default":
        assert frac == pytest.approx(1 / 169, abs=1e-9) and np.array_equal(rgbe, unfiltered)
    else:
        lit = unfiltered[..., :3].any(axis=2)                                 # `wide` is a 204x10 strip, mostly background
        assert 0.2 < frac < 0.95 and (rgbe != unfiltered).any(axis=2)[lit].mean() > (0.5 if name == "wide" else 0.9)   # 3 passes only in `wide`: many zero variances
    assert ((wts == 0) | (wts >= 0.05)).all() and wts.max() <= 1.0 and (wts.reshape(-1, 13, 13)[:, 6, 6] == 1.0).all()


def test_non_local_means_stale_weights_golden(orc):
    """UpdateWeightPeriodicity > 1: weights computed on an earlier frame are applied to the current one (NonLocalMeansFilter.cu:207-224)."""
    img0, var0 = _nlm_inputs("a", "_early")
    img, var = _nlm_inputs("a")
    _, w0 = orc.nlm_filter(img0, var0, 0.45, 1.0)
    stale = orc.nlm_filter(img, var, 0.45, 1.0, weights=w0)[0]
    assert np.array_equal(stale, NLM_GOLD["a_rgbe_stale"]) and not np.array_equal(stale, NLM_GOLD["a_rgbe"])


def test_non_local_means_live_reference(orc):
    """Fresh inputs (not in the goldens) through oracle/_ref when it is built."""
    import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    w, h = 37, 23
    img = np.zeros((h, w), api.PIXEL_DTYPE); var = np.zeros(w * h, ctl.VARIANCE_DTYPE)
    base = rng.uniform(0, 2, (h // 14 + 1, w // 14 + 1, 3)).repeat(14, 0).repeat(14, 1)[:h, :w]   # flat regions + noise: weights in every regime
    n = 5
    samples = base[None] + rng.normal(0, 0.15, (n, h, w, 3))
    img["rgb"] = samples.sum(0).astype(np.float32); img["weight_sum"] = n
    lum = (samples * [0.212671, 0.715160, 0.072169]).sum(-1).reshape(n, -1)
    var["sum_x"] = lum.sum(0); var["sum_x2"] = (lum ** 2).sum(0); var["num_samples_var"] = n; var["iterations_done"] = n
    for k, s2 in ((0.45, 1.0), (0.7, 0.3)):
        a, wa = orc.nlm_filter(img, var, k, s2); b, wb = rb.nlm_filter(img, var, k, s2)
        assert np.array_equal(a, b) and np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
        assert 0.05 < (wa > 0).mean() < 0.99
