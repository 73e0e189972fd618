"""Analytic pins of the radiance loop — evidence that does not depend on reading the reference the same way twice.

The reference cannot be run here and ships no golden image, so besides the chi-square BSDF tests
(test_bsdf_chi2.py) the estimator as a whole is checked against answers known in closed form:

  * white furnace: inside a closed cube whose walls all emit E and reflect Lambert rho, every path of at most D bounces
    carries exactly E * sum_{k=0..D} rho^k.  With NEE and Russian roulette off each PIXEL must show that value (this
    pins Lambert f / pdf = rho, the emitter term, one-sided emission, the depth cap, film normalisation); with NEE on
    (light sampling + MIS weights, pt.rs:230-258,297-323) and with roulette on (pt.rs:843-850) the estimator must stay
    unbiased: the image mean has to agree within 4 standard errors;
  * NEE on vs NEE off on cbox: two different estimators of the same integral (emitter hits only vs light sampling + MIS)
    must have equal means.
The CPU suite runs them through the oracle (and the furnace through tests/hostsim, bit for bit); the GPU suite runs the
same checks through the CUDA path.
"""
import ctypes as C
import os

import numpy as np
import pytest

import scene_variants as sv
from conftest import measured

HERE = os.path.dirname(os.path.abspath(__file__))
RHO, EMIT, DEPTH = 0.5, 1.0, 6
FURNACE = EMIT * sum(RHO ** k for k in range(DEPTH + 1))  # 1.984375, exactly representable


def _furnace_task(akr, spp, use_nee, rr_depth):
    t = akr.RenderTask.from_file(os.path.join(os.path.dirname(HERE), "scenes", "cbox", "pt.json"))
    t.pt.spp, t.pt.spp_per_pass, t.pt.max_depth, t.pt.use_nee, t.pt.rr_depth = spp, spp, DEPTH, use_nee, rr_depth
    return t


def _mean_check(rgb, expected, what, sigmas=4.0):
    """|image mean - expected| <= sigmas * standard error, the standard error taken from the spread of the pixel values
    (pixels use different blue-noise offsets and permutations: independent estimates of the same number)."""
    px = rgb.reshape(-1, 3).astype(np.float64)
    mean, se = px.mean(axis=0), px.std(axis=0, ddof=1) / np.sqrt(len(px))
    z = np.abs(mean - expected) / np.maximum(se, 1e-12)
    measured(f"{what}: mean {mean.round(5)} vs {np.round(expected, 5)}, z = {z.round(2)} (<= {sigmas})")
    assert (z <= sigmas).all(), (mean, expected, se)


# ---- CPU: oracle + hostsim -------------------------------------------------------------------------------------------
def test_white_furnace_oracle_exact_per_pixel(akr, oracle, tables, tmp_path):
    w = h = 32
    scene = akr.load_scene(sv.write_furnace(tmp_path, RHO, EMIT)).set_resolution(w, h)
    pmj, bn = tables
    task = _furnace_task(akr, 16, use_nee=0, rr_depth=100)
    film, st, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn)
    rgb = oracle.resolve(film, w * h)
    err = float(np.abs(rgb - FURNACE).max())
    measured(f"white furnace (oracle, no NEE, no RR) 32x32@16: max |pixel - {FURNACE}| = {err:.2e} (<= 1e-5)")
    assert err <= 1e-5
    assert st.segments == w * h * 16 * (DEPTH + 1) and st.shadow_rays == 0 and st.n_lights == 1  # no path ever leaves or dies early


@pytest.mark.parametrize("use_nee,rr_depth", [(1, 100), (0, 1), (1, 1)], ids=["nee", "roulette", "nee+roulette"])
def test_white_furnace_oracle_unbiased(akr, oracle, tables, tmp_path, use_nee, rr_depth):
    w = h = 32
    scene = akr.load_scene(sv.write_furnace(tmp_path, RHO, EMIT)).set_resolution(w, h)
    pmj, bn = tables
    task = _furnace_task(akr, 64, use_nee, rr_depth)
    film, _, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn)
    _mean_check(oracle.resolve(film, w * h), FURNACE, f"white furnace (oracle, use_nee={use_nee}, rr_depth={rr_depth}) 32x32@64")


@pytest.mark.parametrize("fused", [0, 1], ids=["queued", "fused"])
def test_white_furnace_hostsim_bitwise(akr, oracle, tables, tmp_path, fused):
    """The kernels' per-thread bodies (both pipelines) on the all-emissive scene: every triangle is a light, every hit
    takes the emitter-MIS branch, the light alias table has 12 entries."""
    from test_hostsim_parity import run_hostsim
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    lib.hostsim_set_pipeline(fused)
    try:
        w = h = 24
        scene = akr.load_scene(sv.write_furnace(tmp_path, RHO, EMIT)).set_resolution(w, h)
        pmj, bn = tables
        for use_nee, rr in ((0, 100), (1, 2)):
            task = _furnace_task(akr, 8, use_nee, rr)
            ofilm, ost, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn)
            film, _, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
            assert np.array_equal(film, ofilm)
            assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
    finally:
        lib.hostsim_set_pipeline(0)


@pytest.mark.parametrize("fused", [0, 1], ids=["queued", "fused"])
def test_textured_white_furnace_pins_the_texture_driven_path(akr, oracle, tables, tmp_path, fused):
    """The furnace again, with the albedo read from a constant-valued image texture: the material is texture-driven (shader
    interpreted per hit, general shade class, bilinear fetches whose weights (1 - t) + t must sum to exactly 1), and the
    answer is the same closed form — exact in the oracle, and the kernels' bodies reproduce the oracle bit for bit."""
    from test_hostsim_parity import run_hostsim
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    w = h = 24
    scene = akr.load_scene(sv.write_furnace(tmp_path, RHO, EMIT, textured=True)).set_resolution(w, h)
    pmj, bn = tables
    task = _furnace_task(akr, 8, use_nee=0, rr_depth=100)
    ofilm, ost, _ = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn)
    err = float(np.abs(oracle.resolve(ofilm, w * h) - FURNACE).max())
    measured(f"textured white furnace (oracle) 24x24@8: max |pixel - {FURNACE}| = {err:.2e} (<= 1e-5)")
    assert err <= 1e-5 and ost.segments == w * h * 8 * (DEPTH + 1)
    lib.hostsim_set_pipeline(fused)
    try:
        film, _, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
        assert st.any_dynamic == 1 and st.any_alpha == 0
        assert np.array_equal(film, ofilm) and st.segments == ost.segments
    finally:
        lib.hostsim_set_pipeline(0)


def test_nee_on_off_equal_means_oracle(oracle, tables, cbox, cbox_task):
    w = h = 48
    scene = cbox(w, h)
    pmj, bn = tables
    on = cbox_task(256, use_nee=1)
    off = cbox_task(1024, use_nee=0)
    a = oracle.resolve(oracle.render(scene.desc, w, h, on.pt, on.sampler, on.filter, pmj, bn)[0], w * h).reshape(h, w, 3)
    b = oracle.resolve(oracle.render(scene.desc, w, h, off.pt, off.sampler, off.filter, pmj, bn)[0], w * h).reshape(h, w, 3)
    _equal_means(a, b, "cbox 48x48 NEE on @256 vs off @1024 (oracle)")


def _equal_means(a, b, what, sigmas=4.0):
    """Image means of two estimators agree within `sigmas` standard errors of their difference (8x8-pixel block means as
    the independent observations)."""
    h, w, _ = a.shape
    blk = lambda x: x.astype(np.float64).reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3)).reshape(-1, 3)
    d = blk(a) - blk(b)
    mean, se = d.mean(axis=0), d.std(axis=0, ddof=1) / np.sqrt(len(d))
    z = np.abs(mean) / se
    rel = np.abs(mean) / blk(a).mean(axis=0)
    measured(f"{what}: relative difference of the means {rel.round(5)}, z = {z.round(2)} (<= {sigmas})")
    assert (z <= sigmas).all() and (rel <= 0.02).all()


# ---- GPU: the same pins through the CUDA path ----------------------------------------------------------------------
@pytest.mark.gpu
def test_white_furnace_gpu(akr, tables, tmp_path):
    w = h = 64
    scene = akr.load_scene(sv.write_furnace(tmp_path, RHO, EMIT)).set_resolution(w, h)
    for fused in (0, 2):  # both pipelines (the furnace is a flat scene: 6 pair primitives)
        pt = akr.PathTracer(0)
        pt.set_engine_options(fused=fused)
        rgb = pt.render(scene, _furnace_task(akr, 16, use_nee=0, rr_depth=100)).to_rgb()
        st = pt.stats()
        err = float(np.abs(rgb - FURNACE).max())
        measured(f"white furnace (GPU, fused={fused}, no NEE, no RR) 64x64@16: max |pixel - {FURNACE}| = {err:.2e} (<= 1e-5)")
        assert err <= 1e-5
        assert st.segments == w * h * 16 * (DEPTH + 1) and st.shadow_rays == 0
        for use_nee, rr in ((1, 100), (1, 1)):
            rgb = pt.render(scene, _furnace_task(akr, 64, use_nee, rr)).to_rgb()
            _mean_check(rgb, FURNACE, f"white furnace (GPU, fused={fused}, use_nee={use_nee}, rr_depth={rr}) 64x64@64")
        pt.close()


@pytest.mark.gpu
def test_nee_on_off_equal_means_gpu(akr, cbox, cbox_task):
    w = h = 96
    scene = cbox(w, h)
    pt = akr.PathTracer(0)
    a = pt.render(scene, cbox_task(256, use_nee=1)).to_rgb()
    b = pt.render(scene, cbox_task(4096, use_nee=0)).to_rgb()
    pt.close()
    _equal_means(a, b, "cbox 96x96 NEE on @256 vs off @4096 (GPU)")
