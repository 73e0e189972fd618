"""The `aov` method (crates/akari_integrator/src/aov.rs): first-hit shading / geometric normal, tangent, bitangent,
albedo + emission and lobe roughness accumulated into the film.  CPU: the kernels' bodies (tests/hostsim) against the
oracle's literal closure tree, bit for bit, over cbox, the Principled / node variants and the clutter scene (per-corner
normals and tangents).  GPU: the CUDA path against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import scene_variants as sv
from conftest import image_rel_l2, measured, rel_l2_per_pixel

HERE = os.path.dirname(os.path.abspath(__file__))
AOVS = ["ns", "ng", "tangent", "bitangent", "albedo", "roughness"]


def _task(akr, aov, spp=4, remap=True):
    return akr.RenderTask.from_json('{"method": {"type": "aov", "spp": %d, "aov": "%s", "remap": %s}, "sampler": {"type": "pmj02bn", "seed": 0},'
                                    ' "film": {"filter": {"type": "gaussian", "radius": 1.5}, "out": "aov.exr"}}' % (spp, aov, "true" if remap else "false"))


def _scenes(akr, tmp_path):
    yield "cbox", akr.load_scene(os.path.join(os.path.dirname(HERE), "scenes", "cbox", "scene.json"))
    yield "principled_mix", akr.load_scene(sv.write_variant(tmp_path, "pm", sv.variant_principled_mix))
    yield "nodes", akr.load_scene(sv.write_variant(tmp_path, "nodes", sv.variant_nodes))
    yield "clutter", akr.load_scene(sv.write_clutter(tmp_path, n_lon=12, n_lat=8))


def test_method_file_parses_aov(akr):
    t = _task(akr, "roughness", spp=7, remap=False)
    assert t.method == "aov" and (t.aov.spp, t.aov.aov, t.aov.remap) == (7, 5, 0)
    d = akr.RenderTask.from_json('{"method": {"type": "aov"}, "film": {"out": "x.exr"}}')
    assert (d.aov.spp, d.aov.aov, d.aov.remap) == (256, 0, 1)  # aov::Config::default (aov.rs:29-36)
    with pytest.raises(akr.AkariError):
        akr.RenderTask.from_json('{"method": {"type": "aov", "aov": "depth"}, "film": {"out": "x.exr"}}')


def test_aov_hostsim_bitwise(akr, oracle, tables, tmp_path):
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    pmj, bn = tables
    table = oracle.albedo_table()
    w = h = 24
    for name, scene in _scenes(akr, tmp_path):
        scene.set_resolution(w, h)
        for aov in AOVS:
            for remap in (True, False) if aov in ("ns", "tangent") else (True,):
                task = _task(akr, aov, remap=remap)
                ref = oracle.render_aov(scene.desc, w, h, task.aov, task.sampler, task.filter, pmj, bn)
                film = np.zeros(7 * w * h, np.float32)
                rc = lib.hostsim_render_aov(scene.desc, C.byref(task.raw.aov), C.byref(task.raw.sampler), C.byref(task.raw.filter), C.c_void_p(pmj.ctypes.data),
                                            C.c_void_p(bn.ctypes.data), C.c_void_p(table.ctypes.data), 0, h, C.c_void_p(film.ctypes.data))
                assert rc == 0, lib.hostsim_last_error()
                assert np.array_equal(film, ref), (name, aov, remap)
        # sanity of the content: geometric normals are unit vectors wherever something was hit
        task = _task(akr, "ng", remap=False)
        f = oracle.render_aov(scene.desc, w, h, task.aov, task.sampler, task.filter, pmj, bn)
        n = f[:3 * w * h].reshape(-1, 3) / 4.0
        ln = np.linalg.norm(n, axis=1)
        assert ((ln < 1.0 + 1e-5)).all() and (ln > 0.3).mean() > 0.5


@pytest.mark.gpu
def test_aov_gpu_parity(akr, oracle, tables, tmp_path):
    pmj, bn = tables
    w = h = 64
    pt = akr.PathTracer(0)
    for name, scene in _scenes(akr, tmp_path):
        scene.set_resolution(w, h)
        worst = 0.0
        for aov in AOVS:
            task = _task(akr, aov, spp=8)
            ref = oracle.resolve(oracle.render_aov(scene.desc, w, h, task.aov, task.sampler, task.filter, pmj, bn), w * h).reshape(h, w, 3)
            got = pt.render_aov(scene, task).to_rgb()
            # a camera sample that lands on the other side of a silhouette changes the pixel by a whole unit vector / 8:
            # count pixels, as the radiance tests do
            bad = float((rel_l2_per_pixel(got, ref) > 1e-3).mean())
            worst = max(worst, bad)
            assert bad <= 2e-3, (name, aov, bad)
        measured(f"aov {name} 64x64@8, six outputs: worst fraction of pixels over 1e-3 = {worst:.3e} (<= 2e-3)")
    pt.close()
