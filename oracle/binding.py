"""ctypes binding of oracle/liboracle.so (the CPU restatement).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class AkrOracleStats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64),
        ("segments", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("seconds", C.c_double),
        ("threads", C.c_uint32),
        ("n_lights", C.c_uint32),
    ]


def build(force=False):
    src = os.path.join(HERE, "akari_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        vp = C.c_void_p
        l.akr_oracle_last_error.restype = C.c_char_p
        l.akr_oracle_render.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_int, vp, vp, C.POINTER(AkrOracleStats)]
        l.akr_oracle_render.restype = C.c_int
        l.akr_oracle_render_aov.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, vp]
        l.akr_oracle_render_aov.restype = C.c_int
        l.akr_oracle_resolve.argtypes = [vp, C.c_size_t, vp]
        l.akr_oracle_resolve.restype = None
        l.akr_oracle_xxhash32_4.argtypes = [C.c_uint32] * 4
        l.akr_oracle_xxhash32_4.restype = C.c_uint32
        l.akr_oracle_permute_element.argtypes = [C.c_uint32] * 4
        l.akr_oracle_permute_element.restype = C.c_uint32
        l.akr_oracle_sampler_stream.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp,
                                                C.c_uint32, vp]
        l.akr_oracle_sampler_stream.restype = C.c_int
        l.akr_oracle_alias_table.argtypes = [vp, C.c_uint32, vp, vp, vp]
        l.akr_oracle_alias_table.restype = None
        l.akr_oracle_alias_sample.argtypes = [vp, vp, vp, C.c_uint32, C.c_float, vp, vp, vp]
        l.akr_oracle_alias_sample.restype = None
        l.akr_oracle_camera_ray.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, vp, vp]
        l.akr_oracle_camera_ray.restype = C.c_int
        l.akr_oracle_scene_lights.argtypes = [vp, vp, vp, vp, C.c_uint32]
        l.akr_oracle_scene_lights.restype = C.c_int
        l.akr_oracle_bsdf_eval.argtypes = [C.c_int, vp, C.c_float, C.c_float, vp, vp, vp, vp]
        l.akr_oracle_bsdf_eval.restype = None
        l.akr_oracle_bsdf_sample.argtypes = [C.c_int, vp, C.c_float, C.c_float, vp, C.c_float, C.c_float, C.c_float, vp, vp]
        l.akr_oracle_bsdf_sample.restype = None
        l.akr_oracle_bsdf_chi2_tables.argtypes = [C.c_int, vp, C.c_float, C.c_float, vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]
        l.akr_oracle_bsdf_chi2_tables.restype = None
        l.akr_oracle_make_albedo_table.argtypes = [vp, C.c_uint32]
        l.akr_oracle_make_albedo_table.restype = None
        _lib = l
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


_albedo_cache = {}


def albedo_table(n=64):
    """Deterministic ggx_dielectric_s table, [16,16,16] f32 (z, y, x order flattened x-fastest)."""
    if n not in _albedo_cache:
        path = os.path.join(HERE, f"_albedo_table_{n}.npy")
        if os.path.exists(path):
            _albedo_cache[n] = np.load(path)
        else:
            t = np.zeros(4096, dtype=np.float32)
            lib().akr_oracle_make_albedo_table(_ptr(t), n)
            try:
                np.save(path, t)
            except OSError:
                pass
            _albedo_cache[n] = t
    return _albedo_cache[n]


def render(scene_desc_ptr, width, height, pt_cfg, sampler_cfg, filter_cfg, pmj, bn, table=None, y0=0, y1=None,
           spp_begin=0, spp_end=None, threads=0, film=None, want_first_hits=False):
    """Render with the oracle.  Returns (film_7n float32, stats, first_hits or None)."""
    if y1 is None:
        y1 = height
    if spp_end is None:
        spp_end = pt_cfg.spp
    if table is None:
        table = albedo_table()
    n = width * (y1 - y0)
    if film is None:
        film = np.zeros(7 * n, dtype=np.float32)
    fh = np.zeros((n, 2), dtype=np.uint32) if want_first_hits else None
    st = AkrOracleStats()
    rc = lib().akr_oracle_render(C.cast(scene_desc_ptr, C.c_void_p), C.byref(pt_cfg), C.byref(sampler_cfg), C.byref(filter_cfg),
                                 _ptr(pmj), _ptr(bn), _ptr(table), y0, y1, spp_begin, spp_end, threads, _ptr(film),
                                 _ptr(fh) if fh is not None else None, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"oracle render failed ({rc}): {lib().akr_oracle_last_error().decode()}")
    return film, st, fh


def resolve(film, n_pixels):
    out = np.zeros(n_pixels * 3, dtype=np.float32)
    lib().akr_oracle_resolve(_ptr(film), n_pixels, _ptr(out))
    return out


def bsdf_chi2_tables(kind, color, roughness, eta, wo, n_samples, seed, theta_res, phi_res):
    """(observed histogram u32[theta_res * phi_res], expected counts f64[...]) of one tap closure for one wo
    (akari_test.rs:31-112).  kind: 0 diffuse, 1 GGX reflection, 2 GGX transmission, 3 GGX conductor."""
    color = np.asarray(color, np.float32)
    wo = np.asarray(wo, np.float32)
    hist = np.zeros(theta_res * phi_res, np.uint32)
    exp = np.zeros(theta_res * phi_res, np.float64)
    lib().akr_oracle_bsdf_chi2_tables(kind, _ptr(color), roughness, eta, _ptr(wo), n_samples, seed, theta_res, phi_res, _ptr(hist), _ptr(exp))
    return hist, exp


def render_aov(scene_desc_ptr, width, height, aov_cfg, sampler_cfg, filter_cfg, pmj, bn, table=None, y0=0, y1=None):
    """The `aov` integrator (aov.rs) with the oracle.  Returns the film (7n float32)."""
    if y1 is None:
        y1 = height
    if table is None:
        table = albedo_table()
    film = np.zeros(7 * width * (y1 - y0), dtype=np.float32)
    rc = lib().akr_oracle_render_aov(C.cast(scene_desc_ptr, C.c_void_p), C.byref(aov_cfg), C.byref(sampler_cfg), C.byref(filter_cfg), _ptr(pmj), _ptr(bn),
                                     _ptr(table), y0, y1, _ptr(film))
    if rc != 0:
        raise RuntimeError(f"oracle aov render failed ({rc}): {lib().akr_oracle_last_error().decode()}")
    return film
