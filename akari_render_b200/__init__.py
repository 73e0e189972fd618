"""akari_render_b200 — Blackwell-native unidirectional path tracer behind AkariRender's `pt` surface.

Python is only the thin host layer used by tests, bench.py and the CLI: it binds the two C-ABI
libraries declared in include/akari_b200.h (CUDA engine) and include/akari_b200_host.h (scene /
method front-end).  Names mirror the reference's user-facing surface:

    reference (Rust)                                         here
    -------------------------------------------------------  ------------------------------
    akari_render::load::load_from_path  (load.rs:63-72)       load_scene(path)
    RenderTask / RenderConfig JSON       (lib.rs:57-109)       RenderTask.from_file / .from_json
    pt::Config                           (pt.rs:916-944)       RenderTask.pt  (AkrPtConfig)
    PathTracer::new + Integrator::render (pt.rs:946-959,1056)  PathTracer(device).render(scene, task)
    Film::data / copy_to_rgba_image      (film.rs:66-76,120)   Film.data / Film.to_rgb()
    util::write_image                    (util/mod.rs:57-127)  write_image(path, rgb)

There is no CPU fallback: constructing a PathTracer without the CUDA library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (AkrEngineOptions, AkrFilterConfig, AkrPtConfig, AkrRenderTask, AkrSamplerConfig, AkrStats, AkrTile)

__all__ = ["load_scene", "Scene", "RenderTask", "PathTracer", "Film", "write_image", "sampler_tables", "AkariError"]

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class AkariError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[akr {code}] {msg}")
        self.code = code


_host = None


def host_lib():
    global _host
    if _host is None:
        _host = _abi.load_host_lib()
    return _host


def _host_check(rc):
    if rc != 0:
        raise AkariError(rc, host_lib().akr_host_last_error().decode("utf-8", "replace"))


_tables = None


def sampler_tables():
    """(pmj02bn u32[5*65536*2], bluenoise u16[48*128*128]) — static tables of the pmj02bn sampler."""
    global _tables
    if _tables is None:
        pmj = np.fromfile(os.path.join(DATA_DIR, "pmj02bn.u32"), dtype=np.uint32)
        bn = np.fromfile(os.path.join(DATA_DIR, "bluenoise.u16"), dtype=np.uint16)
        assert pmj.size == 5 * 65536 * 2 and bn.size == 48 * 128 * 128
        _tables = (pmj, bn)
    return _tables


class Scene:
    """Host-side scene (the plain-array AkrSceneDesc a Rust host would hand over)."""

    def __init__(self, handle):
        self._h = handle

    @property
    def desc(self):
        return host_lib().akr_host_scene_desc(self._h)

    @property
    def resolution(self):
        cam = self.desc.contents.camera
        return cam.width, cam.height

    def set_resolution(self, width, height):
        _host_check(host_lib().akr_host_scene_set_resolution(self._h, width, height))
        return self

    def close(self):
        if self._h:
            host_lib().akr_host_free_scene(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_scene(path):
    h = C.c_void_p()
    _host_check(host_lib().akr_host_load_scene(os.fsencode(path), C.byref(h)))
    return Scene(h)


class RenderTask:
    def __init__(self, raw=None):
        self.raw = raw if raw is not None else AkrRenderTask()
        if raw is None:
            host_lib().akr_host_default_task(C.byref(self.raw))

    @classmethod
    def from_file(cls, path):
        t = AkrRenderTask()
        _host_check(host_lib().akr_host_parse_method_file(os.fsencode(path), C.byref(t)))
        return cls(t)

    @classmethod
    def from_json(cls, text):
        t = AkrRenderTask()
        _host_check(host_lib().akr_host_parse_method_string(text.encode(), C.byref(t)))
        return cls(t)

    @property
    def pt(self):
        return self.raw.pt

    @property
    def sampler(self):
        return self.raw.sampler

    @property
    def filter(self):
        return self.raw.filter

    @property
    def out(self):
        return self.raw.out.decode("utf-8", "replace")

    @property
    def method(self):
        return "aov" if self.raw.method == 1 else "pt"

    @property
    def aov(self):
        return self.raw.aov


class Film:
    """Reference film layout (film.rs:66-76): f32 | rgb 3N | splat 3N | weight N |."""

    def __init__(self, data, width, rows):
        self.data = data
        self.width = width
        self.rows = rows

    def to_rgb(self):
        n = self.width * self.rows
        w = self.data[6 * n:7 * n]
        d = np.where(w == 0.0, np.float32(1.0), w).astype(np.float32)
        rgb = self.data[:3 * n].reshape(n, 3) / d[:, None] + self.data[3 * n:6 * n].reshape(n, 3) * np.float32(1.0)
        return rgb.reshape(self.rows, self.width, 3).astype(np.float32)


def write_image(path, rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = rgb.shape
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    _host_check(host_lib().akr_host_write_image(os.fsencode(path), rgb.ctypes.data_as(C.POINTER(C.c_float)), w, h))


class PathTracer:
    """One CUDA context on one device == the reference's `PathTracer` bound to a luisa `Device`."""

    def __init__(self, device=0, stream=None):
        self._lib = _abi.load_cuda_lib()  # raises when the extension is missing: no silent fallback
        self._ctx = C.c_void_p()
        rc = self._lib.akr_b200_create(device, C.byref(self._ctx))
        if rc != 0:
            raise AkariError(rc, "akr_b200_create failed (no usable CUDA device; there is no CPU fallback)")
        if stream is not None:
            self._check(self._lib.akr_b200_set_stream(self._ctx, C.c_void_p(stream)))
        pmj, bn = sampler_tables()
        self._check(self._lib.akr_b200_upload_sampler_tables(self._ctx, pmj.ctypes.data, bn.ctypes.data))
        self._scene = None
        self._tile = None
        self._res = (0, 0)

    def _check(self, rc):
        if rc != 0:
            raise AkariError(rc, self._lib.akr_b200_last_error(self._ctx).decode("utf-8", "replace"))

    def close(self):
        if self._ctx:
            self._lib.akr_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads ----
    def upload_sampler_tables(self, pmj, bn):
        self._check(self._lib.akr_b200_upload_sampler_tables(self._ctx, pmj.ctypes.data, bn.ctypes.data))

    def upload_albedo_table(self, table):
        table = np.ascontiguousarray(table, dtype=np.float32)
        assert table.size == 4096
        self._check(self._lib.akr_b200_upload_albedo_table(self._ctx, table.ctypes.data))

    def upload_scene(self, scene):
        self._check(self._lib.akr_b200_upload_scene(self._ctx, scene.desc))
        self._scene = scene
        self._res = scene.resolution

    def set_engine_options(self, wave_size=0, profile_stages=0, trace_mode=0, fused=0, smem_node_kb=0, aov_mask=0):
        o = AkrEngineOptions(wave_size, 0, profile_stages, trace_mode, fused, smem_node_kb, aov_mask)
        self._check(self._lib.akr_b200_set_engine_options(self._ctx, C.byref(o)))

    # ---- rendering ----
    def _set_tile(self, tile):
        """tile: None (whole sensor), (y0, y1) (contiguous band) or (y0, y1, block_rows, n_shards, shard) (interleaved)."""
        if tile is None:
            self._tile = AkrTile(0, self._res[1], 0, 0, 0, 0)
            return None
        t = tuple(tile) + (0, 0, 0)[: 5 - len(tile)] if len(tile) < 5 else tuple(tile)
        self._tile = AkrTile(t[0], t[1], t[2], t[3], t[4], 0)
        return self._tile

    @property
    def tile_rows(self):
        return int(self._lib.akr_b200_tile_rows(C.byref(self._tile)))

    def begin(self, task, tile=None):
        t = self._set_tile(tile)
        self._check(self._lib.akr_b200_begin(self._ctx, C.byref(task.raw.pt), C.byref(task.raw.sampler), C.byref(task.raw.filter),
                                             C.byref(t) if t is not None else None))

    def render_pass(self, n_spp, blocking=True):
        self._check(self._lib.akr_b200_render_pass(self._ctx, n_spp, 1 if blocking else 0))

    def render(self, scene, task, tile=None):
        """pt::render (pt.rs:1161-1172): uploads `scene` when it is not the resident one (or its resolution changed),
        renders all passes and returns the Film (host copy)."""
        if scene is not None and (scene is not self._scene or scene.resolution != self._res):
            self.upload_scene(scene)
        t = self._set_tile(tile)
        self._check(self._lib.akr_b200_render_pt(self._ctx, C.byref(task.raw.pt), C.byref(task.raw.sampler), C.byref(task.raw.filter),
                                                 C.byref(t) if t is not None else None))
        return self.download_film()

    def render_aov(self, scene, task, tile=None):
        """aov::render (aov.rs:175-185): the `aov` method of `task` (RenderTask with method type "aov")."""
        if scene is not None and (scene is not self._scene or scene.resolution != self._res):
            self.upload_scene(scene)
        t = self._set_tile(tile)
        self._check(self._lib.akr_b200_render_aov(self._ctx, C.byref(task.raw.aov), C.byref(task.raw.sampler), C.byref(task.raw.filter),
                                                  C.byref(t) if t is not None else None))
        return self.download_film()

    def synchronize(self):
        self._check(self._lib.akr_b200_synchronize(self._ctx))

    def download_film(self):
        rows = self.tile_rows
        n = self._res[0] * rows
        out = np.empty(7 * n, dtype=np.float32)
        self._check(self._lib.akr_b200_download_film(self._ctx, out.ctypes.data, out.size))
        return Film(out, self._res[0], rows)

    def resolve_rgb(self):
        rows = self.tile_rows
        out = np.empty((rows, self._res[0], 3), dtype=np.float32)
        self._check(self._lib.akr_b200_resolve_film(self._ctx, out.ctypes.data, out.size, 0))
        return out

    def resolve_into_device(self, device_ptr, n_floats, rgba=False):
        self._check(self._lib.akr_b200_resolve_film_device(self._ctx, C.c_void_p(device_ptr), n_floats, 1 if rgba else 0))

    def first_hits(self):
        rows = self.tile_rows
        n = self._res[0] * rows
        inst = np.empty(n, dtype=np.uint32)
        prim = np.empty(n, dtype=np.uint32)
        self._check(self._lib.akr_b200_debug_first_hits(self._ctx, inst.ctypes.data, prim.ctypes.data, n))
        return inst, prim

    def stats(self):
        s = AkrStats()
        self._check(self._lib.akr_b200_get_stats(self._ctx, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self._lib.akr_b200_reset_stats(self._ctx))
