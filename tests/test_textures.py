"""Texture-driven shading: image textures (raw float + png), texcoords / extract / mapping, checkerboard,
normal map, separate colour, and the stochastic alpha test of the traversal (scene.rs:49-86) that only texture alpha
channels can trigger.  CPU: host loader, the kernels' bodies (tests/hostsim, both pipelines) bit for bit against the
oracle's per-dispatch SVM interpreter.  GPU: the CUDA path against the oracle at the stated tolerance."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import scene_variants as sv
from conftest import image_rel_l2, measured, rel_l2_per_pixel
from test_hostsim_parity import run_hostsim

HERE = os.path.dirname(os.path.abspath(__file__))


class AkrImage(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("texel_format", C.c_uint32), ("address", C.c_uint32),
                ("filter", C.c_uint32), ("_pad", C.c_uint32)]


def _images(scene):
    d = scene.desc.contents
    arr = C.cast(d.images, C.POINTER(AkrImage))
    return [arr[i] for i in range(d.n_images)]


def test_loader_decodes_png_and_float_images(akr, tmp_path):
    """png: every filter type, flipped vertically like DynamicImage::flipv (load.rs:590), RGBA8; raw float: channels
    padded to RGBA with alpha 1, not flipped (load.rs:556-588); sampler modes mapped as load.rs:680-702."""
    path = sv.write_textured(tmp_path)
    scene = akr.load_scene(path)
    sj = json.load(open(path))
    blob = open(os.path.join(os.path.dirname(path), "Scene.bin"), "rb").read()
    imgs = _images(scene)
    assert len(imgs) == 5
    seen = 0
    for mat in sj["materials"].values():
        for node in mat["shader"]["nodes"].values():
            if node["type"] != "image":
                continue
            im = node["image"]
            v = sj["buffer_views"][im["data"]["id"]]
            raw = blob[v["offset"]:v["offset"] + v["length"]]
            match = [g for g in imgs if (g.width, g.height) == (im["width"], im["height"]) and
                     g.address == {"repeat": 0, "clip": 1, "mirror": 2, "extend": 3}[im["extension"]] and
                     g.filter == 1 and g.texel_format == (1 if im["format"] == "float" else 0)]  # nearest -> LinearPoint = bilinear too (load.rs:694)
            assert len(match) == 1
            g = match[0]
            n = g.width * g.height * 4
            if im["format"] == "float":
                got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_float)), (n,)).reshape(g.height, g.width, 4)
                src = np.frombuffer(raw, np.float32).reshape(im["height"], im["width"], im["channels"])
                assert np.array_equal(got[..., :im["channels"]], src) and (got[..., 3] == 1.0).all()
            else:
                import cv2
                got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_uint8)), (n,)).reshape(g.height, g.width, 4)
                ref = cv2.imdecode(np.frombuffer(raw, np.uint8), cv2.IMREAD_UNCHANGED)  # BGRA, independent decoder
                assert np.array_equal(got, ref[::-1, :, [2, 1, 0, 3]])
            seen += 1
    assert seen == 5


def test_unsupported_nodes_and_formats_are_rejected(akr, tmp_path):
    path = sv.write_textured(tmp_path)
    sj = json.load(open(path))
    bad = json.loads(json.dumps(sj))
    bad["materials"]["floor_001"]["shader"]["nodes"]["tex"]["image"]["format"] = "dds"
    p2 = os.path.join(os.path.dirname(path), "bad.json")
    json.dump(bad, open(p2, "w"))
    with pytest.raises(akr.AkariError):
        akr.load_scene(p2)
    bad = json.loads(json.dumps(sj))
    bad["materials"]["floor_001"]["shader"]["nodes"]["n"] = {"type": "noise", "dim": 2, "scale": {"id": "cb_scale"}}
    bad["materials"]["floor_001"]["shader"]["nodes"]["tex"]["uv"] = {"id": "n"}
    json.dump(bad, open(p2, "w"))
    with pytest.raises(akr.AkariError):
        akr.load_scene(p2)


@pytest.mark.parametrize("fused", [0, 1], ids=["queued", "fused"])
@pytest.mark.parametrize("alpha", [True, False], ids=["alpha_cutout", "opaque"])
def test_textured_scene_hostsim_bitwise(akr, oracle, tables, cbox_task, tmp_path, fused, alpha):
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    lib.hostsim_set_pipeline(fused)
    try:
        w = h = 40
        scene = akr.load_scene(sv.write_textured(tmp_path, alpha_cutout=alpha)).set_resolution(w, h)
        # the loader binds every image bilinearly (load.rs:690-699); keep the ABI's point filter covered: the back wall's png
        for g in _images(scene):  # (ctypes pointer indexing returns views of the descriptor's own records)
            if (g.width, g.height) == (12, 16):
                g.filter = 0
        assert sum(1 for g in _images(scene) if g.filter == 0) == 1
        task = cbox_task(8)
        pmj, bn = tables
        ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
        film, fh, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
        assert np.array_equal(fh, ofh)
        assert (st.segments, st.shadow_rays) == (ost.segments, ost.shadow_rays)
        assert np.array_equal(film, ofilm)
        # opaque textures (every texel alpha 1, no zero address mode) need no alpha test at all: the scene keeps the plain
        # traversal and, being flat, the fused pipeline (scene_build.cpp alpha_varies)
        assert st.any_dynamic == 1 and st.any_alpha == (1 if alpha else 0)
        # the primitive intersector (what the kernels run).  Opaque variant: the usual per-pixel gate.  Alpha cutout: the
        # alpha test hashes the BITS of the barycentrics (scene.rs:57-63), which differ in the last place between two
        # intersectors (as between the reference's own OptiX and Embree back ends), so every stochastic decision is
        # redrawn: the images agree statistically, not per pixel.
        lib.hostsim_set_intersector(1)
        film1, fh1, _ = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
        a, b = akr.Film(film1, w, h).to_rgb(), oracle.resolve(ofilm, w * h).reshape(h, w, 3)
        if alpha:
            ma, mb = a.reshape(-1, 3).mean(axis=0), b.reshape(-1, 3).mean(axis=0)
            assert (np.abs(ma - mb) / mb < 0.05).all(), (ma, mb)
        else:
            assert float((rel_l2_per_pixel(a, b) > 1e-3).mean()) <= 2e-2
    finally:
        lib.hostsim_set_intersector(0)
        lib.hostsim_set_pipeline(0)


def test_alpha_cutout_changes_the_image_and_stays_unbiased_in_alpha(akr, oracle, tables, cbox_task, tmp_path):
    """The stochastic alpha test is live: rays pass through the short box in proportion to 1 - alpha."""
    w = h = 32
    pmj, bn = tables
    task = cbox_task(16)
    a = akr.load_scene(sv.write_textured(tmp_path, alpha_cutout=True)).set_resolution(w, h)
    b = akr.load_scene(sv.write_textured(tmp_path, alpha_cutout=False)).set_resolution(w, h)
    fa, sa, ha = oracle.render(a.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    fb, sb, hb = oracle.render(b.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    box = 6  # instance id of the short box (BTreeMap order)
    hits_a, hits_b = int((ha[:, 0] == box).sum()), int((hb[:, 0] == box).sum())
    assert 0.2 * hits_b < hits_a < 0.9 * hits_b, (hits_a, hits_b)  # alpha patches are 0 / 0.5 / 1 in equal shares
    assert not np.array_equal(fa, fb)


@pytest.mark.gpu
def test_textured_scene_gpu_parity(akr, oracle, tables, cbox_task, tmp_path):
    """Opaque textured variant: per-pixel gate against the oracle."""
    w = h = 128
    scene = akr.load_scene(sv.write_textured(tmp_path, alpha_cutout=False)).set_resolution(w, h)
    task = cbox_task(16)
    pmj, bn = tables
    pt = akr.PathTracer(0)
    pt.set_engine_options(aov_mask=1)
    film = pt.render(scene, task)
    st = pt.stats()
    fh = pt.first_hits()
    pt.close()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    same = float(((fh[0] == ofh[:, 0]) & (fh[1] == ofh[:, 1])).mean())
    a, b = film.to_rgb(), oracle.resolve(ofilm, w * h).reshape(h, w, 3)
    bad = float((rel_l2_per_pixel(a, b) > 1e-3).mean())
    img = image_rel_l2(a, b)
    ds = abs(int(st.segments) - int(ost.segments)) / ost.segments
    measured(f"textured cbox (opaque) 128x128@16: first hits equal {same:.5f} (>= 0.9999); pixels over 1e-3: {bad:.3e} (<= 1e-3); "
             f"image rel-L2 {img:.3e} (<= 1e-3); segments rel diff {ds:.2e} (<= 1e-3)")
    assert same >= 0.9999 and bad <= 1e-3 and img <= 1e-3 and ds <= 1e-3


@pytest.mark.gpu
def test_alpha_cutout_gpu_statistical_parity(akr, oracle, tables, cbox_task, tmp_path):
    """Alpha-cutout variant: the alpha test hashes the bits of the barycentrics, so a different intersector redraws every
    stochastic decision (see the hostsim test); GPU and oracle must agree in distribution: equal 8x8-block means within
    4 standard errors, first-hit counts on the cut-out box within 3 %, ray counts within 1 %."""
    from test_analytic_pins import _equal_means
    w = h = 96
    scene = akr.load_scene(sv.write_textured(tmp_path, alpha_cutout=True)).set_resolution(w, h)
    task = cbox_task(64)
    pmj, bn = tables
    pt = akr.PathTracer(0)
    pt.set_engine_options(aov_mask=1)
    film = pt.render(scene, task)
    st = pt.stats()
    fh = pt.first_hits()
    pt.close()
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    _equal_means(film.to_rgb(), oracle.resolve(ofilm, w * h).reshape(h, w, 3), "textured cbox (alpha cutout) 96x96@64, GPU vs oracle")
    box = 6
    ga, oa = int((fh[0] == box).sum()), int((ofh[:, 0] == box).sum())
    ds = abs(int(st.segments) - int(ost.segments)) / ost.segments
    measured(f"alpha cutout: first hits on the cut-out box gpu/oracle {ga}/{oa} (within 10 %), segments rel diff {ds:.2e} (<= 1e-2)")
    assert abs(ga - oa) <= 0.1 * oa and ds <= 1e-2


def test_loader_decodes_openexr_images(akr, oracle, tables, cbox_task, tmp_path):
    """OpenEXR textures (load.rs:590-610: decode, flipv, to_rgba32f): HALF / FLOAT channels stored alphabetically, ZIP (16-line
    chunks), ZIPS, RLE and uncompressed scanlines, a data window that does not start at the origin, RGB and RGBA.  All four
    codecs are lossless, so the texels must equal the source arrays exactly (rows flipped, missing alpha = 1), and the
    kernels' bodies still equal the oracle bit for bit on the scene.  Truncated files, a bad magic number and the PIZ
    codec are rejected."""
    path, sources = sv.write_exr_textured(tmp_path)
    scene = akr.load_scene(path)
    imgs = _images(scene)
    assert len(imgs) == len(sources) == 5
    for arr in sources.values():
        h, w, c = arr.shape
        match = [g for g in imgs if (g.width, g.height) == (w, h)]
        assert len(match) == 1 and match[0].texel_format == 1
        got = np.ctypeslib.as_array(C.cast(match[0].texels, C.POINTER(C.c_float)), (w * h * 4,)).reshape(h, w, 4)
        assert np.array_equal(got[..., :c], arr[::-1]) and (got[..., 3] == 1.0).all()
    # render parity on the scene (texture-driven materials on five surfaces)
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    w = h = 32
    scene.set_resolution(w, h)
    task = cbox_task(4)
    pmj, bn = tables
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    film, fh, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
    assert np.array_equal(film, ofilm) and st.any_dynamic == 1 and st.any_alpha == 0
    # malformed / unsupported files are rejected, not mis-read
    raw = sv._exr_bytes(np.ones((4, 4, 3), np.float32), "zip", "half")
    for bad, what in ((raw[:40], "truncated"), (b"\x00" + raw[1:], "magic"), (raw.replace(b"compression\0compression\0\x01\0\0\0\x03", b"compression\0compression\0\x01\0\0\0\x04"), "PIZ")):
        sj = json.load(open(path))
        blob = bytearray(open(os.path.join(os.path.dirname(path), "Scene.bin"), "rb").read())
        v = sj["buffer_views"][sj["materials"]["floor_001"]["shader"]["nodes"]["tex"]["image"]["data"]["id"]]
        v["offset"], v["length"] = len(blob), len(bad)
        blob.extend(bad)
        sj["buffers"]["Scene"]["length"] = len(blob)
        d = os.path.join(str(tmp_path), "bad_" + what)
        os.makedirs(d, exist_ok=True)
        open(os.path.join(d, "Scene.bin"), "wb").write(bytes(blob))
        json.dump(sj, open(os.path.join(d, "scene.json"), "w"))
        with pytest.raises(akr.AkariError):
            akr.load_scene(os.path.join(d, "scene.json"))


def test_loader_decodes_tiff_images(akr, tmp_path):
    """TIFF textures (load.rs:590-603: decode, flipv, to_rgba8) written by an independent encoder (OpenCV / libtiff):
    uncompressed, LZW, Deflate and PackBits strips, 8-bit gray / RGB / RGBA and 16-bit RGBA (-> 8 bit by
    round(v * 255 / 65535) like image 0.24's to_rgba8)."""
    import cv2
    rng = np.random.default_rng(5)
    cases = [("floor_001", (rng.random((33, 21, 3)) * 255).astype(np.uint8), 5), ("backWall_001", (rng.random((8, 40, 4)) * 255).astype(np.uint8), 1),
             ("leftWall_001", (rng.random((17, 9)) * 255).astype(np.uint8), 8), ("rightWall_001", (rng.random((12, 12, 3)) * 255).astype(np.uint8), 32773),
             ("ceiling_001", (rng.random((10, 14, 4)) * 65535).astype(np.uint16), 5)]
    cases[3][1][3:9, :, :] = 77  # runs for PackBits
    items, expect = [], {}
    for material, arr, compression in cases:
        bgr = arr if arr.ndim == 2 else arr[..., [2, 1, 0] + ([3] if arr.shape[2] == 4 else [])]
        ok, buf = cv2.imencode(".tiff", bgr, [cv2.IMWRITE_TIFF_COMPRESSION, compression])
        assert ok
        h, w = arr.shape[:2]
        c = 1 if arr.ndim == 2 else arr.shape[2]
        items.append((material, bytes(buf), "tiff", w, h, c))
        a8 = arr if arr.dtype == np.uint8 else ((arr.astype(np.uint32) + 128) // 257).astype(np.uint8)
        rgba = np.full((h, w, 4), 255, np.uint8)
        if c == 1:
            rgba[..., :3] = a8[..., None]
        else:
            rgba[..., :c] = a8
        expect[(w, h)] = rgba[::-1]
    scene = akr.load_scene(sv.write_image_textured(tmp_path, "tiff_textured", items, colorspace="srgb"))
    imgs = _images(scene)
    assert len(imgs) == 5
    for g in imgs:
        got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_uint8)), (g.width * g.height * 4,)).reshape(g.height, g.width, 4)
        assert g.texel_format == 0 and np.array_equal(got, expect[(g.width, g.height)])


def test_16_bit_png_converts_like_to_rgba8(akr, tmp_path):
    """16-bit PNG samples become 8-bit by round(v * 255 / 65535) = (v + 128) // 257 (image 0.24.7 `to_rgba8`), not by
    dropping the low byte."""
    import cv2
    rng = np.random.default_rng(9)
    arr = (rng.random((6, 5, 4)) * 65535).astype(np.uint16)
    arr[0, 0] = (255, 256, 383, 65535)  # values on which truncation and rounding differ
    ok, buf = cv2.imencode(".png", arr[..., [2, 1, 0, 3]])
    assert ok
    scene = akr.load_scene(sv.write_image_textured(tmp_path, "png16", [("floor_001", bytes(buf), "png", 5, 6, 4)], colorspace="srgb"))
    g = _images(scene)[0]
    got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_uint8)), (6 * 5 * 4,)).reshape(6, 5, 4)
    assert np.array_equal(got, (((arr.astype(np.uint32) + 128) // 257).astype(np.uint8))[::-1])


def test_clip_address_mode_keeps_the_alpha_test_on_an_opaque_texture(akr, oracle, tables, cbox_task, tmp_path):
    """An image whose texels are all opaque still yields alpha 0 outside [0, 1)^2 under the `clip` (zero) address mode
    (the sampler returns (0, 0, 0, 0) there), so a surface that samples it beyond the unit square must keep the stochastic
    alpha test: rays pass through those parts.  `alpha_varies` (scene_build.cpp) has to flag it; the kernels' bodies then
    equal the oracle — which alpha-tests every texture-driven material — bit for bit, with hits seen behind the wall."""
    import cv2
    rng = np.random.default_rng(3)
    png = (rng.random((8, 8, 3)) * 255).astype(np.uint8)  # three channels: decoded alpha is 255 everywhere
    ok, buf = cv2.imencode(".png", png)
    assert ok
    path = sv.write_image_textured(tmp_path, "clip_opaque", [("backWall_001", bytes(buf), "png", 8, 8, 3)], colorspace="srgb")
    sj = json.load(open(path))
    nodes = sj["materials"]["backWall_001"]["shader"]["nodes"]
    nodes["tex"]["image"]["extension"] = "clip"
    nodes["tc"] = {"type": "texcoords"}
    nodes["uv"] = {"type": "extract", "node": {"id": "tc"}, "field": "uv"}
    nodes["m_loc"] = {"type": "float3", "value": [-0.5, -0.5, 0.0]}
    nodes["m_rot"] = {"type": "float3", "value": [0.0, 0.0, 0.0]}
    nodes["m_scale"] = {"type": "float3", "value": [2.0, 2.0, 1.0]}
    nodes["map"] = {"type": "mapping", "vector": {"id": "uv"}, "mapping": "point", "location": {"id": "m_loc"}, "rotation": {"id": "m_rot"},
                    "scale": {"id": "m_scale"}}
    nodes["tex"]["uv"] = {"id": "map"}  # uv * 2 - 0.5: the texture covers the middle of the wall, the rim samples outside
    json.dump(sj, open(path, "w"))
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    w = h = 40
    scene = akr.load_scene(path).set_resolution(w, h)
    task = cbox_task(8)
    pmj, bn = tables
    ofilm, ost, ofh = oracle.render(scene.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    film, fh, st = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
    assert st.any_alpha == 1
    assert np.array_equal(fh, ofh) and np.array_equal(film, ofilm)
    back_wall = 3  # instance id
    plain = akr.load_scene(os.path.join(sv.CBOX_DIR, "scene.json")).set_resolution(w, h)
    _, _, pfh = oracle.render(plain.desc, w, h, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    n_clip, n_plain = int((ofh[:, 0] == back_wall).sum()), int((pfh[:, 0] == back_wall).sum())
    assert 0 < n_clip < 0.8 * n_plain, (n_clip, n_plain)  # the rim of the wall lets camera rays through


def test_loader_decodes_jpeg_images(akr, tmp_path):
    """JPEG textures (load.rs:590-603).  JPEG is lossy and decoders differ in their last bit (IDCT, chroma upsampling,
    colour conversion); decode_jpeg follows jpeg-decoder 0.3's published structure (exactness ASSUMED), so
    the test bounds its distance to an independent decoder (OpenCV / libjpeg-turbo): 4:4:4 and gray within 2 / 255 per
    texel, subsampled chroma (4:2:2, 4:2:0, odd sizes: MCU padding, edge replication) within 1 / 255 on average, with
    restart intervals; sequential and progressive scan structures."""
    import cv2
    rng = np.random.default_rng(21)

    def smooth(h, w, c):  # band-limited content (what textures are): sums of low-frequency waves + mild noise
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.zeros((h, w, c), np.float32)
        for k in range(c):
            for _ in range(4):
                fx, fy, ph = rng.random() * 0.25, rng.random() * 0.25, rng.random() * 6.28
                img[..., k] += np.sin(xx * fx + yy * fy + ph)
        img = (img - img.min()) / (img.max() - img.min())
        return (img * 235 + 10 + rng.random((h, w, c)) * 6).clip(0, 255).astype(np.uint8)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    cases = [("floor_001", smooth(40, 56, 3), [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_QUALITY, 92], 2.0, 0.6),
             ("backWall_001", smooth(33, 47, 3), [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_QUALITY, 90], 8.0, 1.0),
             ("leftWall_001", smooth(24, 31, 3), [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, 3], 8.0, 1.0),
             ("rightWall_001", smooth(19, 26, 1)[..., 0], [cv2.IMWRITE_JPEG_QUALITY, 80], 2.0, 0.6),
             ("ceiling_001", smooth(16, 16, 3), [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_QUALITY, 95], 8.0, 1.0)]
    items, expect = [], {}
    for material, arr, params, max_err, mean_err in cases:
        ok, buf = cv2.imencode(".jpg", arr if arr.ndim == 2 else arr[..., ::-1], params)
        assert ok
        h, w = arr.shape[:2]
        items.append((material, bytes(buf), "jpeg", w, h, 1 if arr.ndim == 2 else 3))
        ref = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
        ref = np.repeat(ref[..., None], 3, axis=2) if ref.ndim == 2 else ref[..., ::-1]
        expect[(w, h)] = (ref[::-1].astype(np.int32), max_err, mean_err)
    scene = akr.load_scene(sv.write_image_textured(tmp_path, "jpeg_textured", items, colorspace="srgb"))
    imgs = _images(scene)
    assert len(imgs) == 5
    for g in imgs:
        got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_uint8)), (g.width * g.height * 4,)).reshape(g.height, g.width, 4).astype(np.int32)
        ref, max_err, mean_err = expect[(g.width, g.height)]
        d = np.abs(got[..., :3] - ref)
        assert g.texel_format == 0 and (got[..., 3] == 255).all()
        assert d.max() <= max_err and d.mean() <= mean_err, ((g.width, g.height), int(d.max()), float(d.mean()))
    # progressive files (spectral selection + successive approximation, DC interleaved / AC per component, EOB runs)
    for k, (shape, params) in enumerate([((45, 38, 3), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, 88]),
                                         ((24, 40, 3), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_RST_INTERVAL, 2]),
                                         ((30, 30, 1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, 60])]):
        arr = smooth(*shape)
        arr = arr[..., 0] if shape[2] == 1 else arr
        ok, buf = cv2.imencode(".jpg", arr if arr.ndim == 2 else arr[..., ::-1], params)
        assert ok and b"\xff\xc2" in bytes(buf)
        ref = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
        ref = (np.repeat(ref[..., None], 3, axis=2) if ref.ndim == 2 else ref[..., ::-1])[::-1].astype(np.int32)
        h, w = shape[:2]
        scene = akr.load_scene(sv.write_image_textured(tmp_path, f"jpeg_progressive{k}", [("floor_001", bytes(buf), "jpeg", w, h, shape[2])]))
        g = _images(scene)[0]  # (texels live as long as `scene`)
        got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_uint8)), (w * h * 4,)).reshape(h, w, 4).astype(np.int32)
        d = np.abs(got[..., :3] - ref)
        assert d.max() <= 8 and d.mean() <= 1.0, (k, int(d.max()), float(d.mean()))
    # an Adobe APP14 marker with transform 0 says the three components are R, G, B as they stand (no YCbCr conversion)
    ok, buf = cv2.imencode(".jpg", smooth(16, 24, 3), [S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
    raw = bytes(buf)
    adobe = b"\xff\xee\x00\x0eAdobe\x00\x64\x00\x00\x00\x00\x00"  # version 100, flags 0 / 0, transform 0
    assert raw[2:4] == b"\xff\xe0"  # the JFIF APP0 segment: dropped, libjpeg lets JFIF win over Adobe (jpeg-decoder 0.3 does not look at it)
    patched = raw[:2] + adobe + raw[4 + ((raw[4] << 8) | raw[5]):]
    ref = cv2.imdecode(np.frombuffer(patched, np.uint8), cv2.IMREAD_UNCHANGED)[..., ::-1][::-1].astype(np.int32)
    scene = akr.load_scene(sv.write_image_textured(tmp_path, "jpeg_adobe_rgb", [("floor_001", patched, "jpeg", 24, 16, 3)]))
    g = _images(scene)[0]
    got = np.ctypeslib.as_array(C.cast(g.texels, C.POINTER(C.c_uint8)), (24 * 16 * 4,)).reshape(16, 24, 4).astype(np.int32)
    assert np.abs(got[..., :3] - ref).max() <= 2
    # arithmetic-coded / truncated files are rejected
    ok, buf = cv2.imencode(".jpg", smooth(16, 16, 3))
    for bad in (bytes(buf)[:60], bytes(buf).replace(b"\xff\xc0", b"\xff\xc9", 1)):
        with pytest.raises(akr.AkariError):
            akr.load_scene(sv.write_image_textured(tmp_path, "jpeg_bad", [("floor_001", bad, "jpeg", 16, 16, 3)]))


def test_corrupt_image_files_are_rejected_not_crashed_on(akr, tmp_path):
    """450 mutants (byte flips, truncations, splices, bit flips) of valid png / jpeg / tiff / OpenEXR files through the host
    loader: each one loads or raises AkariError.  (tools/fuzz_image_decoders.py is the long form; 15 000 mutants ran clean
    under AddressSanitizer + UBSan after the IDCT overflow and the size-field allocations it found were fixed.)  Also the
    explicit size-field cases: a png / tiff header that claims an image far larger than the file can hold."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_image_decoders", os.path.join(os.path.dirname(HERE), "tools", "fuzz_image_decoders.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    ok, rejected = fz.run(450, 3, str(tmp_path))
    assert ok + rejected == 450 and ok > 50 and rejected > 50
    import cv2
    import struct
    png = bytearray(bytes(cv2.imencode(".png", np.zeros((8, 8, 3), np.uint8))[1]))
    png[16:24] = struct.pack(">II", 60000, 60000)  # IHDR width / height (the CRC is not checked by this decoder)
    with pytest.raises(akr.AkariError, match="out of range"):
        akr.load_scene(sv.write_image_textured(tmp_path, "huge_png", [("floor_001", bytes(png), "png", 8, 8, 3)]))
