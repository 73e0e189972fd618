"""The two C-ABI libraries load and export every symbol the headers declare; without a GPU the CUDA library
fails loudly instead of falling back to anything (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(akr_(?:b200|host)_\w+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol(akr):
    from akari_render_b200 import _abi
    names = _declared("akari_b200.h")
    assert len(names) >= 18 and set(names) == set(_abi.CUDA_SYMBOLS)
    lib = C.CDLL(_abi.CUDA_LIB)
    for n in names:
        assert hasattr(lib, n), n


def test_host_library_exports_every_declared_symbol(akr):
    from akari_render_b200 import _abi
    names = _declared("akari_b200_host.h")
    assert set(names) == set(_abi.HOST_SYMBOLS)
    lib = C.CDLL(_abi.HOST_LIB)
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_a_device(akr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(akr.AkariError) as e:
        akr.PathTracer(0)
    assert e.value.code == 2  # AKR_ERR_CUDA


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under akari_render_b200/ or include/ may reference it."""
    for base, _, files in os.walk(os.path.join(ROOT, "akari_render_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle's", "") or f in ("akr_math.cuh",), (f,)


def test_write_image_exr_and_pfm_round_trip(akr, tmp_path):
    """util::write_image for `.exr` (util/mod.rs:95-127: linear RGB f32) through the host library, read back with an
    independent decoder (OpenCV's OpenEXR / PFM readers)."""
    import os
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import numpy as np
    cv2 = __import__("pytest").importorskip("cv2")
    rgb = (np.random.default_rng(3).random((9, 13, 3)) * 4.0).astype(np.float32)
    rgb[0, 0] = (0.0, 1e-6, 1e4)
    for ext in ("exr", "pfm"):
        path = str(tmp_path / f"img.{ext}")
        akr.write_image(path, rgb)
        back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if back is None and ext == "exr":
            __import__("pytest").skip("this OpenCV build has no OpenEXR reader")
        assert back is not None and back.dtype == np.float32 and back.shape == rgb.shape
        assert np.array_equal(back[..., ::-1], rgb)  # OpenCV returns BGR


def test_write_image_png_is_the_reference_ldr_conversion(akr, tmp_path):
    """util::write_image for non-`.exr` paths (util/mod.rs:64-94): linear -> sRGB (color.rs:565-571), then
    `(x * 255.0).clamp(0.0, 255.0) as u8` (truncation, NaN -> 0), 8-bit RGB; read back with OpenCV's png decoder."""
    import numpy as np
    cv2 = __import__("pytest").importorskip("cv2")
    rng = np.random.default_rng(8)
    rgb = (rng.random((17, 23, 3)) ** 3 * 1.4).astype(np.float32)
    rgb[0, 0] = (0.0, 0.0031308, 0.0031309)
    rgb[0, 1] = (-1.0, 1.0, 50.0)
    rgb[0, 2] = (np.nan, 0.5, 0.25)
    path = str(tmp_path / "img.png")
    akr.write_image(path, rgb)
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert back is not None and back.dtype == np.uint8 and back.shape == rgb.shape
    lin = rgb.astype(np.float32)
    with np.errstate(invalid="ignore"):
        srgb = np.where(lin <= np.float32(0.0031308), lin * np.float32(12.92), np.power(lin, np.float32(1.0 / 2.4)) * np.float32(1.055) - np.float32(0.055))
        want = np.nan_to_num(np.clip(srgb * np.float32(255.0), 0.0, 255.0), nan=0.0).astype(np.uint8)
    d = np.abs(back[..., ::-1].astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and (d == 0).mean() > 0.995, (int(d.max()), float((d == 0).mean()))  # (powf of numpy vs libm: last-ulp ties)
    # -1 -> 0; 50 -> 255; NaN -> 0; and linear 1.0 -> 254: 1.0 * 1.055f - 0.055f = 0.99999994f, * 255 = 254.99998, truncated (as the reference does)
    assert tuple(int(v) for v in back[0, 1, ::-1]) == (0, 254, 255) and back[0, 2, 2] == 0
    with __import__("pytest").raises(akr.AkariError):
        akr.write_image(str(tmp_path / "img.bmp"), rgb)
