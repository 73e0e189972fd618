"""Host front end (scene_loader.cpp) against restatements of the reference's loader rules: transforms (load.rs:129-171),
buffer kinds (scene.rs:96-117).  CPU only."""
import base64
import ctypes as C
import json
import os
import shutil

import numpy as np
import pytest

import scene_variants as sv


def _axis_angle(axis, a):
    """glam::Mat4::from_axis_angle (Rodrigues), float64."""
    x, y, z = axis
    c, s = np.cos(a), np.sin(a)
    t = 1.0 - c
    m = np.eye(4)
    m[:3, :3] = [[t * x * x + c, t * x * y - s * z, t * x * z + s * y], [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
                 [t * x * z - s * y, t * y * z + s * x, t * z * z + c]]
    return m


def _translation(t):
    m = np.eye(4)
    m[:3, 3] = t
    return m


def _reference_trs(tr, r, s, coord_sys, is_camera):
    """load.rs:134-165, line for line."""
    m = np.eye(4)
    if not is_camera:
        m = np.diag([s[0], s[1], s[2], 1.0]) @ m
    if coord_sys == "Akari":
        m = _axis_angle((0, 0, 1), r[2]) @ m
        m = _axis_angle((1, 0, 0), r[0]) @ m
        m = _axis_angle((0, 1, 0), r[1]) @ m
        m = _translation(tr) @ m
    else:
        if is_camera:
            m = _axis_angle((1, 0, 0), -np.pi / 2) @ m
        m = _axis_angle((1, 0, 0), r[0]) @ m
        m = _axis_angle((0, 0, 1), -r[1]) @ m
        m = _axis_angle((0, 1, 0), r[2]) @ m
        m = _translation((tr[0], tr[2], -tr[1])) @ m
    return m


def _mat(c16):
    return np.array(list(c16), np.float64).reshape(4, 4).T  # the ABI stores columns (glam layout)


def _write(tmp_path, name, scene, blob=None):
    d = os.path.join(str(tmp_path), name)
    os.makedirs(d, exist_ok=True)
    if blob is None:
        shutil.copy(os.path.join(sv.CBOX_DIR, "Scene.bin"), os.path.join(d, "Scene.bin"))
    else:
        open(os.path.join(d, "Scene.bin"), "wb").write(blob)
    p = os.path.join(d, "scene.json")
    json.dump(scene, open(p, "w"))
    return p


@pytest.mark.parametrize("coord_sys", ["Akari", "Blender"])
def test_trs_transforms_follow_load_rs(akr, tmp_path, coord_sys):
    """TRS transforms of instances (scale, then the three axis rotations in the reference's order, then the translation) and
    of the camera (no scale; Blender cameras start looking down) in both coordinate systems."""
    scene = json.load(open(os.path.join(sv.CBOX_DIR, "scene.json")))
    rng = np.random.default_rng(4)
    expect = {}
    for k, name in enumerate(sorted(scene["instances"])):
        tr, r, s = rng.uniform(-1, 1, 3), rng.uniform(-3, 3, 3), rng.uniform(0.5, 1.5, 3)
        scene["instances"][name]["transform"] = {"type": "trs", "data": {"translation": list(tr), "rotation": list(r), "scale": list(s),
                                                                          "coordinate_system": coord_sys}}
        expect[k] = _reference_trs(tr, r, s, coord_sys, False)  # instance ids follow the BTreeMap key order
    tr, r = rng.uniform(-1, 1, 3), rng.uniform(-3, 3, 3)
    scene["camera"]["data"]["transform"]["data"] = {"translation": list(tr), "rotation": list(r), "scale": [2.0, 3.0, 4.0], "coordinate_system": coord_sys}
    loaded = akr.load_scene(_write(tmp_path, "trs_" + coord_sys, scene))
    d = loaded.desc.contents
    for k, m in expect.items():
        assert np.allclose(_mat(d.instances[k].transform), m, rtol=0, atol=2e-6), k
    assert np.allclose(_mat(d.camera.c2w), _reference_trs(tr, r, None, coord_sys, True), rtol=0, atol=2e-6)


def test_matrix_transform_rows_are_matrix_rows(akr, tmp_path):
    """`Transform::Matrix` is `Mat4::from_cols_array_2d(m).transpose()`: the JSON rows are the rows of the matrix."""
    scene = json.load(open(os.path.join(sv.CBOX_DIR, "scene.json")))
    name = sorted(scene["instances"])[3]
    m = np.array([[1.0, 0.25, 0.0, 0.5], [0.0, 0.9, -0.1, 0.125], [0.0, 0.1, 1.1, -0.75], [0.0, 0.0, 0.0, 1.0]])
    scene["instances"][name]["transform"] = {"type": "matrix", "data": m.tolist()}
    loaded = akr.load_scene(_write(tmp_path, "matrix", scene))
    assert np.allclose(_mat(loaded.desc.contents.instances[3].transform), m, rtol=0, atol=1e-7)


def test_base64_and_binary_buffers_load_the_same_bytes(akr, tmp_path):
    """Buffer::EmbeddedBase64 / Buffer::EmbeddedBinary (scene.rs:99-105) against the usual Buffer::Path: the same meshes,
    vertex for vertex and index for index."""
    scene = json.load(open(os.path.join(sv.CBOX_DIR, "scene.json")))
    blob = open(os.path.join(sv.CBOX_DIR, "Scene.bin"), "rb").read()
    ref = akr.load_scene(os.path.join(sv.CBOX_DIR, "scene.json"))

    def meshes(s):
        d = s.desc.contents
        out = []
        for i in range(d.n_meshes):
            m = d.meshes[i]
            v = np.ctypeslib.as_array(C.cast(m.vertices, C.POINTER(C.c_float)), (m.n_vertices * 3,)).copy()
            t = np.ctypeslib.as_array(C.cast(m.indices, C.POINTER(C.c_uint32)), (m.n_triangles * 3,)).copy()
            out.append((v, t))
        return out
    want = meshes(ref)
    for kind, buf in (("base64", {"type": "base64", "data": base64.b64encode(blob).decode()}), ("binary", {"type": "binary", "data": list(blob)})):
        s2 = json.loads(json.dumps(scene))
        s2["buffers"]["Scene"] = buf
        got = meshes(akr.load_scene(_write(tmp_path, "buf_" + kind, s2)))
        assert len(got) == len(want)
        for (v, t), (v0, t0) in zip(got, want):
            assert np.array_equal(v, v0) and np.array_equal(t, t0), kind


def test_deeply_nested_json_and_cyclic_shader_graphs_are_rejected(akr, tmp_path):
    """Both used to overflow the stack of the (recursive) JSON parser / shader compiler: nesting beyond 128 levels is refused
    like serde_json's recursion limit does; a shader node that transitively feeds itself is refused (the reference's
    compiler has no visited set and would recurse until it crashes, compiler.rs:116-337)."""
    d = tmp_path / "deep"
    d.mkdir()
    for text in ("[" * 2_000_000, '{"a":' * 500_000, "[" * 129 + "]" * 129):
        (d / "scene.json").write_text(text)
        with pytest.raises(akr.AkariError, match="recursion limit"):
            akr.load_scene(str(d / "scene.json"))
    scene = json.load(open(os.path.join(sv.CBOX_DIR, "scene.json")))
    g = scene["materials"]["floor_001"]["shader"]
    p = sv._principled_name(g)
    g["nodes"]["a"] = {"type": "spectral_uplift", "rgb": {"id": "b"}}
    g["nodes"]["b"] = {"type": "spectral_uplift", "rgb": {"id": "a"}}
    g["nodes"][p]["base_color"] = {"id": "a"}
    with pytest.raises(akr.AkariError, match="cycle"):
        akr.load_scene(_write(tmp_path, "cyclic", scene))
    scene = json.load(open(os.path.join(sv.CBOX_DIR, "scene.json")))
    g = scene["materials"]["floor_001"]["shader"]
    g["nodes"][sv._principled_name(g)]["normal"] = {"id": sv._principled_name(g)}
    with pytest.raises(akr.AkariError, match="cycle"):
        akr.load_scene(_write(tmp_path, "self_cycle", scene))
    # a diamond (one node feeding two inputs) is not a cycle
    scene = json.load(open(os.path.join(sv.CBOX_DIR, "scene.json")))
    g = scene["materials"]["floor_001"]["shader"]
    p = sv._principled_name(g)
    g["nodes"][p]["coat_roughness"] = dict(g["nodes"][p]["roughness"])
    akr.load_scene(_write(tmp_path, "diamond", scene))


def test_error_messages_with_non_utf8_bytes_still_raise_akari_error(akr, tmp_path):
    """A corrupt file can put arbitrary bytes into the loader's error text; the binding must still raise AkariError."""
    scene = open(os.path.join(sv.CBOX_DIR, "scene.json"), "rb").read().replace(b'"perspective"', b'"persp\xff\xfective"')
    d = tmp_path / "bad_utf8"
    d.mkdir()
    (d / "scene.json").write_bytes(scene)
    shutil.copy(os.path.join(sv.CBOX_DIR, "Scene.bin"), str(d / "Scene.bin"))
    with pytest.raises(akr.AkariError):
        akr.load_scene(str(d / "scene.json"))
