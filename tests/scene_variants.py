"""Builds variants of the cbox fixture (same geometry, edited shader-graph constants) in a temp dir, to
exercise the general Principled closure tree, glass / diffuse / emission nodes and multi-light scenes."""
import copy
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CBOX_DIR = os.path.join(ROOT, "scenes", "cbox")


def _set_input(graph, principled, name, value):
    nodes = graph["nodes"]
    nid = nodes[principled][name]["id"]
    node = nodes[nid]
    if node["type"] == "spectral_uplift":
        node = nodes[node["rgb"]["id"]]
    if node["type"] in ("float", "float3", "rgb"):
        node["value"] = value
    else:
        raise ValueError(node["type"])


def _principled_name(graph):
    for k, v in graph["nodes"].items():
        if v["type"] == "principled":
            return k
    raise KeyError("no principled node")


def edit_principled(scene, material, **inputs):
    g = scene["materials"][material]["shader"]
    p = _principled_name(g)
    for k, v in inputs.items():
        _set_input(g, p, k, v)


def replace_with_node(scene, material, node_type, **consts):
    """Replace the material by a single-closure graph: diffuse / glass / emission."""
    nodes = {}
    refs = {}
    for i, (k, v) in enumerate(consts.items()):
        if isinstance(v, (list, tuple)):
            nodes[f"c{i}"] = {"type": "rgb", "value": list(v), "colorspace": "srgb"}
            nodes[f"u{i}"] = {"type": "spectral_uplift", "rgb": {"id": f"c{i}"}}
            refs[k] = {"id": f"u{i}"}
        else:
            nodes[f"c{i}"] = {"type": "float", "value": float(v)}
            refs[k] = {"id": f"c{i}"}
    nodes["bsdf"] = dict({"type": node_type}, **refs)
    nodes["out"] = {"type": "output", "node": {"id": "bsdf"}}
    scene["materials"][material]["shader"] = {"nodes": nodes, "output": {"id": "out"}, "kind": "surface"}


def write_variant(tmpdir, name, edit):
    scene = json.load(open(os.path.join(CBOX_DIR, "scene.json")))
    scene = copy.deepcopy(scene)
    edit(scene)
    d = os.path.join(str(tmpdir), name)
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(CBOX_DIR, "Scene.bin"), os.path.join(d, "Scene.bin"))
    path = os.path.join(d, "scene.json")
    json.dump(scene, open(path, "w"))
    return path


def variant_principled_mix(scene):
    """Every lobe of the Principled tree gets exercised somewhere in the box."""
    edit_principled(scene, "floor_001", coat_weight=0.6, coat_roughness=0.05, coat_ior=1.5, coat_tint=[0.9, 0.95, 1.0])
    edit_principled(scene, "backWall_001", specular_ior_level=0.5, ior=1.5, roughness=0.3)
    edit_principled(scene, "shortBox_001", transmission_weight=1.0, ior=1.45, roughness=0.15, specular_ior_level=0.5)
    edit_principled(scene, "tallBox_001", metallic=0.5, roughness=0.25)
    edit_principled(scene, "leftWall_001", metallic=1.0, roughness=0.4, coat_weight=0.3, coat_roughness=0.1)
    edit_principled(scene, "rightWall_001", transmission_weight=0.4, ior=1.33, roughness=0.5, specular_ior_level=0.8,
                    specular_tint=[1.0, 0.8, 0.6])
    edit_principled(scene, "ceiling_001", emission_color=[0.2, 0.3, 0.9], emission_strength=0.5)


def variant_nodes(scene):
    """diffuse / glass / emission shader nodes instead of Principled."""
    replace_with_node(scene, "floor_001", "diffuse", color=[0.6, 0.6, 0.2])
    replace_with_node(scene, "shortBox_001", "glass", color=[0.95, 0.95, 1.0], ior=1.5, roughness=0.05)
    replace_with_node(scene, "backWall_001", "emission", color=[0.4, 0.1, 0.1], strength=2.0)


# ---- procedural clutter: a scene big enough that the BVH and the primitives live in global memory ------------------
def _uv_sphere(cx, cy, cz, r, n_lon, n_lat):
    """Indexed UV sphere with per-corner smooth normals and two material slots (alternating latitude bands)."""
    import numpy as np
    verts, idx, normals, uvs, slots = [], [], [], [], []
    for j in range(n_lat + 1):
        th = np.pi * j / n_lat
        for i in range(n_lon):
            ph = 2 * np.pi * i / n_lon
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            verts.append(np.array([cx, cy, cz]) + r * n)
    verts = np.array(verts, np.float32)
    c = np.array([cx, cy, cz], np.float32)

    def nrm(k):
        v = verts[k] - c
        return (v / np.linalg.norm(v)).astype(np.float32)

    for j in range(n_lat):
        for i in range(n_lon):
            a = j * n_lon + i
            b = j * n_lon + (i + 1) % n_lon
            d = (j + 1) * n_lon + i
            e = (j + 1) * n_lon + (i + 1) % n_lon
            tris = []
            if j > 0:
                tris.append((a, b, e))
            if j < n_lat - 1:
                tris.append((a, e, d))
            for t in tris:
                idx.append(t)
                normals.append([nrm(t[0]), nrm(t[1]), nrm(t[2])])
                uvs.append([[i / n_lon, j / n_lat], [(i + 1) / n_lon, j / n_lat], [(i + 1) / n_lon, (j + 1) / n_lat]])
                slots.append(j & 1)
    idx = np.array(idx, np.uint32)
    # per-corner tangents dp/dphi (mesh.rs:558-569); poles have a zero-length derivative -> use the x axis there
    tang = np.zeros((len(idx) * 3, 3), np.float32)
    for t, tri in enumerate(idx):
        for k in range(3):
            v = verts[tri[k]] - c
            tv = np.array([-v[2], 0.0, v[0]], np.float32)
            n2 = float(np.dot(tv, tv))
            tang[3 * t + k] = tv / np.sqrt(n2) if n2 > 1e-12 else np.array([1.0, 0.0, 0.0], np.float32)
    return (verts, idx, np.array(normals, np.float32).reshape(-1, 3), np.array(uvs, np.float32).reshape(-1, 2),
            np.array(slots, np.uint32), tang)


def write_clutter(tmp_path, n_lon=32, n_lat=24):
    """cbox + six tessellated spheres: smooth per-corner normals (one faceted), two materials each (glass-like, rough
    metal, Lambert), tangent buffers on two of them (one with non-finite entries -> fallback), one instance with a
    rotation * non-uniform scale, one mirrored.  ~8.5 K triangles => BVH nodes and primitives exceed the shared-memory
    staging budget (TRACE_BVH mode)."""
    import numpy as np
    scene = json.load(open(os.path.join(CBOX_DIR, "scene.json")))
    blob = bytearray(open(os.path.join(CBOX_DIR, "Scene.bin"), "rb").read())
    scene["buffers"]["Scene"]["path"] = "Scene.bin"
    # a glass-like and a rough-metal material derived from existing ones
    scene["materials"]["clutter_glass"] = copy.deepcopy(scene["materials"]["shortBox_001"])
    edit_principled(scene, "clutter_glass", transmission_weight=1.0, ior=1.45, roughness=0.1, specular_ior_level=0.5)
    scene["materials"]["clutter_metal"] = copy.deepcopy(scene["materials"]["tallBox_001"])
    edit_principled(scene, "clutter_metal", roughness=0.35)
    nview = len(scene["buffer_views"])

    def add_view(arr):
        nonlocal nview
        while len(blob) % 16:
            blob.append(0)
        name = f"buf_view_{nview}"
        nview += 1
        raw = np.ascontiguousarray(arr).tobytes()
        scene["buffer_views"][name] = {"buffer": {"id": "Scene"}, "offset": len(blob), "length": len(raw)}
        blob.extend(raw)
        return {"id": name}

    spheres = [(-0.55, 0.35, 0.45, 0.22), (0.45, 0.85, 0.35, 0.2), (0.0, 1.45, -0.3, 0.25), (-0.3, 1.55, 0.4, 0.15),
               (0.6, 0.25, 0.75, 0.18), (0.1, 0.9, 0.6, 0.12)]
    mats = [("floor_001", "clutter_metal"), ("clutter_glass", "backWall_001"), ("leftWall_001", "rightWall_001"),
            ("clutter_metal", "clutter_glass"), ("ceiling_001", "floor_001"), ("clutter_glass", "clutter_metal")]
    for k, (cx, cy, cz, r) in enumerate(spheres):
        v, i, n, uv, sl, tg = _uv_sphere(cx, cy, cz, r, n_lon, n_lat)
        gname = f"zz_sphere_{k}_mesh"
        tangents = None
        if k in (1, 2):  # tangent buffers (mesh.rs:558-569); sphere 2 has non-finite entries -> the dp/du fallback
            if k == 2:
                tg = tg.copy()
                tg[::7, 1] = np.nan
            tangents = add_view(tg)
        normals = None if k == 4 else add_view(n)  # one faceted sphere: geometric normals, uv-derived tangent frame
        scene["geometries"][gname] = {"type": "mesh", "vertices": add_view(v), "indices": add_view(i), "normals": normals,
                                      "uvs": add_view(uv), "tangents": tangents, "materials": add_view(sl)}
        xf = [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]
        if k == 3:  # rotation about z * non-uniform scale + translation: inverse-transpose normals, det area scaling (mesh.rs:608-628)
            ca, sa = float(np.cos(0.5)), float(np.sin(0.5))
            xf = [[1.2 * ca, -0.7 * sa, 0.0, 0.25], [1.2 * sa, 0.7 * ca, 0.0, -0.55], [0.0, 0.0, 0.9, 0.1], [0.0, 0.0, 0.0, 1.0]]
        if k == 5:  # mirrored instance (negative determinant)
            xf = [[-1.0, 0.0, 0.0, 0.2], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]
        scene["instances"][f"zz_sphere_{k}"] = {
            "geometry": {"id": gname},
            "transform": {"type": "matrix", "data": xf},
            "materials": [{"id": mats[k][0]}, {"id": mats[k][1]}]}
    scene["buffers"]["Scene"]["length"] = len(blob)
    d = os.path.join(str(tmp_path), "clutter")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "Scene.bin"), "wb").write(bytes(blob))
    path = os.path.join(d, "scene.json")
    json.dump(scene, open(path, "w"))
    return path


# ---- white furnace: a closed, uniformly emitting Lambert cube around the camera ----------------------------------------
def write_furnace(tmp_path, albedo=0.5, emission=1.0, half=2.0, textured=False):
    """Closed cube (inward-facing normals) centred on the cbox camera: every surface point emits `emission` and reflects
    Lambert `albedo`.  Radiance seen along any path of at most D bounces is emission * sum_{k=0..D} albedo^k."""
    import numpy as np
    scene = json.load(open(os.path.join(CBOX_DIR, "scene.json")))
    mat = copy.deepcopy(scene["materials"]["floor_001"])
    scene["materials"] = {"furnace": mat}
    edit_principled(scene, "furnace", base_color=[albedo] * 3, emission_color=[1.0, 1.0, 1.0], emission_strength=float(emission))
    c = np.array([0.0, 1.0, 9.0], np.float32)  # camera position in the scene's y-up frame
    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], np.float32) * half + c
    # faces as quads wound so that (v1 - v0) x (v2 - v0) points INTO the cube (emission is one-sided, light/area.rs:44-48)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for q in quads:
        a, b, d, e = q
        n = np.cross(v[b] - v[a], v[d] - v[a])
        if np.dot(n, c - v[a]) < 0:  # flip to face the centre
            b, e = e, b
        tris += [(a, b, d), (a, d, e)]
    idx = np.array(tris, np.uint32)
    uvs = np.tile(np.array([[0, 0], [1, 0], [1, 1]], np.float32), (len(tris), 1))
    blob = bytearray()
    views = {}

    def add_view(arr):
        while len(blob) % 16:
            blob.append(0)
        name = f"buf_view_{len(views)}"
        raw = np.ascontiguousarray(arr).tobytes()
        views[name] = {"buffer": {"id": "Scene"}, "offset": len(blob), "length": len(raw)}
        blob.extend(raw)
        return {"id": name}

    scene["geometries"] = {"furnace_mesh": {"type": "mesh", "vertices": add_view(v), "indices": add_view(idx), "normals": None, "uvs": add_view(uvs),
                                             "tangents": None, "materials": add_view(np.zeros(1, np.uint32))}}
    scene["instances"] = {"furnace": {"geometry": {"id": "furnace_mesh"},
                                      "transform": {"type": "matrix", "data": [[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]]},
                                      "materials": [{"id": "furnace"}]}}
    if textured:  # the same albedo from a constant-valued float image: a texture-driven material with the same analytic answer
        g = scene["materials"]["furnace"]["shader"]
        nodes, p = g["nodes"], _principled_name(g)
        tex = np.full((4, 4, 3), albedo, np.float32)
        nodes["tex"] = {"type": "image", "uv": None,
                        "image": {"data": add_view(tex), "format": "float", "colorspace": "none", "extension": "repeat", "interpolation": "linear",
                                  "width": 4, "height": 4, "channels": 3}}
        nodes["tex_up"] = {"type": "spectral_uplift", "rgb": {"id": "tex"}}
        nodes[p]["base_color"] = {"id": "tex_up"}
    scene["lights"] = {}
    scene["buffer_views"] = views
    scene["buffers"] = {"Scene": {"type": "path", "path": "Scene.bin", "length": len(blob)}}
    d = os.path.join(str(tmp_path), "furnace_textured" if textured else "furnace")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "Scene.bin"), "wb").write(bytes(blob))
    path = os.path.join(d, "scene.json")
    json.dump(scene, open(path, "w"))
    return path


# ---- texture-driven materials: image textures (raw float + png), mapping, checkerboard, normal map, separate colour, alpha ----
def _png_bytes(rgba):
    """Minimal PNG writer (8-bit RGBA, filter type 0..4 cycled per row so that the decoder's un-filtering is exercised)."""
    import struct
    import zlib
    import numpy as np
    h, w, _ = rgba.shape
    raw = bytearray()
    prev = np.zeros((w, 4), np.int32)
    for y in range(h):
        cur = rgba[y].astype(np.int32)
        ft = y % 5
        left = np.vstack([np.zeros((1, 4), np.int32), cur[:-1]])
        upleft = np.vstack([np.zeros((1, 4), np.int32), prev[:-1]])
        if ft == 0:
            enc = cur
        elif ft == 1:
            enc = cur - left
        elif ft == 2:
            enc = cur - prev
        elif ft == 3:
            enc = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            enc = cur - pred
        raw.append(ft)
        raw.extend((enc & 0xFF).astype(np.uint8).tobytes())
        prev = cur

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw), 6)) + chunk(b"IEND", b"")


def write_textured(tmp_path, alpha_cutout=True):
    """cbox with texture-driven Principled inputs on every wall and both boxes:
      floor      base colour = raw-float RGBA image (repeat, linear), mesh uvs
      backWall   base colour = sRGB png through texcoords -> extract(uv) -> mapping(point, scale 3, offset) (mirror, nearest)
      leftWall   base colour = checkerboard(scale 4) of two rgb colours
      rightWall  normal = normal_map(raw-float RGB image, strength 0.7), base colour constant
      shortBox   base colour = png with an alpha channel (0 / 0.5 / 1 patches): stochastic alpha-tested traversal (clip address mode)
      tallBox    roughness = extract(Red) of separate_color(float image), metallic 1
    """
    import numpy as np
    scene = json.load(open(os.path.join(CBOX_DIR, "scene.json")))
    blob = bytearray(open(os.path.join(CBOX_DIR, "Scene.bin"), "rb").read())
    scene["buffers"]["Scene"]["path"] = "Scene.bin"
    nview = len(scene["buffer_views"])

    def add_view(raw):
        nonlocal nview
        while len(blob) % 16:
            blob.append(0)
        name = f"buf_view_{nview}"
        nview += 1
        scene["buffer_views"][name] = {"buffer": {"id": "Scene"}, "offset": len(blob), "length": len(raw)}
        blob.extend(raw)
        return {"id": name}

    rng = np.random.default_rng(7)

    # the fixture's uv buffers are all zero except on the short box: give every mesh a planar per-face parametrisation
    # (drop the triangle's dominant normal axis, normalise by the mesh bounds) so that textures actually vary
    def view_array(ref, dtype, cols):
        v = scene["buffer_views"][ref["id"]]
        return np.frombuffer(bytes(blob[v["offset"]:v["offset"] + v["length"]]), dtype).reshape(-1, cols)
    for gname, g in scene["geometries"].items():
        verts, idx = view_array(g["vertices"], np.float32, 3), view_array(g["indices"], np.uint32, 3)
        lo, ext = verts.min(axis=0), np.maximum(verts.max(axis=0) - verts.min(axis=0), 1e-6)
        uvs = np.zeros((len(idx) * 3, 2), np.float32)
        for t, tri in enumerate(idx):
            n = np.abs(np.cross(verts[tri[1]] - verts[tri[0]], verts[tri[2]] - verts[tri[0]]))
            keep = [a for a in range(3) if a != int(np.argmax(n))]
            for k in range(3):
                q = (verts[tri[k]] - lo) / ext
                uvs[3 * t + k] = (q[keep[0]], q[keep[1]])
        g["uvs"] = add_view(uvs.tobytes())

    def image(arr, fmt, colorspace, extension, interpolation):
        h, w, c = arr.shape
        raw = np.ascontiguousarray(arr, np.float32).tobytes() if fmt == "float" else _png_bytes(arr)
        return {"data": add_view(raw), "format": fmt, "colorspace": colorspace, "extension": extension, "interpolation": interpolation,
                "width": w, "height": h, "channels": c}

    def graph_of(material):
        g = scene["materials"][material]["shader"]
        return g, g["nodes"], _principled_name(g)

    def rgb_const(nodes, name, value):
        nodes[name + "_rgb"] = {"type": "rgb", "value": list(value), "colorspace": "srgb"}
        nodes[name] = {"type": "spectral_uplift", "rgb": {"id": name + "_rgb"}}
        return {"id": name}

    # floor: raw float RGBA (3 channels + alpha 1), bilinear, repeat
    g, nodes, p = graph_of("floor_001")
    tex = (0.15 + 0.7 * rng.random((8, 8, 3))).astype(np.float32)
    nodes["tex"] = {"type": "image", "image": image(tex, "float", "none", "repeat", "linear"), "uv": None}
    nodes["tex_up"] = {"type": "spectral_uplift", "rgb": {"id": "tex"}}
    nodes[p]["base_color"] = {"id": "tex_up"}
    # backWall: sRGB png, nearest, mirror, uv through texcoords -> extract -> mapping
    g, nodes, p = graph_of("backWall_001")
    png = (rng.random((16, 12, 4)) * 255).astype(np.uint8)
    png[..., 3] = 255
    nodes["tc"] = {"type": "texcoords"}
    nodes["uv"] = {"type": "extract", "node": {"id": "tc"}, "field": "uv"}
    nodes["m_loc"] = {"type": "float3", "value": [0.25, -0.4, 0.0]}
    nodes["m_rot"] = {"type": "float3", "value": [0.0, 0.0, 0.0]}
    nodes["m_scale"] = {"type": "float3", "value": [3.0, 2.0, 1.0]}
    nodes["map"] = {"type": "mapping", "vector": {"id": "uv"}, "mapping": "point", "location": {"id": "m_loc"}, "rotation": {"id": "m_rot"},
                    "scale": {"id": "m_scale"}}
    nodes["tex"] = {"type": "image", "image": image(png, "png", "srgb", "mirror", "nearest"), "uv": {"id": "map"}}
    nodes["tex_up"] = {"type": "spectral_uplift", "rgb": {"id": "tex"}}
    nodes[p]["base_color"] = {"id": "tex_up"}
    # leftWall: checkerboard
    g, nodes, p = graph_of("leftWall_001")
    nodes["cb_scale"] = {"type": "float", "value": 4.0}
    nodes["cb"] = {"type": "checkerboard", "vector": None, "scale": {"id": "cb_scale"}, "color1": rgb_const(nodes, "cb1", [0.63, 0.065, 0.05]),
                   "color2": rgb_const(nodes, "cb2", [0.9, 0.9, 0.2])}
    nodes[p]["base_color"] = {"id": "cb"}
    # rightWall: normal map from a float image
    g, nodes, p = graph_of("rightWall_001")
    nm = np.zeros((6, 6, 3), np.float32)
    nm[..., 0] = 0.5 + 0.25 * rng.standard_normal((6, 6)).clip(-1, 1)
    nm[..., 1] = 0.5 + 0.25 * rng.standard_normal((6, 6)).clip(-1, 1)
    nm[..., 2] = 0.9
    nodes["nm_tex"] = {"type": "image", "image": image(nm, "float", "none", "extend", "linear"), "uv": None}
    nodes["nm_strength"] = {"type": "float", "value": 0.7}
    nodes["nm"] = {"type": "normal_map", "normal": {"id": "nm_tex"}, "strength": {"id": "nm_strength"}, "space": "tangent"}
    nodes[p]["normal"] = {"id": "nm"}
    # shortBox: png with alpha patches, clip
    if alpha_cutout:
        g, nodes, p = graph_of("shortBox_001")
        cut = np.zeros((8, 8, 4), np.uint8)
        cut[..., :3] = (rng.random((8, 8, 3)) * 200 + 40).astype(np.uint8)
        cut[..., 3] = rng.choice(np.array([0, 128, 255], np.uint8), size=(8, 8))
        nodes["tex"] = {"type": "image", "image": image(cut, "png", "srgb", "clip", "linear"), "uv": None}
        nodes["tex_up"] = {"type": "spectral_uplift", "rgb": {"id": "tex"}}
        nodes[p]["base_color"] = {"id": "tex_up"}
    # tallBox: roughness from the red channel of a float image
    g, nodes, p = graph_of("tallBox_001")
    rough = np.zeros((4, 4, 3), np.float32)
    rough[..., 0] = 0.1 + 0.5 * rng.random((4, 4))
    nodes["r_tex"] = {"type": "image", "image": image(rough, "float", "none", "repeat", "linear"), "uv": None}
    nodes["r_sep"] = {"type": "separate_color", "mode": "rgb", "color": {"id": "r_tex"}}
    nodes["r_red"] = {"type": "extract", "node": {"id": "r_sep"}, "field": "Red"}
    nodes[p]["roughness"] = {"id": "r_red"}
    scene["buffers"]["Scene"]["length"] = len(blob)
    d = os.path.join(str(tmp_path), "textured_alpha" if alpha_cutout else "textured")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "Scene.bin"), "wb").write(bytes(blob))
    path = os.path.join(d, "scene.json")
    json.dump(scene, open(path, "w"))
    return path


# ---- OpenEXR fixtures (single-part scanline; what decode_exr in scene_loader.cpp reads) -----------------------------
def _exr_bytes(arr, compression="zip", pixel_type="half", data_window_origin=(0, 0)):
    """Minimal OpenEXR writer: arr [h, w, 3 or 4] -> R, G, B(, A) channels (stored alphabetically: A, B, G, R), pixel type
    'half' / 'float', compression 'none' / 'rle' / 'zips' / 'zip', increasing-y line order."""
    import struct
    import zlib
    import numpy as np
    h, w, c = arr.shape
    names = ["R", "G", "B", "A"][:c]
    order = sorted(range(c), key=lambda k: names[k])
    dt = np.float16 if pixel_type == "half" else np.float32
    ptype = 1 if pixel_type == "half" else 2
    comp = {"none": 0, "rle": 1, "zips": 2, "zip": 3}[compression]
    lines_per_chunk = 16 if comp == 3 else 1
    x0, y0 = data_window_origin

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<I", len(data)) + data
    chlist = b"".join(names[k].encode() + b"\0" + struct.pack("<IB3xII", ptype, 0, 1, 1) for k in order) + b"\0"
    box = struct.pack("<4i", x0, y0, x0 + w - 1, y0 + h - 1)
    header = (struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([comp])) +
              attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0") +
              attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) +
              attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")

    def rle(b):
        out, i = bytearray(), 0
        while i < len(b):
            j = i
            while j + 1 < len(b) and b[j + 1] == b[i] and j - i < 126:
                j += 1
            if j - i >= 2:
                out += struct.pack("b", j - i) + bytes([b[i]])
                i = j + 1
            else:
                k = i
                while k < len(b) and k - i < 127 and not (k + 2 < len(b) and b[k] == b[k + 1] == b[k + 2]):
                    k += 1
                out += struct.pack("b", -(k - i)) + bytes(b[i:k])
                i = k
        return bytes(out)
    chunks = []
    for cy in range(0, h, lines_per_chunk):
        raw = b"".join(np.ascontiguousarray(arr[y, :, k], dt).tobytes() for y in range(cy, min(h, cy + lines_per_chunk)) for k in order)
        data = raw
        if comp != 0:
            t = np.frombuffer(raw, np.uint8)
            split = np.concatenate([t[0::2], t[1::2]]).astype(np.int32)  # even bytes, then odd bytes
            pred = split.copy()
            pred[1:] = (split[1:] - split[:-1] + 128 + 256) % 256          # byte predictor
            packed = rle(bytes(pred.astype(np.uint8))) if comp == 1 else zlib.compress(bytes(pred.astype(np.uint8)))
            if len(packed) < len(raw):
                data = packed
        chunks.append((y0 + cy, data))
    table_pos = len(header)
    pos = table_pos + 8 * len(chunks)
    offsets, body = [], b""
    for y, data in chunks:
        offsets.append(pos + len(body))
        body += struct.pack("<iI", y, len(data)) + data
    return header + b"".join(struct.pack("<Q", o) for o in offsets) + body


def write_image_textured(tmp_path, name, items, colorspace="none"):
    """cbox with planar uvs whose listed materials take their base colour from an encoded image:
    items = [(material, encoded bytes, format, width, height, channels)]."""
    import numpy as np
    scene = json.load(open(os.path.join(CBOX_DIR, "scene.json")))
    blob = bytearray(open(os.path.join(CBOX_DIR, "Scene.bin"), "rb").read())
    scene["buffers"]["Scene"]["path"] = "Scene.bin"
    nview = len(scene["buffer_views"])

    def add_view(raw):
        nonlocal nview
        while len(blob) % 16:
            blob.append(0)
        view = f"buf_view_{nview}"
        nview += 1
        scene["buffer_views"][view] = {"buffer": {"id": "Scene"}, "offset": len(blob), "length": len(raw)}
        blob.extend(raw)
        return {"id": view}
    for material, raw, fmt, w, h, c in items:
        g = scene["materials"][material]["shader"]
        nodes, p = g["nodes"], _principled_name(g)
        nodes["tex"] = {"type": "image", "uv": None,
                        "image": {"data": add_view(raw), "format": fmt, "colorspace": colorspace, "extension": "repeat", "interpolation": "linear",
                                  "width": w, "height": h, "channels": c}}
        nodes["tex_up"] = {"type": "spectral_uplift", "rgb": {"id": "tex"}}
        nodes[p]["base_color"] = {"id": "tex_up"}
    for gname, g in scene["geometries"].items():  # planar uvs as in write_textured
        v = scene["buffer_views"][g["vertices"]["id"]]
        verts = np.frombuffer(bytes(blob[v["offset"]:v["offset"] + v["length"]]), np.float32).reshape(-1, 3)
        v = scene["buffer_views"][g["indices"]["id"]]
        idx = np.frombuffer(bytes(blob[v["offset"]:v["offset"] + v["length"]]), np.uint32).reshape(-1, 3)
        lo, ext = verts.min(axis=0), np.maximum(verts.max(axis=0) - verts.min(axis=0), 1e-6)
        uvs = np.zeros((len(idx) * 3, 2), np.float32)
        for t, tri in enumerate(idx):
            n = np.abs(np.cross(verts[tri[1]] - verts[tri[0]], verts[tri[2]] - verts[tri[0]]))
            keep = [a for a in range(3) if a != int(np.argmax(n))]
            for k in range(3):
                q = (verts[tri[k]] - lo) / ext
                uvs[3 * t + k] = (q[keep[0]], q[keep[1]])
        g["uvs"] = add_view(uvs.tobytes())
    scene["buffers"]["Scene"]["length"] = len(blob)
    d = os.path.join(str(tmp_path), name)
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "Scene.bin"), "wb").write(bytes(blob))
    path = os.path.join(d, "scene.json")
    json.dump(scene, open(path, "w"))
    return path


def write_exr_textured(tmp_path):
    """cbox whose floor / back wall / left wall / right wall / ceiling base colours are OpenEXR textures in the supported
    encodings.  Returns (scene path, {material: source array})."""
    import numpy as np
    rng = np.random.default_rng(11)
    sources, items = {}, []
    specs = [("floor_001", (9, 7, 3), "zip", "half", (0, 0)), ("backWall_001", (20, 5, 4), "zip", "float", (3, -2)),
             ("leftWall_001", (4, 6, 3), "none", "float", (0, 0)), ("rightWall_001", (6, 6, 4), "zips", "half", (0, 0)),
             ("ceiling_001", (5, 8, 3), "rle", "half", (0, 0))]
    for material, shape, compression, pixel_type, origin in specs:
        arr = (0.1 + 0.8 * rng.random(shape)).astype(np.float16 if pixel_type == "half" else np.float32).astype(np.float32)
        if material == "ceiling_001":
            arr[:, :, :] = np.round(arr * 4) / 4  # long runs of equal bytes: the RLE path really compresses
            arr[2:, :, 1] = 0.5
        if shape[2] == 4:
            arr[..., 3] = 1.0
        sources[material] = arr
        items.append((material, _exr_bytes(arr, compression, pixel_type, origin), "exr", shape[1], shape[0], shape[2]))
    return write_image_textured(tmp_path, "exr_textured", items), sources
