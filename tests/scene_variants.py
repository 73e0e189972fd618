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
