#!/usr/bin/env python3
"""Mutation fuzzer for the C-ABI trust boundary: a valid AkrSceneDesc is edited IN PLACE (SVM node ops / argument slots /
argument counts, shader refs, constant-blob bytes, material slots, vertex indices, geometry ids, flags, shrunken counts,
non-finite transforms) and handed to build_scene_blob — what akr_b200_upload_scene runs first — through the host
simulation library.  Every mutant must be rejected with an error code or build; never crash or read out of bounds.
(Array lengths are only ever shrunk: a count larger than its array is a caller contract violation no callee can detect.)

  python tools/fuzz_descriptor.py 1500 1
  # ASan + UBSan: AKR_B200_HOST_LIB / AKR_HOSTSIM_LIB = libraries built with -fsanitize=address,undefined (see
  # tools/fuzz_image_decoders.py); round 2: 7 500 mutants clean after one added null check (n_images without an array).
"""
import sys, os, tempfile, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import akari_render_b200._abi as abi
abi.HOST_LIB = os.environ.get("AKR_B200_HOST_LIB", abi.HOST_LIB)
import akari_render_b200 as akr
import scene_variants as sv
sim = C.CDLL(os.environ.get("AKR_HOSTSIM_LIB", os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500; SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 0; random.seed(SEED)
tmp = tempfile.mkdtemp()
paths = [os.path.join(sv.CBOX_DIR, "scene.json"), sv.write_textured(tmp), sv.write_clutter(tmp, n_lon=6, n_lat=4), sv.write_variant(tmp, "n", sv.variant_nodes)]
vals = [0, 1, 2, 3, 7, 15, 16, 25, 63, 64, 65, 255, 1000, 0x7fffffff, 0xffffffff, 0xfffffffe]
ok = rej = 0
for it in range(N):
    sc = akr.load_scene(paths[it % len(paths)])      # fresh copy every time: the descriptor is mutated in place
    d = sc.desc.contents
    for _ in range(random.randint(1, 3)):
        m = random.randrange(7)
        if m == 0:    # an SVM node's op / argument
            if not d.n_shader_kinds: continue
            k = d.shader_kinds[random.randrange(d.n_shader_kinds)]
            if k.n_nodes:
                n = k.nodes[random.randrange(k.n_nodes)]
                if random.random() < 0.3: n.op = random.choice(vals)
                elif random.random() < 0.3: n.n_args = random.choice([0, 1, 2, 4, 25, 26, 31, 32])
                else: n.a[random.randrange(abi.AKR_SVM_MAX_ARGS)] = random.choice(vals)
        elif m == 1:  # a shader ref of an instance
            inst = d.instances[random.randrange(d.n_instances)]
            if inst.n_materials:
                r = inst.materials[random.randrange(inst.n_materials)]
                if random.random() < 0.5: r.shader_kind = random.choice(vals)
                else: r.data_offset = random.choice(vals + [d.shader_data_size - 1, d.shader_data_size, d.shader_data_size - 3])
        elif m == 2:  # constant blob bytes
            if d.shader_data_size: d.shader_data[random.randrange(d.shader_data_size)] = random.randrange(256)
        elif m == 3:  # a material slot / an index
            mesh = d.meshes[random.randrange(d.n_meshes)]
            if mesh.n_material_slots and random.random() < 0.5: mesh.material_slots[random.randrange(mesh.n_material_slots)] = random.choice(vals)
            elif mesh.n_triangles: mesh.indices[random.randrange(mesh.n_triangles * 3)] = random.choice(vals)
        elif m == 4:  # instance geometry id / flags / material count
            inst = d.instances[random.randrange(d.n_instances)]
            c = random.randrange(3)
            if c == 0: inst.geom_id = random.choice(vals)
            elif c == 1: inst.flags = random.choice(vals)
            else: inst.n_materials = random.choice([0, min(1, inst.n_materials), inst.n_materials])
        elif m == 5:  # counts
            c = random.randrange(3)
            if c == 0: d.n_shader_kinds = random.choice([0, min(1, d.n_shader_kinds), d.n_shader_kinds])
            elif c == 1: d.shader_data_size = random.choice([0, 4, d.shader_data_size // 2, d.shader_data_size])
            else: d.n_images = random.choice([0, min(1, d.n_images), d.n_images])
        else:         # transform entries
            inst = d.instances[random.randrange(d.n_instances)]
            inst.transform[random.randrange(16)] = random.choice([0.0, float("nan"), float("inf"), 1e30, -1.0])
    arrs = [(C.c_uint32 * 64)() for _ in range(4)]
    rc = sim.hostsim_material_keys(sc.desc, 64, *arrs)
    if rc < 0: rej += 1
    else: ok += 1
print("descriptor fuzz done: built", ok, "rejected", rej)
