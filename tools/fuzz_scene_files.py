#!/usr/bin/env python3
"""Mutation fuzzer for the scene front end: (a) structural edits of scene.json (wrong types, missing keys, dangling ids, huge
offsets), (b) value edits of Scene.bin (NaN / inf / huge vertices, out-of-range indices).  A mutant must be rejected with
AkariError by the loader, rejected by the scene build (akr_b200_upload_scene's build_scene_blob, reached here through the
host simulation library), or build — never crash or read out of bounds.

  python tools/fuzz_scene_files.py 1000 1
  # under ASan + UBSan: build the two libraries with -fsanitize=address,undefined (see tools/fuzz_image_decoders.py) and set
  # AKR_B200_HOST_LIB / AKR_HOSTSIM_LIB; round 2 ran 4 500 structural + 2 400 value mutants clean after fixing a wrapping
  # `offset + length` bound in the buffer-view check.
"""
import copy
import ctypes as C
import json
import os
import random
import struct
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import akari_render_b200._abi as abi  # noqa: E402
if os.environ.get("AKR_B200_HOST_LIB"):
    abi.HOST_LIB = os.environ["AKR_B200_HOST_LIB"]
import akari_render_b200 as akr  # noqa: E402
import scene_variants as sv  # noqa: E402

WEIRD = [0, -1, 1, 2**31, 2**32 - 1, 2**40, -2**31, 1e30, -1e30, 0.5, "", "x", None, [], {}, [1, 2, 3], True]
SPECIAL = [struct.pack("<f", v) for v in (float("nan"), float("inf"), -float("inf"), 1e38, -1e38, 0.0, -0.0, 1e-45)] + \
          [struct.pack("<I", v) for v in (0xffffffff, 0x7fffffff, 0, 1 << 30)]


def _paths(o, p=()):
    if isinstance(o, dict):
        for k, v in o.items():
            yield p + (k,)
            yield from _paths(v, p + (k,))
    elif isinstance(o, list):
        for i, v in enumerate(o[:8]):
            yield p + (i,)
            yield from _paths(v, p + (i,))


def _mutate_json(s):
    allp = list(_paths(s))
    for _ in range(random.randint(1, 3)):
        p = random.choice(allp)
        try:
            o = s
            for k in p[:-1]:
                o = o[k]
            cur = o[p[-1]]
            m = random.randrange(4)
            if m == 1:
                del o[p[-1]]
            elif m == 2 and isinstance(cur, (int, float)) and not isinstance(cur, bool):
                o[p[-1]] = cur * random.choice([-1, 0, 2, 1000, 1e9]) + random.choice([0, 1, -1])
            elif m == 3 and isinstance(cur, dict) and "id" in cur:
                o[p[-1]] = {"id": random.choice(["nope", "buf_view_0", "buf_view_3", ""])}
            else:
                o[p[-1]] = random.choice(WEIRD)
        except (KeyError, IndexError, TypeError):
            pass


def _mutate_blob(blob):
    for _ in range(random.randint(1, 6)):
        i = random.randrange(0, len(blob) - 4) & ~3
        if random.random() < 0.6:
            blob[i:i + 4] = random.choice(SPECIAL)
        else:
            blob[i + random.randrange(4)] = random.randrange(256)


def run(n, seed, tmp=None):
    random.seed(seed)
    tmp = tmp or tempfile.mkdtemp()
    sim = C.CDLL(os.environ.get("AKR_HOSTSIM_LIB", os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")))
    srcs = [sv.CBOX_DIR, os.path.dirname(sv.write_textured(tmp)), os.path.dirname(sv.write_clutter(tmp, n_lon=6, n_lat=4))]
    built = build_rejected = load_rejected = 0
    d = os.path.join(str(tmp), "mutant")
    os.makedirs(d, exist_ok=True)
    for it in range(n):
        src = srcs[it % len(srcs)]
        scene = json.load(open(os.path.join(src, "scene.json")))
        blob = bytearray(open(os.path.join(src, "Scene.bin"), "rb").read())
        if it & 1:
            _mutate_json(scene)
        else:
            _mutate_blob(blob)
        open(os.path.join(d, "Scene.bin"), "wb").write(bytes(blob))
        json.dump(scene, open(os.path.join(d, "scene.json"), "w"))
        try:
            sc = akr.load_scene(os.path.join(d, "scene.json"))
        except akr.AkariError:
            load_rejected += 1
            continue
        arrs = [(C.c_uint32 * 64)() for _ in range(4)]
        if sim.hostsim_material_keys(sc.desc, 64, *arrs) < 0:  # runs build_scene_blob on the descriptor
            build_rejected += 1
        else:
            built += 1
    return built, build_rejected, load_rejected


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print("fuzz done: built %d, rejected by the scene build %d, rejected by the loader %d" % run(n, seed))
