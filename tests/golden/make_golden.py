#!/usr/bin/env python3
"""Regenerates the committed golden fixtures from the CPU oracle (run from the repo root).

The reference ships no golden images and cannot be executed here, so these fixtures pin
the ORACLE (regression) and give the GPU tests a fixed target that does not depend on re-running it.
  cbox_64x64_16spp.npy         film in the reference layout (7 * 64 * 64 f32)
  cbox_64x64_first_hits.npy    (instance, primitive) of the first hit of sample 0 per pixel
  cbox_256x256_16spp_rgb.npy   resolved RGB of BASELINE config C1 (256x256 @ 16 spp), float16-exact? no: f32
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import akari_render_b200 as akr  # noqa: E402
from oracle import binding as oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
pmj, bn = akr.sampler_tables()
task = akr.RenderTask.from_file(os.path.join(ROOT, "scenes", "cbox", "pt.json"))
task.pt.spp = 16

scene = akr.load_scene(os.path.join(ROOT, "scenes", "cbox", "scene.json")).set_resolution(64, 64)
film, st, fh = oracle.render(scene.desc, 64, 64, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
np.save(os.path.join(HERE, "cbox_64x64_16spp.npy"), film)
np.save(os.path.join(HERE, "cbox_64x64_first_hits.npy"), fh)
json.dump({"segments": int(st.segments), "shadow_rays": int(st.shadow_rays), "samples": int(st.samples)},
          open(os.path.join(HERE, "cbox_64x64_16spp.json"), "w"))

scene = akr.load_scene(os.path.join(ROOT, "scenes", "cbox", "scene.json")).set_resolution(256, 256)
film, st, fh = oracle.render(scene.desc, 256, 256, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
np.save(os.path.join(HERE, "cbox_256x256_16spp_rgb.npy"), oracle.resolve(film, 256 * 256).reshape(256, 256, 3))
np.save(os.path.join(HERE, "cbox_256x256_first_hits.npy"), fh.astype(np.uint8 if fh.max() < 256 else np.uint32))
json.dump({"segments": int(st.segments), "shadow_rays": int(st.shadow_rays), "samples": int(st.samples),
           "n_seg": st.segments / st.samples, "shadow_per_seg": st.shadow_rays / st.segments},
          open(os.path.join(HERE, "cbox_256x256_16spp.json"), "w"))
print("golden fixtures written")

# ---- digests of the oracle's film on the scene variants the round-2 features are tested on (32 x 32 @ 8 spp) -------------
# (a digest, not an array: these scenes exist to pin the oracle against accidental edits; the arrays above serve the GPU tests)
import hashlib  # noqa: E402
import tempfile  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
import scene_variants as sv  # noqa: E402

task.pt.spp = 8
digests = {}
tmp = tempfile.mkdtemp()
for name, path in (("principled_mix", sv.write_variant(tmp, "pm", sv.variant_principled_mix)), ("nodes", sv.write_variant(tmp, "nodes", sv.variant_nodes)),
                   ("textured", sv.write_textured(tmp, alpha_cutout=False)), ("textured_alpha", sv.write_textured(tmp, alpha_cutout=True)),
                   ("clutter", sv.write_clutter(tmp, n_lon=8, n_lat=6)), ("exr_textured", sv.write_exr_textured(tmp)[0])):
    scene = akr.load_scene(path).set_resolution(32, 32)
    film, st, fh = oracle.render(scene.desc, 32, 32, task.pt, task.sampler, task.filter, pmj, bn, want_first_hits=True)
    digests[name] = {"film_sha256": hashlib.sha256(np.ascontiguousarray(film).tobytes()).hexdigest(), "segments": int(st.segments),
                     "shadow_rays": int(st.shadow_rays)}
json.dump(digests, open(os.path.join(HERE, "variant_digests.json"), "w"), indent=1)
print("variant digests written")
