#!/usr/bin/env python3
"""Per-stage CUDA-event times of ONE 64-spp pass (kernel-tuning loop; much cheaper than bench.py).
usage: quick_stage_bench.py [cbox|clutter] [engine option=value ...]   (AKR_B200_CUDA_LIB selects a variant build)"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import akari_render_b200 as akr
which = sys.argv[1] if len(sys.argv) > 1 else "cbox"
opts = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
path = os.path.join(ROOT, "scenes", "cbox", "scene.json")
if which == "clutter":
    import scene_variants
    path = scene_variants.write_clutter(tempfile.mkdtemp())
scene = akr.load_scene(path).set_resolution(1280, 720)
task = akr.RenderTask.from_file(os.path.join(ROOT, "scenes", "cbox", "pt.json")); task.pt.spp = 1024
pt = akr.PathTracer(0)
opts.setdefault("wave_size", 1 << 26)
pt.set_engine_options(**opts); pt.upload_scene(scene)
for prof in (0, 1):
    pt.set_engine_options(profile_stages=prof, **opts); pt.reset_stats(); pt.begin(task); pt.render_pass(64, blocking=True)
st = pt.stats()
names = ["raygen", "trace", "shade_lambert", "shade_conductor", "accumulate", "misc", "shade_general"]
ms = {n: round(st.gpu_ms_kernel[i], 2) for i, n in enumerate(names) if st.gpu_ms_kernel[i] > 0}
print(os.path.basename(os.environ.get("AKR_B200_CUDA_LIB", "default")), which, opts, "total", round(sum(ms.values()), 2), ms, flush=True)
