#!/usr/bin/env python3
"""Samples/s of one 64-spp pass at 1280x720 on a test-suite scene variant:
  quick_scene_bench.py principled_mix|nodes|textured|textured_alpha|clutter|clutter_big [bvh]     (bvh: force the BVH / queued pipeline)"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import akari_render_b200 as akr
import scene_variants as sv
which = sys.argv[1]
trace_mode = 1 if len(sys.argv) > 2 and sys.argv[2] == "bvh" else 0
node_kb = int(os.environ.get("AKR_SMEM_NODE_KB", "0"))  # top-of-tree staging budget of the BVH kernels (0 = default 16 KiB)
d = tempfile.mkdtemp()
path = {"principled_mix": lambda: sv.write_variant(d, "pm", sv.variant_principled_mix), "nodes": lambda: sv.write_variant(d, "n", sv.variant_nodes),
        "textured": lambda: sv.write_textured(d, alpha_cutout=False), "textured_alpha": lambda: sv.write_textured(d, alpha_cutout=True),
        "clutter": lambda: sv.write_clutter(d), "clutter_big": lambda: sv.write_clutter(d, n_lon=160, n_lat=120)}[which]()
scene = akr.load_scene(path).set_resolution(1280, 720)
task = akr.RenderTask.from_file(os.path.join(ROOT, "scenes", "cbox", "pt.json")); task.pt.spp = 1024
pt = akr.PathTracer(0)
pt.set_engine_options(wave_size=1 << 26, trace_mode=trace_mode, smem_node_kb=node_kb); t0 = time.time(); pt.upload_scene(scene); t_upload = time.time() - t0
for prof in (0, 1):
    pt.set_engine_options(wave_size=1 << 26, profile_stages=prof, trace_mode=trace_mode, smem_node_kb=node_kb); pt.reset_stats(); pt.begin(task); pt.render_pass(64, blocking=True)
st = pt.stats()
names = ["raygen", "trace", "shade_lambert", "shade_conductor", "accumulate", "misc", "shade_general"]
ms = {n: round(st.gpu_ms_kernel[i], 2) for i, n in enumerate(names) if st.gpu_ms_kernel[i] > 0}
tot = sum(ms.values())
print(f"upload {t_upload:.2f} s;", os.path.basename(os.environ.get("AKR_B200_CUDA_LIB", "default")), which, "bvh" if trace_mode else "auto", f"node_kb {node_kb}", "total ms", round(tot, 2), "=> M samples/s", round(1280 * 720 * 64 / tot / 1e3, 1), ms, flush=True)
