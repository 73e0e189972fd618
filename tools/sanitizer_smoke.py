import sys, os, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import akari_render_b200 as akr
import scene_variants as sv
task = akr.RenderTask.from_file("scenes/cbox/pt.json"); task.pt.spp = 4; task.pt.spp_per_pass = 4
pt = akr.PathTracer(0)
pt.set_engine_options(aov_mask=1, wave_size=2048)
for name, path in (("cbox", "scenes/cbox/scene.json"), ("textured", sv.write_textured(tempfile.mkdtemp())), ("clutter", sv.write_clutter(tempfile.mkdtemp(), n_lon=8, n_lat=6)),
                   ("textured_opaque", sv.write_textured(tempfile.mkdtemp(), alpha_cutout=False)),
                   ("mix", sv.write_variant(tempfile.mkdtemp(), "pm", sv.variant_principled_mix))):
    scene = akr.load_scene(path).set_resolution(37, 23)
    film = pt.render(scene, task)
    film2 = pt.render(scene, task, tile=(0, 23, 4, 3, 1))
    print(name, float(film.to_rgb().mean()), film2.rows, pt.stats().segments)
    t = akr.RenderTask.from_json('{"method": {"type": "aov", "spp": 2, "aov": "roughness"}, "sampler": {"type": "pmj02bn", "seed": 0}, "film": {"filter": {"type": "box", "radius": 0.5}, "out": "a.exr"}}')
    print(" aov", float(pt.render_aov(scene, t).to_rgb().mean()))
pt.set_engine_options(fused=2, wave_size=2048)
scene = akr.load_scene("scenes/cbox/scene.json").set_resolution(37, 23)
print("queued", float(pt.render(scene, task).to_rgb().mean()))
pt.set_engine_options(trace_mode=1, wave_size=2048)  # BVH + queued pipeline with the general class ordered by k_sort_*
scene = akr.load_scene(sv.write_variant(tempfile.mkdtemp(), "pm", sv.variant_principled_mix)).set_resolution(37, 23)
print("queued general", float(pt.render(scene, task).to_rgb().mean()))
pt.close()
