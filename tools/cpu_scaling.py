"""Thread scaling of the CPU oracle on this host (how many threads the CPU baseline should use)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import akari_render_b200 as akr
from oracle import binding as oracle
scene = akr.load_scene(ROOT + "/scenes/cbox/scene.json").set_resolution(1280, 720)
task = akr.RenderTask.from_file(ROOT + "/scenes/cbox/pt.json"); task.pt.spp = 1024
pmj, bn = akr.sampler_tables()
for t in (1, 8, 16, 32, 64, 128, 256):
    _, st, _ = oracle.render(scene.desc, 1280, 720, task.pt, task.sampler, task.filter, pmj, bn, spp_begin=0, spp_end=2, threads=t)
    print(t, "threads:", round(st.samples / st.seconds / 1e6, 3), "M samples/s", flush=True)
