"""N > 1 host logic on CPU: two `gloo` ranks render one row band each and all_gather the frame.

The per-rank renderer here is the host simulation of the kernels (test infrastructure; no GPU in this
container), the sharding + gather code is the product's (`akari_render_b200.sharding`, the same functions
`bench.py` runs over NCCL).  The assembled frame must equal the single-process frame bit for bit."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import akari_render_b200 as akr
from akari_render_b200.sharding import row_bands, max_band_rows, gather_bands
from test_hostsim_parity import run_hostsim
from oracle import binding as oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
W, H, SPP = 40, 27, 8   # odd height: bands of 13 and 14 rows
scene = akr.load_scene(os.path.join(sys.argv[1], "scenes", "cbox", "scene.json")).set_resolution(W, H)
task = akr.RenderTask.from_file(os.path.join(sys.argv[1], "scenes", "cbox", "pt.json")); task.pt.spp = SPP
lib = C.CDLL(os.path.join(sys.argv[1], "tests", "hostsim", "libhostsim.so")); lib.hostsim_last_error.restype = C.c_char_p
tables, table = akr.sampler_tables(), oracle.albedo_table()
y0, y1 = row_bands(H, world)[rank]
film, _, _ = run_hostsim(lib, scene, task, tables, table, W, H, y0=y0, y1=y1)
band = akr.Film(film, W, y1 - y0).to_rgb()
local = torch.zeros((max_band_rows(H, world), W, 3), dtype=torch.float32)
local[: y1 - y0] = torch.from_numpy(band)
img = gather_bands(local, H, W, rank, world, dist).numpy()
if rank == 0:
    np.save(sys.argv[2], img)
dist.barrier(); dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_row_bands_cover_the_frame():
    from akari_render_b200.sharding import row_bands
    for h in (1, 7, 720, 4096):
        for w in (1, 2, 3, 4, 8):
            b = row_bands(h, w)
            assert b[0][0] == 0 and b[-1][1] == h and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [y1 - y0 for y0, y1 in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_assemble_the_single_process_frame(tmp_path, akr, oracle, tables, cbox, cbox_task):
    out = str(tmp_path / "frame.npy")
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(worker), ROOT, out]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    img = np.load(out)
    # single process, whole frame, same renderer
    import ctypes as C
    from test_hostsim_parity import run_hostsim
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    lib.hostsim_last_error.restype = C.c_char_p
    w, h = 40, 27
    scene, task = cbox(w, h), cbox_task(8)
    film, _, _ = run_hostsim(lib, scene, task, tables, oracle.albedo_table(), w, h)
    ref = akr.Film(film, w, h).to_rgb()
    assert img.shape == ref.shape
    assert np.array_equal(img, ref)
