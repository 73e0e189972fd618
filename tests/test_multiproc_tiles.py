"""N > 1 host logic on CPU: two `gloo` ranks render their interleaved row blocks and all_gather the frame.

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
from akari_render_b200.sharding import interleaved_tile, max_tile_rows, gather_rows, tile_row_indices
from test_hostsim_parity import run_hostsim
from oracle import binding as oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
W, H, SPP = 40, 27, 8   # odd height, 4-row blocks: rank 0 owns 15 rows (4 + 4 + 4 + 3), rank 1 owns 12
scene = akr.load_scene(os.path.join(sys.argv[1], "scenes", "cbox", "scene.json")).set_resolution(W, H)
task = akr.RenderTask.from_file(os.path.join(sys.argv[1], "scenes", "cbox", "pt.json")); task.pt.spp = SPP
lib = C.CDLL(os.path.join(sys.argv[1], "tests", "hostsim", "libhostsim.so")); lib.hostsim_last_error.restype = C.c_char_p
tables, table = akr.sampler_tables(), oracle.albedo_table()
BLOCK = 4
y0, y1, block, shards, shard = interleaved_tile(H, world, rank, BLOCK)
rows = tile_row_indices(H, world, rank, BLOCK)
lib.hostsim_set_tile_interleave(block, shards, shard)
film, _, _ = run_hostsim(lib, scene, task, tables, table, W, H, y0=y0, y1=y1, n_rows=len(rows))
band = akr.Film(film, W, len(rows)).to_rgb()
local = torch.zeros((max_tile_rows(H, world, BLOCK), W, 3), dtype=torch.float32)
local[: len(rows)] = torch.from_numpy(band)
img = gather_rows(local, H, W, rank, world, dist, block_rows=BLOCK).numpy()
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


def test_interleaved_tiles_partition_the_frame(akr):
    """Every row belongs to exactly one shard, the engine's row count (akr_b200_tile_rows) agrees with the host-side
    index lists, the gather index is a permutation, and the block height bench.py picks balances 720 and 4096 rows
    exactly over 1 / 2 / 4 / 8 ranks."""
    import ctypes as C
    import torch
    from akari_render_b200 import _abi
    from akari_render_b200.sharding import gather_index, interleaved_tile, max_tile_rows, pick_block_rows, tile_row_indices
    lib = _abi.load_cuda_lib()
    for h in (1, 7, 27, 720, 4096):
        for world in (1, 2, 3, 4, 8):
            for block in (None, 1, 4, 5):
                seen = []
                for r in range(world):
                    rows = tile_row_indices(h, world, r, block)
                    t = _abi.AkrTile(*interleaved_tile(h, world, r, block), 0)
                    assert lib.akr_b200_tile_rows(C.byref(t)) == len(rows), (h, world, r, block)
                    seen += rows
                assert sorted(seen) == list(range(h))
                idx = gather_index(h, world, block)
                mr = max_tile_rows(h, world, block)
                flat = torch.full((world * mr,), -1, dtype=torch.long)
                for r in range(world):
                    rows = tile_row_indices(h, world, r, block)
                    flat[r * mr: r * mr + len(rows)] = torch.tensor(rows, dtype=torch.long)
                assert torch.equal(flat[idx], torch.arange(h))
    for h in (720, 4096):
        for world in (1, 2, 4, 8):
            sizes = [len(tile_row_indices(h, world, r)) for r in range(world)]
            assert max(sizes) == min(sizes), (h, world, pick_block_rows(h, world), sizes)


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
