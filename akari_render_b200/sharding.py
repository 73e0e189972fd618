"""Image-plane sharding across the GPUs of one box.

Every camera sample depends only on (absolute pixel, sample index, seed) and film pixels are written only by
their own pixel's samples (reference: pt.rs:1100 adds to the unshifted pixel; the filter is importance-sampled,
not splatted), so the frame splits by rows, one rank per GPU, with the scene replicated and no collective on the
hot path.  One all_gather of the resolved RGB rows assembles the final HDR buffer.

Rows are INTERLEAVED in small blocks (row y -> rank (y // block_rows) % world): contiguous bands leave the band that
holds the light source and the ceiling with 8 % more segments per sample on cbox, which capped round 1's 8-GPU
scaling at 0.89.  `AkrTile(y0, y1, block_rows, n_shards, shard)` expresses the split to the engine; each rank's film
holds its own rows packed in increasing y.
"""
import torch


def pick_block_rows(height, world, max_block=8):
    """Largest block height <= max_block that gives every rank the same number of rows; 1 when there is none."""
    for t in range(max_block, 0, -1):
        if height % t == 0 and (height // t) % world == 0:
            return t
    return 1


def interleaved_tile(height, world, rank, block_rows=None):
    """The AkrTile tuple (y0, y1, block_rows, n_shards, shard) of `rank`."""
    t = block_rows or pick_block_rows(height, world)
    return (0, height, t, world, rank)


def tile_row_indices(height, world, rank, block_rows=None):
    """Sensor rows of `rank`'s tile, in the order they are packed in its film."""
    t = block_rows or pick_block_rows(height, world)
    return [y for y in range(height) if (y // t) % world == rank]


def max_tile_rows(height, world, block_rows=None):
    return max(len(tile_row_indices(height, world, r, block_rows)) for r in range(world))


def row_bands(height, world):
    """[(y0, y1)] * world — contiguous, disjoint, covering [0, height); sizes differ by at most one row."""
    edges = [(height * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def max_band_rows(height, world):
    return max(y1 - y0 for y0, y1 in row_bands(height, world))


def gather_rows(local_rgb, height, width, rank, world, dist, block_rows=None, out=None, index=None):
    """all_gather the per-rank [rows_r, width, C] packed rows into one [height, width, C] image on every rank.

    `local_rgb` must be a [max_tile_rows, width, C] tensor whose first rows_r rows hold this rank's rows (padding rows
    are ignored); works with any backend (`nccl` on GPUs, `gloo` in the CPU tests).  `index` (from `gather_index`) can be
    passed in to keep its construction out of a timed region."""
    mr = max_tile_rows(height, world, block_rows)
    assert local_rgb.shape[0] == mr and local_rgb.shape[1] == width, local_rgb.shape
    c = local_rgb.shape[2]
    gathered = out if out is not None else torch.empty((world, mr, width, c), dtype=local_rgb.dtype, device=local_rgb.device)
    if world > 1:
        dist.all_gather_into_tensor(gathered.view(world * mr, width, c), local_rgb.contiguous())  # concatenation form: every backend accepts it
    else:
        gathered[0].copy_(local_rgb)
    if index is None:
        index = gather_index(height, world, block_rows, device=local_rgb.device)
    return gathered.view(world * mr, width, c).index_select(0, index)


def gather_index(height, world, block_rows=None, device="cpu"):
    """index[y] = position of sensor row y in the concatenated [world * max_tile_rows] gather buffer."""
    mr = max_tile_rows(height, world, block_rows)
    idx = torch.empty(height, dtype=torch.long)
    for r in range(world):
        for l, y in enumerate(tile_row_indices(height, world, r, block_rows)):
            idx[y] = r * mr + l
    return idx.to(device)


def gather_bands(local_rgb, height, width, rank, world, dist, out=None):
    """Contiguous-band variant (kept for hosts that shard by AkrTile(y0, y1))."""
    bands = row_bands(height, world)
    mr = max_band_rows(height, world)
    assert local_rgb.shape == (mr, width, 3), local_rgb.shape
    gathered = out if out is not None else torch.empty((world, mr, width, 3), dtype=local_rgb.dtype, device=local_rgb.device)
    if world > 1:
        dist.all_gather_into_tensor(gathered.view(world * mr, width, 3), local_rgb.contiguous())
    else:
        gathered[0].copy_(local_rgb)
    return torch.cat([gathered[r, : bands[r][1] - bands[r][0]] for r in range(world)], dim=0)
