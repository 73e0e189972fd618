"""Image-plane sharding across the GPUs of one box.

Every camera sample depends only on (absolute pixel, sample index, seed) and film pixels are written only by
their own pixel's samples (reference: pt.rs:1100 adds to the unshifted pixel; the filter is importance-sampled,
not splatted), so the frame splits into contiguous row bands, one rank per GPU, with the scene replicated and no
collective on the hot path.  One all_gather of the resolved RGB bands assembles the final HDR buffer.
"""
import torch


def row_bands(height, world):
    """[(y0, y1)] * world — contiguous, disjoint, covering [0, height); sizes differ by at most one row."""
    edges = [(height * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def max_band_rows(height, world):
    return max(y1 - y0 for y0, y1 in row_bands(height, world))


def gather_bands(local_rgb, height, width, rank, world, dist, out=None):
    """all_gather the per-rank [rows_r, width, 3] bands into one [height, width, 3] image on every rank.

    `local_rgb` must be a [max_band_rows, width, 3] tensor whose first rows_r rows hold this rank's band (padding
    rows are ignored); works with any backend (`nccl` on GPUs, `gloo` in the CPU tests)."""
    bands = row_bands(height, world)
    mr = max_band_rows(height, world)
    assert local_rgb.shape == (mr, width, 3), local_rgb.shape
    gathered = out if out is not None else torch.empty((world, mr, width, 3), dtype=local_rgb.dtype, device=local_rgb.device)
    if world > 1:
        dist.all_gather_into_tensor(gathered.view(world * mr, width, 3), local_rgb.contiguous())  # concatenation form: every backend accepts it
    else:
        gathered[0].copy_(local_rgb)
    return torch.cat([gathered[r, : bands[r][1] - bands[r][0]] for r in range(world)], dim=0)
