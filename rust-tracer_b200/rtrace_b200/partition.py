"""Row partition of one frame over the ranks of a job, and its inverse.

The reference shards a frame into 64x64 buckets on a thread pool (render.rs:273-298).
Across GPUs the unit is a row: rank g renders rows g, g+G, g+2G, ... (interleaved,
so every rank gets the same mix of empty and dense rows of the sphere flake), packed
densely into a band of ceil(H/G) rows; rank 0 gathers the bands and de-interleaves.
Pure index arithmetic -- used by bench.py on NCCL and by the gloo test on CPU.
"""


def band_rows(height, rank, world):
    """Number of image rows rank `rank` renders."""
    return (height - rank + world - 1) // world if height > rank else 0


def band_capacity(height, world):
    """Rows of the (padded) band buffer every rank allocates: equal sizes for the gather."""
    return (height + world - 1) // world


def band_spec(height, rank, world):
    """(row_start, row_stride, row_count) for Renderer.render_rows."""
    return rank, world, band_rows(height, rank, world)


def deinterleave(bands, height):
    """bands[g][r] is image row r*G + g.  Works on numpy arrays and torch tensors."""
    world = len(bands)
    cap = bands[0].shape[0]
    if hasattr(bands[0], "new_empty"):   # torch
        import torch
        frame = torch.stack(list(bands), dim=1).reshape((cap * world,) + tuple(bands[0].shape[1:]))
    else:
        import numpy as np
        frame = np.stack(list(bands), axis=1).reshape((cap * world,) + tuple(bands[0].shape[1:]))
    return frame[:height]


def block_band_spec(height, rank, world, block=16):
    """(row_start, row_stride, row_block, row_count) for Renderer.render_row_blocks: rank `rank` renders
    blocks of `block` consecutive rows, `world` blocks apart, so whole tiles stay contiguous in the image."""
    start, stride = rank * block, world * block
    count = sum(min(block, height - y0) for y0 in range(start, height, stride))
    return start, stride, block, count


def block_band_rows(height, rank, world, block=16):
    """The image rows of that band, in local-row order."""
    start, stride = rank * block, world * block
    return [y for y0 in range(start, height, stride) for y in range(y0, min(y0 + block, height))]
