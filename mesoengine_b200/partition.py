"""Host-side description of the multi-GPU split, shared by bench.py and the tests.

Index arithmetic only -- the same mapping the kernels use (k_raymarch.cu: tile t belongs to rank t % world and is the
(t // world)-th tile of that rank; k_mesh.cu: chunk c belongs to rank c % world).  No compute happens here.
"""
import numpy as np

TILE_W, TILE_H = 32, 8
TILE_PX = TILE_W * TILE_H


def tile_grid(width, height):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def tiles_per_rank(width, height, world):
    tx, ty = tile_grid(width, height)
    return (tx * ty + world - 1) // world


def rank_tiles(width, height, rank, world):
    """Global tile indices rendered by `rank`, in the order they are packed."""
    tx, ty = tile_grid(width, height)
    return np.arange(rank, tx * ty, world, dtype=np.int64)


def tile_rect(width, height, tile):
    tx, _ = tile_grid(width, height)
    x0, y0 = (tile % tx) * TILE_W, (tile // tx) * TILE_H
    return x0, y0, min(x0 + TILE_W, width), min(y0 + TILE_H, height)


def pack_tiles(frame, rank, world):
    """Row-major frame (H, W) of records -> this rank's packed tiles (tiles_per_rank, TILE_H, TILE_W); what
    MESO_LAYOUT_TILES produces.  Pixels outside the frame (ragged edge tiles) stay zero."""
    h, w = frame.shape[:2]
    out = np.zeros((tiles_per_rank(w, h, world), TILE_H, TILE_W) + frame.shape[2:], dtype=frame.dtype)
    for j, t in enumerate(rank_tiles(w, h, rank, world)):
        x0, y0, x1, y1 = tile_rect(w, h, int(t))
        out[j, : y1 - y0, : x1 - x0] = frame[y0:y1, x0:x1]
    return out


def compose_tiles(gathered, width, height):
    """(world, tiles_per_rank, TILE_H, TILE_W) gathered packed tiles -> row-major frame; what
    meso_compose_tiles_device does on the GPU."""
    world = gathered.shape[0]
    frame = np.zeros((height, width) + gathered.shape[4:], dtype=gathered.dtype)
    tx, ty = tile_grid(width, height)
    for t in range(tx * ty):
        x0, y0, x1, y1 = tile_rect(width, height, t)
        frame[y0:y1, x0:x1] = gathered[t % world, t // world, : y1 - y0, : x1 - x0]
    return frame


def rank_chunks(n_chunks, rank, world):
    """Chunk indices meshed by `rank`."""
    return np.arange(rank, n_chunks, world, dtype=np.int64)
