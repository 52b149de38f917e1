"""Deterministic synthetic scenes and cameras shared by bench.py and the tests.

World unit = one block (BlockSize 1); a grid of N^3 voxels is (N/128)^3 chunks.  The reference generators take no RNG:
the "seed" of every scene is its formula.
"""
import math

import numpy as np

REF_SPHERE = (100.0, 0.0, 0.0, 50.0)  # Runtimes/Helper/GeneratorHelper.h:134


def sphere_scene(n_voxels):
    """V-sphere: the reference sphere scaled by s = N/1024.  N = 1024 is the reference sphere itself (centre (100,0,0),
    radius 50 blocks, fully inside the grid); other N keep radius 50 s and centre the sphere in the grid.
    Returns (origin_chunk, dims_chunks, params)."""
    c = n_voxels // 128
    s = n_voxels / 1024.0
    ox = int(round(100.0 * s / 16.0 - c / 2.0))
    cx = 100.0 if n_voxels == 1024 else (ox + c / 2.0) * 16.0
    return (ox, -c // 2, -c // 2), (c, c, c), (cx, 0.0, 0.0, 50.0 * s)


def terrain_scene(n_voxels, height_chunks=None):
    """V-terrain: TestGenerator SDF over an N x H x N grid centred on y = 0 (the surface crosses mid-volume)."""
    c = n_voxels // 128
    h = height_chunks or c
    return (0, -h // 2, 0), (c, h, c), None


def grid_center_world(origin, dims):
    return tuple((origin[i] + dims[i] / 2.0) * 16.0 for i in range(3))


def orbit_eyes(origin, dims, samples=8, radius_frac=0.75):
    """Eyes on a Fibonacci sphere (Runtimes/Helper/VoxelMathHelper.h:49-71) of radius radius_frac * N (in voxels ->
    world units /8) around the grid centre, each looking at the centre."""
    ctr = np.array(grid_center_world(origin, dims))
    n_world = max(dims) * 16.0
    phi = math.pi * (math.sqrt(5.0) - 1.0)
    eyes = []
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        r = math.sqrt(max(0.0, 1 - y * y))
        th = phi * i
        d = np.array([math.cos(th) * r, y, math.sin(th) * r])
        d = d / np.linalg.norm(d)
        # avoid looking exactly along the up axis (+Z): lookAt degenerates there
        if abs(d[2]) > 0.999:
            d = np.array([0.05, 0.02, math.copysign(1.0, d[2])])
            d = d / np.linalg.norm(d)
        eyes.append(tuple((ctr + d * radius_frac * n_world).tolist()))
    return eyes, tuple(ctr.tolist())
