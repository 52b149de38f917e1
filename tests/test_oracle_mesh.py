"""CPU-only property checks of the meshing / carve definitions in the oracle (neither exists in the reference)."""
import numpy as np
import pytest

import scenes


def _dense(vol, n):
    """Dense bool[x,y,z] of an n^3-voxel oracle volume, from its canonical export."""
    occ, full = vol.occ(), vol.full()
    keys, payload = vol.export_partial()
    dims = vol.dims
    d = np.zeros((n, n, n), dtype=bool)
    pay = {int(k): p for k, p in zip(keys, payload)}
    for c in range(vol.nchunks):
        cx, cy, cz = c % dims[0], (c // dims[0]) % dims[1], c // (dims[0] * dims[1])
        ob = np.unpackbits(occ[c].view(np.uint8), bitorder="little")
        fb = np.unpackbits(full[c].view(np.uint8), bitorder="little")
        for b in np.nonzero(ob)[0]:
            bx, by, bz = cx * 16 + (b & 15), cy * 16 + ((b >> 4) & 15), cz * 16 + (b >> 8)
            if fb[b]:
                d[bx * 8:bx * 8 + 8, by * 8:by * 8 + 8, bz * 8:bz * 8 + 8] = True
            else:
                bits = np.unpackbits(pay[c * 4096 + int(b)].view(np.uint8), bitorder="little").reshape(8, 8, 8)  # [z,y,x]
                d[bx * 8:bx * 8 + 8, by * 8:by * 8 + 8, bz * 8:bz * 8 + 8] = bits.transpose(2, 1, 0).astype(bool)
    return d


def _exposed_faces(d):
    """Set of (x,y,z,face) unit faces: solid voxel whose neighbour across the face is empty (outside = empty)."""
    p = np.pad(d, 1)
    out = set()
    for face, (ax, s) in enumerate([(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]):
        nb = np.roll(p, -s, axis=ax)[1:-1, 1:-1, 1:-1]
        xs, ys, zs = np.nonzero(d & ~nb)
        out.update(zip(xs.tolist(), ys.tolist(), zs.tolist(), [face] * len(xs)))
    return out


def _expand(quads):
    out = set()
    x = quads["w0"] & 0xFFFF; y = quads["w0"] >> 16; z = quads["w1"] & 0xFFFF
    face = (quads["w1"] >> 16) & 7; w = (quads["w1"] >> 24) & 0xFF; h = quads["w2"]
    n = 0
    for xi, yi, zi, f, wi, hi in zip(x.tolist(), y.tolist(), z.tolist(), face.tolist(), w.tolist(), h.tolist()):
        ax = f >> 1
        for dv in range(hi):
            for du in range(wi):
                if ax == 0:
                    out.add((xi, yi + du, zi + dv, f))
                elif ax == 1:
                    out.add((xi + du, yi, zi + dv, f))
                else:
                    out.add((xi + du, yi + dv, zi, f))
                n += 1
    return out, n


@pytest.mark.parametrize("case", ["sphere_voxel", "clipped_sphere", "terrain_block"])
def test_quads_reexpand_to_exposed_faces(orc, case):
    if case == "sphere_voxel":
        origin, dims = (1, -1, -1), (1, 1, 1)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, (24.0, -8.0, -8.0, 6.5), granularity=orc.GRAN_VOXEL)
    elif case == "clipped_sphere":
        origin, dims, params = scenes.sphere_scene(128)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    else:
        origin, dims = (2, -1, 3), (1, 1, 1)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_BLOCK)
    quads = vol.mesh()
    faces, area = _expand(quads)
    exp = _exposed_faces(_dense(vol, 128))
    assert area == len(faces)            # quads never overlap
    assert faces == exp                  # and cover exactly the exposed faces
    assert vol.count_exposed_faces() == len(exp)
    assert np.all(((quads["w1"] >> 24) & 0xFF) <= 8) and np.all(quads["w2"] <= 8) and np.all(quads["w3"] == 0)
    # greedy merge really merges: far fewer quads than unit faces on these shapes
    assert len(quads) < len(exp)
    # canonical sort is idempotent and a permutation
    s = orc.sort_quads(quads)
    assert orc.sort_quads(s).tobytes() == s.tobytes() and len(s) == len(quads)


def test_mesh_bricks_is_a_partition_of_mesh(orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    occ = vol.occ()
    keys = []
    for c in range(vol.nchunks):
        bits = np.unpackbits(occ[c].view(np.uint8), bitorder="little")
        keys += [c * 4096 + int(b) for b in np.nonzero(bits)[0]]
    a = orc.sort_quads(vol.mesh())
    b = orc.sort_quads(vol.mesh_bricks(np.array(keys, dtype=np.uint64)))
    assert a.tobytes() == b.tobytes()


def test_carve_sphere_removes_exactly_the_voxels_inside(orc):
    origin, dims = (0, 0, 0), (1, 1, 1)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, (8.0, 8.0, 8.0, 7.0), granularity=orc.GRAN_VOXEL)
    before = _dense(vol, 128)
    center, radius = (40, 70, 64), 21
    dirty = vol.carve_sphere(center, radius)
    after = _dense(vol, 128)
    X, Y, Z = np.meshgrid(np.arange(128), np.arange(128), np.arange(128), indexing="ij")
    ins = ((2 * X + 1 - 2 * center[0]) ** 2 + (2 * Y + 1 - 2 * center[1]) ** 2 + (2 * Z + 1 - 2 * center[2]) ** 2) < (2 * radius) ** 2
    assert np.array_equal(after, before & ~ins)
    changed = before & ins
    bx, by, bz = np.nonzero(changed.reshape(16, 8, 16, 8, 16, 8).any(axis=(1, 3, 5)))
    exp_dirty = np.sort((bx + 16 * by + 256 * bz).astype(np.uint64))
    assert np.array_equal(dirty, exp_dirty)
    assert vol.count_voxels() == int(after.sum())
    # idempotent: carving the same sphere again changes nothing
    assert len(vol.carve_sphere(center, radius)) == 0
