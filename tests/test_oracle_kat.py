"""Known-answer tests that pin the CPU oracle (CPU-only).

The reference ships no tests or golden vectors, so the oracle is pinned three ways:
  * against fixtures parsed out of the reference's own sources (tests/golden/*.json, tools/gen_golden_from_reference.py);
  * against the hand-derivable facts listed in SURVEY.md section 4, evaluated here with independent numpy / pure-Python
    restatements of the cited formulas;
  * against closed forms (projection matrix entries, index formulas).
"""
import json
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


# ---- fixtures generated from the reference sources ---------------------------------------------------------------

def test_triplanar_faces_match_reference_table(orc):
    g = _golden("triplanar_faces.json")
    assert g["fan_indices"] == [0, 1, 2, 3, 4, 5, 6, 1]                      # TriplePlanarCube.h:36-43
    assert g["octant_bits"] == [["x", "1"], ["y", "2"], ["z", "4"]]          # SimpleVoxel.cpp:129-136
    for row in g["octants"]:
        assert sorted(orc.triplanar_faces(row["octant"])) == row["faces"]
        # the fan is anchored at the cube corner nearest the camera: sign + on axis i iff bit i of the octant id
        assert row["anchor_corner"] == [1 if (row["octant"] >> i) & 1 else -1 for i in range(3)]
    assert sorted(orc.triplanar_faces(0)) == [0, 2, 4]                       # row 0 = faces x-, y-, z-


def test_reference_constants(orc):
    g = _golden("constants.json")
    assert (g["BlockResolution"], g["ChunkResolution"], g["BlockSize"]) == (8, 16, 1.0)
    assert tuple(g["sphere"]) == orc.REF_SPHERE
    assert g["ChunkOccupancyDepth"] == 4 and g["ChunkInnerVoxelCullDepthThreshold"] == 1
    # Hash(p) = fract(sin(dot(p, k)) * m)  (VoxelMathHelper.h:30-33), libm sin
    kx, ky, kz, m = g["hash"]
    for p in [(1.0, 0.0, 0.0), (3.0, -7.0, 11.0), (-120.0, 45.0, 8.0)]:
        v = math.sin((p[0] * kx + p[1] * ky) + p[2] * kz) * m
        assert orc.hash3(*p, sin_mode=orc.SIN_LIBM) == v - math.floor(v)
    a, b, c, d = g["terrain"]
    assert (a, b, c, d) == (0.5, 0.1, 10.3, 0.4)
    assert g["camera_start"] == [5.0, 2.0, 2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    # erosion uses the full Moore neighbourhood, self excluded
    offs = {tuple(o) for o in g["offsets26"]}
    assert len(offs) == 26 and (0, 0, 0) not in offs and all(max(map(abs, o)) == 1 for o in offs)


# ---- a8: sphere generator ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("loc", [(3, 0, 0), (6, -1, 2), (9, 0, 0), (4, -3, -1), (0, 0, 0)])
def test_generate_sphere_chunk(orc, loc):
    """GeneratorHelper.h:120-150: block (X,Y,Z) of chunk loc is solid iff |p - (100,0,0)| - 50 < 0 at the block MIN corner,
    fp64; emission order X outer, Z inner."""
    got = orc.generate_chunk(orc.SDF_SPHERE, orc.REF_SPHERE, loc)
    X, Y, Z = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
    px = np.float64(loc[0]) * 1.0 * 16.0 + X * 1.0
    py = np.float64(loc[1]) * 1.0 * 16.0 + Y * 1.0
    pz = np.float64(loc[2]) * 1.0 * 16.0 + Z * 1.0
    d = np.sqrt(((px - 100.0) ** 2 + py ** 2) + pz ** 2) - 50.0
    exp = np.stack([X[d < 0], Y[d < 0], Z[d < 0]], axis=1).astype(np.uint8)  # C-order of 'ij' meshgrid = X outer, Z inner
    assert np.array_equal(got, exp)


def test_reference_sphere_block_count(orc):
    origin, dims = (2, -4, -4), (8, 8, 8)
    total = 0
    for cz in range(dims[2]):
        for cy in range(dims[1]):
            for cx in range(dims[0]):
                total += len(orc.generate_chunk(orc.SDF_SPHERE, orc.REF_SPHERE, (origin[0] + cx, origin[1] + cy, origin[2] + cz)))
    ax = np.arange(origin[0] * 16, (origin[0] + dims[0]) * 16, dtype=np.float64)
    ay = np.arange(origin[1] * 16, (origin[1] + dims[1]) * 16, dtype=np.float64)
    X, Y, Z = np.meshgrid(ax, ay, ay, indexing="ij")
    exp = int((np.sqrt(((X - 100.0) ** 2 + Y ** 2) + Z ** 2) - 50.0 < 0).sum())
    assert total == exp == 523155


# ---- a9: terrain noise (independent pure-Python restatement of GeneratorHelper.h:19-87) ------------------------------

def _py_hash(x, y, z):
    v = math.sin((x * 127.1 + y * 311.7) + z * 74.7) * 43758.5453123
    return v - math.floor(v)


def _py_noised(x):
    p = [math.floor(v) for v in x]
    w = [v - math.floor(v) for v in x]
    u = [((wi * wi) * wi) * ((wi * ((wi * 6.0) - 15.0)) + 10.0) for wi in w]
    du = [((30.0 * wi) * wi) * ((wi * (wi - 2.0)) + 1.0) for wi in w]
    a = _py_hash(p[0], p[1], p[2]); b = _py_hash(p[0] + 1, p[1], p[2]); c = _py_hash(p[0], p[1] + 1, p[2])
    d = _py_hash(p[0] + 1, p[1] + 1, p[2]); e = _py_hash(p[0], p[1], p[2] + 1); f = _py_hash(p[0] + 1, p[1], p[2] + 1)
    g = _py_hash(p[0], p[1] + 1, p[2] + 1); h = _py_hash(p[0] + 1, p[1] + 1, p[2] + 1)
    k0 = a; k1 = b - a; k2 = c - a; k3 = e - a
    k4 = a - b - c + d; k5 = a - c - e + g; k6 = a - b - e + f; k7 = -a + b + c - d + e - f - g + h
    val = -1.0 + 2.0 * (k0 + k1 * u[0] + k2 * u[1] + k3 * u[2] + k4 * u[0] * u[1] + k5 * u[1] * u[2] + k6 * u[2] * u[0] + k7 * u[0] * u[1] * u[2])
    gx = (2.0 * du[0]) * (k1 + k4 * u[1] + k6 * u[2] + k7 * u[1] * u[2])
    gy = (2.0 * du[1]) * (k2 + k5 * u[2] + k4 * u[0] + k7 * u[2] * u[0])
    gz = (2.0 * du[2]) * (k3 + k6 * u[0] + k5 * u[1] + k7 * u[0] * u[1])
    return [gx, gy, gz, val]


def _py_displacement(p_in):
    p = list(p_in); mgn = 0.5; d = 0.0; s = 1.0
    for i in range(5):
        rnd = _py_noised([p[0] + 10.0, p[1] + 10.0, p[2] + 10.0])
        d += rnd[3] * mgn
        for k in range(3):
            p[k] *= 2.0
            p[k] += (rnd[k] * 0.2) * s
        if i == 2:
            s *= -1.0
        mgn *= 0.5
    p = [v * 32.0 for v in p_in]
    for i in range(4):
        rnd = _py_noised(p)
        d += rnd[3] * mgn
        p = [v * 2.0 for v in p]
        mgn *= 0.5
    return d


def test_terrain_noise_bit_exact_vs_python_restatement(orc):
    rng = np.random.default_rng(7)
    for _ in range(40):
        q = rng.uniform(-40, 40, 3)
        assert np.array_equal(orc.noised(q, orc.SIN_LIBM), np.array(_py_noised(list(q))))
        assert orc.displacement(q, orc.SIN_LIBM) == _py_displacement(list(q))
        x, y, z = q * 10
        assert orc.sdf(orc.SDF_TERRAIN, None, x, y, z, orc.SIN_LIBM) == (y * .5 + _py_displacement([x * .1, y * .1, z * .1]) * 10.3) * .4


def test_displacement_is_bounded(orc):
    """The exact terrain cull of the voxeliser relies on |displacement| < 0.9981."""
    rng = np.random.default_rng(3)
    m = max(abs(orc.displacement(rng.uniform(-500, 500, 3), orc.SIN_PORTABLE)) for _ in range(4000))
    assert m < 0.9981


def test_portable_sin(orc):
    """The device-reproducible sin stays within 1 ulp-ish of libm over the argument range the terrain hash uses, and the
    two never disagree about a block's solidity on a sample of the terrain (SURVEY.md hard part 2)."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.linspace(-10, 10, 2001), rng.uniform(-1e7, 1e7, 20000)])
    assert max(abs(orc.sin_portable(x) - math.sin(x)) for x in xs) < 2.3e-16
    flips = 0
    for loc in [(0, 0, 0), (3, -1, 2), (-5, 0, 7)]:
        a = orc.generate_chunk(orc.SDF_TERRAIN, None, loc, sin_mode=orc.SIN_LIBM)
        b = orc.generate_chunk(orc.SDF_TERRAIN, None, loc, sin_mode=orc.SIN_PORTABLE)
        flips += 0 if np.array_equal(a, b) else 1
    assert flips == 0


# ---- a5-a7, a10: occupancy, erode mips, cull, instance records ------------------------------------------------------

def _np_mips(blocks, depth=4):
    occ = np.zeros((16, 16, 16), dtype=bool)  # [x,y,z]
    for x, y, z in blocks:
        occ[x, y, z] = True
    mips = [occ]
    for _ in range(1, depth):
        last = mips[-1]
        cur = np.zeros_like(occ)
        for x, y, z in blocks:
            if min(x, y, z) < 1 or max(x, y, z) > 14:
                continue
            nb = last[x - 1:x + 2, y - 1:y + 2, z - 1:z + 2].copy()
            nb[1, 1, 1] = True  # self excluded (BinaryOccupancyVolume.h:45-62)
            cur[x, y, z] = nb.all()
        mips.append(cur)
    return mips


def _bits(mask_words):
    """64 u64 words -> bool[x,y,z] with bit index x + 16 y + 256 z (VoxelMathHelper.h:73-76)."""
    b = np.unpackbits(np.ascontiguousarray(mask_words).view(np.uint8), bitorder="little").astype(bool)
    return b.reshape(16, 16, 16).transpose(2, 1, 0)  # flat index = z*256 + y*16 + x


@pytest.mark.parametrize("case", ["sphere", "random", "full", "single"])
def test_erode_mips_and_instances(orc, case):
    if case == "sphere":
        blocks = orc.generate_chunk(orc.SDF_SPHERE, orc.REF_SPHERE, (4, 0, 0))
    elif case == "random":
        rng = np.random.default_rng(5)
        sel = rng.random((16, 16, 16)) < 0.93
        X, Y, Z = np.nonzero(sel)
        blocks = np.stack([X, Y, Z], axis=1).astype(np.uint8)
    elif case == "full":
        X, Y, Z = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
        blocks = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.uint8)
    else:
        blocks = np.array([[7, 8, 9]], dtype=np.uint8)
    mips = orc.erode_mips(blocks, depth=4)
    exp = _np_mips([tuple(int(v) for v in b) for b in blocks])
    for d in range(4):
        assert np.array_equal(_bits(mips[d]), exp[d]), d
    inst = orc.emit_instances(blocks, mips, threshold=1, chunk_index=77, stamp=9)
    keep = [tuple(int(v) for v in b) for b in blocks if not exp[1][tuple(int(v) for v in b)]]
    assert len(inst) == len(keep)
    assert [tuple(r["BlockLocation"][:3]) for r in inst] == keep            # generator order is preserved
    assert np.all(inst["BlockLocation"][:, 3] == 255) and np.all(inst["ChunkIndex"] == 77) and np.all(inst["BlockFrameStamp"] == 9)
    if case == "full":
        assert len(inst) == 16 ** 3 - 14 ** 3                                # only the chunk shell survives the cull
    raw = inst.view(np.uint32).reshape(-1, 3)
    x, y, z = keep[0]
    assert int(raw[0, 1]) == x | (y << 8) | (z << 16) | (255 << 24)          # little-endian unpack of SimpleVoxel.cpp:79-85


# ---- a15: camera -----------------------------------------------------------------------------------------------------

def test_perspective_reverse_z_closed_form(orc):
    fov, a, n, f = math.radians(60.0), 1280.0 / 720.0, 0.1, 1000.0
    m = orc.perspective(fov, a, f, n).reshape(4, 4)  # reverse-Z: near/far swapped (VoxelCamera.cpp:15)
    t = math.tan(fov / 2)
    assert m[0][0] == pytest.approx(1 / (a * t), rel=1e-6) and m[1][1] == pytest.approx(1 / t, rel=1e-6)
    assert m[2][2] == pytest.approx(n / (f - n), rel=1e-6) and m[2][3] == -1.0
    assert m[3][2] == pytest.approx(f * n / (f - n), rel=1e-6)
    assert np.count_nonzero(m) == 5


def test_camera_uniform_reference_start_pose(orc):
    cam = orc.camera_uniform((5, 2, 2), (0, 0, 0), width=1280, height=720)
    v = cam["View"][0].reshape(4, 4)
    eye = -(v[:3, :3] @ v[3, :3])                      # -(R^T t) with glm's column-major storage
    assert np.allclose(eye, [5, 2, 2], atol=1e-5)
    fwd = -np.array([v[0][2], v[1][2], v[2][2]])
    assert np.allclose(fwd, -np.array([5, 2, 2]) / np.linalg.norm([5, 2, 2]), atol=1e-6)
    assert list(cam["CameraChunkLocation"][0]) == [0, 0, 0, 0] and list(cam["SubCameraLocation"][0]) == [5, 2, 2, 0]
    # re-centring (VoxelMathHelper.h:17-22): chunk = floor(pos/16), fract = pos - chunk*16
    cam = orc.camera_uniform((-3.5, 40.25, 16.0), (0, 0, 0), width=64, height=64)
    assert list(cam["CameraChunkLocation"][0][:3]) == [-1, 2, 1] and list(cam["SubCameraLocation"][0][:3]) == [12, 8, 0]


def test_fibonacci_sphere(orc):
    pts = orc.fibonacci_sphere(8)
    assert np.allclose(np.linalg.norm(pts, axis=1), 1.0)
    assert np.allclose(pts[0], [0, 1, 0]) and np.allclose(pts[-1], [0, -1, 0], atol=1e-12)


# ---- volume -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [128, 256])
def test_fast_sphere_classification_is_exact(orc, n):
    import scenes
    origin, dims, params = scenes.sphere_scene(n)
    a = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL, fast=False)
    b = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL, fast=True)
    assert np.array_equal(a.occ(), b.occ()) and np.array_equal(a.full(), b.full())
    (ka, pa), (kb, pb) = a.export_partial(), b.export_partial()
    assert np.array_equal(ka, kb) and np.array_equal(pa, pb)


def test_volume_export_import_roundtrip(orc):
    import scenes
    origin, dims, params = scenes.sphere_scene(256)
    a = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params)
    k, p = a.export_partial()
    b = orc.Volume(origin, dims).import_(a.occ(), a.full(), k, p)
    assert a.count_voxels() == b.count_voxels()
    rng = np.random.default_rng(1)
    for x, y, z in rng.integers(0, 256, (2000, 3)):
        assert a.get_voxel(int(x), int(y), int(z)) == b.get_voxel(int(x), int(y), int(z))
    # voxel-granular brick payload: bit (x + 8 y) of slice z  <=> sdf(min corner of the voxel) < 0
    for x, y, z in rng.integers(0, 256, (300, 3)):
        wx, wy, wz = origin[0] * 16 + x / 8.0, origin[1] * 16 + y / 8.0, origin[2] * 16 + z / 8.0
        assert a.get_voxel(int(x), int(y), int(z)) == int(orc.sdf(orc.SDF_SPHERE, params, wx, wy, wz) < 0)
