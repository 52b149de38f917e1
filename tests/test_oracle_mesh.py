"""CPU-only property checks of the meshing / carve definitions in the oracle (neither exists in the reference)."""
import numpy as np
import pytest

import scenes


def _dense(vol, n):
    """Dense bool[x,y,z] of an n^3-voxel oracle volume, from its canonical export."""
    occ, full = vol.occ(), vol.full()
    keys, payload = vol.export_partial()
    dims = vol.dims
    d = np.zeros((n, n, n) if np.isscalar(n) else tuple(n), dtype=bool)
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
    level = (quads["w1"] >> 19) & 1
    wq, hq = (quads["w1"] >> 24) & 0xFF, quads["w2"]
    assert np.all(quads["w3"] == 0) and np.all((quads["w1"] >> 20) & 0xF == 0)
    assert np.all(wq[level == 0] <= 8) and np.all(hq[level == 0] <= 8)                   # voxel level: inside one brick
    assert np.all(wq[level == 1] % 8 == 0) and np.all(hq[level == 1] % 8 == 0)           # brick level: whole bricks ...
    assert np.all(wq[level == 1] <= 128) and np.all(hq[level == 1] <= 128)               # ... inside one chunk
    if case == "terrain_block":
        assert np.all(level == 1)        # a block-granular scene has only full bricks: everything merges at the brick level
    if case == "sphere_voxel":
        assert np.any(level == 0)
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
    b = orc.sort_quads(np.concatenate([vol.mesh_bricks(np.array(keys, dtype=np.uint64)), vol.mesh_chunk_faces(np.arange(vol.nchunks))]))
    assert a.tobytes() == b.tobytes()
    assert orc.sort_quads(vol.remesh(np.array(keys, dtype=np.uint64))).tobytes() == a.tobytes()


def test_brick_level_merge_of_flat_faces(orc):
    """A slab of 16 x 16 x 2 full bricks in one chunk: two 128 x 128 quads and four 128 x 16 ones -- not 256 + 256 + 4 x 32
    brick faces.  With a partial brick on top of one of them that brick's top face drops back to the voxel level."""
    dims = (1, 1, 1)
    occ = np.zeros((1, 64), dtype=np.uint64); full = np.zeros((1, 64), dtype=np.uint64)
    for z in (3, 4):
        occ[0, z * 4:z * 4 + 4] = ~np.uint64(0); full[0, z * 4:z * 4 + 4] = ~np.uint64(0)
    vol = orc.Volume((0, 0, 0), dims).import_(occ, full, np.zeros(0, np.uint64), np.zeros((0, 8), np.uint64))
    q = vol.mesh()
    assert len(q) == 6 and np.all((q["w1"] >> 19) & 1 == 1)
    area = (((q["w1"] >> 24) & 0xFF).astype(np.int64) * q["w2"]).sum()
    assert area == vol.count_exposed_faces() == 2 * 128 * 128 + 4 * 128 * 16
    top = q[((q["w1"] >> 16) & 7) == 5][0]
    assert (top["w0"], top["w1"] & 0xFFFF, (top["w1"] >> 24) & 0xFF, top["w2"]) == (0, 4 * 8 + 7, 128, 128)
    # a partial brick (one voxel) above brick (5, 6, 4)
    b = 5 + 16 * 6 + 256 * 5
    occ[0, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
    pay = np.zeros((1, 8), np.uint64); pay[0, 0] = 1
    vol2 = orc.Volume((0, 0, 0), dims).import_(occ, full, np.array([b], np.uint64), pay)
    q2 = vol2.mesh()
    faces, area2 = _expand(q2)
    assert area2 == len(faces) == vol2.count_exposed_faces()
    lvl0 = q2[(q2["w1"] >> 19) & 1 == 0]
    # the full brick below shows 63 of its 64 top faces at the voxel level; the lone voxel has 5 exposed faces
    below = lvl0[(lvl0["w1"] & 0xFFFF) == 4 * 8 + 7]
    assert (((below["w1"] >> 24) & 0xFF).astype(np.int64) * below["w2"]).sum() == 63
    assert len(lvl0) - len(below) == 5


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


def _expand_dirty(dirty, dims):
    """Dirty bricks + their six neighbours inside the grid (what meso_remesh_dirty re-meshes), as sorted keys."""
    out = set()
    nb = (dims[0] * 16, dims[1] * 16, dims[2] * 16)
    for key in dirty.tolist():
        c, b = key >> 12, key & 4095
        cx, cy, cz = c % dims[0], (c // dims[0]) % dims[1], c // (dims[0] * dims[1])
        bx, by, bz = cx * 16 + (b & 15), cy * 16 + ((b >> 4) & 15), cz * 16 + (b >> 8)
        for dx, dy, dz in ((0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            x, y, z = bx + dx, by + dy, bz + dz
            if 0 <= x < nb[0] and 0 <= y < nb[1] and 0 <= z < nb[2]:
                cc = (x >> 4) + dims[0] * ((y >> 4) + dims[1] * (z >> 4))
                out.add(cc * 4096 + (x & 15) + 16 * (y & 15) + 256 * (z & 15))
    return np.array(sorted(out), dtype=np.uint64)


@pytest.mark.parametrize("case", ["voxel_sphere", "block_terrain"])
def test_remesh_replacement_rule_keeps_the_list_current(orc, case):
    """The contract of the dirty re-mesh: drop the voxel-level quads whose corner lies in a re-meshed brick and the brick-level
    quads whose corner lies in a chunk that holds one, add what the re-mesh returns -- the result is the full mesh of the edited
    volume (no quad outside that set changes)."""
    if case == "voxel_sphere":
        origin, dims, params = scenes.sphere_scene(256)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
        carves = [((128 - 90, 128, 128), 24), ((128, 228, 138), 40), ((128, 128, 128), 30)]
    else:
        origin, dims = (0, -1, 0), (2, 2, 2)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_BLOCK)
        carves = [((100, 128, 100), 37), ((128, 120, 128), 60), ((10, 140, 250), 25)]
    current = vol.mesh()
    for center, radius in carves:
        dirty = vol.carve_sphere(center, radius)
        assert len(dirty) > 0
        keys = _expand_dirty(dirty, dims)
        fresh = vol.remesh(keys)
        x = (current["w0"] & 0xFFFF).astype(np.int64); y = (current["w0"] >> 16).astype(np.int64); z = (current["w1"] & 0xFFFF).astype(np.int64)
        level = (current["w1"] >> 19) & 1
        chunk = (x >> 7) + dims[0] * ((y >> 7) + dims[1] * (z >> 7))
        brick = chunk * 4096 + ((x >> 3) & 15) + 16 * ((y >> 3) & 15) + 256 * ((z >> 3) & 15)
        stale = np.where(level == 1, np.isin(chunk, np.unique(keys >> np.uint64(12)).astype(np.int64)), np.isin(brick, keys.astype(np.int64)))
        current = np.concatenate([current[~stale], fresh])
        assert orc.sort_quads(current.copy()).tobytes() == orc.sort_quads(vol.mesh()).tobytes()


def test_brick_level_across_chunk_and_grid_borders(orc):
    """Block-granular terrain over 2 x 2 x 1 chunks: every quad is brick-level, none crosses a chunk border, and together they
    cover exactly the exposed faces computed independently from the dense volume (grid border = empty outside)."""
    origin, dims = (0, -1, 0), (2, 2, 1)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_BLOCK)
    quads = vol.mesh()
    assert len(quads) > 0 and np.all((quads["w1"] >> 19) & 1 == 1)
    faces, area = _expand(quads)
    exp = _exposed_faces(_dense(vol, (256, 256, 128)))
    assert area == len(faces) and faces == exp and vol.count_exposed_faces() == len(exp)
    x = (quads["w0"] & 0xFFFF).astype(np.int64); y = (quads["w0"] >> 16).astype(np.int64); z = (quads["w1"] & 0xFFFF).astype(np.int64)
    f = (quads["w1"] >> 16) & 7; w = ((quads["w1"] >> 24) & 0xFF).astype(np.int64); h = quads["w2"].astype(np.int64)
    ax = f >> 1
    u0 = np.where(ax == 0, y, x); v0 = np.where(ax == 2, y, z)
    assert np.all((u0 >> 7) == ((u0 + w - 1) >> 7)) and np.all((v0 >> 7) == ((v0 + h - 1) >> 7))     # inside one chunk
    assert np.all(u0 % 8 == 0) and np.all(v0 % 8 == 0) and np.all(w % 8 == 0) and np.all(h % 8 == 0)
    # far fewer quads than brick faces
    assert len(quads) * 2 < (w // 8 * (h // 8)).sum()


def _greedy_loop(rows):
    """The definition (oracle/orc_mesh.c): rows ascending, lowest set bit first, run length, then as many following rows as hold
    the whole run."""
    r = list(rows)
    out = []
    for v in range(8):
        while r[v]:
            u0 = (r[v] & -r[v]).bit_length() - 1
            w = 0
            while u0 + w < 8 and (r[v] >> (u0 + w)) & 1:
                w += 1
            m = ((1 << w) - 1) << u0
            h = 1
            while v + h < 8 and (r[v + h] & m) == m:
                r[v + h] &= ~m
                h += 1
            r[v] &= ~m
            out.append((u0, v, w, h))
    return out


def _greedy_branch_free(img):
    """The step the mesh kernel runs once per lane and loop iteration (k_mesh.cu:take_quad), restated on a 64-bit integer: lowest
    set bit = corner; run length from its row; height = index of the first following row that misses a bit of the run (rows past
    the image shift in as zeros); the covered rows cleared in one go."""
    M64 = (1 << 64) - 1
    out = []
    while img:
        p = (img & -img).bit_length() - 1
        v, u0, sh = p >> 3, p & 7, p & ~7
        t = img >> sh
        row = t & 0xFF
        inv = (~(row >> u0)) & 0xFFFFFFFF
        w = (inv & -inv).bit_length() - 1
        m = ((1 << w) - 1) << u0
        mrep = (m * 0x0101010101010101) & M64
        miss = ~t & mrep & M64
        h = ((miss & -miss).bit_length() - 1) >> 3 if miss else 8
        clr = mrep if h == 8 else mrep & ((1 << (8 * h)) - 1)
        img &= ~(clr << sh) & M64
        out.append((u0, v, w, h))
    return out


def test_branch_free_greedy_step_equals_the_definition():
    rng = np.random.default_rng(3)
    images = [0xFFFFFFFFFFFFFFFF, 0x0000001818000000, 0xAA55AA55AA55AA55, 0x8000000000000001, 0x00FF00FF00FF00FF, 0xFF818181818181FF]
    images += [int(x) for x in rng.integers(0, 2 ** 63, size=3000, dtype=np.int64).astype(np.uint64)]
    # dense and sparse images exercise long runs / tall quads and isolated faces
    images += [int(a | b) for a, b in zip(rng.integers(0, 2 ** 63, size=1500, dtype=np.int64).astype(np.uint64), rng.integers(0, 2 ** 63, size=1500, dtype=np.int64).astype(np.uint64))]
    images += [int(a & b & c) for a, b, c in zip(*(rng.integers(0, 2 ** 63, size=1500, dtype=np.int64).astype(np.uint64) for _ in range(3)))]
    for img in images:
        rows = [(img >> (8 * v)) & 0xFF for v in range(8)]
        assert _greedy_branch_free(img) == _greedy_loop(rows), hex(img)


def test_word_shift_form_of_the_brick_level_exposure():
    """Pass C of the mesh (k_mesh.cu:mesh_chunk_faces_kernel) forms "full brick whose neighbour on the minus / plus side is absent"
    on whole 64-brick words (word = z*4 + y/4, four 16-bit x-rows) by shifting the neighbour's occupancy onto the brick's bit.
    The same arithmetic restated on Python integers against the per-brick definition, on random 3 x 3 x 3 chunk neighbourhoods."""
    M64 = (1 << 64) - 1
    X0, X15 = 0x0001000100010001, 0x8000800080008000
    rng = np.random.default_rng(9)
    for trial in range(6):
        dims = (3, 3, 3)
        occ = rng.integers(0, 2, size=(48, 48, 48)).astype(bool) if trial else np.ones((48, 48, 48), bool)
        if trial == 2:
            occ = rng.random((48, 48, 48)) < 0.9
        full = occ & (rng.random((48, 48, 48)) < 0.7)

        def word(a, cx, cy, cz, w):        # 64-brick word w of chunk (cx, cy, cz) of bit volume a, zero outside the grid
            if not (0 <= cx < 3 and 0 <= cy < 3 and 0 <= cz < 3):
                return 0
            z, yq = w >> 2, w & 3
            v = 0
            for j in range(4):
                for x in range(16):
                    if a[cx * 16 + x, cy * 16 + yq * 4 + j, cz * 16 + z]:
                        v |= 1 << (16 * j + x)
            return v

        cx = cy = cz = 1
        for w in rng.choice(64, size=12, replace=False).tolist() + [0, 3, 60, 63]:
            z, yq = w >> 2, w & 3
            F, me = word(full, cx, cy, cz, w), word(occ, cx, cy, cz, w)
            om = [((me << 1) & ~X0 & M64) | ((word(occ, cx - 1, cy, cz, w) >> 15) & X0),
                  ((me >> 1) & ~X15) | ((word(occ, cx + 1, cy, cz, w) << 15) & X15),
                  ((me << 16) & M64) | ((word(occ, cx, cy, cz, w - 1) if yq > 0 else word(occ, cx, cy - 1, cz, z * 4 + 3)) >> 48),
                  (me >> 16) | (((word(occ, cx, cy, cz, w + 1) if yq < 3 else word(occ, cx, cy + 1, cz, z * 4 + 0)) << 48) & M64),
                  word(occ, cx, cy, cz, w - 4) if z > 0 else word(occ, cx, cy, cz - 1, 15 * 4 + yq),
                  word(occ, cx, cy, cz, w + 4) if z < 15 else word(occ, cx, cy, cz + 1, 0 * 4 + yq)]
            for d, (dx, dy, dz) in enumerate(((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))):
                e = F & ~om[d] & M64
                for j in range(4):
                    for x in range(16):
                        bx, by, bz = 16 + x, 16 + yq * 4 + j, 16 + z
                        nb = occ[bx + dx, by + dy, bz + dz]      # (the centre chunk's neighbours are inside the 3^3 grid)
                        assert bool((e >> (16 * j + x)) & 1) == bool(full[bx, by, bz] and not nb), (trial, w, d, x, j)
