"""K6 on the GPU through the C ABI: resident-set selection, importance, streaming generation -- bit-exact against
oracle/orc_resident.c and the oracle voxeliser."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from mesoengine_b200 import capi as _capi
    return _capi


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


FORWARDS = [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (0.3, -0.8, 0.52), (-2.0, 1.0, 0.5), (0.0, 0.0, 0.0), (1e-3, -1.0, 1e-3)]


@pytest.mark.parametrize("fwd", FORWARDS)
def test_select_view_chunks_default_config(ctx, capi, orc, fwd):
    """Reference defaults (radius 24 / 6, 120 degrees): same candidates, same importance bits, same order."""
    got = ctx.select_view_chunks(fwd)
    want = orc.select_view_chunks(fwd)
    assert got.shape == want.shape
    assert np.array_equal(got["Offset"], want["Offset"])
    assert np.array_equal(got["Importance"].view(np.uint32), want["Importance"].view(np.uint32))


@pytest.mark.parametrize("cfg", [(1, 1, 120.0, 0), (2, 1, 60.0, 0), (7, 3, 90.0, 0), (12, 12, 170.0, 0), (31, 4, 10.0, 0), (24, 6, 120.0, 1), (5, 0, 120.0, 1)])
def test_select_view_chunks_configs(ctx, capi, orc, cfg):
    F, B, angle, mode = cfg
    fwd = (0.6, 0.2, -0.77)
    got = ctx.select_view_chunks(fwd, capi.view_config(F, B, angle, mode))
    want = orc.select_view_chunks(fwd, F, B, angle, mode)
    assert got.shape == want.shape and got.shape[0] >= 7     # F = 1: the radius test leaves the 6-neighbourhood + centre
    assert np.array_equal(got["Offset"], want["Offset"])
    assert np.array_equal(got["Importance"].view(np.uint32), want["Importance"].view(np.uint32))


def test_select_with_baked_direction_reproduces_reference_table(ctx, capi, orc):
    """The reference answers with the table baked for the nearest of 256 Fibonacci directions (ChunkManager.h:106-124)."""
    fwd = (0.2, 0.1, 0.97)
    d, idx = capi.baked_direction(256, fwd)
    dirs = orc.fibonacci_sphere_f32(256)
    assert idx == orc.nearest_direction(dirs, fwd)
    assert np.array_equal(d.view(np.uint32), dirs[idx].view(np.uint32))
    got = ctx.select_view_chunks(d)
    want = orc.select_view_chunks(dirs[idx])
    assert np.array_equal(got["Offset"], want["Offset"]) and np.array_equal(got["Importance"].view(np.uint32), want["Importance"].view(np.uint32))


def test_chunk_importance(ctx, orc):
    rng = np.random.default_rng(11)
    loc = rng.integers(-80, 80, size=(5000, 3)).astype(np.int32)
    loc[:200] = rng.integers(-3, 4, size=(200, 3))          # the +-2 cube and its rim
    cam = (3, -2, 5)
    loc[:200] += np.array(cam, dtype=np.int32)
    loc[200] = cam
    fwd = (0.3, -0.8, 0.52)
    got = ctx.chunk_importance(cam, fwd, loc)
    want = np.array([orc.chunk_importance(cam, fwd, l) for l in loc], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert ctx.chunk_importance(cam, fwd, np.zeros((0, 3), dtype=np.int32)).shape == (0,)


def _expected_update(orc, loaded, origin, dims, cam, fwd, view, max_new):
    """UpdateLoadingQueue's dispatch loop over the oracle's sorted desired set (ChunkManager.h:229-283)."""
    cand = orc.select_view_chunks(fwd, *view)
    new = []
    missing = 0
    in_window = 0
    for off in cand["Offset"]:
        p = off + np.array(cam) - np.array(origin)
        if np.any(p < 0) or np.any(p >= np.array(dims)):
            continue
        in_window += 1
        slot = int(p[0] + dims[0] * (p[1] + dims[1] * p[2]))
        if loaded[slot]:
            continue
        if len(new) < max_new:
            new.append(slot)
        else:
            missing += 1
    for s in new:
        loaded[s] = True
    return new, missing, cand.shape[0], in_window


def _masked(orc, full_vol, origin, dims, loaded):
    """The fully generated oracle volume with every chunk that is not loaded emptied."""
    occ = full_vol.occ().reshape(-1, 64).copy()
    full = full_vol.full().reshape(-1, 64).copy()
    occ[~loaded] = 0
    full[~loaded] = 0
    k, p = full_vol.export_partial()
    keep = loaded[(k >> np.uint64(12)).astype(np.int64)]
    v = orc.Volume(origin, dims)
    v.import_(occ, full, k[keep], p[keep])
    return v


@pytest.mark.parametrize("case", ["terrain_blocks", "terrain_voxels", "sphere_voxels"])
def test_stream_updates_match_oracle(ctx, capi, orc, case):
    if case == "terrain_blocks":
        origin, dims, kind, params, gran = (-6, -2, -6), (12, 4, 12), orc.SDF_TERRAIN, None, orc.GRAN_BLOCK
        view, max_new = (5, 2, 120.0, 0), 60
    elif case == "terrain_voxels":
        origin, dims, kind, params, gran = (-2, -1, -1), (3, 2, 3), orc.SDF_TERRAIN, None, orc.GRAN_VOXEL
        view, max_new = (2, 1, 120.0, 0), 4
    else:
        origin, dims, kind, params, gran = (2, -4, -4), (8, 8, 8), orc.SDF_SPHERE, (100.0, 0.0, 0.0, 50.0), orc.GRAN_VOXEL
        view, max_new = (6, 2, 100.0, 0), 90
    nchunks = int(np.prod(dims))
    full_vol = orc.Volume(origin, dims).voxelize(kind, params, granularity=gran, sin_mode=orc.SIN_PORTABLE)
    ctx.scene_create(origin, dims, 1 << 18)
    ctx.stream_begin(kind, params, gran)
    assert not ctx.stream_loaded(nchunks).any()
    loaded = np.zeros(nchunks, dtype=bool)
    centre = tuple(int(origin[i] + dims[i] // 2) for i in range(3))
    moves = [(centre, (0.0, 0.0, 1.0)), (centre, (0.0, 0.0, 1.0)), (centre, (1.0, 0.2, 0.0)),
             ((centre[0] + 1, centre[1], centre[2] - 1), (-0.5, 0.1, -0.8)), (centre, (0.0, 1.0, 0.0))]
    total_generated = 0
    for cam, fwd in moves:
        new, missing, n_cand, in_window = _expected_update(orc, loaded, origin, dims, cam, fwd, view, max_new)
        st = ctx.stream_update(cam, fwd, max_new, capi.view_config(*view))
        assert (int(st["generated"]), int(st["missing"]), int(st["candidates"]), int(st["in_window"])) == (len(new), missing, n_cand, in_window)
        total_generated += len(new)
        assert np.array_equal(ctx.stream_loaded(nchunks), loaded)
        want = _masked(orc, full_vol, origin, dims, loaded)
        occ, full, keys, payload = ctx.volume_download()
        assert np.array_equal(occ, want.occ()) and np.array_equal(full, want.full())
        k2, p2 = want.export_partial()
        assert np.array_equal(keys, k2) and np.array_equal(payload, p2)
    assert total_generated > 0 and loaded.any()
    # the streamed volume renders and meshes exactly like the oracle's masked volume
    want = _masked(orc, full_vol, origin, dims, loaded)
    W, H = 160, 96
    ctr = [(origin[i] + dims[i] / 2.0) * 16.0 for i in range(3)]
    eye = (ctr[0] + 90.0, ctr[1] + 60.0, ctr[2] + 75.0)
    cam_u = orc.camera_uniform(eye, ctr, width=W, height=H)
    rec = ctx.raymarch(cam_u, W, H)
    rs = orc.ray_setup(cam_u, origin, W, H)
    ref = want.raymarch(rs, W, H)
    assert np.array_equal(rec.view(np.uint32).reshape(-1), ref.view(np.uint32).reshape(-1))
    n_inst = ctx.build_occupancy(3)
    table, mips, inst = ctx.download_occupancy(n_inst)
    t2, m2, i2 = want.build_occupancy(3)
    assert np.array_equal(mips, m2) and np.array_equal(inst, i2) and np.array_equal(table, t2)
    q = ctx.mesh(1 << 22)
    q2 = want.mesh()
    assert np.array_equal(orc.sort_quads(q.copy()), orc.sort_quads(q2.copy()))


def test_stream_until_complete_equals_full_voxelize(ctx, capi, orc):
    """Enough updates with a view radius covering the window generate every chunk: identical to meso_voxelize_sdf."""
    origin, dims = (0, -1, 0), (3, 2, 3)
    ctx.scene_create(origin, dims, 1 << 18)
    ctx.stream_begin(capi.SDF_TERRAIN, None, capi.GRAN_VOXEL)
    view = capi.view_config(6, 6, 120.0, 1)
    n = 0
    for _ in range(10):
        st = ctx.stream_update((1, 0, 1), (0.0, 0.0, 1.0), 4, view)
        n += int(st["generated"])
        if st["missing"] == 0:
            break
    assert n == 18 and ctx.stream_loaded(18).all()
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_VOXEL, sin_mode=orc.SIN_PORTABLE)
    occ, full, keys, payload = ctx.volume_download()
    assert np.array_equal(occ, vol.occ()) and np.array_equal(full, vol.full())
    k2, p2 = vol.export_partial()
    assert np.array_equal(keys, k2) and np.array_equal(payload, p2)
    # an update with nothing left to do generates nothing and changes nothing
    st = ctx.stream_update((1, 0, 1), (1.0, 0.0, 0.0), 4, view)
    assert int(st["generated"]) == 0 and int(st["missing"]) == 0


def test_stream_errors(ctx, capi):
    ctx.scene_create((0, 0, 0), (1, 1, 1), 1 << 10)
    ctx.voxelize_sdf(capi.SDF_TERRAIN, None, capi.GRAN_BLOCK)       # ends any stream
    with pytest.raises(capi.MesoError):
        ctx.stream_update((0, 0, 0), (0, 0, 1), 4)
    with pytest.raises(capi.MesoError):
        ctx.select_view_chunks((0, 0, 1), capi.view_config(0, 0, 120.0, 0))
    with pytest.raises(capi.MesoError):
        ctx.select_view_chunks((0, 0, 1), capi.view_config(4, 2, 120.0, 7))


def test_block_importance(ctx, orc):
    """CalculateBlockImportance on the device == the oracle's literal restatement (dead near branch included)."""
    rng = np.random.default_rng(5)
    cam = (3, -2, 5)
    chunks = (rng.integers(-6, 7, size=(3000, 3)) + np.array(cam)).astype(np.int32)
    chunks[:300] = np.array(cam) + rng.integers(-1, 2, size=(300, 3))       # where a live near branch would have fired
    blocks = rng.integers(0, 16, size=(3000, 3)).astype(np.uint8)
    fwd = (0.3, -0.8, 0.52)
    got = ctx.block_importance(cam, fwd, chunks, blocks)
    want = np.array([orc.block_importance(cam, fwd, c, b) for c, b in zip(chunks, blocks)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert not (got == 1.0e6).any()                                          # the reference's near branch is dead code


@pytest.mark.parametrize("gran", ["block", "voxel"])
def test_moving_window_follows_the_camera(ctx, capi, orc, gran):
    """meso_stream_recentre: the window follows the camera over three window widths (and back diagonally).  After every
    move + update the resident volume equals a fresh voxelisation of the window where it now is -- chunks that stayed were
    moved, not regenerated (the generated count says so), chunks that left were evicted and their payload slots reused
    (the pool's high-water mark stays bounded), chunks that entered were generated by the next update."""
    g = capi.GRAN_BLOCK if gran == "block" else capi.GRAN_VOXEL
    og = orc.GRAN_BLOCK if gran == "block" else orc.GRAN_VOXEL
    dims = (5, 3, 3)
    cam = [2, 0, 1]
    ctx.scene_create((0, -1, 0), dims, 1 << 18)
    ctx.stream_begin(capi.SDF_TERRAIN, None, g)
    view = capi.view_config(8, 8, 120.0, 1)           # radius covers the whole window: everything in it is desired
    nchunks = int(np.prod(dims))

    def fill():
        total = 0
        for _ in range(8):
            st = ctx.stream_update(cam, (0.0, 0.0, 1.0), 64, view)
            total += int(st["generated"])
            if st["missing"] == 0:
                break
        return total

    def check():
        origin = tuple(c - d // 2 for c, d in zip(cam, dims))
        assert tuple(ctx.origin) == origin
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=og, sin_mode=orc.SIN_PORTABLE)
        occ, full, keys, payload = ctx.volume_download()
        assert np.array_equal(occ, vol.occ()) and np.array_equal(full, vol.full())
        k2, p2 = vol.export_partial()
        assert np.array_equal(keys, k2) and np.array_equal(payload, p2)
        return vol, origin

    ctx.stream_recentre(cam)                          # the window is already centred on the camera: nothing moves
    assert fill() <= nchunks
    vol, origin = check()
    partial_max = len(vol.export_partial()[0])
    # block granularity: 15 = three window widths along x; voxel granularity (the oracle voxelises every window afresh on the
    # CPU, ~8 s each): one window width, then the same diagonal / multi-chunk / full-width jumps
    moves = [(1, 0, 0)] * (15 if gran == "block" else 5) + [(-2, 1, 1), (-3, -1, -1), (0, 0, 2), (-5, 0, 0)]
    for mv in moves:
        cam = [c + m for c, m in zip(cam, mv)]
        assert ctx.stream_recentre(cam) == mv
        entered = nchunks - int(np.prod([max(0, d - abs(m)) for d, m in zip(dims, mv)]))
        assert fill() == entered                      # only what entered was generated; what stayed was moved
        vol, origin = check()
        partial_max = max(partial_max, len(vol.export_partial()[0]))
    handed_out, free = ctx.pool_stats()
    if gran == "voxel":
        assert partial_max > 0 and handed_out <= 2 * partial_max, (handed_out, free, partial_max)   # slots of evicted chunks were reused
    # the frame through the moved window == the oracle's frame of the same region
    w, h = 128, 72
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    camu = orc.camera_uniform(eyes[1], ctr, width=w, height=h)
    ref = vol.raymarch(orc.ray_setup(camu, origin, w, h), w, h, shadow=True)
    assert ctx.raymarch(camu, w, h, shadow=True).tobytes() == ref.tobytes()
    q = ctx.mesh(1 << 22)
    assert np.array_equal(orc.sort_quads(q.copy()), orc.sort_quads(vol.mesh()))


def test_debug_instances_and_counters(ctx, capi, orc):
    """Debug visualisation parity: one FGPUSimpleInstanceData per resident chunk (Marker 1 = has blocks, 0 = empty chunk) in
    FIVec3Comparator order, and RenderManagerInfo's counters, for a static window and while streaming."""
    assert capi.SimpleInstanceData.itemsize == 48 and capi.DebugStats.itemsize == 40
    origin, dims = (0, -2, 0), (3, 4, 2)
    ctx.scene_create(origin, dims, 1 << 12)
    ctx.voxelize_sdf(capi.SDF_TERRAIN, None, capi.GRAN_BLOCK)
    n_inst = ctx.build_occupancy(stamp=2)
    occ, _, _, _ = ctx.volume_download()
    inst = ctx.debug_chunk_instances()
    locs = sorted((origin[0] + x, origin[1] + y, origin[2] + z) for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0]))
    assert [tuple(int(v) for v in r["ChunkLocation"]) for r in inst] == locs
    for r in inst:
        x, y, z = (int(r["ChunkLocation"][i]) - origin[i] for i in range(3))
        has = bool(occ[x + dims[0] * (y + dims[1] * z)].any())
        assert float(r["Marker"]) == (1.0 if has else 0.0)
        assert tuple(r["Position"]) == (16.0, 16.0, 16.0) and float(r["Scale"]) == np.float32(1.6) and tuple(r["Rotation"]) == (1.0, 0.0, 0.0, 0.0)
    st = ctx.debug_stats()
    assert int(st["LoadedChunk"]) == 24 and int(st["LoadedChunkWithBlocks"]) == int(sum(bool(o.any()) for o in occ)) and int(st["LoadedBlock"]) == n_inst
    # streaming: only generated chunks are resident
    ctx.stream_begin(capi.SDF_TERRAIN, None, capi.GRAN_BLOCK)
    view = capi.view_config(8, 8, 120.0, 1)
    s1 = ctx.stream_update((1, 0, 1), (0.0, 0.0, 1.0), 5, view)
    st = ctx.debug_stats()
    assert int(st["LoadedChunk"]) == 5 and int(st["NewlyAddedVisibleChunk"]) == 5 and int(st["MissingChunk"]) == int(s1["missing"]) == 19
    assert int(st["VisibleChunk"]) == int(s1["candidates"])
    loaded = ctx.stream_loaded(24)
    inst = ctx.debug_chunk_instances()
    assert len(inst) == 5
    for r in inst:
        x, y, z = (int(r["ChunkLocation"][i]) - origin[i] for i in range(3))
        assert loaded[x + dims[0] * (y + dims[1] * z)]
