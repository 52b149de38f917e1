"""GPU parity tests: every kernel, called through the C ABI (libmeso_b200.so), against the CPU oracle on the same
inputs.  Bar: bit-exact (integer/bit outputs and, because both sides avoid fma, also t and rgba)."""
import os

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


def _make(ctx, orc, origin, dims, kind, params, gran, max_bricks=1 << 18):
    ctx.scene_create(origin, dims, max_bricks)
    ctx.voxelize_sdf(kind, params, gran)
    vol = orc.Volume(origin, dims).voxelize(kind, params, granularity=gran, sin_mode=orc.SIN_PORTABLE)
    return vol


def _assert_volume_equal(ctx, vol):
    occ, full, keys, payload = ctx.volume_download()
    assert np.array_equal(occ, vol.occ())
    assert np.array_equal(full, vol.full())
    k2, p2 = vol.export_partial()
    assert np.array_equal(keys, k2)
    assert np.array_equal(payload, p2)


# ---- K1 -------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [128, 256])
def test_voxelize_sphere_voxel(ctx, orc, n):
    origin, dims, params = scenes.sphere_scene(n)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    assert vol.num_partial() > 0
    _assert_volume_equal(ctx, vol)


def test_voxelize_reference_sphere_blocks(ctx, orc):
    """The reference configuration itself: GenerateSphere over the 8^3 chunks that contain it, one sample per block."""
    origin, dims, params = scenes.sphere_scene(1024)
    assert params == scenes.REF_SPHERE
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_BLOCK)
    assert vol.num_partial() == 0
    _assert_volume_equal(ctx, vol)
    assert vol.count_voxels() == 523155 * 512  # block count of the reference sphere (tests/test_oracle_kat.py)


def test_voxelize_terrain_voxel(ctx, orc):
    origin, dims = (0, -1, 0), (1, 2, 1)
    vol = _make(ctx, orc, origin, dims, orc.SDF_TERRAIN, None, orc.GRAN_VOXEL)
    assert 0 < vol.num_partial()
    _assert_volume_equal(ctx, vol)


def test_voxelize_terrain_blocks_with_cull_band(ctx, orc):
    """Tall grid: bricks above/below the +-20.58 band are decided by the exact cull, the rest sampled."""
    origin, dims = (3, -3, -2), (2, 6, 2)
    vol = _make(ctx, orc, origin, dims, orc.SDF_TERRAIN, None, orc.GRAN_BLOCK)
    _assert_volume_equal(ctx, vol)


def test_voxelize_terrain_voxel_cull_band(ctx, orc):
    """Voxel granularity across the upper cull boundary (y = 20.58 world units lies in chunk y=1)."""
    origin, dims = (5, 1, 7), (1, 1, 1)
    vol = _make(ctx, orc, origin, dims, orc.SDF_TERRAIN, None, orc.GRAN_VOXEL)
    _assert_volume_equal(ctx, vol)


def test_volume_upload_roundtrip(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    ctx.scene_create(origin, dims, 1 << 16)
    k, p = vol.export_partial()
    ctx.volume_upload(vol.occ(), vol.full(), k, p)
    _assert_volume_equal(ctx, vol)


def test_empty_volume(ctx, orc):
    """Edge case: nothing solid anywhere (sphere far outside the grid)."""
    origin, dims = (0, 0, 0), (2, 1, 1)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, (1000.0, 0, 0, 5.0), orc.GRAN_VOXEL)
    _assert_volume_equal(ctx, vol)
    assert ctx.build_occupancy(7) == 0
    cam = orc.camera_uniform((5, 2, 2), (16, 8, 8), width=64, height=32)
    rec = ctx.raymarch(cam, 64, 32)
    assert np.all(rec["w1"] == 0x0007FFFF) and np.all(rec["rgba"] == 0xFF000000)
    assert len(ctx.mesh(16)) == 0


# ---- K2 -------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["ref_sphere_blocks", "sphere_voxel_256", "terrain_blocks"])
def test_occupancy_mips_instances(ctx, orc, case):
    if case == "ref_sphere_blocks":
        origin, dims, params = scenes.sphere_scene(1024)
        vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_BLOCK)
    elif case == "sphere_voxel_256":
        origin, dims, params = scenes.sphere_scene(256)
        vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    else:
        origin, dims = (0, -2, 0), (3, 4, 3)
        vol = _make(ctx, orc, origin, dims, orc.SDF_TERRAIN, None, orc.GRAN_BLOCK)
    n = ctx.build_occupancy(stamp=42)
    table, mips, inst = ctx.download_occupancy(n)
    t2, m2, i2 = vol.build_occupancy(stamp=42)
    assert np.array_equal(mips, m2)
    assert table.tobytes() == t2.tobytes()
    assert n == len(i2)
    assert inst.tobytes() == i2.tobytes()  # same records in the same (generator) order
    if case == "ref_sphere_blocks":
        assert n == 201936


# ---- K4 -------------------------------------------------------------------------------------------------------

def _cams(orc, origin, dims, w, h, extra=()):
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cams = [orc.camera_uniform(e, ctr, width=w, height=h) for e in eyes]
    for e, c in extra:
        cams.append(orc.camera_uniform(e, c, width=w, height=h))
    return cams


def _compare_frames(ctx, orc, vol, cams, w, h, shadow=True):
    for i, cam in enumerate(cams):
        rec = ctx.raymarch(cam, w, h, shadow=shadow)
        rs = orc.ray_setup(cam, vol.origin, w, h)
        ref = vol.raymarch(rs, w, h, shadow=shadow, mode=orc.DDA_HIER)
        if rec.tobytes() != ref.tobytes():
            a, b = orc.unpack_records(rec), orc.unpack_records(ref)
            bad = {k: int((a[k] != b[k]).sum()) for k in a if k != "t"}
            bad["t"] = int((rec["t"].view(np.uint32) != ref["t"].view(np.uint32)).sum())
            raise AssertionError("camera %d: mismatching fields %s" % (i, bad))


def test_raymarch_sphere_voxel_256(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 320, 184
    ctr = scenes.grid_center_world(origin, dims)
    extra = [((ctr[0] + 1.3, ctr[1] + 0.2, ctr[2] - 0.7), (ctr[0] + 30, ctr[1] + 5, ctr[2] + 3)),   # eye inside the solid
             ((ctr[0] - 13.1, ctr[1] + 0.3, ctr[2] + 0.2), ctr),                                    # skimming just outside
             ((5.0, 2.0, 2.0), (0.0, 0.0, 0.0))]                                                     # reference start pose (VoxelCamera.h:23)
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, w, h, extra), w, h)


def test_raymarch_no_shadow(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, 160, 96)[:3], 160, 96, shadow=False)


def test_raymarch_reference_blocks(ctx, orc):
    """Reference semantics: block-granular reference sphere (every brick all-ones)."""
    origin, dims, params = scenes.sphere_scene(1024)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_BLOCK)
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, 256, 144)[:4], 256, 144)


def test_raymarch_terrain(ctx, orc):
    origin, dims = (0, -1, 0), (2, 2, 2)
    vol = _make(ctx, orc, origin, dims, orc.SDF_TERRAIN, None, orc.GRAN_BLOCK)
    w, h = 256, 144
    ctr = scenes.grid_center_world(origin, dims)
    extra = [((ctr[0] - 14.0, 9.0, ctr[2] - 13.0), (ctr[0] + 10, -3.0, ctr[2] + 12))]  # skim over the surface from inside the grid
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, w, h, extra), w, h)


def test_raymarch_ragged_frame(ctx, orc):
    """Frame size not a multiple of the 32x8 tile."""
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, 101, 37)[:3], 101, 37)


def test_raymarch_stats_and_u(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 320, 184
    cam = _cams(orc, origin, dims, w, h)[2]
    st = ctx.raymarch_stats(cam, w, h)
    _, ref = vol.raymarch(orc.ray_setup(cam, vol.origin, w, h), w, h, stats=True)
    # `steps` is not a parity quantity: the kernel also skips 512^3 regions and 32^3 cells, the oracle walks 128/8/1
    for k in ("primary", "shadow", "hits", "touched_chunks", "touched_bricks", "u_bytes"):
        assert int(st[k]) == int(ref[k]), k
    assert 0 < int(st["steps"]) <= int(ref["steps"])


def test_raymarch_tile_partition_composes(ctx, capi, orc):
    """Multi-GPU data path on one device: every rank's packed tiles, gathered and composed, equal the 1-GPU frame."""
    import torch
    origin, dims, params = scenes.sphere_scene(256)
    _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 320, 184
    cam = _cams(orc, origin, dims, w, h)[1]
    ctx.set_partition(0, 1)
    full = ctx.raymarch(cam, w, h).copy()
    for world in (2, 3, 8):
        tpr = capi.tiles_per_rank(w, h, world)
        gathered = torch.zeros((world, tpr, 256, 4), dtype=torch.int32, device="cuda")
        for rank in range(world):
            ctx.set_partition(rank, world)
            ctx.raymarch_device(cam, w, h, gathered[rank].data_ptr(), layout=capi.LAYOUT_TILES)
        ctx.set_partition(0, 1)
        frame = torch.zeros((h, w, 4), dtype=torch.int32, device="cuda")
        ctx.compose_tiles_device(gathered.data_ptr(), world, w, h, frame.data_ptr())
        ctx.sync()
        torch.cuda.synchronize()
        got = frame.cpu().numpy().view(np.uint32).reshape(h, w, 4)
        exp = full.view(np.uint32).reshape(h, w, 4)
        assert np.array_equal(got, exp), world


def test_frame_ring_async_equals_sync(ctx, capi, orc):
    """meso_raymarch_async / meso_frame_wait (4-slot ring, copy of frame k overlapping frame k+1) returns exactly the
    frames meso_raymarch does, slot reuse included."""
    import torch
    origin, dims, params = scenes.sphere_scene(256)
    _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 320, 184
    cams = _cams(orc, origin, dims, w, h)
    expect = [ctx.raymarch(c, w, h).copy() for c in cams]
    hosts = [torch.empty((h, w, 4), dtype=torch.int32).pin_memory() for _ in range(4)]
    views = [t.numpy().view(capi.HitRecord).reshape(h, w) for t in hosts]
    for k in range(8):
        slot = k % 4
        if k >= 4:
            ctx.frame_wait(slot)
            assert views[slot].tobytes() == expect[k - 4].tobytes()
        ctx.raymarch_async(cams[k], w, h, views[slot], slot)
    for k in range(4, 8):
        ctx.frame_wait(k % 4)
        assert views[k % 4].tobytes() == expect[k].tobytes()
    with pytest.raises(capi.MesoError):
        ctx.raymarch_async(cams[0], w, h, views[0], 7)   # slot out of range -> ArgumentOutOfRange-class error


def test_edits_are_ordered_behind_frames_in_flight(ctx, capi, orc):
    """meso_carve_sphere between two meso_raymarch_async frames must not race the frame that was started before it (join_frames in meso_capi.cu).
    Frame A (started before the carve) shows the uncarved volume, frame B the carved one; repeated to give a race a chance."""
    import torch
    origin, dims, params = scenes.sphere_scene(256)
    w, h = 640, 368
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[2], ctr, width=w, height=h)
    hosts = [torch.empty((h, w, 4), dtype=torch.int32).pin_memory() for _ in range(2)]
    views = [t.numpy().view(capi.HitRecord).reshape(h, w) for t in hosts]
    for rep in range(5):
        vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
        before = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True)
        # the carve sits where this camera looks: radius grows per repetition so every round changes many pixels
        hit = np.argwhere(((before["w1"] >> 20) & 1) == 1)
        py, px = hit[len(hit) // 2]
        center = (int(before["w0"][py, px] & 0xFFFF), int(before["w0"][py, px] >> 16), int(before["w1"][py, px] & 0xFFFF))
        ctx.raymarch_async(cam, w, h, views[0], 0)
        ctx.carve_sphere(center, 20 + 6 * rep)            # enqueued while frame 0 may still be tracing
        ctx.raymarch_async(cam, w, h, views[1], 1)
        vol.carve_sphere(center, 20 + 6 * rep)
        after = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True)
        ctx.frame_wait(0); ctx.frame_wait(1)
        assert before.tobytes() != after.tobytes()
        assert views[0].tobytes() == before.tobytes(), "the carve raced the frame started before it (rep %d)" % rep
        assert views[1].tobytes() == after.tobytes(), rep


def test_rgba8_output_is_the_colour_word_of_the_records(ctx, capi, orc):
    """MESO_FLAG_RGBA8 (the reference's RGBA_UN8 offscreen colour target): 4 B per pixel, equal to record.rgba, through the
    synchronous banded call (ragged height: bands end on different rows) and through the frame ring; oracle-checked."""
    import torch
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 322, 531
    cam = _cams(orc, origin, dims, w, h)[2]
    ref = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h)
    img = ctx.raymarch(cam, w, h, rgba8=True)
    assert img.dtype == np.uint32 and img.shape == (h, w)
    assert np.array_equal(img, ref["rgba"])
    assert np.array_equal(ctx.raymarch(cam, w, h)["rgba"], img)
    pinned = torch.empty((h, w), dtype=torch.int32).pin_memory()
    view = pinned.numpy().view(np.uint32)
    ctx.raymarch_async(cam, w, h, view, 1, rgba8=True)
    ctx.frame_wait(1)
    assert np.array_equal(view, img)
    with pytest.raises(capi.MesoError):   # packed tile layout carries whole records only
        ctx.raymarch_device(cam, w, h, ctx.device_alloc(w * h * 16), layout=capi.LAYOUT_TILES, flags_extra=capi.FLAG_RGBA8)


def test_host_registered_frame_receives_kernel_stores(ctx, capi, orc, tmp_path):
    """Fused gather into host memory on one device: a file-backed shared mapping is registered (meso_host_register), two
    'ranks' (tile partitions 0/2 and 1/2) store their tiles into it straight from the kernel, and the mapping then holds
    exactly the frame of the synchronous call."""
    origin, dims, params = scenes.sphere_scene(256)
    _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 320, 184
    cam = _cams(orc, origin, dims, w, h)[3]
    expect = ctx.raymarch(cam, w, h).copy()
    path = "/dev/shm/meso_test_%d" % os.getpid()
    try:
        with open(path, "wb") as f:
            f.truncate(w * h * 16)
        shm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(w * h * 16,))
        shm[:] = 0xEE
        dptr = ctx.host_register(shm)
        for r in (0, 1):
            ctx.set_partition(r, 2)
            ctx.raymarch_device(cam, w, h, dptr)
        ctx.set_partition(0, 1)
        ctx.sync()
        assert shm.view(capi.HitRecord).reshape(h, w).tobytes() == expect.tobytes()
        ctx.host_unregister(shm)
        del shm
    finally:
        if os.path.exists(path):
            os.unlink(path)
    with pytest.raises(capi.MesoError):
        ctx.host_unregister(np.zeros(64, dtype=np.uint8))   # never registered


def test_device_alloc_download(ctx, capi, orc):
    """The buffer path of the fused multi-GPU gather on one device: library-owned frame, kernel stores, download."""
    origin, dims, params = scenes.sphere_scene(256)
    _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    w, h = 320, 184
    cam = _cams(orc, origin, dims, w, h)[5]
    ptr = ctx.device_alloc(w * h * 16)
    handle = ctx.ipc_export(ptr)
    assert handle.shape == (capi.IPC_HANDLE_BYTES,) and handle.any()
    ctx.raymarch_device(cam, w, h, ptr)
    got = np.zeros((h, w), dtype=capi.HitRecord)
    ctx.download(got, ptr)
    ctx.device_free(ptr)
    assert got.tobytes() == ctx.raymarch(cam, w, h).tobytes()


# ---- K3 -------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["sphere_voxel_256", "sphere_blocks", "terrain_voxel"])
def test_mesh_quads(ctx, orc, case):
    if case == "sphere_voxel_256":
        origin, dims, params = scenes.sphere_scene(256)
        vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    elif case == "sphere_blocks":
        origin, dims, params = scenes.sphere_scene(512)
        vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_BLOCK)
    else:
        origin, dims = (0, -1, 0), (1, 2, 1)
        vol = _make(ctx, orc, origin, dims, orc.SDF_TERRAIN, None, orc.GRAN_VOXEL)
    ref = vol.mesh()
    got = ctx.mesh(len(ref) + 1024)
    assert len(got) == len(ref)
    assert orc.sort_quads(got).tobytes() == orc.sort_quads(ref).tobytes()
    area = int((((got["w1"] >> 24) & 0xFF).astype(np.int64) * got["w2"].astype(np.int64)).sum())
    assert area == vol.count_exposed_faces()


def test_mesh_worst_case_bricks(ctx, orc):
    """Hand-built volume through meso_volume_upload: checkerboard bricks (1536 unmergeable quads each: the kernel's
    staging area overflows and the two-pass path runs), random-noise bricks, full bricks next to them, grid borders."""
    origin, dims = (0, 0, 0), (2, 1, 1)
    nch = 2
    rng = np.random.default_rng(11)
    occ = np.zeros((nch, 64), dtype=np.uint64)
    full = np.zeros((nch, 64), dtype=np.uint64)
    keys, payload = [], []
    checker = np.array([0xAA55AA55AA55AA55 if z % 2 == 0 else 0x55AA55AA55AA55AA for z in range(8)], dtype=np.uint64)
    for c in range(nch):
        for b in rng.choice(4096, size=300, replace=False):
            b = int(b)
            kind = rng.integers(0, 3)
            occ[c, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
            if kind == 0:
                full[c, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
            else:
                keys.append(c * 4096 + b)
                payload.append(checker if kind == 1 else rng.integers(1, 2 ** 63, size=8, dtype=np.int64).astype(np.uint64))
    # two checkerboard bricks side by side and one at the grid corner
    for c, b in ((0, 0), (0, 1), (1, 4095)):
        if not (int(occ[c, b >> 6]) >> (b & 63)) & 1:
            occ[c, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
            keys.append(c * 4096 + b); payload.append(checker)
    order = np.argsort(np.array(keys, dtype=np.uint64))
    keys = np.array(keys, dtype=np.uint64)[order]
    payload = np.stack(payload)[order]
    vol = orc.Volume(origin, dims).import_(occ, full, keys, payload)
    ctx.scene_create(origin, dims, 1 << 12)
    ctx.volume_upload(occ, full, keys, payload)
    ref = vol.mesh()
    got = ctx.mesh(len(ref) + 64)
    assert len(ref) > 1536 * 3
    assert orc.sort_quads(got).tobytes() == orc.sort_quads(ref).tobytes()
    # the same hand-built volume also raymarches identically (bricks with arbitrary payloads)
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, 160, 96)[:3], 160, 96)


def test_mesh_brick_level(ctx, orc):
    """The second merge level: faces of full bricks towards absent bricks, merged per chunk.  A flat slab (one quad per chunk
    face), a 3-D checkerboard of full bricks (2048 bricks x 6 unmergeable faces per chunk: the pass's staging area overflows),
    and a random full / partial / absent mix across chunk borders and the grid border."""
    origin, dims = (0, 0, 0), (2, 2, 1)
    nch = 4
    rng = np.random.default_rng(5)
    occ = np.zeros((nch, 64), dtype=np.uint64); full = np.zeros((nch, 64), dtype=np.uint64)
    keys, payload = [], []
    # chunk 0: a slab of 16 x 16 x 2 full bricks; chunk 1: checkerboard of full bricks; chunks 2, 3: random mix
    for z in (3, 4):
        occ[0, z * 4:z * 4 + 4] = ~np.uint64(0); full[0, z * 4:z * 4 + 4] = ~np.uint64(0)
    for b in range(4096):
        if ((b & 15) + ((b >> 4) & 15) + (b >> 8)) & 1:
            occ[1, b >> 6] |= np.uint64(1) << np.uint64(b & 63); full[1, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
    for c in (2, 3):
        kind = rng.integers(0, 4, size=4096)             # 0,1 absent / 2 full / 3 partial
        for b in np.nonzero(kind >= 2)[0]:
            b = int(b)
            occ[c, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
            if kind[b] == 2:
                full[c, b >> 6] |= np.uint64(1) << np.uint64(b & 63)
            else:
                keys.append(c * 4096 + b)
                payload.append(rng.integers(1, 2 ** 63, size=8, dtype=np.int64).astype(np.uint64) if rng.integers(0, 2) else np.array([0, 0, 0, 1 << 27, 0, 0, 0, 0], dtype=np.uint64))
    keys = np.array(keys, dtype=np.uint64); payload = np.stack(payload)
    vol = orc.Volume(origin, dims).import_(occ, full, keys, payload)
    ctx.scene_create(origin, dims, 1 << 13)
    ctx.volume_upload(occ, full, keys, payload)
    ref = vol.mesh()
    got = ctx.mesh(len(ref) + 64)
    assert orc.sort_quads(got).tobytes() == orc.sort_quads(ref).tobytes()
    level = (got["w1"] >> 19) & 1
    cx = (got["w0"] & 0xFFFF) >> 7; cy = (got["w0"] >> 16) >> 7
    in0 = (cx == 0) & (cy == 0)
    # the slab: its two 128 x 128 faces are one quad each (its +x / +y sides touch chunks 1 and 2 and merge less)
    big = got[in0 & (level == 1) & (((got["w1"] >> 24) & 0xFF) == 128) & (got["w2"] == 128)]
    assert len(big) == 2 and sorted(((big["w1"] >> 16) & 7).tolist()) == [4, 5]
    area = int((((got["w1"] >> 24) & 0xFF).astype(np.int64) * got["w2"].astype(np.int64)).sum())
    assert area == vol.count_exposed_faces()
    # the partition over ranks covers both levels
    parts = []
    for r in range(3):
        ctx.set_partition(r, 3)
        parts.append(ctx.mesh(len(ref) + 64))
    ctx.set_partition(0, 1)
    assert orc.sort_quads(np.concatenate(parts)).tobytes() == orc.sort_quads(ref).tobytes()
    # edits: a carve through the slab and the checkerboard turns full bricks partial / absent; the re-mesh returns both levels
    for center, radius in (((100, 100, 40), 30), ((200, 60, 64), 45), ((128, 128, 64), 20)):
        nd = ctx.carve_sphere(center, radius)
        assert np.array_equal(ctx.download_dirty(nd), vol.carve_sphere(center, radius))
        quads, rkeys = ctx.remesh_dirty(1 << 20, 1 << 20)
        assert orc.sort_quads(quads).tobytes() == orc.sort_quads(vol.remesh(np.sort(rkeys))).tobytes()
    ref = vol.mesh()
    assert orc.sort_quads(ctx.mesh(len(ref) + 64)).tobytes() == orc.sort_quads(ref).tobytes()


def test_mesh_partition_union(ctx, orc):
    """Brick ranges over 'ranks' (chunk % world): the union of the per-rank quad lists is the 1-GPU list."""
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    ref = orc.sort_quads(vol.mesh())
    parts = []
    for rank in range(3):
        ctx.set_partition(rank, 3)
        parts.append(ctx.mesh(len(ref) + 16))
    ctx.set_partition(0, 1)
    got = orc.sort_quads(np.concatenate(parts))
    assert got.tobytes() == ref.tobytes()


def test_mesh_shared_list_fused_gather(ctx, capi, orc):
    """The buffer path of the fused quad gather on one device: three 'ranks' append their chunks' quads to ONE list behind
    ONE counter (meso_mesh_device_shared); the list is the 1-GPU mesh."""
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    ref = orc.sort_quads(vol.mesh())
    cap = len(ref) + 64
    qptr, cptr = ctx.device_alloc(cap * 16), ctx.device_alloc(8)
    ctx.device_memset(cptr, 0, 8)
    for rank in range(3):
        ctx.set_partition(rank, 3)
        ctx.mesh_device_shared(qptr, cptr, cap)
    ctx.set_partition(0, 1)
    cnt = np.zeros(1, dtype=np.uint64)
    ctx.download(cnt, cptr)
    assert int(cnt[0]) == len(ref)
    got = np.zeros(len(ref), dtype=capi.Quad)
    ctx.download(got, qptr, len(ref) * 16)
    ctx.device_free(qptr); ctx.device_free(cptr)
    assert orc.sort_quads(got).tobytes() == ref.tobytes()


# ---- K5 -------------------------------------------------------------------------------------------------------

def test_carve_and_remesh(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _make(ctx, orc, origin, dims, orc.SDF_SPHERE, params, orc.GRAN_VOXEL)
    n = 256
    carves = [((n // 2 - 90, n // 2, n // 2), 24), ((n // 2, n // 2 + 100, n // 2 + 10), 40), ((3, 5, 250), 9), ((n // 2, n // 2, n // 2), 30)]
    for center, radius in carves:
        nd = ctx.carve_sphere(center, radius)
        dirty = ctx.download_dirty(nd)
        ref_dirty = vol.carve_sphere(center, radius)
        assert np.array_equal(dirty, ref_dirty)
        _assert_volume_equal(ctx, vol)
        quads, keys = ctx.remesh_dirty(1 << 20, 1 << 20)
        keys = np.sort(keys)
        ref_q = vol.remesh(keys)      # voxel-level quads of the bricks + brick-level quads of the chunks that hold them
        assert orc.sort_quads(quads).tobytes() == orc.sort_quads(ref_q).tobytes()
    # after the edits a full re-mesh and a frame still agree with the oracle
    ref = vol.mesh()
    got = ctx.mesh(len(ref) + 16)
    assert orc.sort_quads(got).tobytes() == orc.sort_quads(ref).tobytes()
    _compare_frames(ctx, orc, vol, _cams(orc, origin, dims, 160, 96)[:2], 160, 96)


def test_upload_blocks_reference_records(ctx, capi, orc):
    """meso_volume_upload_blocks: the reference's own upload records (FGPUChunk table + FGPUBlock pool, ChunkPool.h:662-679).
    Blocks of the reference sphere as records == the device generator's volume; invalid records (INT_MAX index, stale
    stamp, invalid chunk, chunk outside the window) are skipped exactly like the vertex shader skips them; merge adds."""
    origin, dims = (2, -4, -4), (8, 8, 8)
    ctx.scene_create(origin, dims, 1 << 12)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, orc.REF_SPHERE, granularity=orc.GRAN_BLOCK)
    occ = vol.occ()
    n_chunks = int(np.prod(dims))
    table = np.zeros(n_chunks + 2, dtype=capi.GPUChunk)
    blocks = []
    stamp = 7
    for c in range(n_chunks):
        bits = np.unpackbits(occ[c].view(np.uint8), bitorder="little")
        idx = np.nonzero(bits)[0]
        loc = (origin[0] + c % dims[0], origin[1] + (c // dims[0]) % dims[1], origin[2] + c // (dims[0] * dims[1]))
        table[c] = ((loc if len(idx) else (2**31 - 1,) * 3), stamp if len(idx) else 0)
        for b in idx:
            blocks.append((c, (b & 15, (b >> 4) & 15, b >> 8, 255), stamp))
    n_good = len(blocks)
    table[n_chunks] = ((100, 100, 100), stamp)            # valid record, outside the window
    table[n_chunks + 1] = ((2**31 - 1,) * 3, stamp)       # invalid chunk record
    blocks += [(2**31 - 1, (1, 1, 1, 255), stamp), (0, (1, 1, 1, 255), stamp + 1), (n_chunks, (1, 1, 1, 255), stamp),
               (n_chunks + 1, (1, 1, 1, 255), stamp), (n_chunks + 5, (1, 1, 1, 255), stamp)]
    blocks = np.array(blocks, dtype=capi.GPUBlock)
    rng = np.random.default_rng(3)
    blocks = blocks[rng.permutation(len(blocks))]          # the pool order is thread-scheduled in the reference
    assert ctx.volume_upload_blocks(table, blocks) == n_good
    o2, f2, keys, _ = ctx.volume_download()
    assert np.array_equal(o2, occ) and np.array_equal(f2, occ) and len(keys) == 0
    n = ctx.build_occupancy(stamp=3)
    _, _, inst = ctx.download_occupancy(n)
    _, _, i_ref = vol.build_occupancy(stamp=3)
    assert inst.tobytes() == i_ref.tobytes()
    # the frame through the uploaded records == the oracle's
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[1], ctr, width=128, height=72)
    ref = vol.raymarch(orc.ray_setup(cam, origin, 128, 72), 128, 72, shadow=True)
    assert ctx.raymarch(cam, 128, 72, shadow=True).tobytes() == ref.tobytes()
    # delta upload: half the records, then the other half merged
    good = blocks[(blocks["ChunkIndex"] < n_chunks) & (blocks["BlockFrameStamp"] == stamp)]
    assert ctx.volume_upload_blocks(table, good[: n_good // 2]) == n_good // 2
    assert ctx.volume_upload_blocks(table, good[n_good // 2:], merge=True) == n_good - n_good // 2
    o3, f3, _, _ = ctx.volume_download()
    assert np.array_equal(o3, occ) and np.array_equal(f3, occ)
    # replace empties the window
    assert ctx.volume_upload_blocks(table, blocks[:0]) == 0
    o4, _, _, _ = ctx.volume_download()
    assert not o4.any()


def test_volume_upload_rejects_inconsistent_input(ctx, capi, orc):
    """meso_volume_upload validates what the kernels rely on: one key per occ && !full brick, ascending, naming a partial brick."""
    origin, dims, params = scenes.sphere_scene(256)
    ctx.scene_create(origin, dims, 1 << 16)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    keys, payload = vol.export_partial()
    occ, full = vol.occ(), vol.full()
    ctx.volume_upload(occ, full, keys, payload)
    with pytest.raises(capi.MesoError):
        ctx.volume_upload(occ, full, keys[:-1], payload[:-1])            # a partial brick without a payload
    with pytest.raises(capi.MesoError):
        ctx.volume_upload(occ, full, keys[::-1].copy(), payload)          # not ascending
    bad = keys.copy(); bad[0] = bad[1]
    with pytest.raises(capi.MesoError):
        ctx.volume_upload(occ, full, bad, payload)                        # duplicate
    f2 = full.copy(); f2.reshape(-1)[int(keys[0]) >> 6] |= np.uint64(1) << np.uint64(int(keys[0]) & 63)
    with pytest.raises(capi.MesoError):
        ctx.volume_upload(occ, f2, keys, payload)                         # a key naming a full brick
    ctx.volume_upload(occ, full, keys, payload)                           # still usable afterwards
    o2, f3, k2, p2 = ctx.volume_download()
    assert np.array_equal(o2, occ) and np.array_equal(k2, keys) and np.array_equal(p2, payload)


def test_pool_overflow_leaves_a_consistent_volume(ctx, capi, orc):
    """Voxelise / carve with too small a payload pool: the call fails, and what is left renders and meshes without
    touching memory it does not own (no occ && !full brick without a payload slot)."""
    origin, dims, params = scenes.sphere_scene(256)
    ctx.scene_create(origin, dims, 64)                     # far too few payload slots
    with pytest.raises(capi.MesoError):
        ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
    occ, full, keys, payload = ctx.volume_download()
    assert len(keys) <= 64
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[0], ctr, width=96, height=60)
    vol = orc.Volume(origin, dims).import_(occ, full, keys, payload)
    ref = vol.raymarch(orc.ray_setup(cam, origin, 96, 60), 96, 60, shadow=True)
    assert ctx.raymarch(cam, 96, 60, shadow=True).tobytes() == ref.tobytes()
    # carve: a block-granular sphere has no payload at all; carving its surface needs slots -> overflow, bricks left full
    ctx.scene_create(origin, dims, 8)
    ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_BLOCK)
    with pytest.raises(capi.MesoError):
        ctx.carve_sphere((128, 128, 60), 30)
    occ, full, keys, payload = ctx.volume_download()
    assert len(keys) <= 8
    vol = orc.Volume(origin, dims).import_(occ, full, keys, payload)
    ref = vol.raymarch(orc.ray_setup(cam, origin, 96, 60), 96, 60, shadow=True)
    assert ctx.raymarch(cam, 96, 60, shadow=True).tobytes() == ref.tobytes()
    q = ctx.mesh(1 << 20)
    assert orc.sort_quads(q).tobytes() == orc.sort_quads(vol.mesh()).tobytes()


def test_present_rgba8_texture_ranges(ctx, capi, orc):
    """meso_present_rgba8: the colour words of the records, tightly packed for a TextureRangeDesc-style range (whole frame, a
    ragged interior rectangle, a single pixel, the last row); ranges outside the frame are rejected."""
    origin, dims, params = scenes.sphere_scene(256)
    ctx.scene_create(origin, dims, 1 << 16)
    ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    w, h = 331, 187
    cam = orc.camera_uniform(eyes[4], ctr, width=w, height=h)
    rec = ctx.raymarch(cam, w, h, shadow=True)
    assert np.array_equal(ctx.present_rgba8(cam, w, h), rec["rgba"])
    for (x, y, rw, rh) in ((37, 21, 200, 100), (150, 90, 1, 1), (0, h - 1, w, 1), (w - 5, 0, 5, h)):
        got = ctx.present_rgba8(cam, w, h, rect=(x, y, rw, rh))
        assert got.shape == (rh, rw) and np.array_equal(got, rec["rgba"][y:y + rh, x:x + rw])
    with pytest.raises(capi.MesoError):
        ctx.present_rgba8(cam, w, h, rect=(w - 4, 0, 5, 5))


@pytest.mark.parametrize("kind", ["sphere", "terrain"])
def test_generator_lod_mipmap_level(ctx, capi, orc, kind):
    """MipmapLevel of the generator plug-in: one sample per (2^m)^3 blocks.  Level 0 == the plain block generator; every
    level == the oracle's; a coarser level never has more distinct block groups than a finer one."""
    if kind == "sphere":
        origin, dims, k, ok, params = (2, -4, -4), (8, 8, 8), capi.SDF_SPHERE, orc.SDF_SPHERE, orc.REF_SPHERE
    else:
        origin, dims, k, ok, params = (0, -2, 0), (3, 4, 3), capi.SDF_TERRAIN, orc.SDF_TERRAIN, None
    ctx.scene_create(origin, dims, 1 << 10)
    ctx.voxelize_sdf(k, params, capi.GRAN_BLOCK)
    occ0 = ctx.volume_download()[0]
    for m in range(5):
        ctx.voxelize_sdf_lod(k, params, m)
        occ, full, keys, _ = ctx.volume_download()
        vol = orc.Volume(origin, dims).voxelize_lod(ok, params, mip=m)
        assert np.array_equal(occ, vol.occ()) and np.array_equal(full, vol.full()) and len(keys) == 0
        if m == 0:
            assert np.array_equal(occ, occ0)
    with pytest.raises(capi.MesoError):
        ctx.voxelize_sdf_lod(k, params, 5)
    # the coarse volume renders and meshes like any other
    ctx.voxelize_sdf_lod(k, params, 2)
    vol = orc.Volume(origin, dims).voxelize_lod(ok, params, mip=2)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[1], ctr, width=128, height=72)
    assert ctx.raymarch(cam, 128, 72, shadow=True).tobytes() == vol.raymarch(orc.ray_setup(cam, origin, 128, 72), 128, 72, shadow=True).tobytes()
    assert orc.sort_quads(ctx.mesh(1 << 20)).tobytes() == orc.sort_quads(vol.mesh()).tobytes()
