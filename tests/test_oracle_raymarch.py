"""CPU-only checks of the raymarch definition in the oracle.

  * ORC_DDA_HIER (what the GPU implements) == ORC_DDA_FLAT (the one-voxel-per-step definition), bit for bit;
  * on block-granular volumes the DDA reproduces what the reference's instanced draw resolves per pixel
    (orc_ref_instanced_pixel restates Samples/SimpleVoxel.cpp:146-192,220-224): same block, same face, distance within
    1e-5 relative, colour within 1 LSB -- the tolerances BASELINE.json's north_star states;
  * record conventions (miss record, face ids, shadow rule).
"""
import numpy as np
import pytest

import scenes


def _cams(orc, origin, dims, w, h, n=8):
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    return [orc.camera_uniform(e, ctr, width=w, height=h) for e in eyes[:n]]


@pytest.mark.parametrize("case", ["sphere_voxel", "sphere_block", "terrain_block", "clipped_sphere"])
def test_hier_equals_flat(orc, case):
    if case == "sphere_voxel":
        origin, dims, params = scenes.sphere_scene(256)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    elif case == "sphere_block":
        origin, dims, params = scenes.sphere_scene(512)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_BLOCK)
    elif case == "terrain_block":
        origin, dims = (0, -1, 0), (2, 2, 2)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_BLOCK)
    else:  # the grid cuts the sphere in half: rays enter solid matter through the grid boundary
        origin, dims, params = scenes.sphere_scene(128)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    w, h = 160, 90
    ctr = scenes.grid_center_world(origin, dims)
    cams = _cams(orc, origin, dims, w, h)
    cams.append(orc.camera_uniform((ctr[0] + 0.37, ctr[1] + 0.21, ctr[2] - 0.4), (ctr[0] + 9, ctr[1] + 2, ctr[2] + 1), width=w, height=h))  # inside
    cams.append(orc.camera_uniform((ctr[0], ctr[1], ctr[2] + 30.0), ctr, up=(0, 1, 0), width=w, height=h))  # axis-aligned view: zero direction components
    for cam in cams:
        rs = orc.ray_setup(cam, origin, w, h)
        a = vol.raymarch(rs, w, h, mode=orc.DDA_HIER)
        b = vol.raymarch(rs, w, h, mode=orc.DDA_FLAT)
        assert a.tobytes() == b.tobytes()
        # unaligned empty cubes from a cell distance field (what the CUDA walk skips with, built here by brute force
        # and with another cap): the records do not depend on which empty boxes are skipped
        c = vol.raymarch(rs, w, h, mode=orc.DDA_BOX)
        assert c.tobytes() == b.tobytes()


def test_box_mode_really_skips(orc):
    """ORC_DDA_BOX takes fewer steps than the aligned hierarchy for the same (byte-identical) frame."""
    origin, dims, params = scenes.sphere_scene(512)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL, fast=True)
    w, h = 128, 72
    cam = _cams(orc, origin, dims, w, h)[1]
    rs = orc.ray_setup(cam, origin, w, h)
    a, sa = vol.raymarch(rs, w, h, mode=orc.DDA_HIER, stats=True)
    c, sc = vol.raymarch(rs, w, h, mode=orc.DDA_BOX, stats=True)
    assert a.tobytes() == c.tobytes()
    assert int(sc["steps"]) < int(sa["steps"])
    for k in ("primary", "shadow", "hits", "touched_bricks"):
        assert int(sc[k]) == int(sa[k])


def test_record_conventions(orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    w, h = 96, 54
    cam = _cams(orc, origin, dims, w, h)[0]
    rec = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h)
    u = orc.unpack_records(rec)
    miss = u["hit"] == 0
    assert miss.any() and (~miss).any()
    assert np.all(rec["w0"][miss] == 0xFFFFFFFF) and np.all(rec["w1"][miss] == 0x0007FFFF)
    assert np.all(np.isinf(rec["t"][miss])) and np.all(rec["rgba"][miss] == 0xFF000000)          # clear (0,0,0,1)
    hit = ~miss
    assert np.all(u["face"][hit] < 6) and np.all(rec["rgba"][hit] >> 24 == 0)                    # alpha 0 on hits
    for ch in range(3):  # colour = local*0.5+0.5 (x shade 0.5 or 1): every channel within [0.125, 0.75]
        c = (rec["rgba"][hit] >> (8 * ch)) & 0xFF
        assert c.min() >= 31 and c.max() <= 192
    # the hit voxel is solid and the voxel in front of the hit face is empty
    ys, xs = np.nonzero(hit)
    for y, x in list(zip(ys, xs))[::37]:
        vx, vy, vz, f = int(u["x"][y, x]), int(u["y"][y, x]), int(u["z"][y, x]), int(u["face"][y, x])
        assert vol.get_voxel(vx, vy, vz) == 1
        n = [0, 0, 0]; n[f >> 1] = 1 if f & 1 else -1
        assert vol.get_voxel(vx + n[0], vy + n[1], vz + n[2]) == 0
    # shadow rule: faces turned away from the light are always shadowed
    light = np.array([0.3, 0.5, 0.8])
    away = hit & (np.where(u["face"] & 1, 1.0, -1.0) * light[(u["face"] >> 1).clip(0, 2)] <= 0)
    assert np.all(u["shadow"][away] == 1)
    no_shadow = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=False)
    assert np.all(orc.unpack_records(no_shadow)["shadow"] == 0)
    assert np.array_equal(no_shadow["w0"], rec["w0"]) and np.array_equal(no_shadow["t"], rec["t"])


@pytest.mark.parametrize("eye_idx", [0, 2, 5])
def test_dda_matches_reference_instanced_draw(orc, eye_idx):
    """Reference semantics: block-granular sphere, instance list after the hidden-block cull, per-pixel nearest of the
    three camera-facing faces (fp64)  ==  DDA first hit.  Pixels whose ray passes within 1e-4 of a face edge (raster
    tie-break territory) are skipped."""
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_BLOCK)
    table, _, inst = vol.build_occupancy(stamp=5)
    w, h = 64, 36
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[eye_idx], ctr, width=w, height=h)
    rec = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=False)
    u = orc.unpack_records(rec)
    scene = orc.default_scene_config()
    cam_chunk = cam["CameraChunkLocation"][0][:3]
    checked = 0
    for py in range(h):
        for px in range(w):
            hit, blk, face, t, margin, rgba = orc.ref_instanced_pixel(cam, scene, table, inst, w, h, px, py)
            if hit and margin < 1e-4:
                continue
            assert hit == bool(u["hit"][py, px]), (px, py)
            if not hit:
                continue
            checked += 1
            # DDA voxel -> block in camera-chunk-relative block units (the shader's ViewChunkRelativeBlockOffset)
            vox = np.array([u["x"][py, px], u["y"][py, px], u["z"][py, px]], dtype=np.int64)
            dda_blk = (vox >> 3) + (np.array(origin) - cam_chunk) * 16
            assert np.array_equal(dda_blk, blk), (px, py)
            assert int(u["face"][py, px]) == face
            # distance: DDA t is along the normalised ray in voxel units; the restated draw's t is view depth
            v = cam["View"][0].reshape(4, 4); p = cam["Projection"][0].reshape(4, 4)
            fx = (px + 0.5) * 2.0 / w - 1.0; fy = 1.0 - (py + 0.5) * 2.0 / h
            dlen = np.sqrt((fx / p[0][0]) ** 2 + (fy / p[1][1]) ** 2 + 1.0)
            assert float(rec["t"][py, px]) / 8.0 == pytest.approx(t * dlen, rel=1e-5)
            got = [(int(rec["rgba"][py, px]) >> (8 * c)) & 0xFF for c in range(4)]
            exp = [int(np.floor(float(rgba[c]) * 255.0 + 0.5)) for c in range(4)]
            assert all(abs(g - e) <= 1 for g, e in zip(got, exp)), (px, py, got, exp)
            assert got[face >> 1] in ((191, 192) if face & 1 else (63, 64))   # the pinned channel identifies the face
    assert checked > 300


@pytest.mark.parametrize("cfg", [dict(), dict(probe=False), dict(directional=True), dict(df_shift=4), dict(df_shift=4, directional=True),
                                 dict(brick_cap=4), dict(cell2=False), dict(cell2=4), dict(directional=True, brick_cap=8, cell2=2), dict(df_shift=3, df_cap=16, directional=True, brick_cap=0)])
def test_step_model_walks_produce_the_same_records(orc, cfg):
    """ORC_DDA_MODEL (the step-count model of candidate acceleration structures, tools/step_model.py) skips different
    boxes for every configuration and must still produce the records of the plain hierarchical walk, byte for byte --
    which also proves every box its fields certify is really empty."""
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    w, h = 96, 54
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    orc.step_model(vol, **cfg)
    orc.step_model_counts(reset=True)
    for eye in (eyes[1], eyes[6], (ctr[0] + 1.5, ctr[1] + 0.5, ctr[2] + 9.0)):     # two orbit cameras and one close to the surface
        cam = orc.camera_uniform(eye, ctr, width=w, height=h)
        rs = orc.ray_setup(cam, origin, w, h)
        ref = vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_HIER)
        got = vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_MODEL)
        assert got.tobytes() == ref.tobytes()
    counts = orc.step_model_counts()
    assert counts.sum() > 0 and (cfg.get("cell2", True) or counts[1] == 0)


@pytest.mark.parametrize("scene", ["sphere", "terrain", "carved"])
def test_cube_tables_built_with_the_gpu_algorithms_drive_the_same_walk(orc, scene):
    """csrc/k_cubes.cu's three tables, built on the CPU with the same algorithms and layouts (relaxation rounds for the
    cells, shell tests for bricks and 2^3 cells, 2 bits per octant): a walk that READS them produces the records of the
    plain walk and takes exactly the steps of the walk that computes its cubes on the fly -- the builders, the packing and
    the indexing agree with the definition of the cubes."""
    if scene == "terrain":
        origin, dims = (0, -1, 0), (2, 2, 2)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_VOXEL)
    else:
        origin, dims, params = scenes.sphere_scene(256)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
        if scene == "carved":
            vol.carve_sphere((128, 128, 40), 30)
            vol.carve_sphere((60, 128, 128), 45)
    w, h = 96, 54
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    tables = orc.cube_tables(vol)
    assert tables[0].max() <= 32 and (tables[0] > 1).any() and (tables[1] != 0).any() and (tables[2] != 0).any()
    orc.step_model(vol, df_shift=5, df_cap=32, probe=False, directional=True, brick_cap=4, cell2=4)
    try:
        for eye in (eyes[0], eyes[3], eyes[6]):
            cam = orc.camera_uniform(eye, ctr, width=w, height=h)
            rs = orc.ray_setup(cam, origin, w, h)
            ref = vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_HIER)
            orc.cube_tables_use(vol, None)
            orc.step_model_counts(reset=True)
            fly = vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_MODEL)
            c_fly = orc.step_model_counts()
            orc.cube_tables_use(vol, tables)
            tab = vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_MODEL)
            c_tab = orc.step_model_counts()
            assert fly.tobytes() == ref.tobytes() and tab.tobytes() == ref.tobytes()
            assert c_tab.tolist() == c_fly.tolist()
    finally:
        orc.cube_tables_use(vol, None)


def test_cube_tables_stay_valid_after_carves(orc):
    """Tables built BEFORE a carve, used AFTER it: a carve only removes voxels, so every certified cube is still empty; full
    bricks that became partial get new payload slots whose (zero) entries mean "one cell".  Records == plain walk."""
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
    tables = orc.cube_tables(vol, extra_slots=1 << 15)
    slots_before = vol.num_partial_slots()
    for center, radius in (((128, 128, 40), 30), ((60, 128, 128), 45), ((128, 200, 128), 12)):
        vol.carve_sphere(center, radius)
    assert vol.num_partial_slots() > slots_before and vol.num_partial_slots() <= slots_before + (1 << 15)
    w, h = 96, 54
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    orc.step_model(vol, df_shift=5, df_cap=32, probe=False, directional=True, brick_cap=4, cell2=4)
    orc.cube_tables_use(vol, tables)
    try:
        for eye in (eyes[0], eyes[2], eyes[5], (ctr[0] - 2.0, ctr[1], ctr[2] - 9.5)):
            cam = orc.camera_uniform(eye, ctr, width=w, height=h)
            rs = orc.ray_setup(cam, origin, w, h)
            ref = vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_HIER)
            assert vol.raymarch(rs, w, h, shadow=True, mode=orc.DDA_MODEL).tobytes() == ref.tobytes()
    finally:
        orc.cube_tables_use(vol, None)
