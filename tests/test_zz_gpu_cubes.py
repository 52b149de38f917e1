"""The per-octant forward-cube tables of the raymarch walk (csrc/k_cubes.cu; the default walk whenever they are current).

What these tests pin: (1) frames through the cubes are byte-identical to the oracle's (any certified-empty box is a legal skip, so
a wrong table shows up as a wrong record); (2) the kernel takes the steps the oracle's step model takes with the same cubes
(directional cells cap 32, brick cubes <= 4, 2^3-cell cubes <= 4) within 0.2 % -- the kernel does not look the cell table
up again while a ray stays inside one 32^3 cell, the model does at every step, so a few steps differ in kind; (3) the tables survive carves (which only remove voxels) and are invalidated by every call that may add
voxels."""
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


def _scene(ctx, orc, kind, origin, dims, params, gran):
    ctx.scene_create(origin, dims, 1 << 18)
    ctx.voxelize_sdf(kind, params, gran)
    return orc.Volume(origin, dims).voxelize(kind, params, granularity=gran, sin_mode=orc.SIN_PORTABLE)


def _frames_equal(ctx, orc, vol, origin, dims, w, h, eyes_extra=()):
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    for eye in list(eyes) + list(eyes_extra):
        cam = orc.camera_uniform(eye, ctr, width=w, height=h)
        ref = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True)
        got = ctx.raymarch(cam, w, h, shadow=True, cubes=True)
        assert got.tobytes() == ref.tobytes(), eye
        assert ctx.raymarch(cam, w, h, shadow=True, cubes=False).tobytes() == ref.tobytes()


def test_cubes_frames_equal_oracle_sphere(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _scene(ctx, orc, orc.SDF_SPHERE, origin, dims, params, orc.GRAN_VOXEL)
    ctx.build_cubes()
    ctr = scenes.grid_center_world(origin, dims)
    _frames_equal(ctx, orc, vol, origin, dims, 160, 96, eyes_extra=[(ctr[0] + 1.5, ctr[1] + 0.5, ctr[2] + 9.0)])


def test_cubes_frames_equal_oracle_terrain(ctx, orc):
    origin, dims = (0, -1, 0), (2, 2, 2)
    vol = _scene(ctx, orc, orc.SDF_TERRAIN, origin, dims, None, orc.GRAN_VOXEL)
    _frames_equal(ctx, orc, vol, origin, dims, 128, 72)


@pytest.mark.parametrize("scene", ["sphere", "terrain"])
def test_cube_tables_equal_the_cpu_tables(ctx, orc, scene):
    """cell and brick tables (independent of payload-slot order) against oracle/orc_raymarch.c:orc_cube_tables, bit for bit."""
    if scene == "sphere":
        origin, dims, params = scenes.sphere_scene(256)
        vol = _scene(ctx, orc, orc.SDF_SPHERE, origin, dims, params, orc.GRAN_VOXEL)
    else:
        origin, dims = (0, -1, 0), (2, 2, 3)
        vol = _scene(ctx, orc, orc.SDF_TERRAIN, origin, dims, None, orc.GRAN_VOXEL)
    ctx.build_cubes()
    cell, brick = ctx.download_cubes()
    ref_cell, ref_brick, _ = orc.cube_tables(vol)
    assert np.array_equal(cell, ref_cell)
    assert np.array_equal(brick, ref_brick)


def test_cubes_steps_match_the_step_model(ctx, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _scene(ctx, orc, orc.SDF_SPHERE, origin, dims, params, orc.GRAN_VOXEL)
    orc.step_model(vol, df_shift=5, df_cap=32, probe=False, directional=True, brick_cap=4, cell2=4)
    w, h = 160, 96
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    for eye in eyes[:4]:
        cam = orc.camera_uniform(eye, ctr, width=w, height=h)
        orc.step_model_counts(reset=True)
        vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True, mode=orc.DDA_MODEL)
        model = orc.step_model_counts()
        st = ctx.raymarch_stats(cam, w, h, shadow=True, cubes=True)
        assert abs(int(st["steps"]) - int(model.sum())) <= 0.002 * int(model.sum()), (eye, st, model)
        field = ctx.raymarch_stats(cam, w, h, shadow=True, cubes=False)
        assert int(st["steps"]) < int(field["steps"])


def test_cubes_survive_carves_and_are_invalidated_when_voxels_may_be_added(ctx, capi, orc):
    origin, dims, params = scenes.sphere_scene(256)
    vol = _scene(ctx, orc, orc.SDF_SPHERE, origin, dims, params, orc.GRAN_VOXEL)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[2], ctr, width=96, height=60)
    # built with the volume; a stream update may add voxels anywhere: the tables are gone until they are rebuilt
    ref0 = vol.raymarch(orc.ray_setup(cam, origin, 96, 60), 96, 60, shadow=True)
    assert ctx.raymarch(cam, 96, 60, shadow=True, cubes=True).tobytes() == ref0.tobytes()
    # carves only remove voxels: the tables stay valid (conservative), incl. full bricks that become partial (new slots)
    for center, radius in (((128, 128, 40), 30), ((60, 128, 128), 45), ((128, 200, 128), 12)):
        ctx.carve_sphere(center, radius)
        vol.carve_sphere(center, radius)
        for eye in (eyes[2], eyes[5]):
            cam = orc.camera_uniform(eye, ctr, width=96, height=60)
            ref = vol.raymarch(orc.ray_setup(cam, origin, 96, 60), 96, 60, shadow=True)
            assert ctx.raymarch(cam, 96, 60, shadow=True, cubes=True).tobytes() == ref.tobytes()
    ctx.build_cubes()                                     # a rebuild after the edits gives the same frames with longer steps
    ref = vol.raymarch(orc.ray_setup(cam, origin, 96, 60), 96, 60, shadow=True)
    assert ctx.raymarch(cam, 96, 60, shadow=True, cubes=True).tobytes() == ref.tobytes()
    ctx.stream_begin(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)      # voxels may be added piecemeal from here on: the tables are gone
    with pytest.raises(capi.MesoError):
        ctx.raymarch(cam, 96, 60, cubes=True)
    ctx.raymarch(cam, 96, 60)                                        # the default falls back to the distance field
    with pytest.raises(capi.MesoError):
        ctx.raymarch_device(cam, 96, 60, 1, flags_extra=capi.FLAG_CUBES | capi.FLAG_NO_CUBES)   # exclusive flags: rejected before any launch
