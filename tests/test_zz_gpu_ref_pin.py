"""K1 + K2 on the GPU against rows produced by the REFERENCE'S OWN CODE (tests/golden/ref_build.npz, see
tests/test_ref_pin.py): block lists in generator order, the four erode mips, the hidden-block cull, the emitted
FGPUBlock instances and the FGPUChunk table of every chunk of the reference sphere (448 chunks, 523 155 blocks,
201 936 instances) and of 48 terrain chunks -- no oracle in between.  The checker is the one the CPU suite feeds with
the oracle's outputs (refprobe.check_grid_against_golden)."""
import numpy as np
import pytest

import refprobe

pytestmark = pytest.mark.gpu

GOLD = dict(np.load(refprobe.GOLDEN))


@pytest.fixture(scope="module")
def ctx():
    from mesoengine_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _run(ctx, origin, dims, kind, params, stamp):
    from mesoengine_b200 import capi
    ctx.scene_create(origin, dims, 1 << 16)
    ctx.voxelize_sdf(kind, params, capi.GRAN_BLOCK)
    occ = ctx.volume_download()[0]
    n = ctx.build_occupancy(stamp=stamp)
    table, mips, inst = ctx.download_occupancy(n)
    return occ, mips, table, inst[:n]


def test_reference_sphere_on_gpu_equals_reference_rows(ctx):
    from mesoengine_b200 import capi, scenes
    origin, dims = (2, -4, -4), (8, 8, 8)
    occ, mips, table, inst = _run(ctx, origin, dims, capi.SDF_SPHERE, scenes.REF_SPHERE, 9)
    assert len(inst) == 201936
    assert refprobe.check_grid_against_golden(GOLD, "sphere", refprobe.SPHERE_CHUNKS, origin, dims, occ, mips, table, inst, 9) == 448


def test_terrain_on_gpu_equals_reference_rows(ctx):
    """The device evaluates TestGenerator's hash with the portable sin; on these 48 chunks
    (196 608 samples, 98 057 solid) every block comes out as the reference's libm build decided it."""
    from mesoengine_b200 import capi
    origin, dims = refprobe.TERRAIN_GRID
    occ, mips, table, inst = _run(ctx, origin, dims, capi.SDF_TERRAIN, None, 5)
    assert refprobe.check_grid_against_golden(GOLD, "terrain", refprobe.TERRAIN_CHUNKS, origin, dims, occ, mips, table, inst, 5) == 48


DRAW = dict(np.load(refprobe.DRAW_GOLDEN))


@pytest.mark.parametrize("eye_idx", refprobe.DRAW_EYES)
def test_raymarch_on_gpu_against_the_executed_reference_shaders(ctx, eye_idx):
    """K4 against the frame the reference's own vertex / fragment shader text produced (tests/golden/ref_draw.npz, see
    oracle/ref_glsl_driver.cpp): same hit / miss, block, face, colour within 1 LSB away from face edges."""
    from mesoengine_b200 import capi, scenes
    origin, dims, params = scenes.sphere_scene(256)
    ctx.scene_create(origin, dims, 1 << 16)
    ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_BLOCK)
    cam = np.frombuffer(DRAW[f"eye{eye_idx}_camera"].tobytes(), dtype=capi.Camera).copy()
    rec = ctx.raymarch(cam, refprobe.DRAW_W, refprobe.DRAW_H, shadow=False)
    hits, misses, skipped = refprobe.check_records_against_ref_draw(DRAW, eye_idx, rec, origin)
    assert hits > 15000 and misses > 12000 and skipped < 0.03 * rec.size, (hits, misses, skipped)   # 256 x 144: 1.2-1.6 % of the pixels lie within 1e-3 of a face edge


@pytest.mark.parametrize("i,mode", [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (0, 1), (1, 1)])
def test_select_view_on_gpu_equals_reference_queue(ctx, i, mode):
    """K6 selection against FChunkManageHelper::GetDesiredShowChunkLocationByView / ...Simple as the reference build ran
    them: same candidate set with the same importance bits, and the same importance sequence in pop order (the order
    among equal importances is unspecified in std::priority_queue)."""
    from mesoengine_b200 import capi
    got = ctx.select_view_chunks(refprobe.VIEWS[i], capi.view_config(24, 6, 120.0, mode))
    assert len(got) == int(GOLD[f"view{i}_mode{mode}_count"][0])
    canon = got[np.lexsort((got["Offset"][:, 2], got["Offset"][:, 1], got["Offset"][:, 0]))]
    assert canon.dtype.itemsize == 16
    assert np.array_equal(refprobe._sha(canon), GOLD[f"view{i}_mode{mode}_set_sha1"])
    assert np.array_equal(refprobe._sha(got["Importance"].copy()), GOLD[f"view{i}_mode{mode}_pop_importance_sha1"])


def test_chunk_importance_on_gpu_equals_reference(ctx):
    """K6 importance against FImportanceComputeInfo::CalculateChunkImportance as the reference build ran it
    (30 cameras x 50 chunks, incl. the camera's own chunk where the reference relies on max(0, NaN) = 0)."""
    cam, fwd, loc = GOLD["chunk_importance_cam"], GOLD["chunk_importance_fwd"], GOLD["chunk_importance_loc"]
    want = GOLD["chunk_importance"]
    for g in range(0, len(want), 50):
        got = ctx.chunk_importance(cam[g], fwd[g], loc[g:g + 50])
        assert np.array_equal(got.view(np.uint32), want[g:g + 50].view(np.uint32)), g


def test_block_importance_on_gpu_equals_reference(ctx):
    """FImportanceComputeInfo::CalculateBlockImportance as the reference build ran it (tests/golden/ref_build.npz) against
    the device kernel, bit for bit -- including the dead near branch the reference's unsigned comparison produces."""
    cam, fwd = GOLD["chunk_importance_cam"], GOLD["chunk_importance_fwd"]
    chunk, block, want = GOLD["block_importance_chunk"], GOLD["block_importance_block"], GOLD["block_importance"]
    for g in range(0, len(want), 50):      # 30 cameras x 50 blocks
        got = ctx.block_importance(cam[g], fwd[g], chunk[g:g + 50], block[g:g + 50])
        assert np.array_equal(got.view(np.uint32), want[g:g + 50].view(np.uint32))
