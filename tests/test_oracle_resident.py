"""K6 oracle (resident-set selection): hand-derivable facts and an independent numpy fp32 restatement of
FChunkManageHelper::GetDesiredShowChunkLocationByView / FImportanceComputeInfo (ChunkManagerHelper.h:26-150)."""
import json
import os

import numpy as np
import pytest


F32 = np.float32
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "constants.json")))


def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def np_select_by_view(fwd, F=24, B=6, angle=120.0):
    """Vectorised fp32 restatement, written from the reference text independently of oracle/orc_resident.c."""
    fwd = np.asarray(fwd, dtype=F32)
    r = np.arange(-F, F + 1, dtype=np.int32)
    X, Y, Z = np.meshgrid(r, r, r, indexing="ij")
    off = np.stack([X, Y, Z], -1).reshape(-1, 3)
    o = off.astype(F32)
    with np.errstate(all="ignore"):
        d2 = _dot(o, o)
        ln = np.sqrt(d2)
        inside = ~(ln.astype(np.float64) > F + 1e-6)
        core = (np.abs(off) <= 1).all(-1)
        inv = F32(1.0) / np.sqrt(d2)
        dirs = o * inv[:, None]
        fn = fwd * (F32(1.0) / np.sqrt(_dot(fwd, fwd)))
        thr_view = max(F32(np.cos(F32(F32(angle) * F32(0.01745329251994329576923690768489)) * F32(0.5), dtype=F32)), F32(0.01))
        a = _dot(np.broadcast_to(fn, dirs.shape), dirs)
        a = np.where(a < F32(0), F32(0), a)                     # std::max(dot, 0)
        a = np.where(a > thr_view, F32(1), a / thr_view)
        a = np.minimum(np.maximum(a, F32(0)), F32(1))
        thr = a * F32(F) + (F32(1) - a) * F32(B)
        ok = inside & (core | (ln < thr))
        near = (np.abs(off) <= 2).all(-1)
        ang = _dot(dirs, np.broadcast_to(fwd, dirs.shape))
        ang = np.where(F32(0) < ang, ang, F32(0))               # std::max(0, dot)
        ang = (ang - F32(0.5)) * F32(2)
        ang = np.where(ang < F32(0.75), F32(0.75), ang)
        dist = F32(64) - ln
        dist = np.where(F32(0.25) < dist, dist, F32(0.25))
        imp = np.where(near, F32(1.0e6), ang * dist).astype(F32)
    return off[ok], imp[ok]


def test_scene_config_constants_match_reference():
    c = GOLD["scene_config"]
    assert c["ViewForwardLoadChunkSize"] == 24 and c["ViewBackwardLoadChunkSize"] == 6
    assert c["ViewChunkAngle"] == 120.0 and c["BakeVisibilityViewNum"] == 256
    i = GOLD["importance"]
    assert i == {"Far": 64.0, "Near": 1.0e6, "near_cube": 2, "angle_term": [0.5, 2.0, 0.75], "distance_floor": 0.25}


def test_importance_known_answers(orc):
    fwd = (0.0, 0.0, 1.0)
    assert orc.chunk_importance((0, 0, 0), fwd, (2, -2, 2)) == 1.0e6           # +-2 cube
    assert orc.chunk_importance((5, 5, 5), fwd, (7, 3, 5)) == 1.0e6
    assert orc.chunk_importance((0, 0, 0), fwd, (0, 0, 10)) == 54.0            # max((1-.5)*2,.75)=1 ; 64-10
    assert orc.chunk_importance((0, 0, 0), fwd, (0, 0, -10)) == 0.75 * 54.0    # behind: angle term clamps to .75
    assert orc.chunk_importance((0, 0, 0), fwd, (0, 0, 100)) == 0.25           # beyond Far: distance term clamps to .25
    assert orc.chunk_importance((0, 0, 0), fwd, (3, 0, 0)) == 0.75 * 61.0


def test_block_importance_near_branch_is_dead_in_reference(orc):
    # ChunkManagerHelper.h:55-57 compares int against an unsigned product; no offset passes, so even the camera's own
    # block is scored by the far formula: normalize(0) = NaN -> max(0, NaN) = 0 -> .75 * (1024 - 0)
    fwd = (0.0, 0.0, 1.0)
    v = orc.block_importance((0, 0, 0), fwd, (0, 0, 0), (0, 0, 0))
    assert v == 0.75 * 1024.0
    assert orc.block_importance((0, 0, 0), fwd, (0, 0, 1), (0, 0, 0)) == 1.0 * (1024.0 - 16.0)


@pytest.mark.parametrize("fwd", [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (0.3, -0.8, 0.52), (-2.0, 1.0, 0.5)])
def test_select_by_view_matches_numpy_restatement(orc, fwd):
    got = orc.select_view_chunks(fwd)
    off, imp = np_select_by_view(fwd)
    assert got.shape[0] == off.shape[0]
    key = lambda o: (o[:, 0].astype(np.int64) + 64) * 16384 + (o[:, 1] + 64) * 128 + (o[:, 2] + 64)
    a = np.argsort(key(got["Offset"]))
    b = np.argsort(key(off))
    assert np.array_equal(got["Offset"][a], off[b])
    assert np.array_equal(got["Importance"][a].view(np.uint32), imp[b].view(np.uint32))
    # canonical order: importance descending, ties in loop order
    assert np.all(np.diff(got["Importance"]) <= 0)
    same = np.diff(got["Importance"]) == 0
    assert np.all(np.diff(key(got["Offset"]))[same] > 0)


def test_select_by_view_structure(orc):
    got = orc.select_view_chunks((0.0, 0.0, 1.0))
    off, imp = got["Offset"], got["Importance"]
    assert 10000 < got.shape[0] < 30000                     # SURVEY: "about 15 k" entries per direction
    assert np.all(imp[:125] == 1.0e6) and imp[125] < 1.0e6  # the whole +-2 cube is inside the backward radius 6
    ln = np.sqrt((off.astype(np.float64) ** 2).sum(-1))
    behind = off[:, 2] <= 0
    assert ln[behind].max() < 6.0                           # dot <= 0 -> backward radius
    assert ln.max() < 24.0 and ln[~behind].max() > 23.0     # in-cone chunks reach the forward radius
    sel = {tuple(o) for o in off.tolist()}
    assert (0, 0, 23) in sel and (0, 0, 24) not in sel and (0, 0, -5) in sel and (0, 0, -6) not in sel


def test_select_simple_mode(orc):
    got = orc.select_view_chunks((0.0, 0.0, 1.0), mode=1)
    r = np.arange(-24, 25)
    X, Y, Z = np.meshgrid(r, r, r, indexing="ij")
    ln = np.sqrt((X * X + Y * Y + Z * Z).astype(np.float32))
    assert got.shape[0] == int((ln < 24).sum())
    assert np.all(got["Importance"][:27] == 1.0e6)
    assert got["Importance"][27] == 0.5 and tuple(got["Offset"][27]) == (-2, 0, 0)   # nearest non-core: 1 / |(+-2,0,0)|


def test_small_radius_and_degenerate_forward(orc):
    got = orc.select_view_chunks((0.0, 1.0, 0.0), forward_load=2, backward_load=1)
    assert np.all(got["Importance"] == 1.0e6)
    assert {tuple(o) for o in got["Offset"].tolist()} >= {(x, y, z) for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)}
    # zero forward vector: normalize -> NaN, std::max/min keep the NaN, `len < NaN` is false -> only the 3^3 core
    z = orc.select_view_chunks((0.0, 0.0, 0.0))
    off, _ = np_select_by_view((0.0, 0.0, 0.0))
    assert z.shape[0] == off.shape[0] == 27


def test_abi_baked_direction_is_a_host_function_and_matches_oracle(orc):
    """meso_baked_direction needs no device: GetFibonacciSphere<float> + nearest, bit-identical to the oracle."""
    from mesoengine_b200 import capi
    dirs = orc.fibonacci_sphere_f32(256)
    rng = np.random.default_rng(5)
    for q in rng.normal(size=(64, 3)).astype(np.float32):
        d, idx = capi.baked_direction(256, q)
        assert idx == orc.nearest_direction(dirs, q)
        assert np.array_equal(d.view(np.uint32), dirs[idx].view(np.uint32))
    with pytest.raises(capi.MesoError):
        capi.baked_direction(1, (0, 0, 1))


def test_fibonacci_f32_and_nearest(orc):
    d = orc.fibonacci_sphere_f32(256)
    assert d.dtype == np.float32 and d.shape == (256, 3)
    assert np.allclose(np.linalg.norm(d.astype(np.float64), axis=1), 1.0, atol=1e-6)
    assert d[0, 1] == 1.0 and d[-1, 1] == -1.0
    for k in (0, 1, 17, 128, 255):
        assert orc.nearest_direction(d, d[k]) == k
    assert orc.nearest_direction(d, (0.0, 10.0, 0.0)) == 0
