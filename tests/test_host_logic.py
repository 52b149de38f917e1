"""CPU-only: host-side mirrors (camera, ray setup, tile partition) against the oracle / plain numpy."""
import numpy as np
import pytest

import scenes


def test_camera_mirror_matches_oracle(orc):
    from mesoengine_b200 import camera
    for n in (256, 1024, 4096):
        origin, dims, _ = scenes.sphere_scene(n)
        eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
        for e in eyes + [(5.0, 2.0, 2.0)]:
            for (w, h) in ((3840, 2160), (1280, 720), (101, 37)):
                assert orc.camera_uniform(e, ctr, width=w, height=h).tobytes() == camera.camera_uniform(e, ctr, w, h).tobytes()


def test_camera_callbacks_and_recentering():
    from mesoengine_b200.camera import FVoxelCamera
    cam = FVoxelCamera((5.0, 2.0, 2.0), (0.0, 0.0, 0.0))
    cam.InitializeVoxelCamera(60.0, 0.1, 1000.0, True)
    calls = []
    cam.CameraChunkUpdateCallback = lambda: calls.append("chunk")
    cam.CameraUpdateCallback = lambda: calls.append("view")
    cam.UpdateCamera(16.0)
    assert calls == ["view"]                       # first forward vector, no chunk change (VoxelCamera.cpp:42-59)
    cam.Position = np.array([17.5, -3.0, 2.0], dtype=np.float32)
    cam.UpdateCamera(16.0)
    assert calls == ["view", "chunk", "view"]
    assert list(cam.CameraChunkLocation) == [1, -1, 0] and np.allclose(cam.Position, [1.5, 13.0, 2.0])
    u = cam.GetCameraUniform(1280.0, 720.0)
    assert list(u["SubCameraLocation"][0][:3]) == [1.0, 13.0, 2.0]


def test_ray_setup_abi_matches_oracle(orc):
    from mesoengine_b200 import capi
    origin, dims, _ = scenes.sphere_scene(1024)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    for e in eyes:
        cam = orc.camera_uniform(e, ctr, width=1920, height=1080)
        for light in ((0.3, 0.5, 0.8), (-1.0, 0.2, 0.1)):
            assert orc.ray_setup(cam, origin, 1920, 1080, light).tobytes() == capi.ray_setup(cam, origin, 1920, 1080, light).tobytes()


@pytest.mark.parametrize("wh", [(3840, 2160), (320, 184), (101, 37)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_partition_roundtrip(wh, world):
    from mesoengine_b200 import capi, partition
    w, h = wh
    assert partition.tiles_per_rank(w, h, world) == capi.tiles_per_rank(w, h, world)
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 2 ** 31, (h, w, 4), dtype=np.int64).astype(np.uint32)
    gathered = np.stack([partition.pack_tiles(frame, r, world) for r in range(world)])
    assert np.array_equal(partition.compose_tiles(gathered, w, h), frame)
    tiles = np.concatenate([partition.rank_tiles(w, h, r, world) for r in range(world)])
    tx, ty = partition.tile_grid(w, h)
    assert sorted(tiles.tolist()) == list(range(tx * ty))


def test_scenes_are_inside_their_grids():
    for n in (256, 512, 1024, 4096):
        origin, dims, (cx, cy, cz, r) = scenes.sphere_scene(n)
        for c, o, d in zip((cx, cy, cz), origin, dims):
            assert o * 16 <= c - r and c + r <= (o + d) * 16
    assert scenes.sphere_scene(1024)[2] == scenes.REF_SPHERE
