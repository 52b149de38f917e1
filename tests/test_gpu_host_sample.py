"""GPU: the C++ host mirror (mesoengine_b200/host: headless SimpleVoxel over the C ABI) produces the same frame as the
Python plumbing and as the oracle for the reference configuration (GenerateSphere, one sample per block)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mesoengine_b200", "host", "SimpleVoxel")


def test_cpp_sample_matches_python_path_and_oracle(orc, tmp_path):
    from mesoengine_b200 import camera, capi
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.dirname(EXE), "CXX=g++"])
    w, h = 320, 180
    eye, target = (20.5, -61.25, 33.0), (100.0, 0.0, 0.0)
    out = tmp_path / "frame.bin"
    args = [EXE, "3", str(w), str(h)] + [repr(float(v)) for v in eye + target] + [str(out)]
    log = subprocess.check_output(args, text=True)
    assert "blocks=201936" in log                      # instances after the hidden-block cull of the reference sphere
    got = np.fromfile(out, dtype=capi.HitRecord).reshape(h, w)

    origin, dims = (2, -4, -4), (8, 8, 8)
    cam = camera.camera_uniform(eye, target, w, h)
    ctx = capi.Context(0)
    ctx.scene_create(origin, dims, 1 << 16)
    ctx.voxelize_sdf(capi.SDF_SPHERE, orc.REF_SPHERE, capi.GRAN_BLOCK)
    rec = ctx.raymarch(cam, w, h, shadow=True)
    ctx.close()
    assert got.tobytes() == rec.tobytes()
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, orc.REF_SPHERE, granularity=orc.GRAN_BLOCK)
    ref = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True)
    assert got.tobytes() == ref.tobytes()
    assert int(((got["w1"] >> 20) & 1).sum()) > 1000


@pytest.mark.parametrize("frames", [1, 3])
def test_cpp_sample_streaming_matches_python_streaming(frames, tmp_path):
    """--stream: FChunkManage::UpdateChunks / UpdateLoadingQueue over meso_stream_update (256 chunks per frame, baked view
    direction).  After one frame only part of the window is generated; the frame must still equal the Python path's."""
    from mesoengine_b200 import camera, capi
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.dirname(EXE), "CXX=g++"])
    w, h = 320, 180
    eye, target = (20.5, -61.25, 33.0), (100.0, 0.0, 0.0)
    out = tmp_path / "frame.bin"
    args = [EXE, "--stream", str(frames), str(w), str(h)] + [repr(float(v)) for v in eye + target] + [str(out)]
    log = subprocess.check_output(args, text=True)
    got = np.fromfile(out, dtype=capi.HitRecord).reshape(h, w)

    origin, dims = (2, -4, -4), (8, 8, 8)
    cam = camera.camera_uniform(eye, target, w, h)
    view = np.asarray(cam["View"]).reshape(-1)
    fwd = (-view[0 * 4 + 2], -view[1 * 4 + 2], -view[2 * 4 + 2])          # FVoxelCamera::GetForwardVector
    cam_chunk = np.asarray(cam["CameraChunkLocation"]).reshape(-1)[:3]
    d, _ = capi.baked_direction(256, fwd)
    ctx = capi.Context(0)
    ctx.scene_create(origin, dims, 1 << 16)
    ctx.stream_begin(capi.SDF_SPHERE, (100.0, 0.0, 0.0, 50.0), capi.GRAN_BLOCK)
    loaded = 0
    for _ in range(frames):
        st = ctx.stream_update(cam_chunk, d, 256)
        loaded += int(st["generated"])
    rec = ctx.raymarch(cam, w, h, shadow=True)
    n_loaded = int(ctx.stream_loaded(512).sum())
    ctx.close()
    assert "loaded=%d missing=%d" % (loaded, int(st["missing"])) in log
    assert n_loaded == loaded and (frames > 1 or loaded == 256)
    assert got.tobytes() == rec.tobytes()


def test_cpp_sample_cpu_generator_callback_equals_device_generator(tmp_path):
    """--cpu-generator: the reference's std::function GeneratorType callback (FGeneratorHelper::GenerateSphere on host
    worker threads) -> FChunk.Blocks -> FGPUChunk / FGPUBlock records -> meso_volume_upload_blocks.  Same instances
    after the device cull and the same frame as the device generator (meso_voxelize_sdf)."""
    from mesoengine_b200 import camera, capi
    subprocess.check_call(["make", "-C", os.path.dirname(EXE), "CXX=g++"])
    w, h = 320, 180
    eye, target = (20.5, -61.25, 33.0), (100.0, 0.0, 0.0)
    frames = {}
    for mode in ("--cpu-generator", None):
        out = tmp_path / ("frame%s.bin" % (mode or ""))
        args = [EXE] + ([mode] if mode else []) + ["2", str(w), str(h)] + [repr(float(v)) for v in eye + target] + [str(out)]
        log = subprocess.check_output(args, text=True)
        assert "blocks=201936" in log, log
        frames[mode] = np.fromfile(out, dtype=capi.HitRecord).reshape(h, w)
    assert frames["--cpu-generator"].tobytes() == frames[None].tobytes()


@pytest.mark.parametrize("gpus", [2, 3])
def test_cpp_sample_on_a_group_of_gpus_equals_one_gpu(gpus, tmp_path):
    """--gpus N: the C++ host renders through meso_group_* (replicated volume, slab gather into host memory).  Members share
    devices when the box has fewer GPUs than N.  Same frame as one GPU."""
    import torch
    from mesoengine_b200 import capi
    subprocess.check_call(["make", "-C", os.path.dirname(EXE), "CXX=g++"])
    w, h = 320, 180
    eye, target = (20.5, -61.25, 33.0), (100.0, 0.0, 0.0)
    env = dict(os.environ, MESO_SAMPLE_DEVICE_COUNT=str(torch.cuda.device_count()))
    frames = {}
    for mode in (["--gpus", str(gpus)], []):
        out = tmp_path / ("frame%d.bin" % len(mode))
        args = [EXE] + mode + ["2", str(w), str(h)] + [repr(float(v)) for v in eye + target] + [str(out)]
        log = subprocess.check_output(args, text=True, env=env)
        assert "blocks=201936" in log, log
        frames[len(mode)] = np.fromfile(out, dtype=capi.HitRecord).reshape(h, w)
    assert frames[2].tobytes() == frames[0].tobytes()
