"""GPU: N GPUs behind one handle (meso_group_*, csrc/meso_group.cu).  Members may share a device, so the slab gather, the
prefix-offset quad gather and the sharded dirty re-mesh are exercised on one GPU too (devices [0, 0, 0]); with two or more
GPUs in the box the same tests run across real peers."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from mesoengine_b200 import capi as _capi
    return _capi


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.mark.parametrize("devices", _device_lists() if __import__("torch").cuda.is_available() else [[0]])
def test_group_frame_mesh_edit_equal_single_gpu_and_oracle(capi, orc, devices):
    origin, dims, params = scenes.sphere_scene(256)
    g = capi.Group(devices)
    one = capi.Context(devices[0])
    try:
        g.scene_create(origin, dims, 1 << 16)
        g.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
        one.scene_create(origin, dims, 1 << 16)
        one.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL)
        eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
        # frames: ragged size (not a multiple of the tile, slabs of unequal height), records and RGBA8
        for (w, h) in ((200, 117), (320, 184)):
            for eye in (eyes[1], eyes[6]):
                cam = orc.camera_uniform(eye, ctr, width=w, height=h)
                ref = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True)
                got = g.raymarch(cam, w, h, shadow=True)
                assert got.tobytes() == ref.tobytes()
                assert got.tobytes() == one.raymarch(cam, w, h, shadow=True).tobytes()
                img = g.raymarch(cam, w, h, shadow=True, rgba8=True)
                assert np.array_equal(img, ref["rgba"])
        # frame ring: four frames in flight, each equal to its synchronous frame
        w, h = 256, 144
        outs = [np.zeros((h, w), dtype=capi.HitRecord) for _ in range(4)]
        cams = [orc.camera_uniform(eyes[k], ctr, width=w, height=h) for k in range(4)]
        for k in range(4):
            g.raymarch_async(cams[k], w, h, outs[k], k, shadow=True)
        for k in range(4):
            g.frame_wait(k)
            assert outs[k].tobytes() == one.raymarch(cams[k], w, h, shadow=True).tobytes()
        # quads: concatenation of the members' lists == the single-GPU mesh == the oracle's, after the canonical sort
        ref_q = orc.sort_quads(vol.mesh())
        q, counts = g.mesh(len(ref_q) + 64)
        assert int(counts.sum()) == len(ref_q) and (counts > 0).all()
        assert orc.sort_quads(q).tobytes() == ref_q.tobytes()
        # device-resident gather on member 0: segments written by the members' kernels, then compacted in place
        import torch
        cap = 2 * len(ref_q) + 3 * len(devices)
        with torch.cuda.device(devices[0]):
            dq = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
            n, seg = g.mesh_device(dq.data_ptr(), cap, compact=True)
            torch.cuda.synchronize()
            got_q = dq[:n].cpu().numpy().view(capi.Quad).reshape(-1)
        assert n == len(ref_q) and np.array_equal(seg, counts)
        assert orc.sort_quads(got_q).tobytes() == ref_q.tobytes()
        # edit: replicated carve, re-mesh sharded by key, the frame afterwards
        centre = (int(dims[0] * 64 - 90), int(dims[1] * 64), int(dims[2] * 64))
        nd = g.carve_sphere(centre, 20)
        assert nd == one.carve_sphere(centre, 20)
        dirty = vol.carve_sphere(centre, 20)
        assert nd == len(dirty)
        rq = g.remesh_dirty(1 << 16)
        oq, _ = one.remesh_dirty(1 << 16, 1 << 13)
        assert orc.sort_quads(rq).tobytes() == orc.sort_quads(oq).tobytes()
        cam = orc.camera_uniform(eyes[2], ctr, width=w, height=h)
        ref = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=True)
        assert g.raymarch(cam, w, h, shadow=True).tobytes() == ref.tobytes()
        # a member is an ordinary context between group calls
        m1 = g.member(1)
        occ, full, keys, payload = m1.volume_download()
        assert np.array_equal(occ, vol.occ()) and np.array_equal(full, vol.full())
    finally:
        one.close()
        g.close()


def test_slab_layout_through_the_single_context_entry_point(capi, orc):
    """meso_raymarch_device_slabs on one context with world = 1: three slabs in separate buffers reassemble the frame."""
    import torch
    origin, dims, params = scenes.sphere_scene(256)
    ctx = capi.Context(0)
    try:
        ctx.scene_create(origin, dims, 1 << 16)
        ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
        eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
        w, h, rows = 208, 100, 40
        cam = orc.camera_uniform(eyes[3], ctr, width=w, height=h)
        slabs = [torch.zeros((rows, w, 4), dtype=torch.int32, device="cuda") for _ in range(3)]
        ctx.raymarch_device_slabs(cam, w, h, [t.data_ptr() for t in slabs], rows, shadow=True)
        ctx.sync()
        got = torch.cat(slabs, dim=0)[:h].cpu().numpy().view(capi.HitRecord).reshape(h, w)
        assert got.tobytes() == ctx.raymarch(cam, w, h, shadow=True).tobytes()
        with pytest.raises(capi.MesoError):
            ctx.raymarch_device_slabs(cam, w, h, [t.data_ptr() for t in slabs], 36, shadow=True)      # not a multiple of the tile height
        with pytest.raises(capi.MesoError):
            ctx.raymarch_device_slabs(cam, w, h, [t.data_ptr() for t in slabs[:2]], rows, shadow=True)  # does not cover the frame
    finally:
        ctx.close()
