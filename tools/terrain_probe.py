"""GPU probe: the reference's own scene (TestGenerator terrain) through K1 -> K4 at 1024 x 512 x 1024 voxels, 1080p, primary
+ shadow, block-granular (reference semantics) and voxel-granular; CUDA events, 8 orbit cameras.  Not a bench.py value."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes, camera

W, H = 1920, 1080
origin, dims, _ = scenes.terrain_scene(1024, height_chunks=4)
ctx = capi.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
out = {}
for name, gran in (("blocks", capi.GRAN_BLOCK), ("voxels", capi.GRAN_VOXEL)):
    ctx.scene_create(origin, dims, 1 << 21)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); ctx.voxelize_sdf(capi.SDF_TERRAIN, None, gran); b.record(s); torch.cuda.synchronize()
    t_vox = a.elapsed_time(b)
    eyes, ctr = scenes.orbit_eyes(origin, dims)
    cams = [camera.camera_uniform(e, ctr, W, H) for e in eyes]
    frame = torch.empty((H, W, 4), dtype=torch.int32, device="cuda")
    rays = []
    for c in cams:
        st = ctx.raymarch_stats(c, W, H)
        rays.append(int(st["primary"]) + int(st["shadow"]))
    for c in cams:
        ctx.raymarch_device(c, W, H, frame.data_ptr())
    ts = []
    for c in cams:
        ctx.flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); ctx.raymarch_device(c, W, H, frame.data_ptr()); b.record(s); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    occ, full, keys, _ = ctx.volume_download()
    out[name] = {"voxelize_ms": t_vox, "blocks": int(np.unpackbits(occ.view(np.uint8)).sum()), "partial_bricks": int(len(keys)),
                 "ms_per_frame": float(np.mean(ts)), "Mrays_per_s": float(sum(rays) / (sum(ts) * 1e-3) / 1e6),
                 "rays_per_frame": float(np.mean(rays))}
print(json.dumps(out))
