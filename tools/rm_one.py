"""One raymarch launch per configuration named in RM_ONE (comma list of v10, v10+cubes; MESO_CUBES_LEVEL picks the level) on the bench
workload, camera 0 -- the command ncu wraps (-k regex:raymarch)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mesoengine_b200 import camera, capi, scenes
N, W, H = int(os.environ.get("RM_ONE_N", "4096")), int(os.environ.get("RM_ONE_W", "3840")), int(os.environ.get("RM_ONE_H", "2160"))
origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
ctx.scene_create(origin, dims, max_bricks=(1 << 20) if N >= 4096 else (1 << 18))
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
cam = camera.camera_uniform(eyes[int(os.environ.get("RM_ONE_CAM", "0"))], ctr, W, H)
frame = torch.empty((H, W, 4), dtype=torch.int32, device="cuda")
for name in os.environ.get("RM_ONE", "v10+cubes").split(","):
    ctx.raymarch_device(cam, W, H, frame.data_ptr(), shadow=True, layout=capi.LAYOUT_FRAME, flags_extra=capi.FLAG_CUBES if "+cubes" in name else capi.FLAG_NO_CUBES)
    ctx.sync()
ctx.close()
