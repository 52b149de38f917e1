"""Scratch probe (GPU): where does the end-to-end frame time go?  kernel alone / D2H alone / banded meso_raymarch."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes, camera

n, W, H = 4096, 3840, 2160
origin, dims, params = scenes.sphere_scene(n)
ctx = capi.Context(0)
ctx.scene_create(origin, dims, 1 << 20)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
cams = [camera.camera_uniform(e, ctr, W, H) for e in eyes]
frame = torch.empty((H, W, 4), dtype=torch.int32, device="cuda")
host = torch.empty((H, W, 4), dtype=torch.int32).pin_memory()
host_np = host.numpy().view(capi.HitRecord).reshape(H, W)

def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); ctx.sync()
    t0 = time.perf_counter()
    for k in range(reps): fn(k)
    torch.cuda.synchronize(); ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e3

def kern(k=0):
    ctx.raymarch_device(cams[k % 8], W, H, frame.data_ptr()); ctx.sync()
def copy(k=0):
    host.copy_(frame, non_blocking=True); torch.cuda.synchronize()
print("kernel alone      %.3f ms" % t(kern))
print("D2H alone         %.3f ms" % t(copy))
for b in (1, 2, 4, 8, 16):
    os.environ["MESO_E2E_BANDS"] = str(b)
    print("meso_raymarch bands=%2d  %.3f ms" % (b, t(lambda k=0: ctx.raymarch(cams[k % 8], W, H, out=host_np))))
