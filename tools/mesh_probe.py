"""GPU probe: one meshing pass on the 4096^3 V-sphere (used under ncu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
ctx.scene_create(origin, dims, 1 << 20)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
cap = 1 << 25
quads = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
for _ in range(3):
    n = ctx.mesh_device(quads.data_ptr(), cap)
print("quads", n)
