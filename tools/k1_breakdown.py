"""One voxelize_sdf call at N^3 for an ncu launch list (per-kernel durations of K1 and its derived-data passes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
ctx.scene_create(origin, dims, 1 << 20)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
ctx.sync()
