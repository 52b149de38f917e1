"""GPU probe: SIMT accounting of the raymarch kernel on bench.py's cfg3 workload (meso_raymarch_stats, STATS build):
steps per ray, steps per hierarchy level, and how much of every warp's longest lane the other lanes use."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes, camera

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W, H = (3840, 2160) if N >= 4096 else (1920, 1080)
origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
ctx.scene_create(origin, dims, 1 << 20)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
eyes, ctr = scenes.orbit_eyes(origin, dims)
rows = []
for i, eye in enumerate(eyes):
    cam = camera.camera_uniform(eye, ctr, W, H)
    st = ctx.raymarch_stats(cam, W, H)
    sp, ss = int(st["steps_primary"]), int(st["steps"]) - int(st["steps_primary"])
    rows.append({"camera": i, "primary": int(st["primary"]), "shadow": int(st["shadow"]), "hits": int(st["hits"]),
                 "steps_per_primary": sp / max(1, int(st["primary"])), "steps_per_shadow": ss / max(1, int(st["shadow"])),
                 "lane_use_primary": sp / max(1, int(st["warp_slots_primary"])), "lane_use_shadow": ss / max(1, int(st["warp_slots_shadow"])),
                 "warp_slots_primary": int(st["warp_slots_primary"]), "warp_slots_shadow": int(st["warp_slots_shadow"]),
                 "level_steps(voxel,cell2,brick,df_le2,df_gt2)": [int(x) for x in st["level_steps"]]})
    print(json.dumps(rows[-1]))
tot = {k: sum(r[k] for r in rows) for k in ("primary", "shadow", "warp_slots_primary", "warp_slots_shadow")}
lv = np.sum([r["level_steps(voxel,cell2,brick,df_le2,df_gt2)"] for r in rows], axis=0)
print(json.dumps({"all_cameras": tot, "level_share": (lv / lv.sum()).round(4).tolist(),
                  "slots_share_shadow": tot["warp_slots_shadow"] / (tot["warp_slots_primary"] + tot["warp_slots_shadow"])}))
