"""GPU probe: time the full mesh (passes A, B, C) of the 4096^3 V-sphere with the library MESO_SO names; prints one JSON line with
an order-independent fingerprint of the quad list so that A/B builds can be checked against each other."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
ctx.scene_create(origin, dims, 1 << 20)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
cap = 1 << 25
quads = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
n = ctx.mesh_device(quads.data_ptr(), cap)
ts = []
for _ in range(10):
    ctx.flush_l2()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); ctx.mesh_device(quads.data_ptr(), cap, want_count=False); b.record(s); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
q = quads[:n].to(torch.int64) & 0xFFFFFFFF
fp = int((q * torch.tensor([1, 3, 5, 7], device="cuda", dtype=torch.int64)).sum().item()) & ((1 << 62) - 1)
print(json.dumps({"so": os.environ.get("MESO_SO", "default"), "N": N, "quads": int(n), "fingerprint": fp, "ms_min": min(ts), "ms_med": float(np.median(ts))}))
