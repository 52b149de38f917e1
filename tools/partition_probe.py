"""One GPU playing every rank of an N-GPU job in turn: ms per frame of rank r's tile subset in the 4-frames-in-flight ring
(no rendezvous, no peers).  Separates what the partition itself costs (imbalance between ranks, the per-launch floor, the
loss of cache locality) from what the exchange costs in the real N-GPU run."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mesoengine_b200 import camera, capi, scenes
N, W, H = 4096, 3840, 2160
origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.scene_create(origin, dims, max_bricks=1 << 20)
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
cams = [camera.camera_uniform(e, ctr, W, H) for e in eyes]
R = 4
streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(R - 1)]
frames = [torch.empty((H, W, 4), dtype=torch.int32, device=dev) for _ in range(R)]
out = {}
for world in (1, 2, 4, 8):
    per_rank = []
    for rank in range(world):
        ctx.set_partition(rank, world)
        def step(k):
            s = streams[k % R]
            ctx.set_stream(s.cuda_stream)
            ctx.raymarch_device(cams[k % 8], W, H, frames[k % R].data_ptr(), shadow=True, layout=capi.LAYOUT_FRAME)
            ctx.set_stream(stream.cuda_stream)
        for k in range(8): step(k)
        torch.cuda.synchronize()
        K = 80
        ev0 = torch.cuda.Event(enable_timing=True); ev0.record(stream)
        for s in streams[1:]: s.wait_event(ev0)
        for k in range(K): step(k)
        evs = []
        for s in streams:
            e = torch.cuda.Event(enable_timing=True); e.record(s); evs.append(e)
        torch.cuda.synchronize()
        per_rank.append(max(ev0.elapsed_time(e) for e in evs) / K)
    out[str(world)] = {"ms_per_frame_by_rank": per_rank, "max": max(per_rank), "mean": float(np.mean(per_rank)), "ideal": None}
base = out["1"]["max"]
for w in out: out[w]["ideal"] = base / int(w)
print(json.dumps(out))
