"""GPU A/B of raymarch kernel variants on the bench workload (4096^3 V-sphere, 3840x2160, primary + shadow, 8 cameras).

One process per library build (MESO_SO) and cube level (MESO_CUBES_LEVEL=1..3, read once per process); inside it the walk
is switched per launch between the distance field (MESO_FLAG_NO_CUBES) and the forward cubes.  For every configuration: kernel alone with L2 flushed,
the 4-frames-in-flight loop of bench.py, step counters, and a frame hash that must equal the first configuration's.

    python tools/rm_ab.py [N=4096] [W=3840] [H=2160]            # appends JSON lines to gpurun_out/rm_ab.jsonl
"""
import hashlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mesoengine_b200 import camera, capi, scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
H = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
LIGHT = (0.3, 0.5, 0.8)
CONFIGS = [("v10", "v10", 0, False), ("v10+cubes", "v10", 0, True)]
if os.environ.get("RM_AB_CONFIGS"):
    keep = os.environ["RM_AB_CONFIGS"].split(",")
    CONFIGS = [c for c in CONFIGS if c[0] in keep]

origin, dims, params = scenes.sphere_scene(N)
ctx = capi.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.scene_create(origin, dims, max_bricks=(1 << 20) if N >= 4096 else (1 << 18))
ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
ctx.sync(); t0 = time.perf_counter(); ctx.build_cubes(); ctx.sync(); t_cubes = time.perf_counter() - t0
eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
cams = [camera.camera_uniform(e, ctr, W, H) for e in eyes]
R = 4
streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(R - 1)]
frames = [torch.empty((H, W, 4), dtype=torch.int32, device=dev) for _ in range(R)]
out = open(os.path.join(ROOT, "gpurun_out", "rm_ab.jsonl"), "a")
ref_hash = None
for name, kern, level, cubes in CONFIGS:
    fx = capi.FLAG_CUBES if cubes else capi.FLAG_NO_CUBES
    st = [ctx.raymarch_stats(c, W, H, shadow=True, light=LIGHT, cubes=cubes) for c in cams]
    rays = [int(s["primary"]) + int(s["shadow"]) for s in st]
    hashes = []
    for c in cams:
        ctx.raymarch_device(c, W, H, frames[0].data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME, flags_extra=fx)
        torch.cuda.synchronize()
        hashes.append(hashlib.sha1(frames[0].cpu().numpy().tobytes()).hexdigest()[:12])
    if ref_hash is None:
        ref_hash = hashes
    # kernel alone, L2 flushed
    ks = []
    for rep in range(3):
        for c in cams:
            ctx.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.raymarch_device(c, W, H, frames[0].data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME, flags_extra=fx)
            b.record(stream); torch.cuda.synchronize()
            ks.append(a.elapsed_time(b))
    # 4 frames in flight, 64 frames
    def step(k):
        s = streams[k % R]
        ctx.set_stream(s.cuda_stream)
        ctx.raymarch_device(cams[k % 8], W, H, frames[k % R].data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME, flags_extra=fx)
        ctx.set_stream(stream.cuda_stream)
    for k in range(8):
        step(k)
    torch.cuda.synchronize()
    K = 64
    ev0 = torch.cuda.Event(enable_timing=True); ev0.record(stream)
    for s in streams[1:]:
        s.wait_event(ev0)
    for k in range(K):
        step(k)
    evs = []
    for s in streams:
        e = torch.cuda.Event(enable_timing=True); e.record(s); evs.append(e)
    torch.cuda.synchronize()
    ms = max(ev0.elapsed_time(e) for e in evs) / K
    line = {"so": os.path.basename(capi.SO_PATH), "config": name, "N": N, "res": [W, H], "kernel_ms_alone": float(np.mean(ks)), "kernel_ms_alone_min": float(np.min(ks)),
            "ms_per_frame_4_in_flight": ms, "mrays_s": float(np.mean(rays)) / ms / 1e3, "frames_equal_first_config": hashes == ref_hash,
            "steps_per_ray": float(np.mean([int(s["steps"]) / (int(s["primary"]) + int(s["shadow"])) for s in st])),
            "warp_slots_per_ray": float(np.mean([(int(s["warp_slots_primary"]) + int(s["warp_slots_shadow"])) / (int(s["primary"]) + int(s["shadow"])) for s in st])),
            "build_cubes_s": t_cubes}
    print(json.dumps(line), flush=True)
    out.write(json.dumps(line) + "\n"); out.flush()
ctx.close()
