"""GPU probe: time K1 (voxelise), K2 (occupancy/instances), K3 (mesh), K5 (carve) on a large grid with CUDA events and
report achieved algorithmic GB/s (DESIGN.md section 4 byte counts) against the measured HBM peak."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesoengine_b200 import capi, scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
peak = 6532.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
origin, dims, params = scenes.sphere_scene(N)
nch = int(np.prod(dims))
ctx = capi.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
ctx.scene_create(origin, dims, 1 << 20)

def timed(fn, reps=5, flush=True):
    ts = []
    for _ in range(reps):
        if flush: ctx.flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); fn(); b.record(s); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

out = {}
t = timed(lambda: ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL), reps=3)
occ, full, keys, _ = ctx.volume_download()
npart = len(keys)
blocks = int(np.unpackbits(occ.view(np.uint8)).sum())
alg = 64.0 * npart + 1024.0 * nch + 16.0 * 1024 * nch / 1024  # payload + occ/full words (+ derived of)
out["K1 voxelize sphere voxel-gran"] = (t, alg)
n_inst = ctx.build_occupancy(1)
t = timed(lambda: ctx.build_occupancy(1), reps=5)
out["K2 occupancy+instances (incl. host count readback)"] = (t, 512.0 * nch + 3 * 512.0 * nch + 16.0 * nch + 12.0 * n_inst)
cap = 1 << 25
quads = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
nq = ctx.mesh_device(quads.data_ptr(), cap)
t = timed(lambda: ctx.mesh_device(quads.data_ptr(), cap, want_count=False), reps=5)
out["K3 mesh (passes A + B + C)   "] = (t, 64.0 * npart + 1024.0 * nch + 16.0 * nq + 4)
c = [int(v) for v in (np.array(dims) * 64)]
c[0] -= int(0.39 * N)  # on the sphere surface facing -x
r = 96
nd = ctx.carve_sphere(c, r)
bricks_aabb = (2 * r // 8 + 1) ** 3
c2 = [c[0], c[1] + 300, c[2] + 100]
t = timed(lambda: ctx.carve_sphere(c2, r), reps=1, flush=True)
out["K5 carve r=%d (first application)" % r] = (t, 128.0 * bricks_aabb + 8.0 * nd)
print("grid %d^3: %d chunks, %d blocks, %d partial bricks, %d instances, %d quads, carve dirty %d" % (N, nch, blocks, npart, n_inst, nq, nd))
for k, ((tmin, tmed), alg) in out.items():
    print("%-55s min %8.3f ms  med %8.3f ms   alg %9.2f MB   %8.1f GB/s  = %.3f of measured HBM peak" % (k, tmin, tmed, alg / 1e6, alg / tmin / 1e6, alg / tmin / 1e6 / peak))
