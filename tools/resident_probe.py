"""GPU probe for K6 (resident-set selection + streaming generation), CUDA-event timed.

  1. desired set for one view with the reference's defaults (radius 24 / 6, 120 degrees): kernels only, and the whole
     meso_select_view_chunks call (8-byte count readback + 240 KB list copy); the oracle's time for the same set beside it
     (the reference bakes 256 of these at start-up: ChunkManagerHelper.h:201-234).
  2. the reference's own streaming loop on its own scene (TestGenerator terrain, camera at the origin chunk, view radius
     24): meso_stream_update with max_new = MaxUnsyncedLoadChunkCount = 256 until nothing is missing, then a camera turn.
Prints one JSON object; nothing here is a bench.py value."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from mesoengine_b200 import capi

gran = capi.GRAN_VOXEL if (len(sys.argv) > 1 and sys.argv[1] == "voxel") else capi.GRAN_BLOCK
ctx = capi.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
origin, dims = (-24, -3, -24), (49, 6, 49)
nch = int(np.prod(dims))
ctx.scene_create(origin, dims, 1 << 22)
ctx.stream_begin(capi.SDF_TERRAIN, None, gran)
view = capi.view_config()
out = {"window_chunks": nch, "granularity": "voxel" if gran == capi.GRAN_VOXEL else "block"}


def ev(fn, reps=20):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); fn(); b.record(s); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(min(ts)), float(np.median(ts))


fwd = (0.2, 0.1, 0.97)
cand = ctx.select_view_chunks(fwd, view)
out["candidates"] = int(cand.shape[0])
out["select_kernels_ms(min,med)"] = ev(lambda: ctx.stream_update((0, 0, 0), fwd, 0, view, wait=False))
t0 = time.perf_counter()
for _ in range(20):
    ctx.select_view_chunks(fwd, view)
out["select_call_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
try:
    import orc
    t0 = time.perf_counter(); want = orc.select_view_chunks(fwd); out["oracle_select_ms"] = (time.perf_counter() - t0) * 1e3
    out["select_equal_to_oracle"] = bool(np.array_equal(want["Offset"], cand["Offset"]) and np.array_equal(want["Importance"].view(np.uint32), cand["Importance"].view(np.uint32)))
except Exception as e:  # oracle not built on this box
    out["oracle_select_ms"] = None

# streaming loop
updates = []
total = 0
for k in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); ctx.stream_update((0, 0, 0), fwd, 256, view, wait=False); b.record(s); torch.cuda.synchronize()
    st = ctx.stream_stats()
    updates.append(a.elapsed_time(b)); total += int(st["generated"])
    if int(st["missing"]) == 0:
        break
out["stream_fill"] = {"updates": len(updates), "chunks_generated": total, "in_window": int(st["in_window"]),
                      "ms_per_update_of_256(min,med,max)": (float(min(updates)), float(np.median(updates)), float(max(updates))),
                      "total_ms": float(sum(updates)), "chunks_per_s": total / (sum(updates) * 1e-3)}
# camera turns by 90 degrees: only the newly visible chunks are generated
turn = []
gen = 0
fwd2 = (0.97, 0.1, -0.2)
for k in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); ctx.stream_update((0, 0, 0), fwd2, 256, view, wait=False); b.record(s); torch.cuda.synchronize()
    st = ctx.stream_stats()
    turn.append(a.elapsed_time(b)); gen += int(st["generated"])
    if int(st["missing"]) == 0:
        break
out["stream_turn_90deg"] = {"updates": len(turn), "chunks_generated": gen, "total_ms": float(sum(turn))}
# steady state: nothing to do
out["stream_idle_update_ms(min,med)"] = ev(lambda: ctx.stream_update((0, 0, 0), fwd2, 256, view, wait=False))
occ, full, keys, _ = ctx.volume_download()
out["blocks"] = int(np.unpackbits(occ.view(np.uint8)).sum()); out["partial_bricks"] = int(len(keys))
n_inst = ctx.build_occupancy(1)
out["instances"] = int(n_inst)
print(json.dumps(out))
