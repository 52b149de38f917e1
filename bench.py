#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native voxel hot path.

Metric (BASELINE.json): Mrays/s of the primary + hard-shadow voxel raymarch, whole job over N GPUs.
Workload (default, every N): BASELINE.json configs[3] -- 4096^3 sparse-brick scene (V-sphere, voxel-granular),
3840x2160 primary rays + one shadow ray per lit-facing hit, eight orbit cameras cycled per step, screen tiles
(32x8) interleaved over the ranks; every rank's kernel stores its tile records straight into rank 0's frame over
NVLink peer memory (fused gather; `--gather nccl` = all_gather + compose).  Four frames in flight on alternating
streams, like the reference's kNumBufferedFrames = 4.  `--workload cfg1` runs configs[1] (1024^3, 1920x1080).
A "step" is one frame.

    python bench.py --gpus 1 --steps 100 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P bench.py --gpus 8
    python bench.py --impl reference        # the CPU restatement of the reference path (oracle/), all host threads

Prints ONE JSON line (rank 0).  Timing: exactly K steps between two barrier + synchronize points, CUDA events on the
launching streams, max over ranks.  Every step writes its own 132.7 MB frame buffer and re-reads the scene (footprint
above the 126 MB L2); the separate kernel-alone loop behind `roofline` flushes L2 between launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LIGHT = (0.3, 0.5, 0.8)
WORKLOADS = {
    # name: (N voxels, width, height, description)
    "cfg3": (4096, 3840, 2160, "BASELINE.json configs[3]: 4096^3 sparse-brick V-sphere, 3840x2160 primary + shadow rays, 8 orbit cameras"),
    "cfg1": (1024, 1920, 1080, "BASELINE.json configs[1]: 1024^3 sparse-brick V-sphere, 1920x1080 primary + shadow rays, 8 orbit cameras"),
    "cfg0": (256, 1280, 720, "BASELINE.json configs[0]: 256^3 V-sphere, 1280x720 primary + shadow rays, 8 orbit cameras"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload, key=None):
    """dram bytes per launch of the raymarch kernel from the committed ncu capture (profiles/), or None.
    key="<workload>_warp_instructions": smsp__inst_executed.sum of the same capture."""
    p = os.path.join(ROOT, "profiles", "raymarch_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(key or workload)
    return None


def issue_roofline(ctx, world, workload, kernel_ms, clocks):
    """What actually bounds the walk: warp instructions of one launch (ncu smsp__inst_executed.sum of the committed
    capture, 1-GPU launch) against the issue peak = SMs x 4 schedulers x SM clock.  Explanatory, next to the HBM figure."""
    w = load_traffic(workload, workload + "_warp_instructions")
    if w is None or world != 1:
        return None
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = ctx.sm_count() * 4 * mhz * 1e6
    out = {"warp_instructions_per_launch": w, "peak_warp_instructions_per_s": peak, "achieved_per_s": w / (kernel_ms * 1e-3),
           "frac": w / (kernel_ms * 1e-3) / peak, "source": "profiles/raymarch_traffic.json (ncu capture of the shipped kernel)"}
    ti, lanes, rays = (load_traffic(workload, workload + k) for k in ("_thread_instructions", "_active_threads_per_instruction", "_rays_in_capture"))
    if ti and rays:
        out["thread_instructions_per_ray"] = ti / rays          # round 1: ~3 500
        out["active_threads_per_instruction"] = lanes
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_cameras(scene, width, height):
    from mesoengine_b200 import camera, scenes
    origin, dims, _ = scene
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    return [camera.camera_uniform(e, ctr, width, height) for e in eyes]


def sample_rows(height, n_bands=8, rows=8):
    """The bounded CPU sample: n_bands bands of `rows` scanlines spread evenly over the frame, as one list of scanlines."""
    out = []
    for b in range(n_bands):
        y0 = int((b + 0.5) * height / n_bands) - rows // 2
        y0 = max(0, min(height - rows, y0))
        out.extend(range(y0, y0 + rows))
    return out


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement (oracle/), all host threads, bounded sample per step
# ------------------------------------------------------------------------------------------------------------------

def cpu_build_volume(orc, scene):
    origin, dims, params = scene
    return orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL, fast=True)


def cpu_step(orc, vol, cam, width, height, rows, nthreads, out=None):
    """One step of the CPU arm: the sampled scanlines of one frame in ONE parallel region over (row, 128-pixel span) items
    (oracle/orc_raymarch.c:orc_raymarch_rows), so every one of `nthreads` threads has work for the whole call."""
    rs = orc.ray_setup(cam, vol.origin, width, height, LIGHT)
    rec, st = vol.raymarch_rows(rs, width, height, rows, shadow=True, mode=orc.DDA_HIER, nthreads=nthreads, out=out)
    return int(st["primary"]) + int(st["shadow"]), rec


def cpu_thread_scaling(orc, vol, cam, width, height, rows, nthreads):
    """Mrays/s of the same sample on 1 thread and on all of them: what `cores` in cpu_baseline actually buys."""
    out = {}
    for nt in sorted({1, max(1, nthreads // 2), nthreads}):
        r = rows if nt > 1 else rows[::8]            # one thread gets an eighth of the sample
        t0 = time.perf_counter()
        rays, _ = cpu_step(orc, vol, cam, width, height, r, nt)
        out[str(nt)] = rays / (time.perf_counter() - t0) / 1e6
    out["speedup_all_over_1"] = out[str(nthreads)] / out["1"]
    return out


def reference_code_timings(nthreads):
    """The REFERENCE'S OWN CODE, where it could be compiled (oracle/_ref/libmeso_ref.so: the reference's hot-path headers
    and shader text, built in the authoring container from /root/reference, see oracle/ref_driver.cpp): its CPU
    generator-worker body (GenerateSphere + CalculateOccupancyErodeMipmaps + hidden-block test, ChunkManager.h:160-170)
    over the 448 chunks of the reference sphere on all host threads, and its instanced draw (the voxel VS/FS text through a
    scalar software pipeline) over the resulting 201 936 instances on one thread.  Reported beside the restated path;
    neither is the timed headline workload (the reference has no voxel-in-brick level and no shadow rays)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import numpy as np
        import orc
        import refprobe
        from concurrent.futures import ThreadPoolExecutor
        if not os.path.exists(refprobe.REF_SO):
            return {"unavailable": "oracle/_ref/libmeso_ref.so not present (it is built where /root/reference exists)"}
        ref = refprobe.RefBackend(refprobe.REF_SO)
        chunks = refprobe.SPHERE_CHUNKS

        def gen(loc):
            xyz, _, cull = ref.generate_chunk(0, loc)      # ctypes releases the GIL; the call uses locals only
            return len(xyz), int(len(xyz) - cull.sum())
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            rows = list(ex.map(gen, chunks))
        dt_gen = time.perf_counter() - t0
        blocks, inst = sum(r[0] for r in rows), sum(r[1] for r in rows)
        out = {"kind": "reference", "generate": {"chunks": len(chunks), "blocks": blocks, "instances": inst, "cores": nthreads,
                                                 "s": dt_gen, "chunks_per_s": len(chunks) / dt_gen,
                                                 "what": "GenerateSphere + CalculateOccupancyErodeMipmaps(16, 4) + bShouldVoxelOccupancyCull(.., 1) per chunk"}}
        # the draw: block-granular reference sphere (the reference's own scene), 1280x720 (BASELINE.json configs[0])
        origin, dims = (2, -4, -4), (8, 8, 8)
        vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, orc.REF_SPHERE, granularity=orc.GRAN_BLOCK)
        table, _, instances = vol.build_occupancy(stamp=1)
        w, h = 1280, 720
        cam = orc.camera_uniform((5.0, 2.0, 2.0), (100.0, 0.0, 0.0), width=w, height=h)   # reference start pose, looking at the sphere
        t0 = time.perf_counter()
        depth, instance, color, normal, behind = refprobe.ref_draw(ref.lib, cam, orc.default_scene_config(), table, instances,
                                                                  ref.triplanar_indices(), w, h)
        dt_draw = time.perf_counter() - t0
        out["instanced_draw"] = {"instances": int(len(instances)), "resolution": [w, h], "cores": 1, "s": dt_draw,
                                 "mpixels_per_s": w * h / dt_draw / 1e6, "covered_pixels": int((instance >= 0).sum()),
                                 "vertices_behind_camera": behind,
                                 "what": "cmdDrawIndexed(8, instances) through the reference's vertex/fragment shader text, scalar software pipeline"}
        return out
    except Exception as e:  # the reported baseline must never take the arm down
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # oracle only: this arm must not load the product library (tests/scenes.py loads the scene formulas by file path)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    import scenes
    n, width, height, desc = WORKLOADS[args.workload]
    scene = scenes.sphere_scene(n)
    origin, dims, _ = scene
    nthreads = orc.hw_threads()
    vol = cpu_build_volume(orc, scene)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cams = [orc.camera_uniform(e, ctr, width=width, height=height) for e in eyes]
    n_bands = max(8, min(64, 2 * nthreads))            # >= 16 items per thread whatever the box
    rows = sample_rows(height, n_bands, 8)
    buf = np.zeros((height, width), dtype=orc.HitRecord)
    for k in range(max(1, min(args.warmup, 2))):
        cpu_step(orc, vol, cams[k % 8], width, height, rows, nthreads, out=buf)
    t0 = time.perf_counter()
    rays = 0
    for k in range(args.steps):
        r, _ = cpu_step(orc, vol, cams[k % 8], width, height, rows, nthreads, out=buf)
        rays += r
    dt = time.perf_counter() - t0
    value = rays / dt / 1e6
    scaling = cpu_thread_scaling(orc, vol, cams[0], width, height, rows, nthreads)
    sample = "%d bands x 8 rows x %d px per step (%.1f%% of the frame) in one parallel region, cameras cycled" % (n_bands, width, 100.0 * len(rows) / height)
    line = {
        "impl": "reference", "metric": "Mrays/s voxel raymarch (primary + shadow)", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "scene": "V-sphere %d^3 voxel-granular" % n, "resolution": [width, height],
                   "note": "the reference application cannot be built here (SURVEY.md 8c) and has no voxel-in-brick level or shadow rays; the timed path is the CPU restatement (oracle/), which oracle/_ref pins against the reference's own code; 'reference_code' times that code itself"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": nthreads, "kind": "port", "sample": sample, "thread_scaling_mrays_s": scaling},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_code": reference_code_timings(nthreads),
    }
    emit_line(line)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from mesoengine_b200 import capi, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the voxel path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")   # host-side barriers around the single-process group leg (no GPU spin)
    dev = torch.device("cuda", local_rank)

    n, width, height, desc = WORKLOADS[args.workload]
    scene = scenes.sphere_scene(n)
    origin, dims, params = scene
    ctx = capi.Context(local_rank)
    # one side stream carries everything: our kernels (through the C ABI), NCCL, the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # frames in flight: the reference keeps kNumBufferedFrames = 4 per-frame buffers (Samples/SimpleVoxel.cpp:15); consecutive
    # frames go to alternating streams / frame buffers so the long tail of one frame (a few grazing rays) overlaps the next
    R = max(1, min(8, args.frames_in_flight))
    streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(R - 1)]
    ctx.scene_create(origin, dims, max_bricks=(1 << 20) if n >= 4096 else (1 << 18))
    t_build = time.perf_counter()
    ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)  # replicated on every rank
    ctx.sync()
    t_build = time.perf_counter() - t_build
    cams = make_cameras(scene, width, height)

    # per-camera ray counts and algorithmic bytes (full frame, instrumented launch outside the timed region)
    ctx.set_partition(0, 1)
    stats = [ctx.raymarch_stats(c, width, height, shadow=True, light=LIGHT) for c in cams]
    rays_cam = [int(s["primary"]) + int(s["shadow"]) for s in stats]
    u_cam = [int(s["u_bytes"]) for s in stats]
    ctx.set_partition(rank, world)

    px = width * height
    tpr = capi.tiles_per_rank(width, height, world)
    gather = "single GPU"
    frame_ptrs = [None] * R        # device pointers every rank stores its tile records to (fused gather), one per frame in flight
    frame_owner_ptrs = [None] * R
    if world > 1 and args.gather == "p2p":
        # Fused gather: rank 0 owns the frame (cudaMalloc through the C ABI, exported over CUDA IPC); every rank's
        # raymarch kernel stores its tile records straight into it over NVLink.  No gather pass, no compose pass.
        # every rank issues every broadcast whatever fails where (a rank that left the loop early would leave the others
        # waiting in a collective); failures are recorded and agreed on by one MIN all-reduce afterwards
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        for i in range(R):
            handle = torch.zeros(capi.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
            if rank == 0:
                try:
                    frame_owner_ptrs[i] = ctx.device_alloc(px * 16)
                    handle.copy_(torch.from_numpy(ctx.ipc_export(frame_owner_ptrs[i])))
                except Exception as e:  # e.g. IPC not permitted in this container
                    sys.stderr.write("bench: p2p gather unavailable on rank 0 (%s); falling back to NCCL all_gather\n" % e)
                    ok.zero_()
            dist.broadcast(handle, src=0)
            if rank == 0:
                frame_ptrs[i] = frame_owner_ptrs[i]
            elif int(ok.item()) == 1:
                try:
                    frame_ptrs[i] = ctx.ipc_open(handle.cpu().numpy())
                except Exception as e:
                    sys.stderr.write("bench: p2p gather unavailable on rank %d (%s); falling back to NCCL all_gather\n" % (rank, e))
                    ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            gather = "p2p"
        else:   # release what was opened / allocated before falling back
            for i in range(R):
                if rank != 0 and frame_ptrs[i] is not None:
                    try:
                        ctx.ipc_close(frame_ptrs[i])
                    except Exception:
                        pass
            dist.barrier()
            if rank == 0:
                for i in range(R):
                    if frame_owner_ptrs[i] is not None:
                        ctx.device_free(frame_owner_ptrs[i])
            frame_ptrs = [None] * R
            frame_owner_ptrs = [None] * R
    # rendezvous of the frame loops: arrival words (signal / wait kernels over peer memory, no collective) unless --rendezvous nccl
    arrival = None
    if world > 1:
        arrival = ArrivalWords(ctx, capi, torch, dist, dev, rank, world)
        if not arrival.ok or args.rendezvous == "nccl":
            arrival.close(dist)
            arrival = None
    if world == 1 or gather != "p2p":
        frames = [torch.empty((height, width, 4), dtype=torch.int32, device=dev) for _ in range(R)]
    if world > 1 and gather != "p2p":
        gather = "nccl"
        tiles = [torch.empty((tpr, 256, 4), dtype=torch.int32, device=dev) for _ in range(R)]
        gathered = [torch.empty((world, tpr, 256, 4), dtype=torch.int32, device=dev) for _ in range(R)]
    flags = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(R)]

    def step(k, slot=0):
        """One frame, enqueued on the stream of ring slot `slot`."""
        cam = cams[k % 8]
        s = streams[slot]
        ctx.set_stream(s.cuda_stream)
        with torch.cuda.stream(s):
            if world == 1:
                ctx.raymarch_device(cam, width, height, frames[slot].data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
            elif gather == "p2p" and arrival is not None:
                # word `slot` of rank 0: every rank adds 1 behind its kernel (stream order); rank 0's stream waits for all of them
                ctx.raymarch_device(cam, width, height, frame_ptrs[slot], shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
                ctx.signal_device([arrival.words[0] + 4 * slot])
                arrival.gen[slot] += world
                if rank == 0:
                    ctx.wait_device(arrival.words[0] + 4 * slot, arrival.gen[slot])
            elif gather == "p2p":
                ctx.raymarch_device(cam, width, height, frame_ptrs[slot], shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
                dist.all_reduce(flags[slot])   # stream-ordered 4-byte rendezvous: when it completes on rank 0 every tile has landed
            else:
                ctx.raymarch_device(cam, width, height, tiles[slot].data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_TILES)
                dist.all_gather_into_tensor(gathered[slot], tiles[slot])
                ctx.compose_tiles_device(gathered[slot].data_ptr(), world, width, height, frames[slot].data_ptr())
        ctx.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for k in range(args.warmup):
        step(k, k % R)
    barrier()
    if sampler:
        sampler.start()

    # ---- timed region: exactly K steps between two barrier + synchronize points, CUDA events on the launching streams.
    # R frames in flight; every step writes its own 132.7 MB frame buffer (R of them cycled) and re-reads ~12 MB of scene,
    # so the per-step footprint exceeds the 126 MB L2 on its own -- no artificial flush inside the region. ----
    # The region is run `--regions` times (default 5), each one exactly K steps between two barrier + synchronize points; the
    # line reports the MEDIAN region (at 8 GPUs a 20-step region is 4 ms: one region alone is noise-sensitive), all of them
    # are listed in config.regions_ms.
    region_ms = []
    launches = 0
    for reg in range(max(1, args.regions)):
        launches0 = ctx.launch_count()
        ctx.flush_l2()
        barrier()
        ev_start = torch.cuda.Event(enable_timing=True)
        ev_start.record(stream)
        for s in streams[1:]:
            s.wait_event(ev_start)
        for k in range(args.steps):
            step(k, k % R)
        ev_end = []
        for s in streams:
            e = torch.cuda.Event(enable_timing=True)
            e.record(s)
            ev_end.append(e)
        barrier()
        launches = ctx.launch_count() - launches0 - 1  # minus the one flush before the region
        ms_r = max(ev_start.elapsed_time(e) for e in ev_end)
        total_ms = torch.tensor([ms_r], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        region_ms.append(float(total_ms.item()))
    ms = float(np.median(region_ms))
    rays_total = sum(rays_cam[k % 8] for k in range(args.steps))
    value = rays_total / (ms * 1e-3) / 1e6

    # ---- 1 == N check (outside the timed region): the gathered frame equals the frame one GPU renders on its own ----
    gather_verified = None
    if world > 1:
        step(0, 0)
        barrier()
        if rank == 0:
            got = np.empty((height, width), dtype=capi.HitRecord)
            if gather == "p2p":
                ctx.download(got, frame_owner_ptrs[0])
            else:
                got = frames[0].cpu().numpy().view(capi.HitRecord).reshape(height, width)
            ctx.set_partition(0, 1)
            ref = ctx.raymarch(cams[0], width, height, shadow=True, light=LIGHT)
            ctx.set_partition(rank, world)
            gather_verified = bool(got.tobytes() == ref.tobytes())
        barrier()

    # ---- meshing on N GPUs (BASELINE.json: meshed voxels/s at 1/2/4/8): the resident scene, fused quad gather ----
    mesh_multi = None
    if world > 1 and not args.no_mesh:
        mesh_multi = bench_mesh_multi(ctx, capi, torch, dist, dev, rank, world, stream, n)

    # ---- dominant kernel alone (this rank's tiles), for the roofline ----
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world == 1:
        kptr, klayout = frames[0].data_ptr(), capi.LAYOUT_FRAME
    elif gather == "p2p":
        kptr, klayout = frame_ptrs[0], capi.LAYOUT_FRAME
    else:
        kptr, klayout = tiles[0].data_ptr(), capi.LAYOUT_TILES
    for k in range(args.steps):
        ctx.flush_l2()
        kev[k][0].record(stream)
        ctx.raymarch_device(cams[k % 8], width, height, kptr, shadow=True, light=LIGHT, layout=klayout)
        kev[k][1].record(stream)
    barrier()
    kms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    hbm_peak, peak_src = load_peaks()
    # algorithmic bytes per launch: 16 B per pixel of this rank + scene bytes touched (U, counted once)
    alg_bytes = float(np.mean([16.0 * px / world + u_cam[k % 8] for k in range(args.steps)]))
    achieved = alg_bytes / (kms * 1e-3) / 1e9

    # ---- end to end through the host-buffer API: camera in host memory -> records in pinned host memory ----
    # N = 1: the frame-ring API (meso_raymarch_async / meso_frame_wait, 4 slots = the reference's kNumBufferedFrames): the
    # 132.7 MB device-to-host copy of frame k overlaps the traversal of frame k+1; every frame's records are consumed
    # (first and last record read on the host) before its slot is reused.  N > 1: frame gathered on rank 0, then copied.
    RING = 4
    hosts = [torch.empty((height, width, 4), dtype=torch.int32).pin_memory() for _ in range(RING if world == 1 else (R if rank == 0 else 1))]
    hosts_np = [h.numpy().view(capi.HitRecord).reshape(height, width) for h in hosts]
    host, host_np = hosts[0], hosts_np[0]
    e2e_steps = max(8, min(args.steps, 40))
    consumed = 0

    # N > 1: fused gather straight into HOST memory.  One shared-memory segment holds the R frames in flight; every rank
    # maps it, registers it with its own GPU (meso_host_register) and its raymarch kernel stores its tiles there over its
    # own PCIe link while tracing: N links instead of rank 0's one, no device-side frame, no copy.  Falls back to
    # "gather on rank 0, then copy" if the segment cannot be created or registered.
    host_fused = False
    shm = None
    if world > 1 and not args.no_host_fused:
        shm_path = "/dev/shm/meso_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getuid())
        # every rank runs the same two collectives whatever fails where (no rank may be left waiting in one of them)
        created = torch.zeros(1, dtype=torch.int32, device=dev)
        if rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < R * px * 16 + (64 << 20):   # a short tmpfs would SIGBUS on first touch
                    raise OSError("/dev/shm has %d MB free, need %d MB" % (st.f_bavail * st.f_frsize >> 20, R * px * 16 >> 20))
                with open(shm_path, "wb") as f:
                    f.truncate(R * px * 16)
                created.fill_(1)
            except Exception as e:
                sys.stderr.write("bench: cannot create the shared host frame (%s)\n" % e)
        dist.broadcast(created, src=0)
        ok = torch.zeros(1, dtype=torch.int32, device=dev)
        if int(created.item()) == 1:
            try:
                shm = np.memmap(shm_path, dtype=np.uint8, mode="r+", shape=(R * px * 16,))
                shm_dptr = ctx.host_register(shm)
                ok.fill_(1)
            except Exception as e:
                sys.stderr.write("bench: host-fused gather unavailable on rank %d (%s)\n" % (rank, e))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        host_fused = int(ok.item()) == 1
        if not host_fused:
            if shm is not None:
                try:
                    ctx.host_unregister(shm)
                except Exception:
                    pass
                shm = None
            if rank == 0 and os.path.exists(shm_path):
                os.unlink(shm_path)
        if host_fused:
            shm_views =[shm[i * px * 16:(i + 1) * px * 16].view(capi.HitRecord).reshape(height, width) for i in range(R)]

    # N > 1, default: SLAB gather.  Rank r owns rows [r * rows, (r + 1) * rows) of every frame in flight in its own device
    # memory, exported to all ranks over CUDA IPC; every rank's kernel stores each record straight into the owner's slab
    # over NVLink (MESO_LAYOUT_SLABS: the all-to-all is fused into the store), and after the rendezvous every rank copies its
    # contiguous slab into the shared host frame with ONE large DMA over its own PCIe link -- instead of 512-byte stores
    # from N GPUs interleaved inside every 4 KB host page.  Falls back to the host-fused stores above.
    slabs_ok = False
    slab_rows = ((((height + capi.TILE_H - 1) // capi.TILE_H) + world - 1) // world) * capi.TILE_H
    slab_mine, slab_ptrs = [None] * R, [[None] * world for _ in range(R)]
    if host_fused and not args.no_slabs:
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        handles = torch.zeros((R, capi.IPC_HANDLE_BYTES), dtype=torch.uint8, device=dev)
        try:
            for i in range(R):
                slab_mine[i] = ctx.device_alloc(slab_rows * width * 16)
                handles[i].copy_(torch.from_numpy(ctx.ipc_export(slab_mine[i])))
        except Exception as e:
            sys.stderr.write("bench: slab gather unavailable on rank %d (%s)\n" % (rank, e))
            ok.zero_()
        allh = torch.zeros((world, R, capi.IPC_HANDLE_BYTES), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, handles)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            allh_np = allh.cpu().numpy()
            try:
                for i in range(R):
                    for r in range(world):
                        slab_ptrs[i][r] = slab_mine[i] if r == rank else ctx.ipc_open(allh_np[r, i])
            except Exception as e:
                sys.stderr.write("bench: slab gather unavailable on rank %d (%s)\n" % (rank, e))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        slabs_ok = int(ok.item()) == 1
    my_row0 = rank * slab_rows
    my_rows = max(0, min(slab_rows, height - my_row0))
    flags2 = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(R)]

    # The timed loop again with the frame left DISTRIBUTED (device only, no copy to the host): every rank ends up with its
    # rows of the frame instead of rank 0 receiving (N-1)/N of 132.7 MB per frame.  Shows how much of `value` at large N is
    # rank 0's NVLink ingress rather than tracing (reported as config.distributed_slabs, not the headline).
    distributed = None
    if slabs_ok and arrival is not None:
        def slab_step(k):
            slot = k % R
            s = streams[slot]
            ctx.set_stream(s.cuda_stream)
            ctx.raymarch_device_slabs(cams[k % 8], width, height, slab_ptrs[slot], slab_rows, shadow=True, light=LIGHT)
            ctx.signal_device(arrival.all_ranks(28 + slot % 4))
            arrival.gen[28 + slot % 4] += world
            ctx.wait_device(arrival.mine + 4 * (28 + slot % 4), arrival.gen[28 + slot % 4])
            ctx.set_stream(stream.cuda_stream)
        for k in range(args.warmup):
            slab_step(k)
        regs = []
        for _ in range(3):
            barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for s in streams[1:]:
                s.wait_event(e0)
            for k in range(args.steps):
                slab_step(k)
            ends = []
            for s in streams:
                e = torch.cuda.Event(enable_timing=True)
                e.record(s)
                ends.append(e)
            barrier()
            t = torch.tensor([max(e0.elapsed_time(e) for e in ends)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            regs.append(float(t.item()))
        dms = float(np.median(regs)) / args.steps
        distributed = {"ms_per_step": dms, "mrays_s": rays_total / args.steps / dms / 1e3,
                       "rank0_ingress_gbs_in_the_gathered_loop": 16.0 * px * (world - 1) / world / (ms / args.steps * 1e-3) / 1e9,
                       "note": "same kernels and ring, MESO_LAYOUT_SLABS: every rank keeps rows [r*H/N, (r+1)*H/N) of the frame (all-to-all over NVLink, arrival words), nothing is gathered on one GPU"}

    def slab_frame(k, slot, rgba8=False):
        """One frame through the slab gather on ring slot `slot`'s stream: kernel -> rendezvous (every rank's slab is
        complete) -> this rank's slab to the shared host frame -> rendezvous (every slab has landed)."""
        bpp = 4 if rgba8 else 16
        s = streams[slot]
        ctx.set_stream(s.cuda_stream)
        with torch.cuda.stream(s):
            off = slot * px * 16 + my_row0 * width * bpp
            if arrival is not None:
                # words 8 + slot: "kernels done" (all-to-all: every rank's slab receives rows from every rank);
                # words 16 + slot: "slab has reached the host" (all-to-all: the next frame on this slot overwrites peers' slabs)
                ia, ib = 8 + slot, 16 + slot
                ctx.raymarch_device_slabs(cams[k % 8], width, height, slab_ptrs[slot], slab_rows, shadow=True, light=LIGHT,
                                          flags_extra=capi.FLAG_RGBA8 if rgba8 else 0)
                ctx.signal_device(arrival.all_ranks(ia))
                arrival.gen[ia] += world
                ctx.wait_device(arrival.mine + 4 * ia, arrival.gen[ia])
                if my_rows > 0:
                    ctx.download_async(shm[off:off + my_rows * width * bpp], slab_mine[slot])
                ctx.signal_device(arrival.all_ranks(ib))
                arrival.gen[ib] += world
                ctx.wait_device(arrival.mine + 4 * ib, arrival.gen[ib])
            else:
                ctx.raymarch_device_slabs(cams[k % 8], width, height, slab_ptrs[slot], slab_rows, shadow=True, light=LIGHT,
                                          flags_extra=capi.FLAG_RGBA8 if rgba8 else 0)
                dist.all_reduce(flags[slot])
                if my_rows > 0:
                    ctx.download_async(shm[off:off + my_rows * width * bpp], slab_mine[slot])
                dist.all_reduce(flags2[slot])
            e = torch.cuda.Event()
            e.record(s)
        ctx.set_stream(stream.cuda_stream)
        return e

    def e2e_run(nsteps, rgba8=False):
        nonlocal consumed
        if world > 1 and host_fused:
            pend = [None] * R
            for k in range(nsteps):
                slot = k % R
                if pend[slot] is not None:
                    pend[slot].synchronize()
                    if rank == 0:
                        if rgba8:
                            consumed += int(shm[slot * px * 16]) + int(shm[slot * px * 16 + 4 * px - 1])
                        else:
                            consumed += int(shm_views[slot]["w1"][0, 0]) + int(shm_views[slot]["w1"][-1, -1])
                if slabs_ok:
                    pend[slot] = slab_frame(k, slot, rgba8)
                    continue
                s = streams[slot]
                ctx.set_stream(s.cuda_stream)
                with torch.cuda.stream(s):
                    ctx.raymarch_device(cams[k % 8], width, height, shm_dptr + slot * px * 16, shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME,
                                        flags_extra=capi.FLAG_RGBA8 if rgba8 else 0)
                    dist.all_reduce(flags[slot])   # stream-ordered rendezvous: behind it every rank's tiles are in host memory
                    e = torch.cuda.Event()
                    e.record(s)
                    pend[slot] = e
                ctx.set_stream(stream.cuda_stream)
            for slot in range(R):
                if pend[slot] is not None:
                    pend[slot].synchronize()
        elif world == 1:
            for k in range(nsteps):
                slot = k % RING
                if k >= RING:
                    ctx.frame_wait(slot)
                    consumed += int(hosts_np[slot]["w1"][0, 0]) + int(hosts_np[slot]["w1"][-1, -1])
                ctx.raymarch_async(cams[k % 8], width, height, hosts_np[slot], slot, shadow=True, light=LIGHT)
            for slot in range(RING):
                ctx.frame_wait(slot)
                consumed += int(hosts_np[slot]["w1"][0, 0]) + int(hosts_np[slot]["w1"][-1, -1])
        else:
            # ring over the R frames in flight: rank 0 enqueues the copy of the gathered frame right behind the frame's
            # rendezvous on that frame's stream, so it overlaps the next frame; a slot is consumed before it is reused
            pend = [None] * R
            for k in range(nsteps):
                slot = k % R
                if pend[slot] is not None:
                    pend[slot].synchronize()
                    if rank == 0:
                        consumed += int(hosts_np[slot]["w1"][0, 0]) + int(hosts_np[slot]["w1"][-1, -1])
                step(k, slot)
                with torch.cuda.stream(streams[slot]):
                    if rank == 0:
                        if gather == "p2p":
                            ctx.set_stream(streams[slot].cuda_stream)
                            ctx.download_async(hosts_np[slot], frame_owner_ptrs[slot])
                            ctx.set_stream(stream.cuda_stream)
                        else:
                            hosts[slot].copy_(frames[slot], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(streams[slot])
                    pend[slot] = e
            for slot in range(R):
                if pend[slot] is not None:
                    pend[slot].synchronize()

    e2e_run(RING)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_rays = sum(rays_cam[k % 8] for k in range(e2e_steps))
    e2e_value = e2e_rays / float(e2e_dt.item()) / 1e6
    host_fused_verified = None
    edit_multi = None
    e2e_rgba8 = None
    if host_fused:
        # the same ring with MESO_FLAG_RGBA8: 4 B per pixel into the shared host frame
        e2e_run(R, rgba8=True)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps, rgba8=True)
        barrier()
        dt8 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt8, op=dist.ReduceOp.MAX)
        e2e_rgba8 = {"value": e2e_rays / float(dt8.item()) / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": 4 * px,
                     "note": "MESO_FLAG_RGBA8 through the host-fused gather: only the colour word of every record travels"}
        # the frame the ranks assembled in host memory equals the frame one GPU renders on its own
        if slabs_ok:
            slab_frame(0, 0).synchronize()
        else:
            ctx.raymarch_device(cams[0], width, height, shm_dptr, shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
        barrier()
        if rank == 0:
            ctx.set_partition(0, 1)
            ref = ctx.raymarch(cams[0], width, height, shadow=True, light=LIGHT)
            ctx.set_partition(rank, world)
            host_fused_verified = bool(shm_views[0].tobytes() == ref.tobytes())
        barrier()
        # BASELINE.json configs[4] on N GPUs: carve replicated on every rank, render split by tiles into the shared frame
        edit_multi = None
        if not args.no_mesh:
            slab_info = dict(ptrs=slab_ptrs, mine=slab_mine, rows=slab_rows, row0=my_row0, my_rows=my_rows, shm=shm, px=px, R=R,
                             flags2=flags2, streams=streams, arrival=arrival, world=world) if slabs_ok else None
            edit_multi = bench_edit_loop_multi(ctx, capi, cams, width, height, shm_views[0], shm_dptr, flags[0], stream, rank, dist, torch, slabs=slab_info)
            if slabs_ok:
                # the same loop delivering the image as RGBA8 (the reference's TEXOffscreenColor format): the records stay on the
                # devices for the pick, their colour words are packed and copied -- 33 MB instead of 133 MB per frame into the host
                e8 = bench_edit_loop_multi(ctx, capi, cams, width, height, shm_views[0], shm_dptr, flags[0], stream, rank, dist, torch, slabs=slab_info, rgba8=True)
                if edit_multi is not None and e8 is not None:
                    edit_multi["rgba8"] = {k: e8[k] for k in ("fps", "ms_per_frame", "ms_carve", "ms_remesh_dirty", "ms_render_to_host", "host_bytes_per_frame")}
        barrier()
        if slabs_ok:
            for i in range(R):
                for r in range(world):
                    if r != rank:
                        ctx.ipc_close(slab_ptrs[i][r])
            dist.barrier()
            for i in range(R):
                ctx.device_free(slab_mine[i])
        ctx.host_unregister(shm)
        del shm_views
        shm = None
        dist.barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass
    # same loop with MESO_FLAG_RGBA8: the reference's own output format (RGBA_UN8 colour target), 4 B instead of 16 B per pixel
    if world == 1:
        imgs = [torch.empty((height, width), dtype=torch.int32).pin_memory() for _ in range(RING)]
        imgs_np = [t.numpy().view(np.uint32) for t in imgs]

        def rgba_run(nsteps):
            nonlocal consumed
            for k in range(nsteps):
                slot = k % RING
                if k >= RING:
                    ctx.frame_wait(slot)
                    consumed += int(imgs_np[slot][0, 0]) + int(imgs_np[slot][-1, -1])
                ctx.raymarch_async(cams[k % 8], width, height, imgs_np[slot], slot, shadow=True, light=LIGHT, rgba8=True)
            for slot in range(RING):
                ctx.frame_wait(slot)
                consumed += int(imgs_np[slot][0, 0]) + int(imgs_np[slot][-1, -1])

        rgba_run(RING)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rgba_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_rgba8 = {"value": e2e_rays / (time.perf_counter() - t0) / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": 4 * px,
                     "note": "MESO_FLAG_RGBA8: only the colour word of every record is written and copied (the reference's RGBA_UN8 offscreen target)"}
        del imgs
    clocks = sampler.stop() if sampler else None

    line = None
    if rank == 0:
        line = {
            "metric": "Mrays/s voxel raymarch (primary + shadow)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "scene": "V-sphere %d^3 voxel-granular, %d chunks, replicated per GPU" % (n, int(np.prod(dims))),
                       "resolution": [width, height], "rays_per_frame_mean": rays_total / args.steps,
                       "partition": ("single GPU, row-major frame" if world == 1 else
                                     ("32x8 screen tiles, tile %% %d == rank; " % world) +
                                     ("fused gather: every rank's kernel stores its records into rank 0's frame over NVLink peer memory; rendezvous = " + ("an arrival word on rank 0 that every rank increments behind its kernel (one-thread signal kernel, system-scope atomic over NVLink), waited for on rank 0's stream: no collective in the frame loop" if arrival is not None else "4-byte NCCL all-reduce")
                                      if gather == "p2p" else "NCCL all_gather of packed tile records + compose kernel")),
                       "cache": "no flush inside the timed region: every step writes its own 132.7 MB frame (%d frame buffers cycled) and re-reads the scene, a per-step footprint above the 126 MB L2; the kernel-alone roofline loop flushes L2 (256 MiB write) between launches" % R,
                       "frames_in_flight": R, "distributed_slabs": distributed, "regions_ms": region_ms, "regions_note": "each region = exactly `steps` steps between barrier + synchronize points, max over ranks; value / ms_per_step are the median region",
                       "gather": gather, "gather_verified_equal_to_1gpu_frame": gather_verified,
                       "scene_build_s": t_build,
                       "walk": "mirrored-space stateless DDA over per-octant forward cubes (32^3 cells, bricks), built with the volume"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 160, "d2h_bytes_per_step": 16 * px,
                    "steps": e2e_steps, "rgba8": e2e_rgba8,
                    "note": ("meso_raymarch_async()/meso_frame_wait() frame ring of 4: FGPUUniformCamera from host memory (kernel parameters), records copied to pinned host memory, copy of frame k overlapping frame k+1"
                             if world == 1 else
                             (("slab gather: every rank's kernel stores each record into the device slab of the rank that owns its rows (peer memory over NVLink, MESO_LAYOUT_SLABS), 4-byte NCCL all-reduce as the rendezvous, then every rank copies its contiguous slab into the shared host frame with one DMA over its own PCIe link; frames consumed on rank 0"
                               if slabs_ok else
                               "fused gather into host memory: one shared-memory segment registered by every rank (meso_host_register); each rank's kernel stores its tile records there over its own PCIe link, 4-byte NCCL all-reduce as the rendezvous, frames consumed on rank 0")
                              if host_fused else
                              "frame gathered on rank 0 (fused p2p stores or NCCL), then copied to pinned host memory on that frame's stream, overlapping the next frame in flight")),
                    "host_fused": host_fused, "slab_gather": slabs_ok, "host_fused_verified_equal_to_1gpu_frame": host_fused_verified},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "raymarch10_kernel<false, 3, false>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": load_traffic(args.workload), "peak_source": peak_src,
                         "kernel_ms": kms, "kernel_ms_is": "isolated: one launch per CUDA-event pair, L2 flushed (256 MiB write) before each; NOT the per-step cost",
                         "kernel_ms_in_loop": ms / args.steps, "kernel_ms_in_loop_is": "timed region / steps with %d frames in flight on alternating streams (the tail of frame k overlaps frame k+1)" % R,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture committed under profiles/ (profiles/raymarch_traffic.json names the file); not measured in this run",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "issue": issue_roofline(ctx, world, args.workload, kms, clocks),
                         "note": "divergence/instruction-issue-bound traversal; working set is L2-resident (SURVEY.md 8d); 'issue' relates the ncu-counted warp instructions of one launch to the SMs' issue peak"},
        }

    # ---- BASELINE.json configs[4]: interactive edit loop (1 GPU leg; carve is replicated compute on N GPUs) ----
    if world == 1 and not args.no_mesh and rank == 0:
        line["edit_loop"] = bench_edit_loop(ctx, capi, cams, width, height)
    if world > 1 and rank == 0 and edit_multi is not None:
        line["edit_loop"] = edit_multi
    if world > 1 and rank == 0 and mesh_multi is not None:
        line["mesh"] = {"sphere_%d_voxels" % n: mesh_multi}

    # ---- secondary metric of BASELINE.json: meshed voxels/s (configs[2]-style, 1 GPU leg only) ----
    if world == 1 and not args.no_mesh and rank == 0:
        line["mesh"] = bench_mesh(ctx, capi, scenes, torch, stream, args, dev, n_work=n)

    # ---- the reference's own streaming loop on its own scene (1 GPU leg) ----
    if world == 1 and not args.no_mesh and rank == 0:
        line["stream"] = bench_stream(ctx, capi, torch, stream)

    # ---- BASELINE.json configs[0] and configs[1] at their stated sizes (1 GPU leg): Mrays/s + sampled parity vs the oracle ----
    if world == 1 and not args.no_mesh and rank == 0 and args.workload == "cfg3":
        line["configs"] = {name: bench_small_config(ctx, capi, scenes, torch, stream, name, args) for name in ("cfg0", "cfg1")}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample, outputs byte-compared ----
    if world == 1 and not args.no_cpu and rank == 0:
        line["cpu_baseline"] = cpu_baseline(ctx, capi, scene, cams, width, height, args)

    # ---- the same job through the single-process group API (meso_group_*: the C++ host's way in), rank 0 drives all N GPUs
    #      while the other ranks wait on a host-side barrier ----
    if world > 1 and not args.no_group:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                line["group_api"] = bench_group(capi, torch, scene, cams, rays_cam, width, height, world, args)
            except Exception as e:
                line["group_api"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        dist.barrier(group=cpu_group)
    if world > 1:
        if arrival is not None:
            timed_out = ctx.wait_timed_out()
            if rank == 0:
                line["config"]["rendezvous_wait_timed_out"] = timed_out
            arrival.close(dist)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit_line(line)
    ctx.close()
    return 0


def bench_group(capi, torch, scene, cams, rays_cam, width, height, world, args):
    """N GPUs behind one handle in ONE process (meso_group_*, csrc/meso_group.cu): what a C++ MesoEngine host calls.  Frames
    through the group's ring (slab gather fused into the kernels' stores, cross-device event waits, one DMA per member),
    quads gathered to the host at prefix offsets and on member 0 in segments, the edit loop through the synchronous calls."""
    origin, dims, params = scene
    g = capi.Group(list(range(world)))
    try:
        g.scene_create(origin, dims, max_bricks=(1 << 20) if dims[0] >= 32 else (1 << 18))
        g.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
        g.sync()
        RING = 4
        hosts = [torch.empty((height, width, 4), dtype=torch.int32).pin_memory() for _ in range(RING)]
        hosts_np = [h.numpy().view(capi.HitRecord).reshape(height, width) for h in hosts]
        consumed = 0

        def run(nsteps):
            nonlocal consumed
            for k in range(nsteps):
                slot = k % RING
                if k >= RING:
                    g.frame_wait(slot)
                    consumed += int(hosts_np[slot]["w1"][0, 0]) + int(hosts_np[slot]["w1"][-1, -1])
                g.raymarch_async(cams[k % 8], width, height, hosts_np[slot], slot, shadow=True, light=LIGHT)
            for slot in range(RING):
                g.frame_wait(slot)
                consumed += int(hosts_np[slot]["w1"][0, 0]) + int(hosts_np[slot]["w1"][-1, -1])

        run(RING)
        nsteps = max(8, min(args.steps, 40))
        t0 = time.perf_counter()
        run(nsteps)
        dt = time.perf_counter() - t0
        rays = sum(rays_cam[k % 8] for k in range(nsteps))
        # the group's frame == member 0 rendering the whole frame on its own
        got = g.raymarch(cams[0], width, height, shadow=True, light=LIGHT)
        m0 = g.member(0)
        m0.set_partition(0, 1)
        ref = m0.raymarch(cams[0], width, height, shadow=True, light=LIGHT)
        m0.set_partition(0, world)
        out = {"e2e": {"value": rays / dt / 1e6, "unit": "Mrays/s", "steps": nsteps, "d2h_bytes_per_step": 16 * width * height,
                       "frame_equal_to_1gpu_frame": bool(got.tobytes() == ref.tobytes())}}
        if not args.no_mesh:
            cap = 1 << 25
            qpin = torch.empty((cap, 4), dtype=torch.int32).pin_memory()
            qbuf = qpin.numpy().view(capi.Quad).reshape(-1)
            q, counts = g.mesh(cap, out=qbuf)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); g.mesh(cap, out=qbuf); ts.append(time.perf_counter() - t0)
            out["mesh_to_host"] = {"quads": int(len(q)), "ms": min(ts) * 1e3, "per_member": [int(x) for x in counts],
                                   "note": "members mesh their chunks, then copy their lists into pinned host memory at prefix offsets, all at once"}
            del qbuf, qpin
            dq = m0.device_alloc(cap * 16)
            n, _ = g.mesh_device(dq, cap, compact=False)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); g.mesh_device(dq, cap, compact=False); ts.append(time.perf_counter() - t0)
            ts2 = []
            for _ in range(3):
                t0 = time.perf_counter(); g.mesh_device(dq, cap, compact=True); ts2.append(time.perf_counter() - t0)
            m0.device_free(dq)
            out["mesh_on_member0"] = {"quads": int(n), "ms_segments": min(ts) * 1e3, "ms_compacted": min(ts2) * 1e3,
                                      "note": "wall clock of the call incl. the 8-byte count read-backs"}
            # edit loop, synchronous group calls
            frames = 32
            rec = g.raymarch(cams[0], width, height, shadow=True, light=LIGHT, out=hosts_np[0])
            t_all = time.perf_counter()
            nd_total = 0
            for k in range(frames):
                r = rec[height // 2, width // 2]
                if (int(r["w1"]) >> 20) & 1:
                    center = [int(r["w0"]) & 0xFFFF, int(r["w0"]) >> 16, int(r["w1"]) & 0xFFFF]
                    nd_total += g.carve_sphere(center, 24)
                    g.remesh_dirty(1 << 16)
                rec = g.raymarch(cams[(k + 1) % 8], width, height, shadow=True, light=LIGHT, out=hosts_np[0])
            t_all = time.perf_counter() - t_all
            out["edit_loop"] = {"frames": frames, "fps": frames / t_all, "ms_per_frame": t_all / frames * 1e3, "dirty_bricks_per_frame": nd_total / frames}
        out["note"] = "one process, one handle: meso_group_create over %d devices, peer access, no NCCL; what host/Samples/SimpleVoxel --gpus N uses" % world
        return out
    finally:
        g.close()


class ArrivalWords:
    """32-bit arrival words in every rank's device memory, opened by all ranks over CUDA IPC: the rendezvous of the N-GPU
    frame loops without a collective.  words[r] = base device address of rank r's block of `n` words (own or peer mapping).
    Every rank runs the same collectives here whatever fails where; .ok tells whether all ranks succeeded."""

    def __init__(self, ctx, capi, torch, dist, dev, rank, world, n=32):
        self.ctx, self.rank, self.world, self.n = ctx, rank, world, n
        self.words = [None] * world
        self.mine = None
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        h = torch.zeros(capi.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        try:
            self.mine = ctx.device_alloc(4 * n)
            ctx.device_memset(self.mine, 0, 4 * n)
            ctx.sync()
            h.copy_(torch.from_numpy(ctx.ipc_export(self.mine)))
        except Exception as e:
            sys.stderr.write("bench: arrival words unavailable on rank %d (%s)\n" % (rank, e))
            ok.zero_()
        allh = torch.zeros((world, capi.IPC_HANDLE_BYTES), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, h)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            hs = allh.cpu().numpy()
            try:
                for r in range(world):
                    self.words[r] = self.mine if r == rank else ctx.ipc_open(hs[r])
            except Exception as e:
                sys.stderr.write("bench: arrival words unavailable on rank %d (%s)\n" % (rank, e))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.ok = int(ok.item()) == 1
        self.gen = [0] * n          # launches signalled so far per word index (identical on every rank)

    def all_ranks(self, i):
        return [w + 4 * i for w in self.words]

    def close(self, dist):
        for r in range(self.world):
            if r != self.rank and self.words[r] is not None:
                try:
                    self.ctx.ipc_close(self.words[r])
                except Exception:
                    pass
        dist.barrier()
        if self.mine is not None:
            self.ctx.device_free(self.mine)


def bench_small_config(ctx, capi, scenes, torch, stream, name, args):
    """One of BASELINE.json's smaller raymarch configurations at its stated size: device-timed Mrays/s (kernel per frame, L2
    flushed: these scenes fit the L2) and, unless --no-cpu, the oracle's records on sampled scanlines byte-compared."""
    n, width, height, desc = WORKLOADS[name]
    scene = scenes.sphere_scene(n)
    origin, dims, params = scene
    ctx.scene_create(origin, dims, max_bricks=1 << 18)
    ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
    cams = make_cameras(scene, width, height)
    st = [ctx.raymarch_stats(c, width, height, shadow=True, light=LIGHT) for c in cams]
    rays = [int(s["primary"]) + int(s["shadow"]) for s in st]
    frame = torch.empty((height, width, 4), dtype=torch.int32, device="cuda")
    steps = 24
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(4):
        ctx.raymarch_device(cams[k % 8], width, height, frame.data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
    for k in range(steps):
        ctx.flush_l2()
        ev[k][0].record(stream)
        ctx.raymarch_device(cams[k % 8], width, height, frame.data_ptr(), shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    out = {"workload": desc, "mrays_s": float(np.mean([rays[k % 8] for k in range(steps)])) / ms / 1e3, "ms_per_frame": ms,
           "rays_per_frame_mean": float(np.mean(rays))}
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        vol = cpu_build_volume(orc, scene)
        rows = sample_rows(height, 8, 8)
        mism, cpu_rays, cpu_dt = 0, 0, 0.0
        nthreads = orc.hw_threads()
        for k in (0, 3, 6):
            t0 = time.perf_counter()
            r, rec = cpu_step(orc, vol, cams[k], width, height, rows, nthreads)
            cpu_dt += time.perf_counter() - t0
            cpu_rays += r
            gpu = ctx.raymarch(cams[k], width, height, shadow=True, light=LIGHT)
            mism += int((gpu[rows].view(np.uint32).reshape(-1, 4) != rec[rows].view(np.uint32).reshape(-1, 4)).any(axis=1).sum())
        out["parity_mismatches_on_sampled_rows"] = mism
        out["cpu_mrays_s"] = cpu_rays / cpu_dt / 1e6
        out["cpu_sample"] = "3 cameras x 64 scanlines x %d px, oracle on %d threads" % (width, nthreads)
    return out


def bench_edit_loop(ctx, capi, cams, width, height, frames=48):
    """SURVEY.md 8d edit loop: frame k carves a sphere (r = 24 voxels) at the hit point of the centre pixel of camera
    C_{k mod 8}, re-meshes the dirty bricks (+ their six neighbours) and re-renders 4K.  Host-API calls, wall clock."""
    import torch
    t_carve = t_mesh = t_render = 0.0
    dirty_total = quads_total = 0
    pinned = torch.empty((height, width, 4), dtype=torch.int32).pin_memory()
    out = pinned.numpy().view(capi.HitRecord).reshape(height, width)
    rec = ctx.raymarch(cams[0], width, height, shadow=True, light=LIGHT, out=out)
    t_all = time.perf_counter()
    for k in range(frames):
        r = rec[height // 2, width // 2]
        if (int(r["w1"]) >> 20) & 1:
            center = [int(r["w0"]) & 0xFFFF, int(r["w0"]) >> 16, int(r["w1"]) & 0xFFFF]
            t0 = time.perf_counter()
            nd = ctx.carve_sphere(center, 24)
            t1 = time.perf_counter()
            quads, keys = ctx.remesh_dirty(1 << 16, 1 << 13)
            t2 = time.perf_counter()
            t_carve += t1 - t0; t_mesh += t2 - t1
            dirty_total += nd; quads_total += len(quads)
        t0 = time.perf_counter()
        rec = ctx.raymarch(cams[(k + 1) % 8], width, height, shadow=True, light=LIGHT, out=out)
        t_render += time.perf_counter() - t0
    t_all = time.perf_counter() - t_all
    return {"frames": frames, "fps": frames / t_all, "ms_per_frame": t_all / frames * 1e3, "ms_carve": t_carve / frames * 1e3,
            "ms_remesh_dirty": t_mesh / frames * 1e3, "ms_render_to_host": t_render / frames * 1e3,
            "dirty_bricks_per_frame": dirty_total / frames, "quads_per_frame": quads_total / frames,
            "note": "carve r=24 voxels at the centre-pixel hit, re-mesh dirty bricks + neighbours, re-render 3840x2160 to host memory (synchronous API)"}


def bench_edit_loop_multi(ctx, capi, cams, width, height, frame, frame_dptr, flag, stream, rank, dist, torch, frames=48, slabs=None, rgba8=False):
    """The edit loop on N GPUs.  Every rank reads the centre-pixel hit of the last frame, carves the same sphere into its
    replica of the volume (replicated compute), re-meshes ITS SHARE of the dirty bricks (sharded by key hash) and renders its
    tiles of the next 4K frame.  With the slab gather the loop is pipelined: the only thing frame k+1's carve needs from
    frame k is one 16-byte record, read from the owning rank's slab right behind the kernels' rendezvous, so the slab-to-host
    DMAs of frame k run behind the carve, re-mesh and traversal of frame k+1 (two slab sets alternate).  Wall clock between
    two barriers, every frame's records in host memory at the end."""
    dist.barrier()
    torch.cuda.synchronize()
    py, pxl = height // 2, width // 2
    pick = np.zeros(1, dtype=capi.HitRecord)
    t_carve = t_mesh = t_render = 0.0
    dirty_total = quads_total = 0

    if slabs is None:
        def render(k):
            ctx.raymarch_device(cams[k % 8], width, height, frame_dptr, shadow=True, light=LIGHT, layout=capi.LAYOUT_FRAME)
            with torch.cuda.stream(stream):
                dist.all_reduce(flag)
            stream.synchronize()
            return frame[py, pxl]
        finish = lambda: None
    else:
        R, rows, px = slabs["R"], slabs["rows"], slabs["px"]
        owner = py // rows
        copy_streams = [torch.cuda.Stream() for _ in range(2)]
        dma_done = [None, None]
        # rgba8: the records stay in the slabs (the pick needs them), only their colour words are packed and sent to the host
        packed = [torch.empty((max(slabs["my_rows"], 1), width), dtype=torch.int32, device="cuda") for _ in range(2)] if rgba8 else None

        def render(k):
            slot = k % 2
            if dma_done[slot] is not None:
                dma_done[slot].synchronize()          # this slab set's previous frame has left for the host
            arr = slabs.get("arrival")
            if arr is not None:
                ia, ib = 24 + slot, 26 + slot
                # every rank's copy of the frame that used this slab set two frames ago has left for the host (the kernels of
                # this frame store into peers' slabs): wait for all ranks' "slab has reached the host" signals of that frame
                if arr.gen[ib] > 0:
                    ctx.wait_device(arr.mine + 4 * ib, arr.gen[ib])
                ctx.raymarch_device_slabs(cams[k % 8], width, height, slabs["ptrs"][slot], rows, shadow=True, light=LIGHT)
                ctx.signal_device(arr.all_ranks(ia))
                arr.gen[ia] += slabs["world"]
                ctx.wait_device(arr.mine + 4 * ia, arr.gen[ia])      # every rank's kernel is done: all slabs of this frame are complete
                with torch.cuda.stream(stream):
                    ev = torch.cuda.Event()
                    ev.record(stream)
            else:
                ctx.raymarch_device_slabs(cams[k % 8], width, height, slabs["ptrs"][slot], rows, shadow=True, light=LIGHT)
                with torch.cuda.stream(stream):
                    # this rank's DMA of the PREVIOUS frame (other slab set) is ordered before the rendezvous, so that the
                    # rendezvous also tells every rank that all slabs of that set have left: the frame after this one may then
                    # overwrite them (other ranks' kernels store into this rank's slab)
                    if dma_done[1 - slot] is not None:
                        stream.wait_event(dma_done[1 - slot])
                    dist.all_reduce(flag)                 # every rank's kernel is done: all slabs of this frame are complete
                    ev = torch.cuda.Event()
                    ev.record(stream)
            # the pick: one record from the owning rank's slab (peer memory), on the main stream
            ctx.download(pick, slabs["ptrs"][slot][owner] + ((py - owner * rows) * width + pxl) * 16, 16)
            # the frame itself leaves on a copy stream, behind the next frame's work
            cs = copy_streams[slot]
            cs.wait_event(ev)
            if slabs["my_rows"] > 0:
                ctx.set_stream(cs.cuda_stream)
                if rgba8:
                    off = slot * px * 16 + slabs["row0"] * width * 4
                    ctx.pack_rgba8_device(slabs["mine"][slot], slabs["my_rows"] * width, packed[slot].data_ptr())
                    ctx.download_async(slabs["shm"][off:off + slabs["my_rows"] * width * 4], packed[slot].data_ptr())
                else:
                    off = slot * px * 16 + slabs["row0"] * width * 16
                    ctx.download_async(slabs["shm"][off:off + slabs["my_rows"] * width * 16], slabs["mine"][slot])
                ctx.set_stream(stream.cuda_stream)
            if arr is not None:
                ctx.set_stream(cs.cuda_stream)
                ctx.signal_device(arr.all_ranks(26 + slot))          # "my slab of this frame has reached the host", to every rank
                ctx.set_stream(stream.cuda_stream)
                arr.gen[26 + slot] += slabs["world"]
            e = torch.cuda.Event()
            e.record(cs)
            dma_done[slot] = e
            return pick[0]

        def finish():
            for e in dma_done:
                if e is not None:
                    e.synchronize()

    r = render(0)
    t_all = time.perf_counter()
    for k in range(frames):
        if (int(r["w1"]) >> 20) & 1:
            center = [int(r["w0"]) & 0xFFFF, int(r["w0"]) >> 16, int(r["w1"]) & 0xFFFF]
            t0 = time.perf_counter()
            nd = ctx.carve_sphere(center, 24)
            t1 = time.perf_counter()
            quads, keys = ctx.remesh_dirty(1 << 16, 1 << 13)     # this rank's share of the dirty bricks (+ neighbours)
            quads_total += len(quads)
            t2 = time.perf_counter()
            t_carve += t1 - t0; t_mesh += t2 - t1
            dirty_total += nd
        t0 = time.perf_counter()
        r = render(k + 1)
        t_render += time.perf_counter() - t0
    finish()
    t_all = time.perf_counter() - t_all
    dist.barrier()
    qt = torch.tensor([quads_total], dtype=torch.int64, device="cuda")
    dist.all_reduce(qt)
    return {"frames": frames, "fps": frames / t_all, "ms_per_frame": t_all / frames * 1e3, "ms_carve": t_carve / frames * 1e3,
            "ms_remesh_dirty": t_mesh / frames * 1e3, "ms_render_to_host": t_render / frames * 1e3,
            "dirty_bricks_per_frame": dirty_total / frames, "quads_per_frame": int(qt.item()) / frames,
            "pipelined": slabs is not None, "host_bytes_per_frame": (4 if rgba8 else 16) * width * height,
            "note": "rank 0's clock; carve r=24 voxels at the centre-pixel hit on every rank's replica, dirty bricks re-meshed sharded over the ranks by key hash, 3840x2160 re-rendered by all ranks"
                    + (" through the slab gather: the next carve waits only for the 16-byte pick behind the kernels' rendezvous, the slab-to-host DMAs of frame k overlap frame k+1"
                       if slabs is not None else " into the shared host frame (host-fused stores), one rendezvous per frame")}


def bench_mesh_multi(ctx, capi, torch, dist, dev, rank, world, stream, n_voxels, steps=8):
    """Meshing on N GPUs: chunk c belongs to rank c % world (replicated volume), quads gathered on rank 0.

    Default gather = SEGMENTS: rank 0 owns the list (exported over CUDA IPC); rank r's mesh kernel writes its quads straight
    into segment r = [r * cap / N, ...) of it over NVLink while it is meshing, reserving slots on a counter in its OWN
    memory (no cross-GPU atomic, no second pass); the ranks' 8-byte counts are all-gathered.  The result is the list a
    renderer draws from (N ranges); `compact_ms` is what closing the gaps on rank 0 costs when one contiguous list is wanted.
    Also timed: the round-1 form (one shared counter, system-scope atomics over NVLink) and the per-GPU lists without any
    gather.  Time = CUDA events on every rank around kernel + rendezvous, max over ranks.  Returns None without IPC."""
    cap = 1 << 25
    seg = cap // world
    # every rank runs the same collectives whatever fails where
    hq = torch.zeros(capi.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    hc = torch.zeros(capi.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    ok = torch.zeros(1, dtype=torch.int32, device=dev)
    qptr0 = cptr0 = qptr = cptr = None
    if rank == 0:
        try:
            qptr0, cptr0 = ctx.device_alloc(cap * 16), ctx.device_alloc(8)
            hq.copy_(torch.from_numpy(ctx.ipc_export(qptr0)))
            hc.copy_(torch.from_numpy(ctx.ipc_export(cptr0)))
            ok.fill_(1)
        except Exception as e:
            sys.stderr.write("bench: fused quad gather unavailable (%s)\n" % e)
    dist.broadcast(ok, src=0)
    dist.broadcast(hq, src=0)
    dist.broadcast(hc, src=0)
    if int(ok.item()) == 1:
        try:
            qptr = qptr0 if rank == 0 else ctx.ipc_open(hq.cpu().numpy())
            cptr = cptr0 if rank == 0 else ctx.ipc_open(hc.cpu().numpy())
        except Exception as e:
            sys.stderr.write("bench: fused quad gather unavailable on rank %d (%s)\n" % (rank, e))
            ok.fill_(0)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        if rank != 0:
            for p in (qptr, cptr):
                if p is not None:
                    ctx.ipc_close(p)
        dist.barrier()
        if rank == 0:
            for p in (qptr0, cptr0):
                if p is not None:
                    ctx.device_free(p)
        return None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    mine = torch.zeros(1, dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)

    def timed(body):
        with torch.cuda.stream(stream):
            dist.all_reduce(flag)            # start together
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            body()
            b.record(stream)
        stream.synchronize()
        return a.elapsed_time(b)

    def avg(body):
        timed(body)
        ts = [timed(body) for _ in range(steps)]
        t = torch.tensor([sum(ts) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # (1) segments: kernel writes into rank 0's list at this rank's segment, count stays local; rendezvous = the all-gather
    def seg_body():
        ctx.mesh_device(qptr + rank * seg * 16, seg, want_count=False)
        ctx.mesh_count_device(mine.data_ptr())                 # 8-byte count -> device tensor (stream-ordered, no host wait)
        dist.all_gather_into_tensor(counts, mine)              # behind it every rank's quads are in the list
    ms_seg = avg(seg_body)
    seg_counts = [int(x) for x in counts.cpu()]
    got = None
    if rank == 0:
        parts = []
        for r in range(world):
            part = np.zeros((seg_counts[r], 4), dtype=np.uint32)
            if seg_counts[r]:
                ctx.download(part, qptr + r * seg * 16, seg_counts[r] * 16)
            parts.append(part)
        got = np.concatenate(parts, axis=0)
    # what closing the gaps costs (rank 0, device-local copies on its own stream)
    compact_ms = None
    if rank == 0:
        def compact():
            at = seg_counts[0]
            for r in range(1, world):
                n, src, dst = seg_counts[r], r * seg, at
                done = 0
                while done < n and src > dst:
                    stepn = min(n - done, src - dst)       # chunks no longer than the gap never overlap their source
                    ctx.device_copy(qptr + (dst + done) * 16, qptr + (src + done) * 16, stepn * 16)
                    done += stepn
                at += n
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in ev:
            ctx.mesh_device(qptr, seg, want_count=False)        # restore segment 0..: contents irrelevant for the timing
            a.record(stream); compact(); b.record(stream)
        stream.synchronize()
        compact_ms = min(a.elapsed_time(b) for a, b in ev)
    dist.barrier()

    # (2) round 1's form: one shared counter on rank 0, system-scope atomics over NVLink
    def shared_body():
        ctx.mesh_device_shared(qptr, cptr, cap)
        dist.all_reduce(flag)
    def shared_timed():
        with torch.cuda.stream(stream):
            if rank == 0:
                ctx.device_memset(cptr, 0, 8)
        return timed(shared_body)
    shared_timed()
    ts = [shared_timed() for _ in range(max(2, steps // 2))]
    t_sh = torch.tensor([sum(ts) / len(ts)], dtype=torch.float64, device=dev)
    dist.all_reduce(t_sh, op=dist.ReduceOp.MAX)

    # (3) the same partition with the lists left where they are made (one list per GPU, no gather)
    local = torch.empty((seg + (1 << 20), 4), dtype=torch.int32, device=dev)
    ms_local = avg(lambda: ctx.mesh_device(local.data_ptr(), local.shape[0], want_count=False))
    del local

    out = None
    if rank == 0:
        # the same mesh on one GPU, compared through an order-independent fingerprint (the list order is scheduling-dependent)
        ctx.set_partition(0, 1)
        ref_t = torch.empty((cap, 4), dtype=torch.int32, device=dev)
        nref = ctx.mesh_device(ref_t.data_ptr(), cap)
        ref = ref_t[:nref].cpu().numpy().view(np.uint32)
        ctx.set_partition(rank, world)
        fp = lambda q: (int(q.shape[0]), [int(x) for x in q.astype(np.uint64).sum(axis=0)], [int(x) for x in np.bitwise_xor.reduce(q, axis=0)])
        nq = int(got.shape[0])
        out = {"scene_voxels": n_voxels, "quads": nq, "ms": ms_seg, "meshed_voxels_per_s": float(n_voxels) ** 3 / (ms_seg * 1e-3),
               "equal_to_1gpu_mesh": bool(fp(got) == fp(ref)), "segment_counts": seg_counts, "compact_ms": compact_ms,
               "shared_counter": {"ms": float(t_sh.item()), "note": "round-1 form: one counter on rank 0, a system-scope atomicAdd per warp over NVLink"},
               "sharded_lists": {"ms": ms_local, "meshed_voxels_per_s": float(n_voxels) ** 3 / (ms_local * 1e-3),
                                 "note": "same partition, every rank keeps its own quad list (no gather): max over ranks of the mesh kernels"},
               "note": "chunk c -> rank c % N; segmented gather: every rank's mesh kernel stores its quads into its own segment of rank 0's list over NVLink (local counter, 16 B stores), counts all-gathered; fingerprint = (count, column sums, column xors)"}
        del ref_t
    dist.barrier()
    if rank != 0:
        ctx.ipc_close(qptr); ctx.ipc_close(cptr)
    dist.barrier()
    if rank == 0:
        ctx.device_free(qptr0); ctx.device_free(cptr0)
    return out


def bench_stream(ctx, capi, torch, stream):
    """K6: FChunkManage::UpdateChunks + UpdateLoadingQueue on the reference's defaults (TestGenerator terrain, one sample
    per block, view radius 24 / 6 chunks, 120 degrees, 256 chunks per update) in a 49 x 6 x 49-chunk window around the
    camera chunk: updates until the desired set is resident, then a 90-degree turn.  CUDA events on the launching stream."""
    origin, dims = (-24, -3, -24), (49, 6, 49)
    ctx.scene_create(origin, dims, 1 << 16)
    ctx.stream_begin(capi.SDF_TERRAIN, None, capi.GRAN_BLOCK)
    view = capi.view_config()

    def until_resident(fwd):
        ms, gen, n = 0.0, 0, 0
        while n < 400:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ctx.stream_update((0, 0, 0), fwd, 256, view, wait=False); b.record(stream)
            torch.cuda.synchronize()
            st = ctx.stream_stats()
            ms += a.elapsed_time(b); gen += int(st["generated"]); n += 1
            if int(st["missing"]) == 0:
                break
        return ms, gen, n, st

    ms0, gen0, n0, st0 = until_resident((0.2, 0.1, 0.97))
    ms1, gen1, n1, _ = until_resident((0.97, 0.1, -0.2))
    idle = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.stream_update((0, 0, 0), (0.97, 0.1, -0.2), 256, view, wait=False); b.record(stream)
        torch.cuda.synchronize()
        idle.append(a.elapsed_time(b))
    return {"window_chunks": int(np.prod(dims)), "desired_set": int(st0["candidates"]), "desired_in_window": int(st0["in_window"]),
            "fill": {"updates": n0, "chunks": gen0, "ms": ms0, "chunks_per_s": gen0 / (ms0 * 1e-3)},
            "turn_90deg": {"updates": n1, "chunks": gen1, "ms": ms1},
            "ms_update_nothing_missing": float(np.median(idle)),
            "note": "select (117 649 offsets -> desired set, rank-sorted) + dispatch list + generation of <= 256 chunks + derived data per update, no host round trip inside an update"}


def bench_mesh(ctx, capi, scenes, torch, stream, args, dev, n_work=1024):
    """Face-cull + greedy meshing throughput (meshed voxels/s = N^3 / time): BASELINE.json configs[2] (1024^3 terrain), the
    1024^3 sphere, and the raymarch workload's own scene (the one the N-GPU mesh leg uses)."""
    out = {}
    cases = [("terrain_1024_blocks", 1024, capi.SDF_TERRAIN, capi.GRAN_BLOCK, scenes.terrain_scene(1024)),
             ("sphere_1024_voxels", 1024, capi.SDF_SPHERE, capi.GRAN_VOXEL, scenes.sphere_scene(1024))]
    if n_work != 1024:
        cases.append(("sphere_%d_voxels" % n_work, n_work, capi.SDF_SPHERE, capi.GRAN_VOXEL, scenes.sphere_scene(n_work)))
    for name, n, kind, gran, scene in cases:
        origin, dims, params = scene
        ctx.scene_create(origin, dims, (1 << 20) if n > 1024 else (1 << 18))
        ctx.voxelize_sdf(kind, params, gran)
        cap = (1 << 25) if n > 1024 else (1 << 24)
        quads = torch.empty((cap, 4), dtype=torch.int32, device=dev)
        nq = ctx.mesh_device(quads.data_ptr(), cap)
        steps = max(5, min(args.steps, 30))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            ctx.flush_l2()
            evs[k][0].record(stream)
            ctx.mesh_device(quads.data_ptr(), cap, want_count=False)
            evs[k][1].record(stream)
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        occ, full, keys, _ = ctx.volume_download()
        populated = int(np.unpackbits(occ.view(np.uint8)).sum())
        alg = 64.0 * len(keys) + 1024.0 * int(np.prod(dims)) + 16.0 * nq + 4
        # second merge level (oracle/orc_mesh.c): faces of full bricks towards absent bricks, merged per chunk -- how many quads
        # they became and how many 8x8 brick faces (one quad each before this level existed) those quads cover
        q1 = quads[:nq, 1]
        lvl = ((q1 >> 19) & 1) == 1
        brick_faces = int((((q1[lvl] >> 27) & 0x1F).to(torch.int64) * (quads[:nq, 2][lvl] >> 3).to(torch.int64)).sum().item())
        out[name] = {"meshed_voxels_per_s": n ** 3 / (ms * 1e-3), "populated_voxels_per_s": populated * 512 / (ms * 1e-3),
                     "quads_per_s": int(nq) / (ms * 1e-3), "ms": ms, "quads": int(nq), "populated_bricks": populated,
                     "partial_bricks": int(len(keys)), "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9,
                     "brick_level": {"quads": int(lvl.sum().item()), "brick_faces_covered": brick_faces}}
        if name == "terrain_1024_blocks" and not args.no_cpu:
            # BASELINE.json configs[2], CPU beside GPU: the oracle meshes the very same volume on all host threads; the two
            # quad lists are compared in full after the canonical sort (bit-exact), not sampled
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import orc
            occ_f, full_f, keys_f, payload_f = ctx.volume_download()
            vol = orc.Volume(origin, dims).import_(occ_f, full_f, keys_f, payload_f)
            nthreads = orc.hw_threads()
            vol.mesh(nthreads)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); ref_q = vol.mesh(nthreads); ts.append(time.perf_counter() - t0)
            got_q = quads[:nq].cpu().numpy().view(np.uint32).reshape(-1, 4).copy().view(orc.Quad).reshape(-1)
            equal = bool(len(ref_q) == nq and orc.sort_quads(got_q).tobytes() == orc.sort_quads(ref_q).tobytes())
            out[name]["cpu_baseline"] = {"ms": min(ts) * 1e3, "meshed_voxels_per_s": n ** 3 / min(ts), "cores": nthreads, "kind": "port",
                                         "quad_lists_bit_exact_after_canonical_sort": equal,
                                         "note": "oracle/orc_mesh.c (count pass + emit pass) on the same volume, all host threads"}
        del quads
    return out


def cpu_baseline(ctx, capi, scene, cams, width, height, args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    origin, dims, params = scene
    # re-create the raymarch scene (the mesh leg replaced it) and hand the oracle the very same volume
    ctx.scene_create(origin, dims, max_bricks=(1 << 20) if dims[0] >= 32 else (1 << 18))
    ctx.voxelize_sdf(capi.SDF_SPHERE, params, capi.GRAN_VOXEL)
    occ, full, keys, payload = ctx.volume_download()
    vol = orc.Volume(origin, dims).import_(occ, full, keys, payload)
    nthreads = orc.hw_threads()
    n_bands = max(8, min(64, 2 * nthreads))
    rows = sample_rows(height, n_bands, 8)
    buf = np.zeros((height, width), dtype=orc.HitRecord)
    cpu_step(orc, vol, cams[0], width, height, rows, nthreads, out=buf)
    rays = 0
    mism = 0
    n_steps = 16
    dt = 0.0
    for k in range(n_steps):
        t1 = time.perf_counter()
        r, rec = cpu_step(orc, vol, cams[k % 8], width, height, rows, nthreads, out=buf)
        dt += time.perf_counter() - t1
        rays += r
        # byte-compare the same scanlines of the GPU frame (outside the CPU timing)
        gpu = ctx.raymarch(cams[k % 8], width, height, shadow=True, light=LIGHT)
        mism += int((gpu[rows].view(np.uint32).reshape(-1, 4) != rec[rows].view(np.uint32).reshape(-1, 4)).any(axis=1).sum())
    scaling = cpu_thread_scaling(orc, vol, cams[0], width, height, rows, nthreads)
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": nthreads, "kind": "port",
            "sample": "%d steps x %d bands x 8 rows x %d px in one parallel region per step (cameras cycled); records byte-compared with the GPU frame: %d mismatching pixels"
                      % (n_steps, n_bands, width, mism),
            "thread_scaling_mrays_s": scaling, "parity_mismatches": mism}


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner with printf when
    NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the rest of the process and the line goes out
    through a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--regions", type=int, default=5, help="number of K-step timed regions; the median is reported")
    ap.add_argument("--no-mesh", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--frames-in-flight", type=int, default=4, help="frame ring depth of the timed loop (reference: kNumBufferedFrames = 4)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="multi-GPU frame gather (p2p falls back to nccl if IPC is unavailable)")
    ap.add_argument("--rendezvous", default="arrival", choices=["arrival", "nccl"], help="N > 1 frame loops: arrival words over peer memory, or NCCL all-reduce")
    ap.add_argument("--no-group", action="store_true", help="N > 1: skip the single-process group-API leg")
    ap.add_argument("--no-slabs", action="store_true", help="N > 1 e2e: host-fused 512-byte stores instead of the slab gather")
    ap.add_argument("--no-host-fused", action="store_true", help="N > 1 e2e: gather on rank 0 and copy instead of storing straight into shared host memory")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (world == 1 and args.gpus > 1 and args.impl == "ours"):    # (the torchrun re-launch below passes stdout through)
        claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if world != args.gpus:
        if args.gpus == 1 and world == 1:
            pass
        elif world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
