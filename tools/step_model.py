#!/usr/bin/env python
"""Step-count model of raymarch acceleration structures, on the CPU (no GPU needed).

The raymarch kernel is bound by instruction issue under divergence, and every experiment of round 1 obeyed one rule:
time follows the number of steps as long as no new code path is added (profiles/README.md).  Any certified-empty box is
a legal skip that leaves the records unchanged, so candidate structures can be compared by their STEP COUNTS before a
kernel is written: the oracle's ORC_DDA_MODEL walker (oracle/orc_raymarch.c) takes a configuration (field cell size,
cap, probe-ahead, per-octant forward cubes, brick-level cubes, 2^3 cells), produces the same records as every other
walk (tests/test_oracle_raymarch.py) and counts its steps by kind.

The first configuration is the shipped v8 kernel; its modelled counts are printed next to the counts the GPU measured
(profiles/r1_divergence_4096_v8.jsonl, STATS build) as the calibration of the model.

    python tools/step_model.py [--n 4096] [--cams 0,1,2] [--out profiles/r1_step_model.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = [
    ("v8 shipped: 32^3 field cap 32 + probe, brick, 2^3 cell, voxel", dict()),
    ("v8 without probe-ahead", dict(probe=False)),
    ("per-octant forward cubes, 32^3 cells (one lookup, no probe)", dict(directional=True)),
    ("16^3 field cap 32 + probe", dict(df_shift=4)),
    ("16^3 per-octant forward cubes", dict(df_shift=4, directional=True)),
    ("8^3 (brick) per-octant forward cubes cap 32: replaces field AND brick level", dict(df_shift=3, directional=True)),
    ("v8 + brick-level forward cubes up to 4 bricks", dict(brick_cap=4)),
    ("32^3 per-octant cubes + brick-level forward cubes up to 4 bricks", dict(directional=True, brick_cap=4)),
    ("32^3 per-octant cubes + brick-level forward cubes up to 8 bricks", dict(directional=True, brick_cap=8)),
    ("v8 + forward cubes of up to 4 empty 2^3 cells inside the brick", dict(cell2=4)),
    ("32^3 per-octant cubes + brick cubes <= 4 + 2^3-cell cubes <= 4", dict(directional=True, brick_cap=4, cell2=4)),
    ("v8 without the 2^3-cell level (v7-like)", dict(cell2=False)),
]
KINDS = ["voxel", "cell2", "brick", "field_le2", "field_gt2", "entry"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--cams", default="0,1,2,3,4,5,6,7")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_step_model.json"))
    args = ap.parse_args()
    import orc
    from mesoengine_b200 import scenes
    w, h = (3840, 2160) if args.n >= 4096 else (1920, 1080)
    scene = scenes.sphere_scene(args.n)
    origin, dims, params = scene
    t0 = time.time()
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL, fast=True)
    print("scene %d^3 built in %.1f s" % (args.n, time.time() - t0), flush=True)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cams = [int(c) for c in args.cams.split(",")]
    rows, stride = 8, 64
    bands = [(y, y + rows) for y in range(stride // 2, h - rows, stride)]
    scale = h / float(len(bands) * rows)
    light = (0.3, 0.5, 0.8)

    measured = {}
    mpath = os.path.join(ROOT, "profiles", "r1_divergence_4096_v8.jsonl")
    if args.n == 4096 and os.path.exists(mpath):
        for line in open(mpath):
            d = json.loads(line)
            if "camera" in d:
                measured[d["camera"]] = d

    results = []
    ref_records = {}
    step_img = np.zeros((h, w, 2), dtype=np.uint32)
    orc.set_step_image(step_img, w)

    def warp_slots(y0, y1):
        """32 x the longest lane of every warp (8 x 4 pixels, as the kernel maps them), primary and shadow phase."""
        blk = step_img[y0:y1].reshape((y1 - y0) // 4, 4, w // 8, 8, 2)
        mx = blk.max(axis=(1, 3)).astype(np.int64)
        return 32 * int(mx[..., 0].sum()), 32 * int(mx[..., 1].sum())

    for name, cfg in CONFIGS:
        t0 = time.time()
        orc.step_model(vol, **cfg)
        t_build = time.time() - t0
        per_cam = {}
        for ci in cams:
            cam = orc.camera_uniform(eyes[ci], ctr, width=w, height=h)
            rs = orc.ray_setup(cam, origin, w, h, light)
            orc.step_model_counts(reset=True)
            prim = shad = 0
            sig = 0
            slots_p = slots_s = 0
            for (y0, y1) in bands:
                rec, st = vol.raymarch(rs, w, h, rect=(0, y0, w, y1), shadow=True, mode=orc.DDA_MODEL, stats=True)
                prim += int(st["primary"]); shad += int(st["shadow"])
                sig ^= hash(rec[y0:y1].tobytes())
                a, b = warp_slots(y0, y1)
                slots_p += a; slots_s += b
            # every configuration must see the same frame
            assert ref_records.setdefault(ci, sig) == sig, "records differ between configurations"
            c = orc.step_model_counts()
            per_cam[ci] = {"primary": prim, "shadow": shad, "steps": {k: int(v) for k, v in zip(KINDS, c)},
                           "warp_slots_primary": slots_p, "warp_slots_shadow": slots_s}
        tot = {k: sum(per_cam[ci]["steps"][k] for ci in cams) for k in KINDS}
        rays = sum(per_cam[ci]["primary"] + per_cam[ci]["shadow"] for ci in cams)
        total_steps = sum(tot.values())
        slots = sum(per_cam[ci]["warp_slots_primary"] + per_cam[ci]["warp_slots_shadow"] for ci in cams)
        results.append({"config": name, "params": cfg, "field_build_s": t_build, "rays_sampled": rays,
                        "steps_per_ray": total_steps / rays, "warp_slots_per_ray": slots / rays, "lane_use": total_steps / slots, "by_kind_per_ray": {k: tot[k] / rays for k in KINDS}, "per_camera": per_cam})
        print("%-82s %6.2f steps/ray  %6.2f warp slots/ray  %s" % (name, total_steps / rays, slots / rays, "  ".join("%s %.2f" % (k, tot[k] / rays) for k in KINDS)), flush=True)

    calib = None
    if measured:
        base = results[0]
        rows_out = []
        for ci in cams:
            if ci not in measured:
                continue
            m = measured[ci]
            mv = dict(zip(["voxel", "cell2", "brick", "field_le2", "field_gt2"], m["level_steps(voxel,cell2,brick,df_le2,df_gt2)"]))
            mrays = m["primary"] + m["shadow"]
            mod = base["per_camera"][ci]
            mo_rays = mod["primary"] + mod["shadow"]
            rows_out.append({"camera": ci, "gpu_warp_slots_per_ray": (m["warp_slots_primary"] + m["warp_slots_shadow"]) / mrays,
                             "model_warp_slots_per_ray": (mod["warp_slots_primary"] + mod["warp_slots_shadow"]) / mo_rays,
                             "gpu_steps_per_ray": sum(mv.values()) / mrays,
                             "model_steps_per_ray": sum(v for k, v in mod["steps"].items() if k != "entry") / mo_rays,
                             "gpu_by_kind_per_ray": {k: v / mrays for k, v in mv.items()},
                             "model_by_kind_per_ray": {k: v / mo_rays for k, v in mod["steps"].items()}})
        calib = rows_out
        for r in rows_out:
            print("camera %d: GPU %.2f steps/ray, model %.2f;  GPU %.2f warp slots/ray, model %.2f" %
                  (r["camera"], r["gpu_steps_per_ray"], r["model_steps_per_ray"], r["gpu_warp_slots_per_ray"], r["model_warp_slots_per_ray"]))
    base_spr = results[0]["steps_per_ray"]
    base_slots = results[0]["warp_slots_per_ray"]
    for r in results:
        r["relative_to_shipped"] = r["steps_per_ray"] / base_spr
        r["warp_slots_relative_to_shipped"] = r["warp_slots_per_ray"] / base_slots
    orc.set_step_image(None)
    with open(args.out, "w") as f:
        json.dump({"scene": "V-sphere %d^3 voxel-granular" % args.n, "resolution": [w, h], "sample": "%d bands x %d rows (1/%.2f of the frame), primary + shadow" % (len(bands), rows, scale),
                   "cameras": cams, "configs": results, "calibration_vs_gpu_stats_build": calib}, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
