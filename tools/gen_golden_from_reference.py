#!/usr/bin/env python
"""Generates tests/golden/*.json from the reference SOURCES at /root/reference (read-only, not available on the GPU box).

The reference ships no tests or fixtures, but a few of its data tables and constants pin behaviour
of the hot path.  This script parses them straight out of the reference files and commits them as small fixtures, so the
oracle is checked against the reference's own data and not against a re-typed copy:

  triplanar_faces.json   the 56-corner table + fan index list (Samples/SimpleVoxel.cpp:87-127,
                         Runtimes/Shape/TriplePlanarCube.h:36-43) reduced to: for each octant id, the set of cube faces
                         covered by the six fan triangles, and the local corner each fan is anchored at.
  constants.json         scene constants the path depends on (Runtimes/Voxel/VoxelSceneConfig.h:20-50, sphere generator
                         centre/radius GeneratorHelper.h:134, hash constants VoxelMathHelper.h:32, camera defaults
                         VoxelWindowsInstance.h:22-28, VoxelCamera.h:23, erode offsets BinaryOccupancyVolume.h:47-60).
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def read(rel):
    with open(os.path.join(REF, rel), encoding="utf-8", errors="replace") as f:
        return f.read()


def triplanar():
    src = read("Samples/SimpleVoxel.cpp")
    body = src[src.index("const vec3 TriplanarPositions[56]"):]
    body = body[: body.index(");")]
    sign = {"-BOX_SIZE": -1, "BOX_SIZE": 1}
    corners = [[sign[a.strip()], sign[b.strip()], sign[c.strip()]]
               for a, b, c in re.findall(r"vec3\(\s*(-?BOX_SIZE)\s*,\s*(-?BOX_SIZE)\s*,\s*(-?BOX_SIZE)\s*\)", body)]
    assert len(corners) == 56, len(corners)
    tp = read("Runtimes/Shape/TriplePlanarCube.h")
    idx_body = tp[tp.index("IndexData ="):]
    idx_body = idx_body[: idx_body.index("};")]
    fan = [int(x) for x in re.findall(r"\b(\d+)\b", idx_body)]
    assert len(fan) == 8, fan
    out = {"fan_indices": fan, "octants": []}
    for octant in range(8):
        row = corners[octant * 7: octant * 7 + 7]
        v = [row[i] for i in fan]
        faces = set()
        for k in range(1, len(v) - 1):  # triangle fan: (v0, vk, vk+1)
            tri = [v[0], v[k], v[k + 1]]
            shared = [ax for ax in range(3) if tri[0][ax] == tri[1][ax] == tri[2][ax]]
            assert len(shared) == 1, (octant, tri)
            ax = shared[0]
            faces.add(2 * ax + (1 if tri[0][ax] > 0 else 0))
        out["octants"].append({"octant": octant, "anchor_corner": row[0], "faces": sorted(faces)})
    # octant id bits (GetOctantId): bit0 x<0, bit1 y<0, bit2 z<0
    oid = src[src.index("int GetOctantId"):]
    oid = oid[: oid.index("return id")]
    out["octant_bits"] = re.findall(r"v\.([xyz]) < 0\.0\) id \|= (\d)", oid)
    return out


def constants():
    cfg = read("Runtimes/Voxel/VoxelSceneConfig.h")
    gen = read("Runtimes/Helper/GeneratorHelper.h")
    mh = read("Runtimes/Helper/VoxelMathHelper.h")
    inst = read("Runtimes/Instance/VoxelWindowsInstance.h")
    cam = read("Runtimes/Instance/VoxelCamera.h")
    occ = read("Runtimes/Voxel/Occupancy/BinaryOccupancyVolume.h")
    out = {}
    for name in ("BlockResolution", "ChunkResolution", "ChunkOccupancyDepth", "ChunkInnerVoxelCullDepthThreshold"):
        out[name] = int(re.search(name + r"\s*=\s*(\d+)", cfg).group(1))
    out["BlockSize"] = float(re.search(r"BlockSize\s*=\s*([\d.]+)f", cfg).group(1))
    out["MaxBlockCount"] = eval(re.search(r"MaxBlockCount\s*=\s*([\d *]+);", cfg).group(1))
    out["MaxChunkCount"] = eval(re.search(r"MaxChunkCount\s*=\s*([\d *]+);", cfg).group(1))
    m = re.search(r"BlockCenterLocation - dvec3\{([\d.]+), ([\d.]+), ([\d.]+)\}\) - ([\d.]+);", gen)
    out["sphere"] = [float(m.group(i)) for i in range(1, 5)]
    m = re.search(r"tvec3<T, glm::defaultp>\(([\d.]+), ([\d.]+), ([\d.]+)\)\)\) \* static_cast<T>\(([\d.]+)\)", mh)
    out["hash"] = [float(m.group(i)) for i in range(1, 5)]
    m = re.search(r"\(BlockCenterLocation\.y \* (\.\d+) \+ \(displacement\(BlockCenterLocation \* (\.\d+)\)\) \* ([\d.]+)\) \* (\.\d+)", gen)
    out["terrain"] = [float(m.group(i)) for i in range(1, 5)]
    out["CameraFOV"] = float(re.search(r"CameraFOV = ([\d.]+)f", inst).group(1))
    out["CameraNear"] = float(re.search(r"CameraNear = ([\d.]+)f", inst).group(1))
    out["CameraFar"] = float(re.search(r"CameraFar = ([\d.]+)f", inst).group(1))
    out["Windows"] = [int(re.search(r"WindowsWidth = (\d+)", inst).group(1)), int(re.search(r"WindowsHeight = (\d+)", inst).group(1))]
    m = re.search(r"CameraPositioner_FirstPerson\(vec3\(([\d.]+)f, ([\d.]+)f, ([\d.]+)f\), vec3\(([\d.]+)f, ([\d.]+)f, ([\d.]+)f\), vec3\(([\d.]+)f, ([\d.]+)f, ([\d.]+)f\)\)", cam)
    out["camera_start"] = [float(m.group(i)) for i in range(1, 10)]
    b26 = occ[occ.index("Get26Offsets"):occ.index("Get6Offsets")]
    out["offsets26"] = [[int(a), int(b), int(c)] for a, b, c in re.findall(r"\{(-?\d), (-?\d), (-?\d)\}", b26)]
    b6 = occ[occ.index("Get6Offsets"):occ.index("ErodeSingleVoxel")]
    out["offsets6"] = [[int(a), int(b), int(c)] for a, b, c in re.findall(r"\{(-?\d), (-?\d), (-?\d)\}", b6)]
    assert len(out["offsets26"]) == 26 and len(out["offsets6"]) == 6
    # K6: resident-set selection constants (VoxelSceneConfig.h:33-41, ChunkManagerHelper.h:26-44)
    sc = {}
    for name in ("BakeVisibilityViewNum", "ViewForwardLoadChunkSize", "ViewBackwardLoadChunkSize", "MaxChunkCheckTimes",
                 "MaxUnsyncedLoadChunkCount", "ChunkTaskPerCore"):
        sc[name] = int(re.search(name + r"\s*=\s*(\d+)", cfg).group(1))
    sc["ViewChunkAngle"] = float(re.search(r"ViewChunkAngle\s*=\s*([\d.]+)f", cfg).group(1))
    sc["ChunkOverrideMode"] = re.search(r"ChunkOverrideMode\s*=\s*EChunkOverrideMode::(\w+)", cfg).group(1)
    out["scene_config"] = sc
    mh2 = read("Runtimes/Voxel/Chunk/ChunkManagerHelper.h")
    imp = mh2[mh2.index("CalculateChunkImportance"):mh2.index("CalculateBlockImportance")]
    out["importance"] = {
        "Far": float(re.search(r"Far = ([\d.]+)f", imp).group(1)),
        "Near": float(re.search(r"Importance = ([\d.]+e\d+)f", imp).group(1)),
        "near_cube": int(re.search(r"CurrentOffset\.x >= -(\d)", imp).group(1)),
        "angle_term": [float(x) for x in re.search(r"CameraForwardVector\)\) - ([\d.]+)f\) \* ([\d.]+)f, ([\d.]+)f\)", imp).groups()],
        "distance_floor": float(re.search(r"std::max\(([\d.]+)f, Far - Distance\)", imp).group(1)),
    }
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("triplanar_faces.json", triplanar), ("constants.json", constants)):
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(fn(), f, indent=1, sort_keys=True)
        print("wrote", os.path.join(OUT, name))
