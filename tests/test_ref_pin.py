"""The oracle pinned by the REFERENCE'S OWN CODE.

oracle/_ref/libmeso_ref.so is the reference's hot-path headers (GeneratorHelper.h, VoxelMathHelper.h, Chunk.h,
BinaryOccupancyVolume.h, ChunkManagerHelper.h, NearestMap.h, Comparator.h, TriplePlanarCube.h, GPUStructures.h,
VoxelSceneConfig.h) compiled unmodified from /root/reference against stand-ins for the un-vendored third-party headers
(oracle/ref_shim/, oracle/ref_driver.cpp).  tests/golden/ref_build.npz holds what that build answered on the battery
in tests/refprobe.py (tools/gen_golden_from_ref_build.py).  Here:

  * the repo oracle must answer the same battery bit for bit (runs anywhere; this is the pin);
  * where the reference build is present, it must still reproduce the committed file (the file is not stale);
  * whole-grid K1/K2 outputs in the layout the GPU produces are compared with the reference's rows chunk by chunk
    (the same checker the GPU test uses, fed by the oracle here).
"""
import os

import numpy as np
import pytest

import refprobe

GOLD = dict(np.load(refprobe.GOLDEN))


def _diff(a, b):
    bad = []
    for k in sorted(set(a) | set(b)):
        if k not in a or k not in b:
            bad.append(k + " (missing)")
            continue
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape or x.dtype != y.dtype or x.tobytes() != y.tobytes():
            bad.append(k)
    return bad


@pytest.fixture(scope="module")
def orc_answers(orc):
    return refprobe.probe(refprobe.OrcBackend())


def test_oracle_equals_reference_build_bit_for_bit(orc_answers):
    assert _diff(orc_answers, GOLD) == []


def test_reference_build_reproduces_committed_vectors():
    so = refprobe.build_ref()
    if so is None:
        pytest.skip("no /root/reference and no prebuilt oracle/_ref/libmeso_ref.so here; the committed vectors stand in")
    assert _diff(refprobe.probe(refprobe.RefBackend(so)), GOLD) == []


def test_hand_derived_known_answers_hold_in_the_reference_build():
    """SURVEY.md section 4's de-facto KATs, now answered by the reference's own GenerateSphere + cull."""
    assert int(GOLD["sphere_counts"].sum()) == 523155
    assert int(GOLD["sphere_instances"].sum()) == 201936
    assert GOLD["layouts"].tolist()[:12] == [12, 0, 4, 8, 16, 0, 12, 160, 64, 128, 144, 16]
    assert GOLD["triplanar_indices"].tolist() == [0, 1, 2, 3, 4, 5, 6, 1]
    assert int(GOLD["view0_mode0_count"][0]) == 19661
    # terrain rows are not trivial: full, empty and mixed chunks all occur
    tc = GOLD["terrain_counts"]
    assert (tc == 4096).any() and (tc == 0).any() and ((tc > 0) & (tc < 4096)).sum() >= 20


def test_probe_covers_edge_inputs():
    """empty block list, solid chunk, camera's own chunk (max(0, NaN)), non-normalised and axis-aligned views."""
    assert GOLD["erode_mips"].shape == (9, 4, 512)
    assert not GOLD["erode_mips"][7].any()                      # the empty list
    solid = np.unpackbits(GOLD["erode_mips"][3], axis=1, bitorder="little")
    assert solid[0].sum() == 4096 and solid[1].sum() == 14 ** 3 and solid[2].sum() == 12 ** 3 and solid[3].sum() == 10 ** 3
    assert (GOLD["chunk_importance_loc"][7] == GOLD["chunk_importance_cam"][7]).all()
    assert GOLD["chunk_importance"][7] == np.float32(1.0e6)      # own chunk sits inside the +-2 cube
    assert np.isfinite(GOLD["chunk_importance"]).all()


@pytest.mark.parametrize("sin_mode", ["libm", "portable"])
def test_whole_grid_outputs_match_reference_rows(orc, sin_mode):
    """The layout K1/K2 produce (block masks, mips 1..3, chunk table, instance list), from the oracle, against the
    reference's per-chunk rows.  `portable` is the sin the GPU uses: on these chunks it decides every block as libm does."""
    mode = orc.SIN_LIBM if sin_mode == "libm" else orc.SIN_PORTABLE
    origin, dims = (2, -4, -4), (8, 8, 8)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, orc.REF_SPHERE, granularity=orc.GRAN_BLOCK, sin_mode=mode)
    table, mips, inst = vol.build_occupancy(stamp=9)
    n = refprobe.check_grid_against_golden(GOLD, "sphere", refprobe.SPHERE_CHUNKS, origin, dims, vol.occ(), mips, table, inst, 9)
    assert n == 448 and len(inst) == 201936

    origin, dims = refprobe.TERRAIN_GRID
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_TERRAIN, None, granularity=orc.GRAN_BLOCK, sin_mode=mode)
    table, mips, inst = vol.build_occupancy(stamp=5)
    assert refprobe.check_grid_against_golden(GOLD, "terrain", refprobe.TERRAIN_CHUNKS, origin, dims, vol.occ(), mips, table, inst, 5) == 48


# ---- a13/a14: the reference's own vertex + fragment shader text, executed ---------------------------------------------
DRAW = dict(np.load(refprobe.DRAW_GOLDEN))


def test_reference_shader_build_reproduces_committed_frames(orc):
    so = refprobe.build_ref()
    if so is None:
        pytest.skip("no reference build here; the committed frames stand in")
    assert _diff(refprobe.probe_draw(refprobe.RefBackend(so), orc), DRAW) == []


def test_shader_parks_invalid_instances():
    """ChunkIndex == INT_MAX, or a stale frame stamp: gl_Position = (0,0,-1,1) and nothing else written (SimpleVoxel.cpp:147-163)."""
    for row in DRAW["vs_invalid_instance"]:
        assert row.tolist() == [0.0, 0.0, -1.0, 1.0] + [0.0] * 6


@pytest.mark.parametrize("eye_idx", refprobe.DRAW_EYES)
def test_restated_draw_and_dda_match_the_executed_reference_shaders(orc, eye_idx):
    """Per pixel, three answers: (R) the reference's VS/FS text run through a fan rasteriser with depth Greater,
    (O) the oracle's closed-form restatement of that draw (orc_ref_instanced_pixel), (D) the DDA the CUDA kernel is
    bit-identical to.  Same hit / miss, same block, same face, colour within 1 LSB, and (R) depth == the projection of
    (O)'s hit within 1e-5 relative.  Pixels whose ray passes within 1e-3 block units of a face edge are the
    rasteriser's tie-break territory and are skipped (a few per cent)."""
    w, h = refprobe.DRAW_W, refprobe.DRAW_H
    origin, dims, vol, table, inst, cams = refprobe.draw_scene(orc)
    cam = cams[refprobe.DRAW_EYES.index(eye_idx)]
    r_inst, r_col, r_nrm, r_depth = (DRAW[f"eye{eye_idx}_{k}"] for k in ("instance", "color", "normal", "depth"))
    rec = vol.raymarch(orc.ray_setup(cam, origin, w, h), w, h, shadow=False)
    u = orc.unpack_records(rec)
    scene = orc.default_scene_config()
    cam_chunk = cam["CameraChunkLocation"][0][:3].astype(np.int64)
    P = cam["Projection"][0].reshape(4, 4).astype(np.float64)        # [col][row]
    checked = skipped = 0
    for py in range(h):
        for px in range(w):
            hit, blk, face, t, margin, rgba = orc.ref_instanced_pixel(cam, scene, table, inst, w, h, px, py)
            if hit and margin < 1e-3:
                skipped += 1
                continue
            k = int(r_inst[py, px])
            if not hit:
                # a miss may still sit within a hair of a silhouette edge; the DDA must agree with the restatement,
                # the rasteriser is only required to agree when it also misses or the DDA is the odd one out nowhere
                assert not u["hit"][py, px], (px, py)
                if k >= 0:
                    skipped += 1
                else:
                    assert r_col[py, px].tolist() == [0.0, 0.0, 0.0, 1.0] and r_depth[py, px] == 0.0
                continue
            assert k >= 0, (px, py)
            checked += 1
            b = inst[k]
            r_blk = (table[b["ChunkIndex"]]["ChunkLocation"].astype(np.int64) - cam_chunk) * 16 + b["BlockLocation"][:3].astype(np.int64)
            assert np.array_equal(r_blk, blk), (px, py, r_blk, blk)                                # (R) block == (O) block
            n = r_nrm[py, px]
            ax = face >> 1
            assert abs(float(n[ax]) - (0.5 if face & 1 else -0.5)) < 1e-6, (px, py, n, face)       # (R) face == (O) face
            assert all(abs(float(n[c])) < 0.5 - 1e-4 for c in range(3) if c != ax)
            r8 = [int(np.floor(float(r_col[py, px, c]) * 255.0 + 0.5)) for c in range(4)]
            o8 = [int(np.floor(float(rgba[c]) * 255.0 + 0.5)) for c in range(4)]
            assert all(abs(a - e) <= 1 for a, e in zip(r8, o8)), (px, py, r8, o8)                  # (R) colour ~ (O) colour
            # (R) depth: reverse-Z z/w of a point at view depth t  ->  (P[2][2]*(-t) + P[3][2]) / t
            assert float(r_depth[py, px]) == pytest.approx((P[2][2] * (-t) + P[3][2]) / t, rel=1e-5)
            # (D) the DDA record
            assert u["hit"][py, px]
            vox = np.array([u["x"][py, px], u["y"][py, px], u["z"][py, px]], dtype=np.int64)
            assert np.array_equal((vox >> 3) + (np.array(origin) - cam_chunk) * 16, r_blk), (px, py)
            assert int(u["face"][py, px]) == face
            d8 = [(int(rec["rgba"][py, px]) >> (8 * c)) & 0xFF for c in range(4)]
            assert all(abs(a - e) <= 1 for a, e in zip(d8, r8)), (px, py, d8, r8)                  # (D) colour ~ (R) colour
    assert checked > 15000 and skipped < 0.03 * w * h, (checked, skipped)


@pytest.mark.parametrize("eye_idx", refprobe.DRAW_EYES)
def test_dda_frame_against_reference_draw_whole_frame(orc, eye_idx):
    """The checker tests/test_zz_gpu_ref_pin.py applies to the CUDA kernel's frame, fed here with the oracle's DDA frame."""
    origin, dims, vol, table, inst, cams = refprobe.draw_scene(orc)
    cam = np.frombuffer(DRAW[f"eye{eye_idx}_camera"].tobytes(), dtype=orc.Camera)
    assert cam.tobytes() == cams[refprobe.DRAW_EYES.index(eye_idx)].tobytes()
    rec = vol.raymarch(orc.ray_setup(cam, origin, refprobe.DRAW_W, refprobe.DRAW_H), refprobe.DRAW_W, refprobe.DRAW_H, shadow=False)
    hits, misses, skipped = refprobe.check_records_against_ref_draw(DRAW, eye_idx, rec, origin)
    assert hits > 2000 and misses > 1500 and skipped < 0.15 * rec.size, (hits, misses, skipped)


# ---- the product's pure-host ABI functions against the reference build (no device needed) ------------------------------
def test_product_baked_direction_equals_reference_nearest_map():
    """meso_baked_direction (libmeso_b200.so, host code) == TNearestMap::Query over GetFibonacciSphere<float>(256), the
    table lookup of ChunkManager.h:106-124, on the 400 queries the reference build answered."""
    from mesoengine_b200 import capi
    dirs = GOLD["fibonacci_f32_256"]
    for q, want in zip(GOLD["nearest_direction_query"], GOLD["nearest_direction"]):
        d, idx = capi.baked_direction(256, q)
        assert idx == int(want)
        assert d.view(np.uint32).tolist() == dirs[idx].view(np.uint32).tolist()


def test_host_mirror_declarations_equal_the_reference_headers():
    """mesoengine_b200/host/MesoHost.h compiled side by side with the reference's VoxelSceneConfig.h / VoxelMathHelper.h:
    every FVoxelSceneConfig field (type, default, offset), EChunkOverrideMode, ConvertToChunkLocation bits."""
    import subprocess
    if refprobe.build_ref() is None or not os.path.isdir(refprobe.REF_ROOT):
        pytest.skip("needs /root/reference (compares against the reference's own headers)")
    exe = os.path.join(os.path.dirname(refprobe.REF_SO), "host_mirror_check")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.startswith("ok:"), out.stdout + out.stderr


def test_host_objects_equal_the_reference_objects():
    """FChunk / FBinaryOccupancyVolume / FOccupancyHelper / FGeneratorHelper::GenerateSphere / GeneratorType of the host
    mirror next to the reference's own (oracle/ref_host_objects_check.cpp): block lists in order, four mips bit for bit,
    the cull test incl. out-of-chunk locations, clamped and boundary reads, both erosion stencils."""
    import subprocess
    if refprobe.build_ref() is None or not os.path.isdir(refprobe.REF_ROOT):
        pytest.skip("needs /root/reference (compares against the reference's own headers)")
    exe = os.path.join(os.path.dirname(refprobe.REF_SO), "host_objects_check")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok:") and "523155 blocks" in out.stdout, out.stdout + out.stderr
