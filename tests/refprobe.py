"""One probe, two backends: the same battery of inputs is pushed through
  * RefBackend -- the REFERENCE'S OWN headers compiled into oracle/_ref/libmeso_ref.so (oracle/ref_driver.cpp), and
  * OrcBackend -- the repo's CPU oracle (oracle/libmeso_oracle.so via tests/orc.py),
and both produce the same dictionary of numpy arrays.  tools/gen_golden_from_ref_build.py stores RefBackend's
dictionary as tests/golden/ref_build.npz; tests/test_ref_pin.py checks OrcBackend against it (anywhere) and RefBackend
against it (where the reference build exists).  TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"
REF_SO = os.path.join(_ROOT, "oracle", "_ref", "libmeso_ref.so")
GOLDEN = os.path.join(_ROOT, "tests", "golden", "ref_build.npz")

Candidate = np.dtype([("Importance", "<f4"), ("Offset", "<i4", (3,))])
Camera160 = np.dtype([("Projection", "<f4", (16,)), ("View", "<f4", (16,)), ("CameraChunkLocation", "<i4", (4,)),
                      ("SubCameraLocation", "<f4", (4,))])      # FGPUUniformCamera, GPUStructures.h:36-41


def build_ref():
    """Compile the reference headers where they lie (only possible where /root/reference exists). Returns the .so path
    or None."""
    if os.path.isdir(os.path.join(REF_ROOT, "Runtimes")):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


class RefBackend:
    name = "reference build"

    def __init__(self, so=None):
        self.lib = C.CDLL(so or REF_SO)
        L = self.lib
        L.ref_hash3.restype = C.c_double
        L.ref_hash3.argtypes = [C.c_double] * 3
        L.ref_displacement.restype = C.c_double
        L.ref_chunk_importance.restype = C.c_float
        L.ref_block_importance.restype = C.c_float
        L.ref_select_view_chunks.restype = C.c_int64
        L.ref_bake_and_query.restype = C.c_int64
        L.ref_nearest_direction.restype = C.c_uint32
        L.ref_truncate_frame_stamp.restype = C.c_uint32
        L.ref_truncate_frame_stamp.argtypes = [C.c_uint64]

    def hash3(self, x, y, z):
        return self.lib.ref_hash3(float(x), float(y), float(z))

    def noised(self, x):
        xin = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(4)
        self.lib.ref_noised(_p(xin), _p(out))
        return out

    def displacement(self, p):
        pin = np.ascontiguousarray(p, dtype=np.float64)
        return self.lib.ref_displacement(_p(pin))

    def generate_chunk(self, kind, loc):
        """-> blocks (n,3) u8 in generator order, mips (4,4096) u8 at x+16y+256z, cull flags (n,) u8 (threshold 1)."""
        loc = np.ascontiguousarray(loc, dtype=np.int32)
        xyz = np.zeros((4096, 3), dtype=np.uint8)
        mips = np.zeros((4, 4096), dtype=np.uint8)
        cull = np.zeros(4096, dtype=np.uint8)
        n = self.lib.ref_generate_chunk(C.c_int(kind), _p(loc), C.c_float(1.0), C.c_int(16), C.c_int(4), C.c_int(1),
                                        _p(xyz), _p(mips), _p(cull))
        return xyz[:n].copy(), mips, cull[:n].copy()

    def erode_blocks(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.uint8).reshape(-1, 3)
        mips = np.zeros((4, 4096), dtype=np.uint8)
        cull = np.zeros(max(len(xyz), 1), dtype=np.uint8)
        self.lib.ref_erode_blocks(_p(xyz), C.c_int(len(xyz)), C.c_int(16), C.c_int(4), C.c_int(1), _p(mips), _p(cull))
        return mips, cull[:len(xyz)].copy()

    def erode_offsets(self, use26):
        out = np.zeros((26, 3), dtype=np.int32)
        n = self.lib.ref_erode_offsets(C.c_int(use26), _p(out))
        return out[:n].copy()

    def index_helpers(self, loc, res):
        out = np.zeros(4, dtype=np.uint32)
        self.lib.ref_index_helpers(_p(np.ascontiguousarray(loc, dtype=np.int32)), _p(np.ascontiguousarray(res, dtype=np.int32)), _p(out))
        return out

    def convert_to_chunk_location(self, pos, chunk_size):
        f = np.zeros(3, dtype=np.float32)
        c = np.zeros(3, dtype=np.int32)
        self.lib.ref_convert_to_chunk_location(_p(np.ascontiguousarray(pos, dtype=np.float32)), C.c_float(chunk_size), _p(f), _p(c))
        return f, c

    def fibonacci_f32(self, n):
        out = np.zeros((n, 3), dtype=np.float32)
        self.lib.ref_fibonacci_sphere_f32(C.c_uint32(n), C.c_int(1), _p(out))
        return out

    def fibonacci_f64(self, n, normalize):
        out = np.zeros((n, 3), dtype=np.float64)
        self.lib.ref_fibonacci_sphere_f64(C.c_uint32(n), C.c_int(1 if normalize else 0), _p(out))
        return out

    def chunk_importance(self, cam, fwd, loc):
        return self.lib.ref_chunk_importance(_p(np.ascontiguousarray(cam, dtype=np.int32)), _p(np.ascontiguousarray(fwd, dtype=np.float32)),
                                             _p(np.ascontiguousarray(loc, dtype=np.int32)))

    def block_importance(self, cam, fwd, chunk, block):
        return self.lib.ref_block_importance(_p(np.ascontiguousarray(cam, dtype=np.int32)), _p(np.ascontiguousarray(fwd, dtype=np.float32)),
                                             _p(np.ascontiguousarray(chunk, dtype=np.int32)), _p(np.ascontiguousarray(block, dtype=np.uint8)),
                                             C.c_uint32(16))

    def select_view(self, fwd, forward, backward, angle, mode):
        side = 2 * forward + 1
        out = np.zeros(side ** 3, dtype=Candidate)
        n = self.lib.ref_select_view_chunks(_p(np.ascontiguousarray(fwd, dtype=np.float32)), C.c_uint32(forward), C.c_uint32(backward),
                                            C.c_float(angle), C.c_int(mode), _p(out), C.c_int64(len(out)))
        return out[:n].copy()

    def nearest_direction(self, dirs, q):
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        return int(self.lib.ref_nearest_direction(_p(d), C.c_uint32(len(d)), _p(np.ascontiguousarray(q, dtype=np.float32))))

    def bake_and_query(self, samples, forward, backward, angle, q):
        first = np.zeros(1, dtype=Candidate)
        n = self.lib.ref_bake_and_query(C.c_uint32(samples), C.c_uint32(forward), C.c_uint32(backward), C.c_float(angle),
                                        _p(np.ascontiguousarray(q, dtype=np.float32)), _p(first))
        return int(n), first

    def truncate_frame_stamp(self, s):
        return int(self.lib.ref_truncate_frame_stamp(int(s)))

    def sort_ivec3(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        self.lib.ref_sort_ivec3(_p(a), C.c_int64(len(a)))
        return a

    def triplanar_indices(self):
        out = np.zeros(16, dtype=np.uint16)
        n = self.lib.ref_triplanar_indices(_p(out), C.c_int(16))
        return out[:n].copy()

    def layouts(self):
        out = np.zeros(16, dtype=np.uint32)
        self.lib.ref_layouts(_p(out))
        return out

    def scene_config_defaults(self):
        out = np.zeros(14, dtype=np.float64)
        self.lib.ref_scene_config_defaults(_p(out))
        return out


class OrcBackend:
    """The repo oracle behind the same method names.  Reference-exact settings: libm sin, the reference sphere."""
    name = "repo oracle"

    def __init__(self):
        import orc
        self.orc = orc

    def hash3(self, x, y, z):
        return self.orc.hash3(x, y, z, self.orc.SIN_LIBM)

    def noised(self, x):
        return self.orc.noised(x, self.orc.SIN_LIBM)

    def displacement(self, p):
        return self.orc.displacement(p, self.orc.SIN_LIBM)

    @staticmethod
    def _mips_bytes(mips_words):
        """(depth, 64) u64 words, bit x+16y+256z -> (depth, 4096) bytes at the same index."""
        b = np.unpackbits(np.ascontiguousarray(mips_words).view(np.uint8), bitorder="little")
        return b.reshape(mips_words.shape[0], 4096)

    def generate_chunk(self, kind, loc):
        o = self.orc
        xyz = o.generate_chunk(kind, o.REF_SPHERE if kind == o.SDF_SPHERE else None, loc, sin_mode=o.SIN_LIBM)
        mips = o.erode_mips(xyz, depth=4)
        inst = o.emit_instances(xyz, mips, threshold=1)
        return xyz, self._mips_bytes(mips), self._cull_from_instances(xyz, inst)

    @staticmethod
    def _cull_from_instances(xyz, inst):
        """cull flag per block = the block did not come out of orc_emit_instances (order-preserving match)."""
        cull = np.ones(len(xyz), dtype=np.uint8)
        j = 0
        for i in range(len(xyz)):
            if j < len(inst) and tuple(inst["BlockLocation"][j][:3]) == tuple(xyz[i]):
                cull[i] = 0
                j += 1
        assert j == len(inst), "instances are not an ordered subsequence of the block list"
        return cull

    def erode_blocks(self, xyz):
        o = self.orc
        xyz = np.ascontiguousarray(xyz, dtype=np.uint8).reshape(-1, 3)
        mips = o.erode_mips(xyz, depth=4)
        inst = o.emit_instances(xyz, mips, threshold=1) if len(xyz) else np.zeros(0, dtype=o.GPUBlock)
        return self._mips_bytes(mips), self._cull_from_instances(xyz, inst)

    def erode_offsets(self, use26):
        import json
        g = json.load(open(os.path.join(_ROOT, "tests", "golden", "constants.json")))
        return np.asarray(g["offsets26" if use26 else "offsets6"], dtype=np.int32)

    def index_helpers(self, loc, res):
        # bit index used by every oracle mask (orc_internal.h orc_bidx) and the bound tests of orc_occupancy.c, restated
        x, y, z = (int(v) for v in loc)
        rx, ry, rz = (int(v) for v in res)
        cx, cy, cz = (max(min(v, r - 1), 0) for v, r in ((x, rx), (y, ry), (z, rz)))
        oob = x < 0 or x >= rx or y < 0 or y >= ry or z < 0 or z >= rz
        oobt = x < 1 or x >= rx - 1 or y < 1 or y >= ry - 1 or z < 1 or z >= rz - 1
        return np.array([(x + y * rx + z * rx * ry) & 0xFFFFFFFF, cx + cy * rx + cz * rx * ry, oob, oobt], dtype=np.uint32)

    def convert_to_chunk_location(self, pos, chunk_size):
        f = np.zeros(3, dtype=np.float32)
        c = np.zeros(3, dtype=np.int32)
        self.orc.lib.orc_convert_to_chunk_location(_p(np.ascontiguousarray(pos, dtype=np.float32)), C.c_float(chunk_size), _p(f), _p(c))
        return f, c

    def fibonacci_f32(self, n):
        return self.orc.fibonacci_sphere_f32(n)

    def fibonacci_f64(self, n, normalize):
        return self.orc.fibonacci_sphere(n, normalize)

    def chunk_importance(self, cam, fwd, loc):
        return self.orc.chunk_importance(cam, fwd, loc)

    def block_importance(self, cam, fwd, chunk, block):
        return self.orc.block_importance(cam, fwd, chunk, block, 16)

    def select_view(self, fwd, forward, backward, angle, mode):
        return self.orc.select_view_chunks(fwd, forward, backward, angle, mode)

    def nearest_direction(self, dirs, q):
        return self.orc.nearest_direction(dirs, q)

    def bake_and_query(self, samples, forward, backward, angle, q):
        dirs = self.orc.fibonacci_sphere_f32(samples)
        d = dirs[self.orc.nearest_direction(dirs, q)]
        sel = self.orc.select_view_chunks(d, forward, backward, angle, 0)
        return len(sel), sel[:1]

    def truncate_frame_stamp(self, s):
        return int(s) % (1 << 32)       # the stamp the kernels write is the caller's u32 (include/meso_cuda.h)

    def sort_ivec3(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]

    def triplanar_indices(self):
        import json
        g = json.load(open(os.path.join(_ROOT, "tests", "golden", "triplanar_faces.json")))
        return np.asarray(g["fan_indices"], dtype=np.uint16)

    def layouts(self):
        o = self.orc
        blk, chk, cam = o.GPUBlock, o.GPUChunk, o.Camera
        return np.array([blk.itemsize, blk.fields["ChunkIndex"][1], blk.fields["BlockLocation"][1], blk.fields["BlockFrameStamp"][1],
                         chk.itemsize, chk.fields["ChunkLocation"][1], chk.fields["ChunkFrameStamp"][1],
                         cam.itemsize, cam.fields["View"][1], cam.fields["CameraChunkLocation"][1], cam.fields["SubCameraLocation"][1],
                         o.SceneConfig.itemsize, 0x7FFFFFFF, 0x7FFFFFFF, 255, 0], dtype=np.uint32)

    def scene_config_defaults(self):
        import json
        g = json.load(open(os.path.join(_ROOT, "tests", "golden", "constants.json")))   # parsed from the reference text
        c = g["scene_config"]
        return np.array([g["BlockResolution"], g["BlockSize"], g["ChunkResolution"], g["MaxBlockCount"], g["MaxChunkCount"],
                         c["BakeVisibilityViewNum"], c["ViewForwardLoadChunkSize"], c["ViewBackwardLoadChunkSize"],
                         c["MaxUnsyncedLoadChunkCount"], c["ViewChunkAngle"], g["ChunkOccupancyDepth"],
                         g["ChunkInnerVoxelCullDepthThreshold"], g["ChunkResolution"] * g["BlockSize"],
                         c["MaxChunkCheckTimes"]], dtype=np.float64)


# ---- the battery --------------------------------------------------------------------------------------------------
SPHERE_CHUNKS = [(x, y, z) for x in range(3, 10) for y in range(-4, 4) for z in range(-4, 4)]       # covers the whole sphere
TERRAIN_CHUNKS = [(x, y, z) for x in (-3, 0, 2, 40) for y in (-2, -1, 0, 1) for z in (-2, 0, 5)]     # 48 chunks around y = 0
TERRAIN_GRID = ((-3, -2, -2), (44, 4, 8))      # the smallest chunk window holding TERRAIN_CHUNKS
VIEWS = [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (-0.3, 0.8, 0.52), (0.57735026, 0.57735026, 0.57735026), (0.0, -2.0, 0.0)]


def _chunk_rows(b, kind, chunks):
    counts, inst, h_blocks, h_mips, h_cull = [], [], [], [], []
    for loc in chunks:
        xyz, mips, cull = b.generate_chunk(kind, loc)
        counts.append(len(xyz))
        inst.append(int(len(xyz) - cull.sum()))
        h_blocks.append(_sha(xyz))
        h_mips.append(_sha(np.asarray(mips, dtype=np.uint8)))
        h_cull.append(_sha(cull))
    return (np.array(counts, dtype=np.int32), np.array(inst, dtype=np.int32), np.stack(h_blocks), np.stack(h_mips), np.stack(h_cull))


def probe(b):
    """-> {name: array}.  Deterministic (fixed seeds); every value is compared bit for bit."""
    r = {}
    rng = np.random.default_rng(20241017)
    pts = np.concatenate([rng.uniform(-60, 60, (600, 3)), rng.integers(-2000, 2000, (600, 3)).astype(np.float64),
                          rng.uniform(-1e4, 1e4, (300, 3))])
    r["hash3_in"] = pts
    r["hash3"] = np.array([b.hash3(*p) for p in pts])
    npts = rng.uniform(-40, 40, (400, 3))
    r["noised"] = np.stack([b.noised(p) for p in npts])
    dpts = np.concatenate([rng.uniform(-30, 30, (300, 3)), (rng.integers(-400, 400, (200, 3)) * 0.1)])
    r["displacement"] = np.array([b.displacement(p) for p in dpts])

    (r["sphere_counts"], r["sphere_instances"], r["sphere_blocks_sha1"], r["sphere_mips_sha1"], r["sphere_cull_sha1"]) = \
        _chunk_rows(b, 0, SPHERE_CHUNKS)
    (r["terrain_counts"], r["terrain_instances"], r["terrain_blocks_sha1"], r["terrain_mips_sha1"], r["terrain_cull_sha1"]) = \
        _chunk_rows(b, 1, TERRAIN_CHUNKS)
    # one chunk of each kind in full (not only hashed), so a mismatch can be looked at
    xyz, mips, cull = b.generate_chunk(0, (4, -1, 0))
    r["sphere_chunk_4_-1_0_blocks"], r["sphere_chunk_4_-1_0_mips"], r["sphere_chunk_4_-1_0_cull"] = xyz, np.packbits(mips, axis=1, bitorder="little"), cull
    xyz, mips, cull = b.generate_chunk(1, (0, -1, 0))
    r["terrain_chunk_0_-1_0_blocks"], r["terrain_chunk_0_-1_0_mips"], r["terrain_chunk_0_-1_0_cull"] = xyz, np.packbits(mips, axis=1, bitorder="little"), cull

    # erosion on arbitrary block lists: densities, a solid chunk, a solid chunk with single holes, border slabs, empty
    lists = []
    grid = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3).astype(np.uint8)
    for dens in (0.5, 0.9, 0.99):
        lists.append(grid[rng.random(4096) < dens])
    lists.append(grid)
    holes = np.ones(4096, dtype=bool)
    holes[rng.integers(0, 4096, 12)] = False
    lists.append(grid[holes])
    lists.append(grid[(grid[:, 0] < 3)])
    lists.append(grid[(grid[:, 2] > 11) | (grid[:, 1] == 0)])
    lists.append(grid[:0])
    shuffled = grid[rng.permutation(4096)][:3000]           # generator order is not assumed by the erosion
    lists.append(shuffled)
    em, ec = [], []
    for li in lists:
        mips, cull = b.erode_blocks(li)
        em.append(np.packbits(np.asarray(mips, dtype=np.uint8), axis=1, bitorder="little"))
        ec.append(_sha(cull))
    r["erode_mips"] = np.stack(em)
    r["erode_cull_sha1"] = np.stack(ec)
    r["erode_offsets_26"] = b.erode_offsets(1)
    r["erode_offsets_6"] = b.erode_offsets(0)

    locs = [(x, y, z) for x in (-1, 0, 1, 7, 14, 15, 16) for y in (-1, 0, 8, 15, 16) for z in (-2, 0, 1, 15, 17)]
    r["index_helpers"] = np.stack([b.index_helpers(l, (16, 16, 16)) for l in locs])
    r["index_helpers_ragged"] = np.stack([b.index_helpers(l, (5, 9, 3)) for l in locs])

    pos = np.concatenate([rng.uniform(-100, 100, (300, 3)), rng.integers(-8, 8, (100, 3)) * 16.0,
                          np.array([[0.0, -0.0, 16.0], [15.999999, -1e-7, -16.0], [5.0, 2.0, 2.0], [1e6, -1e6, 0.5]])]).astype(np.float32)
    fc = [b.convert_to_chunk_location(p, 16.0) for p in pos]
    r["chunk_location_fract"] = np.stack([f for f, _ in fc])
    r["chunk_location_chunk"] = np.stack([c for _, c in fc])

    r["fibonacci_f32_256"] = b.fibonacci_f32(256)
    r["fibonacci_f64_8"] = b.fibonacci_f64(8, True)
    r["fibonacci_f64_100_raw"] = b.fibonacci_f64(100, False)

    # 30 cameras x 50 chunk locations (inputs are stored too: the GPU test feeds them to meso_chunk_importance)
    cams = np.repeat(rng.integers(-50, 50, (30, 3)).astype(np.int32), 50, axis=0)
    fw30 = rng.normal(size=(30, 3)).astype(np.float32)
    fw30[::3] /= np.linalg.norm(fw30[::3], axis=1, keepdims=True)
    fw = np.repeat(fw30, 50, axis=0)
    offs = np.concatenate([rng.integers(-3, 4, (500, 3)), rng.integers(-70, 70, (1000, 3))]).astype(np.int32)
    offs = offs[rng.permutation(1500)]
    offs[7] = 0                                              # the camera's own chunk: max(0, NaN)
    r["chunk_importance_cam"], r["chunk_importance_fwd"], r["chunk_importance_loc"] = cams, fw, cams + offs
    r["chunk_importance"] = np.array([b.chunk_importance(c, f, l) for c, f, l in zip(cams, fw, cams + offs)], dtype=np.float32)
    blocks = rng.integers(0, 16, (1500, 3)).astype(np.uint8)
    r["block_importance_chunk"], r["block_importance_block"] = (cams + offs // 8).astype(np.int32), blocks    # inputs, for the GPU test
    r["block_importance"] = np.array([b.block_importance(c, f, c + o, bl) for c, f, o, bl in zip(cams, fw, offs // 8, blocks)], dtype=np.float32)

    for i, v in enumerate(VIEWS):
        for mode in (0, 1):
            if mode == 1 and i > 1:
                continue
            sel = b.select_view(v, 24, 6, 120.0, mode)
            key = np.lexsort((sel["Offset"][:, 2], sel["Offset"][:, 1], sel["Offset"][:, 0]))
            canon = sel[key]
            r[f"view{i}_mode{mode}_count"] = np.array([len(sel)], dtype=np.int64)
            r[f"view{i}_mode{mode}_set_sha1"] = _sha(canon)                                   # the set with its importances
            r[f"view{i}_mode{mode}_pop_importance_sha1"] = _sha(sel["Importance"].copy())     # the pop order's importances
            r[f"view{i}_mode{mode}_head"] = canon[:256].copy().view(np.uint8)
    small = b.select_view((0.2, -0.4, 0.9), 6, 2, 90.0, 0)       # a small set, stored whole
    r["view_small"] = small[np.lexsort((small["Offset"][:, 2], small["Offset"][:, 1], small["Offset"][:, 0]))].view(np.uint8)

    dirs = b.fibonacci_f32(256)
    qs = rng.normal(size=(400, 3)).astype(np.float32)
    r["nearest_direction_query"] = qs
    r["nearest_direction"] = np.array([b.nearest_direction(dirs, q) for q in qs], dtype=np.uint32)
    n, first = b.bake_and_query(16, 8, 3, 120.0, (0.3, 0.1, -0.9))
    r["bake_query_count"] = np.array([n], dtype=np.int64)
    r["bake_query_first_importance"] = np.asarray(first["Importance"], dtype=np.float32).reshape(1)

    r["truncate_frame_stamp"] = np.array([b.truncate_frame_stamp(s) for s in (0, 1, 0xFFFFFFFF, 0x100000000, 0x123456789AB)], dtype=np.uint64)
    tri = rng.integers(-5, 5, (200, 3)).astype(np.int32)
    r["sort_ivec3"] = b.sort_ivec3(tri)
    r["triplanar_indices"] = b.triplanar_indices()
    r["layouts"] = b.layouts()
    r["scene_config_defaults"] = b.scene_config_defaults()
    return r


# ---- whole-grid outputs (the format K1/K2 produce, on the GPU or in the oracle) against the reference build's rows ----
_GEN_ORDER = None


def _gen_order():
    """bit index x+16y+256z of every block in the reference's generator order (X outer, Z inner; GeneratorHelper.h:96-100)."""
    global _GEN_ORDER
    if _GEN_ORDER is None:
        X, Y, Z = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
        _GEN_ORDER = (X + 16 * Y + 256 * Z).reshape(-1), np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.uint8)
    return _GEN_ORDER


def check_grid_against_golden(gold, prefix, chunks, origin, dims, occ, mips123, table, inst, stamp):
    """occ: (nchunks*64,) u64 block masks; mips123: (nchunks,3,64) u64; table: GPUChunk[nchunks]; inst: GPUBlock[n]
    (ChunkIndex = slot, generator order inside a chunk, slots ascending).  Every chunk of `chunks` is compared with the
    row the reference's own GenerateSphere / TestGenerator + CalculateOccupancyErodeMipmaps +
    bShouldVoxelOccupancyCull(.., 1) produced for it: block list, four mips, surviving instances, table entry.
    Returns the number of chunks compared."""
    occ = np.ascontiguousarray(occ, dtype=np.uint64).reshape(-1, 64)
    order, xyz_all = _gen_order()
    inst_chunk = np.asarray(inst["ChunkIndex"])
    assert np.all(np.diff(inst_chunk.astype(np.int64)) >= 0), "instances are not grouped by ascending chunk slot"
    starts = np.searchsorted(inst_chunk, np.arange(occ.shape[0]), side="left")
    ends = np.searchsorted(inst_chunk, np.arange(occ.shape[0]), side="right")
    for row, loc in enumerate(chunks):
        c = [loc[i] - origin[i] for i in range(3)]
        assert all(0 <= c[i] < dims[i] for i in range(3)), "grid does not contain golden chunk %s" % (loc,)
        slot = c[0] + dims[0] * (c[1] + dims[1] * c[2])
        bits0 = np.unpackbits(occ[slot].view(np.uint8), bitorder="little")
        present = bits0[order].astype(bool)
        xyz = xyz_all[present]
        assert len(xyz) == int(gold[prefix + "_counts"][row]), "chunk %s: block count" % (loc,)
        assert np.array_equal(_sha(xyz), gold[prefix + "_blocks_sha1"][row]), "chunk %s: block list" % (loc,)
        m = np.concatenate([occ[slot][None, :], np.asarray(mips123[slot], dtype=np.uint64)], 0)
        mb = np.unpackbits(np.ascontiguousarray(m).view(np.uint8), bitorder="little").reshape(4, 4096)
        assert np.array_equal(_sha(mb), gold[prefix + "_mips_sha1"][row]), "chunk %s: erode mips" % (loc,)
        mine = inst[starts[slot]:ends[slot]]
        assert len(mine) == int(gold[prefix + "_instances"][row]), "chunk %s: instance count" % (loc,)
        # cull flags in generator order, rebuilt from the emitted instances (an ordered subsequence of the blocks)
        keys = xyz[:, 0].astype(np.int64) * 256 + xyz[:, 1].astype(np.int64) * 16 + xyz[:, 2]          # ascending in generator order
        ik = mine["BlockLocation"][:, 0].astype(np.int64) * 256 + mine["BlockLocation"][:, 1].astype(np.int64) * 16 + mine["BlockLocation"][:, 2]
        assert np.all(np.diff(ik) > 0), "chunk %s: instances not in generator order" % (loc,)
        cull = np.ones(len(xyz), dtype=np.uint8)
        pos = np.searchsorted(keys, ik)
        assert np.all(pos < len(keys)) and np.array_equal(keys[pos], ik), "chunk %s: instance of an absent block" % (loc,)
        cull[pos] = 0
        assert np.array_equal(_sha(cull), gold[prefix + "_cull_sha1"][row]), "chunk %s: hidden-block cull" % (loc,)
        if len(mine):
            assert np.all(mine["BlockLocation"][:, 3] == 255) and np.all(mine["BlockFrameStamp"] == stamp)
        t = table[slot]
        if len(xyz):
            assert tuple(t["ChunkLocation"]) == tuple(loc) and int(t["ChunkFrameStamp"]) == stamp
        else:
            assert tuple(t["ChunkLocation"]) == (0x7FFFFFFF,) * 3      # FGPUChunk's default = invalid (Chunk.h:29)
    return len(chunks)


# ---- the reference's voxel shaders, executed (oracle/ref_glsl_driver.cpp) ---------------------------------------------
DRAW_GOLDEN = os.path.join(_ROOT, "tests", "golden", "ref_draw.npz")
DRAW_W, DRAW_H, DRAW_EYES, DRAW_STAMP = 256, 144, (0, 2, 5), 5


def draw_scene(orc):
    """The scene of the draw comparison: the block-granular 256^3 V-sphere after the hidden-block cull (FGPUChunk table +
    FGPUBlock instances from the oracle, themselves pinned by the reference build above) and three orbit cameras."""
    import scenes
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_BLOCK)
    table, _, inst = vol.build_occupancy(stamp=DRAW_STAMP)
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cams = [orc.camera_uniform(eyes[i], ctr, width=DRAW_W, height=DRAW_H) for i in DRAW_EYES]
    return origin, dims, vol, table, inst, cams


def ref_draw(lib, cam, scene_cfg, table, inst, indices, w, h):
    """cmdDrawIndexed through the reference's own VS/FS text.  -> depth, instance, colour, normal, n_behind"""
    depth = np.zeros((h, w), dtype=np.float32)
    instance = np.zeros((h, w), dtype=np.int32)
    color = np.zeros((h, w, 4), dtype=np.float32)
    normal = np.zeros((h, w, 3), dtype=np.float32)
    idx = np.ascontiguousarray(indices, dtype=np.uint16)
    lib.ref_draw_instanced.restype = C.c_int64
    behind = lib.ref_draw_instanced(_p(cam), _p(scene_cfg), _p(np.ascontiguousarray(table)), _p(np.ascontiguousarray(inst)),
                                    C.c_int64(len(inst)), _p(idx), C.c_int(len(idx)), C.c_int(w), C.c_int(h),
                                    _p(depth), _p(instance), _p(color), _p(normal))
    return depth, instance, color, normal, int(behind)


def probe_draw(backend, orc):
    """-> {name: array}: per camera, what the reference's shaders + the fixed-function rules put into every pixel."""
    _, _, _, table, inst, cams = draw_scene(orc)
    scene_cfg = orc.default_scene_config()
    idx = backend.triplanar_indices()
    out = {}
    for i, cam in zip(DRAW_EYES, cams):
        depth, instance, color, normal, behind = ref_draw(backend.lib, cam, scene_cfg, table, inst, idx, DRAW_W, DRAW_H)
        assert behind == 0
        out[f"eye{i}_depth"], out[f"eye{i}_instance"], out[f"eye{i}_color"], out[f"eye{i}_normal"] = depth, instance, color, normal
        # what the draw was given, so that a consumer needs nothing but this file: the camera block and, per covered
        # pixel, the winning instance's ViewChunkRelativeBlockOffset (SimpleVoxel.cpp:177) from its FGPUBlock / FGPUChunk
        out[f"eye{i}_camera"] = np.frombuffer(cam.tobytes(), dtype=np.uint8).copy()
        k = np.maximum(instance, 0)
        cam_chunk = cam["CameraChunkLocation"][0][:3].astype(np.int64)
        blk = (table["ChunkLocation"][inst["ChunkIndex"][k]].astype(np.int64) - cam_chunk) * 16 + inst["BlockLocation"][k][..., :3].astype(np.int64)
        out[f"eye{i}_block"] = np.where((instance >= 0)[..., None], blk, 0).astype(np.int32)
    out["vs_invalid_instance"] = np.stack([_vs(backend.lib, cams[0], scene_cfg, table, 0x7FFFFFFF, 0, DRAW_STAMP, 0),
                                           _vs(backend.lib, cams[0], scene_cfg, table, int(inst["ChunkIndex"][0]), 0, DRAW_STAMP + 1, 3)])
    return out


def _vs(lib, cam, scene_cfg, table, chunk_index, packed_loc, stamp, vertex):
    out = np.zeros(10, dtype=np.float32)
    lib.ref_vs_invoke(_p(cam), _p(scene_cfg), _p(np.ascontiguousarray(table)), C.c_uint32(chunk_index), C.c_uint32(packed_loc),
                      C.c_uint32(stamp), C.c_int(vertex), _p(out))
    return out


def check_records_against_ref_draw(draw, eye, rec, origin_chunk, edge=1e-3):
    """rec: (H,W) hit records of the DDA (from the CUDA kernel, or from the oracle it is bit-identical to) for camera
    draw[f"eye{eye}_camera"] over the draw_scene() volume.  Against the frame the reference's own shaders produced:
    same hit / miss, same block, same face, colour within 1 LSB -- on every pixel farther than `edge` block units from
    a face edge (for hits: read off the interpolated varying, which is the local position - 0.5) and, for misses, not
    adjacent to a covered pixel.  Returns (#hits compared, #misses compared, #skipped)."""
    inst, nrm, col = draw[f"eye{eye}_instance"], draw[f"eye{eye}_normal"].astype(np.float64), draw[f"eye{eye}_color"]
    cam = np.frombuffer(draw[f"eye{eye}_camera"].tobytes(), dtype=Camera160)[0]
    cam_chunk = cam["CameraChunkLocation"][:3].astype(np.int64)
    w0, w1 = rec["w0"].astype(np.int64), rec["w1"].astype(np.int64)
    d_hit = ((w1 >> 20) & 1).astype(bool)
    d_face = (w1 >> 16) & 7
    vox = np.stack([w0 & 0xFFFF, w0 >> 16, w1 & 0xFFFF], -1)
    d_blk = (vox >> 3) + (np.asarray(origin_chunk, dtype=np.int64) - cam_chunk) * 16
    d_rgba = np.stack([(rec["rgba"].astype(np.int64) >> (8 * c)) & 0xFF for c in range(4)], -1)

    covered = inst >= 0
    a = np.abs(nrm)
    ax = a.argmax(-1)
    top = np.take_along_axis(a, ax[..., None], -1)[..., 0]
    second = np.sort(a, -1)[..., 1]
    safe_hit = covered & (np.abs(top - 0.5) < 1e-6) & (second < 0.5 - edge)
    grown = covered.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            sh = np.zeros_like(covered)
            ys, yd = (slice(max(dy, 0), covered.shape[0] + min(dy, 0)), slice(max(-dy, 0), covered.shape[0] + min(-dy, 0)))
            xs, xd = (slice(max(dx, 0), covered.shape[1] + min(dx, 0)), slice(max(-dx, 0), covered.shape[1] + min(-dx, 0)))
            sh[yd, xd] = covered[ys, xs]
            grown |= sh
    safe_miss = ~grown
    assert d_hit[safe_hit].all(), "DDA misses where the reference draw covers the pixel"
    assert not d_hit[safe_miss].any(), "DDA hits where the reference draw leaves the clear colour"
    assert np.array_equal(d_blk[safe_hit], draw[f"eye{eye}_block"].astype(np.int64)[safe_hit]), "block"
    sign = np.take_along_axis(nrm, ax[..., None], -1)[..., 0] > 0
    assert np.array_equal(d_face[safe_hit], (2 * ax + sign)[safe_hit]), "face"
    r8 = np.floor(col.astype(np.float64) * 255.0 + 0.5).astype(np.int64)
    assert (np.abs(d_rgba - r8)[safe_hit] <= 1).all(), "colour differs by more than 1 LSB"
    miss_rgba = d_rgba[safe_miss]
    assert (miss_rgba == np.array([0, 0, 0, 255])).all(), "miss colour is not the clear colour (0,0,0,1)"
    return int(safe_hit.sum()), int(safe_miss.sum()), int((~safe_hit & ~safe_miss).sum())
