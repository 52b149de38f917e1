"""ctypes binding of the CPU oracle (oracle/libmeso_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (mesoengine_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_DIR, "libmeso_oracle.so")


def build(force=False):
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "CC=gcc"], stdout=subprocess.DEVNULL)
    return _SO


build()
lib = C.CDLL(_SO)

GPUBlock = np.dtype([("ChunkIndex", "<u4"), ("BlockLocation", "u1", (4,)), ("BlockFrameStamp", "<u4")])
GPUChunk = np.dtype([("ChunkLocation", "<i4", (3,)), ("ChunkFrameStamp", "<u4")])
HitRecord = np.dtype([("w0", "<u4"), ("w1", "<u4"), ("t", "<f4"), ("rgba", "<u4")])
Quad = np.dtype([("w0", "<u4"), ("w1", "<u4"), ("w2", "<u4"), ("w3", "<u4")])
Camera = np.dtype([("Projection", "<f4", (16,)), ("View", "<f4", (16,)), ("CameraChunkLocation", "<i4", (4,)),
                   ("SubCameraLocation", "<f4", (4,))])
SceneConfig = np.dtype([("BlockSize", "<f4"), ("BlockResolution", "<u4"), ("ChunkSize", "<f4"), ("ChunkResolution", "<u4")])
RaySetup = np.dtype([("o", "<f4", (3,)), ("two_over_w", "<f4"), ("U", "<f4", (3,)), ("two_over_h", "<f4"),
                     ("V", "<f4", (3,)), ("pad0", "<f4"), ("F", "<f4", (3,)), ("pad1", "<f4"),
                     ("L", "<f4", (3,)), ("pad2", "<f4")])
RayStats = np.dtype([("primary", "<u8"), ("shadow", "<u8"), ("hits", "<u8"), ("steps", "<u8"),
                     ("touched_chunks", "<u8"), ("touched_bricks", "<u8"), ("u_bytes", "<u8")])
assert GPUBlock.itemsize == 12 and GPUChunk.itemsize == 16 and Camera.itemsize == 160 and RaySetup.itemsize == 80

SDF_SPHERE, SDF_TERRAIN = 0, 1
GRAN_BLOCK, GRAN_VOXEL = 0, 1
SIN_LIBM, SIN_PORTABLE = 0, 1
FLAG_SHADOW = 1
DDA_FLAT, DDA_HIER, DDA_BOX, DDA_MODEL = 0, 1, 2, 3
REF_SPHERE = (100.0, 0.0, 0.0, 50.0)   # GeneratorHelper.h:134


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


def _i3(v):
    return np.ascontiguousarray(v, dtype=np.int32)


def _d4(v):
    a = np.zeros(4, dtype=np.float64)
    if v is not None:
        v = np.asarray(v, dtype=np.float64)
        a[: len(v)] = v
    return a


lib.orc_sin_portable.restype = C.c_double
lib.orc_sin_portable.argtypes = [C.c_double]
lib.orc_hash3.restype = C.c_double
lib.orc_hash3.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
lib.orc_displacement.restype = C.c_double
lib.orc_sdf.restype = C.c_double
lib.orc_sdf.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
lib.orc_volume_create.restype = C.c_void_p
lib.orc_volume_num_chunks.restype = C.c_int64
lib.orc_volume_occ.restype = C.c_void_p
lib.orc_volume_full.restype = C.c_void_p
for _n in ("orc_volume_num_partial", "orc_volume_export_partial", "orc_volume_count_voxels", "orc_volume_build_occupancy",
           "orc_mesh", "orc_mesh_bricks", "orc_mesh_chunk_faces", "orc_count_exposed_faces", "orc_carve_sphere"):
    getattr(lib, _n).restype = C.c_int64


def sin_portable(x):
    return lib.orc_sin_portable(float(x))


def hash3(x, y, z, sin_mode=SIN_LIBM):
    return lib.orc_hash3(float(x), float(y), float(z), sin_mode)


def noised(x, sin_mode=SIN_LIBM):
    xin = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(4, dtype=np.float64)
    lib.orc_noised(_p(xin), C.c_int(sin_mode), _p(out))
    return out


def displacement(p, sin_mode=SIN_LIBM):
    pin = np.ascontiguousarray(p, dtype=np.float64)
    return lib.orc_displacement(_p(pin), C.c_int(sin_mode))


def sdf(kind, params, x, y, z, sin_mode=SIN_LIBM):
    return lib.orc_sdf(kind, _p(_d4(params)), sin_mode, float(x), float(y), float(z))


def generate_chunk(kind, params, chunk_loc, sin_mode=SIN_LIBM, block_size=1.0, chunk_res=16):
    out = np.zeros((chunk_res ** 3, 3), dtype=np.uint8)
    n = lib.orc_generate_chunk(C.c_int(kind), _p(_d4(params)), C.c_int(sin_mode), _p(_i3(chunk_loc)),
                               C.c_float(block_size), C.c_int(chunk_res), _p(out))
    return out[:n].copy()


def erode_mips(blocks_xyz, depth=4, use26=True):
    b = np.ascontiguousarray(blocks_xyz, dtype=np.uint8).reshape(-1, 3)
    mips = np.zeros((depth, 64), dtype=np.uint64)
    lib.orc_erode_mips(_p(b), C.c_int(len(b)), C.c_int(depth), C.c_int(1 if use26 else 0), _p(mips))
    return mips


def emit_instances(blocks_xyz, mips, threshold=1, chunk_index=0, stamp=1):
    b = np.ascontiguousarray(blocks_xyz, dtype=np.uint8).reshape(-1, 3)
    out = np.zeros(max(len(b), 1), dtype=GPUBlock)
    n = lib.orc_emit_instances(_p(b), C.c_int(len(b)), _p(np.ascontiguousarray(mips)), C.c_int(threshold),
                               C.c_uint32(chunk_index), C.c_uint32(stamp), _p(out))
    return out[:n].copy()


def perspective(fovy, aspect, z_near, z_far):
    m = np.zeros(16, dtype=np.float32)
    lib.orc_perspective_rh_zo(C.c_float(fovy), C.c_float(aspect), C.c_float(z_near), C.c_float(z_far), _p(m))
    return m


def camera_uniform(eye, center, up=(0, 0, 1), fov_deg=60.0, z_near=0.1, z_far=1000.0, reverse_z=True,
                   width=1280, height=720, chunk_size=16.0):
    cam = np.zeros(1, dtype=Camera)
    lib.orc_camera_uniform(_p(_f3(eye)), _p(_f3(center)), _p(_f3(up)), C.c_float(fov_deg), C.c_float(z_near),
                           C.c_float(z_far), C.c_int(1 if reverse_z else 0), C.c_float(width), C.c_float(height),
                           C.c_float(chunk_size), _p(cam))
    return cam


def ray_setup(cam, origin_chunk, width, height, light=(0.3, 0.5, 0.8)):
    rs = np.zeros(1, dtype=RaySetup)
    lib.orc_ray_setup(_p(cam), _p(_i3(origin_chunk)), C.c_int(width), C.c_int(height), _p(_f3(light)), _p(rs))
    return rs


def fibonacci_sphere(samples, normalize=True):
    out = np.zeros((samples, 3), dtype=np.float64)
    lib.orc_fibonacci_sphere(C.c_uint32(samples), C.c_int(1 if normalize else 0), _p(out))
    return out


def triplanar_faces(octant):
    out = np.zeros(3, dtype=np.int32)
    lib.orc_triplanar_faces(C.c_int(octant), _p(out))
    return out.tolist()


def hw_threads():
    return int(lib.orc_hardware_threads())


# ---- K6: resident-set selection -------------------------------------------------------------
ChunkCandidate = np.dtype([("Importance", "<f4"), ("Offset", "<i4", (3,))])
assert ChunkCandidate.itemsize == 16
lib.orc_chunk_importance.restype = C.c_float
lib.orc_block_importance.restype = C.c_float
lib.orc_select_view_chunks.restype = C.c_int64
lib.orc_nearest_direction.restype = C.c_uint32


def chunk_importance(cam_chunk, fwd, loc):
    return float(lib.orc_chunk_importance(_p(_i3(cam_chunk)), _p(_f3(fwd)), _p(_i3(loc))))


def block_importance(cam_chunk, fwd, chunk, block, chunk_resolution=16):
    b = np.asarray(block, dtype=np.uint8)
    return float(lib.orc_block_importance(_p(_i3(cam_chunk)), _p(_f3(fwd)), _p(_i3(chunk)), _p(b), C.c_uint32(chunk_resolution)))


def select_view_chunks(fwd, forward_load=24, backward_load=6, view_angle=120.0, mode=0):
    side = 2 * forward_load + 1
    out = np.zeros(side ** 3, dtype=ChunkCandidate)
    n = lib.orc_select_view_chunks(_p(_f3(fwd)), C.c_uint32(forward_load), C.c_uint32(backward_load), C.c_float(view_angle),
                                   C.c_int(mode), _p(out), C.c_int64(out.shape[0]))
    assert n >= 0
    return out[:n].copy()


def fibonacci_sphere_f32(samples):
    out = np.zeros((samples, 3), dtype=np.float32)
    lib.orc_fibonacci_sphere_f32(C.c_uint32(samples), _p(out))
    return out


def nearest_direction(dirs, q):
    d = np.ascontiguousarray(dirs, dtype=np.float32)
    return int(lib.orc_nearest_direction(_p(d), C.c_uint32(d.shape[0]), _p(_f3(q))))


class Volume:
    def __init__(self, origin_chunk, dims_chunks):
        self.origin = _i3(origin_chunk)
        self.dims = _i3(dims_chunks)
        self.h = C.c_void_p(lib.orc_volume_create(_p(self.origin), _p(self.dims)))
        self.nchunks = int(np.prod(self.dims))

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_volume_destroy(self.h)
            self.h = None

    def voxelize(self, kind, params=None, granularity=GRAN_VOXEL, sin_mode=SIN_PORTABLE, nthreads=None, fast=False):
        self._params = _d4(params)
        lib.orc_volume_voxelize_ex(self.h, C.c_int(kind), _p(self._params), C.c_int(granularity), C.c_int(sin_mode),
                                   C.c_int(nthreads or hw_threads()), C.c_int(1 if fast else 0))
        return self

    def voxelize_lod(self, kind, params=None, mip=0, sin_mode=SIN_PORTABLE, nthreads=None):
        self._params = _d4(params)
        lib.orc_volume_voxelize_lod(self.h, C.c_int(kind), _p(self._params), C.c_int(sin_mode), C.c_int(nthreads or hw_threads()), C.c_int(mip))
        return self

    def occ(self):
        ptr = lib.orc_volume_occ(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(self.nchunks, 64)).copy()

    def full(self):
        ptr = lib.orc_volume_full(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(self.nchunks, 64)).copy()

    def num_partial(self):
        return int(lib.orc_volume_num_partial(self.h))

    def num_partial_slots(self):
        lib.orc_volume_pool_slots.restype = C.c_int64
        return int(lib.orc_volume_pool_slots(self.h))

    def export_partial(self):
        n = self.num_partial()
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        payload = np.zeros((max(n, 1), 8), dtype=np.uint64)
        lib.orc_volume_export_partial(self.h, _p(keys), _p(payload), C.c_int64(n))
        return keys[:n], payload[:n]

    def import_(self, occ, full, keys, payload):
        occ = np.ascontiguousarray(occ, dtype=np.uint64)
        full = np.ascontiguousarray(full, dtype=np.uint64)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        payload = np.ascontiguousarray(payload, dtype=np.uint64)
        lib.orc_volume_import(self.h, _p(occ), _p(full), _p(keys), _p(payload), C.c_int64(len(keys)))
        return self

    def get_voxel(self, x, y, z):
        return int(lib.orc_volume_get_voxel(self.h, C.c_int(x), C.c_int(y), C.c_int(z)))

    def count_voxels(self):
        return int(lib.orc_volume_count_voxels(self.h))

    def build_occupancy(self, stamp=1):
        table = np.zeros(self.nchunks, dtype=GPUChunk)
        mips = np.zeros((self.nchunks, 3, 64), dtype=np.uint64)
        n = int(lib.orc_volume_build_occupancy(self.h, C.c_uint32(stamp), _p(table), _p(mips), None, C.c_int64(0)))
        inst = np.zeros(max(n, 1), dtype=GPUBlock)
        lib.orc_volume_build_occupancy(self.h, C.c_uint32(stamp), _p(table), _p(mips), _p(inst), C.c_int64(n))
        return table, mips, inst[:n]

    def raymarch(self, rs, width, height, rect=None, shadow=True, mode=DDA_HIER, nthreads=None, stats=False):
        x0, y0, x1, y1 = rect if rect is not None else (0, 0, width, height)
        rec = np.zeros((height, width), dtype=HitRecord)
        st = np.zeros(1, dtype=RayStats)
        lib.orc_raymarch(self.h, _p(rs), C.c_int(width), C.c_int(height), C.c_int(x0), C.c_int(y0), C.c_int(x1),
                         C.c_int(y1), C.c_uint32(FLAG_SHADOW if shadow else 0), C.c_int(mode),
                         C.c_int(nthreads or hw_threads()), _p(rec), _p(st) if stats else None)
        return (rec, st[0]) if stats else rec

    def raymarch_rows(self, rs, width, height, rows, shadow=True, mode=DDA_HIER, nthreads=None, out=None):
        """Scanlines `rows` of the frame in one parallel region -> (records [height, width] with those rows filled, stats)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        rec = out if out is not None else np.zeros((height, width), dtype=HitRecord)
        st = np.zeros(1, dtype=RayStats)
        lib.orc_raymarch_rows(self.h, _p(rs), C.c_int(width), C.c_int(height), _p(rows), C.c_int(len(rows)),
                              C.c_uint32(FLAG_SHADOW if shadow else 0), C.c_int(mode), C.c_int(nthreads or hw_threads()), _p(rec), _p(st))
        return rec, st[0]

    def mesh(self, nthreads=None):
        n = int(lib.orc_mesh(self.h, C.c_int(nthreads or hw_threads()), None, C.c_int64(0)))
        q = np.zeros(max(n, 1), dtype=Quad)
        lib.orc_mesh(self.h, C.c_int(nthreads or hw_threads()), _p(q), C.c_int64(n))
        return q[:n]

    def mesh_bricks(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = int(lib.orc_mesh_bricks(self.h, _p(keys), C.c_int64(len(keys)), None, C.c_int64(0)))
        q = np.zeros(max(n, 1), dtype=Quad)
        lib.orc_mesh_bricks(self.h, _p(keys), C.c_int64(len(keys)), _p(q), C.c_int64(n))
        return q[:n]

    def mesh_chunk_faces(self, chunks):
        """Brick-level quads (full-brick faces towards absent bricks, merged per chunk) of the listed chunks."""
        chunks = np.ascontiguousarray(chunks, dtype=np.int64)
        n = int(lib.orc_mesh_chunk_faces(self.h, _p(chunks), C.c_int64(len(chunks)), None, C.c_int64(0)))
        q = np.zeros(max(n, 1), dtype=Quad)
        lib.orc_mesh_chunk_faces(self.h, _p(chunks), C.c_int64(len(chunks)), _p(q), C.c_int64(n))
        return q[:n]

    def remesh(self, keys):
        """What a re-mesh of the listed bricks returns: their voxel-level quads and the brick-level quads of every chunk that
        holds one of them."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        chunks = np.unique(keys >> np.uint64(12)).astype(np.int64)
        return np.concatenate([self.mesh_bricks(keys), self.mesh_chunk_faces(chunks)])

    def count_exposed_faces(self):
        return int(lib.orc_count_exposed_faces(self.h))

    def carve_sphere(self, center, radius):
        cap = 1 << 20
        dirty = np.zeros(cap, dtype=np.uint64)
        n = int(lib.orc_carve_sphere(self.h, _p(_i3(center)), C.c_int32(radius), _p(dirty), C.c_int64(cap)))
        assert n <= cap
        return dirty[:n].copy()


def step_model(vol, df_shift=5, df_cap=32, probe=True, directional=False, brick_cap=0, cell2=True):
    """Configure the step-count model (ORC_DDA_MODEL) and build its field for `vol`. Instrumentation, not parity."""
    lib.orc_step_model_config(C.c_int(df_shift), C.c_int(df_cap), C.c_int(1 if probe else 0), C.c_int(1 if directional else 0),
                              C.c_int(brick_cap), C.c_int(int(cell2)))
    if lib.orc_step_model_build(vol.h) != 0:
        raise MemoryError("orc_step_model_build")


def set_step_image(img, width=0):
    """img: (H, W, 2) uint32 array receiving every pixel's primary / shadow step counts, or None to switch it off."""
    lib.orc_debug_set_step_image(_p(img) if img is not None else None, C.c_int(width))


def step_model_counts(reset=True):
    """-> steps by kind: voxel, 2^3 cell, brick, field step <= 2 cells, field step > 2 cells, grid entry."""
    out = np.zeros(6, dtype=np.uint64)
    lib.orc_step_model_counts(_p(out), C.c_int(1 if reset else 0))
    return out


def cube_tables(vol, extra_slots=0):
    """The forward-cube tables of csrc/k_cubes.cu built on the CPU: (cell [8, ncells] u8, brick [nchunks*4096] u16,
    cell2 [(pool_n + extra_slots)*64] u16; the extra entries stay zero = "one cell", as on the GPU, for payload slots a
    later carve allocates)."""
    ncells = int(np.prod(vol.dims.astype(np.int64) * 4))
    cell = np.zeros((8, ncells), dtype=np.uint8)
    brick = np.zeros(vol.nchunks * 4096, dtype=np.uint16)
    cell2 = np.zeros((max(int(vol.num_partial_slots()), 1) + extra_slots) * 64, dtype=np.uint16)
    if lib.orc_cube_tables(vol.h, _p(cell), _p(brick), _p(cell2)) != 0:
        raise MemoryError("orc_cube_tables")
    return cell, brick, cell2


def cube_tables_use(vol, tables):
    """Make DDA_MODEL read `tables` (as returned by cube_tables; keep them alive) or, with None, compute cubes on the fly."""
    if tables is None:
        lib.orc_cube_tables_use(vol.h, None, None, None)
    else:
        lib.orc_cube_tables_use(vol.h, _p(tables[0]), _p(tables[1]), _p(tables[2]))


def sort_quads(q):
    q = np.ascontiguousarray(q).copy()
    lib.orc_sort_quads(_p(q), C.c_int64(len(q)))
    return q


def ref_instanced_pixel(cam, scene, chunks, blocks, width, height, px, py):
    blk = np.zeros(3, dtype=np.int32)
    face = C.c_int(7)
    t = C.c_double(0)
    margin = C.c_double(0)
    rgba = np.zeros(4, dtype=np.float32)
    hit = lib.orc_ref_instanced_pixel(_p(cam), _p(scene), _p(chunks), C.c_int64(len(chunks)), _p(blocks),
                                      C.c_int64(len(blocks)), C.c_int(width), C.c_int(height), C.c_int(px), C.c_int(py),
                                      _p(blk), C.byref(face), C.byref(t), C.byref(margin), _p(rgba))
    return bool(hit), blk, face.value, t.value, margin.value, rgba


def default_scene_config():
    s = np.zeros(1, dtype=SceneConfig)
    s["BlockSize"] = 1.0
    s["BlockResolution"] = 8
    s["ChunkSize"] = 16.0
    s["ChunkResolution"] = 16
    return s


def unpack_records(rec):
    """-> dict of arrays: x,y,z,face,shadow,hit,t,rgba"""
    w0 = rec["w0"]; w1 = rec["w1"]
    return dict(x=w0 & 0xFFFF, y=w0 >> 16, z=w1 & 0xFFFF, face=(w1 >> 16) & 7, shadow=(w1 >> 19) & 1,
                hit=(w1 >> 20) & 1, t=rec["t"], rgba=rec["rgba"])
