"""ctypes binding of libmeso_b200.so (include/meso_cuda.h).

Plumbing only: every call goes straight through the C ABI to the hand-written CUDA kernels.  There is no Python or
CPU implementation behind any of these methods; if the shared library is missing the import fails, and if no
sm_100 device is present `Context()` raises with the library's own message.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MESO_SO: measurement tools load an A/B build of the same library (tools/rm_ab.py); unset everywhere else
SO_PATH = os.environ.get("MESO_SO") or os.path.join(_HERE, "libmeso_b200.so")

GPUBlock = np.dtype([("ChunkIndex", "<u4"), ("BlockLocation", "u1", (4,)), ("BlockFrameStamp", "<u4")])
GPUChunk = np.dtype([("ChunkLocation", "<i4", (3,)), ("ChunkFrameStamp", "<u4")])
SimpleInstanceData = np.dtype([("Position", "<f4", (3,)), ("ChunkLocation", "<i4", (3,)), ("Scale", "<f4"), ("Rotation", "<f4", (4,)), ("Marker", "<f4")])
DebugStats = np.dtype([("VisibleChunk", "<u4"), ("LoadedChunk", "<u4"), ("LoadedChunkWithBlocks", "<u4"), ("NewlyAddedVisibleChunk", "<u4"),
                       ("MissingChunk", "<u4"), ("PayloadSlotsHandedOut", "<u4"), ("PayloadSlotsFree", "<u4"), ("_pad", "<u4"), ("LoadedBlock", "<i8")])
HitRecord = np.dtype([("w0", "<u4"), ("w1", "<u4"), ("t", "<f4"), ("rgba", "<u4")])
Quad = np.dtype([("w0", "<u4"), ("w1", "<u4"), ("w2", "<u4"), ("w3", "<u4")])
Camera = np.dtype([("Projection", "<f4", (16,)), ("View", "<f4", (16,)), ("CameraChunkLocation", "<i4", (4,)),
                   ("SubCameraLocation", "<f4", (4,))])
SceneConfig = np.dtype([("BlockSize", "<f4"), ("BlockResolution", "<u4"), ("ChunkSize", "<f4"), ("ChunkResolution", "<u4")])
RaySetup = np.dtype([("o", "<f4", (3,)), ("two_over_w", "<f4"), ("U", "<f4", (3,)), ("two_over_h", "<f4"),
                     ("V", "<f4", (3,)), ("pad0", "<f4"), ("F", "<f4", (3,)), ("pad1", "<f4"),
                     ("L", "<f4", (3,)), ("pad2", "<f4")])
RayStats = np.dtype([("primary", "<u8"), ("shadow", "<u8"), ("hits", "<u8"), ("steps", "<u8"),
                     ("touched_chunks", "<u8"), ("touched_bricks", "<u8"), ("u_bytes", "<u8"),
                     ("steps_primary", "<u8"), ("warp_slots_primary", "<u8"), ("warp_slots_shadow", "<u8"),
                     ("level_steps", "<u8", (5,))])
ChunkCandidate = np.dtype([("Importance", "<f4"), ("Offset", "<i4", (3,))])          # FTempChunkDataType
ViewConfig = np.dtype([("ViewForwardLoadChunkSize", "<u4"), ("ViewBackwardLoadChunkSize", "<u4"), ("ViewChunkAngle", "<f4"),
                       ("Mode", "<u4")])
StreamStats = np.dtype([("generated", "<u4"), ("missing", "<u4"), ("candidates", "<u4"), ("in_window", "<u4")])

SDF_SPHERE, SDF_TERRAIN = 0, 1
GRAN_BLOCK, GRAN_VOXEL = 0, 1
FLAG_SHADOW, FLAG_RGBA8, FLAG_CUBES, FLAG_NO_CUBES = 1, 2, 4, 8
LAYOUT_FRAME, LAYOUT_TILES, LAYOUT_SLABS = 0, 1, 2
TILE_W, TILE_H = 32, 8

# every symbol include/meso_cuda.h declares (tests/test_abi.py checks the header against this and the .so)
SYMBOLS = [
    "meso_last_error", "meso_abi_version", "meso_ctx_create", "meso_ctx_destroy", "meso_ctx_set_stream", "meso_ctx_use_own_stream", "meso_ctx_sync",
    "meso_ctx_set_partition", "meso_device_sm_count", "meso_scene_create", "meso_voxelize_sdf", "meso_volume_upload",
    "meso_volume_num_partial", "meso_volume_download", "meso_build_occupancy", "meso_download_chunk_table",
    "meso_download_mips", "meso_download_instances", "meso_ray_setup", "meso_raymarch", "meso_raymarch_device",
    "meso_raymarch_stats", "meso_compose_tiles_device", "meso_tiles_per_rank", "meso_mesh", "meso_mesh_device",
    "meso_carve_sphere", "meso_download_dirty", "meso_remesh_dirty", "meso_host_alloc", "meso_host_free",
    "meso_flush_l2", "meso_launch_count",
    "meso_raymarch_async", "meso_frame_wait", "meso_device_alloc", "meso_device_free", "meso_ipc_export", "meso_ipc_open", "meso_ipc_close", "meso_download", "meso_download_async",
    "meso_select_view_chunks", "meso_chunk_importance", "meso_baked_direction", "meso_stream_begin", "meso_stream_update",
    "meso_stream_update_async", "meso_stream_stats", "meso_stream_loaded",
    "meso_host_register", "meso_host_unregister", "meso_mesh_device_shared", "meso_device_memset",
    "meso_device_copy", "meso_build_cubes", "meso_download_cubes", "meso_volume_upload_blocks", "meso_raymarch_device_slabs", "meso_mesh_count_device", "meso_stream_recentre", "meso_pool_stats", "meso_block_importance",
    "meso_signal_device", "meso_present_rgba8", "meso_pack_rgba8_device", "meso_debug_chunk_instances", "meso_debug_stats", "meso_voxelize_sdf_lod", "meso_wait_device", "meso_wait_timed_out",
    "meso_group_create", "meso_group_destroy", "meso_group_size", "meso_group_ctx", "meso_group_sync", "meso_group_scene_create",
    "meso_group_voxelize_sdf", "meso_group_volume_upload_blocks", "meso_group_carve_sphere", "meso_group_raymarch",
    "meso_group_raymarch_async", "meso_group_frame_wait", "meso_group_mesh", "meso_group_mesh_device", "meso_group_remesh_dirty",
]
IPC_HANDLE_BYTES = 64
UPLOAD_MERGE = 1


class MesoError(RuntimeError):
    pass


def load():
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "mesoengine_b200: %s is missing -- build it with `python -m mesoengine_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % SO_PATH)
    lib = C.CDLL(SO_PATH)
    lib.meso_last_error.restype = C.c_char_p
    lib.meso_tiles_per_rank.restype = C.c_int64
    lib.meso_launch_count.restype = C.c_int64
    lib.meso_launch_count.argtypes = [C.c_void_p]
    lib.meso_device_sm_count.argtypes = [C.c_void_p]
    lib.meso_group_ctx.restype = C.c_void_p
    lib.meso_group_ctx.argtypes = [C.c_void_p, C.c_int]
    return lib


lib = load()


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _ck(code):
    if code != 0:
        raise MesoError("libmeso_b200 error %d: %s" % (code, lib.meso_last_error().decode()))


def default_scene_config():
    s = np.zeros(1, dtype=SceneConfig)
    s["BlockSize"] = 1.0
    s["BlockResolution"] = 8
    s["ChunkSize"] = 16.0
    s["ChunkResolution"] = 16
    return s


def ray_setup(cam, origin_chunk, width, height, light=(0.3, 0.5, 0.8)):
    """Pure host function of the ABI (no device needed)."""
    rs = np.zeros(1, dtype=RaySetup)
    o = np.ascontiguousarray(origin_chunk, dtype=np.int32)
    l = np.ascontiguousarray(light, dtype=np.float32)
    _ck(lib.meso_ray_setup(_p(cam), _p(o), C.c_int(width), C.c_int(height), _p(l), _p(rs)))
    return rs


def tiles_per_rank(width, height, world):
    return int(lib.meso_tiles_per_rank(C.c_int(width), C.c_int(height), C.c_int(world)))


def view_config(forward_load=24, backward_load=6, view_angle=120.0, mode=0):
    """The view fields of FVoxelSceneConfig (VoxelSceneConfig.h:34-35,40) with the reference's defaults."""
    v = np.zeros(1, dtype=ViewConfig)
    v["ViewForwardLoadChunkSize"] = forward_load
    v["ViewBackwardLoadChunkSize"] = backward_load
    v["ViewChunkAngle"] = view_angle
    v["Mode"] = mode
    return v


def baked_direction(samples, forward):
    """Pure host function: the baked Fibonacci direction the reference's TNearestMap returns for `forward`."""
    f = np.ascontiguousarray(forward, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    idx = C.c_uint32(0)
    _ck(lib.meso_baked_direction(C.c_uint32(samples), _p(f), _p(out), C.byref(idx)))
    return out, idx.value


class Context:
    """One GPU, one stream.  Mirrors the role of lvk::IContext for the voxel path."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _ck(lib.meso_ctx_create(C.c_int(device), C.byref(h)))
        self.h = h
        self.origin = None
        self.dims = None
        self.nchunks = 0

    def close(self):
        if getattr(self, "h", None) and lib is not None:  # `lib` is already None during interpreter shutdown
            lib.meso_ctx_destroy(self.h)
        self.h = None

    def __del__(self):
        self.close()

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t handle (0 = CUDA default stream)."""
        _ck(lib.meso_ctx_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def use_own_stream(self):
        _ck(lib.meso_ctx_use_own_stream(self.h))

    def sync(self):
        _ck(lib.meso_ctx_sync(self.h))

    def set_partition(self, rank, world):
        _ck(lib.meso_ctx_set_partition(self.h, C.c_int(rank), C.c_int(world)))
        self.rank, self.world = rank, world

    def sm_count(self):
        return int(lib.meso_device_sm_count(self.h))

    def launch_count(self):
        return int(lib.meso_launch_count(self.h))

    def scene_create(self, origin_chunk, dims_chunks, max_bricks, cfg=None):
        cfg = default_scene_config() if cfg is None else cfg
        self.origin = np.ascontiguousarray(origin_chunk, dtype=np.int32)
        self.dims = np.ascontiguousarray(dims_chunks, dtype=np.int32)
        self.nchunks = int(np.prod(self.dims.astype(np.int64)))
        _ck(lib.meso_scene_create(self.h, _p(cfg), _p(self.origin), _p(self.dims), C.c_uint32(max_bricks)))

    def voxelize_sdf(self, kind, params=None, granularity=GRAN_VOXEL):
        p = np.zeros(4, dtype=np.float64)
        if params is not None:
            p[: len(params)] = np.asarray(params, dtype=np.float64)
        _ck(lib.meso_voxelize_sdf(self.h, C.c_int(kind), _p(p), C.c_int(granularity)))

    def voxelize_sdf_lod(self, kind, params, mipmap_level):
        p = np.ascontiguousarray(params if params is not None else [0, 0, 0, 0], dtype=np.float64)
        _ck(lib.meso_voxelize_sdf_lod(self.h, C.c_int(kind), _p(p), C.c_uint32(mipmap_level)))

    def volume_upload(self, occ, full, keys, payload):
        occ = np.ascontiguousarray(occ, dtype=np.uint64)
        full = np.ascontiguousarray(full, dtype=np.uint64)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        payload = np.ascontiguousarray(payload, dtype=np.uint64)
        _ck(lib.meso_volume_upload(self.h, _p(occ), _p(full), _p(keys), _p(payload), C.c_int64(len(keys))))

    def volume_upload_blocks(self, chunks, blocks, merge=False):
        """The reference's own records: FGPUChunk table + FGPUBlock pool (ChunkPool.h:662-679).  Returns the accepted count."""
        chunks = np.ascontiguousarray(chunks, dtype=GPUChunk)
        blocks = np.ascontiguousarray(blocks, dtype=GPUBlock)
        n = C.c_int64(0)
        _ck(lib.meso_volume_upload_blocks(self.h, _p(chunks), C.c_int64(len(chunks)), _p(blocks), C.c_int64(len(blocks)),
                                          C.c_uint32(UPLOAD_MERGE if merge else 0), C.byref(n)))
        return n.value

    def volume_download(self):
        occ = np.zeros((self.nchunks, 64), dtype=np.uint64)
        full = np.zeros((self.nchunks, 64), dtype=np.uint64)
        n = C.c_int64(0)
        _ck(lib.meso_volume_download(self.h, _p(occ), _p(full), None, None, C.c_int64(0), C.byref(n)))
        keys = np.zeros(max(n.value, 1), dtype=np.uint64)
        payload = np.zeros((max(n.value, 1), 8), dtype=np.uint64)
        if n.value:
            _ck(lib.meso_volume_download(self.h, _p(occ), _p(full), _p(keys), _p(payload), C.c_int64(n.value), C.byref(n)))
        return occ, full, keys[: n.value], payload[: n.value]

    def build_occupancy(self, stamp=1):
        n = C.c_int64(0)
        _ck(lib.meso_build_occupancy(self.h, C.c_uint32(stamp), C.byref(n)))
        return n.value

    def download_occupancy(self, n_inst):
        table = np.zeros(self.nchunks, dtype=GPUChunk)
        mips = np.zeros((self.nchunks, 3, 64), dtype=np.uint64)
        inst = np.zeros(max(n_inst, 1), dtype=GPUBlock)
        _ck(lib.meso_download_chunk_table(self.h, _p(table)))
        _ck(lib.meso_download_mips(self.h, _p(mips)))
        _ck(lib.meso_download_instances(self.h, _p(inst), C.c_int64(max(n_inst, 1))))
        return table, mips, inst[:n_inst]

    def build_cubes(self):
        """(Re)build the per-octant forward-cube tables that FLAG_CUBES reads (opt-in raymarch path)."""
        _ck(lib.meso_build_cubes(self.h))

    def download_cubes(self):
        """-> (cell [8, ncells] u8, brick [nchunks*4096] u16) as built by build_cubes (debug / tests)."""
        ncells = int(np.prod(self.dims.astype(np.int64) * 4))
        cell = np.zeros((8, ncells), dtype=np.uint8)
        brick = np.zeros(self.nchunks * 4096, dtype=np.uint16)
        _ck(lib.meso_download_cubes(self.h, _p(cell), _p(brick)))
        return cell, brick

    @staticmethod
    def _cubes_flag(cubes):
        """cubes=None: the library's default (forward cubes whenever the tables are current); True insists on them; False
        forces the distance-field walk."""
        return 0 if cubes is None else (FLAG_CUBES if cubes else FLAG_NO_CUBES)

    def raymarch(self, cam, width, height, shadow=True, light=(0.3, 0.5, 0.8), out=None, rgba8=False, cubes=None):
        """End-to-end call: camera from host memory, records (or, rgba8=True, the colour image as uint32) into host memory."""
        rec = out if out is not None else np.zeros((height, width), dtype=np.uint32 if rgba8 else HitRecord)
        l = np.ascontiguousarray(light, dtype=np.float32)
        flags = (FLAG_SHADOW if shadow else 0) | (FLAG_RGBA8 if rgba8 else 0) | self._cubes_flag(cubes)
        _ck(lib.meso_raymarch(self.h, _p(cam), C.c_int(width), C.c_int(height), C.c_uint32(flags), _p(l), _p(rec)))
        return rec

    def present_rgba8(self, cam, width, height, rect=None, shadow=True, light=(0.3, 0.5, 0.8)):
        """Tightly packed RGBA8 of the range rect = (x, y, w, h) (None = whole frame): the data of IContext::upload(TextureHandle, ...)."""
        x, y, w, h = rect if rect is not None else (0, 0, width, height)
        out = np.zeros((h, w), dtype=np.uint32)
        l = np.ascontiguousarray(light, dtype=np.float32)
        rg = np.array([x, y, w, h], dtype=np.uint32)
        _ck(lib.meso_present_rgba8(self.h, _p(cam), C.c_int(width), C.c_int(height), C.c_uint32(FLAG_SHADOW if shadow else 0), _p(l),
                                   _p(rg) if rect is not None else None, _p(out)))
        return out

    def raymarch_async(self, cam, width, height, out, slot, shadow=True, light=(0.3, 0.5, 0.8), rgba8=False, cubes=None):
        """Frame-ring call: enqueue frame + copy into `out` (pinned numpy array); pair with frame_wait(slot)."""
        l = np.ascontiguousarray(light, dtype=np.float32)
        flags = (FLAG_SHADOW if shadow else 0) | (FLAG_RGBA8 if rgba8 else 0) | self._cubes_flag(cubes)
        _ck(lib.meso_raymarch_async(self.h, _p(cam), C.c_int(width), C.c_int(height), C.c_uint32(flags), _p(l), _p(out), C.c_int(slot)))

    def frame_wait(self, slot):
        _ck(lib.meso_frame_wait(self.h, C.c_int(slot)))

    def raymarch_device(self, cam, width, height, d_records, shadow=True, light=(0.3, 0.5, 0.8), layout=LAYOUT_FRAME, flags_extra=0):
        l = np.ascontiguousarray(light, dtype=np.float32)
        _ck(lib.meso_raymarch_device(self.h, _p(cam), C.c_int(width), C.c_int(height),
                                     C.c_uint32((FLAG_SHADOW if shadow else 0) | flags_extra), _p(l), C.c_void_p(d_records), C.c_int(layout)))

    def debug_chunk_instances(self):
        out = np.zeros(max(self.nchunks, 1), dtype=SimpleInstanceData)
        n = C.c_int64(0)
        _ck(lib.meso_debug_chunk_instances(self.h, _p(out), C.c_int64(len(out)), C.byref(n)))
        return out[: n.value]

    def debug_stats(self):
        st = np.zeros(1, dtype=DebugStats)
        _ck(lib.meso_debug_stats(self.h, _p(st)))
        return st[0]

    def pack_rgba8_device(self, d_records, n, d_rgba8):
        _ck(lib.meso_pack_rgba8_device(self.h, C.c_void_p(int(d_records)), C.c_int64(n), C.c_void_p(int(d_rgba8))))

    def signal_device(self, words):
        arr = (C.c_void_p * len(words))(*[C.c_void_p(int(p)) for p in words])
        _ck(lib.meso_signal_device(self.h, arr, C.c_int(len(words))))

    def wait_device(self, d_word, target):
        _ck(lib.meso_wait_device(self.h, C.c_void_p(int(d_word)), C.c_uint32(target & 0xFFFFFFFF)))

    def wait_timed_out(self):
        v = C.c_int(0)
        _ck(lib.meso_wait_timed_out(self.h, C.byref(v)))
        return bool(v.value)

    def stream_recentre(self, camera_chunk):
        """Moving window: the grid follows the camera; returns the origin's displacement in chunks."""
        cc = np.ascontiguousarray(camera_chunk, dtype=np.int32)
        moved = np.zeros(3, dtype=np.int32)
        _ck(lib.meso_stream_recentre(self.h, _p(cc), _p(moved)))
        if self.origin is not None:
            self.origin = tuple(int(o) + int(m) for o, m in zip(self.origin, moved))
        return tuple(int(m) for m in moved)

    def pool_stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        _ck(lib.meso_pool_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def block_importance(self, camera_chunk, forward, chunk_locations, block_locations, chunk_resolution=16):
        cc = np.ascontiguousarray(camera_chunk, dtype=np.int32)
        f = np.ascontiguousarray(forward, dtype=np.float32)
        cl = np.ascontiguousarray(chunk_locations, dtype=np.int32).reshape(-1, 3)
        bl = np.ascontiguousarray(block_locations, dtype=np.uint8).reshape(-1, 3)
        out = np.zeros(len(cl), dtype=np.float32)
        _ck(lib.meso_block_importance(self.h, _p(cc), _p(f), _p(cl), _p(bl), C.c_int64(len(cl)), C.c_uint32(chunk_resolution), _p(out)))
        return out

    def raymarch_device_slabs(self, cam, width, height, slab_ptrs, rows_per_slab, shadow=True, light=(0.3, 0.5, 0.8), flags_extra=0):
        """MESO_LAYOUT_SLABS: slab_ptrs[k] = device address (own or peer) of rows [k * rows_per_slab, ...) of the frame."""
        l = np.ascontiguousarray(light, dtype=np.float32)
        arr = (C.c_void_p * len(slab_ptrs))(*[C.c_void_p(int(p)) for p in slab_ptrs])
        _ck(lib.meso_raymarch_device_slabs(self.h, _p(cam), C.c_int(width), C.c_int(height), C.c_uint32((FLAG_SHADOW if shadow else 0) | flags_extra),
                                           _p(l), arr, C.c_int(len(slab_ptrs)), C.c_int(rows_per_slab)))

    def raymarch_stats(self, cam, width, height, shadow=True, light=(0.3, 0.5, 0.8), cubes=None):
        st = np.zeros(1, dtype=RayStats)
        l = np.ascontiguousarray(light, dtype=np.float32)
        _ck(lib.meso_raymarch_stats(self.h, _p(cam), C.c_int(width), C.c_int(height),
                                    C.c_uint32((FLAG_SHADOW if shadow else 0) | self._cubes_flag(cubes)), _p(l), _p(st)))
        return st[0]

    def compose_tiles_device(self, d_tiles, world, width, height, d_frame):
        _ck(lib.meso_compose_tiles_device(self.h, C.c_void_p(d_tiles), C.c_int(world), C.c_int(width), C.c_int(height),
                                          C.c_void_p(d_frame)))

    def mesh(self, cap):
        q = np.zeros(max(cap, 1), dtype=Quad)
        n = C.c_int64(0)
        _ck(lib.meso_mesh(self.h, _p(q), C.c_int64(cap), C.byref(n)))
        if n.value > cap:
            raise MesoError("mesh: %d quads exceed cap %d" % (n.value, cap))
        return q[: n.value]

    def mesh_device(self, d_quads, cap, want_count=True):
        n = C.c_int64(0)
        _ck(lib.meso_mesh_device(self.h, C.c_void_p(d_quads), C.c_int64(cap), C.byref(n) if want_count else None))
        return n.value

    def mesh_count_device(self, d_count_out):
        """8-byte count of the last mesh_device -> device buffer, stream-ordered (no host wait)."""
        _ck(lib.meso_mesh_count_device(self.h, C.c_void_p(int(d_count_out))))

    def mesh_device_shared(self, d_quads, d_counter, cap):
        """Fused quad gather: append this rank's quads to a list (and 8-byte counter) that may live in a peer GPU."""
        _ck(lib.meso_mesh_device_shared(self.h, C.c_void_p(d_quads), C.c_void_p(d_counter), C.c_int64(cap)))

    def carve_sphere(self, center, radius):
        c = np.ascontiguousarray(center, dtype=np.int32)
        n = C.c_int64(0)
        _ck(lib.meso_carve_sphere(self.h, _p(c), C.c_int32(radius), C.byref(n)))
        return n.value

    def download_dirty(self, n):
        keys = np.zeros(max(n, 1), dtype=np.uint64)
        _ck(lib.meso_download_dirty(self.h, _p(keys), C.c_int64(max(n, 1))))
        return keys[:n]

    def remesh_dirty(self, cap, cap_keys):
        q = np.zeros(max(cap, 1), dtype=Quad)
        keys = np.zeros(max(cap_keys, 1), dtype=np.uint64)
        n = C.c_int64(0)
        nk = C.c_int64(0)
        _ck(lib.meso_remesh_dirty(self.h, _p(q), C.c_int64(cap), C.byref(n), _p(keys), C.c_int64(cap_keys), C.byref(nk)))
        if n.value > cap:
            raise MesoError("remesh: %d quads exceed cap %d" % (n.value, cap))
        return q[: n.value], keys[: nk.value]

    def flush_l2(self):
        _ck(lib.meso_flush_l2(self.h))

    # ---- K6: resident-set selection / streaming generation ----
    def select_view_chunks(self, forward, view=None):
        """FChunkManageHelper::GetDesiredShowChunkLocationByView in priority-queue pop order."""
        view = view_config() if view is None else view
        f = np.ascontiguousarray(forward, dtype=np.float32)
        side = 2 * int(view["ViewForwardLoadChunkSize"][0]) + 1
        out = np.zeros(side ** 3, dtype=ChunkCandidate)
        n = C.c_int64(0)
        _ck(lib.meso_select_view_chunks(self.h, _p(f), _p(view), _p(out), C.c_int64(out.shape[0]), C.byref(n)))
        return out[: n.value]

    def chunk_importance(self, camera_chunk, forward, locations):
        loc = np.ascontiguousarray(locations, dtype=np.int32).reshape(-1, 3)
        out = np.zeros(loc.shape[0], dtype=np.float32)
        cc = np.ascontiguousarray(camera_chunk, dtype=np.int32)
        f = np.ascontiguousarray(forward, dtype=np.float32)
        _ck(lib.meso_chunk_importance(self.h, _p(cc), _p(f), _p(loc), C.c_int64(loc.shape[0]), _p(out)))
        return out

    def stream_begin(self, kind, params=None, granularity=GRAN_VOXEL):
        p = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        _ck(lib.meso_stream_begin(self.h, C.c_int(kind), _p(p), C.c_int(granularity)))

    def stream_update(self, camera_chunk, forward, max_new, view=None, wait=True):
        """FChunkManage::UpdateChunks + UpdateLoadingQueue; returns StreamStats (wait=True) or None (enqueue only)."""
        view = view_config() if view is None else view
        cc = np.ascontiguousarray(camera_chunk, dtype=np.int32)
        f = np.ascontiguousarray(forward, dtype=np.float32)
        if not wait:
            _ck(lib.meso_stream_update_async(self.h, _p(cc), _p(f), _p(view), C.c_uint32(max_new)))
            return None
        st = np.zeros(1, dtype=StreamStats)
        _ck(lib.meso_stream_update(self.h, _p(cc), _p(f), _p(view), C.c_uint32(max_new), _p(st)))
        return st[0]

    def stream_stats(self):
        st = np.zeros(1, dtype=StreamStats)
        _ck(lib.meso_stream_stats(self.h, _p(st)))
        return st[0]

    def stream_loaded(self, nchunks):
        """bool per chunk slot of the window: generated."""
        words = np.zeros((nchunks + 31) // 32, dtype=np.uint32)
        _ck(lib.meso_stream_loaded(self.h, _p(words), C.c_int64(words.shape[0])))
        return np.unpackbits(words.view(np.uint8), bitorder="little")[:nchunks].astype(bool)

    # ---- fused gather into host memory ----
    def host_register(self, host_array):
        """Page-lock + map a host numpy buffer (e.g. an mmap of a shared-memory file); returns the device address."""
        d = C.c_void_p()
        _ck(lib.meso_host_register(self.h, _p(host_array), C.c_size_t(host_array.nbytes), C.byref(d)))
        return d.value

    def host_unregister(self, host_array):
        _ck(lib.meso_host_unregister(self.h, _p(host_array)))

    # ---- peer memory (fused gather) ----
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        _ck(lib.meso_device_alloc(self.h, C.c_size_t(nbytes), C.byref(p)))
        return p.value

    def device_memset(self, dptr, value, nbytes):
        _ck(lib.meso_device_memset(self.h, C.c_void_p(dptr), C.c_int(value), C.c_size_t(nbytes)))

    def device_copy(self, dst_dptr, src_dptr, nbytes):
        """device -> device on the context's stream (enqueue only); either side may be an ipc_open()ed peer pointer."""
        _ck(lib.meso_device_copy(self.h, C.c_void_p(dst_dptr), C.c_void_p(src_dptr), C.c_size_t(nbytes)))

    def device_free(self, dptr):
        _ck(lib.meso_device_free(self.h, C.c_void_p(dptr)))

    def ipc_export(self, dptr):
        h = np.zeros(IPC_HANDLE_BYTES, dtype=np.uint8)
        _ck(lib.meso_ipc_export(self.h, C.c_void_p(dptr), _p(h)))
        return h

    def ipc_open(self, handle):
        h = np.ascontiguousarray(handle, dtype=np.uint8)
        p = C.c_void_p()
        _ck(lib.meso_ipc_open(self.h, _p(h), C.byref(p)))
        return p.value

    def ipc_close(self, peer_dptr):
        _ck(lib.meso_ipc_close(self.h, C.c_void_p(peer_dptr)))

    def download_async(self, host_array, dptr, nbytes=None):
        """device -> pinned host, enqueued on the context's stream (no wait)."""
        n = host_array.nbytes if nbytes is None else nbytes
        _ck(lib.meso_download_async(self.h, _p(host_array), C.c_void_p(dptr), C.c_size_t(n)))

    def download(self, host_array, dptr, nbytes=None):
        """device -> host (numpy array or raw host address), synchronous on the context's stream."""
        n = host_array.nbytes if nbytes is None else nbytes
        _ck(lib.meso_download(self.h, _p(host_array), C.c_void_p(dptr), C.c_size_t(n)))


class Group:
    """N GPUs of one box behind one handle (meso_group_*): member r renders tiles t % N == r and meshes chunks c % N == r of
    a replicated volume.  `devices` may name the same GPU more than once."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _ck(lib.meso_group_create(devs, C.c_int(len(devices)), C.byref(h)))
        self.h = h
        self.n = len(devices)

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib.meso_group_destroy(self.h)
        self.h = None

    def __del__(self):
        self.close()

    def member(self, rank):
        """A Context view of member `rank` (owned by the group: do not close it)."""
        c = Context.__new__(Context)
        c.h = None
        p = lib.meso_group_ctx(self.h, C.c_int(rank))
        if not p:
            raise MesoError("meso_group_ctx: rank out of range")
        c.h = C.c_void_p(p)
        c.origin, c.dims, c.nchunks = self.origin, self.dims, self.nchunks
        c.close = lambda: None
        return c

    def sync(self):
        _ck(lib.meso_group_sync(self.h))

    def scene_create(self, origin_chunk, dims_chunks, max_bricks, cfg=None):
        cfg = default_scene_config() if cfg is None else cfg
        o = np.ascontiguousarray(origin_chunk, dtype=np.int32)
        d = np.ascontiguousarray(dims_chunks, dtype=np.int32)
        _ck(lib.meso_group_scene_create(self.h, _p(cfg), _p(o), _p(d), C.c_uint32(max_bricks)))
        self.origin, self.dims, self.nchunks = tuple(int(x) for x in o), tuple(int(x) for x in d), int(np.prod(d))

    def voxelize_sdf(self, kind, params, granularity):
        p = np.ascontiguousarray(params if params is not None else [0, 0, 0, 0], dtype=np.float64)
        _ck(lib.meso_group_voxelize_sdf(self.h, C.c_int(kind), _p(p), C.c_int(granularity)))

    def carve_sphere(self, center, radius):
        c = np.ascontiguousarray(center, dtype=np.int32)
        n = C.c_int64(0)
        _ck(lib.meso_group_carve_sphere(self.h, _p(c), C.c_int32(radius), C.byref(n)))
        return n.value

    def _flags(self, shadow, rgba8, cubes):
        return (FLAG_SHADOW if shadow else 0) | (FLAG_RGBA8 if rgba8 else 0) | Context._cubes_flag(cubes)

    def raymarch(self, cam, width, height, shadow=True, light=(0.3, 0.5, 0.8), out=None, rgba8=False, cubes=None):
        rec = out if out is not None else (np.zeros((height, width), dtype=np.uint32) if rgba8 else np.zeros((height, width), dtype=HitRecord))
        l = np.ascontiguousarray(light, dtype=np.float32)
        _ck(lib.meso_group_raymarch(self.h, _p(cam), C.c_int(width), C.c_int(height), C.c_uint32(self._flags(shadow, rgba8, cubes)), _p(l), _p(rec)))
        return rec

    def raymarch_async(self, cam, width, height, out, slot, shadow=True, light=(0.3, 0.5, 0.8), rgba8=False, cubes=None):
        l = np.ascontiguousarray(light, dtype=np.float32)
        _ck(lib.meso_group_raymarch_async(self.h, _p(cam), C.c_int(width), C.c_int(height), C.c_uint32(self._flags(shadow, rgba8, cubes)), _p(l), _p(out),
                                          C.c_int(slot)))

    def frame_wait(self, slot):
        _ck(lib.meso_group_frame_wait(self.h, C.c_int(slot)))

    def mesh(self, cap, out=None):
        """-> (quads concatenated in member order, per-member counts); out: optional (pinned) Quad buffer of >= cap records"""
        q = out if out is not None else np.zeros(max(cap, 1), dtype=Quad)
        n = C.c_int64(0)
        counts = np.zeros(self.n, dtype=np.int64)
        _ck(lib.meso_group_mesh(self.h, _p(q), C.c_int64(cap), C.byref(n), _p(counts)))
        return q[: n.value], counts

    def mesh_device(self, d_quads, cap, compact=True):
        n = C.c_int64(0)
        counts = np.zeros(self.n, dtype=np.int64)
        _ck(lib.meso_group_mesh_device(self.h, C.c_void_p(int(d_quads)), C.c_int64(cap), C.byref(n), _p(counts), C.c_int(1 if compact else 0)))
        return n.value, counts

    def remesh_dirty(self, cap):
        q = np.zeros(max(cap, 1), dtype=Quad)
        n = C.c_int64(0)
        _ck(lib.meso_group_remesh_dirty(self.h, _p(q), C.c_int64(cap), C.byref(n)))
        return q[: n.value]
