// meso_capi.cu -- the C ABI of include/meso_cuda.h: context, device memory, stream plumbing, host-side ray setup.
// No CPU fallback anywhere: every compute entry point needs a live CUDA device and fails loudly otherwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "meso_ctx.cuh"

static thread_local std::string g_err;

int meso_fail(int code, const std::string& msg) { g_err = msg; return code; }
static int fail(int code, const std::string& msg) { return meso_fail(code, msg); }

extern "C" {

const char* meso_last_error(void) { return g_err.c_str(); }
int meso_abi_version(void) { return 1; }

static int ctx_init(MesoCtx* c);

int meso_ctx_create(int device, MesoCtx** out) {
  if (!out) return fail(MESO_ERR_ARGUMENT, "meso_ctx_create: out is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(MESO_ERR_RUNTIME, std::string("meso_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
  if (device < 0 || device >= n) return fail(MESO_ERR_ARGUMENT, "meso_ctx_create: device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MESO_ERR_RUNTIME, std::string("meso_ctx_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                      std::to_string(prop.minor) + "; libmeso_b200 carries sm_100a code only");
  MesoCtx* c = new MesoCtx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  const int r = ctx_init(c);
  if (r != MESO_OK) { meso_ctx_destroy(c); return r; }   // g_err keeps the failing call's message
  *out = c;
  return MESO_OK;
}

static int ctx_init(MesoCtx* c) {
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaMalloc(&c->d_tmp_count, 16));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 16; i++) CK(cudaEventCreateWithFlags(&c->band_done[i], cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) CK(cudaStreamCreateWithFlags(&c->band_stream[i], cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->band_fork, cudaEventDisableTiming));
  for (int i = 0; i < MESO_FRAME_RING; i++) {
    CK(cudaEventCreateWithFlags(&c->ring_traced[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ring_copied[i], cudaEventDisableTiming));
  }
  CK(cudaMalloc(&c->d_overflow, sizeof(int)));
  CK(cudaMemset(c->d_overflow, 0, sizeof(int)));
  CK(cudaMalloc(&c->d_timeout, sizeof(int))); CK(cudaMemset(c->d_timeout, 0, sizeof(int)));
  CK(cudaHostAlloc(&c->h_stage, MESO_STAGE_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
  CK(cudaHostGetDevicePointer((void**)&c->d_stage, c->h_stage, 0));
  return MESO_OK;
}

int meso_small_read(MesoCtx* c, void* host_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return MESO_OK;
  CK(cudaSetDevice(c->device));   // a kernel launch needs the stream's device current (a group drives several from one thread)
  if (bytes > MESO_STAGE_BYTES || ((uintptr_t)d_src & 3)) {
    CK(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MESO_OK;
  }
  launch_peek(c->lc(), c->d_stage, d_src, bytes);
  CK_LAST("small read");
  CK(cudaStreamSynchronize(c->stream));
  memcpy(host_dst, c->h_stage, bytes);
  return MESO_OK;
}

static void free_scene(MesoCtx* c) {
  DVolume& v = c->v;
  cudaFree(v.occ); cudaFree(v.full); cudaFree(v.of); cudaFree(v.cells); cudaFree(v.region_any); cudaFree(v.mips); cudaFree(v.bptr); cudaFree(v.pool);
  cudaFree(v.chunk_any); cudaFree(v.chunk_full); cudaFree(v.pool_count); cudaFree(v.pool_free); cudaFree(v.pool_free_count); cudaFree(c->d_shift_scratch); c->d_shift_scratch = nullptr; cudaFree(v.df); cudaFree(v.df_tmp); cudaFree(v.pool_cm); cudaFree(v.words); cudaFree(v.n_words);
  cudaFree(c->d_table); cudaFree(c->d_counts); cudaFree(c->d_offsets); cudaFree(c->d_total); cudaFree(c->d_block_totals); cudaFree(c->d_inst);
  c->d_block_totals = nullptr;
  cudaFree(c->d_frame); cudaFree(c->d_stats); cudaFree(c->d_touch_chunk); cudaFree(c->d_touch_brick);
  cudaFree(c->d_work); cudaFree(c->d_work_count); cudaFree(c->d_quad_count); cudaFree(c->d_quads);
  cudaFree(c->d_dirty); cudaFree(c->d_dirty_count); cudaFree(c->d_keys); cudaFree(c->d_keys_count); cudaFree(c->d_mark);
  cudaFree(c->d_chunk_mark); cudaFree(c->d_chunk_list); c->d_chunk_mark = nullptr; c->d_chunk_list = nullptr; c->d_chunk_count = nullptr;
  cudaFree(c->d_loaded); cudaFree(c->d_stream_list); cudaFree(c->d_stream_stats);
  cudaFree(c->d_cube_cell); cudaFree(c->d_cube_cellp); cudaFree(c->d_cube_brick); cudaFree(c->d_cube_cell2);
  c->d_cube_cell = nullptr; c->d_cube_cellp = nullptr; c->d_cube_brick = nullptr; c->d_cube_cell2 = nullptr; c->cubes = CubeTables{}; c->cubes_valid = false;
  c->d_loaded = nullptr; c->d_stream_list = nullptr; c->stream_list_cap = 0; c->d_stream_stats = nullptr; c->streaming = false;
  v = DVolume{};
  c->d_table = nullptr; c->d_counts = c->d_offsets = nullptr; c->d_total = nullptr; c->d_inst = nullptr;
  c->d_frame = nullptr; c->frame_px = 0; c->d_stats = nullptr; c->d_touch_chunk = c->d_touch_brick = nullptr;
  c->d_work = nullptr; c->d_work_count = nullptr; c->d_quad_count = nullptr; c->d_quads = nullptr; c->cap_quads = 0;
  c->d_dirty = nullptr; c->d_dirty_count = nullptr; c->d_keys = nullptr; c->d_keys_count = nullptr; c->d_mark = nullptr;
  c->has_scene = false;
}

int meso_ctx_destroy(MesoCtx* c) {
  if (!c) return MESO_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_scene(c);
  cudaFree(c->d_flush); cudaFree(c->d_tmp_count); cudaFree(c->d_overflow);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  cudaFree(c->d_timeout);
  cudaFree(c->d_sel_keys); cudaFree(c->d_sel_out); cudaFree(c->d_sel_count);
  for (int i = 0; i < 16; i++) if (c->band_done[i]) cudaEventDestroy(c->band_done[i]);
  cudaStreamDestroy(c->copy_stream);
  for (int i = 0; i < 2; i++) if (c->band_stream[i]) cudaStreamDestroy(c->band_stream[i]);
  if (c->band_fork) cudaEventDestroy(c->band_fork);
  for (int i = 0; i < MESO_FRAME_RING; i++) {
    cudaFree(c->d_ring[i]);
    if (c->ring_traced[i]) cudaEventDestroy(c->ring_traced[i]);
    if (c->ring_copied[i]) cudaEventDestroy(c->ring_copied[i]);
  }
  cudaStreamDestroy(c->own_stream);
  delete c;
  return MESO_OK;
}

int meso_ctx_set_stream(MesoCtx* c, void* s) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  c->stream = (cudaStream_t)s;  // NULL is the CUDA (legacy) default stream, a perfectly valid choice
  return MESO_OK;
}
int meso_ctx_use_own_stream(MesoCtx* c) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  c->stream = c->own_stream;
  return MESO_OK;
}

int meso_ctx_sync(MesoCtx* c) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}

int meso_ctx_set_partition(MesoCtx* c, int rank, int world) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  if (world < 1 || rank < 0 || rank >= world) return fail(MESO_ERR_ARGUMENT, "meso_ctx_set_partition: need 0 <= rank < world");
  c->rank = rank; c->world = world;
  return MESO_OK;
}

int meso_device_sm_count(MesoCtx* c) { return c ? c->sm_count : 0; }
int64_t meso_launch_count(MesoCtx* c) { return c ? c->launches : 0; }

static int scene_alloc(MesoCtx* c, const MesoGPUUniformSceneConfig* cfg, const int32_t origin[3], const int32_t dims[3], uint32_t max_bricks);

int meso_scene_create(MesoCtx* c, const MesoGPUUniformSceneConfig* cfg, const int32_t origin[3], const int32_t dims[3], uint32_t max_bricks) {
  const int r = scene_alloc(c, cfg, origin, dims, max_bricks);
  // an allocation that failed half-way (out of memory at a large grid) leaves no scene and no device memory behind
  if (r != MESO_OK && c && !c->has_scene) free_scene(c);
  return r;
}

static int scene_alloc(MesoCtx* c, const MesoGPUUniformSceneConfig* cfg, const int32_t origin[3], const int32_t dims[3], uint32_t max_bricks) {
  if (!c || !cfg || !origin || !dims) return fail(MESO_ERR_ARGUMENT, "meso_scene_create: null argument");
  // the kernels are specialised for the reference's constants (VoxelSceneConfig.h:22-24)
  if (cfg->BlockResolution != 8 || cfg->ChunkResolution != 16 || cfg->BlockSize != 1.0f || cfg->ChunkSize != 16.0f)
    return fail(MESO_ERR_ARGUMENT, "meso_scene_create: only BlockResolution 8, ChunkResolution 16, BlockSize 1, ChunkSize 16 are supported");
  for (int i = 0; i < 3; i++)
    if (dims[i] < 1 || dims[i] > 512) return fail(MESO_ERR_ARGUMENT, "meso_scene_create: dims_chunks out of range [1,512]");
  if (max_bricks == 0) return fail(MESO_ERR_ARGUMENT, "meso_scene_create: max_bricks must be > 0");
  // 32-bit word / cell / brick indices inside the kernels (64 words, 64 cells, 4096 bricks per chunk): 2^20 chunks = a 12 800^3-voxel window
  if ((int64_t)dims[0] * dims[1] * dims[2] > (1ll << 20)) return fail(MESO_ERR_ARGUMENT, "meso_scene_create: more than 2^20 chunks");
  // 32-bit brick / 2^3-cell indices in the raymarch (chunk * 4096 + brick, slot * 64 + cell)
  if (max_bricks > (1u << 26)) return fail(MESO_ERR_ARGUMENT, "meso_scene_create: max_bricks above 2^26");
  CK(cudaSetDevice(c->device));
  if (c->has_scene) { cudaStreamSynchronize(c->stream); free_scene(c); }
  c->cfg = *cfg;
  DVolume& v = c->v;
  for (int i = 0; i < 3; i++) { v.dims[i] = dims[i]; v.nvox[i] = dims[i] * MESO_CV; v.origin[i] = origin[i]; }
  v.nchunks = (int64_t)dims[0] * dims[1] * dims[2];
  v.max_bricks = max_bricks;
  v.chunk_words = (int)((v.nchunks + 31) / 32);
  const size_t nc = (size_t)v.nchunks;
  CK(cudaMalloc(&v.occ, nc * 64 * 8)); CK(cudaMalloc(&v.full, nc * 64 * 8)); CK(cudaMalloc(&v.mips, nc * 3 * 64 * 8));
  CK(cudaMalloc(&v.of, nc * 64 * 16)); CK(cudaMemsetAsync(v.of, 0, nc * 64 * 16, c->stream));
  for (int i = 0; i < 3; i++) v.rdims[i] = (dims[i] + 3) / 4;
  v.region_words = (v.rdims[0] * v.rdims[1] * v.rdims[2] + 31) / 32;
  CK(cudaMalloc(&v.cells, nc * 8)); CK(cudaMemsetAsync(v.cells, 0, nc * 8, c->stream));
  for (int i = 0; i < 3; i++) v.ddims[i] = dims[i] * 4;
  CK(cudaMalloc(&v.df, nc * 64)); CK(cudaMalloc(&v.df_tmp, nc * 64));
  CK(cudaMemsetAsync(v.df, MESO_DF_K + 1, nc * 64, c->stream));   // empty volume: nothing within reach anywhere
  CK(cudaMalloc(&v.region_any, (size_t)v.region_words * 4)); CK(cudaMemsetAsync(v.region_any, 0, (size_t)v.region_words * 4, c->stream));
  CK(cudaMalloc(&v.bptr, nc * MESO_BLOCKS * 4));
  CK(cudaMalloc(&v.pool, (size_t)max_bricks * 64));
  CK(cudaMalloc(&v.pool_cm, (size_t)max_bricks * 8));
  CK(cudaMalloc(&v.words, nc * 64 * 4)); CK(cudaMalloc(&v.n_words, 4));
  CK(cudaMalloc(&v.chunk_any, (size_t)v.chunk_words * 4)); CK(cudaMalloc(&v.chunk_full, (size_t)v.chunk_words * 4));
  CK(cudaMalloc(&v.pool_count, 4));
  CK(cudaMalloc(&v.pool_free, (size_t)max_bricks * 4)); CK(cudaMalloc(&v.pool_free_count, 4));
  CK(cudaMemsetAsync(v.pool_free_count, 0, 4, c->stream));
  CK(cudaMemsetAsync(v.occ, 0, nc * 64 * 8, c->stream)); CK(cudaMemsetAsync(v.full, 0, nc * 64 * 8, c->stream));
  CK(cudaMemsetAsync(v.mips, 0, nc * 3 * 64 * 8, c->stream));
  CK(cudaMemsetAsync(v.bptr, 0xFF, nc * MESO_BLOCKS * 4, c->stream));
  CK(cudaMemsetAsync(v.chunk_any, 0, (size_t)v.chunk_words * 4, c->stream));
  CK(cudaMemsetAsync(v.chunk_full, 0, (size_t)v.chunk_words * 4, c->stream));
  CK(cudaMemsetAsync(v.pool_count, 0, 4, c->stream));
  CK(cudaMalloc(&c->d_table, nc * sizeof(MesoGPUChunk)));
  CK(cudaMalloc(&c->d_counts, nc * 4)); CK(cudaMalloc(&c->d_offsets, nc * 4)); CK(cudaMalloc(&c->d_total, 8));
  CK(cudaMalloc(&c->d_block_totals, 1024 * 4));
  CK(cudaMalloc(&c->d_stats, sizeof(RayStatsDev)));
  CK(cudaMalloc(&c->d_touch_chunk, nc)); CK(cudaMalloc(&c->d_touch_brick, max_bricks));
  CK(cudaMalloc(&c->d_work_count, 16)); c->d_chunk_count = c->d_work_count + 1;   // work, chunk and full-brick counters are adjacent: one memset clears them
  CK(cudaMalloc(&c->d_quad_count, 8));
  c->cap_dirty = 1u << 22;
  CK(cudaMalloc(&c->d_dirty, (size_t)c->cap_dirty * 8)); CK(cudaMalloc(&c->d_dirty_count, 4));
  CK(cudaMalloc(&c->d_keys, (size_t)c->cap_dirty * 8)); CK(cudaMalloc(&c->d_keys_count, 4));
  const size_t mark_words = (nc * MESO_BLOCKS + 31) / 32;
  CK(cudaMalloc(&c->d_mark, mark_words * 4)); CK(cudaMemsetAsync(c->d_mark, 0, mark_words * 4, c->stream));
  CK(cudaMalloc(&c->d_chunk_mark, ((nc + 31) / 32) * 4)); CK(cudaMemsetAsync(c->d_chunk_mark, 0, ((nc + 31) / 32) * 4, c->stream));
  CK(cudaMalloc(&c->d_chunk_list, nc * 4));
  c->cap_inst = 0; c->n_inst = 0; c->n_dirty = 0;
  c->has_scene = true;
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}

// Frames started with meso_raymarch_async are traced on the band streams and may still be reading the volume when the
// call returns.  Every entry point that rewrites the volume in place orders its work on the context's stream behind the
// traversal (not the host copy) of the frames still in flight; with no frame in flight this does nothing.
int meso_join_frames(MesoCtx* c) {
  for (int i = 0; i < MESO_FRAME_RING; i++)
    if (c->ring_busy[i]) CK(cudaStreamWaitEvent(c->stream, c->ring_traced[i], 0));
  return MESO_OK;
}
#define JOIN_FRAMES(c) do { const int jr__ = meso_join_frames(c); if (jr__ != MESO_OK) return jr__; } while (0)

static int build_cubes(MesoCtx* c);

int meso_overflow_finish(MesoCtx* c, const char* what) {
  int h = 0;
  const int rr = meso_small_read(c, &h, c->d_overflow, sizeof(int));
  if (rr != MESO_OK) return rr;
  if (h) {
    cudaMemsetAsync(c->d_overflow, 0, sizeof(int), c->stream);
    return fail(MESO_ERR_RUNTIME, std::string(what) + ": brick payload pool exhausted (raise max_bricks in meso_scene_create)");
  }
  return MESO_OK;
}

static int voxelize_enqueue_lod(MesoCtx* c, int kind, const double params[4], int granularity, int mip);
int meso_voxelize_enqueue(MesoCtx* c, int kind, const double params[4], int granularity) { return voxelize_enqueue_lod(c, kind, params, granularity, 0); }
static int voxelize_enqueue_lod(MesoCtx* c, int kind, const double params[4], int granularity, int mip) {
  NEED_SCENE(c);
  if (kind != MESO_SDF_SPHERE && kind != MESO_SDF_TERRAIN) return fail(MESO_ERR_ARGUMENT, "meso_voxelize_sdf: unknown sdf kind");
  if (granularity != MESO_GRAN_BLOCK && granularity != MESO_GRAN_VOXEL) return fail(MESO_ERR_ARGUMENT, "meso_voxelize_sdf: unknown granularity");
  if (kind == MESO_SDF_SPHERE && !params) return fail(MESO_ERR_ARGUMENT, "meso_voxelize_sdf: sphere needs params");
  c->streaming = false;  // the whole grid is regenerated: a stream in progress ends (meso_stream_begin starts a new one)
  c->cubes_valid = false;
  JOIN_FRAMES(c);
  launch_voxelize(c->lc(), c->v, kind, params, granularity, c->d_overflow, mip);
  CK_LAST("voxelize");
  return build_cubes(c);   // derived data of a whole-grid generation, like the distance field (enqueue only)
}

int meso_voxelize_sdf(MesoCtx* c, int kind, const double params[4], int granularity) {
  const int r = meso_voxelize_enqueue(c, kind, params, granularity);
  if (r != MESO_OK) return r;
  const int ro = meso_overflow_finish(c, "meso_voxelize_sdf");
  if (ro != MESO_OK) c->cubes_valid = false;   // bricks were dropped: whatever the tables say is about another volume
  return ro;
}

int meso_voxelize_sdf_lod(MesoCtx* c, int kind, const double params[4], uint32_t mipmap_level) {
  if (mipmap_level > 4) return fail(MESO_ERR_ARGUMENT, "meso_voxelize_sdf_lod: MipmapLevel out of range [0,4] (16 blocks per chunk axis)");
  const int r = voxelize_enqueue_lod(c, kind, params, MESO_GRAN_BLOCK, (int)mipmap_level);
  if (r != MESO_OK) return r;
  return meso_overflow_finish(c, "meso_voxelize_sdf_lod");
}

int meso_volume_upload(MesoCtx* c, const uint64_t* occ, const uint64_t* full, const uint64_t* keys, const uint64_t* payload, int64_t n) {
  NEED_SCENE(c);
  if (!occ || !full || n < 0 || (n > 0 && (!keys || !payload))) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: null argument");
  if ((uint64_t)n > c->v.max_bricks) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: more partial bricks than max_bricks");
  const size_t nc = (size_t)c->v.nchunks;
  // every occ && !full brick needs exactly one key (its payload), and a key may only name such a brick: a partial brick
  // without a payload slot would be dereferenced out of bounds by the raymarch, mesh and gather kernels
  int64_t n_need = 0;
  for (size_t i = 0; i < nc * 64; i++) {
    if (full[i] & ~occ[i]) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: a brick is marked full but not occupied");
    n_need += __builtin_popcountll(occ[i] & ~full[i]);
  }
  if (n_need != n) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: n_partial does not match popcount(occ & ~full)");
  for (int64_t i = 0; i < n; i++) {
    if (keys[i] >= (uint64_t)nc * MESO_BLOCKS) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: key out of range");
    if (i > 0 && keys[i] <= keys[i - 1]) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: keys must be strictly ascending");
    const uint64_t w = keys[i] >> 6, b = keys[i] & 63;
    if (!(((occ[w] & ~full[w]) >> b) & 1ull)) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload: key names a brick that is not partial");
  }
  c->cubes_valid = false;
  JOIN_FRAMES(c);
  uint64_t *d_k = nullptr, *d_p = nullptr;
  auto body = [&]() -> int {
    CK(cudaMemcpyAsync(c->v.occ, occ, nc * 64 * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->v.full, full, nc * 64 * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->v.bptr, 0xFF, nc * MESO_BLOCKS * 4, c->stream));
    const uint32_t n32 = (uint32_t)n;
    CK(cudaMemcpyAsync(c->v.pool_count, &n32, 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->v.pool_free_count, 0, 4, c->stream));
    if (n > 0) {
      CK(cudaMalloc(&d_k, (size_t)n * 8)); CK(cudaMalloc(&d_p, (size_t)n * 64));
      CK(cudaMemcpyAsync(d_k, keys, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(d_p, payload, (size_t)n * 64, cudaMemcpyHostToDevice, c->stream));
      launch_scatter_payload(c->lc(), c->v, d_k, d_p, n);
    }
    launch_volume_finalize(c->lc(), c->v);
    CK_LAST("volume upload");
    const int rc = build_cubes(c);
    if (rc != MESO_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return MESO_OK;
  };
  const int r = body();
  cudaFree(d_k); cudaFree(d_p);   // on every path
  return r;
}

int meso_volume_upload_blocks(MesoCtx* c, const MesoGPUChunk* chunks, int64_t n_chunks, const MesoGPUBlock* blocks, int64_t n_blocks,
                              uint32_t flags, int64_t* n_accepted) {
  NEED_SCENE(c);
  if (n_chunks < 0 || n_blocks < 0 || (n_chunks > 0 && !chunks) || (n_blocks > 0 && !blocks)) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload_blocks: bad argument");
  if (flags & ~(uint32_t)MESO_UPLOAD_MERGE) return fail(MESO_ERR_ARGUMENT, "meso_volume_upload_blocks: unknown flag");
  DVolume& v = c->v;
  const size_t nc = (size_t)v.nchunks;
  c->cubes_valid = false;
  c->streaming = false;
  JOIN_FRAMES(c);
  MesoGPUChunk* d_c = nullptr; MesoGPUBlock* d_b = nullptr;
  auto body = [&]() -> int {
    if (!(flags & MESO_UPLOAD_MERGE)) {   // replace: the window is emptied first (the reference re-uploads whole pools)
      CK(cudaMemsetAsync(v.occ, 0, nc * 64 * 8, c->stream)); CK(cudaMemsetAsync(v.full, 0, nc * 64 * 8, c->stream));
      CK(cudaMemsetAsync(v.bptr, 0xFF, nc * MESO_BLOCKS * 4, c->stream));
      CK(cudaMemsetAsync(v.pool_count, 0, 4, c->stream));
      CK(cudaMemsetAsync(v.pool_free_count, 0, 4, c->stream));
    }
    CK(cudaMemsetAsync(c->d_quad_count, 0, 8, c->stream));   // borrowed as the accepted-block counter
    if (n_blocks > 0 && n_chunks > 0) {
      CK(cudaMalloc(&d_c, (size_t)n_chunks * sizeof(MesoGPUChunk))); CK(cudaMalloc(&d_b, (size_t)n_blocks * sizeof(MesoGPUBlock)));
      CK(cudaMemcpyAsync(d_c, chunks, (size_t)n_chunks * sizeof(MesoGPUChunk), cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(d_b, blocks, (size_t)n_blocks * sizeof(MesoGPUBlock), cudaMemcpyHostToDevice, c->stream));
      launch_scatter_blocks(c->lc(), v, d_c, n_chunks, d_b, n_blocks, c->d_quad_count);
    }
    launch_volume_finalize(c->lc(), v);
    CK_LAST("volume upload blocks");
    const int rc = build_cubes(c);
    if (rc != MESO_OK) return rc;
    unsigned long long acc = 0;
    CK(cudaMemcpyAsync(&acc, c->d_quad_count, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (n_accepted) *n_accepted = (int64_t)acc;
    return MESO_OK;
  };
  const int r = body();
  cudaFree(d_c); cudaFree(d_b);
  return r;
}

int meso_volume_num_partial(MesoCtx* c, int64_t* out) {
  NEED_SCENE(c);
  if (!out) return fail(MESO_ERR_ARGUMENT, "null out");
  // live partial bricks (payload slots retired by carving are not counted)
  const size_t nc = (size_t)c->v.nchunks;
  std::vector<uint64_t> occ(nc * 64), full(nc * 64);
  CK(cudaMemcpyAsync(occ.data(), c->v.occ, nc * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(full.data(), c->v.full, nc * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  int64_t n = 0;
  for (size_t i = 0; i < nc * 64; i++) n += __builtin_popcountll(occ[i] & ~full[i]);
  *out = n;
  return MESO_OK;
}

int meso_volume_download(MesoCtx* c, uint64_t* occ, uint64_t* full, uint64_t* keys, uint64_t* payload, int64_t cap, int64_t* n_partial) {
  NEED_SCENE(c);
  if (!occ || !full || !n_partial) return fail(MESO_ERR_ARGUMENT, "meso_volume_download: null argument");
  const size_t nc = (size_t)c->v.nchunks;
  CK(cudaMemcpyAsync(occ, c->v.occ, nc * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(full, c->v.full, nc * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  int64_t n = 0;
  for (size_t i = 0; i < nc * 64; i++) n += __builtin_popcountll(occ[i] & ~full[i]);
  *n_partial = n;
  if (n == 0 || !keys || !payload) return MESO_OK;
  if (cap < n) return fail(MESO_ERR_ARGUMENT, "meso_volume_download: cap_partial too small");
  uint64_t *d_k = nullptr, *d_p = nullptr;
  CK(cudaMalloc(&d_k, (size_t)n * 8)); CK(cudaMalloc(&d_p, (size_t)n * 64));
  launch_gather_partial(c->lc(), c->v, d_k, d_p, c->d_tmp_count);
  CK_LAST("gather partial");
  std::vector<uint64_t> k((size_t)n), p((size_t)n * 8);
  CK(cudaMemcpyAsync(k.data(), d_k, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(p.data(), d_p, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_k); cudaFree(d_p);
  // canonical order: ascending key (the device gathers in atomic order)
  std::vector<int64_t> idx((size_t)n);
  for (int64_t i = 0; i < n; i++) idx[(size_t)i] = i;
  std::sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return k[(size_t)a] < k[(size_t)b]; });
  for (int64_t i = 0; i < n; i++) {
    keys[i] = k[(size_t)idx[(size_t)i]];
    memcpy(payload + i * 8, p.data() + (size_t)idx[(size_t)i] * 8, 64);
  }
  return MESO_OK;
}

int meso_build_occupancy(MesoCtx* c, uint32_t stamp, int64_t* n_instances) {
  NEED_SCENE(c);
  // pass 1 on the device (mips, per-chunk counts, scan); the 8-byte total sizes the instance buffer, then pass 2 emits
  launch_occupancy_count(c->lc(), c->v, stamp, c->d_table, c->d_counts, c->d_offsets, c->d_total, c->d_block_totals);
  CK_LAST("occupancy count");
  uint64_t total = 0;
  { const int rr = meso_small_read(c, &total, c->d_total, 8); if (rr != MESO_OK) return rr; }
  if ((int64_t)total > c->cap_inst) {
    cudaFree(c->d_inst); c->d_inst = nullptr;
    c->cap_inst = (int64_t)total + (int64_t)total / 8;   // a little headroom for edits
    CK(cudaMalloc(&c->d_inst, (size_t)c->cap_inst * sizeof(MesoGPUBlock)));
  }
  launch_occupancy_emit(c->lc(), c->v, stamp, c->d_counts, c->d_offsets, c->d_inst, c->cap_inst);
  CK_LAST("occupancy emit");
  c->n_inst = (int64_t)total;
  if (n_instances) *n_instances = c->n_inst;
  return MESO_OK;
}

int meso_download_chunk_table(MesoCtx* c, MesoGPUChunk* out) {
  NEED_SCENE(c);
  if (!out) return fail(MESO_ERR_ARGUMENT, "null out");
  CK(cudaMemcpyAsync(out, c->d_table, (size_t)c->v.nchunks * sizeof(MesoGPUChunk), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}
int meso_download_mips(MesoCtx* c, uint64_t* out) {
  NEED_SCENE(c);
  if (!out) return fail(MESO_ERR_ARGUMENT, "null out");
  CK(cudaMemcpyAsync(out, c->v.mips, (size_t)c->v.nchunks * 3 * 64 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}
int meso_download_instances(MesoCtx* c, MesoGPUBlock* out, int64_t cap) {
  NEED_SCENE(c);
  if (!out && c->n_inst > 0) return fail(MESO_ERR_ARGUMENT, "null out");
  if (cap < c->n_inst) return fail(MESO_ERR_ARGUMENT, "meso_download_instances: cap too small");
  if (c->n_inst > 0) CK(cudaMemcpyAsync(out, c->d_inst, (size_t)c->n_inst * sizeof(MesoGPUBlock), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}

// Host-side ray setup, fp32, every operation written out; identical sequence to oracle/orc_camera.c:orc_ray_setup
// (this TU is compiled with -fmad=false / host -ffp-contract=off, see build.py).
int meso_ray_setup(const MesoGPUUniformCamera* cam, const int32_t origin_chunk[3], int width, int height, const float light_dir[3], MesoRaySetup* rs) {
  if (!cam || !origin_chunk || !rs || width <= 0 || height <= 0) return fail(MESO_ERR_ARGUMENT, "meso_ray_setup: bad argument");
  const float* V = cam->View; const float* P = cam->Projection;
  memset(rs, 0, sizeof(*rs));
  volatile float t0 = V[12], t1 = V[13], t2 = V[14];
  for (int i = 0; i < 3; i++) {
    volatile float m0 = V[i * 4 + 0] * t0; volatile float m1 = V[i * 4 + 1] * t1; volatile float m2 = V[i * 4 + 2] * t2;
    volatile float s01 = m0 + m1; volatile float s = s01 + m2;
    volatile float e = -s;
    volatile float off = (float)((cam->CameraChunkLocation[i] - origin_chunk[i]) * MESO_CV);
    volatile float e8 = e * 8.0f;
    rs->o[i] = e8 + off;
    rs->U[i] = V[i * 4 + 0] / P[0];
    rs->V[i] = V[i * 4 + 1] / P[5];
    rs->F[i] = -V[i * 4 + 2];
  }
  rs->two_over_w = 2.0f / (float)width;
  rs->two_over_h = 2.0f / (float)height;
  float lx = light_dir ? light_dir[0] : 0.3f, ly = light_dir ? light_dir[1] : 0.5f, lz = light_dir ? light_dir[2] : 0.8f;
  volatile float xx = lx * lx; volatile float yy = ly * ly; volatile float zz = lz * lz;
  volatile float sxy = xx + yy; volatile float sl = sxy + zz;
  volatile float inv = 1.0f / sqrtf(sl);
  rs->L[0] = lx * inv; rs->L[1] = ly * inv; rs->L[2] = lz * inv;
  return MESO_OK;
}

// Forward-cube tables for the current volume, enqueued on the context's stream.  level: 1 = cell cubes only, 2 = + brick
// cubes (default: fastest measured), 3 = + 2^3-cell cubes (MESO_CUBES_LEVEL, a measurement knob).
static int cubes_level() {
  static const int level = [] { const char* e = getenv("MESO_CUBES_LEVEL"); const int l = e ? atoi(e) : 3; return l < 1 ? 1 : (l > 3 ? 3 : l); }();
  return level;
}
static int build_cubes(MesoCtx* c) {
  const DVolume& v = c->v;
  const int level = cubes_level();
  const size_t ncells = (size_t)v.ddims[0] * v.ddims[1] * v.ddims[2];
  const size_t npcells = (size_t)(v.ddims[0] + 2) * (v.ddims[1] + 2) * (v.ddims[2] + 2);
  if (!c->d_cube_cell) CK(cudaMalloc(&c->d_cube_cell, 8 * ncells));
  if (!c->d_cube_cellp) CK(cudaMalloc(&c->d_cube_cellp, 8 * npcells));
  if (level >= 2 && !c->d_cube_brick) CK(cudaMalloc(&c->d_cube_brick, (size_t)v.nchunks * MESO_BLOCKS * sizeof(uint16_t)));
  if (level >= 3 && !c->d_cube_cell2) CK(cudaMalloc(&c->d_cube_cell2, (size_t)v.max_bricks * 64 * sizeof(uint16_t)));
  for (int i = 0; i < MESO_FRAME_RING; i++)     // frames in flight may be reading the old tables
    if (c->ring_busy[i]) CK(cudaStreamWaitEvent(c->stream, c->ring_traced[i], 0));
  // Entries of payload slots that do not exist yet read as "one cell": a carve only REMOVES voxels, so every cube the
  // tables certify stays empty afterwards, and the slots it allocates (full bricks that became partial) are covered by
  // this zero fill -- the tables survive carves (like the distance field) and are rebuilt only when voxels may be added.
  if (level >= 3) CK(cudaMemsetAsync(c->d_cube_cell2, 0, (size_t)v.max_bricks * 64 * sizeof(uint16_t), c->stream));
  launch_build_cubes(c->lc(), v, c->d_cube_cell, c->d_cube_cellp, level >= 2 ? c->d_cube_brick : nullptr, level >= 3 ? c->d_cube_cell2 : nullptr);
  CK_LAST("build cubes");
  c->cubes.cell = c->d_cube_cell; c->cubes.cellp = c->d_cube_cellp; c->cubes.pd0 = v.ddims[0] + 2; c->cubes.pd01 = (v.ddims[0] + 2) * (v.ddims[1] + 2);
  c->cubes.npcells = (int64_t)npcells; c->cubes.ncells = (int64_t)ncells;
  c->cubes.brick = level >= 2 ? c->d_cube_brick : nullptr;
  c->cubes.cell2 = level >= 3 ? c->d_cube_cell2 : nullptr;
  c->cubes_valid = true;
  return MESO_OK;
}

int meso_build_cubes(MesoCtx* c) {
  NEED_SCENE(c);
  return build_cubes(c);
}

int meso_download_cubes(MesoCtx* c, uint8_t* cell, uint16_t* brick) {
  NEED_SCENE(c);
  if (!c->cubes_valid) return fail(MESO_ERR_ARGUMENT, "meso_download_cubes: call meso_build_cubes first");
  if (cell) CK(cudaMemcpyAsync(cell, c->d_cube_cell, 8 * (size_t)c->cubes.ncells, cudaMemcpyDeviceToHost, c->stream));
  if (brick && !c->cubes.brick) return fail(MESO_ERR_ARGUMENT, "meso_download_cubes: the brick table is not built at MESO_CUBES_LEVEL=1");
  if (brick) CK(cudaMemcpyAsync(brick, c->d_cube_brick, (size_t)c->v.nchunks * MESO_BLOCKS * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}

// The tables a raymarch launch reads: the forward cubes whenever they are current (the default walk), nullptr -> the
// distance-field walk when they are not (streaming updates) or MESO_FLAG_NO_CUBES asks for it; MESO_FLAG_CUBES insists.
int meso_cubes_for(MesoCtx* c, uint32_t flags, const CubeTables** out) {
  *out = nullptr;
  if ((flags & MESO_FLAG_CUBES) && (flags & MESO_FLAG_NO_CUBES)) return fail(MESO_ERR_ARGUMENT, "MESO_FLAG_CUBES and MESO_FLAG_NO_CUBES exclude each other");
  if (flags & MESO_FLAG_NO_CUBES) return MESO_OK;
  if (c->cubes_valid) { *out = &c->cubes; return MESO_OK; }
  if (flags & MESO_FLAG_CUBES) return fail(MESO_ERR_ARGUMENT, "MESO_FLAG_CUBES: the forward-cube tables are not current (meso_build_cubes)");
  return MESO_OK;
}

static int ensure_frame(MesoCtx* c, size_t px) {
  if (c->frame_px >= px) return MESO_OK;
  cudaFree(c->d_frame); c->d_frame = nullptr; c->frame_px = 0;
  CK(cudaMalloc(&c->d_frame, px * sizeof(MesoHitRecord)));
  c->frame_px = px;
  return MESO_OK;
}

int meso_raymarch_device(MesoCtx* c, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3], void* d_records, int layout) {
  NEED_SCENE(c);
  if (!cam || !d_records || width <= 0 || height <= 0 || width > 65536 || height > 65536) return fail(MESO_ERR_ARGUMENT, "meso_raymarch: bad argument");
  if (layout != MESO_LAYOUT_FRAME && layout != MESO_LAYOUT_TILES) return fail(MESO_ERR_ARGUMENT, "meso_raymarch: bad layout");
  if ((flags & MESO_FLAG_RGBA8) && layout != MESO_LAYOUT_FRAME) return fail(MESO_ERR_ARGUMENT, "meso_raymarch: MESO_FLAG_RGBA8 needs MESO_LAYOUT_FRAME");
  MesoRaySetup rs;
  int r = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
  if (r != MESO_OK) return r;
  const CubeTables* cubes = nullptr;
  r = meso_cubes_for(c, flags, &cubes);
  if (r != MESO_OK) return r;
  launch_raymarch(c->lc(), c->v, rs, width, height, flags, c->rank, c->world, layout, (MesoHitRecord*)d_records, nullptr, nullptr, nullptr,
                  0, -1, cubes);
  CK_LAST("raymarch");
  return MESO_OK;
}

int meso_signal_device(MesoCtx* c, void* const* d_words, int n) {
  if (!c || !d_words || n < 1 || n > MESO_MAX_SLABS) return fail(MESO_ERR_ARGUMENT, "meso_signal_device: bad argument");
  CK(cudaSetDevice(c->device));
  SignalTargets t{};
  for (int i = 0; i < n; i++) {
    if (!d_words[i]) return fail(MESO_ERR_ARGUMENT, "meso_signal_device: null word");
    t.word[i] = (unsigned*)d_words[i];
  }
  t.n = n;
  launch_signal(c->lc(), t);
  CK_LAST("signal");
  return MESO_OK;
}
int meso_wait_device(MesoCtx* c, const void* d_word, uint32_t target) {
  if (!c || !d_word) return fail(MESO_ERR_ARGUMENT, "meso_wait_device: bad argument");
  CK(cudaSetDevice(c->device));
  launch_wait(c->lc(), (const unsigned*)d_word, target, c->d_timeout);
  CK_LAST("wait");
  return MESO_OK;
}
int meso_wait_timed_out(MesoCtx* c, int* out) {
  if (!c || !out) return fail(MESO_ERR_ARGUMENT, "meso_wait_timed_out: bad argument");
  int h = 0;
  const int r = meso_small_read(c, &h, c->d_timeout, sizeof(int));
  if (r != MESO_OK) return r;
  *out = h;
  if (h) CK(cudaMemsetAsync(c->d_timeout, 0, sizeof(int), c->stream));
  return MESO_OK;
}

int meso_raymarch_device_slabs(MesoCtx* c, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3],
                               void* const* d_slabs, int n_slabs, int rows_per_slab) {
  NEED_SCENE(c);
  if (!cam || !d_slabs || width <= 0 || height <= 0 || width > 65536 || height > 65536) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_device_slabs: bad argument");
  if (n_slabs < 1 || n_slabs > MESO_MAX_SLABS) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_device_slabs: n_slabs out of range [1, MESO_MAX_SLABS]");
  if (rows_per_slab <= 0 || rows_per_slab % MESO_TILE_H != 0) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_device_slabs: rows_per_slab must be a positive multiple of MESO_TILE_H");
  if ((int64_t)rows_per_slab * n_slabs < height) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_device_slabs: the slabs do not cover the frame");
  FrameMap fm{};
  for (int i = 0; i < n_slabs; i++) {
    if (!d_slabs[i]) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_device_slabs: null slab pointer");
    fm.slab[i] = d_slabs[i];
  }
  fm.rows_per_slab = rows_per_slab; fm.n_slabs = n_slabs;
  MesoRaySetup rs;
  int r = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
  if (r != MESO_OK) return r;
  const CubeTables* cubes = nullptr;
  r = meso_cubes_for(c, flags, &cubes);
  if (r != MESO_OK) return r;
  launch_raymarch(c->lc(), c->v, rs, width, height, flags, c->rank, c->world, MESO_LAYOUT_SLABS, nullptr, nullptr, nullptr, nullptr, 0, -1, cubes, &fm);
  CK_LAST("raymarch (slabs)");
  return MESO_OK;
}

int meso_raymarch(MesoCtx* c, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3], MesoHitRecord* host) {
  NEED_SCENE(c);
  if (!host || !cam) return fail(MESO_ERR_ARGUMENT, "meso_raymarch: null argument");
  if (width <= 0 || height <= 0 || width > 65536 || height > 65536) return fail(MESO_ERR_ARGUMENT, "meso_raymarch: bad size");
  const size_t px = (size_t)width * height;
  const size_t bpp = (flags & MESO_FLAG_RGBA8) ? 4 : sizeof(MesoHitRecord);   // bytes per pixel of the output
  char* const host_b = reinterpret_cast<char*>(host);
  int r = ensure_frame(c, px);
  if (r != MESO_OK) return r;
  if (c->world > 1) {
    // other ranks' tiles stay all-ones; single launch + one copy
    CK(cudaMemsetAsync(c->d_frame, 0xFF, px * bpp, c->stream));
    r = meso_raymarch_device(c, cam, width, height, flags, light, c->d_frame, MESO_LAYOUT_FRAME);
    if (r != MESO_OK) return r;
    CK(cudaMemcpyAsync(host, c->d_frame, px * bpp, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MESO_OK;
  }
  // Single GPU: the frame is rendered in bands of whole tile rows; the device-to-host copy of a finished band runs on
  // the copy stream while the next band is being traced, so the PCIe transfer hides behind the kernel (or vice versa).
  MesoRaySetup rs;
  r = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
  if (r != MESO_OK) return r;
  const CubeTables* cubes = nullptr;
  r = meso_cubes_for(c, flags, &cubes);
  if (r != MESO_OK) return r;
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W, tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  int bands = tiles_y >= 64 ? 4 : (tiles_y >= 16 ? 2 : 1);   // measured on B200 at 4K: 1 -> 5.6 ms, 4 -> 4.0 ms, 16 -> 5.7 ms
  if (const char* e = getenv("MESO_E2E_BANDS")) bands = std::max(1, std::min(16, atoi(e)));  // tuning knob
  const int rows_per_band = (tiles_y + bands - 1) / bands;
  // Band kernels alternate between two compute streams so the long tail of one band (a few slow silhouette tiles)
  // overlaps the body of the next; both are ordered after whatever the caller enqueued on the context's stream.
  CK(cudaEventRecord(c->band_fork, c->stream));
  CK(cudaStreamWaitEvent(c->band_stream[0], c->band_fork, 0));
  CK(cudaStreamWaitEvent(c->band_stream[1], c->band_fork, 0));
  for (int b = 0; b < bands; b++) {
    const int ty0 = b * rows_per_band, ty1 = std::min(tiles_y, ty0 + rows_per_band);
    if (ty0 >= ty1) break;
    LaunchCtx lc = c->lc();
    lc.stream = c->band_stream[b & 1];
    launch_raymarch(lc, c->v, rs, width, height, flags, 0, 1, MESO_LAYOUT_FRAME, c->d_frame, nullptr, nullptr, nullptr,
                    ty0 * tiles_x, (ty1 - ty0) * tiles_x, cubes);
    CK_LAST("raymarch band");
    CK(cudaEventRecord(c->band_done[b], lc.stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->band_done[b], 0));
    const size_t y0 = (size_t)ty0 * MESO_TILE_H, y1 = std::min((size_t)height, (size_t)ty1 * MESO_TILE_H);
    CK(cudaMemcpyAsync(host_b + y0 * width * bpp, reinterpret_cast<const char*>(c->d_frame) + y0 * width * bpp, (y1 - y0) * width * bpp,
                       cudaMemcpyDeviceToHost, c->copy_stream));
  }
  CK(cudaStreamSynchronize(c->copy_stream));   // every band kernel precedes its copy, so this covers both band streams
  return MESO_OK;
}

// The present path: the frame as tightly packed RGBA8 for a texture range, i.e. exactly the `data` that
// lvk::IContext::upload(TextureHandle, const TextureRangeDesc&, const void* data[]) (LVK.h:830) takes for TEXOffscreenColor.
int meso_present_rgba8(MesoCtx* c, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3],
                       const MesoTextureRange* range, void* host_pixels) {
  NEED_SCENE(c);
  if (!cam || !host_pixels) return fail(MESO_ERR_ARGUMENT, "meso_present_rgba8: null argument");
  if (width <= 0 || height <= 0 || width > 65536 || height > 65536) return fail(MESO_ERR_ARGUMENT, "meso_present_rgba8: bad size");
  MesoTextureRange full{0u, 0u, (uint32_t)width, (uint32_t)height};
  const MesoTextureRange rg = range ? *range : full;
  if (rg.width == 0 || rg.height == 0 || (uint64_t)rg.x + rg.width > (uint64_t)width || (uint64_t)rg.y + rg.height > (uint64_t)height)
    return fail(MESO_ERR_ARGUMENT, "meso_present_rgba8: range outside the frame");
  const size_t px = (size_t)width * height;
  int r = ensure_frame(c, px);
  if (r != MESO_OK) return r;
  MesoRaySetup rs;
  r = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
  if (r != MESO_OK) return r;
  const CubeTables* cubes = nullptr;
  r = meso_cubes_for(c, flags, &cubes);
  if (r != MESO_OK) return r;
  // only the tile rows the range touches are traced (tiles are numbered row-major: a band of rows is one index range)
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W;
  const int ty0 = (int)rg.y / MESO_TILE_H, ty1 = ((int)(rg.y + rg.height) + MESO_TILE_H - 1) / MESO_TILE_H;
  launch_raymarch(c->lc(), c->v, rs, width, height, (flags | MESO_FLAG_RGBA8) & ~0u, 0, 1, MESO_LAYOUT_FRAME, c->d_frame, nullptr, nullptr, nullptr,
                  ty0 * tiles_x, (ty1 - ty0) * tiles_x, cubes);
  CK_LAST("present");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(c->d_frame) + (size_t)rg.y * width + rg.x;
  CK(cudaMemcpy2DAsync(host_pixels, (size_t)rg.width * 4, src, (size_t)width * 4, (size_t)rg.width * 4, rg.height, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}

int meso_raymarch_async(MesoCtx* c, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3], MesoHitRecord* host, int slot) {
  NEED_SCENE(c);
  if (!host || !cam) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_async: null argument");
  if (width <= 0 || height <= 0 || width > 65536 || height > 65536) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_async: bad size");
  if (slot < 0 || slot >= MESO_FRAME_RING) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_async: slot out of range");
  const size_t px = (size_t)width * height;
  if (c->ring_px < px) {   // (re)allocate the ring
    CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < MESO_FRAME_RING; i++) { cudaFree(c->d_ring[i]); c->d_ring[i] = nullptr; }
    for (int i = 0; i < MESO_FRAME_RING; i++) CK(cudaMalloc(&c->d_ring[i], px * sizeof(MesoHitRecord)));
    c->ring_px = px;
  }
  if (c->ring_busy[slot]) { CK(cudaEventSynchronize(c->ring_copied[slot])); c->ring_busy[slot] = false; }
  MesoRaySetup rs;
  int r = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
  if (r != MESO_OK) return r;
  const CubeTables* cubes = nullptr;
  r = meso_cubes_for(c, flags, &cubes);
  if (r != MESO_OK) return r;
  // Frames of the ring alternate between the two band streams so that the tail of frame k (a few tiles with grazing
  // rays) overlaps the body of frame k+1; each is ordered after whatever the caller enqueued on the context's stream.
  LaunchCtx lc = c->lc();
  lc.stream = c->band_stream[slot & 1];
  CK(cudaEventRecord(c->band_fork, c->stream));
  CK(cudaStreamWaitEvent(lc.stream, c->band_fork, 0));
  const size_t bpp = (flags & MESO_FLAG_RGBA8) ? 4 : sizeof(MesoHitRecord);
  if (c->world > 1) CK(cudaMemsetAsync(c->d_ring[slot], 0xFF, px * bpp, lc.stream));
  launch_raymarch(lc, c->v, rs, width, height, flags, c->rank, c->world, MESO_LAYOUT_FRAME, c->d_ring[slot], nullptr, nullptr, nullptr,
                  0, -1, cubes);
  CK_LAST("raymarch async");
  CK(cudaEventRecord(c->ring_traced[slot], lc.stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->ring_traced[slot], 0));
  CK(cudaMemcpyAsync(host, c->d_ring[slot], px * bpp, cudaMemcpyDeviceToHost, c->copy_stream));
  CK(cudaEventRecord(c->ring_copied[slot], c->copy_stream));
  c->ring_busy[slot] = true;
  return MESO_OK;
}

int meso_frame_wait(MesoCtx* c, int slot) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  if (slot < 0 || slot >= MESO_FRAME_RING) return fail(MESO_ERR_ARGUMENT, "meso_frame_wait: slot out of range");
  CK(cudaSetDevice(c->device));
  if (c->ring_busy[slot]) { CK(cudaEventSynchronize(c->ring_copied[slot])); c->ring_busy[slot] = false; }
  return MESO_OK;
}

int meso_raymarch_stats(MesoCtx* c, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3], MesoRayStats* out) {
  NEED_SCENE(c);
  if (!cam || !out || width <= 0 || height <= 0) return fail(MESO_ERR_ARGUMENT, "meso_raymarch_stats: bad argument");
  const size_t px = (size_t)width * height;
  int r = ensure_frame(c, px);
  if (r != MESO_OK) return r;
  MesoRaySetup rs;
  r = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
  if (r != MESO_OK) return r;
  CK(cudaMemsetAsync(c->d_stats, 0, sizeof(RayStatsDev), c->stream));
  CK(cudaMemsetAsync(c->d_touch_chunk, 0, (size_t)c->v.nchunks, c->stream));
  CK(cudaMemsetAsync(c->d_touch_brick, 0, c->v.max_bricks, c->stream));
  const CubeTables* cubes = nullptr;
  r = meso_cubes_for(c, flags, &cubes);
  if (r != MESO_OK) return r;
  launch_raymarch(c->lc(), c->v, rs, width, height, flags, c->rank, c->world, MESO_LAYOUT_FRAME, c->d_frame, c->d_stats, c->d_touch_chunk, c->d_touch_brick,
                  0, -1, cubes);
  CK_LAST("raymarch stats");
  RayStatsDev h;
  std::vector<uint8_t> tc((size_t)c->v.nchunks), tb(c->v.max_bricks);
  CK(cudaMemcpyAsync(&h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(tc.data(), c->d_touch_chunk, tc.size(), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(tb.data(), c->d_touch_brick, tb.size(), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memset(out, 0, sizeof(*out));
  out->primary = h.primary; out->shadow = h.shadow; out->hits = h.hits; out->steps = h.steps;
  out->steps_primary = h.steps_primary; out->warp_slots_primary = h.warp_slots_primary; out->warp_slots_shadow = h.warp_slots_shadow;
  for (int i = 0; i < 5; i++) out->level_steps[i] = h.level_steps[i];
  for (uint8_t x : tc) out->touched_chunks += x;
  for (uint8_t x : tb) out->touched_bricks += x;
  // DESIGN.md "algorithmic bytes": chunk any/full bit grids once + 1024 B of block masks per touched chunk + 68 B per touched brick
  out->u_bytes = 2 * (uint64_t)((c->v.nchunks + 7) / 8) + 1024 * out->touched_chunks + 68 * out->touched_bricks;
  return MESO_OK;
}

int meso_pack_rgba8_device(MesoCtx* c, const void* d_records, int64_t n, void* d_rgba8) {
  if (!c || n < 0 || (n > 0 && (!d_records || !d_rgba8))) return fail(MESO_ERR_ARGUMENT, "meso_pack_rgba8_device: bad argument");
  CK(cudaSetDevice(c->device));
  launch_pack_rgba8(c->lc(), (const MesoHitRecord*)d_records, (size_t)n, (uint32_t*)d_rgba8);
  CK_LAST("pack rgba8");
  return MESO_OK;
}

int64_t meso_tiles_per_rank(int width, int height, int world) {
  if (width <= 0 || height <= 0 || world <= 0) return 0;
  const int64_t tx = (width + MESO_TILE_W - 1) / MESO_TILE_W, ty = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  return (tx * ty + world - 1) / world;
}

int meso_compose_tiles_device(MesoCtx* c, const void* d_tiles, int world, int width, int height, void* d_frame) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  if (!d_tiles || !d_frame || world < 1 || width <= 0 || height <= 0) return fail(MESO_ERR_ARGUMENT, "meso_compose_tiles_device: bad argument");
  CK(cudaSetDevice(c->device));
  launch_compose_tiles(c->lc(), (const MesoHitRecord*)d_tiles, world, width, height, (MesoHitRecord*)d_frame);
  CK_LAST("compose tiles");
  return MESO_OK;
}

int meso_ensure_mesh_buffers(MesoCtx* c) {
  if (!c->d_work) {
    // worst case: every block of this rank's chunks
    c->cap_work = c->v.nchunks * MESO_BLOCKS;
    CK(cudaMalloc(&c->d_work, (size_t)c->cap_work * 8));
  }
  return MESO_OK;
}

int meso_mesh_device(MesoCtx* c, void* d_quads, int64_t cap, int64_t* n_quads) {
  NEED_SCENE(c);
  if (cap < 0 || (cap > 0 && !d_quads)) return fail(MESO_ERR_ARGUMENT, "meso_mesh_device: bad argument");
  int r = meso_ensure_mesh_buffers(c);
  if (r != MESO_OK) return r;
  launch_mesh(c->lc(), c->v, c->rank, c->world, c->mesh_scratch(), (MesoQuad*)d_quads, cap, c->d_quad_count);
  CK_LAST("mesh");
  if (n_quads) {
    unsigned long long n = 0;
    const int rr = meso_small_read(c, &n, c->d_quad_count, 8);
    if (rr != MESO_OK) return rr;
    *n_quads = (int64_t)n;
  }
  return MESO_OK;
}

int meso_mesh_count_device(MesoCtx* c, void* d_count_out) {
  NEED_SCENE(c);
  if (!d_count_out) return fail(MESO_ERR_ARGUMENT, "meso_mesh_count_device: null argument");
  CK(cudaMemcpyAsync(d_count_out, c->d_quad_count, 8, cudaMemcpyDeviceToDevice, c->stream));
  return MESO_OK;
}

int meso_mesh_device_shared(MesoCtx* c, void* d_quads, void* d_counter, int64_t cap) {
  NEED_SCENE(c);
  if (cap <= 0 || !d_quads || !d_counter) return fail(MESO_ERR_ARGUMENT, "meso_mesh_device_shared: bad argument");
  int r = meso_ensure_mesh_buffers(c);
  if (r != MESO_OK) return r;
  launch_mesh(c->lc(), c->v, c->rank, c->world, c->mesh_scratch(), (MesoQuad*)d_quads, cap, (unsigned long long*)d_counter,
              /*reset_count=*/false);
  CK_LAST("mesh (shared list)");
  return MESO_OK;
}

int meso_mesh(MesoCtx* c, MesoQuad* host, int64_t cap, int64_t* n_quads) {
  NEED_SCENE(c);
  if (!n_quads) return fail(MESO_ERR_ARGUMENT, "meso_mesh: n_quads is null");
  if (cap > c->cap_quads) {
    cudaFree(c->d_quads); c->d_quads = nullptr; c->cap_quads = 0;
    CK(cudaMalloc(&c->d_quads, (size_t)cap * sizeof(MesoQuad)));
    c->cap_quads = cap;
  }
  int r = meso_mesh_device(c, c->d_quads, cap, n_quads);
  if (r != MESO_OK) return r;
  const int64_t n = std::min<int64_t>(*n_quads, cap);
  if (n > 0 && host) {
    CK(cudaMemcpyAsync(host, c->d_quads, (size_t)n * sizeof(MesoQuad), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return MESO_OK;
}

int meso_carve_enqueue(MesoCtx* c, const int32_t center[3], int32_t radius) {
  NEED_SCENE(c);
  if (!center || radius < 0 || radius > 30000) return fail(MESO_ERR_ARGUMENT, "meso_carve_sphere: bad argument");
  JOIN_FRAMES(c);
  launch_carve(c->lc(), c->v, center, radius, c->d_dirty, c->cap_dirty, c->d_dirty_count, c->d_overflow);
  CK_LAST("carve");
  return MESO_OK;
}
int meso_carve_finish(MesoCtx* c, int64_t* n_dirty) {
  NEED_SCENE(c);
  uint32_t n = 0;
  int r = meso_small_read(c, &n, c->d_dirty_count, 4);
  if (r != MESO_OK) return r;
  r = meso_overflow_finish(c, "meso_carve_sphere");
  if (r != MESO_OK) return r;
  if (n > c->cap_dirty) return fail(MESO_ERR_RUNTIME, "meso_carve_sphere: dirty list overflow");
  c->n_dirty = n;
  if (n_dirty) *n_dirty = n;
  return MESO_OK;
}
int meso_carve_sphere(MesoCtx* c, const int32_t center[3], int32_t radius, int64_t* n_dirty) {
  const int r = meso_carve_enqueue(c, center, radius);
  return r != MESO_OK ? r : meso_carve_finish(c, n_dirty);
}

int meso_download_dirty(MesoCtx* c, uint64_t* keys, int64_t cap) {
  NEED_SCENE(c);
  if (cap < (int64_t)c->n_dirty || (!keys && c->n_dirty)) return fail(MESO_ERR_ARGUMENT, "meso_download_dirty: cap too small");
  if (c->n_dirty) CK(cudaMemcpyAsync(keys, c->d_dirty, (size_t)c->n_dirty * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::sort(keys, keys + c->n_dirty);
  return MESO_OK;
}

int meso_remesh_dirty(MesoCtx* c, MesoQuad* host, int64_t cap, int64_t* n_quads, uint64_t* host_keys, int64_t cap_keys, int64_t* n_keys) {
  NEED_SCENE(c);
  if (!n_quads) return fail(MESO_ERR_ARGUMENT, "meso_remesh_dirty: n_quads is null");
  launch_expand_dirty(c->lc(), c->v, c->d_dirty, c->n_dirty, c->d_keys, c->cap_dirty, c->d_keys_count, c->d_mark);
  CK_LAST("expand dirty");
  uint32_t nk = 0;
  { const int rr = meso_small_read(c, &nk, c->d_keys_count, 4); if (rr != MESO_OK) return rr; }
  if (nk > c->cap_dirty) {
    cudaMemsetAsync(c->d_mark, 0, (((size_t)c->v.nchunks * MESO_BLOCKS + 31) / 32) * 4, c->stream);
    return fail(MESO_ERR_RUNTIME, "meso_remesh_dirty: key list overflow");
  }
  if (n_keys) *n_keys = nk;
  if (host_keys) {
    if (cap_keys < (int64_t)nk) return fail(MESO_ERR_ARGUMENT, "meso_remesh_dirty: cap_keys too small");
    if (nk) { const int rr = meso_small_read(c, host_keys, c->d_keys, (size_t)nk * 8); if (rr != MESO_OK) return rr; }
  }
  if (cap > c->cap_quads) {
    cudaFree(c->d_quads); c->d_quads = nullptr; c->cap_quads = 0;
    CK(cudaMalloc(&c->d_quads, (size_t)cap * sizeof(MesoQuad)));
    c->cap_quads = cap;
  }
  // voxel-level quads of the listed bricks + brick-level quads of the chunks that hold them; sharded by key / chunk hash over the partition
  launch_mesh_list(c->lc(), c->v, c->d_keys, nk, c->d_quads, cap, c->d_quad_count, c->mesh_scratch(), c->rank, c->world);
  CK_LAST("remesh");
  unsigned long long n = 0;
  { const int rr = meso_small_read(c, &n, c->d_quad_count, 8); if (rr != MESO_OK) return rr; }
  *n_quads = (int64_t)n;
  const int64_t m = std::min<int64_t>((int64_t)n, cap);
  if (m > 0 && host) return meso_small_read(c, host, c->d_quads, (size_t)m * sizeof(MesoQuad));
  return MESO_OK;
}

// ---- K6 ---------------------------------------------------------------------------------------------------------
static int check_view(const float* fwd, const MesoViewConfig* vc, const char* who) {
  if (!fwd || !vc) return fail(MESO_ERR_ARGUMENT, std::string(who) + ": null argument");
  if (vc->ViewForwardLoadChunkSize < 1 || vc->ViewForwardLoadChunkSize > 200)
    return fail(MESO_ERR_ARGUMENT, std::string(who) + ": ViewForwardLoadChunkSize out of range [1,200]");
  if (vc->Mode > 1) return fail(MESO_ERR_ARGUMENT, std::string(who) + ": unknown Mode");
  return MESO_OK;
}
static int ensure_select_buffers(MesoCtx* c, const MesoViewConfig& vc) {
  const int64_t need = resident_max_candidates(vc);
  if (!c->d_sel_count) CK(cudaMalloc(&c->d_sel_count, 4));
  if (need > c->sel_cap) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_sel_keys); cudaFree(c->d_sel_out); c->d_sel_keys = nullptr; c->d_sel_out = nullptr; c->sel_cap = 0;
    CK(cudaMalloc(&c->d_sel_keys, resident_sort_scratch_bytes(vc)));
    CK(cudaMalloc(&c->d_sel_out, (size_t)need * sizeof(MesoChunkCandidate)));
    c->sel_cap = need;
  }
  return MESO_OK;
}

int meso_select_view_chunks(MesoCtx* c, const float forward[3], const MesoViewConfig* view, MesoChunkCandidate* host, int64_t cap, int64_t* count) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  int r = check_view(forward, view, "meso_select_view_chunks");
  if (r != MESO_OK) return r;
  if (!count || cap < 0 || (cap > 0 && !host)) return fail(MESO_ERR_ARGUMENT, "meso_select_view_chunks: bad output arguments");
  CK(cudaSetDevice(c->device));
  r = ensure_select_buffers(c, *view);
  if (r != MESO_OK) return r;
  launch_select_view(c->lc(), forward, *view, c->d_sel_keys, c->d_sel_count, c->d_sel_out, c->sel_cap);
  CK_LAST("select view chunks");
  uint32_t n = 0;
  CK(cudaMemcpyAsync(&n, c->d_sel_count, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *count = n;
  const int64_t m = std::min<int64_t>(n, cap);
  if (m > 0) {
    CK(cudaMemcpyAsync(host, c->d_sel_out, (size_t)m * sizeof(MesoChunkCandidate), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return MESO_OK;
}

int meso_chunk_importance(MesoCtx* c, const int32_t cam[3], const float forward[3], const int32_t* locations, int64_t n, float* host_out) {
  if (!c || !cam || !forward || n < 0 || (n > 0 && (!locations || !host_out))) return fail(MESO_ERR_ARGUMENT, "meso_chunk_importance: bad argument");
  if (n == 0) return MESO_OK;
  CK(cudaSetDevice(c->device));
  int32_t* d_loc = nullptr; float* d_out = nullptr;
  CK(cudaMalloc(&d_loc, (size_t)n * 12));
  if (cudaMalloc(&d_out, (size_t)n * 4) != cudaSuccess) { cudaFree(d_loc); return fail(MESO_ERR_RUNTIME, "meso_chunk_importance: out of device memory"); }
  cudaMemcpyAsync(d_loc, locations, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream);
  launch_chunk_importance(c->lc(), d_loc, n, cam, forward, d_out);
  cudaMemcpyAsync(host_out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
  const cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(d_loc); cudaFree(d_out);
  if (e != cudaSuccess) return fail(MESO_ERR_RUNTIME, std::string("meso_chunk_importance: ") + cudaGetErrorString(e));
  return MESO_OK;
}

int meso_baked_direction(uint32_t samples, const float forward[3], float out_direction[3], uint32_t* out_index) {
  if (samples < 2 || !forward || !out_direction) return fail(MESO_ERR_ARGUMENT, "meso_baked_direction: bad argument");
  // GetFibonacciSphere<float> (VoxelMathHelper.h:49-71); volatile temporaries keep the host compiler from contracting
  const float phi = (float)(3.14159265358979323846 * (std::sqrt(5.0) - 1.0));
  uint32_t best = 0;
  float bd = INFINITY, bx = 0, by = 0, bz = 0;
  for (uint32_t i = 0; i < samples; i++) {
    volatile float frac = (float)(int32_t)i / (float)(samples - 1);
    volatile float y2 = frac * 2.0f;
    const float y = 1.0f - y2;
    volatile float yy = y * y;
    const float radius = sqrtf(1.0f - yy);
    volatile float theta = phi * (float)(int32_t)i;
    volatile float x = cosf(theta) * radius;
    volatile float z = sinf(theta) * radius;
    volatile float xx = x * x, zz = z * z, y_sq = y * y;
    volatile float s1 = xx + y_sq;
    const float inv = 1.0f / sqrtf(s1 + zz);
    volatile float nx = x * inv, ny = y * inv, nz = z * inv;
    volatile float dx = nx - forward[0], dy = ny - forward[1], dz = nz - forward[2];
    volatile float dxx = dx * dx, dyy = dy * dy, dzz = dz * dz;
    volatile float s2 = dxx + dyy;
    const float dd = s2 + dzz;
    if (dd < bd) { bd = dd; best = i; bx = nx; by = ny; bz = nz; }
  }
  out_direction[0] = bx; out_direction[1] = by; out_direction[2] = bz;
  if (out_index) *out_index = best;
  return MESO_OK;
}

int meso_stream_begin(MesoCtx* c, int kind, const double params[4], int granularity) {
  NEED_SCENE(c);
  if (kind != MESO_SDF_SPHERE && kind != MESO_SDF_TERRAIN) return fail(MESO_ERR_ARGUMENT, "meso_stream_begin: unknown sdf kind");
  if (granularity != MESO_GRAN_BLOCK && granularity != MESO_GRAN_VOXEL) return fail(MESO_ERR_ARGUMENT, "meso_stream_begin: unknown granularity");
  if (kind == MESO_SDF_SPHERE && !params) return fail(MESO_ERR_ARGUMENT, "meso_stream_begin: sphere needs params");
  DVolume& v = c->v;
  const size_t nc = (size_t)v.nchunks;
  if (!c->d_loaded) CK(cudaMalloc(&c->d_loaded, (size_t)v.chunk_words * 4));
  if (!c->d_stream_stats) CK(cudaMalloc(&c->d_stream_stats, 16));
  c->cubes_valid = false;
  JOIN_FRAMES(c);
  // empty volume: nothing is generated yet (FChunkPool::Initialize, ChunkPool.h:283-345)
  CK(cudaMemsetAsync(v.occ, 0, nc * 64 * 8, c->stream)); CK(cudaMemsetAsync(v.full, 0, nc * 64 * 8, c->stream));
  CK(cudaMemsetAsync(v.of, 0, nc * 64 * 16, c->stream)); CK(cudaMemsetAsync(v.cells, 0, nc * 8, c->stream));
  CK(cudaMemsetAsync(v.mips, 0, nc * 3 * 64 * 8, c->stream));
  CK(cudaMemsetAsync(v.bptr, 0xFF, nc * MESO_BLOCKS * 4, c->stream));
  CK(cudaMemsetAsync(v.chunk_any, 0, (size_t)v.chunk_words * 4, c->stream));
  CK(cudaMemsetAsync(v.chunk_full, 0, (size_t)v.chunk_words * 4, c->stream));
  CK(cudaMemsetAsync(v.region_any, 0, (size_t)v.region_words * 4, c->stream));
  CK(cudaMemsetAsync(v.pool_count, 0, 4, c->stream));
  CK(cudaMemsetAsync(v.pool_free_count, 0, 4, c->stream));
  CK(cudaMemsetAsync(v.df, MESO_DF_K + 1, nc * 64, c->stream));
  CK(cudaMemsetAsync(c->d_loaded, 0, (size_t)v.chunk_words * 4, c->stream));
  CK(cudaMemsetAsync(c->d_stream_stats, 0, 16, c->stream));
  c->stream_kind = kind; c->stream_gran = granularity;
  for (int i = 0; i < 4; i++) c->stream_params[i] = params ? params[i] : 0.0;
  c->streaming = true;
  return MESO_OK;
}

int meso_stream_update_async(MesoCtx* c, const int32_t cam[3], const float forward[3], const MesoViewConfig* view, uint32_t max_new) {
  NEED_SCENE(c);
  if (!c->streaming) return fail(MESO_ERR_ARGUMENT, "meso_stream_update: call meso_stream_begin first");
  if (!cam) return fail(MESO_ERR_ARGUMENT, "meso_stream_update: camera_chunk is null");
  int r = check_view(forward, view, "meso_stream_update");
  if (r != MESO_OK) return r;
  if (max_new > (1u << 20)) return fail(MESO_ERR_ARGUMENT, "meso_stream_update: max_new out of range");
  r = ensure_select_buffers(c, *view);
  if (r != MESO_OK) return r;
  if (max_new > c->stream_list_cap) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_stream_list); c->d_stream_list = nullptr; c->stream_list_cap = 0;
    CK(cudaMalloc(&c->d_stream_list, (size_t)max_new * 4));
    c->stream_list_cap = max_new;
  }
  c->cubes_valid = false;
  JOIN_FRAMES(c);
  launch_select_view(c->lc(), forward, *view, c->d_sel_keys, c->d_sel_count, c->d_sel_out, c->sel_cap);
  launch_stream_worklist(c->lc(), c->v, c->d_sel_out, c->d_sel_count, c->sel_cap, cam, c->d_loaded, max_new, c->d_stream_list, c->d_stream_stats);
  launch_voxelize_list(c->lc(), c->v, c->stream_kind, c->stream_params, c->stream_gran, c->d_overflow, c->d_stream_list, c->d_stream_stats, max_new);
  CK_LAST("stream update");
  return MESO_OK;
}

// The window follows the camera (FChunkPool's eviction in the form a dense window needs, ChunkPool.h:447-622): the new
// origin puts camera_chunk at the window's centre chunk; chunks that leave are evicted and their payload slots recycled,
// chunks that stay are moved to their new slot, chunks that enter are marked not generated.
int meso_stream_recentre(MesoCtx* c, const int32_t camera_chunk[3], int32_t moved[3]) {
  NEED_SCENE(c);
  if (!c->streaming) return fail(MESO_ERR_ARGUMENT, "meso_stream_recentre: call meso_stream_begin first");
  if (!camera_chunk) return fail(MESO_ERR_ARGUMENT, "meso_stream_recentre: camera_chunk is null");
  DVolume& v = c->v;
  int delta[3];
  bool any = false;
  for (int i = 0; i < 3; i++) { delta[i] = (camera_chunk[i] - v.dims[i] / 2) - v.origin[i]; any |= delta[i] != 0; }
  if (moved) for (int i = 0; i < 3; i++) moved[i] = delta[i];
  if (!any) return MESO_OK;
  if (!c->d_shift_scratch) CK(cudaMalloc(&c->d_shift_scratch, (size_t)v.nchunks * MESO_BLOCKS * 4));
  c->cubes_valid = false;
  JOIN_FRAMES(c);
  launch_window_shift(c->lc(), v, delta, c->d_loaded, c->d_shift_scratch);
  CK_LAST("stream recentre");
  return MESO_OK;
}

int meso_pool_stats(MesoCtx* c, int64_t* slots_handed_out, int64_t* slots_free) {
  NEED_SCENE(c);
  uint32_t hw = 0; int fr = 0;
  int r = meso_small_read(c, &hw, c->v.pool_count, 4);
  if (r != MESO_OK) return r;
  r = meso_small_read(c, &fr, c->v.pool_free_count, 4);
  if (r != MESO_OK) return r;
  if (slots_handed_out) *slots_handed_out = hw;
  if (slots_free) *slots_free = fr;
  return MESO_OK;
}

int meso_block_importance(MesoCtx* c, const int32_t cam[3], const float forward[3], const int32_t* chunk_locations, const uint8_t* block_locations,
                          int64_t n, uint32_t chunk_resolution, float* host_out) {
  if (!c || !cam || !forward || n < 0 || (n > 0 && (!chunk_locations || !block_locations || !host_out))) return fail(MESO_ERR_ARGUMENT, "meso_block_importance: bad argument");
  if (n == 0) return MESO_OK;
  CK(cudaSetDevice(c->device));
  int32_t* d_loc = nullptr; uint8_t* d_blk = nullptr; float* d_out = nullptr;
  auto body = [&]() -> int {
    CK(cudaMalloc(&d_loc, (size_t)n * 12)); CK(cudaMalloc(&d_blk, (size_t)n * 3)); CK(cudaMalloc(&d_out, (size_t)n * 4));
    CK(cudaMemcpyAsync(d_loc, chunk_locations, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_blk, block_locations, (size_t)n * 3, cudaMemcpyHostToDevice, c->stream));
    launch_block_importance(c->lc(), d_loc, d_blk, n, cam, forward, chunk_resolution, d_out);
    CK_LAST("block importance");
    CK(cudaMemcpyAsync(host_out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MESO_OK;
  };
  const int r = body();
  cudaFree(d_loc); cudaFree(d_blk); cudaFree(d_out);
  return r;
}

int meso_debug_chunk_instances(MesoCtx* c, MesoGPUSimpleInstanceData* host_out, int64_t cap, int64_t* count) {
  NEED_SCENE(c);
  if (!count || cap < 0 || (cap > 0 && !host_out)) return fail(MESO_ERR_ARGUMENT, "meso_debug_chunk_instances: bad argument");
  MesoGPUSimpleInstanceData* d_out = nullptr;
  const int64_t n_alloc = std::max<int64_t>(std::min<int64_t>(cap, c->v.nchunks), 1);
  auto body = [&]() -> int {
    CK(cudaMalloc(&d_out, (size_t)n_alloc * sizeof(MesoGPUSimpleInstanceData)));
    launch_debug_instances(c->lc(), c->v, c->streaming ? c->d_loaded : nullptr, c->cfg.ChunkSize, d_out, std::min<int64_t>(cap, c->v.nchunks), c->d_tmp_count);
    CK_LAST("debug instances");
    uint32_t n = 0;
    int r = meso_small_read(c, &n, c->d_tmp_count, 4);
    if (r != MESO_OK) return r;
    *count = n;
    const int64_t m = std::min<int64_t>(n, std::min<int64_t>(cap, c->v.nchunks));
    if (m > 0) {
      CK(cudaMemcpyAsync(host_out, d_out, (size_t)m * sizeof(MesoGPUSimpleInstanceData), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      // canonical order: FIVec3Comparator (Helper/Comparator.h:15-23), x then y then z
      std::sort(host_out, host_out + m, [](const MesoGPUSimpleInstanceData& a, const MesoGPUSimpleInstanceData& b) {
        if (a.ChunkLocation[0] != b.ChunkLocation[0]) return a.ChunkLocation[0] < b.ChunkLocation[0];
        if (a.ChunkLocation[1] != b.ChunkLocation[1]) return a.ChunkLocation[1] < b.ChunkLocation[1];
        return a.ChunkLocation[2] < b.ChunkLocation[2];
      });
    }
    return MESO_OK;
  };
  const int r = body();
  cudaFree(d_out);
  return r;
}

int meso_debug_stats(MesoCtx* c, MesoDebugStats* out) {
  NEED_SCENE(c);
  if (!out) return fail(MESO_ERR_ARGUMENT, "meso_debug_stats: null argument");
  memset(out, 0, sizeof(*out));
  const DVolume& v = c->v;
  std::vector<uint32_t> any((size_t)v.chunk_words), loaded((size_t)v.chunk_words, 0xFFFFFFFFu);
  int r = meso_small_read(c, any.data(), v.chunk_any, (size_t)v.chunk_words * 4);
  if (r != MESO_OK) return r;
  if (c->streaming) {
    r = meso_small_read(c, loaded.data(), c->d_loaded, (size_t)v.chunk_words * 4);
    if (r != MESO_OK) return r;
    uint32_t h[4];
    r = meso_small_read(c, h, c->d_stream_stats, 16);
    if (r != MESO_OK) return r;
    out->NewlyAddedVisibleChunk = h[0]; out->MissingChunk = h[1]; out->VisibleChunk = h[2];
  } else {
    out->VisibleChunk = (uint32_t)v.nchunks;
  }
  for (int64_t i = 0; i < v.nchunks; i++) {
    const uint32_t l = (loaded[(size_t)(i >> 5)] >> (i & 31)) & 1u;
    out->LoadedChunk += l;
    out->LoadedChunkWithBlocks += l & ((any[(size_t)(i >> 5)] >> (i & 31)) & 1u);
  }
  uint32_t hw = 0; int fr = 0;
  r = meso_small_read(c, &hw, v.pool_count, 4);
  if (r != MESO_OK) return r;
  r = meso_small_read(c, &fr, v.pool_free_count, 4);
  if (r != MESO_OK) return r;
  out->PayloadSlotsHandedOut = hw; out->PayloadSlotsFree = (uint32_t)std::max(fr, 0);
  out->LoadedBlock = c->n_inst;
  return MESO_OK;
}

int meso_stream_stats(MesoCtx* c, MesoStreamStats* stats) {
  NEED_SCENE(c);
  if (!c->streaming || !stats) return fail(MESO_ERR_ARGUMENT, "meso_stream_stats: no stream or null argument");
  uint32_t h[4];
  CK(cudaMemcpyAsync(h, c->d_stream_stats, 16, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  stats->generated = h[0]; stats->missing = h[1]; stats->candidates = h[2]; stats->in_window = h[3];
  return MESO_OK;
}

int meso_stream_update(MesoCtx* c, const int32_t cam[3], const float forward[3], const MesoViewConfig* view, uint32_t max_new, MesoStreamStats* stats) {
  int r = meso_stream_update_async(c, cam, forward, view, max_new);
  if (r != MESO_OK) return r;
  MesoStreamStats s;
  r = meso_stream_stats(c, &s);
  if (r != MESO_OK) return r;
  if (stats) *stats = s;
  return meso_overflow_finish(c, "meso_stream_update");
}

int meso_stream_loaded(MesoCtx* c, uint32_t* host_words, int64_t n_words) {
  NEED_SCENE(c);
  if (!c->streaming || !host_words || n_words < c->v.chunk_words) return fail(MESO_ERR_ARGUMENT, "meso_stream_loaded: no stream or buffer too small");
  CK(cudaMemcpyAsync(host_words, c->d_loaded, (size_t)c->v.chunk_words * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MESO_OK;
}

int meso_host_register(MesoCtx* c, void* host_ptr, size_t bytes, void** device_ptr) {
  if (!c || !host_ptr || !device_ptr || bytes == 0) return fail(MESO_ERR_ARGUMENT, "meso_host_register: bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  const cudaError_t e = cudaHostGetDevicePointer(device_ptr, host_ptr, 0);
  if (e != cudaSuccess) {
    cudaHostUnregister(host_ptr);
    return fail(MESO_ERR_RUNTIME, std::string("meso_host_register: ") + cudaGetErrorString(e));
  }
  return MESO_OK;
}
int meso_host_unregister(MesoCtx* c, void* host_ptr) {
  if (!c || !host_ptr) return fail(MESO_ERR_ARGUMENT, "meso_host_unregister: bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaHostUnregister(host_ptr));
  return MESO_OK;
}

int meso_device_alloc(MesoCtx* c, size_t bytes, void** dptr) {
  if (!c || !dptr || bytes == 0) return fail(MESO_ERR_ARGUMENT, "meso_device_alloc: bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMalloc(dptr, bytes));   // plain cudaMalloc: exportable with cudaIpcGetMemHandle
  return MESO_OK;
}
int meso_device_memset(MesoCtx* c, void* dptr, int value, size_t bytes) {
  if (!c || !dptr) return fail(MESO_ERR_ARGUMENT, "meso_device_memset: bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(dptr, value, bytes, c->stream));
  return MESO_OK;
}
int meso_device_copy(MesoCtx* c, void* dst, const void* src, size_t bytes) {
  if (!c || (bytes > 0 && (!dst || !src))) return fail(MESO_ERR_ARGUMENT, "meso_device_copy: bad argument");
  CK(cudaSetDevice(c->device));
  if (bytes > 0) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->stream));   // UVA resolves local and peer addresses
  return MESO_OK;
}
int meso_device_free(MesoCtx* c, void* dptr) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  CK(cudaSetDevice(c->device));
  if (dptr) CK(cudaFree(dptr));
  return MESO_OK;
}
int meso_ipc_export(MesoCtx* c, void* dptr, unsigned char handle[MESO_IPC_HANDLE_BYTES]) {
  if (!c || !dptr || !handle) return fail(MESO_ERR_ARGUMENT, "meso_ipc_export: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MESO_IPC_HANDLE_BYTES, "ipc handle size");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle, &h, sizeof(h));
  return MESO_OK;
}
int meso_ipc_open(MesoCtx* c, const unsigned char handle[MESO_IPC_HANDLE_BYTES], void** peer) {
  if (!c || !handle || !peer) return fail(MESO_ERR_ARGUMENT, "meso_ipc_open: bad argument");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CK(cudaIpcOpenMemHandle(peer, h, cudaIpcMemLazyEnablePeerAccess));
  return MESO_OK;
}
int meso_ipc_close(MesoCtx* c, void* peer) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  CK(cudaSetDevice(c->device));
  if (peer) CK(cudaIpcCloseMemHandle(peer));
  return MESO_OK;
}

int meso_download(MesoCtx* c, void* host_dst, const void* dptr, size_t bytes) {
  if (!c || !host_dst || !dptr) return fail(MESO_ERR_ARGUMENT, "meso_download: bad argument");
  CK(cudaSetDevice(c->device));
  return meso_small_read(c, host_dst, dptr, bytes);   // small reads never queue behind a DMA in flight
}

int meso_download_async(MesoCtx* c, void* host_dst, const void* dptr, size_t bytes) {
  if (!c || !host_dst || !dptr) return fail(MESO_ERR_ARGUMENT, "meso_download_async: bad argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(host_dst, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
  return MESO_OK;
}

int meso_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(MESO_ERR_ARGUMENT, "null out");
  CK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return MESO_OK;
}
int meso_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return MESO_OK;
}

int meso_flush_l2(MesoCtx* c) {
  if (!c) return fail(MESO_ERR_ARGUMENT, "null context");
  CK(cudaSetDevice(c->device));
  if (!c->d_flush) {
    c->flush_words = (size_t)64 << 20;  // 256 MiB > 126 MB L2
    CK(cudaMalloc(&c->d_flush, c->flush_words * 4));
  }
  launch_flush(c->lc(), c->d_flush, c->flush_words);
  CK_LAST("flush");
  return MESO_OK;
}

}  // extern "C"
