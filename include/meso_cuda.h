/*
 * meso_cuda.h -- C ABI of the B200-native voxel hot path (libmeso_b200.so).
 *
 * This is the drop-in boundary for the voxel path of yuchengzhong/MesoEngine.  In the reference that path
 * crosses to the device only through lvk::IContext / lvk::ICommandBuffer
 * (ThirdParty/lightweightvk/lvk/LVK.h:720-845); each entry point below cites the reference interface it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the engine-side binding.
 *
 * Conventions
 *   - every function returns MESO_OK (0) or a negative code; meso_last_error() gives the message.  This mirrors
 *     lvk::Result{Code{Ok, ArgumentOutOfRange, RuntimeError}, message} (LVK.h:250-278).  Nothing throws or aborts
 *     across the ABI.
 *   - one caller thread per context (same contract as LVK: "Cannot acquire more than 1 command buffer
 *     simultaneously", lvk/vulkan/VulkanClasses.cpp:3078; all lvk:: calls happen on the main thread).
 *   - host pointers are copied from / written to before the call returns unless the name ends in _device or
 *     _async, in which case the work is only enqueued on the context's stream (meso_ctx_sync waits).
 *   - there is NO CPU fallback: every compute entry point fails with MESO_ERR_RUNTIME if no sm_100 device /
 *     kernel image is available.
 *
 * Units: BlockSize = 1 world unit, 16 blocks per chunk axis, 8 voxels per block axis
 * (Runtimes/Voxel/VoxelSceneConfig.h:22-24); "grid voxel coordinates" are voxels measured from the minimum
 * corner of the resident grid [origin_chunk, origin_chunk + dims_chunks).
 */
#ifndef MESO_CUDA_H
#define MESO_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MESO_API __attribute__((visibility("default")))
#else
#define MESO_API
#endif

#define MESO_OK 0
#define MESO_ERR_ARGUMENT (-1) /* lvk::Result::Code::ArgumentOutOfRange */
#define MESO_ERR_RUNTIME (-2)  /* lvk::Result::Code::RuntimeError (CUDA errors are sticky and land here) */

/* ---- reference layouts, byte for byte ----------------------------------------------------------------- */
/* Runtimes/Voxel/Block/Block.h:21-26 */
typedef struct { uint32_t ChunkIndex; uint8_t BlockLocation[4]; uint32_t BlockFrameStamp; } MesoGPUBlock;
/* Runtimes/Voxel/Chunk/Chunk.h:27-31 */
typedef struct { int32_t ChunkLocation[3]; uint32_t ChunkFrameStamp; } MesoGPUChunk;
/* Runtimes/Shader/GPUStructures.h:36-41 (glm column-major: M[col*4+row]) */
typedef struct {
  float Projection[16];
  float View[16];
  int32_t CameraChunkLocation[4];
  float SubCameraLocation[4];
} MesoGPUUniformCamera;
/* Runtimes/Shader/GPUStructures.h:13-18 */
typedef struct { float BlockSize; uint32_t BlockResolution; float ChunkSize; uint32_t ChunkResolution; } MesoGPUUniformSceneConfig;

/* ---- records produced by this path (DESIGN.md "records") ------------------------------------------------ */
/* w0 = x | y<<16 ; w1 = z | face<<16 | shadow<<19 | hit<<20 ; t in voxel units ; rgba = R | G<<8 | B<<16 | A<<24.
 * face: 0 -X, 1 +X, 2 -Y, 3 +Y, 4 -Z, 5 +Z, 6 = eye inside a solid voxel, 7 = miss.
 * miss = {0xFFFFFFFF, 0x0007FFFF, +inf, 0xFF000000} (clear colour (0,0,0,1): Samples/SimpleVoxel.cpp:315). */
typedef struct { uint32_t w0, w1; float t; uint32_t rgba; } MesoHitRecord;
/* w0 = x | y<<16 ; w1 = z | face<<16 | w<<24 ; w2 = h ; w3 = 0 (reserved: material). */
typedef struct { uint32_t w0, w1, w2, w3; } MesoQuad;
/* Host-derived ray setup (meso_ray_setup); 80 B, passed to the kernel by value. */
typedef struct {
  float o[3]; float two_over_w;
  float U[3]; float two_over_h;
  float V[3]; float pad0;
  float F[3]; float pad1;
  float L[3]; float pad2;
} MesoRaySetup;
typedef struct {
  uint64_t primary, shadow, hits, steps;
  uint64_t touched_chunks, touched_bricks, u_bytes;
} MesoRayStats;

enum { MESO_SDF_SPHERE = 0, MESO_SDF_TERRAIN = 1 };   /* GeneratorHelper.h:120-150 / :90-119 */
enum { MESO_GRAN_BLOCK = 0, MESO_GRAN_VOXEL = 1 };    /* reference: one sample per block; extension: per voxel */
enum { MESO_FLAG_SHADOW = 1 };
enum { MESO_LAYOUT_FRAME = 0, MESO_LAYOUT_TILES = 1 };
#define MESO_TILE_W 32
#define MESO_TILE_H 8

typedef struct MesoCtx MesoCtx;

MESO_API const char* meso_last_error(void);
MESO_API int meso_abi_version(void);

/* ---- context ------------------------------------------------------------------------------------------
 * Replaces lvk::createVulkanContextWithSwapchain (LVK.h:878-882; call site
 * Runtimes/Instance/VoxelWindowsInstance.cpp:104-111) for the voxel path.  One context = one GPU = one stream. */
MESO_API int meso_ctx_create(int device, MesoCtx** out);
MESO_API int meso_ctx_destroy(MesoCtx* ctx);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream; NULL = the CUDA default stream). */
MESO_API int meso_ctx_set_stream(MesoCtx* ctx, void* cuda_stream);
/* Back to the context's own non-blocking stream (the default after meso_ctx_create). */
MESO_API int meso_ctx_use_own_stream(MesoCtx* ctx);
/* lvk::IContext::wait(SubmitHandle) (LVK.h:801) */
MESO_API int meso_ctx_sync(MesoCtx* ctx);
/* Multi-GPU split: this context renders screen tiles t with t % world == rank and meshes
 * chunks c with c % world == rank.  Default (0,1). */
MESO_API int meso_ctx_set_partition(MesoCtx* ctx, int rank, int world);
MESO_API int meso_device_sm_count(MesoCtx* ctx);

/* ---- scene --------------------------------------------------------------------------------------------
 * Replaces the scene-UBO creation/upload (VoxelWindowsInstance.cpp:116-126) and the pool sizing of
 * FChunkPool::Initialize (Runtimes/Voxel/Chunk/ChunkPool.h:250-358): allocates the chunk table, block masks,
 * erode mips, brick pointer table and a brick payload pool of max_bricks x 64 B.  All chunks of the grid are
 * resident ("everything resident" parity mode; no eviction). */
MESO_API int meso_scene_create(MesoCtx* ctx, const MesoGPUUniformSceneConfig* cfg, const int32_t origin_chunk[3],
                               const int32_t dims_chunks[3], uint32_t max_bricks);

/* ---- K1: voxelise --------------------------------------------------------------------------------------
 * Replaces the GeneratorType workers (Runtimes/Voxel/Chunk/ChunkManager.h:61,160-210) running
 * FGeneratorHelper::GenerateSphere / TestGenerator (Runtimes/Helper/GeneratorHelper.h:90-150) for every chunk of
 * the grid.  params = sphere centre xyz + radius in world units (reference: 100,0,0,50); ignored for terrain.
 * Terrain uses the portable fp64 sin; host-generated volumes go through meso_volume_upload. */
MESO_API int meso_voxelize_sdf(MesoCtx* ctx, int kind, const double params[4], int granularity);

/* ---- volume upload / download ---------------------------------------------------------------------------
 * Replaces FChunkPool::UploadChunk/UploadBlock (ChunkPool.h:662-679: whole-buffer lvk::IContext::upload, LVK.h:822)
 * with the canonical sparse form: occ/full = nchunks x 64 words (bit x + 16 y + 256 z), keys[i] = chunk*4096+block
 * ascending, payload = 8 words (z-slices, bit x + 8 y) per partial brick -- the FVolume intent
 * (Runtimes/Voxel/VoxelStructure.h:30-39). */
MESO_API int meso_volume_upload(MesoCtx* ctx, const uint64_t* occ, const uint64_t* full, const uint64_t* keys,
                                const uint64_t* payload, int64_t n_partial);
MESO_API int meso_volume_num_partial(MesoCtx* ctx, int64_t* out);
MESO_API int meso_volume_download(MesoCtx* ctx, uint64_t* occ, uint64_t* full, uint64_t* keys, uint64_t* payload,
                                  int64_t cap_partial, int64_t* n_partial);

/* ---- K2: occupancy / erode mips / hidden-block cull / instance compaction ------------------------------------
 * Replaces FChunk::CalculateOccupancyErodeMipmaps (Runtimes/Voxel/Chunk/Chunk.h:73-94), bShouldVoxelOccupancyCull
 * (:96-100) and the emission loop of FChunkPool::PushToBlockPool (ChunkPool.h:381-445) for all chunks; fills the
 * FGPUChunk table (ChunkPool.h:567).  Instance order: chunk index, then X outer / Z inner (generator order). */
MESO_API int meso_build_occupancy(MesoCtx* ctx, uint32_t frame_stamp, int64_t* n_instances);
MESO_API int meso_download_chunk_table(MesoCtx* ctx, MesoGPUChunk* out /* nchunks */);
MESO_API int meso_download_mips(MesoCtx* ctx, uint64_t* out /* nchunks*3*64: Mip1..Mip3 */);
MESO_API int meso_download_instances(MesoCtx* ctx, MesoGPUBlock* out, int64_t cap);

/* ---- K4: raymarch -----------------------------------------------------------------------------------------
 * Replaces the camera-UBO upload + the instanced draw + depth resolve:
 * VoxelWindowsInstance::RenderStart (VoxelWindowsInstance.cpp:410-419) and SimpleVoxelWindowsInstance::Render
 * (Samples/SimpleVoxel.cpp:352-398: cmdBindVertexBuffer/cmdPushConstants/cmdDrawIndexed(8, MaxBlockCount)).
 * meso_ray_setup is a pure host function (fp32, no fma). */
MESO_API int meso_ray_setup(const MesoGPUUniformCamera* cam, const int32_t origin_chunk[3], int width, int height,
                            const float light_dir[3], MesoRaySetup* out);
/* End-to-end: camera in host memory -> records in host memory (row-major width x height). */
MESO_API int meso_raymarch(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                           const float light_dir[3], MesoHitRecord* host_records);
/* Frame ring, the reference's kNumBufferedFrames (Samples/SimpleVoxel.cpp:15; per-frame buffers indexed by
 * RenderFrameIndex, VoxelWindowsInstance.cpp:404-408): meso_raymarch_async renders into ring slot `slot` (0..3) and
 * starts the copy of the records into host_records (pinned memory recommended) on a copy stream; it returns as soon as
 * the work is enqueued.  meso_frame_wait(slot) returns when that slot's records are in host memory.  Re-using a slot
 * waits for its previous frame first.  The copy of frame k overlaps the traversal of frame k+1. */
#define MESO_FRAME_RING 4
MESO_API int meso_raymarch_async(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                 const float light_dir[3], MesoHitRecord* host_records, int slot);
MESO_API int meso_frame_wait(MesoCtx* ctx, int slot);
/* Enqueue only; d_records is device memory.  MESO_LAYOUT_FRAME: row-major frame, only this rank's tiles are written.
 * MESO_LAYOUT_TILES: this rank's tiles packed as [local_tile][MESO_TILE_H][MESO_TILE_W]. */
MESO_API int meso_raymarch_device(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                  const float light_dir[3], void* d_records, int layout);
/* Same frame with counters; also marks touched chunks/bricks to report the algorithmic bytes U. */
MESO_API int meso_raymarch_stats(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                 const float light_dir[3], MesoRayStats* out);
/* De-interleave gathered tile-packed buffers (world x tiles_per_rank x 256 records) into a row-major frame. */
MESO_API int meso_compose_tiles_device(MesoCtx* ctx, const void* d_tiles, int world, int width, int height, void* d_frame);
MESO_API int64_t meso_tiles_per_rank(int width, int height, int world);

/* ---- K3: face cull + greedy merge -------------------------------------------------------------------------
 * The north-star form of the reference's "mesher" (hidden-block cull + instance compaction, ChunkPool.h:381-445):
 * exposed faces of every occupied brick merged greedily into quads.  Order of the output list is unspecified;
 * compare after a canonical sort. */
MESO_API int meso_mesh(MesoCtx* ctx, MesoQuad* host_quads, int64_t cap, int64_t* n_quads);
MESO_API int meso_mesh_device(MesoCtx* ctx, void* d_quads, int64_t cap, int64_t* n_quads /* host, written after sync */);

/* ---- K5: edit --------------------------------------------------------------------------------------------- */
MESO_API int meso_carve_sphere(MesoCtx* ctx, const int32_t center[3], int32_t radius, int64_t* n_dirty);
MESO_API int meso_download_dirty(MesoCtx* ctx, uint64_t* keys, int64_t cap);
/* Re-mesh only the bricks of the last carve's dirty list and their six neighbours. */
MESO_API int meso_remesh_dirty(MesoCtx* ctx, MesoQuad* host_quads, int64_t cap, int64_t* n_quads, uint64_t* host_keys,
                               int64_t cap_keys, int64_t* n_keys);

/* ---- peer memory (one process per GPU) -------------------------------------------------------------------------
 * Fused gather: every rank's raymarch kernel stores its tile records straight into the frame buffer of the gathering
 * rank over NVLink (16 B stores to a peer mapping), so the transfer overlaps the traversal and no separate gather or
 * compose pass runs.  The gathering rank allocates the frame with meso_device_alloc and exports it; the others open the
 * handle and pass the returned pointer as d_records to meso_raymarch_device(..., MESO_LAYOUT_FRAME).  A stream-ordered
 * barrier between the ranks (e.g. a 4-byte NCCL all-reduce) tells the owner that all tiles have landed. */
#define MESO_IPC_HANDLE_BYTES 64
MESO_API int meso_device_alloc(MesoCtx* ctx, size_t bytes, void** dptr);
MESO_API int meso_device_free(MesoCtx* ctx, void* dptr);
MESO_API int meso_ipc_export(MesoCtx* ctx, void* dptr, unsigned char handle[MESO_IPC_HANDLE_BYTES]);
MESO_API int meso_ipc_open(MesoCtx* ctx, const unsigned char handle[MESO_IPC_HANDLE_BYTES], void** peer_dptr);
MESO_API int meso_ipc_close(MesoCtx* ctx, void* peer_dptr);
/* lvk::IContext::download (LVK.h:831): device -> host on the context's stream, returns when the bytes are in host memory. */
MESO_API int meso_download(MesoCtx* ctx, void* host_dst, const void* dptr, size_t bytes);
/* Same, enqueue only (host_dst should be pinned); meso_ctx_sync or an event on the stream tells when the bytes are there. */
MESO_API int meso_download_async(MesoCtx* ctx, void* host_dst, const void* dptr, size_t bytes);

/* ---- utilities -------------------------------------------------------------------------------------------- */
MESO_API int meso_host_alloc(size_t bytes, void** out); /* pinned */
MESO_API int meso_host_free(void* p);
MESO_API int meso_flush_l2(MesoCtx* ctx);               /* writes a 256 MiB scratch buffer */
/* Number of kernels of this library launched on this context since creation (bench.py's gpu_launches). */
MESO_API int64_t meso_launch_count(MesoCtx* ctx);

#ifdef __cplusplus
}
#endif
#endif
