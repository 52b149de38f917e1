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
/* w0 = x | y<<16 ; w1 = z | face<<16 | level<<19 | w<<24 ; w2 = h ; w3 = 0 (reserved: material).  (x,y,z) = minimum-corner
 * voxel of the quad, w along u, h along v with (u,v) = (y,z) / (x,z) / (x,y) for X / Y / Z faces.  level 0: voxel-level quad
 * inside one brick (w, h <= 8); level 1: brick-level quad -- whole faces of full bricks towards absent bricks, merged inside one
 * chunk (w, h multiples of 8, <= 128). */
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
  /* SIMT accounting of the same launch (profiles/README.md): steps of primary rays only; 32 x the longest lane of every
   * warp, summed (primary phase / shadow phase) -- steps / warp_slots is the lane utilisation the iteration counts allow;
   * steps by kind: 0 voxel, 1 empty brick (8^3), 2 / 3 / 4 distance-field steps over 32^3 cells with cube half-width
   * 0 (one cell) / 1..3 / 4 and more cells. */
  uint64_t steps_primary, warp_slots_primary, warp_slots_shadow;
  uint64_t level_steps[5];
} MesoRayStats;

/* Runtimes/Shader/GPUStructures.h:86-92: the instance record of the reference's chunk-wireframe debug pass (48 B) */
typedef struct { float Position[3]; int32_t ChunkLocation[3]; float Scale; float Rotation[4]; float Marker; } MesoGPUSimpleInstanceData;
/* The counters of FChunkManage::RenderManagerInfo (Runtimes/Voxel/Chunk/ChunkManager.h:466-489) */
typedef struct {
  uint32_t VisibleChunk;          /* DebugVisibleChunkNum: size of the desired set of the last update                       */
  uint32_t LoadedChunk;           /* ChunkPool.CurrentDebugDrawInstanceCount: resident chunks (with blocks or empty)        */
  uint32_t LoadedChunkWithBlocks; /* of which FChunk (the rest are FEmptyChunk)                                             */
  uint32_t NewlyAddedVisibleChunk;/* DebugNewVisibleChunkNum: chunks generated by the last update                           */
  uint32_t MissingChunk;          /* desired chunks inside the window still not generated                                   */
  uint32_t PayloadSlotsHandedOut, PayloadSlotsFree;   /* brick payload pool (no reference counterpart)                      */
  int64_t LoadedBlock;            /* ChunkPool.CurrentBlockCount: instances emitted by the last meso_build_occupancy        */
} MesoDebugStats;

/* K6.  FChunkManageHelper::FTempChunkDataType = std::pair<float, ivec3> (ChunkManagerHelper.h:76): 16 B. */
typedef struct { float Importance; int32_t Offset[3]; } MesoChunkCandidate;
/* The view fields of FVoxelSceneConfig (VoxelSceneConfig.h:34-35,40; defaults 24, 6, 120).  Mode 0 =
 * GetDesiredShowChunkLocationByView, 1 = GetDesiredShowChunkLocationSimple. */
typedef struct { uint32_t ViewForwardLoadChunkSize, ViewBackwardLoadChunkSize; float ViewChunkAngle; uint32_t Mode; } MesoViewConfig;
typedef struct {
  uint32_t generated;    /* chunks generated by this update (<= max_new)                                  */
  uint32_t missing;      /* desired chunks inside the window still not generated after this update       */
  uint32_t candidates;   /* size of the desired set for this view                                         */
  uint32_t in_window;    /* candidates that fall inside the resident grid window                          */
} MesoStreamStats;

enum { MESO_SDF_SPHERE = 0, MESO_SDF_TERRAIN = 1 };   /* GeneratorHelper.h:120-150 / :90-119 */
enum { MESO_GRAN_BLOCK = 0, MESO_GRAN_VOXEL = 1 };    /* reference: one sample per block; extension: per voxel */
/* MESO_FLAG_RGBA8: write only the colour word, 4 B per pixel instead of the 16 B record -- the format of the reference's
 * TEXOffscreenColor (RGBA_UN8, Samples/SimpleVoxel.cpp:293), ready for IContext::upload(TextureHandle, ...); the output
 * buffer is then uint32_t[width * height] (frame layout only). */
enum { MESO_FLAG_SHADOW = 1, MESO_FLAG_RGBA8 = 2,
       /* The walk reads the per-octant forward-cube tables (meso_build_cubes) whenever they are current -- they are built
        * with the volume by meso_voxelize_sdf / meso_volume_upload* and survive carves -- and falls back to the distance
        * field when they are not (after meso_stream_begin / _update).  Identical records either way.
        * MESO_FLAG_CUBES insists on the tables (error if they are not current); MESO_FLAG_NO_CUBES forces the distance-field walk. */
       MESO_FLAG_CUBES = 4, MESO_FLAG_NO_CUBES = 8 };
enum { MESO_LAYOUT_FRAME = 0, MESO_LAYOUT_TILES = 1, MESO_LAYOUT_SLABS = 2 /* meso_raymarch_device_slabs */ };
#define MESO_MAX_SLABS 8
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

/* The generator plug-in's fourth argument (GeneratorType = FChunk(ivec3, float, unsigned char, uint32_t MipmapLevel),
 * ChunkManager.h:61; passed down at :167,233,291 and ignored by both reference generators): level of detail of the
 * generation.  Block granularity, one SDF sample per (2^MipmapLevel)^3 blocks taken at the group's minimum-corner block
 * (MipmapLevel 0 = meso_voxelize_sdf(..., MESO_GRAN_BLOCK); 4 = one sample per chunk).  The definition is this library's
 * ("parity unpinned by reference; bit-exact vs repo oracle"). */
MESO_API int meso_voxelize_sdf_lod(MesoCtx* ctx, int kind, const double params[4], uint32_t mipmap_level);

/* ---- volume upload / download ---------------------------------------------------------------------------
 * Replaces FChunkPool::UploadChunk/UploadBlock (ChunkPool.h:662-679: whole-buffer lvk::IContext::upload, LVK.h:822)
 * with the canonical sparse form: occ/full = nchunks x 64 words (bit x + 16 y + 256 z), keys[i] = chunk*4096+block
 * ascending, payload = 8 words (z-slices, bit x + 8 y) per partial brick -- the FVolume intent
 * (Runtimes/Voxel/VoxelStructure.h:30-39). */
MESO_API int meso_volume_upload(MesoCtx* ctx, const uint64_t* occ, const uint64_t* full, const uint64_t* keys,
                                const uint64_t* payload, int64_t n_partial);
/* The reference's OWN upload records, as FChunkPool::UploadChunk / UploadBlock hand them to lvk::IContext::upload
 * (ChunkPool.h:662-679): the FGPUChunk table and the FGPUBlock pool.  A block counts iff the reference's vertex shader
 * would draw it (Samples/SimpleVoxel.cpp:160-166: ChunkIndex != INT_MAX, chunk record valid, ChunkFrameStamp ==
 * BlockFrameStamp) and its chunk lies inside the window; it becomes an all-solid brick (the reference has no
 * voxel-in-brick level).  flags = 0 replaces the window's content (the reference re-uploads whole pools);
 * MESO_UPLOAD_MERGE adds the blocks to what is resident (delta upload: only the new records travel).  *n_accepted = blocks
 * that counted.  This is how a user-written CPU generator (GeneratorType, ChunkManager.h:61) gets its FChunk.Blocks in. */
enum { MESO_UPLOAD_MERGE = 1 };
MESO_API int meso_volume_upload_blocks(MesoCtx* ctx, const MesoGPUChunk* chunks, int64_t n_chunks, const MesoGPUBlock* blocks,
                                       int64_t n_blocks, uint32_t flags, int64_t* n_accepted);
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
/* (Re)builds the forward-cube tables of the raymarch walk for the current volume (32^3-cell cubes and brick cubes; enqueue
 * only).  Called implicitly by meso_voxelize_sdf and meso_volume_upload*; call it after a series of meso_stream_update to get
 * the faster walk back.  Calls that may ADD voxels to part of the grid (stream begin / update) invalidate the tables; a
 * carve only removes voxels, every certified cube stays empty, and the tables stay valid (conservative).  No reference
 * counterpart: the reference resolves visibility with a rasterised instanced draw (Samples/SimpleVoxel.cpp:352-398). */
MESO_API int meso_build_cubes(MesoCtx* ctx);
/* Debug / test read-back of two of the tables (either pointer may be NULL): cell = 8 * ncells bytes, octant-major, ncells =
 * product of 4 * dims_chunks; brick = nchunks * 4096 uint16 (2 bits per octant: edge - 1), defined for the empty bricks of
 * non-empty 32^3 cells (zero elsewhere: never read). */
MESO_API int meso_download_cubes(MesoCtx* ctx, uint8_t* cell, uint16_t* brick);
/* The present path.  The reference blits TEXOffscreenColor (RGBA_UN8, Samples/SimpleVoxel.cpp:293) to the
 * swapchain in RenderEnd (Runtimes/Instance/VoxelWindowsInstance.cpp:433-473) and its context fills a texture range with
 * lvk::IContext::upload(TextureHandle, const TextureRangeDesc&, const void* data[]) (lvk/LVK.h:830; TextureRangeDesc :683-692:
 * x, y, dimensions).  meso_present_rgba8 produces exactly that `data`: the colour of the pixels [x, x + width) x [y, y + height)
 * of the frame, tightly packed (row pitch = range width * 4), R | G<<8 | B<<16 | A<<24 per pixel, clear colour (0,0,0,1);
 * range = NULL is the whole frame.  Only the screen tiles the range touches are traced.  (CUDA-Vulkan external memory would
 * save the trip through the host; there is no Vulkan in this image, so that interop is not built and not tested.) */
typedef struct { uint32_t x, y, width, height; } MesoTextureRange;
MESO_API int meso_present_rgba8(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                const float light_dir[3], const MesoTextureRange* range, void* host_pixels);
/* Frame ring, the reference's kNumBufferedFrames (Samples/SimpleVoxel.cpp:15; per-frame buffers indexed by
 * RenderFrameIndex, VoxelWindowsInstance.cpp:404-408): meso_raymarch_async renders into ring slot `slot` (0..3) and
 * starts the copy of the records into host_records (pinned memory recommended) on a copy stream; it returns as soon as
 * the work is enqueued.  meso_frame_wait(slot) returns when that slot's records are in host memory.  Re-using a slot
 * waits for its previous frame first.  The copy of frame k overlaps the traversal of frame k+1.  Calls that rewrite the
 * volume (voxelize, upload, carve, stream begin / update) are ordered behind the traversal of every frame still in
 * flight, so an edit between two frames never races a frame that was started before it. */
#define MESO_FRAME_RING 4
MESO_API int meso_raymarch_async(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                 const float light_dir[3], MesoHitRecord* host_records, int slot);
MESO_API int meso_frame_wait(MesoCtx* ctx, int slot);
/* Enqueue only; d_records is device memory.  MESO_LAYOUT_FRAME: row-major frame, only this rank's tiles are written.
 * MESO_LAYOUT_TILES: this rank's tiles packed as [local_tile][MESO_TILE_H][MESO_TILE_W]. */
MESO_API int meso_raymarch_device(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                  const float light_dir[3], void* d_records, int layout);
/* Slab gather, fused into the kernel's store (N GPUs, end to end into host memory): the frame is cut into n_slabs
 * horizontal slabs of rows_per_slab scanlines (a multiple of MESO_TILE_H; the last slab may be shorter), slab k living in
 * d_slabs[k] -- this rank's own buffer for the rows it owns, the owning rank's buffer (peer memory: meso_ipc_open, or
 * another context of the same process / meso_group) for the others.  Every rank still traces its interleaved tiles
 * (t % world == rank, balanced), but each record is stored exactly once into the slab it belongs to over NVLink, so that
 * after the stream-ordered rendezvous every rank holds a CONTIGUOUS part of the frame and copies it to the host with one
 * large DMA over its own PCIe link.  Slab k holds rows_in_slab x width records (or uint32 colours with MESO_FLAG_RGBA8).
 * Enqueue only. */
MESO_API int meso_raymarch_device_slabs(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                        const float light_dir[3], void* const* d_slabs, int n_slabs, int rows_per_slab);
/* Stream-ordered rendezvous between GPUs WITHOUT a collective: arrival words.  A word is 32 bits of device memory (this
 * GPU's or a peer's: meso_device_alloc + meso_ipc_open, or members of a group).  meso_signal_device adds 1 to each of n words,
 * system scope, once everything enqueued before it on the context's stream is complete and visible ("my tiles of this frame
 * are in your buffer", "my slab has reached the host").  meso_wait_device holds the stream until a word in THIS GPU's memory
 * has reached target (wrap-safe comparison), i.e. until `target` signals in total have arrived.  A wait gives up after about
 * two seconds and raises a flag that meso_wait_timed_out returns (and clears) instead of hanging the device.  Enqueue only.
 * (Counting finished CTAs inside the raymarch kernel to fuse the signal into it was measured: 130 k same-address atomics and
 * a fence per CTA cost more than the extra launch.) */
MESO_API int meso_signal_device(MesoCtx* ctx, void* const* d_words, int n);
MESO_API int meso_wait_device(MesoCtx* ctx, const void* d_word, uint32_t target);
MESO_API int meso_wait_timed_out(MesoCtx* ctx, int* out);
/* Same frame with counters; also marks touched chunks/bricks to report the algorithmic bytes U. */
MESO_API int meso_raymarch_stats(MesoCtx* ctx, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                 const float light_dir[3], MesoRayStats* out);
/* The colour words of n records (device), packed into d_rgba8 (uint32[n], device): the image travels to the host at 4 B per
 * pixel while the records stay on the device for picking.  Enqueue only. */
MESO_API int meso_pack_rgba8_device(MesoCtx* ctx, const void* d_records, int64_t n, void* d_rgba8);
/* De-interleave gathered tile-packed buffers (world x tiles_per_rank x 256 records) into a row-major frame. */
MESO_API int meso_compose_tiles_device(MesoCtx* ctx, const void* d_tiles, int world, int width, int height, void* d_frame);
MESO_API int64_t meso_tiles_per_rank(int width, int height, int world);

/* ---- K3: face cull + greedy merge -------------------------------------------------------------------------
 * The north-star form of the reference's "mesher" (hidden-block cull + instance compaction, ChunkPool.h:381-445):
 * exposed faces of every occupied brick merged greedily into quads -- voxel faces inside each brick, and the faces of full
 * bricks towards absent ones across each chunk (the block-granular scenes of the reference are made of those only: a flat
 * chunk face is one quad).  Order of the output list is unspecified; compare after a canonical sort. */
MESO_API int meso_mesh(MesoCtx* ctx, MesoQuad* host_quads, int64_t cap, int64_t* n_quads);
MESO_API int meso_mesh_device(MesoCtx* ctx, void* d_quads, int64_t cap, int64_t* n_quads /* host, written after sync */);
/* The 8-byte quad count of the last meso_mesh_device on this context, copied device -> device into d_count_out (int64) on
 * the context's stream: lets a multi-GPU host all-gather the counts without a host read in between.  Enqueue only. */
MESO_API int meso_mesh_count_device(MesoCtx* ctx, void* d_count_out);
/* Fused gather of the quad lists: meshes this rank's chunks (meso_ctx_set_partition) and appends the quads to a list
 * shared by all ranks -- d_quads and the 8-byte d_counter may be the gathering rank's memory opened with meso_ipc_open:
 * every warp reserves its slots with one system-scope atomicAdd on the counter and stores its 16 B records over NVLink.
 * The counter is NOT reset here (the owner zeroes it before the ranks start); enqueue only. */
MESO_API int meso_mesh_device_shared(MesoCtx* ctx, void* d_quads, void* d_counter, int64_t cap);

/* ---- K5: edit --------------------------------------------------------------------------------------------- */
MESO_API int meso_carve_sphere(MesoCtx* ctx, const int32_t center[3], int32_t radius, int64_t* n_dirty);
MESO_API int meso_download_dirty(MesoCtx* ctx, uint64_t* keys, int64_t cap);
/* Re-mesh only the bricks of the last carve's dirty list and their six neighbours (host_keys: those bricks, chunk*4096+block).
 * Returned: the voxel-level quads of the listed bricks and the brick-level quads of every chunk that holds one of them -- the
 * caller replaces exactly those (a quad's brick and chunk follow from its corner). */
MESO_API int meso_remesh_dirty(MesoCtx* ctx, MesoQuad* host_quads, int64_t cap, int64_t* n_quads, uint64_t* host_keys,
                               int64_t cap_keys, int64_t* n_keys);

/* ---- K6: resident-set selection and streaming generation --------------------------------
 * The desired set for a view: FChunkManageHelper::GetDesiredShowChunkLocationByView / ...Simple
 * (Runtimes/Voxel/Chunk/ChunkManagerHelper.h:89-198), scored by FImportanceComputeInfo::CalculateChunkImportance
 * (:26-44), returned in the pop order of the reference's priority queue (importance descending; ties, which the heap
 * leaves unspecified, in loop order X outer / Z inner).  Computed on the device for the exact forward vector on every
 * call; the reference instead bakes BakeVisibilityViewNum = 256 directions at start-up and looks up the nearest
 * (BakeVisibilityByView :201-234, TNearestMap, ChunkManager.h:113-124) -- pass that direction (meso_baked_direction)
 * to reproduce its table.  Writes min(count, cap) candidates; *count is the full size of the set. */
MESO_API int meso_select_view_chunks(MesoCtx* ctx, const float forward[3], const MesoViewConfig* view, MesoChunkCandidate* host_out,
                                     int64_t cap, int64_t* count);
/* FImportanceComputeInfo::CalculateChunkImportance for n absolute chunk locations (host, n x 3 int32). */
MESO_API int meso_chunk_importance(MesoCtx* ctx, const int32_t camera_chunk[3], const float forward[3], const int32_t* locations,
                                   int64_t n, float* host_out);
/* GetFibonacciSphere<float>(samples)[TNearestMap::Query(forward)] (VoxelMathHelper.h:49-71, NearestMap.h:32-47): the
 * baked direction the reference would use for this forward vector.  Host arithmetic (libm cosf/sinf); no device work. */
MESO_API int meso_baked_direction(uint32_t samples, const float forward[3], float out_direction[3], uint32_t* out_index);
/* Streaming generation into the scene's grid window, all on the device.  meso_stream_begin empties the volume and fixes
 * the generator (FChunkManage::Initialize, ChunkManager.h:90-102).  meso_stream_update is FChunkManage::UpdateChunks +
 * UpdateLoadingQueue + MultiThreadGenerator (ChunkManager.h:134-283): build the desired set for (camera_chunk, forward),
 * walk it in priority order, and generate the first max_new chunks that are inside the window and not generated yet
 * (max_new plays MaxUnsyncedLoadChunkCount, VoxelSceneConfig.h:37); everything already generated stays.  Chunks outside
 * the window are skipped (the window replaces the reference's pool + eviction: it holds every chunk of the grid). */
MESO_API int meso_stream_begin(MesoCtx* ctx, int sdf_kind, const double params[4], int granularity);
MESO_API int meso_stream_update(MesoCtx* ctx, const int32_t camera_chunk[3], const float forward[3], const MesoViewConfig* view,
                                uint32_t max_new, MesoStreamStats* stats);
/* Same, enqueue only: no host wait; stats of the last update can be fetched later with meso_stream_stats. */
MESO_API int meso_stream_update_async(MesoCtx* ctx, const int32_t camera_chunk[3], const float forward[3], const MesoViewConfig* view,
                                      uint32_t max_new);
MESO_API int meso_stream_stats(MesoCtx* ctx, MesoStreamStats* stats);
/* The window follows the camera: moves the resident grid so that camera_chunk is its centre chunk (origin' = camera_chunk -
 * dims / 2).  This is what FChunkPool's slot probing + eviction (ChunkPool.h:447-622: PushToPool / PushChunk /
 * PushEmptyChunk overriding the least important slot) achieve in the reference, in the form a dense window needs: chunks that
 * leave the window are EVICTED -- their brick payload slots go onto a free stack and are handed out again before any new
 * slot -- chunks that stay keep their data (moved to their new slot), chunks that enter are empty and "not generated", so
 * the next meso_stream_update generates them in priority order.  moved (optional) = the origin's displacement in chunks.
 * Ordered behind frames in flight; the forward-cube tables are invalidated (the walk uses the distance field). */
MESO_API int meso_stream_recentre(MesoCtx* ctx, const int32_t camera_chunk[3], int32_t moved[3]);
/* Payload pool: slots handed out so far (high-water mark) and slots on the free stack (returned by evicted chunks). */
MESO_API int meso_pool_stats(MesoCtx* ctx, int64_t* slots_handed_out, int64_t* slots_free);
/* FImportanceComputeInfo::CalculateBlockImportance (ChunkManagerHelper.h:50-70) for n blocks: chunk_locations n x 3 int32
 * (absolute), block_locations n x 3 uint8; restated literally, including the dead near branch (unsigned comparison). */
MESO_API int meso_block_importance(MesoCtx* ctx, const int32_t camera_chunk[3], const float forward[3], const int32_t* chunk_locations,
                                   const uint8_t* block_locations, int64_t n, uint32_t chunk_resolution, float* host_out);
/* Bit per chunk slot of the window: generated (EChunkState != absent in ChunksLookupTable).  words = ceil(nchunks / 32). */
MESO_API int meso_stream_loaded(MesoCtx* ctx, uint32_t* host_words, int64_t n_words);

/* ---- debug visualisation parity ------------------------------------------------------------------
 * The data behind the reference's operator feedback: one FGPUSimpleInstanceData per resident chunk (what
 * FChunkManage::DrawDebugVisibleChunk draws an octahedron for, Runtimes/Voxel/Chunk/ChunkManager.h:402-428, records built at
 * ChunkPool.h:550-561, coloured by Marker in ShaderWireFrame.h:18-22) in FIVec3Comparator order of ChunkLocation, and the
 * counters of RenderManagerInfo (ChunkManager.h:466-489).  Drawing them stays with the engine. */
MESO_API int meso_debug_chunk_instances(MesoCtx* ctx, MesoGPUSimpleInstanceData* host_out, int64_t cap, int64_t* count);
MESO_API int meso_debug_stats(MesoCtx* ctx, MesoDebugStats* out);

/* ---- peer memory (one process per GPU) -------------------------------------------------------------------------
 * Fused gather: every rank's raymarch kernel stores its tile records straight into the frame buffer of the gathering
 * rank over NVLink (16 B stores to a peer mapping), so the transfer overlaps the traversal and no separate gather or
 * compose pass runs.  The gathering rank allocates the frame with meso_device_alloc and exports it; the others open the
 * handle and pass the returned pointer as d_records to meso_raymarch_device(..., MESO_LAYOUT_FRAME).  A stream-ordered
 * barrier between the ranks (e.g. a 4-byte NCCL all-reduce) tells the owner that all tiles have landed. */
#define MESO_IPC_HANDLE_BYTES 64
MESO_API int meso_device_alloc(MesoCtx* ctx, size_t bytes, void** dptr);
MESO_API int meso_device_free(MesoCtx* ctx, void* dptr);
MESO_API int meso_device_memset(MesoCtx* ctx, void* dptr, int value, size_t bytes);   /* enqueue only */
/* Device-to-device copy on the context's stream, enqueue only; either pointer may be a peer GPU's memory opened with
 * meso_ipc_open (the bulk leg of a quad gather at prefix offsets: one transfer per rank instead of one reservation per
 * warp).  The role lvk::ICommandBuffer::cmdCopyBuffer has for device buffers (lvk/LVK.h). */
MESO_API int meso_device_copy(MesoCtx* ctx, void* dst_dptr, const void* src_dptr, size_t bytes);
MESO_API int meso_ipc_export(MesoCtx* ctx, void* dptr, unsigned char handle[MESO_IPC_HANDLE_BYTES]);
MESO_API int meso_ipc_open(MesoCtx* ctx, const unsigned char handle[MESO_IPC_HANDLE_BYTES], void** peer_dptr);
MESO_API int meso_ipc_close(MesoCtx* ctx, void* peer_dptr);
/* lvk::IContext::download (LVK.h:831): device -> host on the context's stream, returns when the bytes are in host memory. */
MESO_API int meso_download(MesoCtx* ctx, void* host_dst, const void* dptr, size_t bytes);
/* Same, enqueue only (host_dst should be pinned); meso_ctx_sync or an event on the stream tells when the bytes are there. */
MESO_API int meso_download_async(MesoCtx* ctx, void* host_dst, const void* dptr, size_t bytes);

/* Fused gather into HOST memory.  meso_host_register page-locks and maps an existing host allocation -- typically a
 * shared-memory segment that every rank's process has mapped -- and returns the device address kernels can store to.
 * Passing that address (plus the frame's offset) as d_records to meso_raymarch_device(..., MESO_LAYOUT_FRAME) makes
 * every rank deliver its own tiles over its own PCIe link, 512 B per tile row, while it is still tracing: N links
 * instead of the gathering rank's one, and no device-side frame at all.  The records are in host memory when the
 * stream-ordered rendezvous behind the kernels (e.g. a 4-byte NCCL all-reduce) has completed. */
MESO_API int meso_host_register(MesoCtx* ctx, void* host_ptr, size_t bytes, void** device_ptr);
MESO_API int meso_host_unregister(MesoCtx* ctx, void* host_ptr);

/* ---- N GPUs of one box behind one handle (single-process hosts) -----------------------------------------------------
 * Replaces lvk::createVulkanContextWithSwapchain (Runtimes/Instance/VoxelWindowsInstance.cpp:104-111) for a host that
 * wants the frame split over several GPUs: a MesoGroup is N member contexts over a REPLICATED volume (member r renders
 * screen tiles t % N == r and meshes chunks c % N == r) with peer access enabled between them.  The exchanges are fused
 * into the kernels' stores and ordered on the devices (cross-device event waits); no host thread, no NCCL:
 *   frames: every member's raymarch kernel stores each record straight into the horizontal slab of the frame it belongs
 *           to (MESO_LAYOUT_SLABS; slab r lives on member r, peer memory over NVLink), then every member copies its
 *           contiguous slab to host_records with one large DMA over its own PCIe link;
 *   quads:  members mesh into their own lists, delivered at prefix offsets (host) or into per-member segments of a list on
 *           member 0 written over NVLink by the mesh kernel itself (device);
 *   edits:  carves are replicated compute; the dirty re-mesh is sharded by key.
 * devices may name the same GPU more than once (two members sharing a device).  One caller thread per group.  Any
 * single-GPU entry point may be used on a member (meso_group_ctx) between group calls. */
typedef struct MesoGroup MesoGroup;
MESO_API int meso_group_create(const int* devices, int n, MesoGroup** out);
MESO_API int meso_group_destroy(MesoGroup* g);
MESO_API int meso_group_size(MesoGroup* g);
MESO_API MesoCtx* meso_group_ctx(MesoGroup* g, int rank);
MESO_API int meso_group_sync(MesoGroup* g);
MESO_API int meso_group_scene_create(MesoGroup* g, const MesoGPUUniformSceneConfig* cfg, const int32_t origin_chunk[3],
                                     const int32_t dims_chunks[3], uint32_t max_bricks);
MESO_API int meso_group_voxelize_sdf(MesoGroup* g, int kind, const double params[4], int granularity);
MESO_API int meso_group_volume_upload_blocks(MesoGroup* g, const MesoGPUChunk* chunks, int64_t n_chunks, const MesoGPUBlock* blocks,
                                             int64_t n_blocks, uint32_t flags, int64_t* n_accepted);
MESO_API int meso_group_carve_sphere(MesoGroup* g, const int32_t center[3], int32_t radius, int64_t* n_dirty);
/* meso_raymarch / meso_raymarch_async / meso_frame_wait over the group (ring of MESO_FRAME_RING frames; host_records should
 * be pinned, e.g. meso_host_alloc, for the copies to overlap the next frame). */
MESO_API int meso_group_raymarch(MesoGroup* g, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                 const float light_dir[3], MesoHitRecord* host_records);
MESO_API int meso_group_raymarch_async(MesoGroup* g, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags,
                                       const float light_dir[3], MesoHitRecord* host_records, int slot);
MESO_API int meso_group_frame_wait(MesoGroup* g, int slot);
/* meso_mesh over the group: host_quads = the members' lists concatenated in member order; counts (optional, N entries) =
 * the per-member sizes. */
MESO_API int meso_group_mesh(MesoGroup* g, MesoQuad* host_quads, int64_t cap, int64_t* n_quads, int64_t* counts);
/* Device-resident gather on member 0: member r writes its quads into segment [r * (cap / N), ...) of d_quads (member 0's
 * memory) from inside its mesh kernel, over NVLink, with a counter in its own memory.  segment_counts (optional, N
 * entries); compact != 0 closes the gaps with device-local copies: one contiguous list of *n_quads records. */
MESO_API int meso_group_mesh_device(MesoGroup* g, void* d_quads_on_member0, int64_t cap, int64_t* n_quads, int64_t* segment_counts,
                                    int compact);
/* meso_remesh_dirty over the group after meso_group_carve_sphere: the members share the work by a hash of the brick / chunk key,
 * host_quads = their lists concatenated.  The carve is replicated, so the re-meshed bricks are the same on every member: take the
 * keys from any one of them (meso_remesh_dirty(meso_group_ctx(g, 0), NULL, 0, &n, keys, cap_keys, &n_keys) returns member 0's
 * share of the quads and ALL the keys). */
MESO_API int meso_group_remesh_dirty(MesoGroup* g, MesoQuad* host_quads, int64_t cap, int64_t* n_quads);

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
