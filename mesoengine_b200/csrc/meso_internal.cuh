// meso_internal.cuh -- private declarations shared by the kernels and the C ABI (libmeso_b200.so).
// Whole library is compiled with -fmad=false: every fp32/fp64 operation rounds exactly as written, which is what
// makes hit voxels / faces / quads bit-identical to the CPU oracle (DESIGN.md "determinism").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/meso_cuda.h"

#define MESO_CR 16
#define MESO_BR 8
#define MESO_CV 128
#define MESO_WORDS 64
#define MESO_BLOCKS 4096
#define MESO_DF_K 31          // window half-width of the distance-field passes; stored values are 0 .. MESO_DF_K + 1

// Device view of the resident volume.  Chunk slot == linear grid index cx + dx*(cy + dy*cz); everything dense per
// chunk except brick payloads (pool, 64 B each, allocated for partial bricks only).
struct DVolume {
  int dims[3];          // chunks
  int nvox[3];          // voxels
  int origin[3];        // chunk coordinates of the grid minimum
  int64_t nchunks;
  uint64_t* occ;        // nchunks*64   block present (Mip0)
  uint64_t* full;       // nchunks*64   brick all-solid
  ulonglong2* of;       // nchunks*64   {occ, full} interleaved: one 16 B load per 64 bricks in the raymarch (derived)
  uint64_t* mips;       // nchunks*3*64 erode Mip1..3
  uint32_t* bptr;       // nchunks*4096 payload slot of partial bricks (0xFFFFFFFF otherwise)
  uint64_t* pool;       // max_bricks*8 words
  uint64_t* pool_cm;    // max_bricks: per payload slot, bit per 2^3-voxel cell of the brick (index x/2 + 4(y/2) + 16(z/2)) (derived)
  uint32_t* chunk_any;  // bit per chunk: has >=1 block
  uint32_t* chunk_full; // bit per chunk: all 4096 bricks full
  uint64_t* cells;      // per chunk: bit per 32^3-voxel cell (4x4x4 bricks), index x + 4y + 16z (derived)
  uint32_t* region_any; // bit per 512^3-voxel region (4x4x4 chunks) (derived)
  uint8_t* df;          // per 32^3 cell: Chebyshev distance (in cells, capped at MESO_DF_K + 1) to the nearest non-empty cell;
                        // 0 = non-empty.  Conservative: never larger than the true distance (derived; rebuilt whenever
                        // voxels may have been ADDED, left alone by the carve, which only removes)
  uint8_t* df_tmp;      // scratch of the separable passes
  uint32_t* words;      // nchunks*64: K1 scratch, occupancy words the per-voxel kernel has to visit (mixed words)
  uint32_t* n_words;    // device counter of `words`
  int ddims[3];         // cells per axis = dims * 4
  int rdims[3];         // regions per axis = ceil(dims / 4)
  int region_words;
  uint32_t* pool_count; // device counter: payload slots handed out so far (high-water mark; slot indices are < this)
  uint32_t* pool_free;  // max_bricks: stack of payload slots returned by evicted chunks (moving window), reused before the bump allocator
  int* pool_free_count; // entries on that stack
  uint32_t max_bricks;
  int chunk_words;      // number of u32 words in chunk_any / chunk_full
};

// Per-octant forward cubes (k_cubes.cu, meso_build_cubes; opt-in raymarch path MESO_FLAG_CUBES).  Octant of a ray =
// (step_x < 0) | (step_y < 0) << 1 | (step_z < 0) << 2; "forward" = towards that octant.  Every entry certifies an EMPTY
// cube that starts at the cell / brick / 2^3 cell and is `edge` units long on each axis (outside the grid counts as empty).
struct CubeTables {
  const uint8_t* cell;    // [8][ncells]       edge in 32^3 cells (1 .. MESO_DF_K + 1); 0 = the cell is not empty
  const uint8_t* cellp;   // [8][npcells]      the same with a one-cell border all around that reads 255 = outside the grid:
                          //                   cell (x, y, z) at (x + 1) + pd0 (y + 1) + pd01 (z + 1); the raymarch's only exit test
  int pd0, pd01;          // row / slice pitch of cellp: ddims[0] + 2, (ddims[0] + 2) (ddims[1] + 2)
  int64_t npcells;
  const uint16_t* brick;  // [nchunks * 4096]  2 bits per octant: edge - 1 in bricks (1..4); defined for empty bricks of non-empty cells
  const uint16_t* cell2;  // [max_bricks * 64] 2 bits per octant: edge - 1 in 2^3 cells (1..4, inside the brick); defined for empty cells
  int64_t ncells;
};

// MESO_LAYOUT_SLABS: where the rows of the frame live (meso_raymarch_device_slabs); slab k = rows [k * rows_per_slab, ...)
struct FrameMap {
  void* slab[MESO_MAX_SLABS];
  int rows_per_slab;
  int n_slabs;
};

struct RayStatsDev {
  unsigned long long primary, shadow, hits, steps;
  unsigned long long steps_primary, warp_slots_primary, warp_slots_shadow;
  unsigned long long level_steps[5];
};

// ---- launch wrappers (one per .cu) -------------------------------------------------------------------------
struct LaunchCtx {
  cudaStream_t stream;
  int sm_count;
  int64_t* launches;
};

void launch_voxelize(const LaunchCtx& lc, const DVolume& v, int kind, const double params[4], int granularity, int* d_overflow, int mip = 0);
void launch_volume_finalize(const LaunchCtx& lc, const DVolume& v, bool rebuild_df = true);  // derived data: of, cells, chunk/region bits, df
void launch_df_build(const LaunchCtx& lc, const DVolume& v);
void launch_coarse_masks(const LaunchCtx& lc, const DVolume& v);
void launch_scatter_payload(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_keys, const uint64_t* d_payload, int64_t n);
void launch_gather_partial(const LaunchCtx& lc, const DVolume& v, uint64_t* d_keys, uint64_t* d_payload, uint32_t* d_count);

void launch_occupancy_count(const LaunchCtx& lc, const DVolume& v, uint32_t stamp, MesoGPUChunk* d_table, uint32_t* d_counts,
                            uint32_t* d_offsets, uint64_t* d_total, uint32_t* d_block_totals /* 1024 words */);
void launch_occupancy_emit(const LaunchCtx& lc, const DVolume& v, uint32_t stamp, const uint32_t* d_counts, const uint32_t* d_offsets,
                           MesoGPUBlock* d_inst, int64_t cap_inst);

void launch_scatter_blocks(const LaunchCtx& lc, const DVolume& v, const MesoGPUChunk* d_chunks, int64_t n_chunks, const MesoGPUBlock* d_blocks,
                           int64_t n_blocks, unsigned long long* d_accepted);

void launch_raymarch(const LaunchCtx& lc, const DVolume& v, const MesoRaySetup& rs, int width, int height, uint32_t flags,
                     int rank, int world, int layout, MesoHitRecord* d_out, RayStatsDev* d_stats, uint8_t* d_touch_chunk,
                     uint8_t* d_touch_brick, int local_tile0 = 0, int local_tile_count = -1, const CubeTables* cubes = nullptr,
                     const FrameMap* slabs = nullptr);
// k_cubes.cu: (re)build the three tables for the current volume
void launch_build_cubes(const LaunchCtx& lc, const DVolume& v, uint8_t* d_cell, uint8_t* d_cellp, uint16_t* d_brick, uint16_t* d_cell2);
void launch_pack_rgba8(const LaunchCtx& lc, const MesoHitRecord* d_records, size_t n, uint32_t* d_out);
void launch_compose_tiles(const LaunchCtx& lc, const MesoHitRecord* d_tiles, int world, int width, int height, MesoHitRecord* d_frame);

// scratch of the mesh passes: brick work list (work_cap entries = every brick of the rank's chunks; partial bricks from the
// front, full bricks with a partial neighbour from the back) + its two counters; chunk list (nchunks entries) + its counter
// (counters: three adjacent words work_count, chunk_count, full_count); one bit per chunk, all-zero between calls (re-mesh:
// chunks that hold a listed brick)
struct MeshScratch { uint64_t* work; int64_t work_cap; uint32_t* work_count; uint32_t* chunk_list; uint32_t* chunk_count; uint32_t* full_count; uint32_t* chunk_mark; };
void launch_mesh(const LaunchCtx& lc, const DVolume& v, int rank, int world, const MeshScratch& ms,
                 MesoQuad* d_quads, int64_t cap, unsigned long long* d_quad_count, bool reset_count = true);
void launch_mesh_list(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_keys, uint32_t n_keys, MesoQuad* d_quads,
                      int64_t cap, unsigned long long* d_quad_count, const MeshScratch& ms, int rank = 0, int world = 1);

void launch_carve(const LaunchCtx& lc, const DVolume& v, const int32_t center[3], int32_t radius, uint64_t* d_dirty,
                  uint32_t cap_dirty, uint32_t* d_dirty_count, int* d_overflow);
void launch_expand_dirty(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_dirty, uint32_t n_dirty, uint64_t* d_keys,
                         uint32_t cap, uint32_t* d_count, uint32_t* d_mark);
void launch_flush(const LaunchCtx& lc, uint32_t* d_scratch, size_t n_words);
struct SignalTargets { unsigned* word[MESO_MAX_SLABS]; int n; };
void launch_signal(const LaunchCtx& lc, const SignalTargets& t);
void launch_wait(const LaunchCtx& lc, const unsigned* d_word, unsigned target, int* d_timeout_flag);
void launch_peek(const LaunchCtx& lc, void* d_dst_mapped, const void* d_src, size_t bytes);   // src 4-byte aligned

void launch_debug_instances(const LaunchCtx& lc, const DVolume& v, const uint32_t* d_loaded, float chunk_size, MesoGPUSimpleInstanceData* d_out,
                            int64_t cap, uint32_t* d_count);
// moving window (k_resident.cu): evict what leaves, shift what stays, see meso_stream_recentre
void launch_window_shift(const LaunchCtx& lc, DVolume& v, const int delta[3], uint32_t* d_loaded, void* d_scratch);
void launch_block_importance(const LaunchCtx& lc, const int32_t* d_chunk_loc, const uint8_t* d_block_loc, int64_t n, const int32_t cam[3], const float fwd[3],
                             uint32_t chunk_resolution, float* d_out);

// K6 (k_resident.cu, k_voxelize.cu)
int64_t resident_max_candidates(const MesoViewConfig& vc);
size_t resident_sort_scratch_bytes(const MesoViewConfig& vc);
void launch_select_view(const LaunchCtx& lc, const float fwd[3], const MesoViewConfig& vc, uint64_t* d_keys, uint32_t* d_count,
                        MesoChunkCandidate* d_out, int64_t cap);
void launch_chunk_importance(const LaunchCtx& lc, const int32_t* d_loc, int64_t n, const int32_t cam[3], const float fwd[3], float* d_out);
void launch_stream_worklist(const LaunchCtx& lc, const DVolume& v, const MesoChunkCandidate* d_cand, const uint32_t* d_count, int64_t cap,
                            const int32_t cam[3], uint32_t* d_loaded, uint32_t max_new, uint32_t* d_list, uint32_t* d_stats);
void launch_voxelize_list(const LaunchCtx& lc, const DVolume& v, int kind, const double params[4], int granularity, int* d_overflow,
                          const uint32_t* d_list, const uint32_t* d_n, uint32_t max_n);

// ---- small device helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t chunk_index(const DVolume& v, int cx, int cy, int cz) {
  return (int64_t)cx + (int64_t)v.dims[0] * ((int64_t)cy + (int64_t)v.dims[1] * (int64_t)cz);
}
__device__ __forceinline__ int block_bit(int x, int y, int z) { return x + 16 * y + 256 * z; }

// A payload slot for a brick that just became partial: a slot an evicted chunk gave back if there is one, else the next
// unused one.  The result may be >= max_bricks (pool exhausted): the caller drops the brick and calls release_bump().
// Pushes onto the free stack happen in their own kernel (evict_chunks_kernel), never concurrently with this.
__device__ __forceinline__ uint32_t alloc_payload_slot(const DVolume& v) {
  // the free stack is empty unless a window move has evicted chunks: look before popping (a plain load that every
  // allocation shares, instead of two more atomics on one hot address -- they cost the 4096^3 voxelise 0.6 ms)
  if (*reinterpret_cast<volatile int*>(v.pool_free_count) > 0) {
    const int n = atomicSub(v.pool_free_count, 1);
    if (n > 0) return v.pool_free[n - 1];
    atomicAdd(v.pool_free_count, 1);
  }
  return atomicAdd(v.pool_count, 1u);
}
__device__ __forceinline__ void release_bump(const DVolume& v) { atomicSub(v.pool_count, 1u); }
