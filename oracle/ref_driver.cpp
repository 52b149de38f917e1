/*
 * ref_driver.cpp -- C entry points over the REFERENCE'S OWN hot-path headers.  TEST INFRASTRUCTURE.
 *
 * The headers named below are compiled unmodified from where they lie under /root/reference (never copied into this
 * repository); the output goes to oracle/_ref/libmeso_ref.so (git-ignored, travels to the GPU box with the snapshot).
 * Third-party headers the reference does not vendor (glm 0.9.9.8, Boost 1.83, minilog, GLFW) are
 * replaced by the stand-ins in oracle/ref_shim/ -- component-wise vector arithmetic, bit storage, a linear nearest
 * scan -- so every line of arithmetic *of the reference itself* below (hash, noise, displacement, both SDF
 * generators, index helpers, erosion, mips, the hidden-block test, importance, the view-cone set, the Fibonacci
 * bake directions, chunk re-centring, canonical comparators) is the reference's code, executed.
 *
 * What this pins: oracle/orc_sdf.c, orc_occupancy.c, orc_resident.c and parts of orc_camera.c
 * (tests/test_ref_pin.py, and the golden vectors tools/gen_golden_from_ref_build.py writes into tests/golden/).
 * The per-pixel result (GLSL in Samples/SimpleVoxel.cpp:72-224) is handled next door in ref_glsl_driver.cpp.
 * What cannot be pinned this way: the camera matrices (Cookbook Camera.h is un-vendored), FChunkPool placement
 * (threads + LVK buffers).
 *
 * Only tests/, tools/gen_golden_from_ref_build.py and __graft_entry__.build() touch this file or its output.
 */
#include <cstdint>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <algorithm>

/* reference headers, in the order the reference's own translation units see them */
#include "Shape/Shape.h"              /* pulls GLFW + LVK, which Timer.h expects to be there already */
#include "Helper/Timer.h"             /* FTimer (ChunkManagerHelper.h:197 relies on it being included first) */
#include "Helper/VoxelMathHelper.h"
#include "Voxel/VoxelSceneConfig.h"
#include "Voxel/Chunk/Chunk.h"
#include "Voxel/Chunk/ChunkManagerHelper.h"
#include "Helper/GeneratorHelper.h"
#include "Helper/Comparator.h"
#include "Shape/TriplePlanarCube.h"

/* ---- the drop-in boundary against the reference's own declarations (INTEGRATION.md "The binding itself") ---------
 * include/meso_cuda.h next to the reference's record types: the reinterpret_casts the binding relies on are layout-exact.
 * (glm's vector types come from ref_shim/; like glm's they are plain packed scalars.) */
#include "../include/meso_cuda.h"
#define SAME_LAYOUT(RefT, MesoT, rf, mf) \
  static_assert(offsetof(RefT, rf) == offsetof(MesoT, mf) && sizeof(RefT::rf) == sizeof(MesoT::mf), #RefT "::" #rf " vs " #MesoT "::" #mf)
static_assert(sizeof(FGPUBlock) == sizeof(MesoGPUBlock) && sizeof(FGPUChunk) == sizeof(MesoGPUChunk), "record sizes");
static_assert(sizeof(FGPUUniformCamera) == sizeof(MesoGPUUniformCamera) && sizeof(FGPUUniformSceneConfig) == sizeof(MesoGPUUniformSceneConfig), "uniform sizes");
SAME_LAYOUT(FGPUBlock, MesoGPUBlock, ChunkIndex, ChunkIndex);
SAME_LAYOUT(FGPUBlock, MesoGPUBlock, BlockLocation, BlockLocation);
SAME_LAYOUT(FGPUBlock, MesoGPUBlock, BlockFrameStamp, BlockFrameStamp);
SAME_LAYOUT(FGPUChunk, MesoGPUChunk, ChunkLocation, ChunkLocation);
SAME_LAYOUT(FGPUChunk, MesoGPUChunk, ChunkFrameStamp, ChunkFrameStamp);
SAME_LAYOUT(FGPUUniformCamera, MesoGPUUniformCamera, Projection, Projection);
SAME_LAYOUT(FGPUUniformCamera, MesoGPUUniformCamera, View, View);
SAME_LAYOUT(FGPUUniformCamera, MesoGPUUniformCamera, CameraChunkLocation, CameraChunkLocation);
SAME_LAYOUT(FGPUUniformCamera, MesoGPUUniformCamera, SubCameraLocation, SubCameraLocation);
SAME_LAYOUT(FGPUUniformSceneConfig, MesoGPUUniformSceneConfig, BlockSize, BlockSize);
SAME_LAYOUT(FGPUUniformSceneConfig, MesoGPUUniformSceneConfig, BlockResolution, BlockResolution);
SAME_LAYOUT(FGPUUniformSceneConfig, MesoGPUUniformSceneConfig, ChunkSize, ChunkSize);
SAME_LAYOUT(FGPUUniformSceneConfig, MesoGPUUniformSceneConfig, ChunkResolution, ChunkResolution);
/* std::pair<float, ivec3> (ChunkManagerHelper.h:76) is what meso_select_view_chunks returns as MesoChunkCandidate */
static_assert(sizeof(FChunkManageHelper::FTempChunkDataType) == sizeof(MesoChunkCandidate), "candidate size");
#undef SAME_LAYOUT

namespace lvk {
/* LVK.h:52 declares it, LVK.cpp defines it; LVK_ASSERT in inline members refers to it in debug builds. */
bool Assert(bool cond, const char*, int, const char*, ...) { return cond; }
}  // namespace lvk

extern "C" {

/* FVoxelMathHelper::Hash<double>(dvec3)  (VoxelMathHelper.h:30-33) */
double ref_hash3(double x, double y, double z) { return FVoxelMathHelper::Hash<double>(glm::dvec3(x, y, z)); }

/* FGeneratorHelper::noised<double>  (GeneratorHelper.h:19-56); out = {dx, dy, dz, value} (.yzwx swizzle) */
void ref_noised(const double x[3], double out4[4]) {
  glm::dvec4 r = FGeneratorHelper::noised<double>(glm::dvec3(x[0], x[1], x[2]));
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

/* FGeneratorHelper::displacement<double>  (GeneratorHelper.h:58-88) */
double ref_displacement(const double p[3]) { return FGeneratorHelper::displacement<double>(glm::dvec3(p[0], p[1], p[2])); }

static FChunk generate(int kind, const int32_t loc[3], float block_size, int chunk_res) {
  glm::ivec3 L(loc[0], loc[1], loc[2]);
  if (kind == 0) return FGeneratorHelper::GenerateSphere(L, block_size, (unsigned char)chunk_res, 0u);
  return FGeneratorHelper::TestGenerator(L, block_size, (unsigned char)chunk_res, 0u);
}

/* FGeneratorHelper::GenerateSphere (kind 0, GeneratorHelper.h:121-150) / TestGenerator (kind 1, :92-120), then
 * FChunk::CalculateOccupancyErodeMipmaps(res, depth) (Chunk.h:73-94) as ChunkManager.h:169 calls it.
 * out_xyz: 3 bytes per block in FChunk::Blocks order; mips (may be NULL): depth * res^3 bytes, one per voxel at
 * FVoxelMathHelper::Convert3DTo1D; cull (may be NULL): per block, FChunk::bShouldVoxelOccupancyCull(loc, threshold)
 * (Chunk.h:96-100; ChunkPool.h:388 passes 1).  Returns the block count. */
int ref_generate_chunk(int kind, const int32_t loc[3], float block_size, int chunk_res, int depth, int cull_threshold,
                       uint8_t* out_xyz, uint8_t* mips, uint8_t* cull) {
  FChunk c = generate(kind, loc, block_size, chunk_res);
  c.CalculateOccupancyErodeMipmaps((uint32_t)chunk_res, (uint32_t)depth);
  const int n = (int)c.Blocks.size();
  for (int i = 0; i < n; ++i) {
    out_xyz[3 * i + 0] = c.Blocks[i].BlockLocation.x;
    out_xyz[3 * i + 1] = c.Blocks[i].BlockLocation.y;
    out_xyz[3 * i + 2] = c.Blocks[i].BlockLocation.z;
  }
  const int r = chunk_res;
  if (mips)
    for (int d = 0; d < depth; ++d)
      for (int z = 0; z < r; ++z)
        for (int y = 0; y < r; ++y)
          for (int x = 0; x < r; ++x) {
            uint32_t idx = FVoxelMathHelper::Convert3DTo1D(glm::ivec3(x, y, z), glm::ivec3(r, r, r));
            mips[(size_t)d * r * r * r + idx] = c.OccupancyVolumeErodeMipmaps[d].Get(glm::ivec3(x, y, z)) ? 1 : 0;
          }
  if (cull)
    for (int i = 0; i < n; ++i)
      cull[i] = c.bShouldVoxelOccupancyCull(glm::ivec3(c.Blocks[i].BlockLocation), (uint32_t)cull_threshold) ? 1 : 0;
  return n;
}

/* Mips and cull flags of an arbitrary block list (erosion edge cases: blocks on the chunk border, single holes). */
void ref_erode_blocks(const uint8_t* xyz, int n, int chunk_res, int depth, int cull_threshold, uint8_t* mips, uint8_t* cull) {
  FChunk c;
  c.ChunkLocation = glm::ivec3(0, 0, 0);
  for (int i = 0; i < n; ++i)
    c.AddBlock(FBlock{.ChunkIndex = 0, .BlockLocation = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, .VolumeIndex = 0});
  c.CalculateOccupancyErodeMipmaps((uint32_t)chunk_res, (uint32_t)depth);
  const int r = chunk_res;
  for (int d = 0; d < depth; ++d)
    for (int z = 0; z < r; ++z)
      for (int y = 0; y < r; ++y)
        for (int x = 0; x < r; ++x)
          mips[(size_t)d * r * r * r + FVoxelMathHelper::Convert3DTo1D(glm::ivec3(x, y, z), glm::ivec3(r, r, r))] =
              c.OccupancyVolumeErodeMipmaps[d].Get(glm::ivec3(x, y, z)) ? 1 : 0;
  if (cull)
    for (int i = 0; i < n; ++i) cull[i] = c.bShouldVoxelOccupancyCull(glm::ivec3(c.Blocks[i].BlockLocation), (uint32_t)cull_threshold) ? 1 : 0;
}

/* FOccupancyHelper::Get26Offsets / Get6Offsets (BinaryOccupancyVolume.h:45-75) */
int ref_erode_offsets(int use26, int32_t* out_xyz) {
  std::vector<glm::ivec3> o = use26 ? FOccupancyHelper::Get26Offsets() : FOccupancyHelper::Get6Offsets();
  for (size_t i = 0; i < o.size(); ++i) { out_xyz[3 * i] = o[i].x; out_xyz[3 * i + 1] = o[i].y; out_xyz[3 * i + 2] = o[i].z; }
  return (int)o.size();
}

/* index helpers (VoxelMathHelper.h:73-112): out = {Convert3DTo1D, Convert3DTo1DClamped, bIsOutOfBound, bIsOutOfBoundThickness(1)} */
void ref_index_helpers(const int32_t loc[3], const int32_t res[3], uint32_t out4[4]) {
  glm::ivec3 L(loc[0], loc[1], loc[2]), R(res[0], res[1], res[2]);
  out4[0] = FVoxelMathHelper::Convert3DTo1D(L, R);
  out4[1] = FVoxelMathHelper::Convert3DTo1DClamped(L, R);
  out4[2] = FVoxelMathHelper::bIsOutOfBound(L, R) ? 1u : 0u;
  out4[3] = FVoxelMathHelper::bIsOutOfBoundThickness(L, R, 1u) ? 1u : 0u;
}

/* FVoxelMathHelper::ConvertToChunkLocation<float> (VoxelMathHelper.h:16-22) */
void ref_convert_to_chunk_location(const float pos[3], float chunk_size, float fract_out[3], int32_t chunk_out[3]) {
  auto [f, c] = FVoxelMathHelper::ConvertToChunkLocation<float>(glm::vec3(pos[0], pos[1], pos[2]), chunk_size);
  fract_out[0] = f.x; fract_out[1] = f.y; fract_out[2] = f.z;
  chunk_out[0] = c.x; chunk_out[1] = c.y; chunk_out[2] = c.z;
}

/* FVoxelMathHelper::GetFibonacciSphere<T> (VoxelMathHelper.h:49-71) */
void ref_fibonacci_sphere_f32(uint32_t samples, int normalize, float* out_xyz) {
  auto p = FVoxelMathHelper::GetFibonacciSphere<float>(samples, normalize != 0);
  for (size_t i = 0; i < p.size(); ++i) { out_xyz[3 * i] = p[i].x; out_xyz[3 * i + 1] = p[i].y; out_xyz[3 * i + 2] = p[i].z; }
}
void ref_fibonacci_sphere_f64(uint32_t samples, int normalize, double* out_xyz) {
  auto p = FVoxelMathHelper::GetFibonacciSphere<double>(samples, normalize != 0);
  for (size_t i = 0; i < p.size(); ++i) { out_xyz[3 * i] = p[i].x; out_xyz[3 * i + 1] = p[i].y; out_xyz[3 * i + 2] = p[i].z; }
}

/* FImportanceComputeInfo (ChunkManagerHelper.h:22-70) */
float ref_chunk_importance(const int32_t cam_chunk[3], const float fwd[3], const int32_t loc[3]) {
  FImportanceComputeInfo info{.CameraChunk = {cam_chunk[0], cam_chunk[1], cam_chunk[2]}, .CameraForwardVector = {fwd[0], fwd[1], fwd[2]}};
  return info.CalculateChunkImportance(glm::ivec3(loc[0], loc[1], loc[2]));
}
float ref_block_importance(const int32_t cam_chunk[3], const float fwd[3], const int32_t chunk[3], const uint8_t block[3],
                           uint32_t chunk_resolution) {
  FImportanceComputeInfo info{.CameraChunk = {cam_chunk[0], cam_chunk[1], cam_chunk[2]}, .CameraForwardVector = {fwd[0], fwd[1], fwd[2]}};
  return info.CalculateBlockImportance(glm::ivec3(chunk[0], chunk[1], chunk[2]), glm::u8vec3(block[0], block[1], block[2]), chunk_resolution);
}

typedef struct { float Importance; int32_t Offset[3]; } RefChunkCandidate;

/* FChunkManageHelper::GetDesiredShowChunkLocationByView (mode 0, ChunkManagerHelper.h:89-150) / ...Simple (mode 1,
 * :151-190).  Candidates in the priority queue's own pop order (ties as libstdc++'s heap leaves them). */
int64_t ref_select_view_chunks(const float fwd[3], uint32_t forward_load, uint32_t backward_load, float view_angle_deg,
                               int mode, RefChunkCandidate* out, int64_t cap) {
  FVoxelSceneConfig cfg;
  cfg.ViewForwardLoadChunkSize = forward_load;
  cfg.ViewBackwardLoadChunkSize = backward_load;
  cfg.ViewChunkAngle = view_angle_deg;
  glm::vec3 F(fwd[0], fwd[1], fwd[2]);
  auto q = mode == 0 ? FChunkManageHelper::GetDesiredShowChunkLocationByView(F, cfg)
                     : FChunkManageHelper::GetDesiredShowChunkLocationSimple(F, cfg);
  int64_t n = 0;
  while (!q.empty()) {
    if (n < cap) { out[n].Importance = q.top().first; out[n].Offset[0] = q.top().second.x; out[n].Offset[1] = q.top().second.y; out[n].Offset[2] = q.top().second.z; }
    ++n;
    q.pop();
  }
  return n;
}

/* FChunkManageHelper::BakeVisibilityByView (ChunkManagerHelper.h:192-230) + TNearestMap::Query (NearestMap.h:33-49):
 * bakes `samples` views, queries the one nearest to q and returns the size of its queue and its first candidate. */
int64_t ref_bake_and_query(uint32_t samples, uint32_t forward_load, uint32_t backward_load, float view_angle_deg,
                           const float q[3], RefChunkCandidate* first) {
  FVoxelSceneConfig cfg;
  cfg.ViewForwardLoadChunkSize = forward_load;
  cfg.ViewBackwardLoadChunkSize = backward_load;
  cfg.ViewChunkAngle = view_angle_deg;
  auto map = FChunkManageHelper::BakeVisibilityByView(cfg, samples, false);
  auto& queue = map.Query(glm::vec3(q[0], q[1], q[2]));
  if (first && !queue.empty()) { first->Importance = queue.top().first; first->Offset[0] = queue.top().second.x; first->Offset[1] = queue.top().second.y; first->Offset[2] = queue.top().second.z; }
  return (int64_t)queue.size();
}

/* TNearestMap<uint32_t> (NearestMap.h) over n directions: index of the stored direction nearest to q */
uint32_t ref_nearest_direction(const float* dirs, uint32_t n, const float q[3]) {
  TNearestMap<uint32_t> m;
  for (uint32_t i = 0; i < n; ++i) m.Insert(glm::vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), i);
  return m.Query(glm::vec3(q[0], q[1], q[2]));
}

/* FChunkManageHelper::TruncateFrameStamp (ChunkManagerHelper.h:232-236) */
uint32_t ref_truncate_frame_stamp(uint64_t stamp) { return FChunkManageHelper::TruncateFrameStamp(stamp); }

/* FIVec3Comparator (Comparator.h:15-23): sorts n ivec3 in place */
void ref_sort_ivec3(int32_t* xyz, int64_t n) {
  std::vector<glm::ivec3> v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[(size_t)i] = glm::ivec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  std::sort(v.begin(), v.end(), FIVec3Comparator());
  for (int64_t i = 0; i < n; ++i) { xyz[3 * i] = v[(size_t)i].x; xyz[3 * i + 1] = v[(size_t)i].y; xyz[3 * i + 2] = v[(size_t)i].z; }
}

/* FTriplePlanarCube::GetGPUMeshes (TriplePlanarCube.h:33-46): the fan's index list */
int ref_triplanar_indices(uint16_t* out, int cap) {
  FTriplePlanarCube cube;
  auto m = cube.GetGPUMeshes(1.0f);
  const auto& idx = std::get<1>(m);
  for (int i = 0; i < (int)idx.size() && i < cap; ++i) out[i] = idx[(size_t)i];
  return (int)std::get<2>(m);
}

/* Record sizes / field offsets and FVoxelSceneConfig defaults as the reference declares them
 * (Block.h:14-26, Chunk.h:27-31, GPUStructures.h:13-41, VoxelSceneConfig.h:20-50). */
void ref_layouts(uint32_t out[16]) {
  out[0] = (uint32_t)sizeof(FGPUBlock);
  out[1] = (uint32_t)offsetof(FGPUBlock, ChunkIndex);
  out[2] = (uint32_t)offsetof(FGPUBlock, BlockLocation);
  out[3] = (uint32_t)offsetof(FGPUBlock, BlockFrameStamp);
  out[4] = (uint32_t)sizeof(FGPUChunk);
  out[5] = (uint32_t)offsetof(FGPUChunk, ChunkLocation);
  out[6] = (uint32_t)offsetof(FGPUChunk, ChunkFrameStamp);
  out[7] = (uint32_t)sizeof(FGPUUniformCamera);
  out[8] = (uint32_t)offsetof(FGPUUniformCamera, View);
  out[9] = (uint32_t)offsetof(FGPUUniformCamera, CameraChunkLocation);
  out[10] = (uint32_t)offsetof(FGPUUniformCamera, SubCameraLocation);
  out[11] = (uint32_t)sizeof(FGPUUniformSceneConfig);
  FGPUChunk c;
  out[12] = (uint32_t)c.ChunkLocation.x;  /* INT_MAX = invalid */
  FGPUBlock b;
  out[13] = b.ChunkIndex;
  out[14] = b.BlockLocation.w;
  out[15] = 0;
}
void ref_scene_config_defaults(double out[14]) {
  FVoxelSceneConfig c;
  out[0] = c.BlockResolution; out[1] = c.BlockSize; out[2] = c.ChunkResolution; out[3] = c.MaxBlockCount;
  out[4] = c.MaxChunkCount; out[5] = c.BakeVisibilityViewNum; out[6] = c.ViewForwardLoadChunkSize;
  out[7] = c.ViewBackwardLoadChunkSize; out[8] = c.MaxUnsyncedLoadChunkCount; out[9] = c.ViewChunkAngle;
  out[10] = c.ChunkOccupancyDepth; out[11] = c.ChunkInnerVoxelCullDepthThreshold; out[12] = c.GetChunkSize();
  out[13] = c.MaxChunkCheckTimes;
}

}  /* extern "C" */
