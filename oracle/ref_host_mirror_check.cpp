/*
 * ref_host_mirror_check.cpp -- the C++ host mirror (mesoengine_b200/host/MesoHost.h) against the reference's own
 * declarations, compiled side by side.  TEST INFRASTRUCTURE (built by `make -C oracle ref` into oracle/_ref/, run by
 * tests/test_ref_pin.py).  The reference's Voxel/VoxelSceneConfig.h and Helper/VoxelMathHelper.h are included unmodified
 * from /root/reference inside namespace ref (glm comes from oracle/ref_shim/), the mirror lives in namespace meso:
 * every field of FVoxelSceneConfig must exist in both with the same type, default and offset, EChunkOverrideMode must
 * have the same enumerators, and ConvertToChunkLocation must return the same bits.
 */
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <tuple>
#include <type_traits>
#include <vector>

#include <glm/glm.hpp>
#include <glm/ext.hpp>

namespace ref {
#include "Voxel/VoxelSceneConfig.h"
}  // namespace ref

#include "../mesoengine_b200/host/MesoHost.h"

static int failures = 0;
#define FIELD(f)                                                                                                        \
  do {                                                                                                                  \
    static_assert(std::is_same_v<decltype(ref::FVoxelSceneConfig::f), decltype(meso::FVoxelSceneConfig::f)>, "type of " #f); \
    if (!(r.f == m.f)) { std::printf("default of %s differs\n", #f); ++failures; }                                      \
    if (offsetof(ref::FVoxelSceneConfig, f) != offsetof(meso::FVoxelSceneConfig, f)) { std::printf("offset of %s differs\n", #f); ++failures; } \
  } while (0)

int main() {
  ref::FVoxelSceneConfig r;
  meso::FVoxelSceneConfig m;
  FIELD(BlockResolution); FIELD(BlockSize); FIELD(ChunkResolution); FIELD(MaxBlockCount); FIELD(MaxVolumeCount);
  FIELD(MaxChunkCount); FIELD(MaxEmptyChunkCount); FIELD(MaxChunkCheckTimes); FIELD(MaxEmptyChunkCheckTimes);
  FIELD(MaxBlockCheckTimes); FIELD(BakeVisibilityViewNum); FIELD(ViewForwardLoadChunkSize); FIELD(ViewBackwardLoadChunkSize);
  FIELD(MaxSyncedLoadChunkCount); FIELD(MaxUnsyncedLoadChunkCount); FIELD(ChunkTaskPerCore); FIELD(ViewChunkAngle);
  FIELD(ChunkOccupancyDepth); FIELD(ChunkInnerVoxelCullDepthThreshold);
  static_assert(sizeof(ref::FVoxelSceneConfig) == sizeof(meso::FVoxelSceneConfig), "FVoxelSceneConfig size");
  static_assert(std::is_same_v<std::underlying_type_t<ref::EChunkOverrideMode>, std::underlying_type_t<meso::EChunkOverrideMode>>, "enum base");
  static_assert((int)ref::EChunkOverrideMode::FindLess == (int)meso::EChunkOverrideMode::FindLess &&
                (int)ref::EChunkOverrideMode::FindMin == (int)meso::EChunkOverrideMode::FindMin &&
                (int)ref::EChunkOverrideMode::OverrideMin == (int)meso::EChunkOverrideMode::OverrideMin, "EChunkOverrideMode");
  if ((int)r.ChunkOverrideMode != (int)m.ChunkOverrideMode) { std::printf("default of ChunkOverrideMode differs\n"); ++failures; }
  if (offsetof(ref::FVoxelSceneConfig, ChunkOverrideMode) != offsetof(meso::FVoxelSceneConfig, ChunkOverrideMode)) { std::printf("offset of ChunkOverrideMode differs\n"); ++failures; }
  if (r.GetChunkSize() != m.GetChunkSize()) { std::printf("GetChunkSize differs\n"); ++failures; }

  // FVoxelMathHelper::ConvertToChunkLocation<float> (VoxelMathHelper.h:16-22) vs the mirror's, bit for bit
  uint32_t s = 12345u;
  int checked = 0;
  for (int i = 0; i < 20000; ++i) {
    float p[3];
    for (float& c : p) { s = s * 1664525u + 1013904223u; c = ((int32_t)(s >> 8) % 4000000) / 1000.0f - 2000.0f; }
    if (i % 7 == 0) p[i % 3] = (float)(((int)(s >> 20) % 64) - 32) * 16.0f;      // exact chunk boundaries
    if (i % 11 == 0) p[(i + 1) % 3] = -1e-7f;
    auto [rf, rc] = ref::FVoxelMathHelper::ConvertToChunkLocation<float>(glm::vec3(p[0], p[1], p[2]), 16.0f);
    meso::vec3 mf; meso::ivec3 mc;
    meso::FVoxelMathHelper::ConvertToChunkLocation(meso::vec3{p[0], p[1], p[2]}, 16.0f, mf, mc);
    const float a[3] = {rf.x, rf.y, rf.z}, b[3] = {mf.x, mf.y, mf.z};
    if (std::memcmp(a, b, sizeof a) != 0 || rc.x != mc.x || rc.y != mc.y || rc.z != mc.z) {
      if (failures < 5) std::printf("ConvertToChunkLocation(%g, %g, %g) differs\n", p[0], p[1], p[2]);
      ++failures;
    }
    ++checked;
  }
  std::printf("%s: %d fields, %d positions, %d failures\n", failures ? "MISMATCH" : "ok", 20, checked, failures);
  return failures ? 1 : 0;
}
