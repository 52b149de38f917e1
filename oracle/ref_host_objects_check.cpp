/*
 * ref_host_objects_check.cpp -- the host OBJECTS of the mirror (mesoengine_b200/host/MesoHost.h: FBinaryOccupancyVolume,
 * FOccupancyHelper, FChunk, FGeneratorHelper::GenerateSphere, the GeneratorType signature) against the reference's own,
 * compiled side by side.  TEST INFRASTRUCTURE (built by `make -C oracle ref` into oracle/_ref/, run by
 * tests/test_ref_pin.py).  The reference's headers are included unmodified from /root/reference in the global namespace
 * (third-party names from oracle/ref_shim/, as in ref_driver.cpp); the mirror lives in namespace meso.
 *
 * Checked: the generator callback type, GenerateSphere's block lists (order included) on every chunk around the reference
 * sphere, all four erode mips bit for bit over all 4096 locations, bShouldVoxelOccupancyCull at depths 1 and 2 including
 * out-of-chunk locations, and Set / GetClamped / GetWithBoundaryCondition of the bit volume on in- and out-of-range
 * locations, ErodeSingleVoxel<true/false> on random volumes.
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <vector>

#include "Shape/Shape.h"
#include "Helper/Timer.h"
#include "Helper/VoxelMathHelper.h"
#include "Voxel/VoxelSceneConfig.h"
#include "Voxel/Chunk/Chunk.h"
#include "Helper/GeneratorHelper.h"

#include "../mesoengine_b200/host/MesoHost.h"

namespace lvk { bool Assert(bool cond, const char*, int, const char*, ...) { return cond; } }

static int failures = 0;
static void fail(const char* what, int a = 0, int b = 0, int c = 0) { if (failures++ < 8) std::printf("MISMATCH %s (%d, %d, %d)\n", what, a, b, c); }

int main() {
  // the plug-in signature: FChunk(ivec3, float, unsigned char, uint32_t) on both sides (ChunkManager.h:61)
  static_assert(std::is_invocable_r_v<meso::FChunk, meso::GeneratorType, meso::ivec3, float, unsigned char, uint32_t>, "GeneratorType");
  static_assert(sizeof(FBlock) == sizeof(meso::FBlock) && offsetof(FBlock, BlockLocation) == offsetof(meso::FBlock, BlockLocation) &&
                offsetof(FBlock, VolumeIndex) == offsetof(meso::FBlock, VolumeIndex), "FBlock layout");

  int chunks = 0, blocks = 0;
  for (int cz = -5; cz <= 4; cz++)
    for (int cy = -5; cy <= 4; cy++)
      for (int cx = 2; cx <= 10; cx++) {
        const FChunk r = FGeneratorHelper::GenerateSphere(glm::ivec3(cx, cy, cz), 1.0f, 16, 0);
        const meso::FChunk m = meso::FGeneratorHelper::GenerateSphere(meso::ivec3{cx, cy, cz}, 1.0f, 16, 0);
        chunks++;
        if (r.Blocks.size() != m.Blocks.size()) { fail("block count", cx, cy, cz); continue; }
        blocks += (int)r.Blocks.size();
        for (size_t i = 0; i < r.Blocks.size(); i++)
          if (r.Blocks[i].BlockLocation.x != m.Blocks[i].BlockLocation[0] || r.Blocks[i].BlockLocation.y != m.Blocks[i].BlockLocation[1] ||
              r.Blocks[i].BlockLocation.z != m.Blocks[i].BlockLocation[2]) { fail("block order", cx, cy, cz); break; }
        if (r.bIsValid() != m.bIsValid() || r.ChunkLocation.x != m.ChunkLocation.x) fail("chunk location", cx, cy, cz);
        if (r.OccupancyVolumeErodeMipmaps.size() != m.OccupancyVolumeErodeMipmaps.size()) { fail("mip count", cx, cy, cz); continue; }
        for (size_t d = 0; d < r.OccupancyVolumeErodeMipmaps.size(); d++)
          for (int z = 0; z < 16; z++) for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++)
            if (r.OccupancyVolumeErodeMipmaps[d].Get(glm::ivec3(x, y, z)) != m.OccupancyVolumeErodeMipmaps[d].Get(meso::ivec3{x, y, z})) fail("mip bit", (int)d, x + 16 * y, z);
        for (uint32_t depth = 1; depth <= 2; depth++)
          for (int z = -1; z <= 16; z++) for (int y = -1; y <= 16; y++) for (int x = -1; x <= 16; x++)
            if (r.bShouldVoxelOccupancyCull(glm::ivec3(x, y, z), depth) != m.bShouldVoxelOccupancyCull(meso::ivec3{x, y, z}, depth)) fail("cull", x, y, z);
      }

  // the bit volume and single-voxel erosion on random content, including clamped / out-of-range locations
  uint32_t s = 2463534242u;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
  for (uint32_t R : {16u, 8u, 5u}) {
    FBinaryOccupancyVolume rv(R); meso::FBinaryOccupancyVolume mv(R);
    for (int i = 0; i < 6000; i++) {
      const int x = (int)(rnd() % (R + 6)) - 3, y = (int)(rnd() % (R + 6)) - 3, z = (int)(rnd() % (R + 6)) - 3;
      const bool v = (rnd() & 3) != 0;
      rv.Set(v, glm::ivec3(x, y, z)); mv.Set(v, meso::ivec3{x, y, z});
    }
    for (int z = -3; z < (int)R + 3; z++) for (int y = -3; y < (int)R + 3; y++) for (int x = -3; x < (int)R + 3; x++) {
      if (rv.GetClamped(glm::ivec3(x, y, z)) != mv.GetClamped(meso::ivec3{x, y, z})) fail("GetClamped", x, y, z);
      for (bool b : {false, true})
        if (rv.GetWithBoundaryCondition(glm::ivec3(x, y, z), b) != mv.GetWithBoundaryCondition(meso::ivec3{x, y, z}, b)) fail("GetWithBoundaryCondition", x, y, z);
    }
    for (int z = 0; z < (int)R; z++) for (int y = 0; y < (int)R; y++) for (int x = 0; x < (int)R; x++) {
      if (FOccupancyHelper::ErodeSingleVoxel<true>(rv, glm::ivec3(x, y, z)) != meso::FOccupancyHelper::ErodeSingleVoxel<true>(mv, meso::ivec3{x, y, z})) fail("Erode26", x, y, z);
      if (FOccupancyHelper::ErodeSingleVoxel<false>(rv, glm::ivec3(x, y, z)) != meso::FOccupancyHelper::ErodeSingleVoxel<false>(mv, meso::ivec3{x, y, z})) fail("Erode6", x, y, z);
    }
  }
  std::printf("%s: %d chunks, %d blocks, %d failures\n", failures ? "MISMATCH" : "ok", chunks, blocks, failures);
  return failures ? 1 : 0;
}
