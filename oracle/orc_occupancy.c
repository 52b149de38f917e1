/* orc_occupancy.c -- CPU ORACLE (test infrastructure): bit volume, erode mips, hidden-block cull,
 * instance emission.  Restates Runtimes/Voxel/Occupancy/BinaryOccupancyVolume.h:5-99,
 * Runtimes/Voxel/Chunk/Chunk.h:73-100, Runtimes/Voxel/Chunk/ChunkPool.h:385-390,438 and
 * Runtimes/Helper/VoxelMathHelper.h:73-112 bit-at-a-time, exactly as the reference walks them. */
#include "orc_internal.h"

/* VoxelMathHelper.h:98-104 */
static int out_of_bound(int x, int y, int z, int r) { return x < 0 || x >= r || y < 0 || y >= r || z < 0 || z >= r; }
/* VoxelMathHelper.h:106-112 (BoundThickness = 1) */
static int out_of_bound_thickness(int x, int y, int z, int r, int th) {
  return x < th || x >= r - th || y < th || y >= r - th || z < th || z >= r - th;
}

/* BinaryOccupancyVolume.h:45-62 (Moore) and :64-75 (Von Neumann), in the reference's order */
static const int8_t OFF26[26][3] = {
    {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1}, {-1, 1, 0}, {-1, 1, 1},
    {0, -1, -1},  {0, -1, 0},  {0, -1, 1},  {0, 0, -1},  {0, 0, 1},  {0, 1, -1}, {0, 1, 0},   {0, 1, 1},
    {1, -1, -1},  {1, -1, 0},  {1, -1, 1},  {1, 0, -1},  {1, 0, 0},  {1, 0, 1},  {1, 1, -1},  {1, 1, 0},  {1, 1, 1}};
static const int8_t OFF6[6][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};

/* BinaryOccupancyVolume.h:76-98 */
static int erode_single(const uint64_t* base, int x, int y, int z, int use26) {
  if (out_of_bound_thickness(x, y, z, ORC_CR, 1)) return 0;
  int n = use26 ? 26 : 6;
  int r = 1;
  for (int i = 0; i < n; i++) {
    const int8_t* o = use26 ? OFF26[i] : OFF6[i];
    r &= orc_getbit(base, orc_bidx(x + o[0], y + o[1], z + o[2]));
  }
  return r;
}

/* Chunk.h:73-94.  mips[d*64 .. d*64+63], d = 0..depth-1; Mip0 = block set. */
void orc_erode_mips(const uint8_t* xyz, int n, int depth, int use26, uint64_t* mips) {
  memset(mips, 0, sizeof(uint64_t) * ORC_WORDS * (size_t)depth);
  for (int i = 0; i < n; i++) orc_setbit(mips, orc_bidx(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), 1);
  for (int d = 1; d < depth; d++) {
    const uint64_t* last = mips + (size_t)(d - 1) * ORC_WORDS;
    uint64_t* cur = mips + (size_t)d * ORC_WORDS;
    for (int i = 0; i < n; i++) {
      int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
      orc_setbit(cur, orc_bidx(x, y, z), erode_single(last, x, y, z, use26));
    }
  }
}

/* Chunk.h:96-100 + ChunkPool.h:385-390,438.  Pool placement / eviction (ChunkPool.h:391-444) is
 * excluded from parity ("everything resident", SURVEY.md section 7 hard part 7): instances are
 * appended in generator order. */
int orc_emit_instances(const uint8_t* xyz, int n, const uint64_t* mips, int threshold, uint32_t chunk_index,
                       uint32_t stamp, OrcGPUBlock* out) {
  const uint64_t* m = mips + (size_t)threshold * ORC_WORDS;
  int k = 0;
  for (int i = 0; i < n; i++) {
    int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    int cull = out_of_bound(x, y, z, ORC_CR) ? 1 : orc_getbit(m, orc_bidx(x, y, z)); /* GetWithBoundaryCondition(loc,true) */
    if (cull) continue;
    out[k].ChunkIndex = chunk_index;
    out[k].BlockLocation[0] = (uint8_t)x; out[k].BlockLocation[1] = (uint8_t)y;
    out[k].BlockLocation[2] = (uint8_t)z; out[k].BlockLocation[3] = 255u;
    out[k].BlockFrameStamp = stamp;
    k++;
  }
  return k;
}
