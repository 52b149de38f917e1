// k_cubes.cu -- per-octant FORWARD CUBES for the raymarch walk.
//
// Chosen with the calibrated step model (tools/step_model.py, profiles/README.md) and measured on B200 (profiles/
// r2_rm_ab_*.jsonl): reading, at a level of the walk, the edge of the largest empty cube that starts at the current unit
// and extends towards the ray's octant takes the walk on the 4096^3 scene from 15.9 to 12.0 steps per ray with cell and
// brick cubes (the default; +14 % Mrays/s over the distance field) and to 10.7 with 2^3-cell cubes on top (not faster: two
// more bytes per step).  oracle/orc_raymarch.c (ORC_DDA_MODEL) states the same cubes on the CPU;
// tests/test_zz_gpu_cubes.py compares the tables and the frames.
//
// A ray only ever needs what lies ahead of it, so the boxes are anchored at the current unit and grow towards the octant
// (step_x < 0) | (step_y < 0) << 1 | (step_z < 0) << 2.  Outside the grid counts as empty (the walk clamps its steps to
// the grid).  Three tables, one per level that skips:
//   cell  [8][ncells]        u8   edge in 32^3 cells, 1 .. 32; 0 = cell not empty          (replaces df + probe-ahead)
//   brick [nchunks * 4096]   u16  2 bits per octant, edge - 1 in bricks (1..4)              (empty bricks of non-empty cells)
//   cell2 [max_bricks * 64]  u16  2 bits per octant, edge - 1 in 2^3 cells (1..4), inside the brick
// The reference has no counterpart (its per-pixel visibility is a rasterised instanced draw, SimpleVoxel.cpp:352-398).
#include "meso_internal.cuh"

#define CUBE_CAP (MESO_DF_K + 1)

__device__ __forceinline__ bool cell_nonempty(const DVolume& v, int x, int y, int z) {
  const int64_t ci = chunk_index(v, x >> 2, y >> 2, z >> 2);
  return (v.cells[ci] >> ((x & 3) + 4 * (y & 3) + 16 * (z & 3))) & 1ull;
}

// cell[o][e] = 1 for empty cells, 0 otherwise
__global__ void __launch_bounds__(256) cube_cell_init_kernel(DVolume v, uint8_t* __restrict__ f, int64_t ncells) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncells) return;
  const int x = (int)(i % v.ddims[0]), y = (int)((i / v.ddims[0]) % v.ddims[1]), z = (int)(i / ((int64_t)v.ddims[0] * v.ddims[1]));
  const uint8_t e = cell_nonempty(v, x, y, z) ? 0 : 1;
#pragma unroll
  for (int o = 0; o < 8; o++) f[(size_t)o * ncells + i] = e;
}

// Round t raises t-1 -> t: the cube of edge t at e is empty iff the cubes of edge t-1 at e and at its seven forward
// neighbours are (neighbours beyond the grid are empty).  In place: a neighbour already raised to t still reads >= t-1.
__global__ void __launch_bounds__(256) cube_cell_pass_kernel(DVolume v, uint8_t* f, int64_t ncells, int t) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 8 * ncells) return;
  const int o = (int)(g / ncells);
  const int64_t i = g - (int64_t)o * ncells;
  uint8_t* fo = f + (size_t)o * ncells;
  if (fo[i] != t - 1) return;
  const int x = (int)(i % v.ddims[0]), y = (int)((i / v.ddims[0]) % v.ddims[1]), z = (int)(i / ((int64_t)v.ddims[0] * v.ddims[1]));
  const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
#pragma unroll
  for (int q = 1; q < 8; q++) {
    const int nx = x + ((q & 1) ? sx : 0), ny = y + ((q & 2) ? sy : 0), nz = z + ((q & 4) ? sz : 0);
    if ((unsigned)nx >= (unsigned)v.ddims[0] || (unsigned)ny >= (unsigned)v.ddims[1] || (unsigned)nz >= (unsigned)v.ddims[2]) continue;
    if (fo[(size_t)nx + (size_t)v.ddims[0] * ((size_t)ny + (size_t)v.ddims[1] * (size_t)nz)] < t - 1) return;
  }
  fo[i] = (uint8_t)t;
}

// brick (bx, by, bz) in grid brick coordinates; outside the grid = empty
__device__ __forceinline__ bool brick_present(const DVolume& v, int bx, int by, int bz) {
  if ((unsigned)bx >= (unsigned)(v.dims[0] * 16) || (unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return false;
  const int64_t ci = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
  const int lx = bx & 15, ly = by & 15, lz = bz & 15;
  const unsigned long long w = __ldg(&v.of[(size_t)ci * 64 + lz * 4 + (ly >> 2)]).x;
  return (w >> (lx + 16 * (ly & 3))) & 1ull;
}

// One thread per brick of every NON-EMPTY 32^3 cell (the brick level of the walk is only consulted there): for each
// octant the largest t <= 4 such that the t^3 bricks starting here towards the octant are all absent.
__global__ void __launch_bounds__(256) cube_brick_kernel(DVolume v, uint16_t* __restrict__ out, int64_t ncells) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ncells * 64) return;
  const int64_t cell = g >> 6;
  const int l = (int)(g & 63);
  const int ex = (int)(cell % v.ddims[0]), ey = (int)((cell / v.ddims[0]) % v.ddims[1]), ez = (int)(cell / ((int64_t)v.ddims[0] * v.ddims[1]));
  if (!cell_nonempty(v, ex, ey, ez)) return;
  const int bx = ex * 4 + (l & 3), by = ey * 4 + ((l >> 2) & 3), bz = ez * 4 + (l >> 4);
  unsigned r = 0;
  if (!brick_present(v, bx, by, bz)) {
    for (int o = 0; o < 8; o++) {
      const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
      int k = 1;
      for (int t = 2; t <= 4; t++) {
        bool empty = true;
        for (int z = 0; z < t && empty; z++)
          for (int y = 0; y < t && empty; y++)
            for (int x = 0; x < t; x++) {
              if (x < t - 1 && y < t - 1 && z < t - 1) continue;      // inside the (t-1)-cube: tested in the previous round
              if (brick_present(v, bx + sx * x, by + sy * y, bz + sz * z)) { empty = false; break; }
            }
        if (!empty) break;
        k = t;
      }
      r |= (unsigned)(k - 1) << (2 * o);
    }
  }
  const int64_t ci = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
  out[(size_t)ci * MESO_BLOCKS + block_bit(bx & 15, by & 15, bz & 15)] = (uint16_t)r;
}

// One thread per 2^3 cell of every payload slot in use: cubes of empty cells inside the brick, from the cell mask pool_cm.
__global__ void __launch_bounds__(256) cube_cell2_kernel(DVolume v, uint16_t* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = g >> 6;
  if (slot >= (int64_t)*v.pool_count || slot >= (int64_t)v.max_bricks) return;
  const int c = (int)(g & 63);
  const unsigned long long cm = __ldg(&v.pool_cm[slot]);
  const int cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
  unsigned r = 0;
  if (!((cm >> c) & 1ull)) {
    for (int o = 0; o < 8; o++) {
      const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
      int k = 1;
      for (int t = 2; t <= 4; t++) {
        bool empty = true;
        for (int z = 0; z < t && empty; z++)
          for (int y = 0; y < t && empty; y++)
            for (int x = 0; x < t; x++) {
              if (x < t - 1 && y < t - 1 && z < t - 1) continue;
              const int qx = cx + sx * x, qy = cy + sy * y, qz = cz + sz * z;
              if ((unsigned)qx > 3u || (unsigned)qy > 3u || (unsigned)qz > 3u || ((cm >> (qx + 4 * qy + 16 * qz)) & 1ull)) { empty = false; break; }
            }
        if (!empty) break;
        k = t;
      }
      r |= (unsigned)(k - 1) << (2 * o);
    }
  }
  out[(size_t)slot * 64 + c] = (uint16_t)r;
}

// cellp: the cell cubes with a one-cell border that reads 255 (outside the grid)
__global__ void __launch_bounds__(256) cube_cell_pad_kernel(DVolume v, const uint8_t* __restrict__ f, uint8_t* __restrict__ fp, int64_t ncells, int64_t npcells) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 8 * npcells) return;
  const int o = (int)(g / npcells);
  const int64_t i = g - (int64_t)o * npcells;
  const int p0 = v.ddims[0] + 2, p1 = v.ddims[1] + 2;
  const int x = (int)(i % p0) - 1, y = (int)((i / p0) % p1) - 1, z = (int)(i / ((int64_t)p0 * p1)) - 1;
  uint8_t val = 255;
  if ((unsigned)x < (unsigned)v.ddims[0] && (unsigned)y < (unsigned)v.ddims[1] && (unsigned)z < (unsigned)v.ddims[2])
    val = f[(size_t)o * ncells + (size_t)x + (size_t)v.ddims[0] * ((size_t)y + (size_t)v.ddims[1] * (size_t)z)];
  fp[g] = val;
}

void launch_build_cubes(const LaunchCtx& lc, const DVolume& v, uint8_t* d_cell, uint8_t* d_cellp, uint16_t* d_brick, uint16_t* d_cell2) {
  const int64_t ncells = (int64_t)v.ddims[0] * v.ddims[1] * v.ddims[2];
  const int64_t npcells = (int64_t)(v.ddims[0] + 2) * (v.ddims[1] + 2) * (v.ddims[2] + 2);
  cube_cell_init_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, lc.stream>>>(v, d_cell, ncells);
  (*lc.launches)++;
  for (int t = 2; t <= CUBE_CAP; t++) {
    cube_cell_pass_kernel<<<(unsigned)((8 * ncells + 255) / 256), 256, 0, lc.stream>>>(v, d_cell, ncells, t);
    (*lc.launches)++;
  }
  cube_cell_pad_kernel<<<(unsigned)((8 * npcells + 255) / 256), 256, 0, lc.stream>>>(v, d_cell, d_cellp, ncells, npcells);
  (*lc.launches)++;
  if (d_brick) {   // only entries of non-empty 32^3 cells are ever read; the rest is zero-filled (0.05 ms at 4096^3)
    cudaMemsetAsync(d_brick, 0, (size_t)v.nchunks * MESO_BLOCKS * sizeof(uint16_t), lc.stream);
    cube_brick_kernel<<<(unsigned)((ncells * 64 + 255) / 256), 256, 0, lc.stream>>>(v, d_brick, ncells);
    (*lc.launches)++;
  }
  if (d_cell2) {
    cube_cell2_kernel<<<(unsigned)(((int64_t)v.max_bricks * 64 + 255) / 256), 256, 0, lc.stream>>>(v, d_cell2);
    (*lc.launches)++;
  }
}
