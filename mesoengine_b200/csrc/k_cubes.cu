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

// Doubling round K -> 2K (K = 1, 2, 4, 8, 16; in place).  Before it f = min(t, K) exactly, t = edge of the largest empty
// forward cube.  A cell with f == K has t >= K; the cube of edge K + j (0 <= j <= K) at e is the union of the eight cubes
// of edge K at e + j * {0,1}^3 (they overlap because j <= K), so it is empty iff f >= K at those eight cells (cells beyond
// the grid are empty).  That predicate is monotone in j: a binary search finds the largest j, and f becomes K + j =
// min(t, 2K).  Raising a cell keeps it >= K, so reading neighbours another thread has already raised changes nothing.
// Five rounds instead of 31 relaxation passes; same table (the relaxation form is what oracle/orc_raymarch.c builds).
__global__ void __launch_bounds__(256) cube_cell_double_kernel(DVolume v, uint8_t* f, int64_t ncells, int K) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 8 * ncells) return;
  const int o = (int)(g / ncells);
  const int64_t i = g - (int64_t)o * ncells;
  uint8_t* fo = f + (size_t)o * ncells;
  if (fo[i] != K) return;
  const int x = (int)(i % v.ddims[0]), y = (int)((i / v.ddims[0]) % v.ddims[1]), z = (int)(i / ((int64_t)v.ddims[0] * v.ddims[1]));
  const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
  auto ok = [&](int j) {
#pragma unroll
    for (int q = 1; q < 8; q++) {
      const int nx = x + ((q & 1) ? sx * j : 0), ny = y + ((q & 2) ? sy * j : 0), nz = z + ((q & 4) ? sz * j : 0);
      if ((unsigned)nx >= (unsigned)v.ddims[0] || (unsigned)ny >= (unsigned)v.ddims[1] || (unsigned)nz >= (unsigned)v.ddims[2]) continue;
      if (fo[(size_t)nx + (size_t)v.ddims[0] * ((size_t)ny + (size_t)v.ddims[1] * (size_t)nz)] < K) return false;
    }
    return true;
  };
  int lo = 0, hi = K;            // ok(lo) holds (j = 0 is the cell itself), find the largest j <= K with ok(j)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ok(mid)) lo = mid; else hi = mid - 1;
  }
  fo[i] = (uint8_t)(K + lo);
}

// brick (bx, by, bz) in grid brick coordinates; outside the grid = empty
__device__ __forceinline__ bool brick_present(const DVolume& v, int bx, int by, int bz) {
  if ((unsigned)bx >= (unsigned)(v.dims[0] * 16) || (unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return false;
  const int64_t ci = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
  const int lx = bx & 15, ly = by & 15, lz = bz & 15;
  const unsigned long long w = __ldg(&v.of[(size_t)ci * 64 + lz * 4 + (ly >> 2)]).x;
  return (w >> (lx + 16 * (ly & 3))) & 1ull;
}

// presence bits of the ten bricks gx0 .. gx0 + 9 of brick row (by, bz), grid brick coordinates; outside the grid = absent.
// A row of ten bricks spans at most two chunks: two 8-byte loads instead of ten brick tests.
__device__ __forceinline__ uint32_t brick_row10(const DVolume& v, int gx0, int by, int bz) {
  if ((unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return 0u;
  const int w = (bz & 15) * 4 + ((by & 15) >> 2), sh = 16 * (by & 3);
  uint64_t bits = 0;   // bit i = brick (cxa * 16 + i)
  const int cxa = gx0 >> 4;           // arithmetic shift: -1 for gx0 < 0
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int cx = cxa + k;
    if ((unsigned)cx < (unsigned)v.dims[0]) {
      const unsigned long long o = __ldg(&v.occ[(size_t)chunk_index(v, cx, by >> 4, bz >> 4) * 64 + w]);
      bits |= (uint64_t)((o >> sh) & 0xFFFFull) << (16 * k);
    }
  }
  return (uint32_t)(bits >> (gx0 - cxa * 16)) & 0x3FFu;
}

// Bricks of every NON-EMPTY 32^3 cell (the brick level of the walk is only consulted there): for each octant the largest
// t <= 4 such that the t^3 bricks starting here towards the octant are all absent.  A lane tests one cell; the warp then
// works through its non-empty cells (2 % of them at 4096^3) one at a time: the brick presence of the cell and three
// bricks around it (10 x 10 rows of 10 bits) is staged in shared memory, then every lane resolves two bricks from it.
__global__ void __launch_bounds__(256) cube_brick_kernel(DVolume v, uint16_t* __restrict__ out, int64_t ncells) {
  __shared__ uint16_t s_rows[8][100];
  const int64_t cell0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t mine = cell0 + lane;
  bool ne = false;
  if (mine < ncells) {
    const int ex = (int)(mine % v.ddims[0]), ey = (int)((mine / v.ddims[0]) % v.ddims[1]), ez = (int)(mine / ((int64_t)v.ddims[0] * v.ddims[1]));
    ne = cell_nonempty(v, ex, ey, ez);
  }
  unsigned todo = __ballot_sync(0xffffffffu, ne);
  uint16_t* rows = s_rows[warp];
  while (todo) {
    const int64_t cell = cell0 + (__ffs(todo) - 1);
    todo &= todo - 1;
    const int ex = (int)(cell % v.ddims[0]), ey = (int)((cell / v.ddims[0]) % v.ddims[1]), ez = (int)(cell / ((int64_t)v.ddims[0] * v.ddims[1]));
    __syncwarp();
    for (int i = lane; i < 100; i += 32) rows[i] = (uint16_t)brick_row10(v, ex * 4 - 3, ey * 4 - 3 + (i % 10), ez * 4 - 3 + (i / 10));
    __syncwarp();
#pragma unroll 1
    for (int l = lane; l < 64; l += 32) {
      const int X = 3 + (l & 3), Y = 3 + ((l >> 2) & 3), Z = 3 + (l >> 4);     // position in the staged 10^3 neighbourhood
      unsigned r = 0;
      if (!((rows[Z * 10 + Y] >> X) & 1u)) {
#pragma unroll 1
        for (int o = 0; o < 8; o++) {
          const int sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
          int k = 1;
          for (int t = 2; t <= 4; t++) {
            const unsigned mx = (o & 1) ? (((1u << t) - 1u) << (X - t + 1)) : (((1u << t) - 1u) << X);   // t bricks from X towards the octant
            unsigned any = 0;
            for (int z = 0; z < t; z++)
              for (int y = 0; y < t; y++) any |= rows[(Z + sz * z) * 10 + (Y + sy * y)] & mx;
            if (any) break;
            k = t;
          }
          r |= (unsigned)(k - 1) << (2 * o);
        }
      }
      const int bx = ex * 4 + (l & 3), by = ey * 4 + ((l >> 2) & 3), bz = ez * 4 + (l >> 4);
      const int64_t ci = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
      out[(size_t)ci * MESO_BLOCKS + block_bit(bx & 15, by & 15, bz & 15)] = (uint16_t)r;
    }
  }
}

// 2^3-cell cubes inside a partial brick, from the 64-bit cell mask pool_cm (bit x + 4y + 16z), bit-parallel: with E1 = ~cm
// (empty cells), "the cube of edge t at c towards the octant is empty and inside the brick" is
//     E_t = A_z(A_y(A_x(E_{t-1}))),   A_a(E) = E & shift_a(E)  (one cell towards the octant, cells beyond the brick read 0)
// and the table entry of cell c is E_2[c] + E_3[c] + E_4[c] (edge - 1).  Eight lanes per slot, one octant each; the lanes then
// exchange their three masks through shared memory and every lane packs eight cells (8 octants x 2 bits each).
__device__ __forceinline__ unsigned long long and_forward(unsigned long long e, int axis, bool neg) {
  // masks of the cells that HAVE a neighbour one step towards the octant on this axis
  const unsigned long long XL = 0x7777777777777777ull, XH = 0xEEEEEEEEEEEEEEEEull;       // x < 3 / x > 0
  const unsigned long long YL = 0x0FFF0FFF0FFF0FFFull, YH = 0xFFF0FFF0FFF0FFF0ull;       // y < 3 / y > 0
  const unsigned long long ZL = 0x0000FFFFFFFFFFFFull, ZH = 0xFFFFFFFFFFFF0000ull;       // z < 3 / z > 0
  if (axis == 0) return neg ? (e & ((e << 1) & XH)) : (e & ((e >> 1) & XL));
  if (axis == 1) return neg ? (e & ((e << 4) & YH)) : (e & ((e >> 4) & YL));
  return neg ? (e & ((e << 16) & ZH)) : (e & ((e >> 16) & ZL));
}
__global__ void __launch_bounds__(256) cube_cell2_kernel(DVolume v, uint16_t* __restrict__ out) {
  __shared__ unsigned long long s_e[32][8][3];
  const int64_t slot = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int o = threadIdx.x & 7, ls = threadIdx.x >> 3;
  const bool live = slot < (int64_t)*v.pool_count && slot < (int64_t)v.max_bricks;
  unsigned long long e = live ? ~__ldg(&v.pool_cm[slot]) : 0ull;
#pragma unroll
  for (int t = 0; t < 3; t++) {
    e = and_forward(e, 0, o & 1);
    e = and_forward(e, 1, o & 2);
    e = and_forward(e, 2, o & 4);
    s_e[ls][o][t] = e;
  }
  __syncthreads();
  if (!live) return;
  // lane o packs cells 8o .. 8o + 7
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const unsigned long long e2 = s_e[ls][q][0], e3 = s_e[ls][q][1], e4 = s_e[ls][q][2];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = 8 * o + k;
      const uint32_t val = (uint32_t)((e2 >> c) & 1ull) + (uint32_t)((e3 >> c) & 1ull) + (uint32_t)((e4 >> c) & 1ull);
      w[k >> 1] |= val << (2 * q + 16 * (k & 1));
    }
  }
  reinterpret_cast<uint4*>(out + (size_t)slot * 64)[o] = make_uint4(w[0], w[1], w[2], w[3]);
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
  static_assert(CUBE_CAP == 32, "five doubling rounds reach exactly 32");
  for (int K = 1; K < CUBE_CAP; K *= 2) {
    cube_cell_double_kernel<<<(unsigned)((8 * ncells + 255) / 256), 256, 0, lc.stream>>>(v, d_cell, ncells, K);
    (*lc.launches)++;
  }
  cube_cell_pad_kernel<<<(unsigned)((8 * npcells + 255) / 256), 256, 0, lc.stream>>>(v, d_cell, d_cellp, ncells, npcells);
  (*lc.launches)++;
  if (d_brick) {   // only entries of non-empty 32^3 cells are ever read; the rest is zero-filled (0.05 ms at 4096^3)
    cudaMemsetAsync(d_brick, 0, (size_t)v.nchunks * MESO_BLOCKS * sizeof(uint16_t), lc.stream);
    cube_brick_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, lc.stream>>>(v, d_brick, ncells);
    (*lc.launches)++;
  }
  if (d_cell2) {
    cube_cell2_kernel<<<(unsigned)(((int64_t)v.max_bricks + 31) / 32), 256, 0, lc.stream>>>(v, d_cell2);
    (*lc.launches)++;
  }
}
