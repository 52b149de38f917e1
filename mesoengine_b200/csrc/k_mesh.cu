// k_mesh.cu -- K3: face culling + greedy merge of exposed voxel faces into compacted quad lists.
//
// The north-star form of the reference's "mesher" (hidden-block cull + instance compaction,
// Runtimes/Voxel/Chunk/ChunkPool.h:381-445; its draw picks faces per view in the VS, Samples/SimpleVoxel.cpp:129-136).
// Neither quads nor a face-neighbour test exist in the reference: the definition is oracle/orc_mesh.c
// ("parity unpinned by reference; bit-exact vs repo oracle" after the canonical sort).
//
// Pass A (worklist): one thread per potential brick; occupied bricks that are not buried (full with six full
// neighbours) are compacted with a warp ballot + one atomic per warp.
// Pass B (mesh): one warp per brick.  Lanes 0..7 own one z-slice each (64 voxels as a u64) and build the six
// exposed-face slice masks with shifts against the neighbour slices; the 48 (direction, layer) 8x8 images are then
// merged greedily by 48 lane-tasks, counted, prefix-summed across the warp, and written behind one atomicAdd.
// HBM-bound integer work: 64 B per populated brick + 16 B per quad (+ neighbour slices, mostly L2 hits).
#include "meso_internal.cuh"

// slice z of the brick at global block coordinates (zeros outside the grid / absent, ones for full bricks)
__device__ __forceinline__ uint64_t fetch_slice(const DVolume& v, int bx, int by, int bz, int z) {
  if ((unsigned)bx >= (unsigned)(v.dims[0] * 16) || (unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return 0ull;
  const int64_t c = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
  const int b = block_bit(bx & 15, by & 15, bz & 15);
  if (!((__ldg(&v.occ[c * 64 + (b >> 6)]) >> (b & 63)) & 1ull)) return 0ull;
  if ((__ldg(&v.full[c * 64 + (b >> 6)]) >> (b & 63)) & 1ull) return ~0ull;
  return __ldg(&v.pool[(size_t)__ldg(&v.bptr[c * MESO_BLOCKS + b]) * 8 + z]);
}
__device__ __forceinline__ bool brick_full(const DVolume& v, int bx, int by, int bz) {
  if ((unsigned)bx >= (unsigned)(v.dims[0] * 16) || (unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return false;
  const int64_t c = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
  const int b = block_bit(bx & 15, by & 15, bz & 15);
  return (__ldg(&v.full[c * 64 + (b >> 6)]) >> (b & 63)) & 1ull;
}

__global__ void __launch_bounds__(256) mesh_worklist_kernel(DVolume v, int rank, int world, uint64_t* work, uint32_t* work_count) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (t < v.nchunks * MESO_BLOCKS) {
    const int64_t c = t >> 12; const int b = (int)(t & 4095);
    if ((c % world) == rank) {
      const uint64_t o = __ldg(&v.occ[c * 64 + (b >> 6)]);
      if ((o >> (b & 63)) & 1ull) {
        keep = true;
        if ((__ldg(&v.full[c * 64 + (b >> 6)]) >> (b & 63)) & 1ull) {
          const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
          const int bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);
          if (brick_full(v, bx - 1, by, bz) && brick_full(v, bx + 1, by, bz) && brick_full(v, bx, by - 1, bz) &&
              brick_full(v, bx, by + 1, bz) && brick_full(v, bx, by, bz - 1) && brick_full(v, bx, by, bz + 1))
            keep = false;  // buried: no exposed face
        }
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m) {
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(work_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) work[base + __popc(m & ((1u << lane) - 1u))] = (uint64_t)t;
  }
}

__device__ __forceinline__ uint32_t gather_col(uint64_t e, int x) {  // bits (x + 8y) -> byte with bit y
  return (uint32_t)((((e >> x) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}

// Greedy merge of one 8x8 image (row v = byte v).  EMIT=false counts only.
template <bool EMIT>
__device__ __forceinline__ int greedy_image(uint64_t img, int dir, int layer, int ox, int oy, int oz, MesoQuad* out, int64_t pos, int64_t cap) {
  int n = 0;
  const int ax = dir >> 1;
#pragma unroll 1
  for (int vv = 0; vv < 8; vv++) {
    uint32_t row = (uint32_t)(img >> (8 * vv)) & 0xFFu;
    while (row) {
      const int u0 = __ffs(row) - 1;
      const int w = __ffs(~(row >> u0)) - 1;
      const uint32_t m = ((1u << w) - 1u) << u0;
      int h = 1;
      while (vv + h < 8 && (((uint32_t)(img >> (8 * (vv + h))) & m) == m)) { img &= ~((uint64_t)m << (8 * (vv + h))); h++; }
      row &= ~m;
      if (EMIT) {
        int x, y, z;
        if (ax == 0) { x = layer; y = u0; z = vv; } else if (ax == 1) { x = u0; y = layer; z = vv; } else { x = u0; y = vv; z = layer; }
        if (pos + n < cap) {
          uint4 q;
          q.x = (uint32_t)(ox + x) | ((uint32_t)(oy + y) << 16);
          q.y = (uint32_t)(oz + z) | ((uint32_t)dir << 16) | ((uint32_t)w << 24);
          q.z = (uint32_t)h; q.w = 0u;
          reinterpret_cast<uint4*>(out)[pos + n] = q;
        }
      }
      n++;
    }
  }
  return n;
}

// Persistent, grid-stride over the work list (count read from device memory: no host round trip between the passes).
__global__ void __launch_bounds__(256) mesh_bricks_kernel(DVolume v, const uint64_t* __restrict__ work, const uint32_t* __restrict__ work_count_ptr,
                                                          uint32_t work_count_imm, MesoQuad* quads, int64_t cap, unsigned long long* quad_count) {
  __shared__ uint64_t s_e[8][6][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_work = work_count_ptr ? *work_count_ptr : work_count_imm;
  for (int64_t item = (int64_t)blockIdx.x * 8 + warp; item < n_work; item += (int64_t)gridDim.x * 8) {
  __syncwarp();
  const uint64_t key = work[item];
  const int64_t c = (int64_t)(key >> 12); const int b = (int)(key & 4095);
  const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  const int bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);

  const int z = lane & 7;
  uint64_t s = 0, nzm = 0, nzp = 0;
  if (lane < 8) s = fetch_slice(v, bx, by, bz, z);
  if (lane == 0) nzm = fetch_slice(v, bx, by, bz - 1, 7);
  if (lane == 7) nzp = fetch_slice(v, bx, by, bz + 1, 0);
  const uint64_t s_dn = __shfl_up_sync(0xffffffffu, s, 1), s_up = __shfl_down_sync(0xffffffffu, s, 1);
  if (lane < 8) {
    const uint64_t C0 = 0x0101010101010101ull;
    const uint64_t xm = fetch_slice(v, bx - 1, by, bz, z), xp = fetch_slice(v, bx + 1, by, bz, z);
    const uint64_t ym = fetch_slice(v, bx, by - 1, bz, z), yp = fetch_slice(v, bx, by + 1, bz, z);
    const uint64_t n_xm = ((s << 1) & ~C0) | ((xm >> 7) & C0);
    const uint64_t n_xp = ((s >> 1) & ~(C0 << 7)) | ((xp & C0) << 7);
    const uint64_t n_ym = (s << 8) | (ym >> 56);
    const uint64_t n_yp = (s >> 8) | (yp << 56);
    const uint64_t n_zm = z > 0 ? s_dn : nzm;
    const uint64_t n_zp = z < 7 ? s_up : nzp;
    s_e[warp][0][z] = s & ~n_xm; s_e[warp][1][z] = s & ~n_xp;
    s_e[warp][2][z] = s & ~n_ym; s_e[warp][3][z] = s & ~n_yp;
    s_e[warp][4][z] = s & ~n_zm; s_e[warp][5][z] = s & ~n_zp;
  }
  __syncwarp();
  // 48 (dir, layer) images over 32 lanes: task = lane and lane + 32
  uint64_t img[2]; int tdir[2], tlay[2]; int cnt = 0;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int task = lane + 32 * k;
    img[k] = 0; tdir[k] = task >> 3; tlay[k] = task & 7;
    if (task < 48) {
      const int dir = tdir[k], l = tlay[k];
      if (dir >= 4) img[k] = s_e[warp][dir][l];
      else {
        uint64_t im = 0;
#pragma unroll
        for (int zz = 0; zz < 8; zz++) {
          const uint64_t e = s_e[warp][dir][zz];
          const uint32_t row = dir < 2 ? gather_col(e, l) : ((uint32_t)(e >> (8 * l)) & 0xFFu);
          im |= (uint64_t)row << (8 * zz);
        }
        img[k] = im;
      }
      if (img[k]) cnt += greedy_image<false>(img[k], dir, l, 0, 0, 0, nullptr, 0, 0);
    }
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) continue;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(quad_count, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 0);
  int64_t pos = (int64_t)base + incl - cnt;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (lane + 32 * k < 48 && img[k]) pos += greedy_image<true>(img[k], tdir[k], tlay[k], bx * 8, by * 8, bz * 8, quads, pos, cap);
  }
}
}

void launch_mesh(const LaunchCtx& lc, const DVolume& v, int rank, int world, uint64_t* d_work, uint32_t* d_work_count,
                 MesoQuad* d_quads, int64_t cap, unsigned long long* d_quad_count) {
  cudaMemsetAsync(d_work_count, 0, sizeof(uint32_t), lc.stream);
  cudaMemsetAsync(d_quad_count, 0, sizeof(unsigned long long), lc.stream);
  const int64_t n = v.nchunks * MESO_BLOCKS;
  mesh_worklist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(v, rank, world, d_work, d_work_count);
  mesh_bricks_kernel<<<lc.sm_count * 8, 256, 0, lc.stream>>>(v, d_work, d_work_count, 0u, d_quads, cap, d_quad_count);
  (*lc.launches) += 2;
}

void launch_mesh_list(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_keys, uint32_t n_keys, MesoQuad* d_quads,
                      int64_t cap, unsigned long long* d_quad_count) {
  cudaMemsetAsync(d_quad_count, 0, sizeof(unsigned long long), lc.stream);
  if (n_keys == 0) return;
  const unsigned grid = (unsigned)min((int64_t)lc.sm_count * 8, ((int64_t)n_keys + 7) / 8);
  mesh_bricks_kernel<<<grid, 256, 0, lc.stream>>>(v, d_keys, nullptr, n_keys, d_quads, cap, d_quad_count);
  (*lc.launches)++;
}
