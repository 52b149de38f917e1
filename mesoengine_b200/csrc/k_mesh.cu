// k_mesh.cu -- K3: face culling + greedy merge of exposed voxel faces into compacted quad lists.
//
// The north-star form of the reference's "mesher" (hidden-block cull + instance compaction,
// Runtimes/Voxel/Chunk/ChunkPool.h:381-445; its draw picks faces per view in the VS, Samples/SimpleVoxel.cpp:129-136).
// Neither quads nor a face-neighbour test exist in the reference: the definition is oracle/orc_mesh.c
// ("parity unpinned by reference; bit-exact vs repo oracle" after the canonical sort).
//
// Pass A (worklist): one thread per 64-brick occupancy word; occupied bricks that are not buried (full with six full
// neighbours, decided with word-wide shifts) are compacted with a warp prefix sum + one atomic per warp.
// Pass B (mesh): one warp per brick.  Lanes 0..6 resolve the seven bricks involved, then lanes 0..7 own one z-slice
// each (64 voxels as a u64, every slice an independent 8 B load) and build the six exposed-face slice masks with
// shifts against the neighbour slices; the 48 (direction, layer) 8x8 images are then merged greedily by 48
// lane-tasks, counted, prefix-summed across the warp, and written behind one atomicAdd.
// HBM-bound integer work: 64 B per populated brick + 16 B per quad (+ neighbour slices, mostly L2 hits).
#include "meso_internal.cuh"

// full word w of chunk (cx,cy,cz), zeros outside the grid
__device__ __forceinline__ uint64_t full_word(const DVolume& v, int cx, int cy, int cz, int w) {
  if ((unsigned)cx >= (unsigned)v.dims[0] || (unsigned)cy >= (unsigned)v.dims[1] || (unsigned)cz >= (unsigned)v.dims[2]) return 0ull;
  return __ldg(&v.full[chunk_index(v, cx, cy, cz) * 64 + w]);
}

// Pass A, one thread per 64-brick occupancy word (word = z*4 + y/4, four 16-bit x-rows): bricks that are full and
// have six full neighbours are buried (no exposed face); the six neighbour masks come from shifted full-words of this
// and the adjacent words / chunks.  Survivors are appended to the work list (warp prefix + one atomic per warp).
__global__ void __launch_bounds__(256) mesh_worklist_kernel(DVolume v, int rank, int world, uint64_t* work, uint32_t* work_count) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t todo = 0;
  int64_t c = 0; int w = 0;
  if (t < v.nchunks * MESO_WORDS) {
    c = t >> 6; w = (int)(t & 63);
    if ((c % world) == rank) {
      const uint64_t O = __ldg(&v.occ[t]);
      if (O) {
        const uint64_t F = __ldg(&v.full[t]);
        uint64_t buried = 0;
        if (F) {
          const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
          const int z = w >> 2, yq = w & 3;
          const uint64_t X0 = 0x0001000100010001ull, X15 = 0x8000800080008000ull;
          const uint64_t fxm = ((F << 1) & ~X0) | ((full_word(v, cx - 1, cy, cz, w) >> 15) & X0);
          const uint64_t fxp = ((F >> 1) & ~X15) | ((full_word(v, cx + 1, cy, cz, w) << 15) & X15);
          const uint64_t ym_src = yq > 0 ? __ldg(&v.full[t - 1]) : full_word(v, cx, cy - 1, cz, z * 4 + 3);
          const uint64_t yp_src = yq < 3 ? __ldg(&v.full[t + 1]) : full_word(v, cx, cy + 1, cz, z * 4 + 0);
          const uint64_t fym = (F << 16) | (ym_src >> 48);
          const uint64_t fyp = (F >> 16) | (yp_src << 48);
          const uint64_t fzm = z > 0 ? __ldg(&v.full[t - 4]) : full_word(v, cx, cy, cz - 1, 15 * 4 + yq);
          const uint64_t fzp = z < 15 ? __ldg(&v.full[t + 4]) : full_word(v, cx, cy, cz + 1, 0 * 4 + yq);
          buried = F & fxm & fxp & fym & fyp & fzm & fzp;
        }
        todo = O & ~buried;
      }
    }
  }
  const int lane = threadIdx.x & 31;
  const uint32_t n = (uint32_t)__popcll(todo);
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(work_count, total);
  base = __shfl_sync(0xffffffffu, base, 0) + incl - n;
  while (todo) {
    const int bit = __ffsll((long long)todo) - 1;
    todo &= todo - 1;
    work[base++] = (uint64_t)c * MESO_BLOCKS + (uint64_t)(w * 64 + bit);
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

#define MQ_CAP 224   // quads a warp can stage in shared memory (a smooth surface brick yields ~20)
#define MQ_FLUSH 128 // staged quads are written out once this many have accumulated (and at the end of the warp's work)

// Same greedy merge, staging the quads in shared memory; slots come from a shared-memory atomic counter.
__device__ __forceinline__ void greedy_stage(uint64_t img, int dir, int layer, int ox, int oy, int oz, uint4* stage, int* count) {
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
      int x, y, z;
      if (ax == 0) { x = layer; y = u0; z = vv; } else if (ax == 1) { x = u0; y = layer; z = vv; } else { x = u0; y = vv; z = layer; }
      const int idx = atomicAdd(count, 1);
      if (idx < MQ_CAP)
        stage[idx] = make_uint4((uint32_t)(ox + x) | ((uint32_t)(oy + y) << 16), (uint32_t)(oz + z) | ((uint32_t)dir << 16) | ((uint32_t)w << 24),
                                (uint32_t)h, 0u);
    }
  }
}

// Persistent, grid-stride over the work list (count read from device memory: no host round trip between the passes).
// A warp takes TWO bricks per iteration, one per half-warp, through the slice phase (lanes 0..6 of each half resolve the
// seven bricks involved, lanes 0..7 own one z-slice each).  The 2 x 48 (direction, layer) images are then built in three
// full rounds, the non-empty ones (typically 6..12 of 48) are queued in shared memory, and the greedy merge runs over the
// queue densely: one round of busy lanes instead of three rounds of mostly idle ones.
__global__ void __launch_bounds__(256) mesh_bricks_kernel(DVolume v, const uint64_t* __restrict__ work, const uint32_t* __restrict__ work_count_ptr,
                                                          uint32_t work_count_imm, MesoQuad* quads, int64_t cap, unsigned long long* quad_count,
                                                          int shard_rank, int shard_world) {
  __shared__ uint64_t s_e[8][2][6][8];
  __shared__ uint64_t s_img[8][96];
  __shared__ uint8_t s_meta[8][96];
  __shared__ int s_org[8][2][3];
  __shared__ uint4 s_q[8][MQ_CAP];
  __shared__ int s_n[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, hl = lane & 15;
  const uint32_t n_work = work_count_ptr ? *work_count_ptr : work_count_imm;
  const int64_t n_pairs = ((int64_t)n_work + 1) >> 1;
  if (lane == 0) s_n[warp] = 0;
  __syncwarp();
  // write the first `count` staged quads behind one reservation and empty the staging area
  auto flush = [&](int count) {
    if (count > 0) {
      unsigned long long sbase = 0;
      if (lane == 0) sbase = atomicAdd_system(quad_count, (unsigned long long)count);   // system scope: the counter may live in a peer GPU
      sbase = __shfl_sync(0xffffffffu, sbase, 0);
      for (int i = lane; i < count; i += 32)
        if ((int64_t)sbase + i < cap) reinterpret_cast<uint4*>(quads)[sbase + i] = s_q[warp][i];
    }
    __syncwarp();
    if (lane == 0) s_n[warp] = 0;
    __syncwarp();
  };
  for (int64_t pair =(int64_t)blockIdx.x * 8 + warp; pair < n_pairs; pair += (int64_t)gridDim.x * 8) {
  __syncwarp();
  const int64_t item = pair * 2 + half;
  bool valid = item < (int64_t)n_work;
  const uint64_t key = valid ? work[item] : 0ull;
  // key lists (dirty re-mesh) are sharded over the ranks by a hash of the key: the list order is scheduling-dependent and
  // differs between the replicas, the key set does not
  if (shard_world > 1 && (int)(((key * 0x9E3779B97F4A7C15ull) >> 40) % (unsigned)shard_world) != shard_rank) valid = false;
  const int64_t c = (int64_t)(key >> 12); const int b = (int)(key & 4095);
  const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  const int bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);
  if (hl == 0) { s_org[warp][half][0] = bx * 8; s_org[warp][half][1] = by * 8; s_org[warp][half][2] = bz * 8; }

  const int z = hl & 7;
  // Step 1: lanes 0..6 of each half resolve the seven bricks involved (self, -x, +x, -y, +y, -z, +z) in parallel: one
  // 16 B {occ,full} load each, then the payload slot of the partial ones.  state: 0 empty / outside, 1 full, 2 partial.
  int st = 0; uint32_t slot = 0;
  if (valid && hl < 7) {
    const int nx = bx + (hl == 1 ? -1 : (hl == 2 ? 1 : 0));
    const int ny = by + (hl == 3 ? -1 : (hl == 4 ? 1 : 0));
    const int nz = bz + (hl == 5 ? -1 : (hl == 6 ? 1 : 0));
    if ((unsigned)nx < (unsigned)(v.dims[0] * 16) && (unsigned)ny < (unsigned)(v.dims[1] * 16) && (unsigned)nz < (unsigned)(v.dims[2] * 16)) {
      const int64_t nc = chunk_index(v, nx >> 4, ny >> 4, nz >> 4);
      const int nb = block_bit(nx & 15, ny & 15, nz & 15);
      const ulonglong2 p = __ldg(&v.of[nc * 64 + (nb >> 6)]);
      if ((p.x >> (nb & 63)) & 1ull) {
        if ((p.y >> (nb & 63)) & 1ull) st = 1;
        else { st = 2; slot = __ldg(&v.bptr[nc * MESO_BLOCKS + nb]); }
      }
    }
  }
  // Step 2: every slice needed is one independent 8 B load (lanes 0..7 of the half: slice z of self and of the four
  // lateral neighbours; lane 0 / 7: the facing slice of the -z / +z neighbour).
  auto slice_of = [&](int which, int zz) -> uint64_t {
    const int s_st = __shfl_sync(0xffffffffu, st, (lane & 16) | which);
    const uint32_t s_slot = __shfl_sync(0xffffffffu, slot, (lane & 16) | which);
    if (s_st == 2 && hl < 8) return __ldg(&v.pool[(size_t)s_slot * 8 + zz]);
    return s_st == 1 ? ~0ull : 0ull;
  };
  uint64_t s = slice_of(0, z);   // every lane takes part in the shuffles
  if (hl >= 8) s = 0ull;
  const uint64_t xm = slice_of(1, z), xp = slice_of(2, z), ym = slice_of(3, z), yp = slice_of(4, z);
  const uint64_t nzm = slice_of(5, 7), nzp = slice_of(6, 0);
  const uint64_t s_dn = __shfl_up_sync(0xffffffffu, s, 1), s_up = __shfl_down_sync(0xffffffffu, s, 1);
  if (hl < 8) {
    const uint64_t C0 = 0x0101010101010101ull;
    const uint64_t n_xm = ((s << 1) & ~C0) | ((xm >> 7) & C0);
    const uint64_t n_xp = ((s >> 1) & ~(C0 << 7)) | ((xp & C0) << 7);
    const uint64_t n_ym = (s << 8) | (ym >> 56);
    const uint64_t n_yp = (s >> 8) | (yp << 56);
    const uint64_t n_zm = z > 0 ? s_dn : nzm;
    const uint64_t n_zp = z < 7 ? s_up : nzp;
    s_e[warp][half][0][z] = s & ~n_xm; s_e[warp][half][1][z] = s & ~n_xp;
    s_e[warp][half][2][z] = s & ~n_ym; s_e[warp][half][3][z] = s & ~n_yp;
    s_e[warp][half][4][z] = s & ~n_zm; s_e[warp][half][5][z] = s & ~n_zp;
  }
  __syncwarp();
  // 2 x 48 (dir, layer) images over 32 lanes in three full rounds; the non-empty ones are queued (order irrelevant)
  int ni = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int task = lane + 32 * k;
    const int h = task >= 48 ? 1 : 0;
    const int t48 = task - 48 * h;
    const int dir = t48 >> 3, l = t48 & 7;
    uint64_t im = 0;
    if (dir >= 4) im = s_e[warp][h][dir][l];
    else {
#pragma unroll
      for (int zz = 0; zz < 8; zz++) {
        const uint64_t e = s_e[warp][h][dir][zz];
        const uint32_t row = dir < 2 ? gather_col(e, l) : ((uint32_t)(e >> (8 * l)) & 0xFFu);
        im |= (uint64_t)row << (8 * zz);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, im != 0ull);
    if (im) {
      const int idx = ni + __popc(bal & ((1u << lane) - 1u));
      s_img[warp][idx] = im;
      s_meta[warp][idx] = (uint8_t)(dir | (l << 3) | (h << 6));
    }
    ni += __popc(bal);
  }
  __syncwarp();
  if (ni == 0) continue;
  // Fast path: greedy passes over the queue stage the quads of both bricks in shared memory (slot = shared atomic, order
  // is irrelevant).  Staged quads accumulate over several iterations and leave in batches of >= MQ_FLUSH: one
  // (system-scope) atomicAdd and one coalesced copy of 16 B records per batch -- the reservation is a round trip over
  // NVLink when the list lives in a peer GPU.
  int before = s_n[warp];
  if (before >= MQ_FLUSH) { flush(before); before = 0; }
  for (int i = lane; i < ni; i += 32) {
    const int m = s_meta[warp][i], h = m >> 6;
    greedy_stage(s_img[warp][i], m & 7, (m >> 3) & 7, s_org[warp][h][0], s_org[warp][h][1], s_org[warp][h][2], s_q[warp], &s_n[warp]);
  }
  __syncwarp();
  if (s_n[warp] <= MQ_CAP) continue;
  // Rare: more quads than the staging area holds -> write out what earlier iterations staged, then count, prefix and emit
  // this pair's quads straight to global memory.
  flush(before);
  int cnt = 0;
  for (int i = lane; i < ni; i += 32) cnt += greedy_image<false>(s_img[warp][i], 0, 0, 0, 0, 0, nullptr, 0, 0);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd_system(quad_count, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 0);
  int64_t pos = (int64_t)base + incl - cnt;
  for (int i = lane; i < ni; i += 32) {
    const int m = s_meta[warp][i], h = m >> 6;
    pos += greedy_image<true>(s_img[warp][i], m & 7, (m >> 3) & 7, s_org[warp][h][0], s_org[warp][h][1], s_org[warp][h][2], quads, pos, cap);
  }
}
  __syncwarp();
  flush(s_n[warp]);
}

void launch_mesh(const LaunchCtx& lc, const DVolume& v, int rank, int world, uint64_t* d_work, uint32_t* d_work_count,
                 MesoQuad* d_quads, int64_t cap, unsigned long long* d_quad_count, bool reset_count) {
  cudaMemsetAsync(d_work_count, 0, sizeof(uint32_t), lc.stream);
  if (reset_count) cudaMemsetAsync(d_quad_count, 0, sizeof(unsigned long long), lc.stream);
  const int64_t n = v.nchunks * MESO_WORDS;
  mesh_worklist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(v, rank, world, d_work, d_work_count);
  mesh_bricks_kernel<<<lc.sm_count * 8, 256, 0, lc.stream>>>(v, d_work, d_work_count, 0u, d_quads, cap, d_quad_count, 0, 1);
  (*lc.launches) += 2;
}

void launch_mesh_list(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_keys, uint32_t n_keys, MesoQuad* d_quads,
                      int64_t cap, unsigned long long* d_quad_count, int rank, int world) {
  cudaMemsetAsync(d_quad_count, 0, sizeof(unsigned long long), lc.stream);
  if (n_keys == 0) return;
  const unsigned grid = (unsigned)min((int64_t)lc.sm_count * 8, ((int64_t)n_keys + 7) / 8);
  mesh_bricks_kernel<<<grid, 256, 0, lc.stream>>>(v, d_keys, nullptr, n_keys, d_quads, cap, d_quad_count, rank, world);
  (*lc.launches)++;
}
