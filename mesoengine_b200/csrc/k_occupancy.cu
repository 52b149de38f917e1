// k_occupancy.cu -- K2: erode mips, hidden-block cull, instance compaction, chunk table.
//
// Replaces, for every chunk of the grid at once:
//   FChunk::CalculateOccupancyErodeMipmaps      Runtimes/Voxel/Chunk/Chunk.h:73-94
//   FOccupancyHelper::ErodeSingleVoxel<true>    Runtimes/Voxel/Occupancy/BinaryOccupancyVolume.h:76-98
//   FChunk::bShouldVoxelOccupancyCull(loc, 1)   Chunk.h:96-100 (call site ChunkPool.h:387)
//   emission loop of FChunkPool::PushToBlockPool Runtimes/Voxel/Chunk/ChunkPool.h:385-390,438
//   FGPUChunk record                             ChunkPool.h:567
// Integer/bit work, HBM-bound: reads 512 B per chunk, writes 3 x 512 B of mips + 12 B per emitted instance.
//
// The reference erodes one voxel at a time with 26 bit reads (and a heap allocation per call).  Here a 16^3 mask is
// 64 u64 words (word = z*4 + y/4, four 16-bit x-rows per word); a thread owns one word and builds the 26-neighbour
// AND from nine shifted words (x shifts inside the word, y shifts across adjacent words, z = word +-4).
#include "meso_internal.cuh"

#define INNER_ROW 0x7FFEull  // x in [1,14]

// rows y+1 aligned onto rows y for word w of mask m (smem, 64 words); y+1 == 16 -> zeros
__device__ __forceinline__ uint64_t rows_up(const uint64_t* m, int w) {
  uint64_t s = m[w] >> 16;
  if ((w & 3) != 3) s |= m[w + 1] << 48;
  return s;
}
__device__ __forceinline__ uint64_t rows_down(const uint64_t* m, int w) {
  uint64_t s = m[w] << 16;
  if ((w & 3) != 0) s |= m[w - 1] >> 48;
  return s;
}
__device__ __forceinline__ uint64_t h3(uint64_t x) { return x & (x << 1) & (x >> 1); }

// AND of the 26 neighbours (self excluded, BinaryOccupancyVolume.h:45-62) for the 64 positions of word w; zero on the
// one-thick chunk shell (bIsOutOfBoundThickness, VoxelMathHelper.h:106-112).
__device__ __forceinline__ uint64_t erode26_word(const uint64_t* m, int w) {
  const int z = w >> 2, yq = w & 3;
  if (z == 0 || z == 15) return 0ull;
  uint64_t inner = INNER_ROW | (INNER_ROW << 16) | (INNER_ROW << 32) | (INNER_ROW << 48);
  if (yq == 0) inner &= ~0xFFFFull;           // y = 0
  if (yq == 3) inner &= ~(0xFFFFull << 48);   // y = 15
  const uint64_t s = m[w];
  uint64_t r = h3(rows_up(m, w)) & h3(rows_down(m, w)) & (s << 1) & (s >> 1);
  r &= h3(m[w - 4]) & h3(rows_up(m, w - 4)) & h3(rows_down(m, w - 4));
  r &= h3(m[w + 4]) & h3(rows_up(m, w + 4)) & h3(rows_down(m, w + 4));
  return r & inner;
}

// One CTA (64 threads) per chunk: mips 1..3, cull mask, instance count, chunk table record.
__global__ void __launch_bounds__(64) occupancy_mips_kernel(DVolume v, uint32_t stamp, MesoGPUChunk* table, uint32_t* counts) {
  const int64_t c = blockIdx.x;
  const int w = threadIdx.x;
  __shared__ uint64_t a[64], b[64];
  __shared__ uint32_t s_cnt[2];
  const uint64_t occ = v.occ[c * 64 + w];
  a[w] = occ;
  __syncthreads();
  // Mip_d is evaluated only at solid block locations (Chunk.h:85-90): AND with the block set.
  const uint64_t m1 = erode26_word(a, w) & occ;
  b[w] = m1;
  v.mips[(c * 3 + 0) * 64 + w] = m1;
  __syncthreads();
  const uint64_t m2 = erode26_word(b, w) & occ;
  a[w] = m2;
  v.mips[(c * 3 + 1) * 64 + w] = m2;
  __syncthreads();
  const uint64_t m3 = erode26_word(a, w) & occ;
  v.mips[(c * 3 + 2) * 64 + w] = m3;
  // emitted <=> block present and !Mip1 (ChunkPool.h:387 with the hard-coded depth 1)
  int n = __popcll(occ & ~m1);
  int blocks = __popcll(occ);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); blocks += __shfl_xor_sync(0xffffffffu, blocks, o); }
  if ((w & 31) == 0) { s_cnt[w >> 5] = (uint32_t)n; }
  __shared__ int s_blocks[2];
  if ((w & 31) == 0) s_blocks[w >> 5] = blocks;
  __syncthreads();
  if (w == 0) {
    counts[c] = s_cnt[0] + s_cnt[1];
    MesoGPUChunk rec;
    if (s_blocks[0] + s_blocks[1] > 0) {
      const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
      rec.ChunkLocation[0] = v.origin[0] + cx; rec.ChunkLocation[1] = v.origin[1] + cy; rec.ChunkLocation[2] = v.origin[2] + cz;
      rec.ChunkFrameStamp = stamp;
    } else {  // FEmptyChunk: no GPU chunk record (Chunk.h:29 invalid location)
      rec.ChunkLocation[0] = rec.ChunkLocation[1] = rec.ChunkLocation[2] = 2147483647;
      rec.ChunkFrameStamp = 0;
    }
    table[c] = rec;
  }
}

// exclusive scan of per-chunk counts in three small launches: 1024 counts per CTA (local exclusive offsets + the CTA's
// total), one CTA scanning the totals (up to 1024 of them: 2^20 chunks), and the add.  (A single 1024-thread CTA walking all
// counts took 38 us at 32 768 chunks -- a fifth of K2.)
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t x, uint32_t* s_warp, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t wv = s_warp[lane];
    uint32_t wi = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
    s_warp[lane] = wi - wv;                    // exclusive per warp
    if (lane == 31) s_warp[32] = wi;           // CTA total
  }
  __syncthreads();
  total = s_warp[32];
  return s_warp[warp] + incl - x;
}
__global__ void __launch_bounds__(1024) scan_local_kernel(const uint32_t* __restrict__ counts, uint32_t* offsets, int64_t n, uint32_t* block_totals) {
  __shared__ uint32_t s_warp[33];
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  uint32_t total;
  const uint32_t ex = block_exclusive_scan_1024(i < n ? counts[i] : 0u, s_warp, total);
  if (i < n) offsets[i] = ex;
  if (threadIdx.x == 0) block_totals[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) scan_totals_kernel(uint32_t* block_totals, int nblocks, uint64_t* total_out) {
  __shared__ uint32_t s_warp[33];
  uint32_t total;
  const uint32_t ex = block_exclusive_scan_1024((int)threadIdx.x < nblocks ? block_totals[threadIdx.x] : 0u, s_warp, total);
  if ((int)threadIdx.x < nblocks) block_totals[threadIdx.x] = ex;
  if (threadIdx.x == 0) *total_out = total;
}
__global__ void __launch_bounds__(1024) scan_add_kernel(uint32_t* offsets, int64_t n, const uint32_t* __restrict__ block_totals) {
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (i < n && blockIdx.x > 0) offsets[i] += block_totals[blockIdx.x];
}

// One CTA (256 threads) per chunk.  Thread t <-> column (X = t>>4, Y = t&15), i.e. thread order is the reference's
// generator order (X outer, Y, Z inner; GeneratorHelper.h:125-129): a block-wide exclusive scan of the per-column
// popcounts gives every emitted block its rank in exactly the order PushToBlockPool walks Chunk.Blocks.  Records are
// staged in shared memory at their rank and copied out as one contiguous, coalesced run of 32-bit words.
__global__ void __launch_bounds__(256) emit_instances_kernel(DVolume v, uint32_t stamp, const uint32_t* __restrict__ counts,
                                                             const uint32_t* __restrict__ offsets, MesoGPUBlock* inst, int64_t cap) {
  const int64_t c = blockIdx.x;
  const uint32_t total = counts[c];
  if (total == 0) return;
  __shared__ uint16_t s_xyz[MESO_BLOCKS];   // X | Y << 4 | Z << 8 of the emitted blocks, at their rank (8 KB: eight CTAs per SM;
                                            // staging the 12-byte records themselves took 48 KB and capped occupancy at four)
  __shared__ uint64_t cull[64];
  __shared__ uint32_t s_warp[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < 64) cull[t] = v.occ[c * 64 + t] & ~v.mips[(c * 3 + 0) * 64 + t];
  __syncthreads();
  const int X = t >> 4, Y = t & 15;
  uint32_t col = 0;
#pragma unroll
  for (int Z = 0; Z < 16; Z++) col |= (uint32_t)((cull[Z * 4 + (Y >> 2)] >> (X + 16 * (Y & 3))) & 1ull) << Z;
  const uint32_t n = __popc(col);
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int k = 0; k < warp; k++) wbase += s_warp[k];
  uint32_t r = wbase + incl - n;
  while (col) {
    const int Z = __ffs(col) - 1;
    col &= col - 1;
    s_xyz[r++] = (uint16_t)(X | (Y << 4) | (Z << 8));
  }
  __syncthreads();
  const int64_t first = (int64_t)offsets[c];
  const int64_t room = cap - first;
  const uint32_t words = (uint32_t)(room <= 0 ? 0 : (room < (int64_t)total ? room : (int64_t)total)) * 3u;
  uint32_t* out = reinterpret_cast<uint32_t*>(inst) + first * 3;
  // one contiguous, coalesced run of 32-bit words; word i belongs to record i / 3: {ChunkIndex, u8vec4(x, y, z, 255), stamp}
  for (uint32_t i = t; i < words; i += 256) {
    const uint32_t rec = i / 3u, part = i - 3u * rec;
    const uint32_t q = s_xyz[rec];
    out[i] = part == 0 ? (uint32_t)c : (part == 1 ? ((q & 15u) | (((q >> 4) & 15u) << 8) | ((q >> 8) << 16) | (255u << 24)) : stamp);
  }
}

// The reference's own upload records (FChunkPool::UploadChunk / UploadBlock, ChunkPool.h:662-679) scattered into the block
// masks: one thread per FGPUBlock.  A block counts iff the vertex shader would draw it (SimpleVoxel.cpp:160-166:
// ChunkIndex valid, chunk record valid, stamps equal) and its chunk lies inside the window; it becomes an all-solid
// brick (the reference has no voxel-in-brick level).  12 B read per block, two 8-byte atomics.
__global__ void __launch_bounds__(256) scatter_blocks_kernel(DVolume v, const MesoGPUChunk* __restrict__ chunks, int64_t n_chunks,
                                                             const MesoGPUBlock* __restrict__ blocks, int64_t n_blocks,
                                                             unsigned long long* accepted) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  bool ok = false;
  if (i < n_blocks) {
    const MesoGPUBlock b = blocks[i];
    if (b.ChunkIndex != 2147483647u && (int64_t)b.ChunkIndex < n_chunks) {
      const MesoGPUChunk ch = chunks[b.ChunkIndex];
      const bool chunk_valid = !(ch.ChunkLocation[0] == 2147483647 && ch.ChunkLocation[1] == 2147483647 && ch.ChunkLocation[2] == 2147483647);
      const int lx = ch.ChunkLocation[0] - v.origin[0], ly = ch.ChunkLocation[1] - v.origin[1], lz = ch.ChunkLocation[2] - v.origin[2];
      if (chunk_valid && ch.ChunkFrameStamp == b.BlockFrameStamp && (unsigned)lx < (unsigned)v.dims[0] && (unsigned)ly < (unsigned)v.dims[1] &&
          (unsigned)lz < (unsigned)v.dims[2] && b.BlockLocation[0] < 16 && b.BlockLocation[1] < 16 && b.BlockLocation[2] < 16) {
        const int64_t c = chunk_index(v, lx, ly, lz);
        const int bit = block_bit(b.BlockLocation[0], b.BlockLocation[1], b.BlockLocation[2]);
        const unsigned long long m = 1ull << (bit & 63);
        atomicOr((unsigned long long*)&v.occ[c * 64 + (bit >> 6)], m);
        atomicOr((unsigned long long*)&v.full[c * 64 + (bit >> 6)], m);
        ok = true;
      }
    }
  }
  const unsigned n = __popc(__ballot_sync(0xffffffffu, ok));
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(accepted, (unsigned long long)n);
}
void launch_scatter_blocks(const LaunchCtx& lc, const DVolume& v, const MesoGPUChunk* d_chunks, int64_t n_chunks, const MesoGPUBlock* d_blocks,
                           int64_t n_blocks, unsigned long long* d_accepted) {
  if (n_blocks <= 0) return;
  scatter_blocks_kernel<<<(unsigned)((n_blocks + 255) / 256), 256, 0, lc.stream>>>(v, d_chunks, n_chunks, d_blocks, n_blocks, d_accepted);
  (*lc.launches)++;
}

// pass 1: mips + per-chunk instance counts + exclusive scan (d_total = number of instances)
void launch_occupancy_count(const LaunchCtx& lc, const DVolume& v, uint32_t stamp, MesoGPUChunk* d_table, uint32_t* d_counts,
                            uint32_t* d_offsets, uint64_t* d_total, uint32_t* d_block_totals) {
  occupancy_mips_kernel<<<(unsigned)v.nchunks, 64, 0, lc.stream>>>(v, stamp, d_table, d_counts);
  const int nblocks = (int)((v.nchunks + 1023) / 1024);      // <= 1024: meso_scene_create caps the grid at 2^20 chunks
  scan_local_kernel<<<nblocks, 1024, 0, lc.stream>>>(d_counts, d_offsets, v.nchunks, d_block_totals);
  scan_totals_kernel<<<1, 1024, 0, lc.stream>>>(d_block_totals, nblocks, d_total);
  scan_add_kernel<<<nblocks, 1024, 0, lc.stream>>>(d_offsets, v.nchunks, d_block_totals);
  (*lc.launches) += 4;
}
// pass 2: compacted FGPUBlock list in generator order
void launch_occupancy_emit(const LaunchCtx& lc, const DVolume& v, uint32_t stamp, const uint32_t* d_counts, const uint32_t* d_offsets,
                           MesoGPUBlock* d_inst, int64_t cap_inst) {
  emit_instances_kernel<<<(unsigned)v.nchunks, 256, 0, lc.stream>>>(v, stamp, d_counts, d_offsets, d_inst, cap_inst);
  (*lc.launches) += 1;
}
