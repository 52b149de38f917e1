// k_carve.cu -- K5: sphere carve edit + dirty-brick list (BASELINE.json configs[4], "interactive edit loop").
// Not in the reference; defined by oracle/orc_mesh.c:orc_carve_sphere -- integer arithmetic only:
// voxel (x,y,z) is removed iff (2x+1-2cx)^2 + (2y+1-2cy)^2 + (2z+1-2cz)^2 < (2r)^2.
// HBM-bound: 64 B read + 64 B written per brick in the sphere's AABB, 8 B per dirty-list entry.
// Mapping: 8 lanes per brick (one z-slice each), 4 bricks per warp; full bricks that become partial take a payload
// slot from the same bump allocator the voxeliser uses.
#include "meso_internal.cuh"
#include <algorithm>

struct CarveBox { int lo[3]; int hi[3]; int c[3]; int radius; };

__global__ void __launch_bounds__(256) carve_kernel(DVolume v, CarveBox box, uint64_t* dirty, uint32_t cap_dirty, uint32_t* dirty_count, int* overflow) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g = t >> 3; const int z = (int)(t & 7);
  const int lane = threadIdx.x & 31;
  const int ex = box.hi[0] - box.lo[0] + 1, ey = box.hi[1] - box.lo[1] + 1, ez = box.hi[2] - box.lo[2] + 1;
  const bool valid = g < (int64_t)ex * ey * ez;
  int bx = 0, by = 0, bz = 0; int64_t c = 0; int b = 0;
  uint64_t s = 0, m = 0; bool was_full = false, occupied = false;
  if (valid) {
    bx = box.lo[0] + (int)(g % ex); by = box.lo[1] + (int)((g / ex) % ey); bz = box.lo[2] + (int)(g / ((int64_t)ex * ey));
    c = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
    b = block_bit(bx & 15, by & 15, bz & 15);
    occupied = (v.occ[c * 64 + (b >> 6)] >> (b & 63)) & 1ull;
    if (occupied) {
      was_full = (v.full[c * 64 + (b >> 6)] >> (b & 63)) & 1ull;
      s = was_full ? ~0ull : v.pool[(size_t)v.bptr[c * MESO_BLOCKS + b] * 8 + z];
      m = s;
      const long long r2 = 4ll * box.radius * box.radius;
      const long long dz = 2ll * (bz * 8 + z) + 1 - 2ll * box.c[2];
#pragma unroll 1
      for (int y = 0; y < 8; y++) {
        const long long dy = 2ll * (by * 8 + y) + 1 - 2ll * box.c[1];
        const long long rem = r2 - dz * dz - dy * dy;
        if (rem <= 0) continue;
#pragma unroll
        for (int x = 0; x < 8; x++) {
          const long long dx = 2ll * (bx * 8 + x) + 1 - 2ll * box.c[0];
          if (dx * dx < rem) m &= ~(1ull << (x + 8 * y));
        }
      }
    }
  }
  // reductions inside the 8-lane group
  unsigned chg = (m != s) ? 1u : 0u, any = (m != 0ull) ? 1u : 0u;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) { chg |= __shfl_xor_sync(0xffffffffu, chg, o); any |= __shfl_xor_sync(0xffffffffu, any, o); }
  uint32_t slot = 0xFFFFFFFFu;
  if (valid && occupied && chg && z == 0) {
    // A full brick that becomes partial needs a payload slot: take it BEFORE touching the masks.  If the pool is
    // exhausted the brick is left exactly as it was (still full, not dirty), the allocator is rolled back and the call
    // reports the overflow -- the volume stays consistent (no occ && !full brick without a payload).
    bool ok = true;
    if (was_full && any) {
      slot = alloc_payload_slot(v);
      if (slot >= v.max_bricks) { release_bump(v); *overflow = 1; slot = 0xFFFFFFFFu; ok = false; }
    }
    if (ok) {
      const uint32_t di = atomicAdd(dirty_count, 1u);
      if (di < cap_dirty) dirty[di] = (uint64_t)c * MESO_BLOCKS + (uint64_t)b;
      const unsigned long long bit = 1ull << (b & 63);
      if (was_full) atomicAnd((unsigned long long*)&v.full[c * 64 + (b >> 6)], ~bit);
      if (!any) {
        atomicAnd((unsigned long long*)&v.occ[c * 64 + (b >> 6)], ~bit);
        v.bptr[c * MESO_BLOCKS + b] = 0xFFFFFFFFu;  // payload slot (if any) is retired, not recycled
      } else if (was_full) {
        v.bptr[c * MESO_BLOCKS + b] = slot;
      } else {
        slot = v.bptr[c * MESO_BLOCKS + b];
      }
    }
  }
  slot = __shfl_sync(0xffffffffu, slot, lane & ~7);
  if (valid && occupied && chg && any && slot != 0xFFFFFFFFu) v.pool[(size_t)slot * 8 + z] = m;
}

// dirty bricks + their six neighbours, de-duplicated through a bit grid over all blocks of the scene
__global__ void __launch_bounds__(256) expand_dirty_kernel(DVolume v, const uint64_t* __restrict__ dirty, uint32_t n_dirty, uint32_t* mark,
                                                           uint64_t* keys, uint32_t cap, uint32_t* count) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t / 7; const int k = (int)(t % 7);
  if (i >= n_dirty) return;
  const uint64_t key = dirty[i];
  const int64_t c = (int64_t)(key >> 12); const int b = (int)(key & 4095);
  const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  int bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);
  if (k == 1) bx--; else if (k == 2) bx++; else if (k == 3) by--; else if (k == 4) by++; else if (k == 5) bz--; else if (k == 6) bz++;
  if ((unsigned)bx >= (unsigned)(v.dims[0] * 16) || (unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return;
  const uint64_t nk = (uint64_t)chunk_index(v, bx >> 4, by >> 4, bz >> 4) * MESO_BLOCKS + (uint64_t)block_bit(bx & 15, by & 15, bz & 15);
  const uint32_t old = atomicOr(&mark[nk >> 5], 1u << (nk & 31));
  if (old & (1u << (nk & 31))) return;
  const uint32_t idx = atomicAdd(count, 1u);
  if (idx < cap) keys[idx] = nk;
}
__global__ void clear_marks_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ count, uint32_t cap, uint32_t* mark) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = min(*count, cap);
  if (i < n) atomicAnd(&mark[keys[i] >> 5], ~(1u << (keys[i] & 31)));
}

__global__ void flush_kernel(uint32_t* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

void launch_carve(const LaunchCtx& lc, const DVolume& v, const int32_t center[3], int32_t radius, uint64_t* d_dirty,
                  uint32_t cap_dirty, uint32_t* d_dirty_count, int* d_overflow) {
  cudaMemsetAsync(d_dirty_count, 0, sizeof(uint32_t), lc.stream);
  CarveBox box;
  int64_t n = 1;
  for (int i = 0; i < 3; i++) {
    long long a = ((long long)center[i] - radius) >> 3, b = ((long long)center[i] + radius) >> 3;
    const long long nb = (long long)v.dims[i] * 16;
    if (a < 0) a = 0;
    if (b > nb - 1) b = nb - 1;
    box.lo[i] = (int)a; box.hi[i] = (int)b; box.c[i] = center[i];
    n *= (b >= a) ? (b - a + 1) : 0;
  }
  box.radius = radius;
  if (n <= 0) return;
  carve_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, lc.stream>>>(v, box, d_dirty, cap_dirty, d_dirty_count, d_overflow);
  (*lc.launches)++;
  launch_volume_finalize(lc, v, /*rebuild_df=*/false);   // voxels were only removed: the distance field stays conservative
}

// d_mark: zero-initialised bit grid over all blocks of the scene (left all-zero again on return)
void launch_expand_dirty(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_dirty, uint32_t n_dirty, uint64_t* d_keys,
                         uint32_t cap, uint32_t* d_count, uint32_t* d_mark) {
  cudaMemsetAsync(d_count, 0, sizeof(uint32_t), lc.stream);
  if (n_dirty == 0) return;
  const int64_t nt = (int64_t)n_dirty * 7;
  expand_dirty_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, lc.stream>>>(v, d_dirty, n_dirty, d_mark, d_keys, cap, d_count);
  const uint32_t bound = (uint32_t)(nt < (int64_t)cap ? nt : (int64_t)cap);
  clear_marks_kernel<<<(bound + 255) / 256, 256, 0, lc.stream>>>(d_keys, d_count, cap, d_mark);
  (*lc.launches) += 2;
}

// Small device -> host reads (counters, a picked record, a few thousand quads) go through a kernel that stores into mapped
// pinned host memory instead of cudaMemcpyAsync: a copy command queues behind whatever large DMA is in flight on the same
// copy engine (a 66 MB slab on its way to the host delayed a 4-byte count by 2 ms in the N-GPU edit loop); stores do not.
__global__ void __launch_bounds__(256) peek_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t bytes) {
  const size_t words = bytes >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  if (blockIdx.x == 0 && threadIdx.x < (bytes & 3)) dst[(words << 2) + threadIdx.x] = src[(words << 2) + threadIdx.x];
}
void launch_peek(const LaunchCtx& lc, void* d_dst_mapped, const void* d_src, size_t bytes) {
  if (bytes == 0) return;
  const unsigned grid = (unsigned)std::min<size_t>((bytes / 4 + 255) / 256 + 1, (size_t)lc.sm_count * 4);
  peek_kernel<<<grid, 256, 0, lc.stream>>>((unsigned char*)d_dst_mapped, (const unsigned char*)d_src, bytes);
  (*lc.launches)++;
}

// Stream-ordered rendezvous between GPUs without a collective: a signal adds one to a word (usually a peer's), a wait spins
// on a word in this GPU's memory until it reaches `target`.  The wait is bounded (about 2 s): a peer that never arrives sets
// the timeout flag instead of hanging the device.
__global__ void signal_kernel(SignalTargets t) {
  __threadfence_system();
  for (int i = 0; i < t.n; i++) atomicAdd_system(t.word[i], 1u);
}
__global__ void wait_kernel(const unsigned* word, unsigned target, int* timeout_flag) {
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
    if ((int)(v - target) >= 0) break;
    if (clock64() - t0 > 4000000000ll) { *timeout_flag = 1; break; }
    __nanosleep(200);
  }
  __threadfence_system();
}
void launch_signal(const LaunchCtx& lc, const SignalTargets& t) {
  signal_kernel<<<1, 1, 0, lc.stream>>>(t);
  (*lc.launches)++;
}
void launch_wait(const LaunchCtx& lc, const unsigned* d_word, unsigned target, int* d_timeout_flag) {
  wait_kernel<<<1, 1, 0, lc.stream>>>(d_word, target, d_timeout_flag);
  (*lc.launches)++;
}

void launch_flush(const LaunchCtx& lc, uint32_t* d_scratch, size_t n_words) {
  flush_kernel<<<lc.sm_count * 4, 256, 0, lc.stream>>>(d_scratch, n_words);
  (*lc.launches)++;
}
