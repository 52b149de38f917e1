// k_resident.cu -- K6: which chunks should be resident for a view, in what order, and which of them to generate now.
//
// Replaces, on the device and per camera change (the reference bakes 256 directions at start-up, "seconds", 60 MB,
// and then uses the table of the nearest baked direction; ChunkManagerHelper.h:201-234):
//   FChunkManageHelper::GetDesiredShowChunkLocationByView / ...Simple   Runtimes/Voxel/Chunk/ChunkManagerHelper.h:89-198
//   FImportanceComputeInfo::CalculateChunkImportance                    ChunkManagerHelper.h:26-44
//   std::priority_queue<pair<float, ivec3>> pop order                   ChunkManagerHelper.h:76-87
//   the dispatch loop of FChunkManage::UpdateLoadingQueue               Runtimes/Voxel/Chunk/ChunkManager.h:229-283
//     (pop in importance order, skip locations already known, stop at the per-update generation limit)
// fp32 with glm's operation order and IEEE sqrt/div (-prec-sqrt, -prec-div, -fmad=false): bit-identical to
// oracle/orc_resident.c.  glm definitions used: dot = (x*x' + y*y') + z*z'; length = sqrt(dot(v, v));
// normalize = v * (1 / sqrt(dot(v, v))).  std::max(a, b) = (a < b) ? b : a and std::min(a, b) = (b < a) ? b : a are
// written out because the reference leans on their NaN behaviour (normalize(0) at the camera's own chunk).
//
// Data: (2F+1)^3 candidates are evaluated (117 649 for F = 24), ~15 k survive.  A survivor is one 64-bit key
// (~importance bits << 32 | loop-order index): ascending key order is importance descending with ties in the reference's
// loop order (X outer, Z inner).  Sorting is a rank sort -- every key counts the keys below it, 8 lanes per key -- which
// for 15 k keys is 28 M comparisons spread over the whole GPU and needs no scratch, passes or host round trip.
#include "meso_internal.cuh"

struct ViewParams {
  int F, B, mode;
  float view_threshold;
  float fwd[3];
};

__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax * bx + ay * by) + az * bz; }

// CalculateChunkImportance for offset (x, y, z) = ChunkLocation - CameraChunk (ChunkManagerHelper.h:29-42)
__device__ __forceinline__ float chunk_importance(int x, int y, int z, float fx, float fy, float fz) {
  if (x >= -2 && x <= 2 && y >= -2 && y <= 2 && z >= -2 && z <= 2) return 1.0e6f;
  const float ox = (float)x, oy = (float)y, oz = (float)z;
  const float d2 = dot3(ox, oy, oz, ox, oy, oz);
  const float inv = 1.0f / sqrtf(d2);
  const float dist = sqrtf(d2);
  const float angle = std_max((std_max(0.0f, dot3(ox * inv, oy * inv, oz * inv, fx, fy, fz)) - 0.5f) * 2.0f, 0.75f);
  const float distance = std_max(0.25f, 64.0f - dist);
  return angle * distance;
}

#define SEL_BUCKETS 4096   // top 12 bits of the key (sign, exponent and three mantissa bits of the inverted importance)
__global__ void __launch_bounds__(256) select_view_kernel(ViewParams p, uint64_t* __restrict__ keys, uint32_t* count, uint32_t* hist) {
  const int side = 2 * p.F + 1;
  const int64_t total = (int64_t)side * side * side;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  bool ok = false;
  float imp = 0.0f;
  if (idx < total) {
    const int Z = (int)(idx % side) - p.F, Y = (int)((idx / side) % side) - p.F, X = (int)(idx / ((int64_t)side * side)) - p.F;
    const float ox = (float)X, oy = (float)Y, oz = (float)Z;
    const float d2 = dot3(ox, oy, oz, ox, oy, oz);
    const float len = sqrtf(d2);
    if (!((double)len > (double)p.F + 1e-6)) {
      const bool core = X >= -1 && X <= 1 && Y >= -1 && Y <= 1 && Z >= -1 && Z <= 1;
      if (p.mode == 0) {
        if (core) ok = true;
        else {
          const float inv = 1.0f / sqrtf(d2);
          const float finv = 1.0f / sqrtf(dot3(p.fwd[0], p.fwd[1], p.fwd[2], p.fwd[0], p.fwd[1], p.fwd[2]));
          float alpha = std_max(dot3(p.fwd[0] * finv, p.fwd[1] * finv, p.fwd[2] * finv, ox * inv, oy * inv, oz * inv), 0.0f);
          const bool in_cone = alpha > p.view_threshold;
          alpha = in_cone ? 1.0f : alpha / p.view_threshold;
          alpha = std_min(std_max(alpha, 0.0f), 1.0f);
          const float thr = alpha * (float)p.F + (1.0f - alpha) * (float)p.B;
          ok = len < thr;
        }
        if (ok) imp = chunk_importance(X, Y, Z, p.fwd[0], p.fwd[1], p.fwd[2]);
      } else {
        if (core) { ok = true; imp = 1.0e6f; }
        else if (len < (float)p.F) { ok = true; imp = 1.0f / len; }
      }
    }
  }
  // warp-aggregated append; the order of the key list is irrelevant (the sort below is total)
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (ok) {
    const uint64_t key = ((uint64_t)(0xFFFFFFFFu - __float_as_uint(imp)) << 32) | (uint64_t)(uint32_t)idx;
    keys[base + __popc(m & ((1u << lane) - 1u))] = key;
    atomicAdd(&hist[key >> 52], 1u);
  }
}

// Sorting ~20 k 64-bit keys: a counting pass on the key's top 12 bits (histogram filled by the select kernel, one-CTA scan,
// scatter into bucket order) followed by a rank sort INSIDE each bucket, 8 lanes per key -- every key counts the keys of its
// bucket below it.  The importances spread over ~70 buckets, so this does ~2 % of the comparisons of ranking against all
// keys (round 1: 400 M comparisons, 158 us -- the second-largest kernel of a stream update); still no passes over the data
// that depend on its size class, and the output is the exact ascending key order.
__global__ void __launch_bounds__(1024) bucket_scan_kernel(uint32_t* hist /* SEL_BUCKETS + 1 */, uint32_t* cursor) {
  __shared__ uint32_t s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t v[4], x = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = hist[threadIdx.x * 4 + k]; x += v[k]; cursor[threadIdx.x * 4 + k] = 0u; }
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
    s_warp[lane] = wi - wv;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - x;
#pragma unroll
  for (int k = 0; k < 4; k++) { hist[threadIdx.x * 4 + k] = run; run += v[k]; }
  if (threadIdx.x == 1023) hist[SEL_BUCKETS] = run;
}
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ count,
                                                             const uint32_t* __restrict__ off, uint32_t* cursor, uint64_t* __restrict__ keys2) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= *count) return;
  const uint64_t k = keys[i];
  const uint32_t b = (uint32_t)(k >> 52);
  keys2[off[b] + atomicAdd(&cursor[b], 1u)] = k;
}
// 32 keys per CTA, 8 lanes per key.  The grid is sized for the worst case; CTAs past *count exit.
__global__ void __launch_bounds__(256) rank_sort_kernel(const uint64_t* __restrict__ keys2, const uint32_t* __restrict__ count,
                                                        const uint32_t* __restrict__ off, MesoChunkCandidate* __restrict__ out, int64_t cap, int F) {
  const uint32_t n = *count;
  if (blockIdx.x * 32u >= n) return;
  const uint32_t k = blockIdx.x * 32u + (threadIdx.x >> 3);
  const uint32_t sub = threadIdx.x & 7;
  const uint64_t my = (k < n) ? keys2[k] : ~0ull;
  const uint32_t b = (uint32_t)(my >> 52);
  const uint32_t lo = (k < n) ? off[b] : 0u, hi = (k < n) ? off[b + 1] : 0u;
  uint32_t rank = 0;
#pragma unroll 4
  for (uint32_t j = lo + sub; j < hi; j += 8) rank += (__ldg(keys2 + j) < my) ? 1u : 0u;
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  rank += __shfl_xor_sync(0xffffffffu, rank, 4);
  rank += lo;
  if (sub == 0 && k < n && (int64_t)rank < cap) {
    const int side = 2 * F + 1;
    const uint32_t idx = (uint32_t)my;
    MesoChunkCandidate c;
    c.Importance = __uint_as_float(0xFFFFFFFFu - (uint32_t)(my >> 32));
    c.Offset[2] = (int)(idx % side) - F; c.Offset[1] = (int)((idx / side) % side) - F; c.Offset[0] = (int)(idx / (side * side)) - F;
    out[rank] = c;
  }
}

// CalculateChunkImportance for a list of absolute chunk locations (the eviction score of resident chunks)
__global__ void chunk_importance_kernel(const int32_t* __restrict__ loc, int64_t n, int cx, int cy, int cz, float fx, float fy, float fz,
                                        float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = chunk_importance(loc[3 * i] - cx, loc[3 * i + 1] - cy, loc[3 * i + 2] - cz, fx, fy, fz);
}

// The dispatch loop of UpdateLoadingQueue (ChunkManager.h:229-283) over the sorted candidates, one CTA: a candidate is
// "not found in ChunksLookupTable" iff its chunk lies inside the grid window and its loaded bit is clear; the first
// max_new such candidates, in priority order, become the generation list and are marked loaded (= EChunkState::Computing).
// stats: [0] chunks listed now, [1] desired chunks still missing after this update, [2] candidates, [3] candidates inside
// the window.
__global__ void __launch_bounds__(1024) stream_worklist_kernel(DVolume v, const MesoChunkCandidate* __restrict__ cand,
                                                               const uint32_t* __restrict__ count, int64_t cap, int cx, int cy, int cz,
                                                               uint32_t* loaded, uint32_t max_new, uint32_t* list, uint32_t* stats) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry, s_in;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_carry = 0; s_in = 0; }
  __syncthreads();
  uint32_t n = *count;
  if ((int64_t)n > cap) n = (uint32_t)cap;
  uint32_t in_grid_local = 0;
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    uint32_t flag = 0, slot = 0;
    if (i < n) {
      const int x = cand[i].Offset[0] + cx - v.origin[0], y = cand[i].Offset[1] + cy - v.origin[1], z = cand[i].Offset[2] + cz - v.origin[2];
      if (x >= 0 && x < v.dims[0] && y >= 0 && y < v.dims[1] && z >= 0 && z < v.dims[2]) {
        slot = (uint32_t)(x + v.dims[0] * (y + v.dims[1] * z));
        in_grid_local++;
        flag = ((loaded[slot >> 5] >> (slot & 31)) & 1u) ^ 1u;
      }
    }
    uint32_t incl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t wv = s_warp[lane];
      uint32_t wi = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
      s_warp[lane] = wi - wv;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t pos = carry + s_warp[warp] + incl - flag;
    if (flag && pos < max_new) {
      list[pos] = slot;
      atomicOr(&loaded[slot >> 5], 1u << (slot & 31));
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = pos + flag;
    __syncthreads();
  }
  atomicAdd(&s_in, in_grid_local);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t missing = s_carry;
    stats[0] = missing < max_new ? missing : max_new;
    stats[1] = missing - stats[0];
    stats[2] = *count;
    stats[3] = s_in;
  }
}

static ViewParams make_params(const float fwd[3], const MesoViewConfig& vc) {
  ViewParams p;
  p.F = (int)vc.ViewForwardLoadChunkSize; p.B = (int)vc.ViewBackwardLoadChunkSize; p.mode = (int)vc.Mode;
  // ChunkManagerHelper.h:97: max(cos(radians(ViewChunkAngle) * 0.5f), 0.01f); libm cosf as in the oracle
  volatile float rad = vc.ViewChunkAngle * 0.01745329251994329576923690768489f;
  volatile float half = rad * 0.5f;
  const float cv = cosf(half);
  p.view_threshold = (cv < 0.01f) ? 0.01f : cv;
  p.fwd[0] = fwd[0]; p.fwd[1] = fwd[1]; p.fwd[2] = fwd[2];
  return p;
}

int64_t resident_max_candidates(const MesoViewConfig& vc) {
  const int64_t side = 2 * (int64_t)vc.ViewForwardLoadChunkSize + 1;
  return side * side * side;
}

// d_keys: scratch of resident_sort_scratch_bytes(): keys | keys in bucket order | bucket offsets | bucket cursors;
// d_count: one u32; d_out: cap candidates, sorted
size_t resident_sort_scratch_bytes(const MesoViewConfig& vc) { return (size_t)resident_max_candidates(vc) * 16 + (size_t)(2 * SEL_BUCKETS + 8) * 4; }
void launch_select_view(const LaunchCtx& lc, const float fwd[3], const MesoViewConfig& vc, uint64_t* d_keys, uint32_t* d_count,
                        MesoChunkCandidate* d_out, int64_t cap) {
  const ViewParams p = make_params(fwd, vc);
  const int64_t total = resident_max_candidates(vc);
  uint64_t* keys2 = d_keys + total;
  uint32_t* hist = reinterpret_cast<uint32_t*>(keys2 + total);
  uint32_t* cursor = hist + SEL_BUCKETS + 4;
  cudaMemsetAsync(d_count, 0, sizeof(uint32_t), lc.stream);
  cudaMemsetAsync(hist, 0, (SEL_BUCKETS + 1) * sizeof(uint32_t), lc.stream);
  select_view_kernel<<<(unsigned)((total + 255) / 256), 256, 0, lc.stream>>>(p, d_keys, d_count, hist);
  bucket_scan_kernel<<<1, 1024, 0, lc.stream>>>(hist, cursor);
  bucket_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, lc.stream>>>(d_keys, d_count, hist, cursor, keys2);
  rank_sort_kernel<<<(unsigned)((total + 31) / 32), 256, 0, lc.stream>>>(keys2, d_count, hist, d_out, cap, p.F);
  (*lc.launches) += 4;
}

// FImportanceComputeInfo::CalculateBlockImportance (ChunkManagerHelper.h:50-70), restated literally including its dead near
// branch: the int offset is compared with `-2 * ChunkResolution`, ChunkResolution being uint32_t, so both sides convert to
// unsigned and no value satisfies `u >= 4294967264 && u <= 32`.  Far branch: the chunk formula with Far = 64 * ChunkResolution.
__global__ void block_importance_kernel(const int32_t* __restrict__ chunk_loc, const uint8_t* __restrict__ block_loc, int64_t n, int cx, int cy, int cz,
                                        float fx, float fy, float fz, uint32_t res, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = (int)res;
  const int x = (chunk_loc[3 * i] - cx) * r + block_loc[3 * i], y = (chunk_loc[3 * i + 1] - cy) * r + block_loc[3 * i + 1],
            z = (chunk_loc[3 * i + 2] - cz) * r + block_loc[3 * i + 2];
  const uint32_t lo = (uint32_t)(-2) * res, hi = 2u * res;
  const bool near_ = ((uint32_t)x >= lo && (uint32_t)x <= hi) && ((uint32_t)y >= lo && (uint32_t)y <= hi) && ((uint32_t)z >= lo && (uint32_t)z <= hi);
  if (near_) { out[i] = 1.0e6f; return; }
  const float ox = (float)x, oy = (float)y, oz = (float)z;
  const float d2 = dot3(ox, oy, oz, ox, oy, oz);
  const float inv = 1.0f / sqrtf(d2);
  const float dist = sqrtf(d2);
  const float angle = std_max((std_max(0.0f, dot3(ox * inv, oy * inv, oz * inv, fx, fy, fz)) - 0.5f) * 2.0f, 0.75f);
  const float distance = std_max(0.25f, 64.0f * (float)res - dist);
  out[i] = angle * distance;
}
void launch_block_importance(const LaunchCtx& lc, const int32_t* d_chunk_loc, const uint8_t* d_block_loc, int64_t n, const int32_t cam[3], const float fwd[3],
                             uint32_t chunk_resolution, float* d_out) {
  if (n <= 0) return;
  block_importance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(d_chunk_loc, d_block_loc, n, cam[0], cam[1], cam[2], fwd[0], fwd[1], fwd[2],
                                                                               chunk_resolution, d_out);
  (*lc.launches)++;
}

// ---- debug visualisation data (ChunkPool.h:550-561, ShaderWireFrame.h:18-22) ------------------------------------------------
// One FGPUSimpleInstanceData per resident chunk -- what the reference's chunk-wireframe pass draws an octahedron for:
// Position = (ChunkSize,)*3, ChunkLocation, Scale = ChunkSize * 0.1, Rotation = identity quaternion, Marker = 1 for a chunk
// with blocks (FChunk) and 0 for an empty one (FEmptyChunk).  loaded == null: every chunk of the window is resident.
__global__ void __launch_bounds__(256) debug_instances_kernel(DVolume v, const uint32_t* __restrict__ loaded, float chunk_size,
                                                              MesoGPUSimpleInstanceData* __restrict__ out, int64_t cap, uint32_t* count) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  bool ok = c < v.nchunks && (!loaded || ((loaded[c >> 5] >> (c & 31)) & 1u));
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (!ok) return;
  const int64_t idx = (int64_t)base + __popc(m & ((1u << lane) - 1u));
  if (idx >= cap) return;
  MesoGPUSimpleInstanceData r;
  r.Position[0] = r.Position[1] = r.Position[2] = chunk_size;
  r.ChunkLocation[0] = v.origin[0] + (int)(c % v.dims[0]);
  r.ChunkLocation[1] = v.origin[1] + (int)((c / v.dims[0]) % v.dims[1]);
  r.ChunkLocation[2] = v.origin[2] + (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  r.Scale = chunk_size * 0.1f;
  r.Rotation[0] = 1.0f; r.Rotation[1] = r.Rotation[2] = r.Rotation[3] = 0.0f;
  r.Marker = ((v.chunk_any[c >> 5] >> (c & 31)) & 1u) ? 1.0f : 0.0f;
  out[idx] = r;
}
void launch_debug_instances(const LaunchCtx& lc, const DVolume& v, const uint32_t* d_loaded, float chunk_size, MesoGPUSimpleInstanceData* d_out,
                            int64_t cap, uint32_t* d_count) {
  cudaMemsetAsync(d_count, 0, sizeof(uint32_t), lc.stream);
  debug_instances_kernel<<<(unsigned)((v.nchunks + 255) / 256), 256, 0, lc.stream>>>(v, d_loaded, chunk_size, d_out, cap, d_count);
  (*lc.launches)++;
}

// ---- moving window (FChunkPool's eviction, ChunkPool.h:447-622, in the form a dense window needs) --------------------------
// The window follows the camera: origin' = origin + delta.  Chunks that leave it are evicted -- their payload slots go on
// the free stack and are handed out again before any new slot (alloc_payload_slot) -- chunks that stay keep their data at
// their new slot (slot = position relative to the origin), chunks that enter start empty and not generated; the next
// meso_stream_update generates them in priority order like any other missing chunk.
__global__ void __launch_bounds__(256) evict_chunks_kernel(DVolume v, int dx, int dy, int dz) {
  const int64_t c = blockIdx.x;     // old slot
  const int x = (int)(c % v.dims[0]), y = (int)((c / v.dims[0]) % v.dims[1]), z = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  const int nx = x - dx, ny = y - dy, nz = z - dz;   // position in the new window
  if ((unsigned)nx < (unsigned)v.dims[0] && (unsigned)ny < (unsigned)v.dims[1] && (unsigned)nz < (unsigned)v.dims[2]) return;   // stays
  for (int w = threadIdx.x; w < 64; w += blockDim.x) {
    uint64_t part = v.occ[c * 64 + w] & ~v.full[c * 64 + w];
    while (part) {
      const int bit = __ffsll((long long)part) - 1;
      part &= part - 1;
      const uint32_t slot = v.bptr[c * MESO_BLOCKS + w * 64 + bit];
      if (slot < v.max_bricks) v.pool_free[atomicAdd(v.pool_free_count, 1)] = slot;
    }
  }
}
// dst[new slot] = src[new slot + delta] for the chunks that stay, `fill` words for those that enter; W 32-bit words per chunk
__global__ void __launch_bounds__(256) shift_chunks_kernel(DVolume v, int dx, int dy, int dz, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                           int words_per_chunk, uint32_t fill) {
  const int64_t n = blockIdx.x;     // new slot
  const int x = (int)(n % v.dims[0]), y = (int)((n / v.dims[0]) % v.dims[1]), z = (int)(n / ((int64_t)v.dims[0] * v.dims[1]));
  const int ox = x + dx, oy = y + dy, oz = z + dz;   // where it was
  const bool stays = (unsigned)ox < (unsigned)v.dims[0] && (unsigned)oy < (unsigned)v.dims[1] && (unsigned)oz < (unsigned)v.dims[2];
  const int64_t o = stays ? (int64_t)ox + (int64_t)v.dims[0] * ((int64_t)oy + (int64_t)v.dims[1] * oz) : 0;
  const uint4 f4 = make_uint4(fill, fill, fill, fill);
  const uint4* s4 = reinterpret_cast<const uint4*>(src + o * words_per_chunk);
  uint4* d4 = reinterpret_cast<uint4*>(dst + n * words_per_chunk);
  for (int i = threadIdx.x; i < words_per_chunk / 4; i += blockDim.x) d4[i] = stays ? s4[i] : f4;
}
__global__ void shift_loaded_kernel(DVolume v, int dx, int dy, int dz, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= v.nchunks) return;
  const int x = (int)(n % v.dims[0]), y = (int)((n / v.dims[0]) % v.dims[1]), z = (int)(n / ((int64_t)v.dims[0] * v.dims[1]));
  const int ox = x + dx, oy = y + dy, oz = z + dz;
  if (!((unsigned)ox < (unsigned)v.dims[0] && (unsigned)oy < (unsigned)v.dims[1] && (unsigned)oz < (unsigned)v.dims[2])) return;
  const int64_t o = (int64_t)ox + (int64_t)v.dims[0] * ((int64_t)oy + (int64_t)v.dims[1] * oz);
  if ((src[o >> 5] >> (o & 31)) & 1u) atomicOr(&dst[n >> 5], 1u << (n & 31));
}
// d_scratch: nchunks * 16 KiB (the largest per-chunk array, bptr).  v.origin is updated (host copy of the descriptor).
void launch_window_shift(const LaunchCtx& lc, DVolume& v, const int delta[3], uint32_t* d_loaded, void* d_scratch) {
  const unsigned nc = (unsigned)v.nchunks;
  evict_chunks_kernel<<<nc, 64, 0, lc.stream>>>(v, delta[0], delta[1], delta[2]);
  struct Arr { void* p; int words; uint32_t fill; };
  const Arr arrs[4] = {{v.occ, 128, 0u}, {v.full, 128, 0u}, {v.mips, 384, 0u}, {v.bptr, MESO_BLOCKS, 0xFFFFFFFFu}};
  for (const Arr& a : arrs) {
    shift_chunks_kernel<<<nc, 256, 0, lc.stream>>>(v, delta[0], delta[1], delta[2], (const uint32_t*)a.p, (uint32_t*)d_scratch, a.words, a.fill);
    cudaMemcpyAsync(a.p, d_scratch, (size_t)nc * a.words * 4, cudaMemcpyDeviceToDevice, lc.stream);
  }
  cudaMemsetAsync(d_scratch, 0, (size_t)v.chunk_words * 4, lc.stream);
  shift_loaded_kernel<<<(nc + 255) / 256, 256, 0, lc.stream>>>(v, delta[0], delta[1], delta[2], d_loaded, (uint32_t*)d_scratch);
  cudaMemcpyAsync(d_loaded, d_scratch, (size_t)v.chunk_words * 4, cudaMemcpyDeviceToDevice, lc.stream);
  (*lc.launches) += 6;
  for (int i = 0; i < 3; i++) v.origin[i] += delta[i];
  launch_volume_finalize(lc, v);   // of / cells / chunk and region bits / distance field of the shifted window
}

void launch_chunk_importance(const LaunchCtx& lc, const int32_t* d_loc, int64_t n, const int32_t cam[3], const float fwd[3], float* d_out) {
  if (n <= 0) return;
  chunk_importance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(d_loc, n, cam[0], cam[1], cam[2], fwd[0], fwd[1], fwd[2], d_out);
  (*lc.launches)++;
}

void launch_stream_worklist(const LaunchCtx& lc, const DVolume& v, const MesoChunkCandidate* d_cand, const uint32_t* d_count, int64_t cap,
                            const int32_t cam[3], uint32_t* d_loaded, uint32_t max_new, uint32_t* d_list, uint32_t* d_stats) {
  stream_worklist_kernel<<<1, 1024, 0, lc.stream>>>(v, d_cand, d_count, cap, cam[0], cam[1], cam[2], d_loaded, max_new, d_list, d_stats);
  (*lc.launches)++;
}
