// k_voxelize.cu -- K1: SDF -> chunk/block/brick occupancy, plus volume upload/download helpers.
//
// Replaces the reference's generator workers (Runtimes/Voxel/Chunk/ChunkManager.h:160-210) running
// FGeneratorHelper::GenerateSphere / TestGenerator (Runtimes/Helper/GeneratorHelper.h:90-150) over every chunk.
// fp64 throughout, glm evaluation order, no fma (-fmad=false): bit-identical to the CPU oracle.
//
// Mapping: one CTA per 64-bit occupancy word (= 64 bricks: 16 x, 4 y, one z of a chunk); 8 warps x 8 bricks;
// a warp evaluates the 512 voxels of a brick 16 per lane, assembles the eight z-slices with two shuffles and
// (for partial bricks) bumps the payload allocator once.  The CTA writes its occ/full words with plain stores.
// Roofline: sphere = HBM-write (N^3/8 B of payload at most), terrain = fp64 ALU (~1.5 kflop + 72 sin per sample).
#include "meso_internal.cuh"

// ---- portable fp64 sin: same operation sequence as oracle/orc_sdf.c:orc_sin_portable ---------------------------
__device__ __forceinline__ double sin_portable(double x) {
  const double PIO2_1 = 0x1.921fb54000000p+0, PIO2_2 = 0x1.10b4611800000p-30, PIO2_3 = 0x1.313198a000000p-61,
               PIO2_4 = 0x1.701b839a25205p-92, TWO_OVER_PI = 0x1.45f306dc9c883p-1;
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double k = floor(x * TWO_OVER_PI + 0.5);
  double r = x - k * PIO2_1;
  r = r - k * PIO2_2;
  r = r - k * PIO2_3;
  r = r - k * PIO2_4;
  double q = k - 4.0 * floor(k * 0.25);
  double z = r * r;
  double s, c;
  {
    double v = z * r;
    double p = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    s = r + v * (S1 + z * p);
  }
  {
    double p = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    c = 1.0 - (0.5 * z - z * p);
  }
  if (q == 0.0) return s;
  if (q == 1.0) return c;
  if (q == 2.0) return -s;
  return -c;
}

// Runtimes/Helper/VoxelMathHelper.h:25-33
__device__ __forceinline__ double hash3(double x, double y, double z) {
  double d = (x * 127.1 + y * 311.7) + z * 74.7;
  double v = sin_portable(d) * 43758.5453123;
  return v - floor(v);
}

// Runtimes/Helper/GeneratorHelper.h:19-53; out = (grad.xyz, value)
__device__ void noised(const double x[3], double out4[4]) {
  double p[3], w[3], u[3], du[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    p[i] = floor(x[i]);
    w[i] = x[i] - floor(x[i]);
    u[i] = ((w[i] * w[i]) * w[i]) * ((w[i] * ((w[i] * 6.0) - 15.0)) + 10.0);
    du[i] = ((30.0 * w[i]) * w[i]) * ((w[i] * (w[i] - 2.0)) + 1.0);
  }
  double a = hash3(p[0] + 0, p[1] + 0, p[2] + 0);
  double b = hash3(p[0] + 1, p[1] + 0, p[2] + 0);
  double c = hash3(p[0] + 0, p[1] + 1, p[2] + 0);
  double d = hash3(p[0] + 1, p[1] + 1, p[2] + 0);
  double e = hash3(p[0] + 0, p[1] + 0, p[2] + 1);
  double f = hash3(p[0] + 1, p[1] + 0, p[2] + 1);
  double g = hash3(p[0] + 0, p[1] + 1, p[2] + 1);
  double h = hash3(p[0] + 1, p[1] + 1, p[2] + 1);
  double k0 = a;
  double k1 = b - a;
  double k2 = c - a;
  double k3 = e - a;
  double k4 = a - b - c + d;
  double k5 = a - c - e + g;
  double k6 = a - b - e + f;
  double k7 = -a + b + c - d + e - f - g + h;
  out4[3] = -1.0 + 2.0 * (k0 + k1 * u[0] + k2 * u[1] + k3 * u[2] + k4 * u[0] * u[1] + k5 * u[1] * u[2] +
                          k6 * u[2] * u[0] + k7 * u[0] * u[1] * u[2]);
  out4[0] = (2.0 * du[0]) * (k1 + k4 * u[1] + k6 * u[2] + k7 * u[1] * u[2]);
  out4[1] = (2.0 * du[1]) * (k2 + k5 * u[2] + k4 * u[0] + k7 * u[2] * u[0]);
  out4[2] = (2.0 * du[2]) * (k3 + k6 * u[0] + k5 * u[1] + k7 * u[0] * u[1]);
}

// Runtimes/Helper/GeneratorHelper.h:56-87
__device__ double displacement(const double p_in[3]) {
  double p[3] = {p_in[0], p_in[1], p_in[2]};
  double mgn = 0.5, d = 0.0, s = 1.0;
  double rnd[4], q[3];
  for (int i = 0; i < 5; i++) {
    q[0] = p[0] + 10.0; q[1] = p[1] + 10.0; q[2] = p[2] + 10.0;
    noised(q, rnd);
    d += rnd[3] * mgn;
#pragma unroll
    for (int k = 0; k < 3; k++) { p[k] *= 2.0; p[k] += (rnd[k] * 0.2) * s; }
    if (i == 2) s *= -1.0;
    mgn *= 0.5;
  }
  p[0] = p_in[0] * 32.0; p[1] = p_in[1] * 32.0; p[2] = p_in[2] * 32.0;  // pow(2.0, 5)
  for (int i = 0; i < 4; i++) {
    noised(p, rnd);
    d += rnd[3] * mgn;
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] *= 2.0;
    mgn *= 0.5;
  }
  return d;
}

struct SdfParams { double p[4]; };

template <int KIND>
__device__ __forceinline__ bool sdf_solid(const SdfParams& sp, double x, double y, double z) {
  if (KIND == MESO_SDF_SPHERE) {  // GeneratorHelper.h:134
    double dx = x - sp.p[0], dy = y - sp.p[1], dz = z - sp.p[2];
    return sqrt((dx * dx + dy * dy) + dz * dz) - sp.p[3] < 0.0;
  } else {                        // GeneratorHelper.h:104
    double q[3] = {x * .1, y * .1, z * .1};
    return (y * .5 + displacement(q) * 10.3) * .4 < 0.0;
  }
}

// |displacement| <= sum of amplitudes (0.998046875) x (1 + rounding) < 0.9981, so the sign of the terrain SDF is
// fixed once |y * 0.5| > 10.3 * 0.9981 = 10.2805; 10.29 leaves a 1e-2 margin.  Exact cull, never approximate.
#define TERRAIN_Y_HALF_BOUND 10.29

// Exact brick classification for the sphere.  sqrt((dx*dx + dy*dy) + dz*dz) - r is monotone non-decreasing in each of
// |dx|, |dy|, |dz| separately (every rounding step is monotone), so over the 8^3 sample lattice of a brick the maximum
// is at the per-axis farthest sample and the minimum at the per-axis nearest: all-solid <=> farthest sample solid,
// all-empty <=> nearest sample not solid.  No tolerance involved (same argument as oracle/orc_volume.c).
// The same argument holds for any box of the sample lattice p_i = lo + i/8, i in [0, n) (the coordinates are exact in
// fp64): per axis the farthest sample is one of the two ends and the nearest is the sample at or just above the centre
// coordinate.  floor((c - lo) * 8) may be off by one when c sits on a sample within rounding, but then that sample is in
// both candidate pairs; and two samples with the same computed |p - c| give the same SDF value, so ties do not matter.
__device__ __forceinline__ void axis_extremes(double lo, int n, double c, double& pn, double& pf) {
  const double p0 = lo, p1 = lo + (double)(n - 1) * 0.125;
  pf = fabs(p1 - c) > fabs(p0 - c) ? p1 : p0;
  const double t = floor((c - lo) * 8.0);
  const int i0 = t < 0.0 ? 0 : (t > (double)(n - 1) ? n - 1 : (int)t);
  const int i1 = min(i0 + 1, n - 1);
  const double q0 = lo + (double)i0 * 0.125, q1 = lo + (double)i1 * 0.125;
  pn = fabs(q1 - c) < fabs(q0 - c) ? q1 : q0;
}
// 1 all-solid, 0 all-empty, -1 mixed, for the box of nx x ny x nz samples with minimum corner (b0, b1, b2)
__device__ __forceinline__ int sphere_box_class(const SdfParams& sp, double b0, double b1, double b2, int nx, int ny, int nz) {
  double nr0, nr1, nr2, fr0, fr1, fr2;
  axis_extremes(b0, nx, sp.p[0], nr0, fr0);
  axis_extremes(b1, ny, sp.p[1], nr1, fr1);
  axis_extremes(b2, nz, sp.p[2], nr2, fr2);
  if (sdf_solid<MESO_SDF_SPHERE>(sp, fr0, fr1, fr2)) return 1;
  if (!sdf_solid<MESO_SDF_SPHERE>(sp, nr0, nr1, nr2)) return 0;
  return -1;
}
__device__ __forceinline__ int sphere_brick_class(const SdfParams& sp, double b0, double b1, double b2) {
  return sphere_box_class(sp, b0, b1, b2, 8, 8, 8);
}

// Pre-pass, one THREAD per 64-brick occupancy word: words the exact box test decides in one piece (sphere: 98 % of
// the words at 4096^3; terrain: everything outside the +-20.58 band) get their occ/full words here; the others are
// appended to the work list of voxelize_voxel_kernel (warp-aggregated).  (Doing this test inside the voxel kernel with
// one CTA per word was 10x slower: 256 threads repeating the same fp64 evaluation for 2.1 M words.)
// chunk_list != null (streaming, K6): thread t -> chunk chunk_list[t >> 6], only below *chunk_n.
template <int KIND>
__global__ void __launch_bounds__(256) classify_words_kernel(DVolume v, SdfParams sp, const uint32_t* __restrict__ chunk_list,
                                                             const uint32_t* __restrict__ chunk_n, int64_t n_chunks_imm,
                                                             uint32_t* __restrict__ words, uint32_t* n_words) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int64_t c = t >> 6;
  const int W = (int)(t & 63);
  bool live = c < (chunk_list ? (int64_t)*chunk_n : n_chunks_imm);
  int cls = 0;
  int64_t word_global = 0;
  if (live) {
    if (chunk_list) c = chunk_list[c];
    word_global = c * 64 + W;
    const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
    const double cs0 = (double)(v.origin[0] + cx) * 1.0 * 16.0, cs1 = (double)(v.origin[1] + cy) * 1.0 * 16.0,
                 cs2 = (double)(v.origin[2] + cz) * 1.0 * 16.0;
    const double b1 = cs1 + (double)(4 * (W & 3));          // the word: bricks X 0..15, Y 4(W&3)..+3, Z W>>2
    if (KIND == MESO_SDF_SPHERE) cls = sphere_box_class(sp, cs0, b1, cs2 + (double)(W >> 2), 128, 32, 8);
    else cls = (b1 * .5 > TERRAIN_Y_HALF_BOUND) ? 0 : (((b1 + 3.875) * .5 < -TERRAIN_Y_HALF_BOUND) ? 1 : -1);
    if (cls >= 0) { v.occ[word_global] = cls ? ~0ull : 0ull; v.full[word_global] = cls ? ~0ull : 0ull; }
  }
  const bool mixed = live && cls < 0;
  const unsigned m = __ballot_sync(0xffffffffu, mixed);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(n_words, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (mixed) words[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)word_global;
}

// Persistent over the list of mixed words (count on the device): CTA per word.
template <int KIND>
__global__ void __launch_bounds__(256) voxelize_voxel_kernel(DVolume v, SdfParams sp, int* overflow, const uint32_t* __restrict__ words,
                                                             const uint32_t* __restrict__ n_words) {
  __shared__ unsigned long long s_occ, s_full;
  __shared__ int s_list[64];
  __shared__ int s_mixed;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vz = lane >> 2, q = lane & 3;
  const uint32_t n = *n_words;
  for (uint32_t wi = blockIdx.x; wi < n; wi += gridDim.x) {
  const int64_t word_global = (int64_t)words[wi];
  const int64_t c = word_global >> 6;
  const int W = (int)(word_global & 63);
  const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  const double cs0 = (double)(v.origin[0] + cx) * 1.0 * 16.0, cs1 = (double)(v.origin[1] + cy) * 1.0 * 16.0,
               cs2 = (double)(v.origin[2] + cz) * 1.0 * 16.0;
  __syncthreads();   // the previous word's shared state has been consumed
  if (threadIdx.x == 0) { s_occ = 0ull; s_full = 0ull; s_mixed = 0; }
  __syncthreads();
  // ---- pass A: one thread per brick classifies it exactly (all-solid / all-empty / mixed) ----
  if (threadIdx.x < 64) {
    const int i = threadIdx.x;
    const int bi = W * 64 + i;
    const int X = bi & 15, Y = (bi >> 4) & 15, Z = bi >> 8;
    const double b0 = cs0 + (double)X * 1.0, b1 = cs1 + (double)Y * 1.0, b2 = cs2 + (double)Z * 1.0;
    int cls = -1;
    if (KIND == MESO_SDF_TERRAIN) {
      if (b1 * .5 > TERRAIN_Y_HALF_BOUND) cls = 0;
      else if ((b1 + 0.875) * .5 < -TERRAIN_Y_HALF_BOUND) cls = 1;
    } else {
      cls = sphere_brick_class(sp, b0, b1, b2);
    }
    const unsigned solid = __ballot_sync(0xffffffffu, cls == 1);
    const unsigned mixed = __ballot_sync(0xffffffffu, cls < 0);
    if (lane == 0) {
      atomicOr(&s_occ, (unsigned long long)solid << (32 * warp));
      atomicOr(&s_full, (unsigned long long)solid << (32 * warp));
    }
    if (cls < 0) s_list[atomicAdd(&s_mixed, 1)] = i;   // order is irrelevant: every mixed brick is evaluated in full
    (void)mixed;
  }
  __syncthreads();
  // ---- pass B: a warp evaluates all 512 samples of a mixed brick, 16 per lane ----
  const int n_mixed = s_mixed;
  for (int k = warp; k < n_mixed; k += 8) {
    const int i = s_list[k];
    const int bi = W * 64 + i;
    const int X = bi & 15, Y = (bi >> 4) & 15, Z = bi >> 8;
    const double b0 = cs0 + (double)X * 1.0, b1 = cs1 + (double)Y * 1.0, b2 = cs2 + (double)Z * 1.0;
    uint32_t bits = 0;
    const double pz = b2 + (double)vz * 0.125;
#pragma unroll 1
    for (int r = 0; r < 2; r++) {
      const int vy = 2 * q + r;
      const double py = b1 + (double)vy * 0.125;
#pragma unroll 1
      for (int vx = 0; vx < 8; vx++) {
        const double px = b0 + (double)vx * 0.125;
        if (sdf_solid<KIND>(sp, px, py, pz)) bits |= 1u << (vx + 8 * r);
      }
    }
    uint64_t s = (uint64_t)bits << (16 * q);
    s |= __shfl_xor_sync(0xffffffffu, s, 1);
    s |= __shfl_xor_sync(0xffffffffu, s, 2);
    const bool any = __any_sync(0xffffffffu, s != 0ull);
    const bool all = __all_sync(0xffffffffu, s == ~0ull);
    // a mixed brick needs a payload slot; if the pool is exhausted the brick is dropped (left empty), the allocator rolled
    // back and the overflow reported: the volume never holds an occ && !full brick without a payload
    bool keep = any;
    if (any && !all) {
      uint32_t slot = 0;
      if (lane == 0) slot = alloc_payload_slot(v);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (slot < v.max_bricks) {
        if (q == 0) v.pool[(size_t)slot * 8 + vz] = s;
        if (lane == 0) v.bptr[c * MESO_BLOCKS + bi] = slot;
      } else {
        keep = false;
        if (lane == 0) { release_bump(v); *overflow = 1; }
      }
    }
    if (lane == 0) {
      if (keep) atomicOr(&s_occ, 1ull << i);
      if (all) atomicOr(&s_full, 1ull << i);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    v.occ[word_global] = s_occ;
    v.full[word_global] = s_full;
  }
  }
}

// Block-granular (reference semantics, GeneratorHelper.h:120-150): one sample at the block min corner; brick all-ones.
// mip > 0 (the generator plug-in's MipmapLevel, ChunkManager.h:61): one sample per (2^mip)^3 blocks, at the group's minimum corner.
template <int KIND>
__global__ void __launch_bounds__(64) voxelize_block_kernel(DVolume v, SdfParams sp, const uint32_t* __restrict__ list,
                                                            const uint32_t* __restrict__ list_n, int mip) {
  int64_t c = blockIdx.x >> 6;
  if (list) {
    if (c >= (int64_t)*list_n) return;
    c = list[c];
  }
  const int W = (int)(blockIdx.x & 63);
  const int64_t word_global = c * 64 + W;
  const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
  const int bi = W * 64 + threadIdx.x;
  const int gm = ~((1 << mip) - 1);
  const int X = (bi & 15) & gm, Y = ((bi >> 4) & 15) & gm, Z = (bi >> 8) & gm;
  const double px = (double)(v.origin[0] + cx) * 1.0 * 16.0 + (double)X * 1.0;
  const double py = (double)(v.origin[1] + cy) * 1.0 * 16.0 + (double)Y * 1.0;
  const double pz = (double)(v.origin[2] + cz) * 1.0 * 16.0 + (double)Z * 1.0;
  const bool solid = sdf_solid<KIND>(sp, px, py, pz);
  const uint32_t m = __ballot_sync(0xffffffffu, solid);
  __shared__ uint32_t s_m[2];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t o = (uint64_t)s_m[0] | ((uint64_t)s_m[1] << 32);
    v.occ[word_global] = o;
    v.full[word_global] = o;
  }
}

// chunk-level any/full bit grids: one warp per chunk
template <bool SET_REGION>
__device__ __forceinline__ void finalize_chunk(const DVolume& v, int64_t c, int lane) {
  uint64_t o0 = v.occ[c * 64 + lane], o1 = v.occ[c * 64 + 32 + lane];
  uint64_t f0 = v.full[c * 64 + lane], f1 = v.full[c * 64 + 32 + lane];
  v.of[c * 64 + lane] = make_ulonglong2(o0, f0);
  v.of[c * 64 + 32 + lane] = make_ulonglong2(o1, f1);
  const bool any = __any_sync(0xffffffffu, (o0 | o1) != 0ull);
  const bool all = __all_sync(0xffffffffu, (f0 & f1) == ~0ull);
  // 32^3-voxel cells (4x4x4 bricks): word w = bz*4 + by/4 feeds cells (x/4 = nibble, y/4 = w & 3, z/4 = w >> 4)
  uint32_t lo = 0, hi = 0;
  {
    const uint32_t r0 = (uint32_t)((o0 | (o0 >> 16) | (o0 >> 32) | (o0 >> 48)) & 0xFFFFull);
    const uint32_t r1 = (uint32_t)((o1 | (o1 >> 16) | (o1 >> 32) | (o1 >> 48)) & 0xFFFFull);
    uint32_t n0 = 0, n1 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { n0 |= ((r0 >> (4 * i)) & 0xFu) ? (1u << i) : 0u; n1 |= ((r1 >> (4 * i)) & 0xFu) ? (1u << i) : 0u; }
    lo = n0 << (4 * (lane & 3) + 16 * (lane >> 4));          // words 0..31  -> z/4 in {0,1}
    hi = n1 << (4 * (lane & 3) + 16 * (lane >> 4));          // words 32..63 -> z/4 in {2,3}
  }
  lo = __reduce_or_sync(0xffffffffu, lo);
  hi = __reduce_or_sync(0xffffffffu, hi);
  if (lane == 0) {
    v.cells[c] = (uint64_t)lo | ((uint64_t)hi << 32);
    if (any) atomicOr(&v.chunk_any[c >> 5], 1u << (c & 31));
    if (all) atomicOr(&v.chunk_full[c >> 5], 1u << (c & 31));
    if (SET_REGION && any) {
      const int cx = (int)(c % v.dims[0]), cy = (int)((c / v.dims[0]) % v.dims[1]), cz = (int)(c / ((int64_t)v.dims[0] * v.dims[1]));
      const int r = (cx >> 2) + v.rdims[0] * ((cy >> 2) + v.rdims[1] * (cz >> 2));
      atomicOr(&v.region_any[r >> 5], 1u << (r & 31));
    }
  }
}

__global__ void __launch_bounds__(256) finalize_kernel(DVolume v) {
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= v.nchunks) return;
  finalize_chunk<false>(v, c, threadIdx.x & 31);
}

// streaming (K6): derived data of the listed, newly generated chunks only; bits are only ever added
__global__ void __launch_bounds__(256) finalize_list_kernel(DVolume v, const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_n) {
  const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= (int64_t)*list_n) return;
  finalize_chunk<true>(v, (int64_t)list[k], threadIdx.x & 31);
}

// bit per 512^3-voxel region (4x4x4 chunks): one thread per region
__global__ void region_kernel(DVolume v) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int nr = v.rdims[0] * v.rdims[1] * v.rdims[2];
  if (r >= nr) return;
  const int rx = r % v.rdims[0], ry = (r / v.rdims[0]) % v.rdims[1], rz = r / (v.rdims[0] * v.rdims[1]);
  bool any = false;
  for (int z = rz * 4; z < min(rz * 4 + 4, v.dims[2]); z++)
    for (int y = ry * 4; y < min(ry * 4 + 4, v.dims[1]); y++)
      for (int x = rx * 4; x < min(rx * 4 + 4, v.dims[0]); x++) {
        const int ci = x + v.dims[0] * (y + v.dims[1] * z);
        any |= (v.chunk_any[ci >> 5] >> (ci & 31)) & 1u;
      }
  if (any) atomicOr(&v.region_any[r >> 5], 1u << (r & 31));
}

// ---- 2^3-cell masks of the partial bricks (raymarch: skipping inside a brick) ----------------------------------
// One thread per payload slot in use: OR the eight slices pairwise in z, then in x and y with shifts, and gather the
// sixteen 2x2 results of every z pair.
__global__ void __launch_bounds__(256) coarse_mask_kernel(DVolume v) {
  const uint32_t slot = blockIdx.x * 256 + threadIdx.x;
  const uint32_t n = min(*v.pool_count, v.max_bricks);
  if (slot >= n) return;
  const uint64_t* p = v.pool + (size_t)slot * 8;
  uint64_t cm = 0;
#pragma unroll
  for (int ez = 0; ez < 4; ez++) {
    uint64_t q = p[2 * ez] | p[2 * ez + 1];
    q |= q >> 1;    // even x: OR of the x pair
    q |= q >> 8;    // even y rows: OR of the y pair
#pragma unroll
    for (int ey = 0; ey < 4; ey++)
#pragma unroll
      for (int ex = 0; ex < 4; ex++) cm |= ((q >> (2 * ex + 16 * ey)) & 1ull) << (ex + 4 * ey + 16 * ez);
  }
  v.pool_cm[slot] = cm;
}
void launch_coarse_masks(const LaunchCtx& lc, const DVolume& v) {
  coarse_mask_kernel<<<(v.max_bricks + 255) / 256, 256, 0, lc.stream>>>(v);
  (*lc.launches)++;
}

// ---- cell distance field (raymarch empty-space skipping) -------------------------------------------------------
// df(c) = min over non-empty 32^3 cells e of max(|cx-ex|, |cy-ey|, |cz-ez|), capped at K + 1: the cube of half-width
// df - 1 cells around c is empty.  The L-infinity transform is separable: three identical passes
//     out(i) = min over |k| <= K of max(in(i + k), |k|)
// along x, y, z, starting from 0 / cap (outside the grid counts as empty).  One thread per cell, 63 byte loads from L1.
__global__ void __launch_bounds__(256) df_seed_kernel(DVolume v) {
  const int64_t n = (int64_t)v.ddims[0] * v.ddims[1] * v.ddims[2];
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int ex = (int)(i % v.ddims[0]), ey = (int)((i / v.ddims[0]) % v.ddims[1]), ez = (int)(i / ((int64_t)v.ddims[0] * v.ddims[1]));
  const int64_t c = (ex >> 2) + (int64_t)v.dims[0] * ((ey >> 2) + (int64_t)v.dims[1] * (ez >> 2));
  const int e = (ex & 3) + 4 * (ey & 3) + 16 * (ez & 3);
  v.df[i] = ((v.cells[c] >> e) & 1ull) ? 0 : (MESO_DF_K + 1);
}
template <int AXIS>
__global__ void __launch_bounds__(256) df_pass_kernel(DVolume v, const uint8_t* __restrict__ in, uint8_t* __restrict__ out) {
  const int64_t n = (int64_t)v.ddims[0] * v.ddims[1] * v.ddims[2];
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int ex = (int)(i % v.ddims[0]), ey = (int)((i / v.ddims[0]) % v.ddims[1]), ez = (int)(i / ((int64_t)v.ddims[0] * v.ddims[1]));
  const int p = AXIS == 0 ? ex : (AXIS == 1 ? ey : ez);
  const int len = v.ddims[AXIS];
  const int64_t stride = AXIS == 0 ? 1 : (AXIS == 1 ? v.ddims[0] : (int64_t)v.ddims[0] * v.ddims[1]);
  int best = in[i];
  for (int a = 1; a <= MESO_DF_K && a < best; a++) {  // outwards: a tap at distance a cannot give less than a
    if (p - a >= 0) best = min(best, max((int)in[i - a * stride], a));
    if (p + a < len) best = min(best, max((int)in[i + a * stride], a));
  }
  out[i] = (uint8_t)best;
}

void launch_df_build(const LaunchCtx& lc, const DVolume& v) {
  const int64_t n = (int64_t)v.ddims[0] * v.ddims[1] * v.ddims[2];
  const unsigned grid = (unsigned)((n + 255) / 256);
  df_seed_kernel<<<grid, 256, 0, lc.stream>>>(v);
  df_pass_kernel<0><<<grid, 256, 0, lc.stream>>>(v, v.df, v.df_tmp);
  df_pass_kernel<1><<<grid, 256, 0, lc.stream>>>(v, v.df_tmp, v.df);
  df_pass_kernel<2><<<grid, 256, 0, lc.stream>>>(v, v.df, v.df_tmp);
  cudaMemcpyAsync(v.df, v.df_tmp, (size_t)n, cudaMemcpyDeviceToDevice, lc.stream);
  (*lc.launches) += 4;
}

__global__ void scatter_payload_kernel(DVolume v, const uint64_t* __restrict__ keys, const uint64_t* __restrict__ payload, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 3; const int j = (int)(t & 7);
  if (i >= n) return;
  v.pool[i * 8 + j] = payload[i * 8 + j];
  if (j == 0) v.bptr[keys[i]] = (uint32_t)i;
}

__global__ void gather_partial_kernel(DVolume v, uint64_t* keys, uint64_t* payload, uint32_t* count) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per block of the grid
  if (t >= v.nchunks * MESO_BLOCKS) return;
  const int64_t c = t >> 12; const int b = (int)(t & 4095);
  const uint64_t o = v.occ[c * 64 + (b >> 6)], f = v.full[c * 64 + (b >> 6)];
  if (!((o & ~f) >> (b & 63) & 1ull)) return;
  const uint32_t idx = atomicAdd(count, 1u);
  keys[idx] = (uint64_t)t;
  const uint64_t* src = v.pool + (size_t)v.bptr[t] * 8;
#pragma unroll
  for (int j = 0; j < 8; j++) payload[(size_t)idx * 8 + j] = src[j];
}

// voxel granularity: word pre-pass (uniform words written, mixed words listed), then the persistent per-voxel kernel over
// the list.  chunk_list == null: chunks [0, n_chunks); else chunk_list[0 .. *chunk_n), launch sized for n_chunks entries.
static void launch_voxelize_words(const LaunchCtx& lc, const DVolume& v, int kind, const SdfParams& sp, int* g_overflow,
                                  const uint32_t* chunk_list, const uint32_t* chunk_n, int64_t n_chunks) {
  cudaMemsetAsync(v.n_words, 0, sizeof(uint32_t), lc.stream);
  const unsigned cgrid = (unsigned)((n_chunks * 64 + 255) / 256);
  const unsigned vgrid = (unsigned)lc.sm_count * 8u;
  if (kind == MESO_SDF_SPHERE) {
    classify_words_kernel<MESO_SDF_SPHERE><<<cgrid, 256, 0, lc.stream>>>(v, sp, chunk_list, chunk_n, n_chunks, v.words, v.n_words);
    voxelize_voxel_kernel<MESO_SDF_SPHERE><<<vgrid, 256, 0, lc.stream>>>(v, sp, g_overflow, v.words, v.n_words);
  } else {
    classify_words_kernel<MESO_SDF_TERRAIN><<<cgrid, 256, 0, lc.stream>>>(v, sp, chunk_list, chunk_n, n_chunks, v.words, v.n_words);
    voxelize_voxel_kernel<MESO_SDF_TERRAIN><<<vgrid, 256, 0, lc.stream>>>(v, sp, g_overflow, v.words, v.n_words);
  }
  (*lc.launches)++;
}

void launch_voxelize(const LaunchCtx& lc, const DVolume& v, int kind, const double params[4], int granularity, int* g_overflow, int mip) {
  SdfParams sp;
  for (int i = 0; i < 4; i++) sp.p[i] = params ? params[i] : 0.0;
  cudaMemsetAsync(v.bptr, 0xFF, sizeof(uint32_t) * MESO_BLOCKS * (size_t)v.nchunks, lc.stream);
  cudaMemsetAsync(v.pool_count, 0, sizeof(uint32_t), lc.stream);
  cudaMemsetAsync(v.pool_free_count, 0, sizeof(int), lc.stream);
  const unsigned grid = (unsigned)(v.nchunks * 64);
  if (granularity == MESO_GRAN_VOXEL) {
    launch_voxelize_words(lc, v, kind, sp, g_overflow, nullptr, nullptr, v.nchunks);
  } else {
    if (kind == MESO_SDF_SPHERE) voxelize_block_kernel<MESO_SDF_SPHERE><<<grid, 64, 0, lc.stream>>>(v, sp, nullptr, nullptr, mip);
    else voxelize_block_kernel<MESO_SDF_TERRAIN><<<grid, 64, 0, lc.stream>>>(v, sp, nullptr, nullptr, mip);
  }
  (*lc.launches)++;
  launch_volume_finalize(lc, v);
}

// K6 streaming: generate only the chunks d_list[0 .. *d_n) (grid slots), at most max_n of them; the rest of the volume is
// untouched.  The launch is sized for max_n; CTAs beyond *d_n exit (no host readback between selection and generation).
void launch_voxelize_list(const LaunchCtx& lc, const DVolume& v, int kind, const double params[4], int granularity, int* g_overflow,
                          const uint32_t* d_list, const uint32_t* d_n, uint32_t max_n) {
  if (max_n == 0) return;
  SdfParams sp;
  for (int i = 0; i < 4; i++) sp.p[i] = params ? params[i] : 0.0;
  const unsigned grid = max_n * 64u;
  if (granularity == MESO_GRAN_VOXEL) {
    launch_voxelize_words(lc, v, kind, sp, g_overflow, d_list, d_n, (int64_t)max_n);
  } else {
    if (kind == MESO_SDF_SPHERE) voxelize_block_kernel<MESO_SDF_SPHERE><<<grid, 64, 0, lc.stream>>>(v, sp, d_list, d_n, 0);
    else voxelize_block_kernel<MESO_SDF_TERRAIN><<<grid, 64, 0, lc.stream>>>(v, sp, d_list, d_n, 0);
  }
  finalize_list_kernel<<<(max_n + 7) / 8, 256, 0, lc.stream>>>(v, d_list, d_n);
  (*lc.launches) += 2;
  launch_coarse_masks(lc, v);
  launch_df_build(lc, v);   // voxels were added: the distance field must not overestimate
}

void launch_volume_finalize(const LaunchCtx& lc, const DVolume& v, bool rebuild_df) {
  cudaMemsetAsync(v.chunk_any, 0, sizeof(uint32_t) * v.chunk_words, lc.stream);
  cudaMemsetAsync(v.chunk_full, 0, sizeof(uint32_t) * v.chunk_words, lc.stream);
  cudaMemsetAsync(v.region_any, 0, sizeof(uint32_t) * v.region_words, lc.stream);
  finalize_kernel<<<(unsigned)((v.nchunks + 7) / 8), 256, 0, lc.stream>>>(v);
  const int nr = v.rdims[0] * v.rdims[1] * v.rdims[2];
  region_kernel<<<(nr + 127) / 128, 128, 0, lc.stream>>>(v);
  (*lc.launches) += 2;
  launch_coarse_masks(lc, v);
  if (rebuild_df) launch_df_build(lc, v);
}

void launch_scatter_payload(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_keys, const uint64_t* d_payload, int64_t n) {
  if (n <= 0) return;
  scatter_payload_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, lc.stream>>>(v, d_keys, d_payload, n);
  (*lc.launches)++;
}

void launch_gather_partial(const LaunchCtx& lc, const DVolume& v, uint64_t* d_keys, uint64_t* d_payload, uint32_t* d_count) {
  cudaMemsetAsync(d_count, 0, sizeof(uint32_t), lc.stream);
  const int64_t n = v.nchunks * MESO_BLOCKS;
  gather_partial_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(v, d_keys, d_payload, d_count);
  (*lc.launches)++;
}
