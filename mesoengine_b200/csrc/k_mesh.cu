// k_mesh.cu -- K3: face culling + greedy merge of exposed voxel faces into compacted quad lists.
//
// The north-star form of the reference's "mesher" (hidden-block cull + instance compaction,
// Runtimes/Voxel/Chunk/ChunkPool.h:381-445; its draw picks faces per view in the VS, Samples/SimpleVoxel.cpp:129-136).
// Neither quads nor a face-neighbour test exist in the reference: the definition is oracle/orc_mesh.c
// ("parity unpinned by reference; bit-exact vs repo oracle" after the canonical sort).
//
// Two levels (definition: oracle/orc_mesh.c).  Voxel level: the exposed voxel faces of a brick, merged inside the brick (passes A
// and B below).  Brick level: the face of a FULL brick towards an ABSENT brick is not a voxel-level face; those faces are merged
// per chunk, direction and brick layer over 16x16-brick images (pass C) -- the block-granular scenes of the reference consist of
// nothing else, and a flat chunk face is one quad instead of 256.
//
// Pass A (worklist): one warp per chunk of the rank; partial bricks, and full bricks with at least one PARTIAL neighbour (decided
// with word-wide shifts), are compacted with a warp prefix sum + one atomic per warp.
// Pass B (mesh): a warp pass = 32 bricks x one axis.  Phase 1, one THREAD per brick: its eight z-slices in registers, the eight
// planes along the axis formed from them, the facing plane of the two neighbours, sixteen (direction, layer) 8x8 images, the
// non-empty ones queued in shared memory.  Phase 2, a converged loop: every lane takes ONE quad off its current image per
// iteration with a branch-free greedy step and fetches the next queued image when its own is used up; quads staged per warp in
// shared memory and written behind one reservation per batch as one bulk asynchronous store.
// HBM-bound integer work on paper: 64 B per populated brick + 16 B per quad (+ neighbour slices, mostly L1 / L2 hits).
#include "meso_internal.cuh"

// chunk index -> coordinates in 32-bit arithmetic (a 64-bit division per thread costs more than the loads it guards)
__device__ __forceinline__ void chunk_coords(const DVolume& v, int64_t c, int& cx, int& cy, int& cz) {
  const uint32_t u = (uint32_t)c, dx = (uint32_t)v.dims[0], dy = (uint32_t)v.dims[1];
  const uint32_t q = u / dx;
  cx = (int)(u - q * dx); cz = (int)(q / dy); cy = (int)(q - (uint32_t)cz * dy);
}
// {occupancy, full} word w of chunk (cx,cy,cz), zeros outside the grid
__device__ __forceinline__ ulonglong2 of_word(const DVolume& v, int cx, int cy, int cz, int w) {
  if ((unsigned)cx >= (unsigned)v.dims[0] || (unsigned)cy >= (unsigned)v.dims[1] || (unsigned)cz >= (unsigned)v.dims[2]) return make_ulonglong2(0ull, 0ull);
  return __ldg(&v.of[chunk_index(v, cx, cy, cz) * 64 + w]);
}
__device__ __forceinline__ uint64_t partial_of(const ulonglong2 p) { return p.x & ~p.y; }

// The six neighbour words of 64-brick word w (= z*4 + y/4, four 16-bit x-rows) of chunk c, each brick's neighbour moved onto
// the brick's own bit: m[2a] = minus side, m[2a+1] = plus side of axis a.  F picks the bit plane: occupancy, full or partial.
template <class F>
__device__ __forceinline__ void neighbour_words(const DVolume& v, int64_t c, int w, F pick, uint64_t m[6]) {
  int cx, cy, cz;
  chunk_coords(v, c, cx, cy, cz);
  const int z = w >> 2, yq = w & 3;
  const ulonglong2* own = v.of + c * 64;
  const uint64_t X0 = 0x0001000100010001ull, X15 = 0x8000800080008000ull;
  const uint64_t me = pick(__ldg(own + w));
  m[0] = ((me << 1) & ~X0) | ((pick(of_word(v, cx - 1, cy, cz, w)) >> 15) & X0);
  m[1] = ((me >> 1) & ~X15) | ((pick(of_word(v, cx + 1, cy, cz, w)) << 15) & X15);
  const uint64_t ym_src = pick(yq > 0 ? __ldg(own + w - 1) : of_word(v, cx, cy - 1, cz, z * 4 + 3));
  const uint64_t yp_src = pick(yq < 3 ? __ldg(own + w + 1) : of_word(v, cx, cy + 1, cz, z * 4 + 0));
  m[2] = (me << 16) | (ym_src >> 48);
  m[3] = (me >> 16) | (yp_src << 48);
  m[4] = pick(z > 0 ? __ldg(own + w - 4) : of_word(v, cx, cy, cz - 1, 15 * 4 + yq));
  m[5] = pick(z < 15 ? __ldg(own + w + 4) : of_word(v, cx, cy, cz + 1, 0 * 4 + yq));
}

__device__ __forceinline__ uint32_t chunk_hash_rank(uint64_t c, int world) { return (uint32_t)(((c * 0x9E3779B97F4A7C15ull) >> 40) % (unsigned)world); }
__device__ __forceinline__ bool chunk_wholly_full(const DVolume& v, int cx, int cy, int cz) {
  if ((unsigned)cx >= (unsigned)v.dims[0] || (unsigned)cy >= (unsigned)v.dims[1] || (unsigned)cz >= (unsigned)v.dims[2]) return false;
  const int64_t ci = chunk_index(v, cx, cy, cz);
  return (__ldg(&v.chunk_full[ci >> 5]) >> (ci & 31)) & 1u;
}

// false: the chunk cannot own a face of either level -- not this rank's, no brick, or wholly full inside wholly full neighbours
__device__ __forceinline__ bool chunk_may_have_faces(const DVolume& v, int64_t c, bool by_hash, int rank, int world) {
  if (by_hash ? (world > 1 && (int)chunk_hash_rank((uint64_t)c, world) != rank) : (world > 1 && ((uint32_t)c % (uint32_t)world) != (uint32_t)rank)) return false;
  if (!((__ldg(&v.chunk_any[c >> 5]) >> (c & 31)) & 1u)) return false;
  int cx, cy, cz;
  chunk_coords(v, c, cx, cy, cz);
  return !(chunk_wholly_full(v, cx, cy, cz) && chunk_wholly_full(v, cx - 1, cy, cz) && chunk_wholly_full(v, cx + 1, cy, cz) && chunk_wholly_full(v, cx, cy - 1, cz) &&
           chunk_wholly_full(v, cx, cy + 1, cz) && chunk_wholly_full(v, cx, cy, cz - 1) && chunk_wholly_full(v, cx, cy, cz + 1));
}

// Pass A, one WARP per chunk of this rank (c = rank + world * i), two 64-brick words per lane.  The chunk is dismissed first by
// its chunk bits, looked up by eight lanes at once (no brick; or wholly full inside six wholly full neighbours).  A full brick
// has voxel-level faces only towards PARTIAL neighbours (towards a full one the face is hidden, towards an absent one it belongs
// to the brick level).  Survivors are appended to the work list (warp prefix + one atomic per warp); the chunks that get this far
// form pass C's list.
__global__ void __launch_bounds__(256) mesh_worklist_kernel(DVolume v, int rank, int world, uint64_t* work, int64_t work_cap, uint32_t* work_count,
                                                            uint32_t* full_count, uint32_t* chunk_list, uint32_t* chunk_count) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)rank + (int64_t)world * ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (c >= v.nchunks) return;
  int cx, cy, cz;
  chunk_coords(v, c, cx, cy, cz);
  bool flag = false;
  if (lane == 0) flag = (__ldg(&v.chunk_any[c >> 5]) >> (c & 31)) & 1u;
  else if (lane < 8) {
    const int k = lane - 2;   // lane 1: the chunk itself; lanes 2..7: its neighbours -x +x -y +y -z +z
    flag = chunk_wholly_full(v, cx + (k == 0 ? -1 : k == 1 ? 1 : 0), cy + (k == 2 ? -1 : k == 3 ? 1 : 0), cz + (k == 4 ? -1 : k == 5 ? 1 : 0));
  }
  const unsigned bits = __ballot_sync(0xffffffffu, flag);
  if (!(bits & 1u) || (bits & 0xFEu) == 0xFEu) return;
  if (lane == 0) chunk_list[atomicAdd(chunk_count, 1u)] = (uint32_t)c;
  uint64_t todo[2], todo_full[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int w = lane + 32 * k;
    const ulonglong2 P = __ldg(&v.of[c * 64 + w]);
    todo[k] = P.x & ~P.y;
    todo_full[k] = 0ull;
    if (P.y) {
      uint64_t pn[6];
      neighbour_words(v, c, w, [](const ulonglong2 p) { return partial_of(p); }, pn);
      todo_full[k] = P.y & (pn[0] | pn[1] | pn[2] | pn[3] | pn[4] | pn[5]);
    }
  }
  // two lists in one buffer: partial bricks from the front, full bricks from the back -- pass B's warp passes are then uniform
  // (a full brick costs two neighbour planes, a partial one sixteen images: mixed in one pass the full lanes would idle)
  const uint32_t n = (uint32_t)(__popcll(todo[0]) + __popcll(todo[1])), nf = (uint32_t)(__popcll(todo_full[0]) + __popcll(todo_full[1]));
  uint32_t incl = n | (nf << 16);                    // both prefix sums in one word (at most 128 bricks per lane, 4096 per warp)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  uint32_t base = 0, base_f = 0;
  if (lane == 0) {
    if (total & 0xFFFFu) base = atomicAdd(work_count, total & 0xFFFFu);
    if (total >> 16) base_f = atomicAdd(full_count, total >> 16);
  }
  base = __shfl_sync(0xffffffffu, base, 0) + (incl & 0xFFFFu) - n;
  base_f = __shfl_sync(0xffffffffu, base_f, 0) + (incl >> 16) - nf;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    uint64_t t = todo[k];
    while (t) {
      const int bit = __ffsll((long long)t) - 1;
      t &= t - 1;
      work[base++] = (uint64_t)c * MESO_BLOCKS + (uint64_t)((lane + 32 * k) * 64 + bit);
    }
    t = todo_full[k];
    while (t) {
      const int bit = __ffsll((long long)t) - 1;
      t &= t - 1;
      work[work_cap - 1 - (int64_t)(base_f++)] = (uint64_t)c * MESO_BLOCKS + (uint64_t)((lane + 32 * k) * 64 + bit);
    }
  }
}

__device__ __forceinline__ uint32_t gather_col(uint64_t e, int x) {  // bits (x + 8y) -> byte with bit y
  return (uint32_t)((((e >> x) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}

#ifndef MQ_CAP
#define MQ_CAP 320    // quads a warp can stage in shared memory (32 bricks x one axis of a smooth surface yield ~230)
#endif
#define MQ_FLUSH (MQ_CAP / 2)  // staged quads are written out once this many have accumulated (and at the end of the warp's work)

// The loop form of the greedy merge of one 8x8 image, as the oracle writes it.  Only the FALLBACK of phase 1 when the warp's
// image queue is full (more than MI_CAP non-empty images in one pass: checkerboard bricks): the image is merged on the spot, quads
// staged in shared memory (slot = shared atomic), and what does not fit the staging area goes straight to the global list.
// (not inlined: it has twelve call sites in the kernel -- inlined they were a third of its 4 500 instructions and the kernel
// stalled 11 % of the time on instruction fetch)
__device__ __noinline__ void greedy_stage(uint64_t img, int dir, int layer, int ox, int oy, int oz, uint4* stage, int* count,
                                             MesoQuad* quads, int64_t cap, unsigned long long* quad_count) {
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
      const uint4 q = make_uint4((uint32_t)(ox + x) | ((uint32_t)(oy + y) << 16), (uint32_t)(oz + z) | ((uint32_t)dir << 16) | ((uint32_t)w << 24),
                                 (uint32_t)h, 0u);
      const int idx = atomicAdd(count, 1);
      if (idx < MQ_CAP) stage[idx] = q;
      else {
        const unsigned long long g = atomicAdd_system(quad_count, 1ull);
        if ((int64_t)g < cap) reinterpret_cast<uint4*>(quads)[g] = q;
      }
    }
    if (vv == 7 || (img >> (8 * (vv + 1))) == 0ull) break;   // nothing left in the rows below
  }
}

// One step of the greedy merge on a 64-bit image (row v = byte v, bit u), branch-free: the lowest set bit is (first non-empty
// row v, lowest u0) = the quad the fixed merge order takes next; its run length w from the row; its height h = index of the
// first following row that misses a bit of the run (rows past the image shift in as zeros, which bounds h by itself); the h rows
// of the run are cleared in one go.  meta = x0 | y0 << 16 | z0 << 32 | dir << 48 | layer << 52 (voxel origin of the brick: the
// sums cannot carry).
__device__ __forceinline__ uint4 take_quad(uint64_t& img, uint64_t meta) {
  const uint32_t axis = (uint32_t)(meta >> 49) & 3u;
  const int p = (__ffsll((long long)img) - 1) & 63;    // (& 63: an exhausted image, img == 0, stays well-defined and 0)
  const int vv = p >> 3, u0 = p & 7, sh = p & ~7;
  const uint64_t t = img >> sh;                       // rows v, v+1, ... in bytes 0, 1, ...
  const uint32_t row = (uint32_t)t & 0xFFu;
  const int w = __ffs((int)~(row >> u0)) - 1;
  const uint32_t m = ((1u << w) - 1u) << u0;
  const uint32_t m4 = m * 0x01010101u;
  const uint64_t mrep = ((uint64_t)m4 << 32) | m4;    // the run in every row
  const uint64_t miss = ~t & mrep;
  const int h = miss ? ((__ffsll((long long)miss) - 1) >> 3) : 8;
  const uint64_t clr = h == 8 ? mrep : (mrep & ((1ull << (8 * h)) - 1ull));
  img &= ~(clr << sh);
  const uint32_t layer = (uint32_t)(meta >> 52) & 7u;
  const uint32_t qx = axis == 0 ? layer : (uint32_t)u0;
  const uint32_t qy = axis == 0 ? (uint32_t)u0 : (axis == 1 ? layer : (uint32_t)vv);
  const uint32_t qz = axis == 2 ? layer : (uint32_t)vv;
  return make_uint4((uint32_t)meta + (qx | (qy << 16)), (((uint32_t)(meta >> 32) & 0x0007FFFFu) + qz) | ((uint32_t)w << 24), (uint32_t)h, 0u);
}

// state of a brick: 0 absent / outside the grid, 1 full, 2 partial (slot = its payload)
__device__ __forceinline__ int brick_state(const DVolume& v, int bx, int by, int bz, uint32_t& slot) {
  if ((unsigned)bx >= (unsigned)(v.dims[0] * 16) || (unsigned)by >= (unsigned)(v.dims[1] * 16) || (unsigned)bz >= (unsigned)(v.dims[2] * 16)) return 0;
  const int64_t nc = chunk_index(v, bx >> 4, by >> 4, bz >> 4);
  const int nb = block_bit(bx & 15, by & 15, bz & 15);
  const ulonglong2 p = __ldg(&v.of[nc * 64 + (nb >> 6)]);
  if (!((p.x >> (nb & 63)) & 1ull)) return 0;
  if ((p.y >> (nb & 63)) & 1ull) return 1;
  slot = __ldg(&v.bptr[nc * MESO_BLOCKS + nb]);
  return 2;
}
__device__ __forceinline__ void load_slices(const DVolume& v, uint32_t slot, uint64_t s[8]) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(v.pool + (size_t)slot * 8);
#pragma unroll
  for (int i = 0; i < 4; i++) { const ulonglong2 q = __ldg(p + i); s[2 * i] = q.x; s[2 * i + 1] = q.y; }
}
// (Staging the 64-byte payloads of a pass in shared memory with one cp.async.bulk per lane + an mbarrier wait + LDS reads was
// built and measured: correct, and slower than four LDG.128 per lane -- 0.79 vs 0.57 ms at 4096^3 with the kernel of that
// time: 8 KB more shared memory per CTA, a 4-way bank conflict on the read-back, and nothing to hide.  profiles/README.md.
// The bulk-copy engine is used where it pays: the staged quad batches leave shared memory as one bulk store each.)
#ifndef MB_BULK_STORE
#define MB_BULK_STORE 1
#endif
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// plane `l` of the brick along `axis`, as the 8x8 image the oracle defines for that axis:
//   axis 0 (x = l): row z, bit y;   axis 1 (y = l): row z, bit x;   axis 2 (z = l): row y, bit x (the slice itself)
template <int AXIS>
__device__ __forceinline__ uint64_t axis_plane(const uint64_t s[8], int l) {
  if (AXIS == 2) {   // select chain, not s[l]: a dynamically indexed register array would live in local memory
    const uint64_t a = (l & 1) ? s[1] : s[0], b = (l & 1) ? s[3] : s[2], c = (l & 1) ? s[5] : s[4], d = (l & 1) ? s[7] : s[6];
    const uint64_t ab = (l & 2) ? b : a, cd = (l & 2) ? d : c;
    return (l & 4) ? cd : ab;
  }
  uint64_t p = 0;
#pragma unroll
  for (int z = 0; z < 8; z++) {
    const uint32_t row = AXIS == 0 ? gather_col(s[z], l) : ((uint32_t)(s[z] >> (8 * l)) & 0xFFu);
    p |= (uint64_t)row << (8 * z);
  }
  return p;
}
// the plane of a neighbouring brick that faces this one: its layer 7 (the neighbour on the minus side) or 0 (plus side)
// (`absent`: what an absent neighbour counts as -- empty for a partial brick, covered for a full one, whose face towards it
// belongs to the brick level)
template <int AXIS>
__device__ __forceinline__ uint64_t neighbour_plane(const DVolume& v, int bx, int by, int bz, int layer, uint64_t absent) {
  uint32_t slot = 0;
  const int st = brick_state(v, bx, by, bz, slot);
  if (st == 0) return absent;
  if (st == 1) return ~0ull;
  if (AXIS == 2) return __ldg(&v.pool[(size_t)slot * 8 + layer]);
  uint64_t s[8];
  load_slices(v, slot, s);
  return axis_plane<AXIS>(s, layer);
}

#ifndef MB_THREADS
#define MB_THREADS 128   // 4 warps per CTA: 20 KB of staged quads + 16 KB of queued images in static shared memory
#endif
#define MB_WARPS (MB_THREADS / 32)
#ifndef MI_CAP
#define MI_CAP 256       // (direction, layer) images a warp can queue per pass (32 bricks x one axis: typically 100-200 non-empty)
#endif
#ifndef MB_MINB
#define MB_MINB 6
#endif

// Phase 1 of a warp pass: the sixteen (direction, layer) exposed-face images of one brick (state st: 0 nothing, 1 full, 2 partial
// with its eight z-slices in s) along AXIS; the non-empty ones go into the warp's queue as {image, origin | direction | layer}
// (a full queue merges the image on the spot).
template <int AXIS>
__device__ __forceinline__ void queue_axis_images(const DVolume& v, int st, const uint64_t s[8], int bx, int by, int bz, ulonglong2* queue, int* qn, uint4* stage, int* count,
                                                  MesoQuad* quads, int64_t cap, unsigned long long* quad_count) {
  if (st == 0) return;
  const uint64_t org = (uint64_t)(uint32_t)(bx * 8) | ((uint64_t)(uint32_t)(by * 8) << 16) | ((uint64_t)(uint32_t)(bz * 8) << 32);
  auto enqueue = [&](uint64_t img, int dir, int l) {
    if (img) {
      const int idx = atomicAdd(qn, 1);
      if (idx < MI_CAP) queue[idx] = make_ulonglong2(img, org | ((uint64_t)dir << 48) | ((uint64_t)l << 52));
      else greedy_stage(img, dir, l, bx * 8, by * 8, bz * 8, stage, count, quads, cap, quad_count);
    }
  };
  if (st == 1) {
    // a full brick shows voxel faces only on its two outer layers, where the neighbour is a partial brick (an absent
    // neighbour counts as covered: that face belongs to the brick level)
    enqueue(~neighbour_plane<AXIS>(v, bx - (AXIS == 0), by - (AXIS == 1), bz - (AXIS == 2), 7, ~0ull), 2 * AXIS, 0);
    enqueue(~neighbour_plane<AXIS>(v, bx + (AXIS == 0), by + (AXIS == 1), bz + (AXIS == 2), 0, ~0ull), 2 * AXIS + 1, 7);
    return;
  }
  const uint64_t nbm = neighbour_plane<AXIS>(v, bx - (AXIS == 0), by - (AXIS == 1), bz - (AXIS == 2), 7, 0ull);
  const uint64_t nbp = neighbour_plane<AXIS>(v, bx + (AXIS == 0), by + (AXIS == 1), bz + (AXIS == 2), 0, 0ull);
  uint64_t prev = nbm, cur = axis_plane<AXIS>(s, 0);
#pragma unroll 1
  for (int l = 0; l < 8; l++) {
    const uint64_t next = l < 7 ? axis_plane<AXIS>(s, l + 1) : nbp;
    enqueue(cur & ~prev, 2 * AXIS, l);
    enqueue(cur & ~next, 2 * AXIS + 1, l);
    prev = cur; cur = next;
  }
}

// Persistent, grid-stride over the work list (count read from device memory: no host round trip between the passes).
// A warp pass = 32 bricks x one axis.  Phase 1, one THREAD per brick: the brick's eight z-slices in registers, the eight
// planes along the axis formed from them (the slices themselves for z, byte / column gathers for y / x), the facing plane of
// the two neighbours, sixteen exposed-face images; the non-empty ones are queued in shared memory.  Phase 2 (`merge` below): a
// converged loop in which every lane takes one quad off its image per iteration and the next queued image when that is used up.
// Quads are staged per warp and leave in batches behind one reservation.
// History (profiles/README.md): round 1's warp-per-two-bricks form spent 55 % of its instructions building slices and images
// cooperatively and ran the greedy loops at 2-6 active lanes; one thread per (brick, axis) without the queue ran them at 4;
// the queue with a data-dependent greedy loop per image (lanes round-robin over the images) at 8; the converged one-quad-per-
// iteration loop runs at 26 (0.58 -> 0.41 ms), and letting the lanes keep their images across passes removes its tail (0.385 ms).
// Work: either pass A's two lists (counts read from device memory: partial bricks work[0 .. np), full bricks with a partial
// neighbour work[work_cap - 1 - i], i < nf) or one list of n_list keys of any state (re-mesh).
__global__ void __launch_bounds__(MB_THREADS, MB_MINB) mesh_bricks_kernel(DVolume v, const uint64_t* __restrict__ work, int64_t work_cap,
                                                                    const uint32_t* __restrict__ partial_count_ptr, const uint32_t* __restrict__ full_count_ptr,
                                                                    uint32_t n_list, MesoQuad* quads, int64_t cap, unsigned long long* quad_count,
                                                                    int shard_rank, int shard_world) {
  __shared__ __align__(128) uint4 s_q[MB_WARPS][MQ_CAP];
  __shared__ ulonglong2 s_img[MB_WARPS][MI_CAP];
  __shared__ int s_n[MB_WARPS], s_in[MB_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_work = partial_count_ptr ? *partial_count_ptr : n_list;
  const uint32_t n_full = full_count_ptr ? *full_count_ptr : 0u;
  const int64_t g_first = (((int64_t)n_work + 31) / 32) * 3;      // (32 bricks, axis) passes over the first list ...
  const int64_t n_groups = g_first + ((int64_t)n_full + 31) / 32; // ... and 32-brick passes, all three axes at once, over the full bricks
  if (lane == 0) { s_n[warp] = 0; s_in[warp] = 0; }
  __syncwarp();
  // write the first `count` staged quads behind one reservation and empty the staging area
  auto flush = [&](int count) {
    if (count > 0) {
#if MB_BULK_STORE
      // the staged batch leaves shared memory as ONE bulk asynchronous store (TMA engine, non-tensor form): lane 0 reserves,
      // makes the lanes' generic-proxy writes visible to the async proxy, issues the copy and waits until the source has
      // been read (the staging area is reused right away)
      __syncwarp();
      if (lane == 0) {
        const unsigned long long sbase = atomicAdd_system(quad_count, (unsigned long long)count);   // system scope: the counter may live in a peer GPU
        const long long room = (long long)cap - (long long)sbase;
        const int n = room <= 0 ? 0 : (room < count ? (int)room : count);
        if (n > 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(reinterpret_cast<uint4*>(quads) + sbase), "r"(smem_u32(&s_q[warp][0])), "r"(n * 16) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
      }
#else
      unsigned long long sbase = 0;
      if (lane == 0) sbase = atomicAdd_system(quad_count, (unsigned long long)count);   // system scope: the counter may live in a peer GPU
      sbase = __shfl_sync(0xffffffffu, sbase, 0);
      for (int i = lane; i < count; i += 32)
        if ((int64_t)sbase + i < cap) reinterpret_cast<uint4*>(quads)[sbase + i] = s_q[warp][i];
#endif
    }
    __syncwarp();
    if (lane == 0) s_n[warp] = 0;
    __syncwarp();
  };
  // Phase 2: ONE QUAD PER LANE PER ITERATION of a converged loop (take_quad).  A lane whose image is used up takes the next
  // queued one (indices handed out with a ballot: the loop is converged, no atomics).  The loop stops as soon as a lane would
  // idle -- the queue has run dry -- and the lanes KEEP the images they are working on across passes: phase 1 of the next pass
  // refills the queue, so the merge runs with (nearly) all lanes busy whatever the quad counts of the images; only the last
  // call drains.  Slots of the staging area come from the same kind of ballot.
  const unsigned lt = (1u << lane) - 1u;
  uint64_t img = 0ull, meta = 0ull;
  auto merge = [&](int ni, bool drain) {
    int n_staged = min(s_n[warp], MQ_CAP);     // phase 1 stages quads itself only when the image queue overflows
    int next = 0;
    for (;;) {
      const unsigned need = __ballot_sync(0xffffffffu, img == 0ull);
      if (need != 0u && next < ni) {
        const int i = next + __popc(need & lt);
        if (img == 0ull && i < ni) { const ulonglong2 e = s_img[warp][i]; img = e.x; meta = e.y; }
        next += __popc(need);
      }
      const unsigned has = __ballot_sync(0xffffffffu, img != 0ull);
      if (has == 0u || (!drain && has != 0xffffffffu && next >= ni)) break;
      if (n_staged + 32 > MQ_CAP) { flush(n_staged); n_staged = 0; }
      if (img != 0ull) s_q[warp][n_staged + __popc(has & lt)] = take_quad(img, meta);
      n_staged += __popc(has);
    }
    __syncwarp();
    if (lane == 0) { s_in[warp] = 0; s_n[warp] = n_staged; }
    __syncwarp();
    if (n_staged >= MQ_FLUSH) flush(n_staged);
    __syncwarp();
  };
  for (int64_t grp = (int64_t)blockIdx.x * MB_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * MB_WARPS) {
    const bool full_pass = grp >= g_first;
    const int axis = full_pass ? 3 : (int)(grp % 3);
    const int64_t item = (full_pass ? grp - g_first : grp / 3) * 32 + lane;
    bool valid = item < (int64_t)(full_pass ? n_full : n_work);
    const uint64_t key = valid ? work[full_pass ? work_cap - 1 - item : item] : 0ull;
    // key lists (dirty re-mesh) are sharded over the ranks by a hash of the key: the list order is scheduling-dependent and
    // differs between the replicas, the key set does not
    if (shard_world > 1 && (int)(((key * 0x9E3779B97F4A7C15ull) >> 40) % (unsigned)shard_world) != shard_rank) valid = false;
    const int b = (int)(key & 4095);
    int cx, cy, cz;
    chunk_coords(v, (int64_t)(key >> 12), cx, cy, cz);
    const int bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);
    uint32_t slot = 0;
    const int st = valid ? brick_state(v, bx, by, bz, slot) : 0;
    uint64_t s[8];
    if (st == 2) load_slices(v, slot, s);
    else {
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = 0ull;    // (unused: a full brick's images come from its neighbours' planes alone)
    }
    if (axis == 0 || full_pass) queue_axis_images<0>(v, st, s, bx, by, bz, s_img[warp], &s_in[warp], s_q[warp], &s_n[warp], quads, cap, quad_count);
    if (axis == 1 || full_pass) queue_axis_images<1>(v, st, s, bx, by, bz, s_img[warp], &s_in[warp], s_q[warp], &s_n[warp], quads, cap, quad_count);
    if (axis == 2 || full_pass) queue_axis_images<2>(v, st, s, bx, by, bz, s_img[warp], &s_in[warp], s_q[warp], &s_n[warp], quads, cap, quad_count);
    __syncwarp();
    merge(min(s_in[warp], MI_CAP), /*drain=*/false);
  }
  merge(0, /*drain=*/true);      // the images the lanes still hold
  __syncwarp();
  flush(min(s_n[warp], MQ_CAP));
}

// ---- pass C: brick level -------------------------------------------------------------------------------------------------------
// One warp per (chunk, axis).  Chunks without bricks, without full bricks, or wholly full with six wholly full neighbours (the bulk of a
// solid's interior; one bit per chunk each) are dismissed first.  Otherwise the lanes form the chunk's 4096-bit images
// "full brick whose neighbour on the minus / plus side is absent" in the native word layout (two words per lane, the neighbour's
// occupancy moved onto the brick's bit: in-chunk words from shared memory, the facing words of the adjacent chunk only where a
// full brick touches the border) and leave them in shared memory; then lane = (side, brick layer) gathers its 16 rows x 16 bits
// and merges them greedily.  Chunks come from all of the volume (c % world == rank) or from a list (re-mesh; sharded over the
// ranks by a hash of the chunk index).  Streaming integer work: 1 KB of {occupancy, full} words per chunk that gets that far.
#define CF_WARPS 4
#define CF_STAGE 96
__global__ void __launch_bounds__(CF_WARPS * 32) mesh_chunk_faces_kernel(DVolume v, const uint32_t* __restrict__ chunk_list, const uint32_t* __restrict__ chunk_count_ptr,
                                                                           int shard_rank, int shard_world, MesoQuad* quads, int64_t cap, unsigned long long* quad_count) {
  __shared__ uint64_t s_o[CF_WARPS][64];
  __shared__ uint64_t s_e[CF_WARPS][2][64];
  __shared__ uint16_t s_rows[CF_WARPS][32][16 + 2];   // +2: lanes start in different banks
  __shared__ uint4 s_q[CF_WARPS][CF_STAGE];
  __shared__ int s_n[CF_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_chunks = (int64_t)*chunk_count_ptr;
  if (lane == 0) s_n[warp] = 0;
  __syncwarp();
  for (int64_t item = (int64_t)blockIdx.x * CF_WARPS + warp; item < n_chunks * 3; item += (int64_t)gridDim.x * CF_WARPS) {
    const int64_t c = (int64_t)chunk_list[item / 3];
    const int axis = (int)(item % 3);
    if (shard_world > 1 && !chunk_may_have_faces(v, c, true, shard_rank, shard_world)) continue;
    int cx, cy, cz;
    chunk_coords(v, c, cx, cy, cz);
    const ulonglong2 P0 = __ldg(&v.of[c * 64 + lane]), P1 = __ldg(&v.of[c * 64 + lane + 32]);
    if (!__any_sync(0xffffffffu, (P0.y | P1.y) != 0ull)) continue;     // no full brick
    __syncwarp();
    s_o[warp][lane] = P0.x; s_o[warp][lane + 32] = P1.x;
    __syncwarp();
    {
      uint64_t any = 0;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int w = lane + 32 * k, z = w >> 2, yq = w & 3;
        const uint64_t F = k ? P1.y : P0.y, me = k ? P1.x : P0.x;
        uint64_t em = 0, ep = 0;
        if (F) {
          uint64_t om, op;
          if (axis == 0) {
            const uint64_t X0 = 0x0001000100010001ull, X15 = 0x8000800080008000ull;
            om = (me << 1) & ~X0; op = (me >> 1) & ~X15;
            if (F & X0) om |= (of_word(v, cx - 1, cy, cz, w).x >> 15) & X0;
            if (F & X15) op |= (of_word(v, cx + 1, cy, cz, w).x << 15) & X15;
          } else if (axis == 1) {
            om = (me << 16) | ((yq > 0 ? s_o[warp][w - 1] : of_word(v, cx, cy - 1, cz, z * 4 + 3).x) >> 48);
            op = (me >> 16) | ((yq < 3 ? s_o[warp][w + 1] : of_word(v, cx, cy + 1, cz, z * 4 + 0).x) << 48);
          } else {
            om = z > 0 ? s_o[warp][w - 4] : of_word(v, cx, cy, cz - 1, 15 * 4 + yq).x;
            op = z < 15 ? s_o[warp][w + 4] : of_word(v, cx, cy, cz + 1, 0 * 4 + yq).x;
          }
          em = F & ~om; ep = F & ~op;
        }
        s_e[warp][0][w] = em; s_e[warp][1][w] = ep;
        any |= em | ep;
      }
      if (!__any_sync(0xffffffffu, any != 0ull)) continue;     // warp-uniform: no full brick of this chunk faces an absent one along this axis
      __syncwarp();
      const int side = lane >> 4, L = lane & 15;
      const uint64_t* e = s_e[warp][side];
      uint16_t* r = s_rows[warp][lane];
      uint32_t nonzero = 0;
      for (int vv = 0; vv < 16; vv++) {
        uint32_t row;
        if (axis == 2) row = (uint32_t)(e[L * 4 + (vv >> 2)] >> (16 * (vv & 3))) & 0xFFFFu;            // layer z: row y, bit x
        else if (axis == 1) row = (uint32_t)(e[vv * 4 + (L >> 2)] >> (16 * (L & 3))) & 0xFFFFu;        // layer y: row z, bit x
        else {                                                                                         // layer x: row z, bit y
          row = 0;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint64_t t = (e[vv * 4 + q] >> L) & 0x0001000100010001ull;
            row |= (uint32_t)((t & 1ull) | ((t >> 15) & 2ull) | ((t >> 30) & 4ull) | ((t >> 45) & 8ull)) << (4 * q);
          }
        }
        r[vv] = (uint16_t)row; nonzero |= row;
      }
      if (nonzero) {
        const int dir = 2 * axis + side;
        const int lv = L * 8 + (side ? 7 : 0);     // the voxel layer that carries the face
        for (int vv = 0; vv < 16; vv++) {
          uint32_t row = r[vv];
          while (row) {
            const int u0 = __ffs(row) - 1;
            const int w = __ffs(~(row >> u0)) - 1;
            const uint32_t m = ((1u << w) - 1u) << u0;
            int h = 1;
            while (vv + h < 16 && ((uint32_t)r[vv + h] & m) == m) { r[vv + h] = (uint16_t)(r[vv + h] & ~m); h++; }
            row &= ~m;
            int x, y, z;
            if (axis == 0) { x = lv; y = u0 * 8; z = vv * 8; } else if (axis == 1) { x = u0 * 8; y = lv; z = vv * 8; } else { x = u0 * 8; y = vv * 8; z = lv; }
            const uint4 q = make_uint4((uint32_t)(cx * 128 + x) | ((uint32_t)(cy * 128 + y) << 16),
                                       (uint32_t)(cz * 128 + z) | ((uint32_t)dir << 16) | (1u << 19) | ((uint32_t)(w * 8) << 24), (uint32_t)(h * 8), 0u);
            const int idx = atomicAdd(&s_n[warp], 1);
            if (idx < CF_STAGE) s_q[warp][idx] = q;
            else {
              const unsigned long long g = atomicAdd_system(quad_count, 1ull);
              if ((int64_t)g < cap) reinterpret_cast<uint4*>(quads)[g] = q;
            }
          }
        }
      }
      __syncwarp();
      const int count = min(s_n[warp], CF_STAGE);
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
    }
  }
}

// chunks that hold a listed brick, de-duplicated through a bit per chunk (mark: zero on entry, zero again after clear_chunk_marks)
__global__ void __launch_bounds__(256) mark_chunks_kernel(const uint64_t* __restrict__ keys, uint32_t n_keys, uint32_t* mark, uint32_t* chunk_list, uint32_t* chunk_count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const uint32_t c = (uint32_t)(keys[i] >> 12);
  if (i > 0 && (uint32_t)(keys[i - 1] >> 12) == c) return;   // cheap filter: runs of keys of one chunk
  const uint32_t old = atomicOr(&mark[c >> 5], 1u << (c & 31));
  if (old & (1u << (c & 31))) return;
  chunk_list[atomicAdd(chunk_count, 1u)] = c;
}
__global__ void clear_chunk_marks_kernel(const uint32_t* __restrict__ chunk_list, const uint32_t* __restrict__ chunk_count, uint32_t* mark) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *chunk_count) atomicAnd(&mark[chunk_list[i] >> 5], ~(1u << (chunk_list[i] & 31)));
}

void launch_mesh(const LaunchCtx& lc, const DVolume& v, int rank, int world, const MeshScratch& ms,
                 MesoQuad* d_quads, int64_t cap, unsigned long long* d_quad_count, bool reset_count) {
  cudaMemsetAsync(ms.work_count, 0, 3 * sizeof(uint32_t), lc.stream);   // work_count, chunk_count and full_count are adjacent words
  if (reset_count) cudaMemsetAsync(d_quad_count, 0, sizeof(unsigned long long), lc.stream);
  const int64_t mine = (v.nchunks - rank + world - 1) / world;     // chunks rank, rank + world, ...: one warp each
  mesh_worklist_kernel<<<(unsigned)((mine + 7) / 8), 256, 0, lc.stream>>>(v, rank, world, ms.work, ms.work_cap, ms.work_count, ms.full_count, ms.chunk_list, ms.chunk_count);
  mesh_bricks_kernel<<<lc.sm_count * 12, MB_THREADS, 0, lc.stream>>>(v, ms.work, ms.work_cap, ms.work_count, ms.full_count, 0u, d_quads, cap, d_quad_count, 0, 1);
  const int64_t groups = (v.nchunks * 3 + CF_WARPS - 1) / CF_WARPS;
  mesh_chunk_faces_kernel<<<(unsigned)min((int64_t)lc.sm_count * 16, groups), CF_WARPS * 32, 0, lc.stream>>>(v, ms.chunk_list, ms.chunk_count, 0, 1, d_quads, cap, d_quad_count);
  (*lc.launches) += 3;
}

void launch_mesh_list(const LaunchCtx& lc, const DVolume& v, const uint64_t* d_keys, uint32_t n_keys, MesoQuad* d_quads,
                      int64_t cap, unsigned long long* d_quad_count, const MeshScratch& ms, int rank, int world) {
  cudaMemsetAsync(d_quad_count, 0, sizeof(unsigned long long), lc.stream);
  if (n_keys == 0) return;
  const int64_t passes = (((int64_t)n_keys + 31) / 32) * 3;
  const unsigned grid = (unsigned)min((int64_t)lc.sm_count * 12, (passes + MB_WARPS - 1) / MB_WARPS);
  mesh_bricks_kernel<<<grid, MB_THREADS, 0, lc.stream>>>(v, d_keys, 0, nullptr, nullptr, n_keys, d_quads, cap, d_quad_count, rank, world);
  // the brick-level quads of every chunk that holds a listed brick
  cudaMemsetAsync(ms.chunk_count, 0, sizeof(uint32_t), lc.stream);
  mark_chunks_kernel<<<(n_keys + 255) / 256, 256, 0, lc.stream>>>(d_keys, n_keys, ms.chunk_mark, ms.chunk_list, ms.chunk_count);
  const int64_t bound = min((int64_t)n_keys, v.nchunks);      // the list cannot be longer than either
  mesh_chunk_faces_kernel<<<(unsigned)min((int64_t)lc.sm_count * 16, (bound * 3 + CF_WARPS - 1) / CF_WARPS), CF_WARPS * 32, 0, lc.stream>>>(
      v, ms.chunk_list, ms.chunk_count, rank, world, d_quads, cap, d_quad_count);
  clear_chunk_marks_kernel<<<(unsigned)((bound + 255) / 256), 256, 0, lc.stream>>>(ms.chunk_list, ms.chunk_count, ms.chunk_mark);
  (*lc.launches) += 4;
}
