// k_raymarch.cu -- K4: per-pixel 3D-DDA, primary + one hard-shadow ray.
//
// Replaces the reference's per-pixel visibility: the 1 M-instance triangle-fan draw + reverse-Z depth test
// (Samples/SimpleVoxel.cpp:146-192 VS, :220-224 FS, dispatch :352-398) -- see DESIGN.md for the equivalence.
//
// The DDA is STATELESS (DESIGN.md "DDA"): the crossing time of integer voxel plane p on axis a is always
//     t_a(p) = (float(p) - o_a) * inv_a          one FSUB + one FMUL, never fused (-fmad=false)
// and crossings are consumed in the total order (t, axis).  Skipping an empty aligned cell (512^3 region, 128^3 chunk,
// 32^3 cell, 8^3 brick) consumes the smallest of the cell's three exit keys and re-derives the other two coordinates
// from the same keys only when the walk has to look finer (advance_axis), so the hierarchical walk visits exactly the
// voxels the oracle's flat walk would: hit voxel, face and t are bit-identical to oracle/orc_raymarch.c.
//
// Mapping: CTA = one 32x8 screen tile (8 warps), warp = 8x4 pixels (a 128 B-aligned 4-line store of 16 B records).
// Region / chunk any-bits are staged in shared memory; 32^3-cell masks, {occ,full} word pairs (16 B loads) and brick
// slices are cached in registers behind tags.  Shadow rays are compacted across the CTA (ballot + prefix sum through
// shared memory) so the second trace runs in dense warps.  Multi-GPU: tile t belongs to rank t % world.
#include "meso_internal.cuh"

#define F_INF __int_as_float(0x7F800000)

struct Ray {
  float o[3], d[3], inv[3];
  int step[3];  // +1 / -1 / 0 (inactive: |d| < 1e-20, never crosses a plane)
  int sgn[3];   // 0 or -1: coordinates are kept mirrored (c ^ sgn) so every step is "+"
  int lim[3];   // mirrored coordinate at which the ray has left the grid on that axis
};

__device__ __forceinline__ float plane_t(const Ray& r, int a, int plane) {
  return __fmul_rn(__fsub_rn((float)plane, r.o[a]), r.inv[a]);
}
__device__ __forceinline__ bool key_less(float t1, int a1, float t2, int a2) { return t1 < t2 || (t1 == t2 && a1 < a2); }

__device__ __forceinline__ void ray_init(Ray& r, const float o[3], const float d[3], const DVolume& v) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    r.o[i] = o[i]; r.d[i] = d[i];
    if (fabsf(d[i]) >= 1e-20f) { r.inv[i] = __fdiv_rn(1.0f, d[i]); r.step[i] = d[i] > 0.0f ? 1 : -1; }
    else { r.inv[i] = 0.0f; r.step[i] = 0; }
    r.sgn[i] = r.step[i] < 0 ? -1 : 0;
    r.lim[i] = r.step[i] > 0 ? v.nvox[i] : (r.step[i] < 0 ? 0 : 0x7FFFFFFF);
  }
}

__device__ __forceinline__ int clamp_floor_to_int(float x) {
  float f = floorf(x);
  f = fminf(fmaxf(f, -1.0e9f), 1.0e9f);
  return (int)f;
}

// True coordinate on axis b after consuming every crossing with key < (ts, as), starting from the (older) true cell
// coordinate cur.
__device__ __forceinline__ int advance_axis(const Ray& r, int b, int cur, float ts, int as) {
  const int st = r.step[b];
  if (st == 0) return cur;
  const float pos = __fadd_rn(r.o[b], __fmul_rn(r.d[b], ts));
  const float fl = floorf(pos);
  int e = (int)fminf(fmaxf(fl, -1.0e9f), 1.0e9f);
  // Fast path (DESIGN.md "advance_axis shortcut"): when the estimated position is farther than eps from every voxel
  // plane, every plane behind it has a computed crossing time < ts and every plane ahead > ts (fp32 error of
  // t_b(p) <= 1.8e-7 |t|, of pos <= 6e-8 (|ts| + |pos|)), so the exact answer is floor(pos).  Otherwise decide with
  // the keys themselves.  Same result either way; the oracle only has the slow path.
  const float fr = __fsub_rn(pos, fl);
  const float eps = __fadd_rn(__fmul_rn(__fadd_rn(fabsf(ts), fabsf(pos)), 5e-7f), 1e-5f);
  if (fr > eps && fr < __fsub_rn(1.0f, eps)) return e;
  if (st > 0) e = max(e, cur); else e = min(e, cur);
  for (;;) {
    const int pa = st > 0 ? e + 1 : e;
    if (key_less(plane_t(r, b, pa), b, ts, as)) e += st; else break;
  }
  for (;;) {
    if (e == cur) break;
    const int pb = st > 0 ? e : e + 1;
    if (!key_less(plane_t(r, b, pb), b, ts, as)) e -= st; else break;
  }
  return e;
}

struct Scene {
  const DVolume* v;
  const uint32_t* s_any;     // shared: bit per chunk
  const uint32_t* s_region;  // shared: bit per 512^3 region
  uint8_t* touch_chunk;
  uint8_t* touch_brick;
};

__device__ __forceinline__ bool inside(const DVolume& v, const int c[3]) {
  return (unsigned)c[0] < (unsigned)v.nvox[0] && (unsigned)c[1] < (unsigned)v.nvox[1] && (unsigned)c[2] < (unsigned)v.nvox[2];
}
__device__ __forceinline__ bool gone(const DVolume& v, const Ray& r, const int c[3]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (r.step[i] > 0) { if (c[i] >= v.nvox[i]) return true; }
    else if (r.step[i] < 0) { if (c[i] < 0) return true; }
    else if (c[i] < 0 || c[i] >= v.nvox[i]) return true;
  }
  return false;
}

struct Trace { bool hit; int c[3]; int axis; float t; unsigned steps; };

// levels: 4 = region (512^3), 3 = chunk (128^3), 2 = cell (32^3), 1 = brick (8^3), 0 = voxel
template <bool STATS>
__device__ __forceinline__ void trace(const Scene& s, const Ray& r, const int c0[3], Trace& tr) {
  const DVolume& v = *s.v;
  int c[3] = {c0[0], c0[1], c0[2]};
  int la = -1; float lt = 0.0f; unsigned steps = 0;
  bool hit = false, alive = true;
  if (!inside(v, c)) {
    if (gone(v, r, c)) alive = false;
    else {
      int a = -1; float ta = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const bool before = (r.step[i] > 0 && c[i] < 0) || (r.step[i] < 0 && c[i] >= v.nvox[i]);
        if (before) {
          const float ti = plane_t(r, i, r.step[i] > 0 ? 0 : v.nvox[i]);
          if (a < 0 || key_less(ta, a, ti, i)) { a = i; ta = ti; }
        }
      }
#pragma unroll
      for (int b = 0; b < 3; b++) {
        if (b == a) c[b] = r.step[b] > 0 ? 0 : v.nvox[b] - 1;
        else c[b] = advance_axis(r, b, c[b], ta, a);
      }
      la = a; lt = ta; steps = 1;
      alive = inside(v, c);
    }
  }
  // mirrored coordinates: cs = c ^ sgn.  Above voxel level only the last stepped axis is exact; the other two hold an
  // older true coordinate inside the same cell of the level that was stepped, and are made exact before looking finer.
  int cs[3];
#pragma unroll
  for (int i = 0; i < 3; i++) cs[i] = c[i] ^ r.sgn[i];
  int gran = 0;  // shift of the level last stepped: non-stepped axes are only valid at this granularity
  int need = 4;
  int ci = 0, wtag = -1, ztag = -1;
  uint32_t slot = 0;
  unsigned long long cellmask = 0, wocc = 0, wfull = 0, slice = 0;

  while (alive) {
    int sh;
    {
#pragma unroll
      for (int i = 0; i < 3; i++) c[i] = cs[i] ^ r.sgn[i];
      bool go = true;  // keep looking finer
      sh = 0;
      if (need >= 4) {
        const int ri = (c[0] >> 9) + v.rdims[0] * ((c[1] >> 9) + v.rdims[1] * (c[2] >> 9));
        if (!((s.s_region[ri >> 5] >> (ri & 31)) & 1u)) { sh = 9; go = false; }
      }
#define MESO_SYNC_IF_COARSER(S)                                                            \
      if (go && gran > (S)) { /* looking finer than the level that was stepped: make the other two axes exact */ \
        _Pragma("unroll") for (int b = 0; b < 3; b++) if (b != la) c[b] = advance_axis(r, b, c[b], lt, la);   \
        _Pragma("unroll") for (int i = 0; i < 3; i++) cs[i] = c[i] ^ r.sgn[i];                                 \
        gran = 0;                                                                            \
      }
      MESO_SYNC_IF_COARSER(7)
      if (go && need >= 3) {
        ci = (c[0] >> 7) + v.dims[0] * ((c[1] >> 7) + v.dims[1] * (c[2] >> 7));
        if (!((s.s_any[ci >> 5] >> (ci & 31)) & 1u)) { sh = 7; go = false; }
        else {
          if (STATS) s.touch_chunk[ci] = 1;
          cellmask = __ldg(&v.cells[ci]);
          wtag = -1;
        }
      }
      MESO_SYNC_IF_COARSER(5)
      if (go && need >= 2) {
        const int e = ((c[0] >> 5) & 3) + 4 * ((c[1] >> 5) & 3) + 16 * ((c[2] >> 5) & 3);
        if (!((cellmask >> e) & 1ull)) { sh = 5; go = false; }
      }
      MESO_SYNC_IF_COARSER(3)
      if (go && need >= 1) {
        const int bx = (c[0] >> 3) & 15, by = (c[1] >> 3) & 15, bz = (c[2] >> 3) & 15;
        const int w = bz * 4 + (by >> 2);
        if (w != wtag) {
          const ulonglong2 p = __ldg(&v.of[(size_t)ci * 64 + w]);
          wocc = p.x; wfull = p.y; wtag = w;
        }
        const int bit = bx + 16 * (by & 3);
        if (!((wocc >> bit) & 1ull)) { sh = 3; go = false; }
        else {
          if ((wfull >> bit) & 1ull) {
            if (gran > 0) {
#pragma unroll
              for (int b = 0; b < 3; b++) if (b != la) c[b] = advance_axis(r, b, c[b], lt, la);
            }
            hit = true; break;
          }
          slot = __ldg(&v.bptr[(size_t)ci * MESO_BLOCKS + (bx + 16 * by + 256 * bz)]);
          if (STATS) s.touch_brick[slot] = 1;
          ztag = -1;
        }
      }
      MESO_SYNC_IF_COARSER(0)
      if (go) {
        const int z = c[2] & 7;
        if (z != ztag) { slice = __ldg(&v.pool[(size_t)slot * 8 + z]); ztag = z; }
        if ((slice >> ((c[0] & 7) + 8 * (c[1] & 7))) & 1ull) { hit = true; break; }
      }
    }
    // ---- one step at level sh: consume the smallest of the three pending keys (ties: lower axis first) ----
    const int mask = (1 << sh) - 1;
    int nx[3]; float tn[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      nx[i] = (cs[i] | mask) + 1;                         // mirrored coordinate after crossing
      const int pl = (nx[i] ^ r.sgn[i]) - r.sgn[i];       // true plane index
      const float t = __fmul_rn(__fsub_rn((float)pl, r.o[i]), r.inv[i]);
      tn[i] = r.step[i] != 0 ? t : F_INF;
    }
    int a = 0; float ta = tn[0];
    if (tn[1] < ta) { a = 1; ta = tn[1]; }
    if (tn[2] < ta) { a = 2; ta = tn[2]; }
    if (!(ta < F_INF)) break;  // zero direction
    int cross = 0, lim = 0, ca = 0;
#pragma unroll
    for (int b = 0; b < 3; b++) if (b == a) { cross = cs[b] ^ nx[b]; cs[b] = nx[b]; ca = nx[b]; lim = r.lim[b]; }
    la = a; lt = ta; steps++;
    gran = sh;
    const unsigned ucross = (unsigned)cross;
    need = (ucross >> 9) ? 4 : ((ucross >> 7) ? 3 : ((ucross >> 5) ? 2 : ((ucross >> 3) ? 1 : 0)));
    if (need >= 3) alive = ca < lim;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) tr.c[i] = hit ? c[i] : (cs[i] ^ r.sgn[i]);
  tr.hit = hit; tr.axis = la; tr.t = lt; tr.steps = steps;
}

__device__ __forceinline__ uint32_t to_un8(float x) { return (uint32_t)__fadd_rn(__fmul_rn(x, 255.0f), 0.5f); }

struct ShadowJob { float p[3]; int c[3]; int owner; };

template <bool STATS>
__global__ void __launch_bounds__(256) raymarch_kernel(DVolume v, MesoRaySetup rs, int width, int height, uint32_t flags,
                                                       int rank, int world, int layout, int tiles_x, int n_tiles,
                                                       MesoHitRecord* __restrict__ out, RayStatsDev* stats,
                                                       uint8_t* touch_chunk, uint8_t* touch_brick) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_any = s_dyn;
  uint32_t* s_region = s_dyn + v.chunk_words;
  __shared__ ShadowJob s_jobs[256];
  __shared__ uint8_t s_shadow[256];
  __shared__ int s_warp_cnt[8];
  for (int i = threadIdx.x; i < v.chunk_words; i += blockDim.x) s_any[i] = v.chunk_any[i];
  for (int i = threadIdx.x; i < v.region_words; i += blockDim.x) s_region[i] = v.region_any[i];
  __syncthreads();
  const int local_tile = blockIdx.x;
  const int tile = local_tile * world + rank;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = (warp & 3) * 8 + (lane & 7), ty = (warp >> 2) * 4 + (lane >> 3);
  const int px = (tile % tiles_x) * MESO_TILE_W + tx;
  const int py = (tile / tiles_x) * MESO_TILE_H + ty;
  const bool valid = tile < n_tiles && px < width && py < height;

  Scene sc; sc.v = &v; sc.s_any = s_any; sc.s_region = s_region; sc.touch_chunk = touch_chunk; sc.touch_brick = touch_brick;

  // ---- phase 1: primary ray ----
  Trace tr; tr.hit = false; tr.axis = -1; tr.t = 0.0f; tr.steps = 0; tr.c[0] = tr.c[1] = tr.c[2] = 0;
  float p[3] = {0.f, 0.f, 0.f};
  int face = 7, shadow = 0;
  bool want_shadow = false;
  unsigned steps = 0, n_shadow = 0;
  if (valid) {
    const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)px, 0.5f), rs.two_over_w), 1.0f);
    const float fy = __fsub_rn(1.0f, __fmul_rn(__fadd_rn((float)py, 0.5f), rs.two_over_h));
    float d[3];
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __fadd_rn(__fadd_rn(__fmul_rn(fx, rs.U[i]), __fmul_rn(fy, rs.V[i])), rs.F[i]);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __fdiv_rn(d[i], len);
    Ray r; ray_init(r, rs.o, d, v);
    int c0[3];
#pragma unroll
    for (int i = 0; i < 3; i++) c0[i] = clamp_floor_to_int(rs.o[i]);
    trace<STATS>(sc, r, c0, tr);
    steps = tr.steps;
    if (tr.hit) {
      face = 6;
#pragma unroll
      for (int i = 0; i < 3; i++) p[i] = __fadd_rn(r.o[i], __fmul_rn(r.d[i], tr.t));
      if (tr.axis >= 0) {
        const int ax = tr.axis;
        int st_ax = 0; float l_ax = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; i++) if (i == ax) { st_ax = r.step[i]; l_ax = rs.L[i]; p[i] = (float)(r.step[i] > 0 ? tr.c[i] : tr.c[i] + 1); }
        face = ax * 2 + (st_ax > 0 ? 0 : 1);
        if (flags & MESO_FLAG_SHADOW) {
          const bool facing = st_ax > 0 ? (l_ax < 0.0f) : (l_ax > 0.0f);
          if (!facing) shadow = 1; else want_shadow = true;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; i++) p[i] = r.o[i];
      }
    }
  }

  // ---- phase 2: shadow rays, compacted across the CTA ----
  if (flags & MESO_FLAG_SHADOW) {
    const unsigned bal = __ballot_sync(0xffffffffu, want_shadow);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const int n = s_warp_cnt[w]; if (w < warp) base += n; total += n; }
    if (want_shadow) {
      const int idx = base + __popc(bal & ((1u << lane) - 1u));
      ShadowJob& j = s_jobs[idx];
      const int ax = tr.axis;
#pragma unroll
      for (int i = 0; i < 3; i++) { j.p[i] = p[i]; j.c[i] = tr.c[i] - ((i == ax) ? ((face & 1) ? -1 : 1) : 0); }
      j.owner = threadIdx.x;
    }
    __syncthreads();
    if ((int)threadIdx.x < total) {
      const ShadowJob j = s_jobs[threadIdx.x];
      Ray sr; ray_init(sr, j.p, rs.L, v);
      Trace st;
      trace<STATS>(sc, sr, j.c, st);
      s_shadow[j.owner] = st.hit ? 1 : 0;
      steps += st.steps; n_shadow = 1;
    }
    __syncthreads();
    if (want_shadow) shadow = s_shadow[threadIdx.x];
  }

  // ---- phase 3: shade + store ----
  if (valid) {
    MesoHitRecord rec;
    if (!tr.hit) {
      rec.w0 = 0xFFFFFFFFu; rec.w1 = 0x0007FFFFu; rec.t = F_INF; rec.rgba = 0xFF000000u;
    } else {
      const float shade = shadow ? 0.5f : 1.0f;
      uint32_t ch[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float local = __fsub_rn(__fmul_rn(p[i], 0.125f), (float)(tr.c[i] >> 3));
        local = fminf(fmaxf(local, 0.0f), 1.0f);
        const float col = __fadd_rn(__fmul_rn(__fsub_rn(local, 0.5f), 0.5f), 0.5f);  // SimpleVoxel.cpp:222
        ch[i] = to_un8(__fmul_rn(col, shade));
      }
      rec.w0 = (uint32_t)tr.c[0] | ((uint32_t)tr.c[1] << 16);
      rec.w1 = (uint32_t)tr.c[2] | ((uint32_t)face << 16) | ((uint32_t)shadow << 19) | (1u << 20);
      rec.t = tr.t;
      rec.rgba = ch[0] | (ch[1] << 8) | (ch[2] << 16);
    }
    size_t dst;
    if (layout == MESO_LAYOUT_FRAME) dst = (size_t)py * width + px;
    else dst = (size_t)local_tile * (MESO_TILE_W * MESO_TILE_H) + ty * MESO_TILE_W + tx;
    reinterpret_cast<uint4*>(out)[dst] = make_uint4(rec.w0, rec.w1, __float_as_uint(rec.t), rec.rgba);
  }

  if (STATS) {
    unsigned long long v0 = valid ? 1 : 0, v1 = n_shadow, v2 = tr.hit ? 1 : 0, v3 = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
    }
    if (lane == 0) {
      atomicAdd(&stats->primary, v0); atomicAdd(&stats->shadow, v1);
      atomicAdd(&stats->hits, v2); atomicAdd(&stats->steps, v3);
    }
  }
}

__global__ void __launch_bounds__(256) compose_tiles_kernel(const uint4* __restrict__ tiles, int world, int width, int height,
                                                            int tiles_x, int n_tiles, int64_t tiles_per_rank, uint4* __restrict__ frame) {
  const int tile = blockIdx.x;
  if (tile >= n_tiles) return;
  const int rank = tile % world; const int64_t local = tile / world;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int px = (tile % tiles_x) * MESO_TILE_W + tx, py = (tile / tiles_x) * MESO_TILE_H + ty;
  if (px >= width || py >= height) return;
  frame[(size_t)py * width + px] = tiles[((size_t)rank * tiles_per_rank + local) * 256 + threadIdx.x];
}

void launch_raymarch(const LaunchCtx& lc, const DVolume& v, const MesoRaySetup& rs, int width, int height, uint32_t flags,
                     int rank, int world, int layout, MesoHitRecord* d_out, RayStatsDev* d_stats, uint8_t* d_touch_chunk,
                     uint8_t* d_touch_brick) {
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W, tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  const int n_tiles = tiles_x * tiles_y;
  const int local_tiles = (n_tiles - rank + world - 1) / world;
  if (local_tiles <= 0) return;
  const size_t smem = sizeof(uint32_t) * ((size_t)v.chunk_words + (size_t)v.region_words);
  if (d_stats)
    raymarch_kernel<true><<<local_tiles, 256, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, tiles_x, n_tiles,
                                                                 d_out, d_stats, d_touch_chunk, d_touch_brick);
  else
    raymarch_kernel<false><<<local_tiles, 256, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, tiles_x, n_tiles,
                                                                  d_out, nullptr, nullptr, nullptr);
  (*lc.launches)++;
}

void launch_compose_tiles(const LaunchCtx& lc, const MesoHitRecord* d_tiles, int world, int width, int height, MesoHitRecord* d_frame) {
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W, tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  const int n_tiles = tiles_x * tiles_y;
  const int64_t tpr = (n_tiles + world - 1) / world;
  compose_tiles_kernel<<<n_tiles, 256, 0, lc.stream>>>(reinterpret_cast<const uint4*>(d_tiles), world, width, height, tiles_x,
                                                       n_tiles, tpr, reinterpret_cast<uint4*>(d_frame));
  (*lc.launches)++;
}
