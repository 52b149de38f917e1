// k_raymarch.cu -- K4: per-pixel 3D-DDA, primary + one hard-shadow ray.
//
// Replaces the reference's per-pixel visibility: the 1 M-instance triangle-fan draw + reverse-Z depth test
// (Samples/SimpleVoxel.cpp:146-192 VS, :220-224 FS, dispatch :352-398) -- see DESIGN.md for the equivalence.
//
// The DDA is STATELESS (DESIGN.md "DDA"): the crossing time of integer voxel plane p on axis a is always
//     t_a(p) = (float(p) - o_a) * inv_a          one FSUB + one FMUL, never fused (-fmad=false)
// and crossings are consumed in the total order (t, axis).  Skipping an empty box -- an aligned 8^3 brick, or the cube
// of 32^3 cells the distance field DVolume.df certifies empty around the current cell -- consumes the smallest of the
// box's three exit keys and re-derives the other two coordinates from the same keys (advance_axis), so the walk visits
// exactly the voxels the oracle's flat walk would: hit voxel, face and t are bit-identical to oracle/orc_raymarch.c.
// Any empty box is a legal skip; how big the boxes are only changes the number of steps, never the result.
//
// Execution model: CTA = one 32x8 screen tile (tile t belongs to rank t % world), warp = 8x4 pixels.  All primary rays
// of a warp start together at the eye and stay at similar levels of the walk; afterwards the lanes that hit a lit-facing
// face trace their shadow rays, starting together again.  Every level of the walk shares ONE step section per
// iteration: separate step code for the distance-field level was measured twice and costs 12 % more warp instructions.
// Distance-field bytes come through L1/L2 (2 MB at 4096^3); {occ,full} word pairs (16 B loads) and brick slices are
// cached in registers behind tags.
// (A persistent "idle lanes pull the next pixel" variant was measured and dropped: mixing rays of different phases in
// one warp cut SIMT efficiency from 20/32 to 8/32 active threads -- profiles/README.md.)
//
// All per-ray state is kept in named scalars (x/y/z members, SEL3 selects), never in indexable arrays: the compiler
// turns "if (i == a) v = arr[i]" chains into a dynamically indexed load, which would push the whole state to local memory.
#include "meso_internal.cuh"
#include <cstdlib>
#include <cstring>

#define F_INF __int_as_float(0x7F800000)
#define RM_THREADS 256
#define SEL3(a, X, Y, Z) ((a) == 0 ? (X) : ((a) == 1 ? (Y) : (Z)))

struct Ray {
  float ox, oy, oz, dx, dy, dz, ix, iy, iz;
  int sx, sy, sz;  // step: +1 / -1 / 0 (inactive: |d| < 1e-20, never crosses a plane); mirror sign = step >> 31
};

__device__ __forceinline__ float plane_t1(float o, float inv, int plane) { return __fmul_rn(__fsub_rn((float)plane, o), inv); }
__device__ __forceinline__ bool key_less(float t1, int a1, float t2, int a2) { return t1 < t2 || (t1 == t2 && a1 < a2); }

__device__ __forceinline__ void dir1(float d, float& inv, int& step) {
  if (fabsf(d) >= 1e-20f) { inv = __fdiv_rn(1.0f, d); step = d > 0.0f ? 1 : -1; }
  else { inv = 0.0f; step = 0; }
}
__device__ __forceinline__ void ray_dir(Ray& r, float dx, float dy, float dz) {
  r.dx = dx; r.dy = dy; r.dz = dz;
  dir1(dx, r.ix, r.sx); dir1(dy, r.iy, r.sy); dir1(dz, r.iz, r.sz);
}

__device__ __forceinline__ int clamp_floor_to_int(float x) {
  float f = floorf(x);
  f = fminf(fmaxf(f, -1.0e9f), 1.0e9f);
  return (int)f;
}

// True coordinate on axis b after consuming every crossing with key < (ts, as), starting from the (older) true cell
// coordinate cur.  (o, d, inv, st) are the ray's components on axis b.
__device__ __forceinline__ int advance_axis(float o, float d, float inv, int st, int b, int cur, float ts, int as) {
  if (st == 0) return cur;
  const float pos = __fadd_rn(o, __fmul_rn(d, ts));
  const float fl = floorf(pos);
  int e = (int)fminf(fmaxf(fl, -1.0e9f), 1.0e9f);
  // Fast path (DESIGN.md "advance_axis shortcut"): when the estimated position is farther than eps from every voxel
  // plane, every plane behind it has a computed crossing time < ts and every plane ahead > ts (fp32 error of
  // t_b(p) <= 1.8e-7 |t|, of pos <= 6e-8 (|ts| + |pos|)), so the exact answer is floor(pos).  Otherwise decide with
  // the keys themselves.  Same result either way; the oracle only has the slow path.  (One launch-wide bound for eps
  // instead of the per-call value saves six instructions and was 2 % slower: more lanes end up in the slow path.)
  const float fr = __fsub_rn(pos, fl);
  const float eps = __fadd_rn(__fmul_rn(__fadd_rn(fabsf(ts), fabsf(pos)), 5e-7f), 1e-5f);
  if (fr > eps && fr < __fsub_rn(1.0f, eps)) return e;
  if (st > 0) e = max(e, cur); else e = min(e, cur);
  for (;;) {
    const int pa = st > 0 ? e + 1 : e;
    if (key_less(plane_t1(o, inv, pa), b, ts, as)) e += st; else break;
  }
  for (;;) {
    if (e == cur) break;
    const int pb = st > 0 ? e : e + 1;
    if (!key_less(plane_t1(o, inv, pb), b, ts, as)) e -= st; else break;
  }
  return e;
}

struct Scene {
  const DVolume* v;
  uint8_t* touch_chunk;
  uint8_t* touch_brick;
  unsigned* lv;              // STATS only: per-thread steps per level
  CubeTables ct;             // CUBES only: per-octant forward cubes (meso_build_cubes)
};

// Walk state of one ray.  cs = c ^ (step >> 31): mirrored coordinates, so every aligned step is "(cs | mask) + 1".
// All three are exact between iterations: a step out of a box bigger than a voxel re-derives the two other axes from
// the consumed key right away (one sync site in the step section; syncing lazily, only when the walk looks finer, needs
// the same ~35 instructions at three places of the cascade, and a warp pays for every place any of its lanes visits).
// need: 3 = the 32^3 cell may have changed (look the distance field up), 2 = the brick changed, 1 = the 2^3 cell
// changed, 0 = same 2^3 cell.
struct Walk {
  int csx, csy, csz;
  int la; float lt;
  int need;
  int ci, wtag, ztag;
  uint32_t slot;
  unsigned long long wocc, wfull, slice, cm;
};

enum { W_CONTINUE = 0, W_HIT = 1, W_EXIT = 2 };

// make the two axes other than w.la exact (true coordinates cx,cy,cz in/out)
__device__ __forceinline__ void sync_axes(const Ray& r, const Walk& w, int& cx, int& cy, int& cz) {
  if (w.la != 0) cx = advance_axis(r.ox, r.dx, r.ix, r.sx, 0, cx, w.lt, w.la);
  if (w.la != 1) cy = advance_axis(r.oy, r.dy, r.iy, r.sy, 1, cy, w.lt, w.la);
  if (w.la != 2) cz = advance_axis(r.oz, r.dz, r.iz, r.sz, 2, cz, w.lt, w.la);
}

// Places the ray at its first cell inside the grid (entry from outside included).  Returns false if it never enters.
__device__ __forceinline__ bool walk_begin(const DVolume& v, const Ray& r, int cx, int cy, int cz, Walk& w, unsigned& steps) {
  const int nx = v.nvox[0], ny = v.nvox[1], nz = v.nvox[2];
  w.la = -1; w.lt = 0.0f;
  bool alive = true;
  const bool in = (unsigned)cx < (unsigned)nx && (unsigned)cy < (unsigned)ny && (unsigned)cz < (unsigned)nz;
  if (!in) {
    bool gone = false;
    gone |= r.sx > 0 ? cx >= nx : (r.sx < 0 ? cx < 0 : (cx < 0 || cx >= nx));
    gone |= r.sy > 0 ? cy >= ny : (r.sy < 0 ? cy < 0 : (cy < 0 || cy >= ny));
    gone |= r.sz > 0 ? cz >= nz : (r.sz < 0 ? cz < 0 : (cz < 0 || cz >= nz));
    if (gone) alive = false;
    else {
      int a = -1; float ta = 0.0f;
      if ((r.sx > 0 && cx < 0) || (r.sx < 0 && cx >= nx)) { const float ti = plane_t1(r.ox, r.ix, r.sx > 0 ? 0 : nx); if (a < 0 || key_less(ta, a, ti, 0)) { a = 0; ta = ti; } }
      if ((r.sy > 0 && cy < 0) || (r.sy < 0 && cy >= ny)) { const float ti = plane_t1(r.oy, r.iy, r.sy > 0 ? 0 : ny); if (a < 0 || key_less(ta, a, ti, 1)) { a = 1; ta = ti; } }
      if ((r.sz > 0 && cz < 0) || (r.sz < 0 && cz >= nz)) { const float ti = plane_t1(r.oz, r.iz, r.sz > 0 ? 0 : nz); if (a < 0 || key_less(ta, a, ti, 2)) { a = 2; ta = ti; } }
      cx = a == 0 ? (r.sx > 0 ? 0 : nx - 1) : advance_axis(r.ox, r.dx, r.ix, r.sx, 0, cx, ta, a);
      cy = a == 1 ? (r.sy > 0 ? 0 : ny - 1) : advance_axis(r.oy, r.dy, r.iy, r.sy, 1, cy, ta, a);
      cz = a == 2 ? (r.sz > 0 ? 0 : nz - 1) : advance_axis(r.oz, r.dz, r.iz, r.sz, 2, cz, ta, a);
      w.la = a; w.lt = ta; steps++;
      alive = (unsigned)cx < (unsigned)nx && (unsigned)cy < (unsigned)ny && (unsigned)cz < (unsigned)nz;
    }
  }
  w.csx = cx ^ (r.sx >> 31); w.csy = cy ^ (r.sy >> 31); w.csz = cz ^ (r.sz >> 31);
  w.need = 3; w.ci = -1; w.wtag = -1; w.ztag = -1; w.slot = 0;
  w.wocc = 0; w.wfull = 0; w.slice = 0; w.cm = 0;
  return alive;
}

// One iteration: classify the current cell from the coarsest level that changed down to the first empty level (or a
// solid voxel), then take one step at that level.  Levels: the distance field over 32^3 cells (a step leaves the whole
// empty cube of half-width df - 1 cells around the current cell, clamped to the grid), bricks (8^3), voxels.
// On W_HIT (cx,cy,cz) is the exact hit voxel.
// CUBES (opt-in, MESO_FLAG_CUBES; chosen by the step model in profiles/README.md, not yet measured on hardware): every
// level reads the edge of the largest empty cube that STARTS at the current cell / brick / 2^3 cell and extends towards
// the ray's octant (CubeTables) where the shipped walk reads a symmetric distance or an occupancy bit.  Same levels, same
// single step section, bigger boxes; any empty box is a legal skip, so the records cannot change.
template <bool STATS, bool CUBES = false>
__device__ __forceinline__ int walk_iter(const Scene& s, const Ray& r, Walk& w, int& cx, int& cy, int& cz, unsigned& steps) {
  const DVolume& v = *s.v;
  const int gx = r.sx >> 31, gy = r.sy >> 31, gz = r.sz >> 31;
  cx = w.csx ^ gx; cy = w.csy ^ gy; cz = w.csz ^ gz;
  bool go = true;  // keep looking finer
  int sh = 0;      // level of the step: 0 voxel, 1 2^3 cell, 3 brick, 5 distance-field cube
  int kdf = 1;     // cells to advance at that level (> 1 only for distance-field steps)
  if (CUBES) {
    const int oct = (gx & 1) | (gy & 2) | (gz & 4);
    if (w.need >= 3) {
      const int ex = cx >> 5, ey = cy >> 5, ez = cz >> 5;
      const int k = (int)__ldg(&s.ct.cell[(size_t)oct * (size_t)s.ct.ncells + (size_t)(ex + v.ddims[0] * (ey + v.ddims[1] * ez))]);
      if (k > 0) { sh = 5; kdf = k; go = false; }
      else {
        const int ci = (cx >> 7) + v.dims[0] * ((cy >> 7) + v.dims[1] * (cz >> 7));
        if (ci != w.ci) { w.ci = ci; w.wtag = -1; }
        if (STATS) {
          const int e = ((cx >> 5) & 3) + 4 * ((cy >> 5) & 3) + 16 * ((cz >> 5) & 3);
          if ((v.cells[ci] >> e) & 1ull) s.touch_chunk[ci] = 1;
        }
      }
    }
    if (go && w.need >= 2) {
      const int bx = (cx >> 3) & 15, by = (cy >> 3) & 15, bz = (cz >> 3) & 15;
      const int wi = bz * 4 + (by >> 2);
      if (wi != w.wtag) {
        const ulonglong2 p = __ldg(&v.of[(size_t)w.ci * 64 + wi]);
        w.wocc = p.x; w.wfull = p.y; w.wtag = wi;
      }
      const int bit = bx + 16 * (by & 3);
      if (!((w.wocc >> bit) & 1ull)) {
        sh = 3; go = false;
        kdf = 1 + (((int)__ldg(&s.ct.brick[(size_t)w.ci * MESO_BLOCKS + (bx + 16 * by + 256 * bz)]) >> (2 * oct)) & 3);
      } else {
        if ((w.wfull >> bit) & 1ull) return W_HIT;
        w.slot = __ldg(&v.bptr[(size_t)w.ci * MESO_BLOCKS + (bx + 16 * by + 256 * bz)]);
        w.cm = __ldg(&v.pool_cm[w.slot]);
        if (STATS) s.touch_brick[w.slot] = 1;
        w.ztag = -1;
      }
    }
    if (go && w.need >= 1) {
      const int ce = ((cx >> 1) & 3) + 4 * ((cy >> 1) & 3) + 16 * ((cz >> 1) & 3);
      if (!((w.cm >> ce) & 1ull)) {
        sh = 1; go = false;
        kdf = 1 + (((int)__ldg(&s.ct.cell2[(size_t)w.slot * 64 + ce]) >> (2 * oct)) & 3);
      }
    }
  } else {
  if (w.need >= 3) {
    const int ex = cx >> 5, ey = cy >> 5, ez = cz >> 5;
    const int df = (int)__ldg(&v.df[ex + v.ddims[0] * (ey + v.ddims[1] * ez)]);
    if (df > 0) {   // the cube of half-width df - 1 cells around this cell is empty
      sh = 5; kdf = df; go = false;
      // Probe ahead: the cell m = c + df along the ray's octant diagonal.  If its own empty cube has half-width >= df it
      // contains this cell too, and the box [c, c + df + df(m) - 1] (forward only) lies inside it: a ray never needs
      // what is behind it, so rays leaving or skimming a surface take steps about twice as long for one more byte.
      const int mx = ex + ((df ^ gx) - gx), my = ey + ((df ^ gy) - gy), mz = ez + ((df ^ gz) - gz);
      if ((unsigned)mx < (unsigned)v.ddims[0] && (unsigned)my < (unsigned)v.ddims[1] && (unsigned)mz < (unsigned)v.ddims[2]) {
        const int d2 = (int)__ldg(&v.df[mx + v.ddims[0] * (my + v.ddims[1] * mz)]);
        if (d2 > df) kdf = df + d2;
      }
    } else {
      const int ci = (cx >> 7) + v.dims[0] * ((cy >> 7) + v.dims[1] * (cz >> 7));
      if (ci != w.ci) { w.ci = ci; w.wtag = -1; }
      if (STATS) {
        const int e = ((cx >> 5) & 3) + 4 * ((cy >> 5) & 3) + 16 * ((cz >> 5) & 3);
        if ((v.cells[ci] >> e) & 1ull) s.touch_chunk[ci] = 1;   // exact cell occupancy (df may be stale-conservative after a carve)
      }
    }
  }
  if (go && w.need >= 2) {
    const int bx = (cx >> 3) & 15, by = (cy >> 3) & 15, bz = (cz >> 3) & 15;
    const int wi = bz * 4 + (by >> 2);
    if (wi != w.wtag) {
      const ulonglong2 p = __ldg(&v.of[(size_t)w.ci * 64 + wi]);
      w.wocc = p.x; w.wfull = p.y; w.wtag = wi;
    }
    const int bit = bx + 16 * (by & 3);
    if (!((w.wocc >> bit) & 1ull)) { sh = 3; go = false; }
    else {
      if ((w.wfull >> bit) & 1ull) return W_HIT;
      w.slot = __ldg(&v.bptr[(size_t)w.ci * MESO_BLOCKS + (bx + 16 * by + 256 * bz)]);
      w.cm = __ldg(&v.pool_cm[w.slot]);
      if (STATS) s.touch_brick[w.slot] = 1;
      w.ztag = -1;
    }
  }
  if (go && w.need >= 1) {
    // 2^3-voxel cells of the partial brick
    // (a 4^3 level on top of this one was measured: 9 % fewer steps, 7 % slower -- one more divergent path per iteration)
    if (!((w.cm >> (((cx >> 1) & 3) + 4 * ((cy >> 1) & 3) + 16 * ((cz >> 1) & 3))) & 1ull)) { sh = 1; go = false; }
  }
  }  // !CUBES
  if (go) {
    const int z = cz & 7;
    if (z != w.ztag) { w.slice = __ldg(&v.pool[(size_t)w.slot * 8 + z]); w.ztag = z; }
    if ((w.slice >> ((cx & 7) + 8 * (cy & 7))) & 1ull) return W_HIT;
  }
  // ---- one step at level sh: consume the smallest of the three pending keys (ties: lower axis first) ----
  // mirrored coordinate after crossing: kdf cells of size 2^sh ahead, never beyond the grid end (nvox if step > 0, else 0);
  // for the aligned levels (kdf = 1) this is (cs | mask) + 1 and the clamp never binds
  const int nxx = min(((w.csx >> sh) + kdf) << sh, v.nvox[0] & ~gx);
  const int nxy = min(((w.csy >> sh) + kdf) << sh, v.nvox[1] & ~gy);
  const int nxz = min(((w.csz >> sh) + kdf) << sh, v.nvox[2] & ~gz);
  const float tx = r.sx != 0 ? plane_t1(r.ox, r.ix, (nxx ^ gx) - gx) : F_INF;               // (n ^ g) - g = true plane index
  const float ty = r.sy != 0 ? plane_t1(r.oy, r.iy, (nxy ^ gy) - gy) : F_INF;
  const float tz = r.sz != 0 ? plane_t1(r.oz, r.iz, (nxz ^ gz) - gz) : F_INF;
  int a = 0; float ta = tx;
  if (ty < ta) { a = 1; ta = ty; }
  if (tz < ta) { a = 2; ta = tz; }
  if (!(ta < F_INF)) return W_EXIT;  // zero direction
  const int olds = SEL3(a, w.csx, w.csy, w.csz);
  const int news = SEL3(a, nxx, nxy, nxz);
  if (a == 0) w.csx = nxx; else if (a == 1) w.csy = nxy; else w.csz = nxz;
  w.la = a; w.lt = ta; steps++;
  if (STATS) s.lv[sh == 0 ? 0 : (sh <= 2 ? 1 : (sh == 3 ? 2 : (kdf <= 2 ? 3 : 4)))]++;
  const unsigned ucross = (unsigned)(olds ^ news);
  w.need = (ucross >> 5) ? 3 : ((ucross >> 3) ? 2 : ((ucross >> 1) ? 1 : 0));
  if (w.need >= 3) {
    const int st = SEL3(a, r.sx, r.sy, r.sz);
    const int nv = SEL3(a, v.nvox[0], v.nvox[1], v.nvox[2]);
    if (news >= (st > 0 ? nv : 0)) return W_EXIT;   // mirrored coordinate at which the ray has left the grid
  }
  // the one place where the two other axes are made exact: right after every step that left a box bigger than a voxel
  if (sh > 0) {
    cx = w.csx ^ gx; cy = w.csy ^ gy; cz = w.csz ^ gz;
    sync_axes(r, w, cx, cy, cz);
    w.csx = cx ^ gx; w.csy = cy ^ gy; w.csz = cz ^ gz;
  }
  return W_CONTINUE;
}

__device__ __forceinline__ uint32_t to_un8(float x) { return (uint32_t)__fadd_rn(__fmul_rn(x, 255.0f), 0.5f); }

// per-lane result of a finished pixel, kept until the next batched retire
struct Done {
  int cx, cy, cz;   // hit voxel
  int face;         // 0..5, 6 inside, 7 miss
  int shadow;
  float t;
  float px, py, pz; // hit point
};

__device__ __forceinline__ uint32_t shade1(float p, int c, float shade) {
  float local = __fsub_rn(__fmul_rn(p, 0.125f), (float)(c >> 3));
  local = fminf(fmaxf(local, 0.0f), 1.0f);
  const float col = __fadd_rn(__fmul_rn(__fsub_rn(local, 0.5f), 0.5f), 0.5f);  // SimpleVoxel.cpp:222
  return to_un8(__fmul_rn(col, shade));
}
__device__ __forceinline__ uint4 shade_record(const Done& dn) {
  if (dn.face == 7) return make_uint4(0xFFFFFFFFu, 0x0007FFFFu, 0x7F800000u, 0xFF000000u);
  const float shade = dn.shadow ? 0.5f : 1.0f;
  const uint32_t r = shade1(dn.px, dn.cx, shade), g = shade1(dn.py, dn.cy, shade), b = shade1(dn.pz, dn.cz, shade);
  return make_uint4((uint32_t)dn.cx | ((uint32_t)dn.cy << 16),
                    (uint32_t)dn.cz | ((uint32_t)dn.face << 16) | ((uint32_t)dn.shadow << 19) | (1u << 20),
                    __float_as_uint(dn.t), r | (g << 8) | (b << 16));
}

// CTA = one 32x8 screen tile; warp = 8x4 pixels (four full 128 B lines per record store).  All primary rays of a warp
// start together from the same eye, so the lanes stay at similar levels of the walk; when all of them are done, the lanes
// that hit a lit-facing face trace that pixel's shadow ray, again all starting together.  Keeping phases apart matters:
// refilling idle lanes with new pixels (v4) or letting a lane continue as its shadow ray inside the primary loop (v9)
// both raise lane occupancy and both cost 30 % or more extra warp instructions (profiles/README.md).
// No shared memory, no barriers: the CTA is only the unit of tile ownership.
#define RM_KERNEL_NAME raymarch_kernel
#define RM_EXTRA_PARAM
#define RM_CUBES false
#define RM_SET_CUBES(sc)
#include "raymarch_kernel.inc"
#undef RM_KERNEL_NAME
#undef RM_EXTRA_PARAM
#undef RM_CUBES
#undef RM_SET_CUBES

// the same frame through the per-octant forward cubes (MESO_FLAG_CUBES)
#define RM_KERNEL_NAME raymarch_cubes_kernel
#define RM_EXTRA_PARAM , CubeTables ct
#define RM_CUBES true
#define RM_SET_CUBES(sc) (sc).ct = ct
#include "raymarch_kernel.inc"
#undef RM_KERNEL_NAME
#undef RM_EXTRA_PARAM
#undef RM_CUBES
#undef RM_SET_CUBES

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
                     uint8_t* d_touch_brick, int local_tile0, int local_tile_count, const CubeTables* cubes) {
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W, tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  const int n_tiles = tiles_x * tiles_y;
  const int all_local = (n_tiles - rank + world - 1) / world;
  if (local_tile_count < 0) local_tile_count = all_local - local_tile0;
  if (local_tile_count <= 0) return;
  const size_t smem = 0;
  // (Dispatching the tiles in a golden-ratio permuted order, to spread the expensive silhouette tiles over the launch,
  // was measured: no gain in the pipelined loop, 2 % slower alone -- neighbouring tiles share distance-field and brick lines.)
  if (cubes) {
    if (d_stats)
      raymarch_cubes_kernel<true><<<local_tile_count, RM_THREADS, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, tiles_x,
                                                                                     n_tiles, local_tile0, d_out, d_stats, d_touch_chunk, d_touch_brick, *cubes);
    else
      raymarch_cubes_kernel<false><<<local_tile_count, RM_THREADS, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, tiles_x,
                                                                                      n_tiles, local_tile0, d_out, nullptr, nullptr, nullptr, *cubes);
    (*lc.launches)++;
    return;
  }
  if (d_stats)
    raymarch_kernel<true><<<local_tile_count, RM_THREADS, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, tiles_x,
                                                                             n_tiles, local_tile0, d_out, d_stats, d_touch_chunk, d_touch_brick);
  else
    raymarch_kernel<false><<<local_tile_count, RM_THREADS, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, tiles_x,
                                                                              n_tiles, local_tile0, d_out, nullptr, nullptr, nullptr);
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
