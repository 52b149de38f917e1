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
#include <cstdio>

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
  const int ox0 = w.csx, oy0 = w.csy, oz0 = w.csz;   // CUBES: cell before the step, all axes
  if (a == 0) w.csx = nxx; else if (a == 1) w.csy = nxy; else w.csz = nxz;
  w.la = a; w.lt = ta; steps++;
  if (STATS) s.lv[sh == 0 ? 0 : (sh <= 2 ? 1 : (sh == 3 ? 2 : (kdf <= 2 ? 3 : 4)))]++;
  unsigned ucross = (unsigned)(olds ^ news);
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
    if (CUBES) {
      // A forward cube of several bricks / 2^3 cells is unaligned: the two synced axes may have crossed a brick, a 32^3
      // cell or a chunk face inside it although the stepping axis did not -- the cached chunk index, word pair and
      // payload slot are only valid for what `need` says, so it has to look at all three axes here.
      ucross = (unsigned)((ox0 ^ w.csx) | (oy0 ^ w.csy) | (oz0 ^ w.csz));
      w.need = (ucross >> 5) ? 3 : ((ucross >> 3) ? 2 : ((ucross >> 1) ? 1 : 0));
    }
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


// =====================================================================================================================
// v10: the same walk, written in MIRRORED SPACE with a branch-free axis sync.
//
// Mirrored space.  With g = step >> 31 (0 or -1) the walk already keeps cs = c ^ g, in which every active axis moves
// towards +.  v10 mirrors the ray as well: om = g ? -o : o, dm = |d|, im = |1/d|.  The key of mirrored plane m is
//     (float(m) - om) * im
// which is bit-identical to the defining key (float(p) - o) * inv of the true plane p = (m ^ g) - g: for g = -1, p = -m,
// float(p) = -float(m), fl(-fm - o) = -fl(fm - om), and (-a) * inv = a * (-inv) exactly.  So no sign handling is left in
// the loop: one I2F, one FSUB, one FMUL per pending key.  Inactive axes (|d| < 1e-20) carry im = +inf: their planes are
// never the smallest key.
//
// Axis sync.  After a step out of a box bigger than a voxel the two other axes need "the number of planes whose key is
// smaller than the consumed key".  A computed key is within 1.8e-7 relative of the real crossing time, so that count lies
// within floor(x -+ 1.8e-7 t) of the real position x = om + dm t, and the estimate floor(fma(dm, t, om)) within
// floor(x -+ 6e-8 |x|): while |o| <= 1e6 (hence t < 2e6 inside a grid of at most 65 536 voxels per axis) the two differ by
// at most one, and testing the keys of the two planes that bound the estimated cell decides it exactly -- fixed cost, no
// data-dependent loop, no divergent slow path.  A ray whose origin is farther out than 1e6 voxels takes the loop form for
// its whole walk (walk10_slow).  Result identical to oracle/orc_raymarch.c either way.
#define RM10_FAR 1.0e6f
struct MRay {
  float ox, oy, oz, dx, dy, dz, ix, iy, iz;
  int gx, gy, gz;
};

__device__ __forceinline__ void mirror1(float o, float d, float inv, int st, float& om, float& dm, float& im, int& g) {
  g = st >> 31;
  om = g ? -o : o;
  dm = fabsf(d);
  im = st != 0 ? fabsf(inv) : F_INF;
}
__device__ __forceinline__ void mirror_ray(const Ray& r, MRay& m) {
  mirror1(r.ox, r.dx, r.ix, r.sx, m.ox, m.dx, m.ix, m.gx);
  mirror1(r.oy, r.dy, r.iy, r.sy, m.oy, m.dy, m.iy, m.gy);
  mirror1(r.oz, r.dz, r.iz, r.sz, m.oz, m.dz, m.iz, m.gz);
}

// loop form (any magnitude), mirrored
__device__ __noinline__ int sync1_slow(float om, float dm, float im, int b, int cur, float ts, int as) {
  if (!(im < F_INF)) return cur;
  const float fl = floorf(__fadd_rn(om, __fmul_rn(dm, ts)));
  int e = max((int)fminf(fmaxf(fl, -1.0e9f), 1.0e9f), cur);
  for (;;) { if (key_less(__fmul_rn(__fsub_rn((float)(e + 1), om), im), b, ts, as)) e++; else break; }
  for (;;) {
    if (e == cur) break;
    if (!key_less(__fmul_rn(__fsub_rn((float)e, om), im), b, ts, as)) e--; else break;
  }
  return e;
}

// mirrored coordinate on axis b after consuming every crossing with key < (ts, as); cur = the (older) exact coordinate.
// tsi = float bits of ts, plus one if a tie with this axis goes to this axis (b < as): for the non-negative keys of a walk,
// "k < ts || (k == ts && b < as)" is the integer comparison bits(k) < tsi (a negative key is a negative integer: smaller,
// as it should be).  Two integer compares and two predicated adds; everything else runs on the FMA / conversion pipes --
// the walk is bound by the half-rate ALU pipe (compares, logic, shifts, selects), ncu profiles/r2_*.
__device__ __forceinline__ int sync1(float om, float dm, float im, int cur, float ts, int tsi) {
  const float pm = fmaf(dm, ts, om);                                   // estimate only: any rounding will do
  const float fl = floorf(pm);
  const float kb = __fmul_rn(__fsub_rn(fl, om), im);                    // key of the plane into the estimated cell
  const float ka = __fmul_rn(__fsub_rn(__fadd_rn(fl, 1.0f), om), im);   // key of the plane out of it
  int e = __float2int_rd(pm);
  if (__float_as_int(ka) < tsi) e += 1;
  if (!(__float_as_int(kb) < tsi)) e -= 1;
  return max(e, cur);   // an inactive axis (im = +inf) can produce NaN keys: it stays where it is
}

struct Walk10 {
  int csx, csy, csz;
  int la; float lt;
  unsigned ux;       // xor of the cell before / after the last step, all axes: >= 32 the 32^3 cell may have changed (look the
                     // field / cubes up), >= 8 the brick, >= 2 the 2^3 cell, else same 2^3 cell
  int ci, wtag, ztag;
  uint32_t slot;
  unsigned long long wocc, wfull, slice, cm;
};

struct Scene10 {
  uint8_t* touch_chunk;
  uint8_t* touch_brick;
  unsigned* lv;
  const uint8_t* cellp;      // CL >= 1: padded cell cubes (CubeTables.cellp)
  const uint16_t* brick;     // CL >= 2
  const uint16_t* cell2;     // CL >= 3
  int oct_off;               // CL >= 1: the ray's octant: octant * npcells + (1, 1, 1) of the padded grid
  int oct2;                  // 2 * octant
  int pd0, pd01;             // CL >= 1: padded cell grid: row / slice pitch
};

// bit i (0..63) of w as a 32-bit value in bit 0.  The empty asm keeps the compiler from widening the test back to 64 bits
// (it otherwise compares a register pair: two more instructions on the half-rate ALU pipe per test).
__device__ __forceinline__ unsigned bit64(unsigned long long w, unsigned i) {
  unsigned s = (unsigned)(w >> i);
  asm("" : "+r"(s));
  return s & 1u;
}

// One ray from its first cell inside the grid to a hit or the grid's end.  CL = how many levels read forward cubes:
// 0 = distance field + probe-ahead (the v8 boxes), 1 = per-octant cell cubes, 2 = + brick cubes, 3 = + 2^3-cell cubes.
// Returns W_HIT with (cx, cy, cz) the exact hit voxel, or W_EXIT.
template <bool STATS, int CL, bool SLOW>
__device__ __forceinline__ int walk10(const DVolume& v, const Scene10& s, const MRay& r, Walk10& w, int& cx, int& cy, int& cz, unsigned& steps) {
  const int gx = r.gx, gy = r.gy, gz = r.gz;
  const int bx_end = v.nvox[0] & ~gx, by_end = v.nvox[1] & ~gy, bz_end = v.nvox[2] & ~gz;   // mirrored plane at which the ray has left the grid
  for (;;) {
    cx = w.csx ^ gx; cy = w.csy ^ gy; cz = w.csz ^ gz;
    int sz = 1, kdf = 1;   // the step: kdf boxes of edge sz (a power of two: 32 cell, 8 brick, 2 cell, 1 voxel)
    do {   // classify the current cell from the coarsest level that changed down to the first empty level (or a solid voxel)
      if (w.ux >= 32u) {
        const int ex = cx >> 5, ey = cy >> 5, ez = cz >> 5;
        if (CL >= 1) {
          // padded table: one border cell all around that reads 255 = "outside the grid" -- the walk has no other exit test
          const int k = (int)__ldg(s.cellp + (unsigned)(s.oct_off + ex + s.pd0 * ey + s.pd01 * ez));
          if (k == 255) return W_EXIT;
          if (k > 0) { sz = 32; kdf = k; break; }
        } else {
          if ((unsigned)cx >= (unsigned)v.nvox[0] || (unsigned)cy >= (unsigned)v.nvox[1] || (unsigned)cz >= (unsigned)v.nvox[2]) return W_EXIT;
          const int e = ex + v.ddims[0] * (ey + v.ddims[1] * ez);
          const int df = (int)__ldg(v.df + e);
          if (df > 0) {
            sz = 32; kdf = df;
            const int mx = ex + ((df ^ gx) - gx), my = ey + ((df ^ gy) - gy), mz = ez + ((df ^ gz) - gz);
            if ((unsigned)mx < (unsigned)v.ddims[0] && (unsigned)my < (unsigned)v.ddims[1] && (unsigned)mz < (unsigned)v.ddims[2]) {
              const int d2 = (int)__ldg(v.df + (mx + v.ddims[0] * (my + v.ddims[1] * mz)));
              if (d2 > df) kdf = df + d2;
            }
            break;
          }
        }
        const int ci = (cx >> 7) + v.dims[0] * ((cy >> 7) + v.dims[1] * (cz >> 7));
        if (ci != w.ci) { w.ci = ci; w.wtag = -1; }
        if (STATS) {
          const int c64 = ((cx >> 5) & 3) + 4 * ((cy >> 5) & 3) + 16 * ((cz >> 5) & 3);
          if ((v.cells[ci] >> c64) & 1ull) s.touch_chunk[ci] = 1;
        }
      }
      if (w.ux >= 8u) {
        const unsigned b12 = ((unsigned)(cx >> 3) & 15u) | (((unsigned)(cy >> 3) & 15u) << 4) | (((unsigned)(cz >> 3) & 15u) << 8);
        const int wi = (int)(b12 >> 6);
        const unsigned bit = b12 & 63u;
#ifdef RM10_DEBUG
        if (w.ci < 0 || w.ci >= (int)v.nchunks) {
          printf("RM10 bad ci=%d c=(%d,%d,%d) cs=(%d,%d,%d) g=(%d,%d,%d) ux=%u steps=%u la=%d lt=%g o=(%g,%g,%g) d=(%g,%g,%g) i=(%g,%g,%g)\n", w.ci, cx, cy, cz,
                 w.csx, w.csy, w.csz, gx, gy, gz, w.ux, steps, w.la, w.lt, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, r.ix, r.iy, r.iz);
          return W_EXIT;
        }
#endif
        if (wi != w.wtag) {
          const ulonglong2 p = __ldg(v.of + ((unsigned)w.ci * 64u + (unsigned)wi));
          w.wocc = p.x; w.wfull = p.y; w.wtag = wi;
        }
        if (!bit64(w.wocc, bit)) {
          sz = 8;
          if (CL >= 2) kdf = 1 + (((int)__ldg(s.brick + ((unsigned)w.ci * (unsigned)MESO_BLOCKS + b12)) >> s.oct2) & 3);
          break;
        }
        if (bit64(w.wfull, bit)) return W_HIT;
        w.slot = __ldg(v.bptr + ((unsigned)w.ci * (unsigned)MESO_BLOCKS + b12));
        w.cm = __ldg(v.pool_cm + w.slot);
        if (STATS) s.touch_brick[w.slot] = 1;
        w.ztag = -1;
      }
      if (w.ux >= 2u) {
        const unsigned ce = ((unsigned)(cx >> 1) & 3u) | (((unsigned)(cy >> 1) & 3u) << 2) | (((unsigned)(cz >> 1) & 3u) << 4);
        if (!bit64(w.cm, ce)) {
          sz = 2;
          if (CL >= 3) kdf = 1 + (((int)__ldg(s.cell2 + (w.slot * 64u + ce)) >> s.oct2) & 3);
          break;
        }
      }
      const int z = cz & 7;
      if (z != w.ztag) { w.slice = __ldg(v.pool + ((size_t)w.slot * 8 + z)); w.ztag = z; }
      if (bit64(w.slice, ((unsigned)cx & 7u) | (((unsigned)cy & 7u) << 3))) return W_HIT;
    } while (0);
    // ---- one step: kdf boxes of edge sz; consume the smallest pending key (ties: lower axis first) ----
    const int nmask = -sz, add = kdf * sz;
    const int nxx = min((w.csx & nmask) + add, bx_end);
    const int nxy = min((w.csy & nmask) + add, by_end);
    const int nxz = min((w.csz & nmask) + add, bz_end);
    const float tx = __fmul_rn(__fsub_rn((float)nxx, r.ox), r.ix);
    const float ty = __fmul_rn(__fsub_rn((float)nxy, r.oy), r.iy);
    const float tz = __fmul_rn(__fsub_rn((float)nxz, r.oz), r.iz);
    const bool yx = ty < tx;
    const float txy = yx ? ty : tx;
    const bool zm = tz < txy;
    const float ta = zm ? tz : txy;
    if (!(ta < F_INF)) return W_EXIT;   // zero direction
    const int ox0 = w.csx, oy0 = w.csy, oz0 = w.csz;
    const int a = zm ? 2 : (yx ? 1 : 0);
    w.la = a; w.lt = ta; steps++;
    if (STATS) s.lv[sz == 1 ? 0 : (sz == 2 ? 1 : (sz == 8 ? 2 : (kdf <= 2 ? 3 : 4)))]++;
    // the two axes that did not step are made exact here, the one place where that happens (a step out of a single voxel
    // cannot have crossed a plane of another axis: those two stay as they are)
    int ex = ox0, ey = oy0, ez = oz0;
    if (sz > 1) {
      if (SLOW) {
        ex = sync1_slow(r.ox, r.dx, r.ix, 0, ox0, ta, a);
        ey = sync1_slow(r.oy, r.dy, r.iy, 1, oy0, ta, a);
        ez = sync1_slow(r.oz, r.dz, r.iz, 2, oz0, ta, a);
      } else {
        const int ti = __float_as_int(ta);
        ex = sync1(r.ox, r.dx, r.ix, ox0, ta, ti + 1);              // axis 0 wins every tie
        ey = sync1(r.oy, r.dy, r.iy, oy0, ta, ti + (zm ? 1 : 0));   // axis 1 wins a tie against axis 2 only
        ez = sync1(r.oz, r.dz, r.iz, oz0, ta, ti);                 // axis 2 never does
      }
    }
    // (written as three selects on `a`: an earlier form -- assign the stepped axis, then conditionally overwrite the two
    // others under zm / yx predicates -- came out of the compiler with the stepped y coordinate reverted to its old value)
    w.csx = a == 0 ? nxx : ex;
    w.csy = a == 1 ? nxy : ey;
    w.csz = a == 2 ? nxz : ez;
    // a forward cube of several bricks / 2^3 cells is unaligned: the synced axes may have crossed a brick, a 32^3 cell or
    // a chunk face inside it, so this looks at all three axes
    w.ux = (unsigned)((ox0 ^ w.csx) | (oy0 ^ w.csy) | (oz0 ^ w.csz));
  }
}

// the whole walk of a ray that starts farther than RM10_FAR from the grid's corner: loop-form axis sync, out of line
template <bool STATS, int CL>
__device__ __noinline__ int walk10_slow(const DVolume& v, const Scene10& s, const MRay& r, Walk10& w, int& cx, int& cy, int& cz, unsigned& steps) {
  return walk10<STATS, CL, true>(v, s, r, w, cx, cy, cz, steps);
}
__device__ __forceinline__ bool ray_is_far(const MRay& r) { return fmaxf(fmaxf(fabsf(r.ox), fabsf(r.oy)), fabsf(r.oz)) > RM10_FAR; }

// first cell of a ray (walk_begin, true space) -> mirrored walk state
__device__ __forceinline__ bool walk10_begin(const DVolume& v, const Ray& r, int cx, int cy, int cz, Walk10& w, unsigned& steps) {
  Walk w0;
  const bool alive = walk_begin(v, r, cx, cy, cz, w0, steps);
  w.csx = w0.csx; w.csy = w0.csy; w.csz = w0.csz; w.la = w0.la; w.lt = w0.lt;
  w.ux = 0xFFFFFFFFu; w.ci = -1; w.wtag = -1; w.ztag = -1; w.slot = 0; w.wocc = 0; w.wfull = 0; w.slice = 0; w.cm = 0;
  return alive;
}

// An inactive axis never moves: its cell is the start cell whatever the origin says (a shadow ray starts from a computed
// hit point, which may round into the neighbouring cell).  Park its origin in the middle of that cell so that the sync's
// estimate stays there; its keys are +-inf either way.
__device__ __forceinline__ void park_inactive(MRay& m, const Walk10& w) {
  if (!(m.ix < F_INF)) m.ox = __fadd_rn((float)w.csx, 0.5f);
  if (!(m.iy < F_INF)) m.oy = __fadd_rn((float)w.csy, 0.5f);
  if (!(m.iz < F_INF)) m.oz = __fadd_rn((float)w.csz, 0.5f);
}

template <int CL>
__device__ __forceinline__ void scene10_octant(Scene10& sc, const CubeTables& ct, const MRay& m) {
  if (CL >= 1) {
    const int oct = (m.gx & 1) | (m.gy & 2) | (m.gz & 4);
    int off = oct * (int)ct.npcells + (1 + ct.pd0 + ct.pd01);   // (+1, +1, +1): skip the border
    asm("" : "+r"(off));   // one register, kept: left to itself the compiler re-derives it from the ray's signs on every lookup
    sc.oct_off = off; sc.oct2 = 2 * oct;
  }
}

// CTA = RM10_THREADS / 32 warps of one 32x8 screen tile (8 warps of 8x4 pixels); 256 / RM10_THREADS CTAs per tile.  Smaller
// CTAs give their registers back as soon as their own slowest warp is done (achieved occupancy, profiles/r2_*).
#ifndef RM10_MINB
#define RM10_MINB 4
#endif
#ifndef RM10_THREADS
#define RM10_THREADS 128
#endif
#define RM10_SPLIT (256 / RM10_THREADS)
template <bool STATS, int CL>
__global__ void __launch_bounds__(RM10_THREADS, RM10_MINB *(256 / RM10_THREADS)) raymarch10_kernel(DVolume v, MesoRaySetup rs, int width, int height, uint32_t flags,
                                                                   int rank, int world, int layout, int tiles_x, int n_tiles, int local_tile0,
                                                                   MesoHitRecord* __restrict__ out, RayStatsDev* stats,
                                                                   uint8_t* touch_chunk, uint8_t* touch_brick, CubeTables ct) {
  const int warp = (threadIdx.x >> 5) + (blockIdx.x % RM10_SPLIT) * (RM10_THREADS / 32), lane = threadIdx.x & 31;
  const int local_tile = local_tile0 + blockIdx.x / RM10_SPLIT;
  const int tile = local_tile * world + rank;
  const int tx = (warp & 3) * 8 + (lane & 7), ty = (warp >> 2) * 4 + (lane >> 3);
  const int px = (tile % tiles_x) * MESO_TILE_W + tx;
  const int py = (tile / tiles_x) * MESO_TILE_H + ty;
  const bool valid = tile < n_tiles && px < width && py < height;
  unsigned lv[5] = {0, 0, 0, 0, 0};
  Scene10 sc; sc.touch_chunk = touch_chunk; sc.touch_brick = touch_brick; sc.lv = lv;
  sc.cellp = ct.cellp; sc.oct_off = 0; sc.brick = ct.brick; sc.cell2 = ct.cell2; sc.oct2 = 0; sc.pd0 = ct.pd0; sc.pd01 = ct.pd01;
  const float Lx = rs.L[0], Ly = rs.L[1], Lz = rs.L[2];

  Done dn; dn.cx = dn.cy = dn.cz = 0; dn.face = 7; dn.shadow = 0; dn.t = 0.f; dn.px = dn.py = dn.pz = 0.f;
  bool want_shadow = false;
  int hit_axis = -1;
  unsigned steps = 0, n_shadow = 0;
  if (valid) {
    const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)px, 0.5f), rs.two_over_w), 1.0f);
    const float fy = __fsub_rn(1.0f, __fmul_rn(__fadd_rn((float)py, 0.5f), rs.two_over_h));
    float dx = __fadd_rn(__fadd_rn(__fmul_rn(fx, rs.U[0]), __fmul_rn(fy, rs.V[0])), rs.F[0]);
    float dy = __fadd_rn(__fadd_rn(__fmul_rn(fx, rs.U[1]), __fmul_rn(fy, rs.V[1])), rs.F[1]);
    float dz = __fadd_rn(__fadd_rn(__fmul_rn(fx, rs.U[2]), __fmul_rn(fy, rs.V[2])), rs.F[2]);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    dx = __fdiv_rn(dx, len); dy = __fdiv_rn(dy, len); dz = __fdiv_rn(dz, len);
    Ray r; r.ox = rs.o[0]; r.oy = rs.o[1]; r.oz = rs.o[2];
    ray_dir(r, dx, dy, dz);
    Walk10 w;
    MRay m;
    int cx = 0, cy = 0, cz = 0, res = W_EXIT;
    const bool alive = walk10_begin(v, r, clamp_floor_to_int(r.ox), clamp_floor_to_int(r.oy), clamp_floor_to_int(r.oz), w, steps);
    mirror_ray(r, m);   // from here on only the mirrored ray is live (the true one is m with the signs put back)
    if (alive) {
      scene10_octant<CL>(sc, ct, m);
      const bool far = ray_is_far(m);
      park_inactive(m, w);
#ifdef RM10_NO_SLOW
      (void)far; res = walk10<STATS, CL, false>(v, sc, m, w, cx, cy, cz, steps);
#else
      res = far ? walk10_slow<STATS, CL>(v, sc, m, w, cx, cy, cz, steps) : walk10<STATS, CL, false>(v, sc, m, w, cx, cy, cz, steps);
#endif
    }
    if (res == W_HIT) {
      dn.t = w.lt; dn.cx = cx; dn.cy = cy; dn.cz = cz;
      // o + d t in true space (o from the launch constants: an inactive axis of m has its origin parked)
      dn.px = __fadd_rn(rs.o[0], __fmul_rn(m.gx ? -m.dx : m.dx, w.lt));
      dn.py = __fadd_rn(rs.o[1], __fmul_rn(m.gy ? -m.dy : m.dy, w.lt));
      dn.pz = __fadd_rn(rs.o[2], __fmul_rn(m.gz ? -m.dz : m.dz, w.lt));
      if (w.la >= 0) {
        const int g_ax = SEL3(w.la, m.gx, m.gy, m.gz);   // the hit axis is an active one: step > 0 <=> g == 0
        const float l_ax = SEL3(w.la, Lx, Ly, Lz);
        const int c_ax = SEL3(w.la, cx, cy, cz);
        const float pl = (float)(c_ax - g_ax);
        if (w.la == 0) dn.px = pl; else if (w.la == 1) dn.py = pl; else dn.pz = pl;
        dn.face = w.la * 2 - g_ax;
        hit_axis = w.la;
        // a crossing at t = 0 on a negative-direction axis: the defining key (+0) * (negative inv) is -0.0, the mirrored one +0.0
        if (g_ax && w.lt == 0.0f) dn.t = -0.0f;
        if (flags & MESO_FLAG_SHADOW) {
          const bool facing = g_ax == 0 ? (l_ax < 0.0f) : (l_ax > 0.0f);
          if (!facing) dn.shadow = 1; else want_shadow = true;
        }
      } else {
        dn.face = 6;
        dn.px = rs.o[0]; dn.py = rs.o[1]; dn.pz = rs.o[2];
      }
    }
  }

  const unsigned steps_p = steps;
  if (flags & MESO_FLAG_SHADOW) {
    if (want_shadow) {
      const int nrm = (dn.face & 1) ? 1 : -1;
      Ray r; r.ox = dn.px; r.oy = dn.py; r.oz = dn.pz;
      ray_dir(r, Lx, Ly, Lz);
      Walk10 w;
      int cx = 0, cy = 0, cz = 0, res = W_EXIT;
      if (walk10_begin(v, r, dn.cx + (hit_axis == 0 ? nrm : 0), dn.cy + (hit_axis == 1 ? nrm : 0), dn.cz + (hit_axis == 2 ? nrm : 0), w, steps)) {
        MRay m; mirror_ray(r, m);
        scene10_octant<CL>(sc, ct, m);
        const bool far = ray_is_far(m);
        park_inactive(m, w);
#ifdef RM10_NO_SLOW
        (void)far; res = walk10<STATS, CL, false>(v, sc, m, w, cx, cy, cz, steps);
#else
        res = far ? walk10_slow<STATS, CL>(v, sc, m, w, cx, cy, cz, steps) : walk10<STATS, CL, false>(v, sc, m, w, cx, cy, cz, steps);
#endif
      }
      dn.shadow = res == W_HIT ? 1 : 0;
      n_shadow = 1;
    }
  }

  if (valid) {
    const size_t dst = layout == MESO_LAYOUT_FRAME ? (size_t)py * width + px
                                                   : (size_t)local_tile * (MESO_TILE_W * MESO_TILE_H) + ty * MESO_TILE_W + tx;
    const uint4 rec = shade_record(dn);
    if (flags & MESO_FLAG_RGBA8) reinterpret_cast<uint32_t*>(out)[dst] = rec.w;
    else reinterpret_cast<uint4*>(out)[dst] = rec;
  }

  if (STATS) {
    unsigned long long v0 = valid ? 1 : 0, v1 = n_shadow, v2 = dn.face != 7 ? 1 : 0, v3 = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
    }
    unsigned long long v4 = steps_p;
    unsigned mp = steps_p, ms = steps - steps_p;
    unsigned long long l0 = lv[0], l1 = lv[1], l2 = lv[2], l3 = lv[3], l4 = lv[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v4 += __shfl_xor_sync(0xffffffffu, v4, o);
      mp = max(mp, __shfl_xor_sync(0xffffffffu, mp, o)); ms = max(ms, __shfl_xor_sync(0xffffffffu, ms, o));
      l0 += __shfl_xor_sync(0xffffffffu, l0, o); l1 += __shfl_xor_sync(0xffffffffu, l1, o); l2 += __shfl_xor_sync(0xffffffffu, l2, o);
      l3 += __shfl_xor_sync(0xffffffffu, l3, o); l4 += __shfl_xor_sync(0xffffffffu, l4, o);
    }
    if (lane == 0) {
      atomicAdd(&stats->primary, v0); atomicAdd(&stats->shadow, v1);
      atomicAdd(&stats->hits, v2); atomicAdd(&stats->steps, v3);
      atomicAdd(&stats->steps_primary, v4);
      atomicAdd(&stats->warp_slots_primary, 32ull * mp); atomicAdd(&stats->warp_slots_shadow, 32ull * ms);
      atomicAdd(&stats->level_steps[0], l0); atomicAdd(&stats->level_steps[1], l1); atomicAdd(&stats->level_steps[2], l2);
      atomicAdd(&stats->level_steps[3], l3); atomicAdd(&stats->level_steps[4], l4);
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
                     uint8_t* d_touch_brick, int local_tile0, int local_tile_count, const CubeTables* cubes) {
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W, tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  const int n_tiles = tiles_x * tiles_y;
  const int all_local = (n_tiles - rank + world - 1) / world;
  if (local_tile_count < 0) local_tile_count = all_local - local_tile0;
  if (local_tile_count <= 0) return;
  const size_t smem = 0;
  // A/B switch while both walks exist: MESO_RM_KERNEL=v8 selects the round-1 kernels, anything else v10;
  // MESO_CUBES_LEVEL = 1..3 = how many levels of the v10 walk read forward cubes under MESO_FLAG_CUBES.
  const bool use_v8 = [] { const char* e = getenv("MESO_RM_KERNEL"); return e && strcmp(e, "v8") == 0; }();
  const int cubes_level = [] { const char* e = getenv("MESO_CUBES_LEVEL"); const int l = e ? atoi(e) : 3; return l < 1 ? 1 : (l > 3 ? 3 : l); }();
  if (!use_v8) {
    const CubeTables ct = cubes ? *cubes : CubeTables{};
    const int cl = cubes ? cubes_level : 0;
#define RM10_LAUNCH(ST, CL)                                                                                                        \
    raymarch10_kernel<ST, CL><<<local_tile_count * RM10_SPLIT, RM10_THREADS, smem, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, \
                                                                                 tiles_x, n_tiles, local_tile0, d_out, d_stats,    \
                                                                                 d_touch_chunk, d_touch_brick, ct)
    if (d_stats) { if (cl == 0) RM10_LAUNCH(true, 0); else if (cl == 1) RM10_LAUNCH(true, 1); else if (cl == 2) RM10_LAUNCH(true, 2); else RM10_LAUNCH(true, 3); }
    else         { if (cl == 0) RM10_LAUNCH(false, 0); else if (cl == 1) RM10_LAUNCH(false, 1); else if (cl == 2) RM10_LAUNCH(false, 2); else RM10_LAUNCH(false, 3); }
#undef RM10_LAUNCH
    (*lc.launches)++;
    return;
  }
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
