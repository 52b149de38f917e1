// k_raymarch.cu -- K4: per-pixel 3D-DDA, primary + one hard-shadow ray.
//
// Replaces the reference's per-pixel visibility: the 1 M-instance triangle-fan draw + reverse-Z depth test
// (Samples/SimpleVoxel.cpp:146-192 VS, :220-224 FS, dispatch :352-398) -- see DESIGN.md for the equivalence.
//
// The DDA is STATELESS (DESIGN.md "DDA"): the crossing time of integer voxel plane p on axis a is always
//     t_a(p) = (float(p) - o_a) * inv_a          one FSUB + one FMUL, never fused (-fmad=false)
// and crossings are consumed in the total order (t, axis).  Skipping an empty box -- an aligned 8^3 brick or 2^3 cell, or
// an unaligned cube of 32^3 cells / bricks that a table certifies empty -- consumes the smallest of the box's three exit
// keys and re-derives the other two coordinates from the same keys (sync1), so the walk visits exactly the voxels the
// oracle's flat walk would: hit voxel, face and t are bit-identical to oracle/orc_raymarch.c.  Any empty box is a legal
// skip; how big the boxes are only changes the number of steps, never the result.
//
// Execution model: 32x8 screen tile = 8 warps of 8x4 pixels (tile t belongs to rank t % world), two 128-thread CTAs per
// tile.  All primary rays of a warp start together at the eye and stay at similar levels of the walk; afterwards the
// lanes that hit a lit-facing face trace their shadow rays, starting together again.  Every level of the walk shares
// ONE step section per iteration (separate step code per level costs more warp instructions than it saves: a warp pays
// for every path any of its lanes takes).  Table bytes come through L1/L2; {occ,full} word pairs (16 B loads), the
// 2^3-cell mask and the brick slice are cached in registers behind tags.
// Measured and dropped (profiles/README.md): persistent warps refilling idle lanes, primary-into-shadow continuation,
// CTA-wide / global shadow-ray compaction (again in round 2 on this kernel: 7.7 vs 8.3 Grays/s), a 4^3 level, per-level step code.
//
// All per-ray state is kept in named scalars (x/y/z members), never in indexable arrays: the compiler turns
// "if (i == a) v = arr[i]" chains into dynamically indexed loads, which pushes the whole state to local memory.
#include "meso_internal.cuh"
#include <cstdlib>
#include <cstring>
#include <algorithm>

#define F_INF __int_as_float(0x7F800000)
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

enum { W_CONTINUE = 0, W_HIT = 1, W_EXIT = 2 };

// Places the ray at its first cell inside the grid (entry from outside included).  Returns false if it never enters.
// (cx, cy, cz) in: floor of the origin; out: that first cell; (la, lt) = the entry crossing, la = -1 if the ray starts inside.
__device__ __forceinline__ bool walk_begin(const DVolume& v, const Ray& r, int& cx, int& cy, int& cz, int& la, float& lt, unsigned& steps) {
  const int nx = v.nvox[0], ny = v.nvox[1], nz = v.nvox[2];
  la = -1; lt = 0.0f;
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
      la = a; lt = ta; steps++;
      alive = (unsigned)cx < (unsigned)nx && (unsigned)cy < (unsigned)ny && (unsigned)cz < (unsigned)nz;
    }
  }
  return alive;
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

// =====================================================================================================================
// The walk, written in MIRRORED SPACE with a branch-free axis sync.
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
// its whole walk -- decided on the host from the eye (launch_raymarch picks the FAR instantiation; a shadow ray starts at a
// hit point inside the grid and is never far).  Result identical to oracle/orc_raymarch.c either way.
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

// first cell of a ray (walk_begin, true space) -> mirrored walk state
__device__ __forceinline__ bool walk10_begin(const DVolume& v, const Ray& r, int cx, int cy, int cz, Walk10& w, unsigned& steps) {
  const bool alive = walk_begin(v, r, cx, cy, cz, w.la, w.lt, steps);
  w.csx = cx ^ (r.sx >> 31); w.csy = cy ^ (r.sy >> 31); w.csz = cz ^ (r.sz >> 31);
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
template <bool STATS, int CL, bool FAR>
__global__ void __launch_bounds__(RM10_THREADS, RM10_MINB *(256 / RM10_THREADS)) raymarch10_kernel(DVolume v, MesoRaySetup rs, int width, int height, uint32_t flags,
                                                                   int rank, int world, int layout, int tiles_x, int n_tiles, int local_tile0,
                                                                   FrameMap fm, RayStatsDev* stats,
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
      park_inactive(m, w);
      res = walk10<STATS, CL, FAR>(v, sc, m, w, cx, cy, cz, steps);
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
        park_inactive(m, w);
        res = walk10<STATS, CL, false>(v, sc, m, w, cx, cy, cz, steps);
      }
      dn.shadow = res == W_HIT ? 1 : 0;
      n_shadow = 1;
    }
  }

  if (valid) {
    // MESO_LAYOUT_SLABS: the frame is cut into horizontal slabs of rows_per_slab scanlines, slab k living in fm.slab[k] --
    // another GPU's memory over NVLink for the rows this rank does not own (the all-to-all of the slab gather, fused into
    // the store: every record is written once, to where it will be copied to the host from)
    void* out = fm.slab[0];
    size_t dst;
    if (layout == MESO_LAYOUT_SLABS) {
      const int k = py / fm.rows_per_slab;
      out = fm.slab[k];
      dst = (size_t)(py - k * fm.rows_per_slab) * width + px;
    } else if (layout == MESO_LAYOUT_FRAME) dst = (size_t)py * width + px;
    else dst = (size_t)local_tile * (MESO_TILE_W * MESO_TILE_H) + ty * MESO_TILE_W + tx;
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
                     uint8_t* d_touch_brick, int local_tile0, int local_tile_count, const CubeTables* cubes, const FrameMap* slabs) {
  const int tiles_x = (width + MESO_TILE_W - 1) / MESO_TILE_W, tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  const int n_tiles = tiles_x * tiles_y;
  const int all_local = (n_tiles - rank + world - 1) / world;
  if (local_tile_count < 0) local_tile_count = all_local - local_tile0;
  if (local_tile_count <= 0) return;
  // `cubes` = the per-octant forward-cube tables (k_cubes.cu) when they are current, else null: the walk then reads the
  // distance field + probe-ahead (CL = 0; streaming updates, which would have to rebuild the tables every frame).
  // cubes_level: 1 = cell cubes, 2 = + brick cubes (default, fastest measured on B200), 3 = + 2^3-cell cubes.
  const CubeTables ct = cubes ? *cubes : CubeTables{};
  FrameMap fm{};
  if (slabs) fm = *slabs; else { fm.slab[0] = d_out; fm.rows_per_slab = 1 << 30; fm.n_slabs = 1; }
  // an eye farther than RM10_FAR voxels from the grid's corner: loop-form axis sync for the primary rays, over the distance field
  const bool far = fmaxf(fmaxf(fabsf(rs.o[0]), fabsf(rs.o[1])), fabsf(rs.o[2])) > RM10_FAR;
  const int cl = (cubes && !far) ? (cubes->cell2 ? 3 : (cubes->brick ? 2 : 1)) : 0;
#define RM10_LAUNCH(ST, CL, FR)                                                                                                    \
  raymarch10_kernel<ST, CL, FR><<<local_tile_count * RM10_SPLIT, RM10_THREADS, 0, lc.stream>>>(v, rs, width, height, flags, rank, world, layout, \
                                                                               tiles_x, n_tiles, local_tile0, fm, d_stats,       \
                                                                               d_touch_chunk, d_touch_brick, ct)
  if (far)          { if (d_stats) RM10_LAUNCH(true, 0, true); else RM10_LAUNCH(false, 0, true); }
  else if (d_stats) { if (cl == 0) RM10_LAUNCH(true, 0, false); else if (cl == 1) RM10_LAUNCH(true, 1, false); else if (cl == 2) RM10_LAUNCH(true, 2, false); else RM10_LAUNCH(true, 3, false); }
  else              { if (cl == 0) RM10_LAUNCH(false, 0, false); else if (cl == 1) RM10_LAUNCH(false, 1, false); else if (cl == 2) RM10_LAUNCH(false, 2, false); else RM10_LAUNCH(false, 3, false); }
#undef RM10_LAUNCH
  (*lc.launches)++;
}

// colour words of n records, packed: what travels to the host when only the image is wanted but the records stay useful on
// the device (picking); 16 B read + 4 B written per pixel
__global__ void __launch_bounds__(256) pack_rgba8_kernel(const uint4* __restrict__ rec, size_t n, uint32_t* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = rec[i].w;
}
void launch_pack_rgba8(const LaunchCtx& lc, const MesoHitRecord* d_records, size_t n, uint32_t* d_out) {
  if (n == 0) return;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)lc.sm_count * 16);
  pack_rgba8_kernel<<<grid, 256, 0, lc.stream>>>(reinterpret_cast<const uint4*>(d_records), n, d_out);
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
