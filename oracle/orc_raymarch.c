/* orc_raymarch.c -- CPU ORACLE (test infrastructure): per-pixel visibility.
 *
 * (1) orc_ref_instanced_pixel restates what the reference's instanced draw resolves per pixel
 *     (Samples/SimpleVoxel.cpp:146-192 VS, :220-224 FS, depth Greater/clear 0 at :322,:382, fan faces per
 *     Runtimes/Shape/TriplePlanarCube.h:36-43 + SimpleVoxel.cpp:87-136) as "nearest of the three camera-facing
 *     faces of every valid instance along the pixel-centre ray", in fp64.
 * (2) orc_raymarch is the repo-DEFINED 3D-DDA ("parity unpinned by reference; bit-exact vs repo oracle").
 *     The DDA is STATELESS: the crossing time of an integer voxel plane p on axis a is always
 *         t_a(p) = (float(p) - o_a) * inv_a            (one fp32 sub, one fp32 mul, no fma)
 *     and crossings are consumed in the total order key = (t, axis).  ORC_DDA_FLAT walks one voxel per step on
 *     an unbounded grid and is the definition; ORC_DDA_HIER skips empty chunks (128^3) / empty bricks (8^3) and
 *     re-derives the other two coordinates from the same keys, so it is bit-identical to FLAT by construction
 *     (tests/test_oracle_raymarch.py checks it).  ORC_DDA_BOX additionally leaves whole UNALIGNED cubes of empty
 *     32^3 cells in one step (cubes from a brute-force Chebyshev distance over cells, cap 8): the kind of skip the
 *     CUDA kernel makes (its own field: separable passes, cap 32, probe-ahead).  FLAT == HIER == BOX, byte for byte,
 *     is the CPU-side evidence that the records do not depend on which empty boxes a walk skips.
 */
#include "orc_internal.h"

/* optional instrumentation: steps taken at voxel / brick / chunk level (DESIGN.md "where the steps go") */
static uint64_t* orc_debug_level_hist = NULL;
void orc_debug_set_level_hist(uint64_t* hist3) { orc_debug_level_hist = hist3; }

/* ---- step-count MODEL of acceleration structures (ORC_DDA_MODEL; instrumentation, never part of parity) ----------
 * Any certified-empty box is a legal skip, so one walker can stand in for candidate kernel designs: it asks a
 * configurable hierarchy for a box around the current voxel, leaves it, and counts the steps by kind.  The records it
 * produces are byte-identical to ORC_DDA_FLAT / HIER (checked in tests/test_oracle_raymarch.py), only the counts
 * differ.  tools/step_model.py calibrates it against the step counts measured on the GPU (profiles/) and evaluates
 * variants (finer field, per-octant forward cubes, brick-level cubes) before any GPU time is spent on them. */
typedef struct {
  int df_shift;      /* log2 of the distance-field cell in voxels: 5 = 32^3 (shipped kernel), 4 = 16^3                   */
  int df_cap;        /* cap of the field in cells (shipped: 32)                                                         */
  int probe;         /* 1: second lookup along the octant diagonal as the shipped kernel does                            */
  int directional;   /* 1: per-octant forward-cube field (largest empty cube starting at the cell, towards the octant)   */
  int brick_cap;     /* > 1: forward cubes of empty bricks at the brick level, up to this many bricks per step           */
  int cell2;         /* 1: 2^3-voxel cells inside partial bricks (shipped); t > 1: forward cubes of up to t empty cells,     */
                     /*    confined to the brick (what a per-brick table of 64 cells x 8 octants x 2 bits would certify)     */
} OrcStepModel;
static OrcStepModel g_model = {5, 32, 1, 0, 0, 1};
static uint64_t g_model_steps[6];   /* 0 voxel, 1 2^3 cell, 2 brick, 3 field step <= 2 cells, 4 field step > 2 cells, 5 entry */
static uint8_t* g_model_sym = NULL;       /* symmetric field, dd cells                      */
static uint8_t* g_model_fwd = NULL;       /* forward-cube field, 8 octants x dd cells       */
static int g_model_dd[3] = {0, 0, 0};
/* table-driven variant: the three tables of csrc/k_cubes.cu, built by orc_cube_tables with the GPU's algorithms */
static const uint8_t* g_tab_cell = NULL; static const uint16_t* g_tab_brick = NULL; static const uint16_t* g_tab_cell2 = NULL;
static int64_t g_tab_ncells = 0;

/* optional instrumentation: steps of every pixel's primary and shadow ray (2 x u32 per pixel, row-major full frame) --
 * lets tools/step_model.py apply the kernel's warp shape (8x4 pixels) and predict warp iterations, not only steps */
static uint32_t* g_step_image = NULL; static int g_step_image_w = 0;
void orc_debug_set_step_image(uint32_t* img, int width) { g_step_image = img; g_step_image_w = width; }

typedef struct { float o[3], d[3], inv[3]; int step[3]; } Ray;
typedef struct { int hit; int c[3]; int axis; float t; uint64_t steps; } Trace;

typedef struct {
  const OrcVolume* v;
  int n[3];               /* grid size in voxels */
  const uint8_t* chunk_any;
  uint8_t* touched_chunk; /* may be NULL */
  uint8_t* touched_brick; /* indexed by payload slot; may be NULL */
  const uint8_t* df;      /* ORC_DDA_BOX only: per 32^3 cell, Chebyshev distance in cells to the nearest non-empty cell */
  int dd[3];              /* cells per axis */
} Scene;

static inline void ray_init(Ray* r, const float o[3], const float d[3]) {
  for (int i = 0; i < 3; i++) {
    r->o[i] = o[i]; r->d[i] = d[i];
    if (fabsf(d[i]) >= 1e-20f) { r->inv[i] = 1.0f / d[i]; r->step[i] = d[i] > 0.0f ? 1 : -1; }
    else { r->inv[i] = 0.0f; r->step[i] = 0; }
  }
}
static inline float plane_t(const Ray* r, int a, int plane) { return ((float)plane - r->o[a]) * r->inv[a]; }
static inline int key_less(float t1, int a1, float t2, int a2) { return t1 < t2 || (t1 == t2 && a1 < a2); }

/* coordinate on axis b after consuming every crossing with key < (ts, as), starting from cell coordinate cur */
static int advance_axis(const Ray* r, int b, int cur, float ts, int as) {
  int st = r->step[b];
  if (st == 0) return cur;
  float pos = r->o[b] + r->d[b] * ts;
  float fl = floorf(pos);
  if (fl > 1.0e9f) fl = 1.0e9f;
  if (fl < -1.0e9f) fl = -1.0e9f;
  int e = (int)fl;
  if (st > 0) { if (e < cur) e = cur; } else { if (e > cur) e = cur; }
  for (;;) {
    int pa = st > 0 ? e + 1 : e;
    if (key_less(plane_t(r, b, pa), b, ts, as)) e += st; else break;
  }
  for (;;) {
    if (e == cur) break;
    int pb = st > 0 ? e : e + 1;
    if (!key_less(plane_t(r, b, pb), b, ts, as)) e -= st; else break;
  }
  return e;
}

static inline int inside(const Scene* s, const int c[3]) {
  return c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[0] < s->n[0] && c[1] < s->n[1] && c[2] < s->n[2];
}
/* the ray can never (re-)enter the grid from cell c */
static inline int gone(const Scene* s, const Ray* r, const int c[3]) {
  for (int i = 0; i < 3; i++) {
    if (r->step[i] > 0) { if (c[i] >= s->n[i]) return 1; }
    else if (r->step[i] < 0) { if (c[i] < 0) return 1; }
    else if (c[i] < 0 || c[i] >= s->n[i]) return 1;
  }
  return 0;
}

/* level of the cell at voxel c (inside the grid): 0 = solid voxel, 1 = empty voxel in a partial brick,
 * 8 = empty brick, 128 = empty chunk */
static inline int cell_level(const Scene* s, const int c[3]) {
  const OrcVolume* v = s->v;
  int64_t ci = orc_cidx(v, c[0] >> 7, c[1] >> 7, c[2] >> 7);
  if (!s->chunk_any[ci]) return ORC_CV;
  if (s->touched_chunk && !s->touched_chunk[ci]) {
    /* A chunk's block masks count as needed once a ray stands in one of its non-empty 32^3 cells (4x4x4 bricks): the
     * unit the GPU walk resolves before it reads them.  occ word w = bz*4 + by/4 holds four 16-bit x-rows. */
    int ex = (c[0] >> 5) & 3, ey = (c[1] >> 5) & 3, ez = (c[2] >> 5) & 3;
    uint64_t m = 0;
    for (int bz = 4 * ez; bz < 4 * ez + 4; bz++) m |= v->occ[ci * ORC_WORDS + bz * 4 + ey] & (0x000F000F000F000Full << (4 * ex));
    if (m) s->touched_chunk[ci] = 1;
  }
  int b = orc_bidx((c[0] >> 3) & 15, (c[1] >> 3) & 15, (c[2] >> 3) & 15);
  if (!orc_getbit(v->occ + ci * ORC_WORDS, b)) return ORC_BR;
  if (orc_getbit(v->full + ci * ORC_WORDS, b)) return 0;
  uint32_t slot = v->bptr[ci][b];
  if (s->touched_brick) s->touched_brick[slot] = 1;
  const uint64_t* p = v->pool + (size_t)slot * 8;
  return (int)((p[c[2] & 7] >> ((c[0] & 7) + 8 * (c[1] & 7))) & 1u) ? 0 : 1;
}

/* leave the axis-aligned box [lo, hi) (voxel planes, already clamped to the grid) that contains c; returns 0 if the ray has no direction */
static int leave_box(const Ray* r, int c[3], const int lo[3], const int hi[3], Trace* tr) {
  int a = -1; float ta = 0.0f; int pl_a = 0;
  for (int i = 0; i < 3; i++) {
    if (!r->step[i]) continue;
    int pl = r->step[i] > 0 ? hi[i] : lo[i];
    float ti = plane_t(r, i, pl);
    if (a < 0 || key_less(ti, i, ta, a)) { a = i; ta = ti; pl_a = pl; }
  }
  if (a < 0) return 0;
  c[a] = r->step[a] > 0 ? pl_a : pl_a - 1;
  for (int b = 0; b < 3; b++) if (b != a) c[b] = advance_axis(r, b, c[b], ta, a);
  tr->axis = a; tr->t = ta; tr->steps++;
  return 1;
}

static int brick_occupied(const Scene* s, int bx, int by, int bz) {   /* brick coordinates in the grid; outside = empty */
  const OrcVolume* v = s->v;
  if (bx < 0 || by < 0 || bz < 0 || bx >= v->dims[0] * 16 || by >= v->dims[1] * 16 || bz >= v->dims[2] * 16) return 0;
  int64_t ci = orc_cidx(v, bx >> 4, by >> 4, bz >> 4);
  return orc_getbit(v->occ + ci * ORC_WORDS, orc_bidx(bx & 15, by & 15, bz & 15));
}

static void model_walk(const Scene* s, const Ray* r, int c[3], Trace* tr) {
  const OrcStepModel* m = &g_model;
  const int sh = m->df_shift;
  const int oct = (r->step[0] < 0 ? 1 : 0) | (r->step[1] < 0 ? 2 : 0) | (r->step[2] < 0 ? 4 : 0);
  const size_t ncell = (size_t)g_model_dd[0] * g_model_dd[1] * g_model_dd[2];
  for (;;) {
    /* field level */
    const int e[3] = {c[0] >> sh, c[1] >> sh, c[2] >> sh};
    const size_t ei = (size_t)e[0] + (size_t)g_model_dd[0] * ((size_t)e[1] + (size_t)g_model_dd[1] * (size_t)e[2]);
    int k = g_tab_cell ? g_tab_cell[(size_t)oct * (size_t)g_tab_ncells + ei] : (m->directional ? g_model_fwd[(size_t)oct * ncell + ei] : g_model_sym[ei]);
    if (k > 0) {
      if (!g_tab_cell && !m->directional && m->probe) {
        int q[3]; int in = 1;
        for (int i = 0; i < 3; i++) { q[i] = e[i] + (r->step[i] < 0 ? -k : k); if (q[i] < 0 || q[i] >= g_model_dd[i]) in = 0; }
        if (in) { int d2 = g_model_sym[(size_t)q[0] + (size_t)g_model_dd[0] * ((size_t)q[1] + (size_t)g_model_dd[1] * (size_t)q[2])]; if (d2 > k) k += d2; }
      }
      int lo[3], hi[3];
      for (int i = 0; i < 3; i++) {
        hi[i] = (e[i] + k) << sh; if (hi[i] > s->n[i]) hi[i] = s->n[i];
        lo[i] = (e[i] - k + 1) << sh; if (lo[i] < 0) lo[i] = 0;
      }
      if (!leave_box(r, c, lo, hi, tr)) return;
      __atomic_fetch_add(&g_model_steps[k <= 2 ? 3 : 4], 1, __ATOMIC_RELAXED);
      if (!inside(s, c)) return;
      continue;
    }
    int L = cell_level(s, c);
    if (L == 0) { tr->hit = 1; return; }
    if (L >= ORC_BR) {
      /* empty brick: aligned brick step, or the largest empty cube of bricks towards the octant (direct test, cap brick_cap) */
      int kb = 1;
      const int b[3] = {c[0] >> 3, c[1] >> 3, c[2] >> 3};
      if (g_tab_brick) {
        const int64_t ci = orc_cidx(s->v, b[0] >> 4, b[1] >> 4, b[2] >> 4);
        kb = 1 + ((g_tab_brick[(size_t)ci * ORC_BLOCKS + orc_bidx(b[0] & 15, b[1] & 15, b[2] & 15)] >> (2 * oct)) & 3);
      } else
      for (int t = 2; t <= m->brick_cap; t++) {
        int empty = 1;
        for (int z = 0; z < t && empty; z++) for (int y = 0; y < t && empty; y++) for (int x = 0; x < t; x++) {
          if (x < t - 1 && y < t - 1 && z < t - 1) continue;   /* the (t-1)-cube was tested already */
          if (brick_occupied(s, b[0] + (r->step[0] < 0 ? -x : x), b[1] + (r->step[1] < 0 ? -y : y), b[2] + (r->step[2] < 0 ? -z : z))) { empty = 0; break; }
        }
        if (!empty) break;
        kb = t;
      }
      int lo[3], hi[3];
      for (int i = 0; i < 3; i++) {
        hi[i] = (b[i] + kb) << 3; if (hi[i] > s->n[i]) hi[i] = s->n[i];
        lo[i] = (b[i] - kb + 1) << 3; if (lo[i] < 0) lo[i] = 0;
      }
      if (!leave_box(r, c, lo, hi, tr)) return;
      __atomic_fetch_add(&g_model_steps[2], 1, __ATOMIC_RELAXED);
      if (!inside(s, c)) return;
      continue;
    }
    /* empty voxel inside a partial brick */
    int size = 1;
    if (m->cell2) {
      const OrcVolume* v = s->v;
      int64_t ci = orc_cidx(v, c[0] >> 7, c[1] >> 7, c[2] >> 7);
      uint32_t slot = v->bptr[ci][orc_bidx((c[0] >> 3) & 15, (c[1] >> 3) & 15, (c[2] >> 3) & 15)];
      const uint64_t* p = v->pool + (size_t)slot * 8;
      const int x0 = c[0] & 6, y0 = c[1] & 6, z0 = c[2] & 6;
      uint64_t mask = (3ull << (x0 + 8 * y0)) | (3ull << (x0 + 8 * (y0 + 1)));
      if (((p[z0] | p[z0 + 1]) & mask) == 0) size = 2;
      if (size == 2 && g_tab_cell2) {
        const int ce = ((c[0] >> 1) & 3) + 4 * ((c[1] >> 1) & 3) + 16 * ((c[2] >> 1) & 3);
        const int kc = 1 + ((g_tab_cell2[(size_t)slot * 64 + ce] >> (2 * oct)) & 3);
        if (kc > 1) {
          int lo2[3], hi2[3];
          for (int i = 0; i < 3; i++) { const int base = c[i] & ~1; hi2[i] = base + 2 * kc; lo2[i] = base + 2 - 2 * kc; }
          if (!leave_box(r, c, lo2, hi2, tr)) return;
          __atomic_fetch_add(&g_model_steps[1], 1, __ATOMIC_RELAXED);
          if (!inside(s, c)) return;
          continue;
        }
      } else if (size == 2 && m->cell2 > 1) {
        /* largest cube of empty 2^3 cells starting at this cell towards the octant, inside the brick */
        const int q0[3] = {(c[0] & 7) >> 1, (c[1] & 7) >> 1, (c[2] & 7) >> 1};
        int kc = 1;
        for (int t = 2; t <= m->cell2; t++) {
          int ok = 1;
          for (int z = 0; z < t && ok; z++) for (int y = 0; y < t && ok; y++) for (int x = 0; x < t; x++) {
            if (x < t - 1 && y < t - 1 && z < t - 1) continue;
            const int qx = q0[0] + (r->step[0] < 0 ? -x : x), qy = q0[1] + (r->step[1] < 0 ? -y : y), qz = q0[2] + (r->step[2] < 0 ? -z : z);
            if (qx < 0 || qy < 0 || qz < 0 || qx > 3 || qy > 3 || qz > 3) { ok = 0; break; }
            const uint64_t mk = (3ull << (2 * qx + 16 * qy)) | (3ull << (2 * qx + 16 * qy + 8));
            if ((p[2 * qz] | p[2 * qz + 1]) & mk) { ok = 0; break; }
          }
          if (!ok) break;
          kc = t;
        }
        if (kc > 1) {
          int lo2[3], hi2[3];
          for (int i = 0; i < 3; i++) {
            const int base = c[i] & ~1;
            hi2[i] = base + 2 * kc; lo2[i] = base + 2 - 2 * kc;
          }
          if (!leave_box(r, c, lo2, hi2, tr)) return;
          __atomic_fetch_add(&g_model_steps[1], 1, __ATOMIC_RELAXED);
          if (!inside(s, c)) return;
          continue;
        }
      }
    }
    int lo[3], hi[3];
    for (int i = 0; i < 3; i++) { lo[i] = c[i] & ~(size - 1); hi[i] = lo[i] + size; }
    if (!leave_box(r, c, lo, hi, tr)) return;
    __atomic_fetch_add(&g_model_steps[size == 2 ? 1 : 0], 1, __ATOMIC_RELAXED);
    if (!inside(s, c)) return;
  }
}

static Trace trace(const Scene* s, const Ray* r, const int c0[3], int mode) {
  Trace tr; tr.hit = 0; tr.axis = -1; tr.t = 0.0f; tr.steps = 0;
  int c[3] = {c0[0], c0[1], c0[2]};
  if (mode == ORC_DDA_FLAT) {
    for (;;) {
      if (inside(s, c)) { if (cell_level(s, c) == 0) { tr.hit = 1; break; } }
      else if (gone(s, r, c)) break;
      int a = -1; float ta = 0.0f;
      for (int i = 0; i < 3; i++) {
        if (!r->step[i]) continue;
        float ti = plane_t(r, i, r->step[i] > 0 ? c[i] + 1 : c[i]);
        if (a < 0 || key_less(ti, i, ta, a)) { a = i; ta = ti; }
      }
      if (a < 0) break; /* zero direction */
      c[a] += r->step[a]; tr.axis = a; tr.t = ta; tr.steps++;
    }
  } else {
    if (!inside(s, c)) {
      if (gone(s, r, c)) goto done;
      int a = -1; float ta = 0.0f;
      for (int i = 0; i < 3; i++) {
        int before = (r->step[i] > 0 && c[i] < 0) || (r->step[i] < 0 && c[i] >= s->n[i]);
        if (!before) continue;
        float ti = plane_t(r, i, r->step[i] > 0 ? 0 : s->n[i]);
        if (a < 0 || key_less(ta, a, ti, i)) { a = i; ta = ti; }
      }
      c[a] = r->step[a] > 0 ? 0 : s->n[a] - 1;
      for (int b = 0; b < 3; b++) if (b != a) c[b] = advance_axis(r, b, c[b], ta, a);
      tr.axis = a; tr.t = ta; tr.steps++;
      if (!inside(s, c)) goto done;
    }
    if (mode == ORC_DDA_MODEL) { if (tr.steps) __atomic_fetch_add(&g_model_steps[5], 1, __ATOMIC_RELAXED); model_walk(s, r, c, &tr); goto done; }
    for (;;) {
      int L = cell_level(s, c);
      if (L == 0) { tr.hit = 1; break; }
      if (mode == ORC_DDA_BOX && L >= ORC_BR) {
        /* Empty brick or chunk: if the distance field certifies an empty cube of cells around this 32^3 cell, leave the
         * whole cube (clamped to the grid) in one step -- an UNALIGNED box, unlike the aligned cells of ORC_DDA_HIER.
         * Same keys, same order: any empty box is a legal skip. */
        int k = s->df[(c[0] >> 5) + s->dd[0] * ((c[1] >> 5) + s->dd[1] * (c[2] >> 5))];
        if (k > 0) {
          int a = -1; float ta = 0.0f; int pl_a = 0;
          for (int i = 0; i < 3; i++) {
            if (!r->step[i]) continue;
            int e = c[i] >> 5, pl;
            if (r->step[i] > 0) { pl = (e + k) * 32; if (pl > s->n[i]) pl = s->n[i]; }
            else { pl = (e - k + 1) * 32; if (pl < 0) pl = 0; }
            float ti = plane_t(r, i, pl);
            if (a < 0 || key_less(ti, i, ta, a)) { a = i; ta = ti; pl_a = pl; }
          }
          if (a < 0) break;
          c[a] = r->step[a] > 0 ? pl_a : pl_a - 1;
          for (int b = 0; b < 3; b++) if (b != a) c[b] = advance_axis(r, b, c[b], ta, a);
          tr.axis = a; tr.t = ta; tr.steps++;
          if (!inside(s, c)) break;
          continue;
        }
      }
      int a = -1; float ta = 0.0f; int pl_a = 0;
      for (int i = 0; i < 3; i++) {
        if (!r->step[i]) continue;
        int base = c[i] & ~(L - 1);
        int pl = r->step[i] > 0 ? base + L : base;
        float ti = plane_t(r, i, pl);
        if (a < 0 || key_less(ti, i, ta, a)) { a = i; ta = ti; pl_a = pl; }
      }
      if (a < 0) break;
      if (L == 1) c[a] += r->step[a];
      else {
        c[a] = r->step[a] > 0 ? pl_a : pl_a - 1;
        for (int b = 0; b < 3; b++) if (b != a) c[b] = advance_axis(r, b, c[b], ta, a);
      }
      tr.axis = a; tr.t = ta; tr.steps++;
      if (orc_debug_level_hist) __atomic_fetch_add(&orc_debug_level_hist[L == 1 ? 0 : (L == ORC_BR ? 1 : 2)], 1, __ATOMIC_RELAXED);
      if (!inside(s, c)) break;
    }
  }
done:
  tr.c[0] = c[0]; tr.c[1] = c[1]; tr.c[2] = c[2];
  return tr;
}

static inline uint32_t to_un8(float x) { return (uint32_t)(x * 255.0f + 0.5f); }

typedef struct {
  Scene sc; const OrcRaySetup* rs; int width, x0, x1, y0; const int32_t* rows; uint32_t flags; int mode; OrcHitRecord* rec;
  uint64_t primary, shadow, hits, steps; pthread_mutex_t mu;
} RmArg;

static void shade_pixel(RmArg* a, int px, int py, uint64_t cnt[4]) {
  const OrcRaySetup* rs = a->rs;
  float fx = ((float)px + 0.5f) * rs->two_over_w - 1.0f;
  float fy = 1.0f - ((float)py + 0.5f) * rs->two_over_h;
  float d[3];
  for (int i = 0; i < 3; i++) d[i] = (fx * rs->U[i] + fy * rs->V[i]) + rs->F[i];
  float len = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  for (int i = 0; i < 3; i++) d[i] = d[i] / len;
  Ray r; ray_init(&r, rs->o, d);
  int c0[3];
  for (int i = 0; i < 3; i++) { float f = floorf(rs->o[i]); if (f > 1.0e9f) f = 1.0e9f; if (f < -1.0e9f) f = -1.0e9f; c0[i] = (int)f; }
  Trace tr = trace(&a->sc, &r, c0, a->mode);
  cnt[0]++; cnt[3] += tr.steps;
  if (g_step_image) { g_step_image[2 * ((size_t)py * g_step_image_w + px)] = (uint32_t)tr.steps; g_step_image[2 * ((size_t)py * g_step_image_w + px) + 1] = 0; }
  OrcHitRecord* out = &a->rec[(size_t)py * a->width + px];
  if (!tr.hit) { out->w0 = 0xFFFFFFFFu; out->w1 = 0x0007FFFFu; out->t = INFINITY; out->rgba = 0xFF000000u; return; }
  cnt[2]++;
  int face = 6; float p[3]; int shadow = 0;
  for (int i = 0; i < 3; i++) p[i] = r.o[i] + r.d[i] * tr.t;
  if (tr.axis >= 0) {
    int ax = tr.axis;
    face = ax * 2 + (r.step[ax] > 0 ? 0 : 1);
    p[ax] = (float)(r.step[ax] > 0 ? tr.c[ax] : tr.c[ax] + 1);
    if (a->flags & ORC_FLAG_SHADOW) {
      int facing = r.step[ax] > 0 ? (rs->L[ax] < 0.0f) : (rs->L[ax] > 0.0f);
      if (!facing) shadow = 1;
      else {
        Ray sr; ray_init(&sr, p, rs->L);
        int sc0[3] = {tr.c[0], tr.c[1], tr.c[2]};
        sc0[ax] -= r.step[ax];
        Trace st = trace(&a->sc, &sr, sc0, a->mode);
        cnt[1]++; cnt[3] += st.steps;
        if (g_step_image) g_step_image[2 * ((size_t)py * g_step_image_w + px) + 1] = (uint32_t)st.steps;
        shadow = st.hit;
      }
    }
  } else { for (int i = 0; i < 3; i++) p[i] = r.o[i]; }
  float shade = shadow ? 0.5f : 1.0f;
  uint32_t ch[3];
  for (int i = 0; i < 3; i++) {
    float local = p[i] * 0.125f - (float)(tr.c[i] >> 3);
    if (local < 0.0f) local = 0.0f;
    if (local > 1.0f) local = 1.0f;
    float col = (local - 0.5f) * 0.5f + 0.5f;   /* SimpleVoxel.cpp:222  Normal*0.5+0.5, Normal = local-0.5 */
    ch[i] = to_un8(col * shade);
  }
  out->w0 = (uint32_t)tr.c[0] | ((uint32_t)tr.c[1] << 16);
  out->w1 = (uint32_t)tr.c[2] | ((uint32_t)face << 16) | ((uint32_t)shadow << 19) | (1u << 20);
  out->t = tr.t;
  out->rgba = ch[0] | (ch[1] << 8) | (ch[2] << 16); /* alpha 0 on hits (SimpleVoxel.cpp:222), 1 on clear (:315) */
}

/* work item = RM_SPAN consecutive pixels of one sampled row: fine enough that a handful of rows keeps every thread busy */
#define RM_SPAN 128
static void rm_range(void* ctx, int64_t b, int64_t e, int tid) {
  (void)tid;
  RmArg* a = (RmArg*)ctx;
  uint64_t cnt[4] = {0, 0, 0, 0};
  const int spans = (a->x1 - a->x0 + RM_SPAN - 1) / RM_SPAN;
  for (int64_t it = b; it < e; it++) {
    const int row = a->rows ? a->rows[it / spans] : a->y0 + (int)(it / spans);
    const int xa = a->x0 + (int)(it % spans) * RM_SPAN;
    const int xb = xa + RM_SPAN < a->x1 ? xa + RM_SPAN : a->x1;
    for (int px = xa; px < xb; px++) shade_pixel(a, px, row, cnt);
  }
  pthread_mutex_lock(&a->mu);
  a->primary += cnt[0]; a->shadow += cnt[1]; a->hits += cnt[2]; a->steps += cnt[3];
  pthread_mutex_unlock(&a->mu);
}

static const uint8_t* chunk_any_cached(const OrcVolume* cv) {
  OrcVolume* v = (OrcVolume*)cv;   /* the cache is not part of the volume's value */
  pthread_mutex_lock(&v->lock);
  if (!v->any_valid) {
    if (!v->any) v->any = (uint8_t*)malloc((size_t)v->nchunks);
    for (int64_t c = 0; c < v->nchunks; c++) {
      uint64_t o = 0;
      for (int w = 0; w < ORC_WORDS; w++) o |= v->occ[c * ORC_WORDS + w];
      v->any[c] = o != 0;
    }
    v->any_valid = 1;
  }
  pthread_mutex_unlock(&v->lock);
  return v->any;
}

static void raymarch_impl(const OrcVolume* v, const OrcRaySetup* rs, int width, int x0, int y0, int x1, int nrows, const int32_t* rows,
                          uint32_t flags, int mode, int nthreads, OrcHitRecord* records, OrcRayStats* stats);

void orc_raymarch(const OrcVolume* v, const OrcRaySetup* rs, int width, int height, int x0, int y0, int x1, int y1,
                  uint32_t flags, int mode, int nthreads, OrcHitRecord* records, OrcRayStats* stats) {
  (void)height;
  raymarch_impl(v, rs, width, x0, y0, x1, y1 - y0, NULL, flags, mode, nthreads, records, stats);
}
/* the same for an explicit list of scanlines (bench.py's bounded CPU sample: one call, one parallel region) */
void orc_raymarch_rows(const OrcVolume* v, const OrcRaySetup* rs, int width, int height, const int32_t* rows, int nrows,
                       uint32_t flags, int mode, int nthreads, OrcHitRecord* records, OrcRayStats* stats) {
  (void)height;
  raymarch_impl(v, rs, width, 0, 0, width, nrows, rows, flags, mode, nthreads, records, stats);
}

static void raymarch_impl(const OrcVolume* v, const OrcRaySetup* rs, int width, int x0, int y0, int x1, int nrows, const int32_t* rows,
                          uint32_t flags, int mode, int nthreads, OrcHitRecord* records, OrcRayStats* stats) {
  RmArg a; memset(&a, 0, sizeof(a));
  a.sc.v = v;
  for (int i = 0; i < 3; i++) a.sc.n[i] = v->dims[i] * ORC_CV;
  a.sc.chunk_any = chunk_any_cached(v);
  uint8_t* df = NULL;
  if (mode == ORC_DDA_BOX) {
    /* cell occupancy from the block masks, then a brute-force Chebyshev distance (cap 8 cells): deliberately not the
     * device's algorithm, cap or data -- the point of this mode is that the records do not depend on the boxes */
    const int dd0 = v->dims[0] * 4, dd1 = v->dims[1] * 4, dd2 = v->dims[2] * 4, CAP = 8;
    const size_t nc = (size_t)dd0 * dd1 * dd2;
    uint8_t* occ_cell = (uint8_t*)calloc(nc, 1);
    df = (uint8_t*)calloc(nc, 1);
    for (int ez = 0; ez < dd2; ez++) for (int ey = 0; ey < dd1; ey++) for (int ex = 0; ex < dd0; ex++) {
      const int64_t ci = orc_cidx(v, ex >> 2, ey >> 2, ez >> 2);
      uint64_t m = 0;
      for (int bz = 4 * (ez & 3); bz < 4 * (ez & 3) + 4; bz++)
        m |= v->occ[ci * ORC_WORDS + bz * 4 + (ey & 3)] & (0x000F000F000F000Full << (4 * (ex & 3)));
      occ_cell[ex + (size_t)dd0 * (ey + (size_t)dd1 * ez)] = m != 0;
    }
    for (int ez = 0; ez < dd2; ez++) for (int ey = 0; ey < dd1; ey++) for (int ex = 0; ex < dd0; ex++) {
      const size_t i = ex + (size_t)dd0 * (ey + (size_t)dd1 * ez);
      int k = 0;
      if (!occ_cell[i]) {
        for (k = 1; k <= CAP; k++) {   /* k = smallest Chebyshev radius with a non-empty cell inside the cube */
          int found = 0;
          for (int z = ez - k; z <= ez + k && !found; z++) for (int y = ey - k; y <= ey + k && !found; y++) for (int x = ex - k; x <= ex + k; x++) {
            if (x < 0 || y < 0 || z < 0 || x >= dd0 || y >= dd1 || z >= dd2) continue;
            if (occ_cell[x + (size_t)dd0 * (y + (size_t)dd1 * z)]) { found = 1; break; }
          }
          if (found) break;
        }
        if (k > CAP) k = CAP + 1;
      }
      df[i] = (uint8_t)k;
    }
    free(occ_cell);
    a.sc.df = df; a.sc.dd[0] = dd0; a.sc.dd[1] = dd1; a.sc.dd[2] = dd2;
  }
  if (stats) {
    a.sc.touched_chunk = (uint8_t*)calloc((size_t)v->nchunks, 1);
    a.sc.touched_brick = (uint8_t*)calloc((size_t)(v->pool_n > 0 ? v->pool_n : 1), 1);
  }
  a.rs = rs; a.width = width; a.x0 = x0; a.x1 = x1; a.y0 = y0; a.rows = rows; a.flags = flags; a.mode = mode; a.rec = records;
  pthread_mutex_init(&a.mu, NULL);
  orc_parallel_for((int64_t)nrows * ((x1 - x0 + RM_SPAN - 1) / RM_SPAN), nthreads, 1, rm_range, &a);
  pthread_mutex_destroy(&a.mu);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->primary = a.primary; stats->shadow = a.shadow; stats->hits = a.hits; stats->steps = a.steps;
    for (int64_t c = 0; c < v->nchunks; c++) stats->touched_chunks += a.sc.touched_chunk[c];
    for (int64_t i = 0; i < v->pool_n; i++) stats->touched_bricks += a.sc.touched_brick[i];
    /* DESIGN.md "algorithmic bytes": chunk-level any/full bit grids once + 1024 B of block masks per touched
     * chunk + (4 B pointer + 64 B payload) per touched partial brick */
    stats->u_bytes = 2 * (uint64_t)((v->nchunks + 7) / 8) + 1024 * stats->touched_chunks + 68 * stats->touched_bricks;
    free(a.sc.touched_chunk); free(a.sc.touched_brick);
  }
  free(df);
}

/* ---- reference-semantics restatement of the instanced draw ------------------------------------ */

/* The fan 0,1,2,3,4,5,6,1 of table row `octant` (SimpleVoxel.cpp:87-127) covers exactly the three cube
 * faces that meet at the row's vertex 0, whose sign on axis i is + when bit i of the octant id is set
 * (tests/golden/triplanar_faces.json, generated from the reference table, pins this).  Face ids as in
 * OrcHitRecord: 2*axis + (positive side). */
void orc_triplanar_faces(int octant, int faces_out[3]) {
  for (int i = 0; i < 3; i++) faces_out[i] = 2 * i + ((octant >> i) & 1);
}

int orc_ref_instanced_pixel(const OrcGPUUniformCamera* cam, const OrcGPUUniformSceneConfig* scene,
                            const OrcGPUChunk* chunks, int64_t n_chunks, const OrcGPUBlock* blocks, int64_t n_blocks,
                            int width, int height, int px, int py, int32_t out_block[3], int* out_face,
                            double* out_t, double* out_margin, float out_rgba[4]) {
  const float* V = cam->View; const float* P = cam->Projection;
  double eye[3], d[3];
  double fx = ((double)px + 0.5) * 2.0 / (double)width - 1.0;
  double fy = 1.0 - ((double)py + 0.5) * 2.0 / (double)height;
  double vv[3] = {fx / (double)P[0], fy / (double)P[5], -1.0};
  for (int i = 0; i < 3; i++) {
    eye[i] = -((double)V[i * 4 + 0] * V[12] + (double)V[i * 4 + 1] * V[13] + (double)V[i * 4 + 2] * V[14]);
    d[i] = (double)V[i * 4 + 0] * vv[0] + (double)V[i * 4 + 1] * vv[1] + (double)V[i * 4 + 2] * vv[2];
  }
  double best_t = INFINITY, best_margin = 0.0; int best_face = 7; int32_t best_blk[3] = {0, 0, 0}; double best_local[3] = {0, 0, 0};
  int found = 0;
  for (int64_t k = 0; k < n_blocks; k++) {
    const OrcGPUBlock* B = &blocks[k];
    if (B->ChunkIndex == (uint32_t)INT_MAX || (int64_t)B->ChunkIndex >= n_chunks) continue;     /* SimpleVoxel.cpp:149 */
    const OrcGPUChunk* C = &chunks[B->ChunkIndex];
    if (C->ChunkFrameStamp != B->BlockFrameStamp) continue;                                     /* :152 */
    int32_t off[3]; float origin[3], rel[3];
    float inv_res = 1.0f / (float)scene->ChunkResolution;                                       /* :178 */
    int octant = 0;
    for (int i = 0; i < 3; i++) {
      off[i] = (C->ChunkLocation[i] - cam->CameraChunkLocation[i]) * (int32_t)scene->ChunkResolution + (int32_t)B->BlockLocation[i]; /* :166-177 */
      origin[i] = ((float)off[i] * inv_res) * scene->ChunkSize;                                 /* :179 */
      rel[i] = origin[i] - cam->SubCameraLocation[i];                                           /* :181 */
      if (rel[i] < 0.0f) octant |= 1 << i;                                                      /* :129-136 */
    }
    int faces[3]; orc_triplanar_faces(octant, faces);
    for (int f = 0; f < 3; f++) {
      int ax = faces[f] >> 1, pos = faces[f] & 1;
      float corner = pos ? 0.5f : -0.5f;
      float plane = (corner + 0.5f) * scene->BlockSize + origin[ax];                            /* :185 */
      if (d[ax] == 0.0) continue;
      double t = ((double)plane - eye[ax]) / d[ax];
      if (!(t > 0.0)) continue;
      double margin = INFINITY, local[3]; int ok = 1;
      for (int b = 0; b < 3; b++) {
        if (b == ax) { local[b] = pos ? 1.0 : 0.0; continue; }
        double q = eye[b] + d[b] * t;
        double lo = (double)origin[b], hi = (double)((-0.5f + 0.5f) * scene->BlockSize + origin[b]) + (double)scene->BlockSize;
        if (q < lo || q > hi) { ok = 0; break; }
        double m = fmin(q - lo, hi - q); if (m < margin) margin = m;
        local[b] = (q - lo) / (double)scene->BlockSize;
      }
      if (!ok) continue;
      if (t < best_t) {
        /* margin also accounts for how close the runner-up is (depth ties at shared edges) */
        best_margin = fmin(margin, found ? fabs(best_t - t) : INFINITY);
        best_t = t; best_face = faces[f]; found = 1;
        for (int b = 0; b < 3; b++) { best_blk[b] = off[b]; best_local[b] = local[b]; }
      } else {
        double gap = fabs(t - best_t); if (gap < best_margin) best_margin = gap;
      }
    }
  }
  if (!found) { out_rgba[0] = out_rgba[1] = out_rgba[2] = 0.0f; out_rgba[3] = 1.0f; return 0; }   /* clear: :315 */
  for (int b = 0; b < 3; b++) { out_block[b] = best_blk[b]; out_rgba[b] = (float)((best_local[b] - 0.5) * 0.5 + 0.5); }
  out_rgba[3] = 0.0f;
  *out_face = best_face; *out_t = best_t; *out_margin = best_margin;
  return 1;
}

/* ---- step-count model: fields and control (instrumentation; see the comment at the top of this file) --------------- */
void orc_step_model_config(int df_shift, int df_cap, int probe, int directional, int brick_cap, int cell2) {
  g_model.df_shift = df_shift; g_model.df_cap = df_cap; g_model.probe = probe; g_model.directional = directional;
  g_model.brick_cap = brick_cap; g_model.cell2 = cell2;
}
void orc_step_model_counts(uint64_t out[6], int reset) {
  for (int i = 0; i < 6; i++) { out[i] = g_model_steps[i]; if (reset) g_model_steps[i] = 0; }
}
/* Builds the field the current configuration needs over cells of 2^df_shift voxels (3 <= df_shift <= 7). */
int orc_step_model_build(const OrcVolume* v) {
  const int sh = g_model.df_shift, cap = g_model.df_cap;
  if (sh < 3 || sh > 7 || cap < 1 || cap > 255) return -1;
  const int bpc = 1 << (sh - 3);                        /* bricks per cell axis */
  int dd[3];
  for (int i = 0; i < 3; i++) dd[i] = (v->dims[i] * 16 + bpc - 1) / bpc;
  const size_t n = (size_t)dd[0] * dd[1] * dd[2];
  free(g_model_sym); free(g_model_fwd); g_model_sym = NULL; g_model_fwd = NULL;
  for (int i = 0; i < 3; i++) g_model_dd[i] = dd[i];
  uint8_t* occ = (uint8_t*)calloc(n, 1);
  if (!occ) return -1;
  for (int64_t c = 0; c < v->nchunks; c++) {
    const int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
    for (int w = 0; w < ORC_WORDS; w++) {
      uint64_t bits = v->occ[c * ORC_WORDS + w];
      while (bits) {
        const int b = w * 64 + __builtin_ctzll(bits); bits &= bits - 1;
        const int bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);
        occ[(size_t)(bx / bpc) + (size_t)dd[0] * ((size_t)(by / bpc) + (size_t)dd[1] * (size_t)(bz / bpc))] = 1;
      }
    }
  }
  if (!g_model.directional) {
    /* Chebyshev distance to the nearest occupied cell, capped: k-fold box dilation, one cell per round */
    uint8_t* dist = (uint8_t*)malloc(n); uint8_t* cur = (uint8_t*)malloc(n); uint8_t* tmp = (uint8_t*)malloc(n);
    if (!dist || !cur || !tmp) { free(occ); free(dist); free(cur); free(tmp); return -1; }
    for (size_t i = 0; i < n; i++) { dist[i] = occ[i] ? 0 : (uint8_t)cap; cur[i] = occ[i]; }
    for (int k = 1; k < cap; k++) {
      /* x */
      for (size_t row = 0; row < (size_t)dd[1] * dd[2]; row++) {
        const uint8_t* a = cur + row * dd[0]; uint8_t* o = tmp + row * dd[0];
        for (int x = 0; x < dd[0]; x++) o[x] = a[x] | (x > 0 ? a[x - 1] : 0) | (x + 1 < dd[0] ? a[x + 1] : 0);
      }
      /* y */
      for (int z = 0; z < dd[2]; z++) for (int y = 0; y < dd[1]; y++) {
        const size_t base = (size_t)dd[0] * ((size_t)y + (size_t)dd[1] * z);
        for (int x = 0; x < dd[0]; x++) {
          uint8_t r = tmp[base + x];
          if (y > 0) r |= tmp[base - dd[0] + x];
          if (y + 1 < dd[1]) r |= tmp[base + dd[0] + x];
          cur[base + x] = r;
        }
      }
      /* z */
      const size_t plane = (size_t)dd[0] * dd[1];
      int changed = 0;
      for (int z = 0; z < dd[2]; z++) for (size_t i = 0; i < plane; i++) {
        const size_t at = (size_t)z * plane + i;
        uint8_t r = cur[at];
        if (z > 0) r |= cur[at - plane];
        if (z + 1 < dd[2]) r |= cur[at + plane];
        tmp[at] = r;
        if (r && dist[at] == cap && !occ[at]) { dist[at] = (uint8_t)k; changed = 1; }
      }
      uint8_t* sw = cur; cur = tmp; tmp = sw;
      (void)changed;
    }
    free(cur); free(tmp);
    g_model_sym = dist;
  } else {
    uint8_t* fwd = (uint8_t*)malloc(8 * n);
    if (!fwd) { free(occ); return -1; }
    for (int o = 0; o < 8; o++) {
      const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
      uint8_t* f = fwd + (size_t)o * n;
      for (int zi = 0; zi < dd[2]; zi++) for (int yi = 0; yi < dd[1]; yi++) for (int xi = 0; xi < dd[0]; xi++) {
        /* visit forward neighbours first: walk against the octant direction */
        const int x = sx > 0 ? dd[0] - 1 - xi : xi, y = sy > 0 ? dd[1] - 1 - yi : yi, z = sz > 0 ? dd[2] - 1 - zi : zi;
        const size_t at = (size_t)x + (size_t)dd[0] * ((size_t)y + (size_t)dd[1] * (size_t)z);
        if (occ[at]) { f[at] = 0; continue; }
        int best = cap;
        for (int q = 1; q < 8; q++) {
          const int nx = x + ((q & 1) ? sx : 0), ny = y + ((q & 2) ? sy : 0), nz = z + ((q & 4) ? sz : 0);
          if (nx < 0 || ny < 0 || nz < 0 || nx >= dd[0] || ny >= dd[1] || nz >= dd[2]) continue;   /* beyond the grid: empty */
          const int fn = f[(size_t)nx + (size_t)dd[0] * ((size_t)ny + (size_t)dd[1] * (size_t)nz)];
          if (fn < best) best = fn;
        }
        f[at] = (uint8_t)(best + 1 > cap ? cap : best + 1);
      }
    }
    g_model_fwd = fwd;
  }
  free(occ);
  return 0;
}


/* ---- the tables of csrc/k_cubes.cu, built on the CPU with the GPU's algorithms (same layouts; cell2 indexed by the
 * oracle's own payload slots) -- reference data for the opt-in forward-cube path and a check of the builders' logic ---- */
int orc_cube_tables(const OrcVolume* v, uint8_t* cell, uint16_t* brick, uint16_t* cell2) {
  const int dd[3] = {v->dims[0] * 4, v->dims[1] * 4, v->dims[2] * 4};
  const int64_t n = (int64_t)dd[0] * dd[1] * dd[2];
  const int cap = 32;
  uint8_t* ne = (uint8_t*)calloc((size_t)n, 1);   /* cell not empty */
  if (!ne) return -1;
  for (int z = 0; z < dd[2]; z++) for (int y = 0; y < dd[1]; y++) for (int x = 0; x < dd[0]; x++) {
    const int64_t ci = orc_cidx(v, x >> 2, y >> 2, z >> 2);
    uint64_t m = 0;
    for (int bz = 4 * (z & 3); bz < 4 * (z & 3) + 4; bz++) m |= v->occ[ci * ORC_WORDS + bz * 4 + (y & 3)] & (0x000F000F000F000Full << (4 * (x & 3)));
    ne[x + (int64_t)dd[0] * (y + (int64_t)dd[1] * z)] = m != 0;
  }
  /* cube_cell_init_kernel + cube_cell_pass_kernel: relaxation rounds t = 2 .. 32, in place */
  for (int o = 0; o < 8; o++) for (int64_t i = 0; i < n; i++) cell[(size_t)o * n + i] = ne[i] ? 0 : 1;
  for (int t = 2; t <= cap; t++)
    for (int o = 0; o < 8; o++) {
      uint8_t* f = cell + (size_t)o * n;
      const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
      for (int z = 0; z < dd[2]; z++) for (int y = 0; y < dd[1]; y++) for (int x = 0; x < dd[0]; x++) {
        const int64_t i = x + (int64_t)dd[0] * (y + (int64_t)dd[1] * z);
        if (f[i] != t - 1) continue;
        int ok = 1;
        for (int q = 1; q < 8 && ok; q++) {
          const int nx = x + ((q & 1) ? sx : 0), ny = y + ((q & 2) ? sy : 0), nz = z + ((q & 4) ? sz : 0);
          if (nx < 0 || ny < 0 || nz < 0 || nx >= dd[0] || ny >= dd[1] || nz >= dd[2]) continue;
          if (f[nx + (int64_t)dd[0] * (ny + (int64_t)dd[1] * nz)] < t - 1) ok = 0;
        }
        if (ok) f[i] = (uint8_t)t;
      }
    }
  /* cube_brick_kernel: bricks of non-empty cells, shell tests up to 4 */
  Scene sc; memset(&sc, 0, sizeof(sc)); sc.v = v;
  memset(brick, 0, (size_t)v->nchunks * ORC_BLOCKS * sizeof(uint16_t));
  for (int ez = 0; ez < dd[2]; ez++) for (int ey = 0; ey < dd[1]; ey++) for (int ex = 0; ex < dd[0]; ex++) {
    if (!ne[ex + (int64_t)dd[0] * (ey + (int64_t)dd[1] * ez)]) continue;
    for (int l = 0; l < 64; l++) {
      const int bx = ex * 4 + (l & 3), by = ey * 4 + ((l >> 2) & 3), bz = ez * 4 + (l >> 4);
      unsigned r = 0;
      if (!brick_occupied(&sc, bx, by, bz)) {
        for (int o = 0; o < 8; o++) {
          const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
          int k = 1;
          for (int t = 2; t <= 4; t++) {
            int empty = 1;
            for (int z = 0; z < t && empty; z++) for (int y = 0; y < t && empty; y++) for (int x = 0; x < t; x++) {
              if (x < t - 1 && y < t - 1 && z < t - 1) continue;
              if (brick_occupied(&sc, bx + sx * x, by + sy * y, bz + sz * z)) { empty = 0; break; }
            }
            if (!empty) break;
            k = t;
          }
          r |= (unsigned)(k - 1) << (2 * o);
        }
      }
      const int64_t ci = orc_cidx(v, bx >> 4, by >> 4, bz >> 4);
      brick[(size_t)ci * ORC_BLOCKS + orc_bidx(bx & 15, by & 15, bz & 15)] = (uint16_t)r;
    }
  }
  /* cube_cell2_kernel: per payload slot, from the 2^3-cell mask */
  for (int64_t slot = 0; slot < v->pool_n; slot++) {
    const uint64_t* p = v->pool + (size_t)slot * 8;
    uint64_t cm = 0;
    for (int c = 0; c < 64; c++) {
      const int x0 = 2 * (c & 3), y0 = 2 * ((c >> 2) & 3), z0 = 2 * (c >> 4);
      const uint64_t mk = (3ull << (x0 + 8 * y0)) | (3ull << (x0 + 8 * (y0 + 1)));
      if ((p[z0] | p[z0 + 1]) & mk) cm |= 1ull << c;
    }
    for (int c = 0; c < 64; c++) {
      const int cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
      unsigned r = 0;
      if (!((cm >> c) & 1ull)) {
        for (int o = 0; o < 8; o++) {
          const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
          int k = 1;
          for (int t = 2; t <= 4; t++) {
            int empty = 1;
            for (int z = 0; z < t && empty; z++) for (int y = 0; y < t && empty; y++) for (int x = 0; x < t; x++) {
              if (x < t - 1 && y < t - 1 && z < t - 1) continue;
              const int qx = cx + sx * x, qy = cy + sy * y, qz = cz + sz * z;
              if (qx < 0 || qy < 0 || qz < 0 || qx > 3 || qy > 3 || qz > 3 || ((cm >> (qx + 4 * qy + 16 * qz)) & 1ull)) { empty = 0; break; }
            }
            if (!empty) break;
            k = t;
          }
          r |= (unsigned)(k - 1) << (2 * o);
        }
      }
      cell2[(size_t)slot * 64 + c] = (uint16_t)r;
    }
  }
  free(ne);
  return 0;
}

/* ORC_DDA_MODEL reads these tables instead of computing cubes on the fly (NULLs switch back). */
void orc_cube_tables_use(const OrcVolume* v, const uint8_t* cell, const uint16_t* brick, const uint16_t* cell2) {
  g_tab_cell = cell; g_tab_brick = brick; g_tab_cell2 = cell2;
  if (cell) {
    g_model.df_shift = 5;
    for (int i = 0; i < 3; i++) g_model_dd[i] = v->dims[i] * 4;
    g_tab_ncells = (int64_t)g_model_dd[0] * g_model_dd[1] * g_model_dd[2];
  }
}
