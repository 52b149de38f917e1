/* orc_mesh.c -- CPU ORACLE (test infrastructure): face culling + greedy merge into quads, sphere carve.
 * Neither exists in the reference ("No quads, no face-neighbour test, no greedy merge");
 * both are DEFINED here -- "parity unpinned by reference; bit-exact vs repo oracle".
 *
 * Meshing unit = one occupied 8^3 brick.  For each of the 6 face directions (ids as in OrcHitRecord) and each
 * of the 8 layers along that axis, the exposed-face mask is an 8x8 bit image (row v, bit u; (u,v) = (y,z) for X,
 * (x,z) for Y, (x,y) for Z).  A face is exposed iff the voxel is solid and its neighbour across the face is empty
 * (neighbour bricks/chunks consulted; outside the grid = empty).  Greedy merge per image, fixed order:
 *   for v = 0..7: while row[v] != 0: u0 = lowest set bit; w = length of the run of ones starting at u0;
 *                 h = 1 + number of following rows that contain the whole run (each cleared as it is taken).
 * These VOXEL-LEVEL quads never cross brick boundaries (w,h <= 8): the brick is the unit of the dirty re-mesh.
 *
 * Second level, over whole bricks (the block-granular scenes of the reference are made of nothing else): the face of a
 * FULL brick towards an ABSENT brick (unoccupied, or outside the grid) is left out of the voxel level and merged per chunk
 * instead.  For each chunk, direction and brick layer along the axis (16 of them) the image is 16 rows x 16 bits of such
 * faces, (u,v) as above in brick units, merged greedily in the same fixed order; a quad covers wb x hb bricks and is
 * recorded in voxel units (corner = the voxel layer carrying the face, w = 8 wb <= 128, h = 8 hb) with bit 19 of w1 set.
 * Brick-level quads never cross chunk boundaries: the chunk is their unit of re-meshing (a re-mesh of listed bricks
 * returns the voxel-level quads of those bricks and the brick-level quads of every chunk that holds one of them).
 * A flat 16 x 16-brick chunk face is therefore 1 quad, not 256.
 *
 * Emission order is unspecified; comparisons use the canonical sort of orc_sort_quads (lexicographic on (w3,w2,w1,w0)
 * as unsigned -- cf. Comparator.h:15-23).
 */
#include "orc_internal.h"

static inline uint8_t gather_col(uint64_t e, int x) { /* bits (x + 8*y), y=0..7 -> byte with bit y */
  return (uint8_t)((((e >> x) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}

/* fills rows[dir][layer][v] for one brick */
static void face_rows(const OrcVolume* v, int64_t bx, int64_t by, int64_t bz, const uint64_t s[8], uint8_t rows[6][8][8]) {
  uint64_t nxm[8], nxp[8], nym[8], nyp[8], nzm[8], nzp[8];
  orc_brick_slices(v, bx - 1, by, bz, nxm); orc_brick_slices(v, bx + 1, by, bz, nxp);
  orc_brick_slices(v, bx, by - 1, bz, nym); orc_brick_slices(v, bx, by + 1, bz, nyp);
  orc_brick_slices(v, bx, by, bz - 1, nzm); orc_brick_slices(v, bx, by, bz + 1, nzp);
  const uint64_t C0 = 0x0101010101010101ull; /* x = 0 column */
  for (int z = 0; z < 8; z++) {
    uint64_t m = s[z];
    uint64_t n_xm = ((m << 1) & ~C0) | ((nxm[z] >> 7) & C0);          /* neighbour at x-1 */
    uint64_t n_xp = ((m >> 1) & ~(C0 << 7)) | ((nxp[z] & C0) << 7);   /* neighbour at x+1 */
    uint64_t n_ym = (m << 8) | (nym[z] >> 56);
    uint64_t n_yp = (m >> 8) | (nyp[z] << 56);
    uint64_t n_zm = z > 0 ? s[z - 1] : nzm[7];
    uint64_t n_zp = z < 7 ? s[z + 1] : nzp[0];
    uint64_t e0 = m & ~n_xm, e1 = m & ~n_xp, e2 = m & ~n_ym, e3 = m & ~n_yp, e4 = m & ~n_zm, e5 = m & ~n_zp;
    for (int l = 0; l < 8; l++) {
      rows[0][l][z] = gather_col(e0, l);            /* X: layer x=l, row v=z, bit u=y */
      rows[1][l][z] = gather_col(e1, l);
      rows[2][l][z] = (uint8_t)(e2 >> (8 * l));     /* Y: layer y=l, row v=z, bit u=x */
      rows[3][l][z] = (uint8_t)(e3 >> (8 * l));
      rows[4][z][l] = (uint8_t)(e4 >> (8 * l));     /* Z: layer z, row v=y=l, bit u=x */
      rows[5][z][l] = (uint8_t)(e5 >> (8 * l));
    }
  }
}

/* 0 absent (or outside the grid), 1 full, 2 partial -- from the occupancy / full bits, not from the payload */
static int brick_state(const OrcVolume* v, int64_t bx, int64_t by, int64_t bz) {
  if (bx < 0 || by < 0 || bz < 0 || bx >= (int64_t)v->dims[0] * ORC_CR || by >= (int64_t)v->dims[1] * ORC_CR || bz >= (int64_t)v->dims[2] * ORC_CR) return 0;
  int64_t c = orc_cidx(v, (int)(bx >> 4), (int)(by >> 4), (int)(bz >> 4));
  int b = orc_bidx((int)(bx & 15), (int)(by & 15), (int)(bz & 15));
  if (!orc_getbit(v->occ + c * ORC_WORDS, b)) return 0;
  return orc_getbit(v->full + c * ORC_WORDS, b) ? 1 : 2;
}
static const int DIR_STEP[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
/* the face of brick (bx,by,bz) towards `dir` belongs to the brick level */
static int brick_level_face(const OrcVolume* v, int64_t bx, int64_t by, int64_t bz, int dir) {
  return brick_state(v, bx, by, bz) == 1 && brick_state(v, bx + DIR_STEP[dir][0], by + DIR_STEP[dir][1], bz + DIR_STEP[dir][2]) == 0;
}

static int64_t mesh_brick(const OrcVolume* v, int64_t bx, int64_t by, int64_t bz, OrcQuad* out, int64_t cap, int64_t n, int64_t* faces) {
  uint64_t s[8];
  orc_brick_slices(v, bx, by, bz, s);
  uint64_t any = 0; for (int z = 0; z < 8; z++) any |= s[z];
  if (!any) return n;
  uint8_t rows[6][8][8];
  face_rows(v, bx, by, bz, s, rows);
  for (int dir = 0; dir < 6; dir++) {
    int ax = dir >> 1;
    const int lifted = brick_level_face(v, bx, by, bz, dir);   /* layer 0 (minus side) / 7 (plus side) goes to the brick level */
    for (int l = 0; l < 8; l++) {
      uint8_t* r = rows[dir][l];
      if (faces) for (int q = 0; q < 8; q++) *faces += __builtin_popcount(r[q]);
      if (lifted && l == ((dir & 1) ? 7 : 0)) continue;
      for (int vv = 0; vv < 8; vv++) {
        while (r[vv]) {
          int u0 = __builtin_ctz(r[vv]);
          int w = __builtin_ctz(~((unsigned)r[vv] >> u0));
          uint8_t m = (uint8_t)(((1u << w) - 1u) << u0);
          int h = 1;
          while (vv + h < 8 && (r[vv + h] & m) == m) { r[vv + h] &= (uint8_t)~m; h++; }
          r[vv] &= (uint8_t)~m;
          int x, y, z;
          if (ax == 0) { x = l; y = u0; z = vv; } else if (ax == 1) { x = u0; y = l; z = vv; } else { x = u0; y = vv; z = l; }
          if (out && n < cap) {
            out[n].w0 = (uint32_t)(bx * 8 + x) | ((uint32_t)(by * 8 + y) << 16);
            out[n].w1 = (uint32_t)(bz * 8 + z) | ((uint32_t)dir << 16) | ((uint32_t)w << 24);
            out[n].w2 = (uint32_t)h; out[n].w3 = 0;
          }
          n++;
        }
      }
    }
  }
  return n;
}

/* brick-level quads of one chunk */
static int64_t mesh_chunk_faces(const OrcVolume* v, int64_t c, OrcQuad* out, int64_t cap, int64_t n) {
  int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
  uint64_t any_full = 0;
  for (int w = 0; w < ORC_WORDS; w++) any_full |= v->full[c * ORC_WORDS + w];
  if (!any_full) return n;
  for (int dir = 0; dir < 6; dir++) {
    int ax = dir >> 1;
    for (int l = 0; l < 16; l++) {
      uint16_t r[16];
      for (int vv = 0; vv < 16; vv++) {
        r[vv] = 0;
        for (int u = 0; u < 16; u++) {
          int x, y, z;
          if (ax == 0) { x = l; y = u; z = vv; } else if (ax == 1) { x = u; y = l; z = vv; } else { x = u; y = vv; z = l; }
          if (brick_level_face(v, cx * 16 + x, cy * 16 + y, cz * 16 + z, dir)) r[vv] |= (uint16_t)(1u << u);
        }
      }
      for (int vv = 0; vv < 16; vv++) {
        while (r[vv]) {
          int u0 = __builtin_ctz(r[vv]);
          int w = __builtin_ctz(~((unsigned)r[vv] >> u0));
          uint16_t m = (uint16_t)(((1u << w) - 1u) << u0);
          int h = 1;
          while (vv + h < 16 && (r[vv + h] & m) == m) { r[vv + h] &= (uint16_t)~m; h++; }
          r[vv] &= (uint16_t)~m;
          int lv = l * 8 + ((dir & 1) ? 7 : 0);   /* the voxel layer that carries the face */
          int x, y, z;
          if (ax == 0) { x = lv; y = u0 * 8; z = vv * 8; } else if (ax == 1) { x = u0 * 8; y = lv; z = vv * 8; } else { x = u0 * 8; y = vv * 8; z = lv; }
          if (out && n < cap) {
            out[n].w0 = (uint32_t)(cx * 128 + x) | ((uint32_t)(cy * 128 + y) << 16);
            out[n].w1 = (uint32_t)(cz * 128 + z) | ((uint32_t)dir << 16) | (1u << 19) | ((uint32_t)(w * 8) << 24);
            out[n].w2 = (uint32_t)(h * 8); out[n].w3 = 0;
          }
          n++;
        }
      }
    }
  }
  return n;
}

int64_t orc_mesh_chunk_faces(const OrcVolume* v, const int64_t* chunks, int64_t nc, OrcQuad* quads, int64_t cap) {
  int64_t n = 0;
  for (int64_t i = 0; i < nc; i++) n = mesh_chunk_faces(v, chunks[i], quads, cap, n);
  return n;
}

int64_t orc_mesh_bricks(const OrcVolume* v, const uint64_t* keys, int64_t nk, OrcQuad* quads, int64_t cap) {
  int64_t n = 0;
  for (int64_t i = 0; i < nk; i++) {
    int64_t c = (int64_t)(keys[i] / ORC_BLOCKS); int b = (int)(keys[i] % ORC_BLOCKS);
    int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
    n = mesh_brick(v, cx * 16 + (b & 15), cy * 16 + ((b >> 4) & 15), cz * 16 + (b >> 8), quads, cap, n, NULL);
  }
  return n;
}

typedef struct { const OrcVolume* v; OrcQuad** bufs; int64_t* counts; int64_t* caps; int64_t faces; pthread_mutex_t mu; int count_only; } MeshArg;
static void mesh_range(void* ctx, int64_t b, int64_t e, int tid) {
  (void)tid;
  MeshArg* a = (MeshArg*)ctx; const OrcVolume* v = a->v;
  int64_t faces = 0;
  for (int64_t c = b; c < e; c++) {
    int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
    /* first pass counts, second pass emits into an exactly-sized per-chunk buffer */
    int64_t n = 0;
    for (int bb = 0; bb < ORC_BLOCKS; bb++)
      if (orc_getbit(v->occ + c * ORC_WORDS, bb))
        n = mesh_brick(v, cx * 16 + (bb & 15), cy * 16 + ((bb >> 4) & 15), cz * 16 + (bb >> 8), NULL, 0, n, &faces);
    n = mesh_chunk_faces(v, c, NULL, 0, n);
    a->counts[c] = n;
    if (!a->count_only && n > 0) {
      a->bufs[c] = (OrcQuad*)malloc(sizeof(OrcQuad) * (size_t)n);
      int64_t m = 0;
      for (int bb = 0; bb < ORC_BLOCKS; bb++)
        if (orc_getbit(v->occ + c * ORC_WORDS, bb))
          m = mesh_brick(v, cx * 16 + (bb & 15), cy * 16 + ((bb >> 4) & 15), cz * 16 + (bb >> 8), a->bufs[c], n, m, NULL);
      mesh_chunk_faces(v, c, a->bufs[c], n, m);
    }
  }
  pthread_mutex_lock(&a->mu); a->faces += faces; pthread_mutex_unlock(&a->mu);
}

static int64_t mesh_all(const OrcVolume* v, int nthreads, OrcQuad* quads, int64_t cap, int count_only, int64_t* faces) {
  MeshArg a; memset(&a, 0, sizeof(a));
  a.v = v; a.count_only = count_only;
  a.bufs = (OrcQuad**)calloc((size_t)v->nchunks, sizeof(OrcQuad*));
  a.counts = (int64_t*)calloc((size_t)v->nchunks, sizeof(int64_t));
  pthread_mutex_init(&a.mu, NULL);
  orc_parallel_for(v->nchunks, nthreads, 1, mesh_range, &a);
  pthread_mutex_destroy(&a.mu);
  int64_t n = 0;
  for (int64_t c = 0; c < v->nchunks; c++) {
    if (a.bufs[c]) {
      for (int64_t i = 0; i < a.counts[c]; i++) if (quads && n + i < cap) quads[n + i] = a.bufs[c][i];
      free(a.bufs[c]);
    }
    n += a.counts[c];
  }
  if (faces) *faces = a.faces;
  free(a.bufs); free(a.counts);
  return n;
}
int64_t orc_mesh(const OrcVolume* v, int nthreads, OrcQuad* quads, int64_t cap) { return mesh_all(v, nthreads, quads, cap, quads == NULL, NULL); }
int64_t orc_count_exposed_faces(const OrcVolume* v) { int64_t f = 0; mesh_all(v, orc_hardware_threads(), NULL, 0, 1, &f); return f; }

static int quad_cmp(const void* pa, const void* pb) {
  const OrcQuad* a = (const OrcQuad*)pa; const OrcQuad* b = (const OrcQuad*)pb;
  if (a->w3 != b->w3) return a->w3 < b->w3 ? -1 : 1;
  if (a->w2 != b->w2) return a->w2 < b->w2 ? -1 : 1;
  if (a->w1 != b->w1) return a->w1 < b->w1 ? -1 : 1;
  if (a->w0 != b->w0) return a->w0 < b->w0 ? -1 : 1;
  return 0;
}
void orc_sort_quads(OrcQuad* q, int64_t n) { qsort(q, (size_t)n, sizeof(OrcQuad), quad_cmp); }

/* ---- K5: sphere carve -------------------------------------------------------------------------
 * A voxel (x,y,z) is removed iff its centre (x+.5,y+.5,z+.5) lies strictly inside the sphere of integer centre
 * and radius in grid voxel units: (2x+1-2cx)^2 + (2y+1-2cy)^2 + (2z+1-2cz)^2 < (2r)^2, all in int64. */
int64_t orc_carve_sphere(OrcVolume* v, const int32_t ctr[3], int32_t radius, uint64_t* dirty, int64_t cap) {
  int64_t nd = 0;
  int64_t lo[3], hi[3];
  for (int i = 0; i < 3; i++) {
    int64_t a = ((int64_t)ctr[i] - radius) >> 3, b = ((int64_t)ctr[i] + radius) >> 3;
    int64_t nb = (int64_t)v->dims[i] * ORC_CR;
    if (a < 0) a = 0;
    if (b > nb - 1) b = nb - 1;
    lo[i] = a; hi[i] = b;
  }
  int64_t r2 = 4 * (int64_t)radius * radius;
  /* iterate in ascending key order: chunk index major (z,y,x), then block index (z,y,x) */
  for (int64_t c = 0; c < v->nchunks; c++) {
    int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
    if (cx * 16 + 15 < lo[0] || cx * 16 > hi[0] || cy * 16 + 15 < lo[1] || cy * 16 > hi[1] || cz * 16 + 15 < lo[2] || cz * 16 > hi[2]) continue;
    for (int b = 0; b < ORC_BLOCKS; b++) {
      int64_t bx = cx * 16 + (b & 15), by = cy * 16 + ((b >> 4) & 15), bz = cz * 16 + (b >> 8);
      if (bx < lo[0] || bx > hi[0] || by < lo[1] || by > hi[1] || bz < lo[2] || bz > hi[2]) continue;
      if (!orc_getbit(v->occ + c * ORC_WORDS, b)) continue;
      uint64_t s[8], t[8]; int changed = 0;
      orc_brick_slices(v, bx, by, bz, s);
      for (int z = 0; z < 8; z++) {
        uint64_t m = s[z];
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
          int64_t dx = 2 * (bx * 8 + x) + 1 - 2 * (int64_t)ctr[0], dy = 2 * (by * 8 + y) + 1 - 2 * (int64_t)ctr[1], dz = 2 * (bz * 8 + z) + 1 - 2 * (int64_t)ctr[2];
          if (dx * dx + dy * dy + dz * dz < r2) m &= ~(1ull << (x + 8 * y));
        }
        t[z] = m; if (m != s[z]) changed = 1;
      }
      if (!changed) continue;
      if (nd < cap && dirty) dirty[nd] = (uint64_t)c * ORC_BLOCKS + (uint64_t)b;
      nd++;
      uint64_t any = 0; for (int z = 0; z < 8; z++) any |= t[z];
      int was_full = orc_getbit(v->full + c * ORC_WORDS, b);
      orc_setbit(v->full + c * ORC_WORDS, b, 0);
      if (!any) { orc_setbit(v->occ + c * ORC_WORDS, b, 0); if (!was_full) v->bptr[c][b] = 0xFFFFFFFFu; continue; }
      if (!v->bptr[c]) { v->bptr[c] = (uint32_t*)malloc(sizeof(uint32_t) * ORC_BLOCKS); memset(v->bptr[c], 0xFF, sizeof(uint32_t) * ORC_BLOCKS); }
      if (was_full || v->bptr[c][b] == 0xFFFFFFFFu) v->bptr[c][b] = orc_alloc_payload(v);
      memcpy(v->pool + (size_t)v->bptr[c][b] * 8, t, sizeof(t));
    }
  }
  return nd;
}
