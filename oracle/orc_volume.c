/* orc_volume.c -- CPU ORACLE (test infrastructure): the chunk -> block -> voxel volume.
 * Block level restates FChunk / FBinaryOccupancyVolume (Runtimes/Voxel/Chunk/Chunk.h:57-101,
 * Runtimes/Voxel/Occupancy/BinaryOccupancyVolume.h:5-39); the 8^3 brick payload follows the layout
 * intent of FVolume (Runtimes/Voxel/VoxelStructure.h:30-39: BlockResolution^3/32 = 16 x u32 per block).
 * The voxel-in-brick level itself is a repo extension ("parity unpinned by reference"). */
#include "orc_internal.h"
#include <unistd.h>

int orc_hardware_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}

/* ---- tiny parallel-for: dynamic scheduling over [0, n) in grains, one atomic fetch-add per grain -------- */
typedef struct { orc_range_fn fn; void* ctx; int64_t n, grain; int64_t* next; int tid; } PfArg;
static void* pf_worker(void* p) {
  PfArg* a = (PfArg*)p;
  for (;;) {
    const int64_t b = __atomic_fetch_add(a->next, a->grain, __ATOMIC_RELAXED);
    if (b >= a->n) break;
    int64_t e = b + a->grain; if (e > a->n) e = a->n;
    a->fn(a->ctx, b, e, a->tid);
  }
  return NULL;
}
void orc_parallel_for(int64_t n, int nthreads, int64_t grain, orc_range_fn fn, void* ctx) {
  if (n <= 0) return;
  if (grain < 1) grain = 1;
  if (nthreads <= 1) { fn(ctx, 0, n, 0); return; }
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; PfArg args[256];
  int64_t next = 0;
  for (int i = 0; i < nthreads; i++) {
    args[i] = (PfArg){fn, ctx, n, grain, &next, i};
    if (i > 0) pthread_create(&th[i], NULL, pf_worker, &args[i]);
  }
  pf_worker(&args[0]);   /* the calling thread works too: nthreads threads busy, not nthreads + 1 */
  for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
}

/* ---- volume ---------------------------------------------------------------------------------- */
OrcVolume* orc_volume_create(const int32_t origin[3], const int32_t dims[3]) {
  OrcVolume* v = (OrcVolume*)calloc(1, sizeof(OrcVolume));
  for (int i = 0; i < 3; i++) { v->origin[i] = origin[i]; v->dims[i] = dims[i]; }
  v->nchunks = (int64_t)dims[0] * dims[1] * dims[2];
  v->occ = (uint64_t*)calloc((size_t)v->nchunks * ORC_WORDS, sizeof(uint64_t));
  v->full = (uint64_t*)calloc((size_t)v->nchunks * ORC_WORDS, sizeof(uint64_t));
  v->bptr = (uint32_t**)calloc((size_t)v->nchunks, sizeof(uint32_t*));
  v->pool_cap = 1024; v->pool_n = 0;
  v->pool = (uint64_t*)malloc(sizeof(uint64_t) * 8 * (size_t)v->pool_cap);
  pthread_mutex_init(&v->lock, NULL);
  return v;
}
void orc_volume_destroy(OrcVolume* v) {
  if (!v) return;
  for (int64_t c = 0; c < v->nchunks; c++) free(v->bptr[c]);
  free(v->bptr); free(v->occ); free(v->full); free(v->pool); free(v->any);
  pthread_mutex_destroy(&v->lock);
  free(v);
}
static void clear_volume(OrcVolume* v) {
  memset(v->occ, 0, sizeof(uint64_t) * ORC_WORDS * (size_t)v->nchunks);
  memset(v->full, 0, sizeof(uint64_t) * ORC_WORDS * (size_t)v->nchunks);
  for (int64_t c = 0; c < v->nchunks; c++) { free(v->bptr[c]); v->bptr[c] = NULL; }
  v->pool_n = 0;
  v->any_valid = 0;
}
uint32_t orc_alloc_payload(OrcVolume* v) { /* caller holds v->lock */
  if (v->pool_n == v->pool_cap) {
    v->pool_cap *= 2;
    v->pool = (uint64_t*)realloc(v->pool, sizeof(uint64_t) * 8 * (size_t)v->pool_cap);
  }
  return (uint32_t)v->pool_n++;
}
static void set_brick(OrcVolume* v, int64_t c, int b, const uint64_t s[8]) {
  uint64_t any = 0, all = ~0ull;
  for (int z = 0; z < 8; z++) { any |= s[z]; all &= s[z]; }
  uint64_t* occ = v->occ + c * ORC_WORDS; uint64_t* full = v->full + c * ORC_WORDS;
  pthread_mutex_lock(&v->lock); /* X slabs of one chunk share occupancy words */
  orc_setbit(occ, b, any != 0);
  orc_setbit(full, b, all == ~0ull);
  if (any != 0 && all != ~0ull) {
    if (!v->bptr[c]) { v->bptr[c] = (uint32_t*)malloc(sizeof(uint32_t) * ORC_BLOCKS); memset(v->bptr[c], 0xFF, sizeof(uint32_t) * ORC_BLOCKS); }
    uint32_t idx = v->bptr[c][b];
    if (idx == 0xFFFFFFFFu) { idx = orc_alloc_payload(v); v->bptr[c][b] = idx; }
    memcpy(v->pool + (size_t)idx * 8, s, sizeof(uint64_t) * 8);
  }
  pthread_mutex_unlock(&v->lock);
}

typedef struct { OrcVolume* v; int kind; const double* params; int gran; int sin_mode; int fast; int mip; } VoxArg;

/* Sphere fast path (ORC_FAST=1 in orc_volume_voxelize_ex).  The fp64 sphere SDF sqrt((dx*dx + dy*dy) + dz*dz) - r is
 * monotone non-decreasing in each of |dx|, |dy|, |dz| separately (every rounding step is monotone), so over an
 * axis-aligned lattice of sample points its maximum is taken at the per-axis farthest sample and its minimum at the
 * per-axis nearest one.  A box of samples is therefore all-solid iff the farthest sample is solid and all-empty iff
 * the nearest sample is not: exact, no tolerance.  tests/test_oracle_kat.py checks it against the brute-force walk. */
static void axis_extremes(double lo, double step, int n, double c, double* nearest, double* farthest) {
  double best_n = 0, best_f = 0, dn = INFINITY, df = -1.0;
  for (int i = 0; i < n; i++) {
    double p = lo + (double)i * step;
    double d = fabs(p - c);
    if (d < dn) { dn = d; best_n = p; }
    if (d > df) { df = d; best_f = p; }
  }
  *nearest = best_n; *farthest = best_f;
}
/* returns 1 = all solid, 0 = all empty, -1 = mixed, for the n^3 samples lo + i*step */
static int sphere_box_class(const double* params, const double lo[3], double step, int n) {
  double nr[3], fr[3];
  for (int k = 0; k < 3; k++) axis_extremes(lo[k], step, n, params[k], &nr[k], &fr[k]);
  if (orc_sdf(ORC_SDF_SPHERE, params, 0, fr[0], fr[1], fr[2]) < 0.0) return 1;
  if (!(orc_sdf(ORC_SDF_SPHERE, params, 0, nr[0], nr[1], nr[2]) < 0.0)) return 0;
  return -1;
}
/* One chunk: block-granular = GeneratorHelper.h:120-150 verbatim (block solid <=> sdf(min corner) < 0, brick
 * all-ones); voxel-granular (extension) samples every voxel min corner p = blockCorner + (vx,vy,vz)*BlockSize/8. */
static void vox_range(void* ctx, int64_t b, int64_t e, int tid) {
  (void)tid;
  VoxArg* a = (VoxArg*)ctx; OrcVolume* v = a->v;
  const double BlockSize = 1.0;
  for (int64_t item = b; item < e; item++) {
    /* work item = (chunk, X slab) for voxel granularity, whole chunk for block granularity */
    int64_t c = a->gran == ORC_GRAN_BLOCK ? item : item / ORC_CR;
    int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
    int32_t loc[3] = {v->origin[0] + cx, v->origin[1] + cy, v->origin[2] + cz};
    if (a->gran == ORC_GRAN_BLOCK) {
      uint8_t xyz[3 * ORC_BLOCKS];
      int n = orc_generate_chunk(a->kind, a->params, a->sin_mode, loc, 1.0f, ORC_CR, xyz);
      uint64_t ones[8]; for (int z = 0; z < 8; z++) ones[z] = ~0ull;
      if (a->mip <= 0) {
        for (int i = 0; i < n; i++) set_brick(v, c, orc_bidx(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), ones);
      } else {
        /* MipmapLevel m (the generator plug-in's fourth argument, ChunkManager.h:61; the reference's generators ignore it --
         * this definition is ours): one sample per (2^m)^3 blocks, taken at the group's minimum-corner block, i.e. block
         * (X, Y, Z) is solid iff the reference generator makes block (X & ~(2^m - 1), ...) solid. */
        uint8_t solid[ORC_BLOCKS]; memset(solid, 0, sizeof solid);
        for (int i = 0; i < n; i++) solid[orc_bidx(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])] = 1;
        const int mask = ~((1 << a->mip) - 1);
        for (int X = 0; X < ORC_CR; X++) for (int Y = 0; Y < ORC_CR; Y++) for (int Z = 0; Z < ORC_CR; Z++)
          if (solid[orc_bidx(X & mask, Y & mask, Z & mask)]) set_brick(v, c, orc_bidx(X, Y, Z), ones);
      }
    } else {
      double cs[3]; for (int k = 0; k < 3; k++) cs[k] = (double)loc[k] * BlockSize * (double)ORC_CR;
      int X = (int)(item % ORC_CR);
      int chunk_class = -1;
      if (a->fast && a->kind == ORC_SDF_SPHERE) chunk_class = sphere_box_class(a->params, cs, BlockSize / 8.0, ORC_CV);
      if (chunk_class == 0) continue;
      for (int Y = 0; Y < ORC_CR; Y++) for (int Z = 0; Z < ORC_CR; Z++) {
        double bc[3] = {cs[0] + (double)X * BlockSize, cs[1] + (double)Y * BlockSize, cs[2] + (double)Z * BlockSize};
        uint64_t s[8];
        if (a->fast && a->kind == ORC_SDF_SPHERE) {
          int cls = chunk_class == 1 ? 1 : sphere_box_class(a->params, bc, BlockSize / 8.0, ORC_BR);
          if (cls == 0) continue;
          if (cls == 1) { for (int z = 0; z < 8; z++) s[z] = ~0ull; set_brick(v, c, orc_bidx(X, Y, Z), s); continue; }
        }
        for (int vz = 0; vz < 8; vz++) {
          uint64_t m = 0;
          for (int vy = 0; vy < 8; vy++) for (int vx = 0; vx < 8; vx++) {
            double px = bc[0] + (double)vx * (BlockSize / 8.0), py = bc[1] + (double)vy * (BlockSize / 8.0), pz = bc[2] + (double)vz * (BlockSize / 8.0);
            if (orc_sdf(a->kind, a->params, a->sin_mode, px, py, pz) < 0.0) m |= 1ull << (vx + 8 * vy);
          }
          s[vz] = m;
        }
        set_brick(v, c, orc_bidx(X, Y, Z), s);
      }
    }
  }
}
void orc_volume_voxelize_lod(OrcVolume* v, int kind, const double params[4], int sin_mode, int nthreads, int mip) {
  clear_volume(v);
  VoxArg a = {v, kind, params, ORC_GRAN_BLOCK, sin_mode, 0, mip < 0 ? 0 : (mip > 4 ? 4 : mip)};
  orc_parallel_for(v->nchunks, nthreads, 1, vox_range, &a);
}
void orc_volume_voxelize_ex(OrcVolume* v, int kind, const double params[4], int gran, int sin_mode, int nthreads, int fast) {
  clear_volume(v);
  VoxArg a = {v, kind, params, gran, sin_mode, fast, 0};
  orc_parallel_for(gran == ORC_GRAN_BLOCK ? v->nchunks : v->nchunks * ORC_CR, nthreads, 1, vox_range, &a);
}
void orc_volume_voxelize(OrcVolume* v, int kind, const double params[4], int gran, int sin_mode, int nthreads) {
  orc_volume_voxelize_ex(v, kind, params, gran, sin_mode, nthreads, 0);
}

int64_t orc_volume_num_chunks(const OrcVolume* v) { return v->nchunks; }
const uint64_t* orc_volume_occ(const OrcVolume* v) { return v->occ; }
const uint64_t* orc_volume_full(const OrcVolume* v) { return v->full; }

int64_t orc_volume_num_partial(const OrcVolume* v) {
  int64_t n = 0;
  for (int64_t i = 0; i < v->nchunks * ORC_WORDS; i++) n += __builtin_popcountll(v->occ[i] & ~v->full[i]);
  return n;
}
int64_t orc_volume_export_partial(const OrcVolume* v, uint64_t* keys, uint64_t* payload, int64_t cap) {
  int64_t n = 0;
  for (int64_t c = 0; c < v->nchunks; c++) {
    const uint64_t* occ = v->occ + c * ORC_WORDS; const uint64_t* full = v->full + c * ORC_WORDS;
    for (int b = 0; b < ORC_BLOCKS; b++) {
      if (orc_getbit(occ, b) && !orc_getbit(full, b)) {
        if (n < cap) {
          keys[n] = (uint64_t)c * ORC_BLOCKS + (uint64_t)b;
          memcpy(payload + n * 8, v->pool + (size_t)v->bptr[c][b] * 8, sizeof(uint64_t) * 8);
        }
        n++;
      }
    }
  }
  return n;
}
void orc_volume_import(OrcVolume* v, const uint64_t* occ, const uint64_t* full, const uint64_t* keys,
                       const uint64_t* payload, int64_t n_partial) {
  clear_volume(v);
  memcpy(v->occ, occ, sizeof(uint64_t) * ORC_WORDS * (size_t)v->nchunks);
  memcpy(v->full, full, sizeof(uint64_t) * ORC_WORDS * (size_t)v->nchunks);
  for (int64_t i = 0; i < n_partial; i++) {
    int64_t c = (int64_t)(keys[i] / ORC_BLOCKS); int b = (int)(keys[i] % ORC_BLOCKS);
    if (!v->bptr[c]) { v->bptr[c] = (uint32_t*)malloc(sizeof(uint32_t) * ORC_BLOCKS); memset(v->bptr[c], 0xFF, sizeof(uint32_t) * ORC_BLOCKS); }
    uint32_t idx = orc_alloc_payload(v);
    v->bptr[c][b] = idx;
    memcpy(v->pool + (size_t)idx * 8, payload + i * 8, sizeof(uint64_t) * 8);
  }
}

void orc_brick_slices(const OrcVolume* v, int64_t bx, int64_t by, int64_t bz, uint64_t out[8]) {
  memset(out, 0, sizeof(uint64_t) * 8);
  if (bx < 0 || by < 0 || bz < 0 || bx >= (int64_t)v->dims[0] * ORC_CR || by >= (int64_t)v->dims[1] * ORC_CR || bz >= (int64_t)v->dims[2] * ORC_CR) return;
  int64_t c = orc_cidx(v, (int)(bx >> 4), (int)(by >> 4), (int)(bz >> 4));
  int b = orc_bidx((int)(bx & 15), (int)(by & 15), (int)(bz & 15));
  if (!orc_getbit(v->occ + c * ORC_WORDS, b)) return;
  if (orc_getbit(v->full + c * ORC_WORDS, b)) { for (int z = 0; z < 8; z++) out[z] = ~0ull; return; }
  memcpy(out, v->pool + (size_t)v->bptr[c][b] * 8, sizeof(uint64_t) * 8);
}
int orc_volume_get_voxel(const OrcVolume* v, int x, int y, int z) {
  if (x < 0 || y < 0 || z < 0) return 0;
  uint64_t s[8];
  orc_brick_slices(v, x >> 3, y >> 3, z >> 3, s);
  return (int)((s[z & 7] >> ((x & 7) + 8 * (y & 7))) & 1u);
}
int64_t orc_volume_count_voxels(const OrcVolume* v) {
  int64_t n = 0;
  for (int64_t c = 0; c < v->nchunks; c++)
    for (int b = 0; b < ORC_BLOCKS; b++) {
      if (!orc_getbit(v->occ + c * ORC_WORDS, b)) continue;
      if (orc_getbit(v->full + c * ORC_WORDS, b)) { n += 512; continue; }
      const uint64_t* s = v->pool + (size_t)v->bptr[c][b] * 8;
      for (int z = 0; z < 8; z++) n += __builtin_popcountll(s[z]);
    }
  return n;
}

/* K2 over the grid: for every chunk rebuild the reference's block list in generator order from the occupancy
 * mask, then run the restated Chunk.h:73-94 mips and ChunkPool.h:385-390,438 emission.  chunk_table follows
 * ChunkPool.h:567 ({location, stamp}); chunks without blocks keep the INT_MAX "invalid" location (Chunk.h:29),
 * i.e. they are FEmptyChunk and own no GPU chunk record. */
int64_t orc_volume_build_occupancy(const OrcVolume* v, uint32_t stamp, OrcGPUChunk* table, uint64_t* mips123,
                                   OrcGPUBlock* inst, int64_t cap) {
  int64_t total = 0;
  uint8_t xyz[3 * ORC_BLOCKS];
  uint64_t mips[4 * ORC_WORDS];
  OrcGPUBlock tmp[ORC_BLOCKS];
  for (int64_t c = 0; c < v->nchunks; c++) {
    const uint64_t* occ = v->occ + c * ORC_WORDS;
    int n = 0;
    for (int X = 0; X < ORC_CR; X++) for (int Y = 0; Y < ORC_CR; Y++) for (int Z = 0; Z < ORC_CR; Z++)
      if (orc_getbit(occ, orc_bidx(X, Y, Z))) { xyz[3 * n] = (uint8_t)X; xyz[3 * n + 1] = (uint8_t)Y; xyz[3 * n + 2] = (uint8_t)Z; n++; }
    int cx = (int)(c % v->dims[0]), cy = (int)((c / v->dims[0]) % v->dims[1]), cz = (int)(c / ((int64_t)v->dims[0] * v->dims[1]));
    if (table) {
      if (n > 0) { table[c].ChunkLocation[0] = v->origin[0] + cx; table[c].ChunkLocation[1] = v->origin[1] + cy; table[c].ChunkLocation[2] = v->origin[2] + cz; table[c].ChunkFrameStamp = stamp; }
      else { table[c].ChunkLocation[0] = table[c].ChunkLocation[1] = table[c].ChunkLocation[2] = INT_MAX; table[c].ChunkFrameStamp = 0; }
    }
    orc_erode_mips(xyz, n, 4, 1, mips);
    if (mips123) memcpy(mips123 + (size_t)c * 3 * ORC_WORDS, mips + ORC_WORDS, sizeof(uint64_t) * 3 * ORC_WORDS);
    int k = orc_emit_instances(xyz, n, mips, 1, (uint32_t)c, stamp, tmp);
    for (int i = 0; i < k; i++) { if (inst && total + i < cap) inst[total + i] = tmp[i]; }
    total += k;
  }
  return total;
}

/* payload slots ever allocated (>= live partial bricks: a carve retires slots without reusing them) */
int64_t orc_volume_pool_slots(const OrcVolume* v) { return v->pool_n; }
