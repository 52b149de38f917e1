/* orc_internal.h -- private to the CPU oracle (test infrastructure; see meso_oracle.h). */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include "meso_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>

#define ORC_CR 16          /* blocks per chunk axis  (VoxelSceneConfig.h:24) */
#define ORC_BR 8           /* voxels per block axis  (VoxelSceneConfig.h:22) */
#define ORC_CV 128         /* voxels per chunk axis */
#define ORC_WORDS 64       /* u64 words per 16^3 bit volume */
#define ORC_BLOCKS 4096

struct OrcVolume {
  int32_t origin[3];
  int32_t dims[3];
  int64_t nchunks;
  uint64_t* occ;    /* nchunks*64: block present (Mip0)                   */
  uint64_t* full;   /* nchunks*64: brick is all-solid (no payload stored)  */
  uint32_t** bptr;  /* per chunk: NULL or 4096 payload indices (0xFFFFFFFF = none) */
  uint64_t* pool;   /* 8 words per payload */
  int64_t pool_n, pool_cap;
  pthread_mutex_t lock;
  uint8_t* any;     /* per chunk: has >= 1 block.  Cache of the raymarch walk; cleared with the volume and rebuilt by the
                       first frame after it.  A carve only removes blocks, so a stale 1 is merely conservative.       */
  int any_valid;
};

/* bit index of block (x,y,z) in a chunk: x + 16*y + 256*z  (VoxelMathHelper.h:73-76) */
static inline int orc_bidx(int x, int y, int z) { return x + ORC_CR * y + ORC_CR * ORC_CR * z; }
static inline int orc_getbit(const uint64_t* m, int i) { return (int)((m[i >> 6] >> (i & 63)) & 1u); }
static inline void orc_setbit(uint64_t* m, int i, int v) {
  if (v) m[i >> 6] |= (1ull << (i & 63)); else m[i >> 6] &= ~(1ull << (i & 63));
}
static inline int64_t orc_cidx(const OrcVolume* v, int cx, int cy, int cz) {
  return (int64_t)cx + (int64_t)v->dims[0] * ((int64_t)cy + (int64_t)v->dims[1] * (int64_t)cz);
}

typedef void (*orc_range_fn)(void* ctx, int64_t begin, int64_t end, int tid);
void orc_parallel_for(int64_t n, int nthreads, int64_t grain, orc_range_fn fn, void* ctx);

/* brick payload (8 z-slices, bit x+8y) of block b in chunk c; all-ones for full, zeros for absent/outside. */
void orc_brick_slices(const OrcVolume* v, int64_t bx, int64_t by, int64_t bz, uint64_t out[8]);
uint32_t orc_alloc_payload(OrcVolume* v);

#endif
