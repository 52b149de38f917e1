/*
 * orc_resident.c -- CPU ORACLE (test infrastructure): which chunks should be resident for a view, and how
 * important each one is.  Restates, in fp32 with glm's operation order:
 *   FImportanceComputeInfo::CalculateChunkImportance   Runtimes/Voxel/Chunk/ChunkManagerHelper.h:26-44
 *   FImportanceComputeInfo::CalculateBlockImportance   ChunkManagerHelper.h:50-69
 *   FChunkManageHelper::GetDesiredShowChunkLocationByView   ChunkManagerHelper.h:89-150
 *   FChunkManageHelper::GetDesiredShowChunkLocationSimple   ChunkManagerHelper.h:151-198
 *   TNearestMap::Query (nearest baked direction)        Runtimes/Voxel/Spatial/NearestMap.h:32-47
 *
 * glm is un-vendored; its published definitions used here:
 *   dot(a,b)  = (a.x*b.x + a.y*b.y) + a.z*b.z        (detail/func_geometric.inl compute_dot<3>)
 *   length(v) = sqrt(dot(v,v))
 *   normalize(v) = v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1 / sqrt(x)
 *   radians(d) = d * 0.01745329251994329576923690768489f
 * normalize of the zero vector yields NaN components; the reference relies on max(0, NaN) == 0 for std::max's
 * (a < b) ? b : a definition.  The same comparisons are written out here.
 *
 * The reference returns a std::priority_queue; its pop order among equal importances is whatever the heap does.
 * The oracle's order is the canonical one: importance descending, ties in loop order (X outer, Y, Z inner).
 */
#include "meso_oracle.h"
#include <math.h>
#include <stdlib.h>

static float std_max(float a, float b) { return (a < b) ? b : a; }  /* std::max(a, b) */
static float std_min(float a, float b) { return (b < a) ? b : a; }  /* std::min(a, b) */

static float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void normalize3(const float v[3], float out[3]) {
  const float inv = 1.0f / sqrtf(dot3(v, v));
  out[0] = v[0] * inv; out[1] = v[1] * inv; out[2] = v[2] * inv;
}

/* shared tail of Calculate{Chunk,Block}Importance (ChunkManagerHelper.h:37-41, 63-67) */
static float importance_far(const int32_t off[3], const float fwd[3], float far) {
  const float o[3] = {(float)off[0], (float)off[1], (float)off[2]};
  float dir[3];
  normalize3(o, dir);
  const float dist = sqrtf(dot3(o, o));
  const float angle = std_max((std_max(0.0f, dot3(dir, fwd)) - 0.5f) * 2.0f, 0.75f);
  const float distance = std_max(0.25f, far - dist);
  return angle * distance;
}

float orc_chunk_importance(const int32_t cam_chunk[3], const float fwd[3], const int32_t loc[3]) {
  const int32_t off[3] = {loc[0] - cam_chunk[0], loc[1] - cam_chunk[1], loc[2] - cam_chunk[2]};
  if (off[0] >= -2 && off[0] <= 2 && off[1] >= -2 && off[1] <= 2 && off[2] >= -2 && off[2] <= 2) return 1.0e6f;
  return importance_far(off, fwd, 64.0f);
}

float orc_block_importance(const int32_t cam_chunk[3], const float fwd[3], const int32_t chunk[3], const uint8_t block[3],
                           uint32_t chunk_resolution) {
  /* ChunkManagerHelper.h:53: (ChunkLocation - CameraChunk) * (int)ChunkResolution + ivec3(BlockLocation) */
  const int32_t r = (int32_t)chunk_resolution;
  const int32_t off[3] = {(chunk[0] - cam_chunk[0]) * r + block[0], (chunk[1] - cam_chunk[1]) * r + block[1],
                          (chunk[2] - cam_chunk[2]) * r + block[2]};
  /* :55-57 compares the int offset with `-2 * ChunkResolution`, ChunkResolution being uint32_t: both sides convert to
   * unsigned, i.e. `u >= 4294967264 && u <= 32` for resolution 16, which no value satisfies -- the near branch of the
   * block variant is dead in the reference.  Restated literally (unsigned arithmetic), not "fixed". */
  const uint32_t lo = (uint32_t)(-2) * chunk_resolution, hi = 2u * chunk_resolution;
  int near_ = 1;
  for (int a = 0; a < 3; a++) {
    const uint32_t u = (uint32_t)off[a];
    if (!(u >= lo && u <= hi)) near_ = 0;
  }
  if (near_) return 1.0e6f;
  return importance_far(off, fwd, 64.0f * (float)chunk_resolution);
}

typedef struct { float imp; uint32_t order; int32_t off[3]; } Cand;
static int cand_cmp(const void* pa, const void* pb) {
  const Cand* a = (const Cand*)pa; const Cand* b = (const Cand*)pb;
  if (a->imp != b->imp) return (a->imp > b->imp) ? -1 : 1;
  return (a->order < b->order) ? -1 : (a->order > b->order);
}

int64_t orc_select_view_chunks(const float fwd[3], uint32_t forward_load, uint32_t backward_load, float view_angle_deg,
                               int mode, OrcChunkCandidate* out, int64_t cap) {
  const int32_t F = (int32_t)forward_load, B = (int32_t)backward_load;
  const float view_threshold = std_max(cosf((view_angle_deg * 0.01745329251994329576923690768489f) * 0.5f), 0.01f);
  const int64_t side = 2 * (int64_t)F + 1;
  Cand* c = (Cand*)malloc((size_t)(side * side * side) * sizeof(Cand));
  if (!c) return -1;
  float fwd_n[3];
  normalize3(fwd, fwd_n);
  const int32_t zero[3] = {0, 0, 0};
  int64_t n = 0;
  uint32_t order = 0;
  for (int32_t X = -F; X <= F; X++)
    for (int32_t Y = -F; Y <= F; Y++)
      for (int32_t Z = -F; Z <= F; Z++, order++) {
        const float o[3] = {(float)X, (float)Y, (float)Z};
        const float len = sqrtf(dot3(o, o));
        if ((double)len > (double)F + 1e-6) continue;                    /* :103 / :165 (int + double literal) */
        const int core = X >= -1 && X <= 1 && Y >= -1 && Y <= 1 && Z >= -1 && Z <= 1;
        int ok = 0;
        float imp = 0.0f;
        if (mode == 0) {                                                 /* ByView, :110-147 */
          if (core) ok = 1;
          else {
            float dir[3];
            normalize3(o, dir);
            float alpha = std_max(dot3(fwd_n, dir), 0.0f);
            const int in_cone = alpha > view_threshold;
            alpha = in_cone ? 1.0f : alpha / view_threshold;
            alpha = std_min(std_max(alpha, 0.0f), 1.0f);
            const float thr = alpha * (float)F + (1.0f - alpha) * (float)B;
            ok = len < thr;
          }
          if (ok) { const int32_t loc[3] = {X, Y, Z}; imp = orc_chunk_importance(zero, fwd, loc); }
        } else {                                                         /* Simple, :170-186 */
          if (core) { ok = 1; imp = 1.0e6f; }
          else if (len < (float)F) { ok = 1; imp = 1.0f / len; }
        }
        if (!ok) continue;
        c[n].imp = imp; c[n].order = order; c[n].off[0] = X; c[n].off[1] = Y; c[n].off[2] = Z;
        n++;
      }
  qsort(c, (size_t)n, sizeof(Cand), cand_cmp);
  for (int64_t i = 0; i < n && i < cap; i++) {
    out[i].Importance = c[i].imp; out[i].Offset[0] = c[i].off[0]; out[i].Offset[1] = c[i].off[1]; out[i].Offset[2] = c[i].off[2];
  }
  free(c);
  return n;
}

/* GetFibonacciSphere<float> (Runtimes/Helper/VoxelMathHelper.h:49-71) in the float instantiation the bake uses
 * (ChunkManagerHelper.h:205).  cosf/sinf are libm's (un-vendored third-party arithmetic). */
void orc_fibonacci_sphere_f32(uint32_t samples, float* out_xyz) {
  const float phi = (float)(3.14159265358979323846 * (sqrt(5.0) - 1.0));
  for (uint32_t i = 0; i < samples; i++) {
    const float y = 1.0f - ((float)(int32_t)i / (float)(samples - 1)) * 2.0f;
    const float radius = sqrtf(1.0f - y * y);
    const float theta = phi * (float)(int32_t)i;
    const float v[3] = {cosf(theta) * radius, y, sinf(theta) * radius};
    normalize3(v, out_xyz + 3 * i);
  }
}

/* nearest baked direction by Euclidean distance (bgi::nearest(q, 1) on points, NearestMap.h:35-37); first index wins
 * ties (the R-tree's choice among exact ties is unspecified). */
uint32_t orc_nearest_direction(const float* dirs, uint32_t n, const float q[3]) {
  uint32_t best = 0;
  float bd = INFINITY;
  for (uint32_t i = 0; i < n; i++) {
    const float d[3] = {dirs[3 * i] - q[0], dirs[3 * i + 1] - q[1], dirs[3 * i + 2] - q[2]};
    const float dd = dot3(d, d);
    if (dd < bd) { bd = dd; best = i; }
  }
  return best;
}
