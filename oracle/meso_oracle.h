/*
 * meso_oracle.h -- CPU ORACLE for the MesoEngine voxel hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may load it.
 * The product path (mesoengine_b200/, include/meso_cuda.h) never links or calls it.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures
 * and its application cannot be built or run in this image (needs glm, Boost 1.83, Vulkan, GLFW,
 * MSVC).  Its hot-path HEADERS are compiled unmodified into oracle/_ref/libmeso_ref.so against
 * stand-ins for the absent third-party headers (ref_driver.cpp, ref_shim/); the functions below that
 * restate reference code cite the file:line they follow and are pinned (a) bit for bit by that
 * build -- generators, hash/noise, index helpers, erosion/mips/cull, importance, view-cone set,
 * Fibonacci directions, re-centring (tests/test_ref_pin.py, tests/golden/ref_build.npz) -- and
 * (b) by the hand-derivable known-answer facts of SURVEY.md section 4 (tests/test_oracle_kat.py).
 * The per-pixel result is pinned by executing the reference's own GLSL text (ref_glsl_driver.cpp:
 * block, face, colour <= 1 LSB, depth 1e-5); the camera matrices (un-vendored Cookbook/glm code)
 * are restated from the published definitions only.
 * Everything the reference does not contain (voxel-in-brick level, DDA order, shadow rays,
 * face-cull + greedy merge, sphere carve) is DEFINED here: "parity unpinned by reference;
 * bit-exact vs repo oracle".
 *
 * All citations are relative to /root/reference.
 */
#ifndef MESO_ORACLE_H
#define MESO_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference layouts ---------------------------------------------------------------- */

/* Runtimes/Voxel/Block/Block.h:21-26  (12 B GPU instance; location little-endian x=bits0-7) */
typedef struct { uint32_t ChunkIndex; uint8_t BlockLocation[4]; uint32_t BlockFrameStamp; } OrcGPUBlock;
/* Runtimes/Voxel/Chunk/Chunk.h:27-31 (16 B; INT_MAX location = invalid) */
typedef struct { int32_t ChunkLocation[3]; uint32_t ChunkFrameStamp; } OrcGPUChunk;
/* Runtimes/Shader/GPUStructures.h:36-41 (160 B; glm column-major mat4: M[col*4+row]) */
typedef struct {
  float Projection[16];
  float View[16];
  int32_t CameraChunkLocation[4];
  float SubCameraLocation[4];
} OrcGPUUniformCamera;
/* Runtimes/Shader/GPUStructures.h:13-18 (16 B) */
typedef struct { float BlockSize; uint32_t BlockResolution; float ChunkSize; uint32_t ChunkResolution; } OrcGPUUniformSceneConfig;

/* ---- repo-defined output records (DESIGN.md "records") ----------------------------------- */

/* 16 B per pixel.  w0 = x | y<<16 ; w1 = z | face<<16 | shadow<<19 | hit<<20 ;
 * (x,y,z) = hit voxel in grid voxel coordinates; face: 0 -X, 1 +X, 2 -Y, 3 +Y, 4 -Z, 5 +Z,
 * 6 = ray started inside a solid voxel, 7 = miss.  Miss = {0xFFFFFFFF, 0x0007FFFF, +inf, 0xFF000000}. */
typedef struct { uint32_t w0, w1; float t; uint32_t rgba; } OrcHitRecord;
/* 16 B per quad.  w0 = x | y<<16 ; w1 = z | face<<16 | level<<19 | w<<24 ; w2 = h ; w3 = 0 (material, reserved).
 * (x,y,z) = minimum-corner voxel of the quad in grid voxel coordinates; w along u, h along v where
 * (u,v) = (y,z) for X faces, (x,z) for Y faces, (x,y) for Z faces.  level 0: voxel-level quad inside one brick
 * (w,h <= 8); level 1: brick-level quad of whole full-brick faces inside one chunk (w,h multiples of 8, <= 128). */
typedef struct { uint32_t w0, w1, w2, w3; } OrcQuad;

/* Ray setup derived on the HOST from FGPUUniformCamera (DESIGN.md "ray setup").  Same 80-byte
 * struct as MesoRaySetup in include/meso_cuda.h; restated independently here. */
typedef struct {
  float o[3];      /* eye in grid voxel coordinates                                   */
  float two_over_w;
  float U[3];      /* right / P[0][0]                                                 */
  float two_over_h;
  float V[3];      /* up / P[1][1]                                                    */
  float pad0;
  float F[3];      /* forward                                                         */
  float pad1;
  float L[3];      /* normalised direction TOWARDS the light (shadow rays)            */
  float pad2;
} OrcRaySetup;

enum { ORC_SDF_SPHERE = 0, ORC_SDF_TERRAIN = 1 };
enum { ORC_GRAN_BLOCK = 0, ORC_GRAN_VOXEL = 1 };
enum { ORC_SIN_LIBM = 0, ORC_SIN_PORTABLE = 1 };
enum { ORC_FLAG_SHADOW = 1 };
enum { ORC_DDA_FLAT = 0, ORC_DDA_HIER = 1, ORC_DDA_BOX = 2, ORC_DDA_MODEL = 3 };  /* BOX: HIER + unaligned empty cubes from a cell distance field; MODEL: see orc_step_model_* */

/* ---- a8/a9: SDFs and per-chunk generators (GeneratorHelper.h:19-150, VoxelMathHelper.h:25-33) */
double orc_sin_portable(double x);
double orc_hash3(double x, double y, double z, int sin_mode);
void   orc_noised(const double x[3], int sin_mode, double out4[4]);
double orc_displacement(const double p[3], int sin_mode);
double orc_sdf(int kind, const double params[4], int sin_mode, double x, double y, double z);
/* Returns block count; out_xyz receives (x,y,z) bytes in generator order (X outer, Z inner). */
int orc_generate_chunk(int kind, const double params[4], int sin_mode, const int32_t chunk_loc[3],
                       float block_size, int chunk_res, uint8_t* out_xyz);

/* ---- a5-a7, a10: occupancy, erode mips, hidden-block cull, instance emission ------------- */
/* mips: depth * 64 words (res fixed 16). (BinaryOccupancyVolume.h:5-99, Chunk.h:73-100) */
void orc_erode_mips(const uint8_t* blocks_xyz, int n_blocks, int depth, int use26, uint64_t* mips);
/* (ChunkPool.h:385-390,438) emits FGPUBlock for blocks with !Mip[threshold]; returns count. */
int orc_emit_instances(const uint8_t* blocks_xyz, int n_blocks, const uint64_t* mips, int threshold_depth,
                       uint32_t chunk_index, uint32_t stamp, OrcGPUBlock* out);

/* ---- a15: camera (VoxelCamera.cpp:12-59; glm 0.9.9.8 perspectiveRH_ZO / lookAtRH restated) */
void orc_perspective_rh_zo(float fovy, float aspect, float z_near, float z_far, float out16[16]);
void orc_look_at_view(const float eye[3], const float center[3], const float up[3], float out16[16]);
void orc_convert_to_chunk_location(const float pos[3], float chunk_size, float fract_out[3], int32_t chunk_out[3]);
void orc_camera_uniform(const float eye_world[3], const float center_world[3], const float up[3], float fov_deg,
                        float z_near, float z_far, int reverse_z, float width, float height, float chunk_size,
                        OrcGPUUniformCamera* out);
void orc_ray_setup(const OrcGPUUniformCamera* cam, const int32_t origin_chunk[3], int width, int height,
                   const float light_dir[3], OrcRaySetup* out);
void orc_fibonacci_sphere(uint32_t samples, int normalize, double* out_xyz);

/* ---- volume (chunk -> block mask -> 8^3 brick payload) ------------------------------------ */
typedef struct OrcVolume OrcVolume;
OrcVolume* orc_volume_create(const int32_t origin_chunk[3], const int32_t dims_chunks[3]);
void orc_volume_destroy(OrcVolume*);
void orc_volume_voxelize(OrcVolume*, int kind, const double params[4], int granularity, int sin_mode, int nthreads);
/* fast = 1: exact hierarchical classification for the sphere (see orc_volume.c); identical output, used at 4096^3. */
/* block granularity with the generator's MipmapLevel argument: one sample per (2^mip)^3 blocks (mip 0..4) */
void orc_volume_voxelize_lod(OrcVolume*, int kind, const double params[4], int sin_mode, int nthreads, int mip);
void orc_volume_voxelize_ex(OrcVolume*, int kind, const double params[4], int granularity, int sin_mode, int nthreads, int fast);
int64_t orc_volume_num_chunks(const OrcVolume*);
const uint64_t* orc_volume_occ(const OrcVolume*);   /* nchunks*64 words */
const uint64_t* orc_volume_full(const OrcVolume*);  /* nchunks*64 words */
int64_t orc_volume_num_partial(const OrcVolume*);
int64_t orc_volume_pool_slots(const OrcVolume*);   /* payload slots ever allocated (sizes the cell2 cube table) */
/* Canonical export of partial bricks sorted by (chunk, block): keys[i] = chunk*4096+block, payload 8 words each. */
int64_t orc_volume_export_partial(const OrcVolume*, uint64_t* keys, uint64_t* payload, int64_t cap);
/* Import from canonical form (used to hand the GPU a host-generated volume and vice versa). */
void orc_volume_import(OrcVolume*, const uint64_t* occ, const uint64_t* full, const uint64_t* keys,
                       const uint64_t* payload, int64_t n_partial);
int orc_volume_get_voxel(const OrcVolume*, int x, int y, int z);
int64_t orc_volume_count_voxels(const OrcVolume*);

/* K2 over the whole grid: chunk table, mips 1..3, instance list. Returns instance count. */
int64_t orc_volume_build_occupancy(const OrcVolume*, uint32_t stamp, OrcGPUChunk* chunk_table,
                                   uint64_t* mips123 /* nchunks*3*64 or NULL */, OrcGPUBlock* instances, int64_t cap);

/* ---- a13/a14: raymarch -------------------------------------------------------------------- */
typedef struct { uint64_t primary, shadow, hits, steps; uint64_t touched_chunks, touched_bricks; uint64_t u_bytes; } OrcRayStats;
/* Render the sub-rectangle [x0,x1) x [y0,y1) into records (row-major, full frame stride `width`). */
void orc_raymarch(const OrcVolume*, const OrcRaySetup*, int width, int height, int x0, int y0, int x1, int y1,
                  uint32_t flags, int dda_mode, int nthreads, OrcHitRecord* records, OrcRayStats* stats);
/* The same for an explicit list of scanlines (records still row-major width x height): one call = one parallel region
 * over (row, 128-pixel span) items -- bench.py's bounded CPU sample. */
void orc_raymarch_rows(const OrcVolume* v, const OrcRaySetup* rs, int width, int height, const int32_t* rows, int nrows,
                       uint32_t flags, int mode, int nthreads, OrcHitRecord* records, OrcRayStats* stats);
/* Reference-semantics restatement of the instanced draw (SimpleVoxel.cpp:146-192,220-224): nearest of the three
 * camera-facing faces of every valid instance along the pixel-centre ray, fp64. Returns 1 on hit.
 * out: block location in camera-chunk-relative block units, face id, t (world units), margin to nearest edge. */
int orc_ref_instanced_pixel(const OrcGPUUniformCamera* cam, const OrcGPUUniformSceneConfig* scene,
                            const OrcGPUChunk* chunks, int64_t n_chunks, const OrcGPUBlock* blocks, int64_t n_blocks,
                            int width, int height, int px, int py, int32_t out_block[3], int* out_face,
                            double* out_t, double* out_margin, float out_rgba[4]);
/* The 56-corner table and fan indices (SimpleVoxel.cpp:87-136, TriplePlanarCube.h:36-43). */
void orc_triplanar_faces(int octant, int faces_out[3]);

/* ---- K3: face cull + greedy merge --------------------------------------------------------- */
int64_t orc_mesh(const OrcVolume*, int nthreads, OrcQuad* quads, int64_t cap);
/* Voxel-level quads of the listed bricks only (keys = chunk*4096+block). */
int64_t orc_mesh_bricks(const OrcVolume*, const uint64_t* keys, int64_t n, OrcQuad* quads, int64_t cap);
/* Brick-level quads of the listed chunks only. */
int64_t orc_mesh_chunk_faces(const OrcVolume*, const int64_t* chunks, int64_t n, OrcQuad* quads, int64_t cap);
/* Number of exposed unit faces (re-expansion property check). */
int64_t orc_count_exposed_faces(const OrcVolume*);
void orc_sort_quads(OrcQuad* q, int64_t n);

/* ---- K5: sphere carve --------------------------------------------------------------------- */
/* Removes voxels whose centre lies strictly inside the sphere (integer grid-voxel centre, radius).
 * dirty receives keys (chunk*4096+block) of bricks whose contents changed, ascending. Returns count. */
int64_t orc_carve_sphere(OrcVolume*, const int32_t center[3], int32_t radius, uint64_t* dirty, int64_t cap);

/* ---- K6: resident-set selection (orc_resident.c) ------------------------ */
/* FChunkManageHelper::FTempChunkDataType = std::pair<float, ivec3> (ChunkManagerHelper.h:76), 16 B */
typedef struct { float Importance; int32_t Offset[3]; } OrcChunkCandidate;
float orc_chunk_importance(const int32_t cam_chunk[3], const float fwd[3], const int32_t loc[3]);
float orc_block_importance(const int32_t cam_chunk[3], const float fwd[3], const int32_t chunk[3], const uint8_t block[3],
                           uint32_t chunk_resolution);
/* mode 0 = GetDesiredShowChunkLocationByView, 1 = ...Simple.  Writes min(count, cap) candidates, importance descending,
 * ties in loop order; returns the full count (-1: out of memory). */
int64_t orc_select_view_chunks(const float fwd[3], uint32_t forward_load, uint32_t backward_load, float view_angle_deg,
                               int mode, OrcChunkCandidate* out, int64_t cap);
void orc_fibonacci_sphere_f32(uint32_t samples, float* out_xyz);
uint32_t orc_nearest_direction(const float* dirs, uint32_t n, const float q[3]);

/* ---- instrumentation: step-count model of acceleration structures (ORC_DDA_MODEL; orc_raymarch.c) ---------------
 * Same records as every other mode; counts steps by kind for a configurable hierarchy.  Not thread-safe to reconfigure
 * while a frame is being traced.  counts: 0 voxel, 1 2^3 cell, 2 brick, 3 field step <= 2 cells, 4 field step > 2 cells,
 * 5 grid-entry steps. */
void orc_step_model_config(int df_shift, int df_cap, int probe, int directional, int brick_cap, int cell2);
int orc_step_model_build(const OrcVolume* v);
void orc_step_model_counts(uint64_t out[6], int reset);
/* per-pixel steps of the primary and the shadow ray (2 x u32 per pixel of the full frame); NULL switches it off */
void orc_debug_set_step_image(uint32_t* img, int width);
/* The three forward-cube tables of csrc/k_cubes.cu (opt-in raymarch path), built with the GPU's algorithms in the GPU's
 * layouts: cell [8][ncells] u8, brick [nchunks*4096] u16, cell2 [pool_n*64] u16 (oracle payload-slot order).
 * orc_cube_tables_use makes ORC_DDA_MODEL read them (NULLs: back to on-the-fly cubes). */
int orc_cube_tables(const OrcVolume* v, uint8_t* cell, uint16_t* brick, uint16_t* cell2);
void orc_cube_tables_use(const OrcVolume* v, const uint8_t* cell, const uint16_t* brick, const uint16_t* cell2);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
