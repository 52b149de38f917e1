/*
 * ref_glsl_driver.cpp -- the reference's voxel VERTEX and FRAGMENT shader text, executed on the CPU.  TEST INFRASTRUCTURE.
 *
 * ref_voxel_vs.inc / ref_voxel_fs.inc are cut out of /root/reference/Samples/SimpleVoxel.cpp at build time
 * (oracle/extract_ref_glsl.py; syntax-only edits) and compiled here against ref_shim/glsl_compat.h.  What is the
 * reference's: instance validity, chunk/block offset arithmetic, octant choice, the 56-corner table, vertex
 * positions, Projection * View * position, the fragment colour.  What is NOT the reference's and is written here from
 * the Vulkan fixed-function rules the draw relies on (SimpleVoxel.cpp:336-347,376-390): triangle-fan assembly over
 * the index list, the viewport transform with LVK's Y flip (lvk/vulkan/VulkanClasses.cpp:2350-2361), pixel-centre
 * sampling, perspective-correct varying interpolation, depth test Greater against a clear of 0, no face culling.
 * Near-plane clipping is not implemented: the call reports how many vertices had w <= 0 and the test keeps the
 * camera outside the geometry.
 *
 * Part of oracle/_ref/libmeso_ref.so; pins orc_ref_instanced_pixel (oracle/orc_raymarch.c) in tests/test_ref_pin.py.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace ref_vs {
using namespace glsl;
/* layout(location = 0..2) in uint ...  (Block.h:62-71 builds these declarations) */
static uint InstanceChunkIndex, InstanceBlockLocation, InstanceBlockFrameStamp;
static int gl_VertexIndex;
static vec4 gl_Position;
static PerVertex vtx;
static PushConstants pc;
#include "ref_voxel_vs.inc"
}  // namespace ref_vs

namespace ref_fs {
using namespace glsl;
static PerVertex vtx;
static vec4 out_FragColor;
#include "ref_voxel_fs.inc"
}  // namespace ref_fs

extern "C" {

/* One vertex-shader invocation.  out = {gl_Position[4], Normal[3], Color[3]}. */
void ref_vs_invoke(const void* camera160, const void* scene16, const void* chunks, uint32_t inst_chunk_index,
                   uint32_t inst_block_location, uint32_t inst_stamp, int vertex_index, float out[10]) {
  using namespace ref_vs;
  std::memcpy(&pc.Camera, camera160, 160);
  std::memcpy(&pc.Scene, scene16, 16);
  pc.Chunks.ChunkData = static_cast<const glsl::GPUChunk*>(chunks);
  InstanceChunkIndex = inst_chunk_index; InstanceBlockLocation = inst_block_location; InstanceBlockFrameStamp = inst_stamp;
  gl_VertexIndex = vertex_index;
  gl_Position = glsl::vec4(); vtx = glsl::PerVertex();
  vs_main();
  out[0] = gl_Position.x; out[1] = gl_Position.y; out[2] = gl_Position.z; out[3] = gl_Position.w;
  out[4] = vtx.Normal.x; out[5] = vtx.Normal.y; out[6] = vtx.Normal.z;
  out[7] = vtx.Color.x; out[8] = vtx.Color.y; out[9] = vtx.Color.z;
}

/* The whole draw: cmdDrawIndexed(8 indices, n_blocks instances) as a triangle fan, depth Greater, clear colour
 * (0,0,0,1), clear depth 0.  blocks = FGPUBlock[n] (12 B: ChunkIndex, packed location, stamp).
 * Outputs per pixel: depth (f32), instance (-1 = clear), colour (4 x f32, fragment shader output), normal (3 x f32,
 * the interpolated varying).  Returns the number of vertices with w <= 0 (must be 0 for the result to be meaningful). */
int64_t ref_draw_instanced(const void* camera160, const void* scene16, const void* chunks, const void* blocks, int64_t n_blocks,
                           const uint16_t* indices, int n_indices, int width, int height,
                           float* depth, int32_t* instance, float* color, float* normal) {
  const size_t npx = (size_t)width * height;
  for (size_t i = 0; i < npx; ++i) {
    depth[i] = 0.0f; instance[i] = -1;
    color[4 * i] = color[4 * i + 1] = color[4 * i + 2] = 0.0f; color[4 * i + 3] = 1.0f;
    normal[3 * i] = normal[3 * i + 1] = normal[3 * i + 2] = 0.0f;
  }
  int64_t behind = 0;
  const uint32_t* B = static_cast<const uint32_t*>(blocks);
  std::vector<float> v((size_t)n_indices * 10);
  for (int64_t k = 0; k < n_blocks; ++k) {
    for (int i = 0; i < n_indices; ++i)
      ref_vs_invoke(camera160, scene16, chunks, B[3 * k], B[3 * k + 1], B[3 * k + 2], (int)indices[i], &v[(size_t)i * 10]);
    /* the VS parks invalid instances at (0,0,-1,1): zero-area triangles outside the depth range, nothing is rasterised */
    double sx[16], sy[16], sz[16], iw[16];
    bool skip = false;
    for (int i = 0; i < n_indices; ++i) {
      const float* p = &v[(size_t)i * 10];
      if (!(p[3] > 0.0f)) { ++behind; skip = true; continue; }
      double w = p[3];
      sx[i] = ((double)p[0] / w + 1.0) * 0.5 * width;
      sy[i] = (1.0 - (double)p[1] / w) * 0.5 * height;     /* negative-height viewport: NDC +y is up on screen */
      sz[i] = (double)p[2] / w;
      iw[i] = 1.0 / w;
    }
    if (skip) continue;
    for (int t = 1; t + 1 < n_indices; ++t) {              /* fan: (0, t, t+1) */
      const int id[3] = {0, t, t + 1};
      double ax = sx[id[0]], ay = sy[id[0]], bx = sx[id[1]], by = sy[id[1]], cx = sx[id[2]], cy = sy[id[2]];
      double area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
      if (area == 0.0) continue;
      int x0 = (int)std::floor(std::fmin(ax, std::fmin(bx, cx))), x1 = (int)std::ceil(std::fmax(ax, std::fmax(bx, cx)));
      int y0 = (int)std::floor(std::fmin(ay, std::fmin(by, cy))), y1 = (int)std::ceil(std::fmax(ay, std::fmax(by, cy)));
      if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > width) x1 = width; if (y1 > height) y1 = height;
      for (int py = y0; py < y1; ++py)
        for (int px = x0; px < x1; ++px) {
          double qx = px + 0.5, qy = py + 0.5;
          double l0 = ((bx - qx) * (cy - qy) - (by - qy) * (cx - qx)) / area;
          double l1 = ((cx - qx) * (ay - qy) - (cy - qy) * (ax - qx)) / area;
          double l2 = 1.0 - l0 - l1;
          if (l0 < 0.0 || l1 < 0.0 || l2 < 0.0) continue;
          double z = l0 * sz[id[0]] + l1 * sz[id[1]] + l2 * sz[id[2]];
          if (z < 0.0 || z > 1.0) continue;                /* depth clip */
          size_t o = (size_t)py * width + px;
          if (!((float)z > depth[o])) continue;            /* CompareOp_Greater */
          double p0 = l0 * iw[id[0]], p1 = l1 * iw[id[1]], p2 = l2 * iw[id[2]], ps = p0 + p1 + p2;
          float nrm[3];
          for (int c = 0; c < 3; ++c)
            nrm[c] = (float)((p0 * v[(size_t)id[0] * 10 + 4 + c] + p1 * v[(size_t)id[1] * 10 + 4 + c] + p2 * v[(size_t)id[2] * 10 + 4 + c]) / ps);
          ref_fs::vtx.Normal = glsl::vec3(nrm[0], nrm[1], nrm[2]);
          ref_fs::fs_main();
          depth[o] = (float)z; instance[o] = (int32_t)k;
          color[4 * o] = ref_fs::out_FragColor.x; color[4 * o + 1] = ref_fs::out_FragColor.y;
          color[4 * o + 2] = ref_fs::out_FragColor.z; color[4 * o + 3] = ref_fs::out_FragColor.w;
          normal[3 * o] = nrm[0]; normal[3 * o + 1] = nrm[1]; normal[3 * o + 2] = nrm[2];
        }
    }
  }
  return behind;
}

}  /* extern "C" */
