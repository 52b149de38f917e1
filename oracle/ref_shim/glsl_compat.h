/*
 * GLSL-in-C++ stand-in -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * oracle/extract_ref_glsl.py cuts the function bodies of the reference's voxel vertex and fragment shaders out of
 * Samples/SimpleVoxel.cpp (from `ivec3 UnpackU8Vec3` to the end of `main`, and the fragment `main`) into
 * a scratch directory (removed after the compile) with three textual changes (array constructor -> braces, `.xyz` -> `.xyz()`, `main` renamed).
 * This header supplies what that text needs to compile as C++: fp32 vectors / matrices with GLSL's operator set
 * (literals such as 0.5 are GLSL floats: every scalar operand is converted to float first), column-major mat4, and the
 * interface variables the shader's declarations (built from C++ string pieces in the reference, SimpleVoxel.cpp:41-62)
 * would have introduced.
 */
#pragma once
#include <climits>
#include <cstdint>
#include <type_traits>

namespace glsl {

typedef unsigned int uint;
template <typename S> using if_scalar = std::enable_if_t<std::is_arithmetic_v<S>>;

struct ivec3;
struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  template <typename S, typename = if_scalar<S>> explicit vec3(S s) : x((float)s), y((float)s), z((float)s) {}
  template <typename A, typename B, typename C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
};
struct ivec3 {
  int x, y, z;
  ivec3() : x(0), y(0), z(0) {}
  template <typename A, typename B, typename C> ivec3(A a, B b, C c) : x((int)a), y((int)b), z((int)c) {}
};
struct vec4 {
  float x, y, z, w;
  vec4() : x(0), y(0), z(0), w(0) {}
  template <typename A, typename B, typename C, typename D> vec4(A a, B b, C c, D d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
  template <typename D, typename = if_scalar<D>> vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
  vec3 xyz() const { return vec3(x, y, z); }
};
struct ivec4 {
  int x, y, z, w;
  ivec3 xyz() const { return ivec3(x, y, z); }
};
struct mat4 { float m[4][4]; /* m[col][row], as std430 stores it */ };

inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator*(const ivec3& a, int s) { return ivec3(a.x * s, a.y * s, a.z * s); }
/* GLSL: int operands are implicitly converted to float when the other operand is a float */
inline vec3 operator*(const ivec3& a, float s) { return vec3((float)a.x * s, (float)a.y * s, (float)a.z * s); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename S, typename = if_scalar<S>> vec3 operator*(const vec3& a, S s) { float f = (float)s; return vec3(a.x * f, a.y * f, a.z * f); }
template <typename S, typename = if_scalar<S>> vec3 operator+(const vec3& a, S s) { float f = (float)s; return vec3(a.x + f, a.y + f, a.z + f); }

/* matrix products exactly as a column-major GLSL implementation without contraction evaluates them:
 * (A*B)[c][r] = sum_k A[k][r] * B[c][k], accumulated k = 0..3 */
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 o;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = a.m[0][r] * b.m[c][0];
      for (int k = 1; k < 4; ++k) s = s + a.m[k][r] * b.m[c][k];
      o.m[c][r] = s;
    }
  return o;
}
inline vec4 operator*(const mat4& a, const vec4& v) {
  float in[4] = {v.x, v.y, v.z, v.w}, o[4];
  for (int r = 0; r < 4; ++r) {
    float s = a.m[0][r] * in[0];
    for (int k = 1; k < 4; ++k) s = s + a.m[k][r] * in[k];
    o[r] = s;
  }
  return vec4(o[0], o[1], o[2], o[3]);
}

/* the blocks the shader reaches through buffer references (GPUStructures.h:13-41, Chunk.h:27-41) */
struct CameraInfo { mat4 Projection; mat4 View; ivec4 CameraChunkLocation; vec4 SubCameraLocation; };
struct SceneInfo { float BlockSize; uint BlockResolution; float ChunkSize; uint ChunkResolution; };
struct GPUChunk { ivec3 ChunkLocation; uint ChunkFrameStamp; };
struct ChunksBuffer { const GPUChunk* ChunkData; };
struct PushConstants { CameraInfo Camera; SceneInfo Scene; ChunksBuffer Chunks; };
struct PerVertex { vec3 Normal; vec3 Color; };

static_assert(sizeof(CameraInfo) == 160 && sizeof(SceneInfo) == 16 && sizeof(GPUChunk) == 16, "std430 layouts");

}  // namespace glsl
