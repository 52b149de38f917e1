/*
 * Minimal stand-in for glm 0.9.9.8 -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * The reference's hot-path headers (Runtimes/Helper/GeneratorHelper.h, VoxelMathHelper.h,
 * Runtimes/Voxel/Chunk/Chunk.h, ChunkManagerHelper.h, Occupancy/BinaryOccupancyVolume.h ...) are compiled
 * UNMODIFIED from /root/reference; glm itself is an un-vendored dependency, so the
 * handful of glm names those headers use are provided here, written from glm's public definitions:
 * component-wise arithmetic, dot = (x*x' + y*y') + z*z' (left to right), length = sqrt(dot(v, v)),
 * normalize = v * (1 / sqrt(dot(v, v))) (glm::inversesqrt), floor / fract component-wise.
 * Nothing here is derived from the repo's own oracle (oracle/orc_*.c): the two are independent
 * statements that tests/test_ref_pin.py compares.
 */
#pragma once
#include <math.h>   /* also brings the float overloads of cos/sin into the global namespace, as MSVC's <cmath> does */
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>
#include <type_traits>

namespace glm {

enum qualifier { packed_highp, defaultp = packed_highp, highp = packed_highp };

template <int L, typename T, qualifier Q = defaultp> struct vec;

template <typename T, qualifier Q>
struct vec<2, T, Q> {
  T x, y;
  constexpr vec() : x(0), y(0) {}
  constexpr explicit vec(T s) : x(s), y(s) {}
  template <typename A, typename B> constexpr vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
  template <typename U, qualifier P> constexpr explicit vec(const vec<2, U, P>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
};

template <typename T, qualifier Q>
struct vec<3, T, Q> {
  T x, y, z;
  constexpr vec() : x(0), y(0), z(0) {}
  constexpr vec(const vec&) = default;
  constexpr vec& operator=(const vec&) = default;
  constexpr explicit vec(T s) : x(s), y(s), z(s) {}
  template <typename A, typename B, typename C>
  constexpr vec(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
  template <typename U, qualifier P, typename = std::enable_if_t<!std::is_same_v<U, T>>>
  constexpr explicit vec(const vec<3, U, P>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  template <typename U> vec& operator+=(const vec<3, U, Q>& v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); z += static_cast<T>(v.z); return *this; }
  template <typename U> vec& operator-=(const vec<3, U, Q>& v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); z -= static_cast<T>(v.z); return *this; }
  template <typename U> vec& operator*=(const vec<3, U, Q>& v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); z *= static_cast<T>(v.z); return *this; }
  template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>> vec& operator+=(U s) { x += static_cast<T>(s); y += static_cast<T>(s); z += static_cast<T>(s); return *this; }
  template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>> vec& operator*=(U s) { x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s); return *this; }
  template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>> vec& operator/=(U s) { x /= static_cast<T>(s); y /= static_cast<T>(s); z /= static_cast<T>(s); return *this; }
};

template <typename T, qualifier Q>
struct vec<4, T, Q> {
  T x, y, z, w;
  constexpr vec() : x(0), y(0), z(0), w(0) {}
  constexpr explicit vec(T s) : x(s), y(s), z(s), w(s) {}
  template <typename A, typename B, typename C, typename D>
  constexpr vec(A a, B b, C c, D d) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
  template <typename A, typename U, qualifier P, typename = std::enable_if_t<std::is_arithmetic_v<A>>>
  constexpr vec(A a, const vec<3, U, P>& v) : x(static_cast<T>(a)), y(static_cast<T>(v.x)), z(static_cast<T>(v.y)), w(static_cast<T>(v.z)) {}
  template <typename D, typename U, qualifier P, typename = std::enable_if_t<std::is_arithmetic_v<D>>>
  constexpr vec(const vec<3, U, P>& v, D d) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(d)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T, qualifier Q = defaultp> using tvec2 = vec<2, T, Q>;
template <typename T, qualifier Q = defaultp> using tvec3 = vec<3, T, Q>;
template <typename T, qualifier Q = defaultp> using tvec4 = vec<4, T, Q>;

using vec2 = vec<2, float>;   using vec3 = vec<3, float>;   using vec4 = vec<4, float>;
using dvec2 = vec<2, double>; using dvec3 = vec<3, double>; using dvec4 = vec<4, double>;
using ivec2 = vec<2, int>;    using ivec3 = vec<3, int>;    using ivec4 = vec<4, int>;
using uvec3 = vec<3, unsigned>; using uvec4 = vec<4, unsigned>;
using u8vec3 = vec<3, uint8_t>; using u8vec4 = vec<4, uint8_t>;

/* column-major 4x4, value-initialised to zero by `= {}` as in glm (GPUStructures.h:38-39 only stores it) */
template <typename T> struct tmat4 { vec<4, T> c[4]; vec<4, T>& operator[](int i) { return c[i]; } const vec<4, T>& operator[](int i) const { return c[i]; } };
using mat4 = tmat4<float>;

/* ---- component-wise operators (vec3) ---- */
#define GLM_SHIM_BINOP(op)                                                                                              \
  template <typename T, qualifier Q> constexpr vec<3, T, Q> operator op(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { \
    return vec<3, T, Q>(a.x op b.x, a.y op b.y, a.z op b.z);                                                            \
  }                                                                                                                     \
  template <typename T, qualifier Q> constexpr vec<3, T, Q> operator op(const vec<3, T, Q>& a, T s) {                   \
    return vec<3, T, Q>(a.x op s, a.y op s, a.z op s);                                                                  \
  }                                                                                                                     \
  template <typename T, qualifier Q> constexpr vec<3, T, Q> operator op(T s, const vec<3, T, Q>& a) {                   \
    return vec<3, T, Q>(s op a.x, s op a.y, s op a.z);                                                                  \
  }
GLM_SHIM_BINOP(+)
GLM_SHIM_BINOP(-)
GLM_SHIM_BINOP(*)
GLM_SHIM_BINOP(/)
#undef GLM_SHIM_BINOP
/* mixed scalar types that appear in the reference: `tvec3<T> / float` (VoxelMathHelper.h:19),
 * `tvec3<T> * float` (:20), `dvec3 * .1` etc. already match T; float-with-double needs the promotion glm does not do
 * either (glm converts the scalar to T) */
template <typename T, qualifier Q, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>>
constexpr vec<3, T, Q> operator/(const vec<3, T, Q>& a, S s) { return a / static_cast<T>(s); }
template <typename T, qualifier Q, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>>
constexpr vec<3, T, Q> operator*(const vec<3, T, Q>& a, S s) { return a * static_cast<T>(s); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator-(const vec<3, T, Q>& a) { return vec<3, T, Q>(-a.x, -a.y, -a.z); }
template <typename T, qualifier Q> constexpr bool operator==(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T, qualifier Q> constexpr bool operator!=(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { return !(a == b); }

/* ---- functions ---- */
template <typename T, qualifier Q> vec<3, T, Q> floor(const vec<3, T, Q>& v) { return vec<3, T, Q>(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
template <typename T, qualifier Q> vec<3, T, Q> fract(const vec<3, T, Q>& v) { return v - floor(v); }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>> T fract(T v) { return v - std::floor(v); }
template <typename T, qualifier Q> T dot(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { vec<3, T, Q> t(a * b); return t.x + t.y + t.z; }
template <typename T, qualifier Q> T dot(const vec<2, T, Q>& a, const vec<2, T, Q>& b) { return a.x * b.x + a.y * b.y; }
template <typename T, qualifier Q> T length(const vec<3, T, Q>& v) { return std::sqrt(dot(v, v)); }
template <typename T> T inversesqrt(T x) { return static_cast<T>(1) / std::sqrt(x); }
template <typename T, qualifier Q> vec<3, T, Q> normalize(const vec<3, T, Q>& v) { return v * inversesqrt(dot(v, v)); }
template <typename T> constexpr T radians(T degrees) { return degrees * static_cast<T>(0.01745329251994329576923690768489); }
template <typename T> constexpr T degrees(T radians) { return radians * static_cast<T>(57.295779513082320876798154814105); }

}  // namespace glm
