/* orc_camera.c -- CPU ORACLE (test infrastructure): camera matrices and ray setup.
 * Restates Runtimes/Instance/VoxelCamera.cpp:12-59, Runtimes/Helper/VoxelMathHelper.h:17-22 and the
 * un-vendored third-party pieces they call: glm 0.9.9.8 perspectiveRH_ZO / lookAtRH / translate and
 * 3D-Graphics-Rendering-Cookbook shared/Camera.h CameraPositioner_FirstPerson::getViewMatrix
 * (restated from their published definitions; parity unpinned -- the kernels and this oracle both
 * consume the explicit P and V matrices, so camera rounding never enters a parity comparison). */
#include "orc_internal.h"

/* glm/ext/matrix_clip_space.inl perspectiveRH_ZO (GLM_FORCE_DEPTH_ZERO_TO_ONE, CMakeLists.txt:5) */
void orc_perspective_rh_zo(float fovy, float aspect, float zNear, float zFar, float m[16]) {
  float tanHalfFovy = tanf(fovy / 2.0f);
  memset(m, 0, sizeof(float) * 16);
  m[0 * 4 + 0] = 1.0f / (aspect * tanHalfFovy);
  m[1 * 4 + 1] = 1.0f / (tanHalfFovy);
  m[2 * 4 + 2] = zFar / (zNear - zFar);
  m[2 * 4 + 3] = -1.0f;
  m[3 * 4 + 2] = -(zFar * zNear) / (zFar - zNear);
}

static void normalize3(float v[3]) { /* glm: v * inversesqrt(dot(v,v)) */
  float inv = 1.0f / sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
}
static void cross3(const float x[3], const float y[3], float r[3]) {
  r[0] = x[1] * y[2] - y[1] * x[2];
  r[1] = x[2] * y[0] - y[2] * x[0];
  r[2] = x[0] * y[1] - y[0] * x[1];
}

/* View = mat4_cast(orientation) * translate(-position) with orientation = lookAtRH(eye, center, up)
 * (Cookbook CameraPositioner_FirstPerson; the mat->quat->mat round trip is skipped). */
void orc_look_at_view(const float eye[3], const float center[3], const float up[3], float m[16]) {
  float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
  normalize3(f);
  float s[3]; cross3(f, up, s); normalize3(s);
  float u[3]; cross3(s, f, u);
  memset(m, 0, sizeof(float) * 16);
  m[0 * 4 + 0] = s[0]; m[1 * 4 + 0] = s[1]; m[2 * 4 + 0] = s[2];
  m[0 * 4 + 1] = u[0]; m[1 * 4 + 1] = u[1]; m[2 * 4 + 1] = u[2];
  m[0 * 4 + 2] = -f[0]; m[1 * 4 + 2] = -f[1]; m[2 * 4 + 2] = -f[2];
  m[15] = 1.0f;
  /* r * translate(-eye): column 3 = ((r0*tx + r1*ty) + r2*tz) + r3 */
  float t[3] = {-eye[0], -eye[1], -eye[2]};
  for (int r = 0; r < 3; r++) m[12 + r] = ((m[0 * 4 + r] * t[0] + m[1 * 4 + r] * t[1]) + m[2 * 4 + r] * t[2]) + 0.0f;
}

/* VoxelMathHelper.h:17-22 */
void orc_convert_to_chunk_location(const float pos[3], float chunk_size, float fract_out[3], int32_t chunk_out[3]) {
  for (int i = 0; i < 3; i++) {
    float c = floorf(pos[i] / chunk_size);
    fract_out[i] = pos[i] - c * chunk_size;
    chunk_out[i] = (int32_t)c;
  }
}

/* VoxelCamera.cpp:4-10,12-17,31-40,42-59 for a camera placed at eye_world looking at center_world. */
void orc_camera_uniform(const float eye_world[3], const float center_world[3], const float up[3], float fov_deg,
                        float z_near, float z_far, int reverse_z, float width, float height, float chunk_size,
                        OrcGPUUniformCamera* out) {
  float fov = (float)(fov_deg * (M_PI / 180.0f));
  float aspect = (float)width / (float)height;
  orc_perspective_rh_zo(fov, aspect, reverse_z ? z_far : z_near, reverse_z ? z_near : z_far, out->Projection);
  float fr[3]; int32_t ch[3];
  orc_convert_to_chunk_location(eye_world, chunk_size, fr, ch);
  /* the orientation is fixed at construction from the un-recentred pose; UpdateCamera only moves the position */
  float tmp[16];
  orc_look_at_view(eye_world, center_world, up, tmp);
  memcpy(out->View, tmp, sizeof(tmp));
  float t[3] = {-fr[0], -fr[1], -fr[2]};
  for (int r = 0; r < 3; r++)
    out->View[12 + r] = ((tmp[0 * 4 + r] * t[0] + tmp[1 * 4 + r] * t[1]) + tmp[2 * 4 + r] * t[2]) + 0.0f;
  for (int i = 0; i < 3; i++) {
    out->CameraChunkLocation[i] = ch[i];
    out->SubCameraLocation[i] = (float)(int32_t)fr[i]; /* VoxelCamera.cpp:38 ivec4(getPosition(), 0.0) */
  }
  out->CameraChunkLocation[3] = 0;
  out->SubCameraLocation[3] = 0.0f;
}

/* DESIGN.md "ray setup": pixel-centre rays from P and V (viewport Y-flip per
 * ThirdParty/lightweightvk/lvk/vulkan/VulkanClasses.cpp:2350-2361: NDC +Y is the top row). */
void orc_ray_setup(const OrcGPUUniformCamera* cam, const int32_t origin_chunk[3], int width, int height,
                   const float light_dir[3], OrcRaySetup* rs) {
  const float* V = cam->View;
  const float* P = cam->Projection;
  memset(rs, 0, sizeof(*rs));
  float t0 = V[12], t1 = V[13], t2 = V[14];
  for (int i = 0; i < 3; i++) {
    float e = -((V[i * 4 + 0] * t0 + V[i * 4 + 1] * t1) + V[i * 4 + 2] * t2); /* -(R^T t) */
    float off = (float)((cam->CameraChunkLocation[i] - origin_chunk[i]) * ORC_CV);
    rs->o[i] = e * 8.0f + off;
    rs->U[i] = V[i * 4 + 0] / P[0];
    rs->V[i] = V[i * 4 + 1] / P[5];
    rs->F[i] = -V[i * 4 + 2];
  }
  rs->two_over_w = 2.0f / (float)width;
  rs->two_over_h = 2.0f / (float)height;
  float inv = 1.0f / sqrtf((light_dir[0] * light_dir[0] + light_dir[1] * light_dir[1]) + light_dir[2] * light_dir[2]);
  for (int i = 0; i < 3; i++) rs->L[i] = light_dir[i] * inv;
}
