// MesoHost.h -- C++ host mirror of the reference's engine-side surface for the voxel path, over the C ABI.
//
// Same type / method names and argument meaning as the reference (paths relative to the reference root), so that
// Samples/SimpleVoxel reads the same and a maintainer can see exactly where lvk:: calls became meso_* calls:
//   FVoxelSceneConfig                     Runtimes/Voxel/VoxelSceneConfig.h:20-50
//   FBlock / FGPUBlock                    Runtimes/Voxel/Block/Block.h:14-26
//   FGPUChunk                             Runtimes/Voxel/Chunk/Chunk.h:27-31
//   FGPUUniformCamera / SceneConfig       Runtimes/Shader/GPUStructures.h:13-56
//   FVoxelMathHelper::ConvertToChunkLocation  Runtimes/Helper/VoxelMathHelper.h:17-22
//   FVoxelCamera                          Runtimes/Instance/VoxelCamera.{h,cpp}
//   FChunkManage (facade)                 Runtimes/Voxel/Chunk/ChunkManager.h:90-102,134-159,211-400
//   VoxelWindowsInstance (headless)       Runtimes/Instance/VoxelWindowsInstance.{h,cpp}: Initialize, RunInstance, hooks
// Header-only like the reference's Runtimes.  No glm/boost/GLFW/Vulkan: a 40-line vector layer replaces glm here.
// There is no CPU compute in this layer: generation, occupancy, meshing and visibility all run in libmeso_b200.so.
#pragma once
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/meso_cuda.h"

namespace meso {

struct vec3 { float x = 0, y = 0, z = 0; };
struct ivec3 { int32_t x = 0, y = 0, z = 0; bool operator!=(const ivec3& o) const { return x != o.x || y != o.y || z != o.z; } };
struct mat4 { float m[16] = {0}; float& at(int col, int row) { return m[col * 4 + row]; } float at(int col, int row) const { return m[col * 4 + row]; } };
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline vec3 normalize(vec3 v) { const float inv = 1.0f / std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z); return {v.x * inv, v.y * inv, v.z * inv}; }

inline void Check(int code, const char* what) {
  if (code != MESO_OK) throw std::runtime_error(std::string(what) + ": " + meso_last_error());
}

enum class EChunkOverrideMode : uint8_t { FindLess = 1 << 0, FindMin = 1 << 1, OverrideMin = 1 << 2 };

struct FVoxelSceneConfig {
  unsigned char BlockResolution = 8;
  float BlockSize = 1.0f;
  unsigned char ChunkResolution = 16;
  uint32_t MaxBlockCount = 65536 * 16;
  uint32_t MaxVolumeCount = 65536 * 16;   // brick payload pool (meso_scene_create max_bricks)
  uint32_t MaxChunkCount = 8192 * 2;
  uint32_t MaxEmptyChunkCount = 8192 * 4;
  uint32_t ChunkOccupancyDepth = 4;
  uint32_t ChunkInnerVoxelCullDepthThreshold = 1;
  EChunkOverrideMode ChunkOverrideMode = EChunkOverrideMode::FindMin;  // unused: everything in the window is resident
  float GetChunkSize() const { return ChunkResolution * BlockSize; }
};

struct FBlock { uint32_t ChunkIndex = INT_MAX; uint8_t BlockLocation[3] = {255u, 255u, 255u}; uint32_t VolumeIndex = INT_MAX; };
using FGPUBlock = MesoGPUBlock;
using FGPUChunk = MesoGPUChunk;
using FGPUUniformCamera = MesoGPUUniformCamera;
using FGPUUniformSceneConfig = MesoGPUUniformSceneConfig;
static_assert(sizeof(FGPUBlock) == 12 && sizeof(FGPUChunk) == 16 && sizeof(FGPUUniformCamera) == 160, "reference layouts");

struct FVoxelMathHelper {
  static void ConvertToChunkLocation(vec3 Position, float ChunkSize, vec3& Fracted, ivec3& Chunk) {
    const float cx = std::floor(Position.x / ChunkSize), cy = std::floor(Position.y / ChunkSize), cz = std::floor(Position.z / ChunkSize);
    Fracted = {Position.x - cx * ChunkSize, Position.y - cy * ChunkSize, Position.z - cz * ChunkSize};
    Chunk = {(int32_t)cx, (int32_t)cy, (int32_t)cz};
  }
};

// First-person camera: orientation fixed by lookAt(position, target, up) at construction (Cookbook
// CameraPositioner_FirstPerson), reverse-Z perspective, chunk-relative re-centring with the two callbacks.
class FVoxelCamera {
 public:
  vec3 Position{5.0f, 2.0f, 2.0f};
  mat4 Orientation;
  float Fov = float(45.0f * (M_PI / 180.0f));
  float Near = 0.1f, Far = 1000.0f;
  bool bReverseZ = true;
  ivec3 CameraChunkLocation{0, 0, 0};
  vec3 CameraForward{0, 0, 0};
  std::function<void()> CameraChunkUpdateCallback, CameraUpdateCallback;

  FVoxelCamera() { SetPose({5.0f, 2.0f, 2.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 1.0f}); }
  void SetPose(vec3 position, vec3 target, vec3 up) {
    Position = position;
    const vec3 f = normalize(target - position), s = normalize(cross(f, up)), u = cross(s, f);
    Orientation = mat4();
    Orientation.at(0, 0) = s.x; Orientation.at(1, 0) = s.y; Orientation.at(2, 0) = s.z;
    Orientation.at(0, 1) = u.x; Orientation.at(1, 1) = u.y; Orientation.at(2, 1) = u.z;
    Orientation.at(0, 2) = -f.x; Orientation.at(1, 2) = -f.y; Orientation.at(2, 2) = -f.z;
    Orientation.at(3, 3) = 1.0f;
  }
  void InitializeVoxelCamera(float FovAngle_ = 45.0f, float Near_ = 0.5f, float Far_ = 1000.0f, float bReverseZ_ = true) {
    Fov = float(FovAngle_ * (M_PI / 180.0f)); bReverseZ = bReverseZ_; Near = Near_; Far = Far_;
  }
  mat4 GetProjectionMatrix(float ViewWidth = 1920.0f, float ViewHeight = 1280.0f) const {
    const float AspectRatio = ViewWidth / ViewHeight;
    const float zNear = bReverseZ ? Far : Near, zFar = bReverseZ ? Near : Far;  // glm perspectiveRH_ZO
    const float t = std::tan(Fov / 2.0f);
    mat4 P;
    P.at(0, 0) = 1.0f / (AspectRatio * t); P.at(1, 1) = 1.0f / t;
    P.at(2, 2) = zFar / (zNear - zFar); P.at(2, 3) = -1.0f; P.at(3, 2) = -(zFar * zNear) / (zFar - zNear);
    return P;
  }
  mat4 GetViewMatrix() const {
    mat4 V = Orientation;
    const float tx = -Position.x, ty = -Position.y, tz = -Position.z;
    for (int r = 0; r < 3; r++) V.at(3, r) = ((V.at(0, r) * tx + V.at(1, r) * ty) + V.at(2, r) * tz) + 0.0f;
    return V;
  }
  vec3 GetForwardVector() const { const mat4 V = GetViewMatrix(); return {-V.at(0, 2), -V.at(1, 2), -V.at(2, 2)}; }
  FGPUUniformCamera GetCameraUniform(float ViewWidth = 1920.0f, float ViewHeight = 1280.0f) const {
    FGPUUniformCamera u{};
    const mat4 P = GetProjectionMatrix(ViewWidth, ViewHeight), V = GetViewMatrix();
    for (int i = 0; i < 16; i++) { u.Projection[i] = P.m[i]; u.View[i] = V.m[i]; }
    u.CameraChunkLocation[0] = CameraChunkLocation.x; u.CameraChunkLocation[1] = CameraChunkLocation.y; u.CameraChunkLocation[2] = CameraChunkLocation.z;
    u.SubCameraLocation[0] = (float)(int32_t)Position.x; u.SubCameraLocation[1] = (float)(int32_t)Position.y; u.SubCameraLocation[2] = (float)(int32_t)Position.z;
    return u;
  }
  void UpdateCamera(const FVoxelSceneConfig& CurrentSceneConfig) {
    vec3 fr; ivec3 off;
    FVoxelMathHelper::ConvertToChunkLocation(Position, CurrentSceneConfig.GetChunkSize(), fr, off);
    const bool moved = off != ivec3{0, 0, 0};
    if (moved) {
      if (CameraChunkUpdateCallback) CameraChunkUpdateCallback();
      CameraChunkLocation = {CameraChunkLocation.x + off.x, CameraChunkLocation.y + off.y, CameraChunkLocation.z + off.z};
    }
    const vec3 fwd = GetForwardVector();
    if (fwd.x != CameraForward.x || fwd.y != CameraForward.y || fwd.z != CameraForward.z || moved) {
      CameraForward = fwd;
      if (CameraUpdateCallback) CameraUpdateCallback();
    }
    Position = fr;
  }
};

// What the generator plug-in point (GeneratorType, ChunkManager.h:61) becomes: the SDF is evaluated on the device, so the
// "generator" is a description, not a callback.
struct FGeneratorDesc {
  int Kind = MESO_SDF_SPHERE;                       // FGeneratorHelper::GenerateSphere / TestGenerator
  double Params[4] = {100.0, 0.0, 0.0, 50.0};       // GeneratorHelper.h:134
  int Granularity = MESO_GRAN_BLOCK;                // reference: one sample per block
};

// Facade with FChunkManage's shape.  The resident set is a fixed window of chunks [WindowOrigin, WindowOrigin+WindowDims)
// ("everything resident"); UpdateLoadingQueue regenerates it on the device when dirty instead of dispatching CPU workers.
class FChunkManage {
 public:
  struct FChunkPoolView {             // the public buffers of FChunkPool (ChunkPool.h:222-223), now device-side counts
    uint32_t MaxBlockCount = 0; int64_t CurrentBlockCount = 0; int64_t ChunkCount = 0;
  } ChunkPool;
  ivec3 WindowOrigin{0, 0, 0}, WindowDims{1, 1, 1};
  uint32_t FrameStamp = 1;
  bool bDirty = true;

  void Initialize(MesoCtx* Ctx_, const FVoxelSceneConfig& VoxelSceneConfig, FGeneratorDesc Generator_, ivec3 WindowOrigin_, ivec3 WindowDims_) {
    Ctx = Ctx_; Generator = Generator_; WindowOrigin = WindowOrigin_; WindowDims = WindowDims_;
    const FGPUUniformSceneConfig cfg{VoxelSceneConfig.BlockSize, (uint32_t)VoxelSceneConfig.BlockResolution, VoxelSceneConfig.GetChunkSize(), (uint32_t)VoxelSceneConfig.ChunkResolution};
    const int32_t o[3] = {WindowOrigin.x, WindowOrigin.y, WindowOrigin.z}, d[3] = {WindowDims.x, WindowDims.y, WindowDims.z};
    Check(meso_scene_create(Ctx, &cfg, o, d, VoxelSceneConfig.MaxVolumeCount), "meso_scene_create");
    ChunkPool.MaxBlockCount = VoxelSceneConfig.MaxBlockCount;
    ChunkPool.ChunkCount = (int64_t)d[0] * d[1] * d[2];
    bDirty = true;
  }
  // camera moved to another chunk / turned: the window is static here, so only the stamp advances (ChunkManager.h:134-159)
  void UpdateChunks(ivec3, vec3, const FVoxelSceneConfig&) { FrameStamp++; }
  // ChunkManager.h:211-400: generate what is missing, then publish chunk table + block instances (K1 + K2 on the device)
  void UpdateLoadingQueue(const FVoxelSceneConfig&, uint32_t /*RenderFrameIndex*/) {
    if (!bDirty) return;
    Check(meso_voxelize_sdf(Ctx, Generator.Kind, Generator.Params, Generator.Granularity), "meso_voxelize_sdf");
    Check(meso_build_occupancy(Ctx, FrameStamp, &ChunkPool.CurrentBlockCount), "meso_build_occupancy");
    bDirty = false;
  }
 private:
  MesoCtx* Ctx = nullptr;
  FGeneratorDesc Generator;
};

struct VoxelInstanceInitialConfig {
  float CameraFOV = 60.0f, CameraNear = 0.1f, CameraFar = 1000.0f;
  int WindowsWidth = 1280, WindowsHeight = 720;
  bool bReverseZ = true;
  uint32_t kNumBufferedFrames = 4;
  int Device = 0;
  FVoxelSceneConfig VoxelSceneConfig;
};

// Headless frame loop with the reference's virtual hooks and call order (VoxelWindowsInstance.cpp:368-393):
// UpdateCamera -> UpdatePhysics -> RenderStart -> Render -> RenderEnd -> UpdateFrameIndex.
class VoxelWindowsInstance {
 public:
  MesoCtx* Context = nullptr;                     // stands where std::unique_ptr<lvk::IContext> LVKContext stood
  int WindowsWidth = 0, WindowsHeight = 0;
  bool bLVKReverseZ = true;
  uint32_t LVKNumBufferedFrames = 3;
  FVoxelSceneConfig VoxelSceneConfig;
  FVoxelCamera WindowsCamera;
  FGPUUniformSceneConfig SceneConfig{};
  std::vector<FGPUUniformCamera> UBOCamera;       // per-frame camera ring (VoxelWindowsInstance.cpp:113-120)
  std::vector<MesoHitRecord> OffscreenRecords;    // stands where TEXOffscreenColor stood (16 B records instead of RGBA8)
  uint32_t RenderGlobalFrameIndex = 0, RenderFrameIndex = 0;

  virtual ~VoxelWindowsInstance() { if (Context) meso_ctx_destroy(Context); }
  virtual void Initialize(const VoxelInstanceInitialConfig& InitialConfig) {
    bLVKReverseZ = InitialConfig.bReverseZ; LVKNumBufferedFrames = InitialConfig.kNumBufferedFrames;
    WindowsWidth = InitialConfig.WindowsWidth; WindowsHeight = InitialConfig.WindowsHeight;
    VoxelSceneConfig = InitialConfig.VoxelSceneConfig;
    InitializeCameraAndScene(InitialConfig);
    Device = InitialConfig.Device;
    InitializeContext();
    InitializeBegin();
    CreateWindowsFrameBuffer();
    InitializeRender();
  }
  virtual void InitializeCameraAndScene(const VoxelInstanceInitialConfig& InitialConfig) {
    WindowsCamera.InitializeVoxelCamera(InitialConfig.CameraFOV, InitialConfig.CameraNear, InitialConfig.CameraFar, bLVKReverseZ);
    WindowsCamera.CameraChunkUpdateCallback = [this]() { WhenCameraChunkUpdate(); };
    WindowsCamera.CameraUpdateCallback = [this]() { WhenCameraUpdate(); };
    SceneConfig = {InitialConfig.VoxelSceneConfig.BlockSize, (uint32_t)InitialConfig.VoxelSceneConfig.BlockResolution,
                   InitialConfig.VoxelSceneConfig.GetChunkSize(), (uint32_t)InitialConfig.VoxelSceneConfig.ChunkResolution};
  }
  virtual void InitializeContext() {               // lvk::createVulkanContextWithSwapchain -> meso_ctx_create
    Check(meso_ctx_create(Device, &Context), "meso_ctx_create");
    UBOCamera.assign(LVKNumBufferedFrames, FGPUUniformCamera{});
  }
  virtual void InitializeBegin() {}
  virtual void CreateWindowsFrameBuffer() { OffscreenRecords.assign((size_t)WindowsWidth * WindowsHeight, MesoHitRecord{}); }
  virtual void InitializeRender() { RenderFrameIndex = 0; RenderGlobalFrameIndex = 0; }
  virtual void RunInstance(uint32_t Frames) {
    for (uint32_t f = 0; f < Frames; f++) {
      UpdateCamera(); UpdatePhysics(); RenderStart(); Render(); RenderEnd(); UpdateFrameIndex();
    }
  }
  virtual void UpdateCamera() { WindowsCamera.UpdateCamera(VoxelSceneConfig); }
  virtual void UpdatePhysics() {}
  virtual void UpdateFrameIndex() { RenderGlobalFrameIndex++; RenderFrameIndex = RenderGlobalFrameIndex % LVKNumBufferedFrames; }
  virtual void RenderStart() { UBOCamera[RenderFrameIndex] = WindowsCamera.GetCameraUniform((float)WindowsWidth, (float)WindowsHeight); }
  virtual void Render() {}
  virtual void RenderEnd() {}
  virtual void WhenCameraChunkUpdate() {}
  virtual void WhenCameraUpdate() {}
 protected:
  int Device = 0;
};

}  // namespace meso
