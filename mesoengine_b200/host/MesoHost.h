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
  uint32_t MaxChunkCheckTimes = 128;       // unused (pool probing, ChunkPool.h:447-622): the window holds every chunk it covers;
  uint32_t MaxEmptyChunkCheckTimes = 128;  // kept so that code written against the reference's struct compiles unchanged
  uint32_t MaxBlockCheckTimes = 16;        // (oracle/ref_host_mirror_check.cpp compares this struct with the reference's)
  uint32_t BakeVisibilityViewNum = 256;   // > 0: snap the forward vector to the nearest of this many baked directions
  uint32_t ViewForwardLoadChunkSize = 24;
  uint32_t ViewBackwardLoadChunkSize = 6;
  uint32_t MaxSyncedLoadChunkCount = 0;
  uint32_t MaxUnsyncedLoadChunkCount = 256;  // chunks generated per UpdateLoadingQueue (one device launch, no CPU workers)
  uint32_t ChunkTaskPerCore = 8;             // unused: batching is the launch itself
  EChunkOverrideMode ChunkOverrideMode = EChunkOverrideMode::FindMin;  // unused: everything in the window is resident
  float ViewChunkAngle = 120.0f;
  uint32_t ChunkOccupancyDepth = 4;
  uint32_t ChunkInnerVoxelCullDepthThreshold = 1;
  float GetChunkSize() const { return ChunkResolution * BlockSize; }
  MesoViewConfig GetViewConfig() const { return MesoViewConfig{ViewForwardLoadChunkSize, ViewBackwardLoadChunkSize, ViewChunkAngle, 0u}; }
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

// FImportanceComputeInfo / FChunkManageHelper (ChunkManagerHelper.h:22-198) over the C ABI (K6 runs on the device).
struct FImportanceComputeInfo {
  ivec3 CameraChunk{0, 0, 0};
  vec3 CameraForwardVector{};
  // CalculateChunkImportance for a batch of absolute chunk locations (ChunkManagerHelper.h:26-48)
  std::vector<float> CalculateChunkImportance(MesoCtx* Ctx, const std::vector<ivec3>& ChunkLocations) const {
    static_assert(sizeof(ivec3) == 12, "ivec3 is three packed int32");
    std::vector<float> out(ChunkLocations.size());
    const int32_t cc[3] = {CameraChunk.x, CameraChunk.y, CameraChunk.z};
    const float f[3] = {CameraForwardVector.x, CameraForwardVector.y, CameraForwardVector.z};
    Check(meso_chunk_importance(Ctx, cc, f, reinterpret_cast<const int32_t*>(ChunkLocations.data()), (int64_t)ChunkLocations.size(), out.data()),
          "meso_chunk_importance");
    return out;
  }
};
struct FChunkManageHelper {
  using FTempChunkDataType = MesoChunkCandidate;  // <Importance, ChunkLocation offset>
  // The reference's priority queue, already in pop order (ChunkManagerHelper.h:89-150)
  static std::vector<FTempChunkDataType> GetDesiredShowChunkLocationByView(MesoCtx* Ctx, vec3 ForwardVector, const FVoxelSceneConfig& VoxelSceneConfig) {
    const MesoViewConfig vc = VoxelSceneConfig.GetViewConfig();
    const int64_t side = 2 * (int64_t)vc.ViewForwardLoadChunkSize + 1;
    std::vector<FTempChunkDataType> out((size_t)(side * side * side));
    const float f[3] = {ForwardVector.x, ForwardVector.y, ForwardVector.z};
    int64_t n = 0;
    Check(meso_select_view_chunks(Ctx, f, &vc, out.data(), (int64_t)out.size(), &n), "meso_select_view_chunks");
    out.resize((size_t)n);
    return out;
  }
};

// Facade with FChunkManage's shape over a window of chunks [WindowOrigin, WindowOrigin+WindowDims) that stands where the
// chunk pool stood.  Two modes:
//   bStreaming = false: everything in the window is generated at the first UpdateLoadingQueue ("everything resident");
//   bStreaming = true : the reference's loop -- UpdateChunks records the view, UpdateLoadingQueue generates the most
//                       important missing chunks of the desired set, at most MaxSynced+MaxUnsyncedLoadChunkCount per
//                       call, on the device (meso_stream_update) instead of dispatching CPU workers.
class FChunkManage {
 public:
  struct FChunkPoolView {             // the public buffers of FChunkPool (ChunkPool.h:222-223), now device-side counts
    uint32_t MaxBlockCount = 0; int64_t CurrentBlockCount = 0; int64_t ChunkCount = 0;
  } ChunkPool;
  ivec3 WindowOrigin{0, 0, 0}, WindowDims{1, 1, 1};
  uint32_t FrameStamp = 1;
  bool bDirty = true;
  bool bStreaming = false;
  bool bDebugDisableUpdateChunk = false;                                   // ChunkManager.h:76
  uint32_t DebugVisibleChunkNum = 0, DebugLoadedChunkNum = 0, DebugMissingChunkNum = 0;  // ChunkManager.h:66-75 counters

  void Initialize(MesoCtx* Ctx_, const FVoxelSceneConfig& VoxelSceneConfig, FGeneratorDesc Generator_, ivec3 WindowOrigin_, ivec3 WindowDims_,
                  bool bStreaming_ = false) {
    Ctx = Ctx_; Generator = Generator_; WindowOrigin = WindowOrigin_; WindowDims = WindowDims_; bStreaming = bStreaming_;
    const FGPUUniformSceneConfig cfg{VoxelSceneConfig.BlockSize, (uint32_t)VoxelSceneConfig.BlockResolution, VoxelSceneConfig.GetChunkSize(), (uint32_t)VoxelSceneConfig.ChunkResolution};
    const int32_t o[3] = {WindowOrigin.x, WindowOrigin.y, WindowOrigin.z}, d[3] = {WindowDims.x, WindowDims.y, WindowDims.z};
    Check(meso_scene_create(Ctx, &cfg, o, d, VoxelSceneConfig.MaxVolumeCount), "meso_scene_create");
    ChunkPool.MaxBlockCount = VoxelSceneConfig.MaxBlockCount;
    ChunkPool.ChunkCount = (int64_t)d[0] * d[1] * d[2];
    if (bStreaming) Check(meso_stream_begin(Ctx, Generator.Kind, Generator.Params, Generator.Granularity), "meso_stream_begin");
    bDirty = true;
  }
  // View direction or camera chunk changed (ChunkManager.h:134-159).  With BakeVisibilityViewNum > 0 the reference answers
  // from the table baked for the nearest Fibonacci direction (ChunkManager.h:106-124): the same direction is used here.
  void UpdateChunks(ivec3 NewChunkLocation, vec3 NewForwardVector, const FVoxelSceneConfig& VoxelSceneConfig) {
    if (bDebugDisableUpdateChunk) return;
    CameraChunk = NewChunkLocation;
    ViewDirection = NewForwardVector;
    if (VoxelSceneConfig.BakeVisibilityViewNum > 1) {
      const float f[3] = {NewForwardVector.x, NewForwardVector.y, NewForwardVector.z};
      float d[3];
      Check(meso_baked_direction(VoxelSceneConfig.BakeVisibilityViewNum, f, d, nullptr), "meso_baked_direction");
      ViewDirection = {d[0], d[1], d[2]};
    }
    bHasView = true;
    FrameStamp++;
  }
  // ChunkManager.h:211-400: generate what is missing, then publish chunk table + block instances (K1 + K2 on the device)
  void UpdateLoadingQueue(const FVoxelSceneConfig& VoxelSceneConfig, uint32_t /*RenderFrameIndex*/) {
    if (bStreaming) {
      if (!bHasView) return;
      const int32_t cc[3] = {CameraChunk.x, CameraChunk.y, CameraChunk.z};
      const float f[3] = {ViewDirection.x, ViewDirection.y, ViewDirection.z};
      const MesoViewConfig vc = VoxelSceneConfig.GetViewConfig();
      MesoStreamStats st{};
      Check(meso_stream_update(Ctx, cc, f, &vc, VoxelSceneConfig.MaxSyncedLoadChunkCount + VoxelSceneConfig.MaxUnsyncedLoadChunkCount, &st), "meso_stream_update");
      DebugVisibleChunkNum = st.candidates; DebugMissingChunkNum = st.missing; DebugLoadedChunkNum += st.generated;
      if (st.generated > 0 || bDirty) Check(meso_build_occupancy(Ctx, FrameStamp, &ChunkPool.CurrentBlockCount), "meso_build_occupancy");
      bDirty = false;
      return;
    }
    if (!bDirty) return;
    Check(meso_voxelize_sdf(Ctx, Generator.Kind, Generator.Params, Generator.Granularity), "meso_voxelize_sdf");
    Check(meso_build_occupancy(Ctx, FrameStamp, &ChunkPool.CurrentBlockCount), "meso_build_occupancy");
    bDirty = false;
  }
 private:
  MesoCtx* Ctx = nullptr;
  FGeneratorDesc Generator;
  ivec3 CameraChunk{0, 0, 0};
  vec3 ViewDirection{};
  bool bHasView = false;
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
