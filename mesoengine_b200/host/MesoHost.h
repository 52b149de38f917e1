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
//   FBinaryOccupancyVolume / FOccupancyHelper  Runtimes/Voxel/Occupancy/BinaryOccupancyVolume.h:5-99
//   FChunkBase / FChunk / FEmptyChunk     Runtimes/Voxel/Chunk/Chunk.h:44-106
//   GeneratorType (the CPU plug-in point) Runtimes/Voxel/Chunk/ChunkManager.h:61; FGeneratorHelper::GenerateSphere GeneratorHelper.h:120-150
//   FChunkManage (facade)                 Runtimes/Voxel/Chunk/ChunkManager.h:90-102,134-159,211-400
//   VoxelWindowsInstance (headless)       Runtimes/Instance/VoxelWindowsInstance.{h,cpp}: Initialize, RunInstance, hooks
// Header-only like the reference's Runtimes.  No glm/boost/GLFW/Vulkan: a 40-line vector layer replaces glm here.
// The product path has no CPU compute in this layer: generation, occupancy, meshing and visibility all run in
// libmeso_b200.so.  The host OBJECTS the north star says stay intact (FChunk, FBinaryOccupancyVolume, a user-written
// GeneratorType callback) are kept as plain containers with the reference's methods; what a callback produces travels
// to the device as the reference's own FGPUChunk / FGPUBlock records (meso_volume_upload_blocks).
#pragma once
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
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

// ---- host objects of the reference, kept intact (plain containers; the device path does not need them) --------------
// BinaryOccupancyVolume.h:5-39.  Bit storage = 64-bit words instead of boost::dynamic_bitset; index x + y R + z R^2
// (FVoxelMathHelper::Convert3DTo1D, VoxelMathHelper.h:73-76); Set / GetClamped clamp the location into the volume.
struct FBinaryOccupancyVolume {
  using FOccupancyValue = bool;
  std::vector<uint64_t> OccupancyVolume;
  uint32_t Resolution = 0;
  FBinaryOccupancyVolume() = default;
  explicit FBinaryOccupancyVolume(uint32_t Resolution_) : OccupancyVolume(((size_t)Resolution_ * Resolution_ * Resolution_ + 63) / 64, 0ull), Resolution(Resolution_) {}
  static int32_t Clamp1(int32_t v, int32_t hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
  uint32_t Index(ivec3 L) const { return (uint32_t)L.x + (uint32_t)L.y * Resolution + (uint32_t)L.z * Resolution * Resolution; }
  uint32_t IndexClamped(ivec3 L) const { const int32_t h = (int32_t)Resolution - 1; return Index({Clamp1(L.x, h), Clamp1(L.y, h), Clamp1(L.z, h)}); }
  bool bIsOutOfBound(ivec3 L) const { const int32_t r = (int32_t)Resolution; return L.x < 0 || L.x >= r || L.y < 0 || L.y >= r || L.z < 0 || L.z >= r; }
  void Set(FOccupancyValue Value, ivec3 Location) {
    const uint32_t i = IndexClamped(Location);
    if (Value) OccupancyVolume[i >> 6] |= 1ull << (i & 63); else OccupancyVolume[i >> 6] &= ~(1ull << (i & 63));
  }
  FOccupancyValue GetClamped(ivec3 Location) const { const uint32_t i = IndexClamped(Location); return (OccupancyVolume[i >> 6] >> (i & 63)) & 1ull; }
  FOccupancyValue Get(ivec3 Location) const { const uint32_t i = Index(Location); return (OccupancyVolume[i >> 6] >> (i & 63)) & 1ull; }
  FOccupancyValue GetWithBoundaryCondition(ivec3 Location, FOccupancyValue BoundaryValue = false) const {
    return bIsOutOfBound(Location) ? BoundaryValue : Get(Location);
  }
};

// BinaryOccupancyVolume.h:45-99: a location within one cell of the border erodes to 0; otherwise the AND of its 26
// (bUseComplex, the default) or 6 neighbours.
struct FOccupancyHelper {
  template <bool bUseComplex = true>
  static bool ErodeSingleVoxel(const FBinaryOccupancyVolume& Volume, ivec3 L) {
    const int32_t r = (int32_t)Volume.Resolution;
    if (L.x < 1 || L.x >= r - 1 || L.y < 1 || L.y >= r - 1 || L.z < 1 || L.z >= r - 1) return false;
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          const int n = (dx != 0) + (dy != 0) + (dz != 0);
          if (n == 0 || (!bUseComplex && n != 1)) continue;
          if (!Volume.Get({L.x + dx, L.y + dy, L.z + dz})) return false;
        }
    return true;
  }
};

struct FChunkBase {                       // Chunk.h:44-55
  ivec3 ChunkLocation{INT_MAX, INT_MAX, INT_MAX};
  uint32_t ChunkFrameStamp = 0;
  bool bIsValid() const { return ChunkLocation != ivec3{INT_MAX, INT_MAX, INT_MAX}; }
};
struct FChunk : public FChunkBase {       // Chunk.h:57-101
  std::vector<FBlock> Blocks;
  std::vector<FBinaryOccupancyVolume> OccupancyVolumeErodeMipmaps;
  void AddBlock(const FBlock& NewBlock) { Blocks.push_back(NewBlock); }
  // Mip0 = the block set; Mip_d, evaluated at block locations only, = erode(Mip_{d-1}).  On the device this is K2
  // (meso_build_occupancy); the host method exists for callers that inspect a chunk before handing it over.
  void CalculateOccupancyErodeMipmaps(const uint32_t Resolution = 16, const uint32_t MaxDepth = 4) {
    OccupancyVolumeErodeMipmaps.clear();
    OccupancyVolumeErodeMipmaps.emplace_back(Resolution);
    for (const FBlock& b : Blocks) OccupancyVolumeErodeMipmaps[0].Set(true, {b.BlockLocation[0], b.BlockLocation[1], b.BlockLocation[2]});
    for (uint32_t d = 1; d < MaxDepth; d++) {
      FBinaryOccupancyVolume Mip(Resolution);
      for (const FBlock& b : Blocks) {
        const ivec3 L{b.BlockLocation[0], b.BlockLocation[1], b.BlockLocation[2]};
        Mip.Set(FOccupancyHelper::ErodeSingleVoxel(OccupancyVolumeErodeMipmaps[d - 1], L), L);
      }
      OccupancyVolumeErodeMipmaps.push_back(std::move(Mip));
    }
  }
  bool bShouldVoxelOccupancyCull(ivec3 BlockLocation, uint32_t ThresholdDepth = 2) const {
    return OccupancyVolumeErodeMipmaps[ThresholdDepth].GetWithBoundaryCondition(BlockLocation, true);
  }
};
struct FEmptyChunk : public FChunkBase {};

// The generator plug-in point, as the reference declares it (ChunkManager.h:61): called once per chunk on host worker
// threads; FChunk.Blocks is what travels to the device.
using GeneratorType = std::function<FChunk(ivec3, float, unsigned char, uint32_t)>;

struct FGeneratorHelper {
  // GeneratorHelper.h:120-150: block min corner against the sphere (100,0,0) r 50, fp64, X outer / Z inner
  static FChunk GenerateSphere(ivec3 StartLocation, float BlockSize, unsigned char ChunkResolution, uint32_t /*MipmapLevel*/) {
    FChunk Result;
    Result.ChunkLocation = StartLocation;
    const double s = (double)BlockSize, cr = (double)ChunkResolution;
    const double ox = (double)StartLocation.x * s * cr, oy = (double)StartLocation.y * s * cr, oz = (double)StartLocation.z * s * cr;
    for (uint32_t X = 0; X < ChunkResolution; X++)
      for (uint32_t Y = 0; Y < ChunkResolution; Y++)
        for (uint32_t Z = 0; Z < ChunkResolution; Z++) {
          const double dx = (ox + (double)X * s) - 100.0, dy = (oy + (double)Y * s) - 0.0, dz = (oz + (double)Z * s) - 0.0;
          if (std::sqrt((dx * dx + dy * dy) + dz * dz) - 50.0 < 0.0) {
            FBlock b; b.ChunkIndex = 0; b.BlockLocation[0] = (uint8_t)X; b.BlockLocation[1] = (uint8_t)Y; b.BlockLocation[2] = (uint8_t)Z; b.VolumeIndex = 0;
            Result.AddBlock(b);
          }
        }
    Result.CalculateOccupancyErodeMipmaps();
    return Result;
  }
};

// The device-side form of the plug-in point: the SDF is evaluated by K1, so the "generator" is a description, not a
// callback.  A GeneratorType callback (above) is the other way in: FChunkManage runs it on host threads and uploads.
struct FGeneratorDesc {
  int Kind = MESO_SDF_SPHERE;                       // FGeneratorHelper::GenerateSphere / TestGenerator
  double Params[4] = {100.0, 0.0, 0.0, 50.0};       // GeneratorHelper.h:134
  int Granularity = MESO_GRAN_BLOCK;                // reference: one sample per block
  uint32_t MipmapLevel = 0;                         // the plug-in's fourth argument (ChunkManager.h:61): (2^level)^3 blocks per sample
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

  // FChunkManage::SetGenerator (ChunkManager.h:62-65): a user-written CPU generator.  When set, UpdateLoadingQueue runs it for
  // every chunk of the window on ThreadCount host threads (the reference's generator workers, SimpleVoxel.cpp:260-261) and
  // uploads the blocks as the reference's own FGPUChunk / FGPUBlock records; K2 (mips, cull, instances) stays on the device.
  void SetGenerator(GeneratorType Generator_, uint32_t ThreadCount_ = 0) {
    HostGenerator = std::move(Generator_);
    const uint32_t hc = std::max(1u, std::thread::hardware_concurrency());
    ThreadCount = ThreadCount_ ? ThreadCount_ : std::max(hc > 2 ? hc - 2 : 1u, 1u);
    bDirty = true;
  }
  // N GPUs: the volume is replicated on every member of the group; Ctx_ is then member 0, which also runs K2.
  void SetGroup(MesoGroup* Group_) { Group = Group_; }
  void Initialize(MesoCtx* Ctx_, const FVoxelSceneConfig& VoxelSceneConfig, FGeneratorDesc Generator_, ivec3 WindowOrigin_, ivec3 WindowDims_,
                  bool bStreaming_ = false) {
    Ctx = Ctx_; Generator = Generator_; WindowOrigin = WindowOrigin_; WindowDims = WindowDims_; bStreaming = bStreaming_;
    if (Group && bStreaming) throw std::runtime_error("FChunkManage: streaming generation runs on one context (no group)");
    const FGPUUniformSceneConfig cfg{VoxelSceneConfig.BlockSize, (uint32_t)VoxelSceneConfig.BlockResolution, VoxelSceneConfig.GetChunkSize(), (uint32_t)VoxelSceneConfig.ChunkResolution};
    const int32_t o[3] = {WindowOrigin.x, WindowOrigin.y, WindowOrigin.z}, d[3] = {WindowDims.x, WindowDims.y, WindowDims.z};
    if (Group) Check(meso_group_scene_create(Group, &cfg, o, d, VoxelSceneConfig.MaxVolumeCount), "meso_group_scene_create");
    else Check(meso_scene_create(Ctx, &cfg, o, d, VoxelSceneConfig.MaxVolumeCount), "meso_scene_create");
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
      if (HostGenerator) throw std::runtime_error("FChunkManage: a host GeneratorType callback needs the static window mode (bStreaming = false)");
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
    if (HostGenerator) {
      RunHostGenerator(VoxelSceneConfig);
      Check(meso_build_occupancy(Ctx, FrameStamp, &ChunkPool.CurrentBlockCount), "meso_build_occupancy");
      bDirty = false;
      return;
    }
    if (Group) Check(meso_group_voxelize_sdf(Group, Generator.Kind, Generator.Params, Generator.Granularity), "meso_group_voxelize_sdf");
    else if (Generator.MipmapLevel > 0 && Generator.Granularity == MESO_GRAN_BLOCK)
      Check(meso_voxelize_sdf_lod(Ctx, Generator.Kind, Generator.Params, Generator.MipmapLevel), "meso_voxelize_sdf_lod");
    else Check(meso_voxelize_sdf(Ctx, Generator.Kind, Generator.Params, Generator.Granularity), "meso_voxelize_sdf");
    Check(meso_build_occupancy(Ctx, FrameStamp, &ChunkPool.CurrentBlockCount), "meso_build_occupancy");
    bDirty = false;
  }
  int64_t DebugUploadedBlockNum = 0;      // blocks the last host-generator pass handed to the device
 private:
  // MultiThreadGeneratorBatched (ChunkManager.h:183-210) without the pool placement: chunks of the window are dealt to the
  // workers round-robin; every worker fills its own record vectors (ChunkIndex = the chunk's row in the table).
  void RunHostGenerator(const FVoxelSceneConfig& VoxelSceneConfig) {
    const int64_t n = (int64_t)WindowDims.x * WindowDims.y * WindowDims.z;
    std::vector<FGPUChunk> Table((size_t)n);
    const uint32_t nt = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ThreadCount, n));
    std::vector<std::vector<FGPUBlock>> Parts(nt);
    std::vector<std::string> Errors(nt);
    auto Work = [&](uint32_t t) {
      try {
        for (int64_t i = t; i < n; i += nt) {
          const ivec3 loc{WindowOrigin.x + (int32_t)(i % WindowDims.x), WindowOrigin.y + (int32_t)((i / WindowDims.x) % WindowDims.y),
                          WindowOrigin.z + (int32_t)(i / ((int64_t)WindowDims.x * WindowDims.y))};
          FChunk c = HostGenerator(loc, VoxelSceneConfig.BlockSize, VoxelSceneConfig.ChunkResolution, Generator.MipmapLevel);
          c.ChunkLocation = loc;   // "just make sure" (ChunkManager.h:167)
          FGPUChunk rec;
          if (c.Blocks.empty()) { rec.ChunkLocation[0] = rec.ChunkLocation[1] = rec.ChunkLocation[2] = INT_MAX; rec.ChunkFrameStamp = 0; }
          else { rec.ChunkLocation[0] = loc.x; rec.ChunkLocation[1] = loc.y; rec.ChunkLocation[2] = loc.z; rec.ChunkFrameStamp = FrameStamp; }
          Table[(size_t)i] = rec;
          for (const FBlock& b : c.Blocks)
            Parts[t].push_back(FGPUBlock{(uint32_t)i, {b.BlockLocation[0], b.BlockLocation[1], b.BlockLocation[2], 255u}, FrameStamp});
        }
      } catch (const std::exception& e) { Errors[t] = e.what(); }
    };
    std::vector<std::thread> Workers;
    for (uint32_t t = 1; t < nt; t++) Workers.emplace_back(Work, t);
    Work(0);
    for (auto& w : Workers) w.join();
    for (const auto& e : Errors) if (!e.empty()) throw std::runtime_error("GeneratorType callback: " + e);
    std::vector<FGPUBlock> Blocks;
    size_t total = 0;
    for (const auto& p : Parts) total += p.size();
    Blocks.reserve(total);
    for (const auto& p : Parts) Blocks.insert(Blocks.end(), p.begin(), p.end());
    if (Group) Check(meso_group_volume_upload_blocks(Group, Table.data(), n, Blocks.data(), (int64_t)Blocks.size(), 0u, &DebugUploadedBlockNum), "meso_group_volume_upload_blocks");
    else Check(meso_volume_upload_blocks(Ctx, Table.data(), n, Blocks.data(), (int64_t)Blocks.size(), 0u, &DebugUploadedBlockNum), "meso_volume_upload_blocks");
  }
  MesoCtx* Ctx = nullptr;
  MesoGroup* Group = nullptr;
  FGeneratorDesc Generator;
  GeneratorType HostGenerator;
  uint32_t ThreadCount = 1;
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
  std::vector<int> Devices;                       // more than one entry: the frame is split over these GPUs (meso_group_*)
  FVoxelSceneConfig VoxelSceneConfig;
};

// Headless frame loop with the reference's virtual hooks and call order (VoxelWindowsInstance.cpp:368-393):
// UpdateCamera -> UpdatePhysics -> RenderStart -> Render -> RenderEnd -> UpdateFrameIndex.
class VoxelWindowsInstance {
 public:
  MesoCtx* Context = nullptr;                     // stands where std::unique_ptr<lvk::IContext> LVKContext stood
  MesoGroup* Group = nullptr;                     // InitialConfig.Devices.size() > 1: N GPUs behind one handle; Context = member 0
  int WindowsWidth = 0, WindowsHeight = 0;
  bool bLVKReverseZ = true;
  uint32_t LVKNumBufferedFrames = 3;
  FVoxelSceneConfig VoxelSceneConfig;
  FVoxelCamera WindowsCamera;
  FGPUUniformSceneConfig SceneConfig{};
  std::vector<FGPUUniformCamera> UBOCamera;       // per-frame camera ring (VoxelWindowsInstance.cpp:113-120)
  std::vector<MesoHitRecord> OffscreenRecords;    // stands where TEXOffscreenColor stood (16 B records instead of RGBA8)
  uint32_t RenderGlobalFrameIndex = 0, RenderFrameIndex = 0;

  virtual ~VoxelWindowsInstance() {
    if (Group) meso_group_destroy(Group);         // owns its members, Context included
    else if (Context) meso_ctx_destroy(Context);
  }
  virtual void Initialize(const VoxelInstanceInitialConfig& InitialConfig) {
    bLVKReverseZ = InitialConfig.bReverseZ; LVKNumBufferedFrames = InitialConfig.kNumBufferedFrames;
    WindowsWidth = InitialConfig.WindowsWidth; WindowsHeight = InitialConfig.WindowsHeight;
    VoxelSceneConfig = InitialConfig.VoxelSceneConfig;
    InitializeCameraAndScene(InitialConfig);
    Device = InitialConfig.Device;
    Devices = InitialConfig.Devices;
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
    if (Devices.size() > 1) {
      Check(meso_group_create(Devices.data(), (int)Devices.size(), &Group), "meso_group_create");
      Context = meso_group_ctx(Group, 0);
    } else {
      Check(meso_ctx_create(Devices.empty() ? Device : Devices[0], &Context), "meso_ctx_create");
    }
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
  std::vector<int> Devices;
};

}  // namespace meso
