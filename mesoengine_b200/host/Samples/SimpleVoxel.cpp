// SimpleVoxel (headless) -- the reference sample (Samples/SimpleVoxel.cpp:229-460) on the B200 path.
// Same structure: a VoxelWindowsInstance subclass that owns an FChunkManage and overrides the frame-loop hooks; the
// instanced draw of Render() (cmdBindVertexBuffer / cmdPushConstants / cmdDrawIndexed(8, MaxBlockCount),
// SimpleVoxel.cpp:352-398) is one call: meso_raymarch().
//
//   SimpleVoxel [--stream | --cpu-generator | --gpus N] [frames] [width height] [eye x y z] [target x y z] [out.bin]
// --gpus N: the frame is split over GPUs 0..N-1 behind one MesoGroup (replicated volume, slab gather into host memory);
//           N may exceed the number of devices in the box (members then share GPUs: device r % count).
// --cpu-generator: the generator is the reference's std::function callback (SimpleVoxel.cpp:263-267) run on host worker
// threads; its FChunk.Blocks reach the device as FGPUChunk / FGPUBlock records (meso_volume_upload_blocks).
// --stream: the reference's loading loop -- chunks are generated as the view asks for them (FChunkManage::UpdateChunks +
// UpdateLoadingQueue, at most MaxUnsyncedLoadChunkCount per frame), not all at once.
// Prints an FNV-1a checksum of the last frame's records (tests/test_gpu_host_sample.py compares it with the Python path).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../MesoHost.h"

using namespace meso;

class SimpleVoxelWindowsInstance : public VoxelWindowsInstance {
 public:
  FChunkManage ChunkManager;
  bool bStream = false;
  bool bCpuGenerator = false;
  uint32_t CameraUpdates = 0, ChunkUpdates = 0;
  double RenderMs = 0.0;

  void InitializeBegin() override {
    VoxelWindowsInstance::InitializeBegin();
    FGeneratorDesc Generator;   // FGeneratorHelper::GenerateSphere, the generator SimpleVoxel.cpp:263-267 wires in
    // the 8^3 chunks around the reference sphere (centre (100,0,0), radius 50 blocks)
    ChunkManager.SetGroup(Group);
    ChunkManager.Initialize(Context, VoxelSceneConfig, Generator, ivec3{2, -4, -4}, ivec3{8, 8, 8}, bStream);
    if (bCpuGenerator) {
      auto GeneratorInstance = [](ivec3 StartLocation, float BlockSize, unsigned char ChunkResolution, uint32_t MipmapLevel) {
        return FGeneratorHelper::GenerateSphere(StartLocation, BlockSize, ChunkResolution, MipmapLevel);
      };
      ChunkManager.SetGenerator(GeneratorInstance);
    }
  }
  void WhenCameraChunkUpdate() override { ChunkUpdates++; }
  void WhenCameraUpdate() override {
    CameraUpdates++;
    ChunkManager.UpdateChunks(WindowsCamera.CameraChunkLocation, WindowsCamera.CameraForward, VoxelSceneConfig);
  }
  void UpdatePhysics() override { ChunkManager.UpdateLoadingQueue(VoxelSceneConfig, RenderFrameIndex); }
  void Render() override {
    const auto t0 = std::chrono::steady_clock::now();
    const float Light[3] = {0.3f, 0.5f, 0.8f};
    if (Group)
      Check(meso_group_raymarch(Group, &UBOCamera[RenderFrameIndex], WindowsWidth, WindowsHeight, MESO_FLAG_SHADOW, Light, OffscreenRecords.data()),
            "meso_group_raymarch");
    else
      Check(meso_raymarch(Context, &UBOCamera[RenderFrameIndex], WindowsWidth, WindowsHeight, MESO_FLAG_SHADOW, Light, OffscreenRecords.data()),
            "meso_raymarch");
    RenderMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

int main(int argc, char* argv[]) {
  bool stream = false, cpu_generator = false;
  if (argc > 1 && std::strcmp(argv[1], "--stream") == 0) { stream = true; argc--; argv++; }
  else if (argc > 1 && std::strcmp(argv[1], "--cpu-generator") == 0) { cpu_generator = true; argc--; argv++; }
  int gpus = 1;
  if (argc > 2 && std::strcmp(argv[1], "--gpus") == 0) { gpus = std::max(1, std::atoi(argv[2])); argc -= 2; argv += 2; }
  const uint32_t frames = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 4;
  VoxelInstanceInitialConfig cfg;
  if (gpus > 1) {
    int count = 1;
    if (const char* e = std::getenv("MESO_SAMPLE_DEVICE_COUNT")) count = std::max(1, std::atoi(e));   // set by the caller (no CUDA headers here)
    for (int r = 0; r < gpus; r++) cfg.Devices.push_back(r % count);
  }
  if (argc > 3) { cfg.WindowsWidth = std::atoi(argv[2]); cfg.WindowsHeight = std::atoi(argv[3]); }
  vec3 eye{5.0f, 2.0f, 2.0f}, target{100.0f, 0.0f, 0.0f};   // the reference start position, turned towards the sphere
  if (argc > 9) {
    eye = {(float)std::atof(argv[4]), (float)std::atof(argv[5]), (float)std::atof(argv[6])};
    target = {(float)std::atof(argv[7]), (float)std::atof(argv[8]), (float)std::atof(argv[9])};
  }
  try {
    SimpleVoxelWindowsInstance Instance;
    Instance.bStream = stream;
    Instance.bCpuGenerator = cpu_generator;
    Instance.WindowsCamera.SetPose(eye, target, {0.0f, 0.0f, 1.0f});
    Instance.Initialize(cfg);
    Instance.RunInstance(frames);
    uint64_t h = 1469598103934665603ull;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(Instance.OffscreenRecords.data());
    const size_t n = Instance.OffscreenRecords.size() * sizeof(MesoHitRecord);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    size_t hits = 0;
    for (const auto& r : Instance.OffscreenRecords) hits += (r.w1 >> 20) & 1u;
    std::printf("frames=%u size=%dx%d blocks=%lld hits=%zu checksum=%016llx ms_per_frame=%.3f camera_updates=%u loaded=%u missing=%u\n", frames, cfg.WindowsWidth,
                cfg.WindowsHeight, (long long)Instance.ChunkManager.ChunkPool.CurrentBlockCount, hits, (unsigned long long)h,
                Instance.RenderMs / frames, Instance.CameraUpdates, Instance.ChunkManager.DebugLoadedChunkNum, Instance.ChunkManager.DebugMissingChunkNum);
    if (argc > 10) {
      FILE* f = std::fopen(argv[10], "wb");
      if (f) { std::fwrite(p, 1, n, f); std::fclose(f); }
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "SimpleVoxel: %s\n", e.what());
    return 1;
  }
  return 0;
}
