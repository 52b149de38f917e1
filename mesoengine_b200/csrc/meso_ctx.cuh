// meso_ctx.cuh -- the context object behind the C ABI, shared by meso_capi.cu (one GPU) and meso_group.cu (N GPUs of one
// process).  Private to libmeso_b200.so.
#pragma once
#include <string>
#include "meso_internal.cuh"

int meso_fail(int code, const std::string& msg);   // records the message for meso_last_error() (thread-local) and returns code
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      (void)cudaGetLastError(); /* reported here: do not let it surface again in a later CK_LAST */      \
      return meso_fail(MESO_ERR_RUNTIME, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    }                                                                                                    \
  } while (0)
#define CK_LAST(what)                                                                                    \
  do {                                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ != cudaSuccess) return meso_fail(MESO_ERR_RUNTIME, std::string(what) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct MesoCtx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int rank = 0, world = 1;
  int64_t launches = 0;
  bool has_scene = false;
  MesoGPUUniformSceneConfig cfg{};
  DVolume v{};
  // K2
  MesoGPUChunk* d_table = nullptr;
  uint32_t* d_counts = nullptr;
  uint32_t* d_offsets = nullptr;
  uint64_t* d_total = nullptr;
  uint32_t* d_block_totals = nullptr;   // scan scratch: one word per 1024 chunks
  MesoGPUBlock* d_inst = nullptr;
  int64_t cap_inst = 0;
  int64_t n_inst = 0;
  // K4
  MesoHitRecord* d_frame = nullptr;
  size_t frame_px = 0;
  RayStatsDev* d_stats = nullptr;
  uint8_t* d_touch_chunk = nullptr;
  uint8_t* d_touch_brick = nullptr;
  // K3
  uint64_t* d_work = nullptr;
  int64_t cap_work = 0;
  uint32_t* d_work_count = nullptr;
  unsigned long long* d_quad_count = nullptr;
  MesoQuad* d_quads = nullptr;
  int64_t cap_quads = 0;
  // K5
  uint64_t* d_dirty = nullptr;
  uint32_t* d_dirty_count = nullptr;
  uint32_t cap_dirty = 0;
  uint32_t n_dirty = 0;
  uint64_t* d_keys = nullptr;
  uint32_t* d_keys_count = nullptr;
  uint32_t* d_mark = nullptr;
  uint32_t* d_chunk_mark = nullptr;    // one bit per chunk, all-zero between calls (re-mesh: chunks that hold a listed brick)
  uint32_t* d_chunk_list = nullptr;    // nchunks entries
  uint32_t* d_chunk_count = nullptr;
  MeshScratch mesh_scratch() const { return MeshScratch{d_work, cap_work, d_work_count, d_chunk_list, d_chunk_count, d_work_count + 2, d_chunk_mark}; }
  int* d_overflow = nullptr;
  // K6
  uint64_t* d_sel_keys = nullptr;         // candidate keys (scratch), sel_cap entries
  MesoChunkCandidate* d_sel_out = nullptr;  // sorted candidates, sel_cap entries
  int64_t sel_cap = 0;
  uint32_t* d_sel_count = nullptr;
  uint32_t* d_loaded = nullptr;           // bit per chunk slot (scene-sized)
  uint32_t* d_stream_list = nullptr;      // generation list of the current update
  uint32_t stream_list_cap = 0;
  uint32_t* d_stream_stats = nullptr;     // 4 words
  int* d_timeout = nullptr;               // set by a wait that gave up
  void* d_shift_scratch = nullptr;        // moving window: nchunks * 16 KiB, allocated by the first meso_stream_recentre
  bool streaming = false;
  int stream_kind = 0, stream_gran = 0;
  double stream_params[4] = {0, 0, 0, 0};
  // forward cubes (MESO_FLAG_CUBES)
  uint8_t* d_cube_cell = nullptr;
  uint8_t* d_cube_cellp = nullptr;
  uint16_t* d_cube_brick = nullptr;
  uint16_t* d_cube_cell2 = nullptr;
  CubeTables cubes{};
  bool cubes_valid = false;
  // misc
  uint32_t* d_flush = nullptr;
  size_t flush_words = 0;
  uint32_t* d_tmp_count = nullptr;
  unsigned char* h_stage = nullptr;         // pinned + mapped staging for small device -> host reads (meso_small_read)
  unsigned char* d_stage = nullptr;         // its device alias

  cudaStream_t copy_stream = nullptr;       // D2H of finished bands overlaps the next band's kernel (meso_raymarch)
  cudaEvent_t band_done[16] = {nullptr};
  cudaStream_t band_stream[2] = {nullptr, nullptr};
  cudaEvent_t band_fork = nullptr;
  // frame ring (meso_raymarch_async): the reference's kNumBufferedFrames
  MesoHitRecord* d_ring[MESO_FRAME_RING] = {nullptr};
  size_t ring_px = 0;
  bool ring_busy[MESO_FRAME_RING] = {false};
  cudaEvent_t ring_traced[MESO_FRAME_RING] = {nullptr};
  cudaEvent_t ring_copied[MESO_FRAME_RING] = {nullptr};

  LaunchCtx lc() { return LaunchCtx{stream, sm_count, &launches}; }
};


#define NEED_SCENE(c)                                                               \
  do {                                                                              \
    if (!(c)) return meso_fail(MESO_ERR_ARGUMENT, "null context");                  \
    if (!(c)->has_scene) return meso_fail(MESO_ERR_ARGUMENT, "no scene: call meso_scene_create first"); \
    CK(cudaSetDevice((c)->device));                                                 \
  } while (0)

// pieces of single-GPU entry points that meso_group.cu runs on all members before waiting for any of them
extern "C" {
int meso_join_frames(MesoCtx* c);                                       // order the context's stream behind frames in flight
int meso_cubes_for(MesoCtx* c, uint32_t flags, const CubeTables** out); // tables a raymarch launch reads
int meso_voxelize_enqueue(MesoCtx* c, int kind, const double params[4], int granularity);
int meso_overflow_finish(MesoCtx* c, const char* what);                 // waits for the stream; error if the payload pool overflowed
int meso_carve_enqueue(MesoCtx* c, const int32_t center[3], int32_t radius);
int meso_carve_finish(MesoCtx* c, int64_t* n_dirty);
int meso_ensure_mesh_buffers(MesoCtx* c);
#define MESO_STAGE_BYTES (1u << 20)
// device -> host, synchronous on the context's stream; up to MESO_STAGE_BYTES through a kernel + mapped staging (never
// queued behind a DMA in flight), larger reads through cudaMemcpyAsync
int meso_small_read(MesoCtx* c, void* host_dst, const void* d_src, size_t bytes);
}
