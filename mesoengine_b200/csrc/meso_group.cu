// meso_group.cu -- N GPUs of one box behind ONE handle, for a single-process host (the C++ engine).
//
// The reference owns one device through one lvk::IContext (Runtimes/Instance/VoxelWindowsInstance.cpp:104-111); a host
// that wants the frame split over the 8 GPUs of a box should not have to re-implement rendezvous and peer-memory
// exchange itself (one context over N devices).  A MesoGroup is N member contexts -- member r
// renders screen tiles t % N == r and meshes chunks c % N == r of a REPLICATED volume -- with peer access
// enabled between them, and three fused exchanges:
//   * frames:  every member's raymarch kernel stores each record straight into the horizontal slab of the frame it
//              belongs to (MESO_LAYOUT_SLABS; the slab lives on the member that owns those rows, peer memory over
//              NVLink), so when the kernels are done every member holds a contiguous part of the frame and copies it to the
//              host with ONE large DMA over its own PCIe link.  Device-side ordering only: cross-device event waits, no
//              host thread in the loop until meso_group_frame_wait.
//   * quads:   members mesh into their own lists; the lists are delivered at prefix offsets, each over its own link.
//   * edits:   carves are replicated compute (identical on every replica, no traffic); the dirty re-mesh is sharded by key.
// One caller thread per group (the contract of include/meso_cuda.h).
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "meso_ctx.cuh"

struct MesoGroup {
  int n = 0;
  MesoCtx* ctx[MESO_MAX_SLABS] = {nullptr};
  // frame ring: slab[slot][r] = member r's part of frame `slot` (rows [r * rows, (r + 1) * rows) of the frame)
  void* slab[MESO_FRAME_RING][MESO_MAX_SLABS] = {{nullptr}};
  size_t slab_bytes = 0;
  int ring_w = 0, ring_h = 0, rows_per_slab = 0;
  cudaEvent_t traced[MESO_FRAME_RING][MESO_MAX_SLABS] = {{nullptr}};
  cudaEvent_t copied[MESO_FRAME_RING][MESO_MAX_SLABS] = {{nullptr}};
  cudaStream_t copy_stream[MESO_MAX_SLABS] = {nullptr};
  bool busy[MESO_FRAME_RING] = {false};
  // mesh: pinned per-member counts
  unsigned long long* h_counts = nullptr;
};

static int fail(int code, const std::string& msg) { return meso_fail(code, msg); }
#define NEED_GROUP(g) do { if (!(g) || (g)->n < 1) return fail(MESO_ERR_ARGUMENT, "null group"); } while (0)

extern "C" {

int meso_group_destroy(MesoGroup* g) {
  if (!g) return MESO_OK;
  for (int r = 0; r < g->n; r++) {
    if (!g->ctx[r]) continue;
    cudaSetDevice(g->ctx[r]->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < MESO_FRAME_RING; s++) {
      cudaFree(g->slab[s][r]);
      if (g->traced[s][r]) cudaEventDestroy(g->traced[s][r]);
      if (g->copied[s][r]) cudaEventDestroy(g->copied[s][r]);
    }
    if (g->copy_stream[r]) cudaStreamDestroy(g->copy_stream[r]);
  }
  if (g->h_counts) cudaFreeHost(g->h_counts);
  for (int r = 0; r < g->n; r++) meso_ctx_destroy(g->ctx[r]);
  delete g;
  return MESO_OK;
}

int meso_group_create(const int* devices, int n, MesoGroup** out) {
  if (!out) return fail(MESO_ERR_ARGUMENT, "meso_group_create: out is null");
  *out = nullptr;
  if (!devices || n < 1 || n > MESO_MAX_SLABS) return fail(MESO_ERR_ARGUMENT, "meso_group_create: need 1..MESO_MAX_SLABS devices");
  MesoGroup* g = new MesoGroup();
  g->n = n;
  auto bail = [&](int code) { meso_group_destroy(g); return code; };
  for (int r = 0; r < n; r++) {
    const int rc = meso_ctx_create(devices[r], &g->ctx[r]);
    if (rc != MESO_OK) return bail(rc);
    meso_ctx_set_partition(g->ctx[r], r, n);
  }
  // peer access between every pair of distinct devices (the same device may appear twice: two members sharing a GPU)
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      const int da = g->ctx[a]->device, db = g->ctx[b]->device;
      if (da == db) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, da, db) != cudaSuccess || !can) {
        (void)cudaGetLastError();
        fail(MESO_ERR_RUNTIME, "meso_group_create: device " + std::to_string(da) + " cannot access device " + std::to_string(db) + " (no peer path)");
        return bail(MESO_ERR_RUNTIME);
      }
      cudaSetDevice(da);
      const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        fail(MESO_ERR_RUNTIME, std::string("meso_group_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        (void)cudaGetLastError();
        return bail(MESO_ERR_RUNTIME);
      }
      (void)cudaGetLastError();
    }
  for (int r = 0; r < n; r++) {
    if (cudaSetDevice(g->ctx[r]->device) != cudaSuccess || cudaStreamCreateWithFlags(&g->copy_stream[r], cudaStreamNonBlocking) != cudaSuccess) {
      fail(MESO_ERR_RUNTIME, "meso_group_create: stream creation failed");
      return bail(MESO_ERR_RUNTIME);
    }
    for (int s = 0; s < MESO_FRAME_RING; s++)
      if (cudaEventCreateWithFlags(&g->traced[s][r], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&g->copied[s][r], cudaEventDisableTiming) != cudaSuccess) {
        fail(MESO_ERR_RUNTIME, "meso_group_create: event creation failed");
        return bail(MESO_ERR_RUNTIME);
      }
  }
  if (cudaHostAlloc(&g->h_counts, sizeof(unsigned long long) * MESO_MAX_SLABS, cudaHostAllocPortable) != cudaSuccess) {
    fail(MESO_ERR_RUNTIME, "meso_group_create: pinned allocation failed");
    return bail(MESO_ERR_RUNTIME);
  }
  *out = g;
  return MESO_OK;
}

int meso_group_size(MesoGroup* g) { return g ? g->n : 0; }
MesoCtx* meso_group_ctx(MesoGroup* g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }

int meso_group_sync(MesoGroup* g) {
  NEED_GROUP(g);
  for (int r = 0; r < g->n; r++) {
    const int rc = meso_ctx_sync(g->ctx[r]);
    if (rc != MESO_OK) return rc;
    CK(cudaStreamSynchronize(g->copy_stream[r]));
  }
  return MESO_OK;
}

// ---- replicated volume operations: enqueue on every member first, then wait for each ---------------------------------
int meso_group_scene_create(MesoGroup* g, const MesoGPUUniformSceneConfig* cfg, const int32_t origin[3], const int32_t dims[3], uint32_t max_bricks) {
  NEED_GROUP(g);
  for (int r = 0; r < g->n; r++) {
    const int rc = meso_scene_create(g->ctx[r], cfg, origin, dims, max_bricks);
    if (rc != MESO_OK) return rc;
  }
  return MESO_OK;
}

int meso_group_voxelize_sdf(MesoGroup* g, int kind, const double params[4], int granularity) {
  NEED_GROUP(g);
  for (int r = 0; r < g->n; r++) {
    const int rc = meso_voxelize_enqueue(g->ctx[r], kind, params, granularity);
    if (rc != MESO_OK) return rc;
  }
  int first = MESO_OK;
  for (int r = 0; r < g->n; r++) {
    const int rc = meso_overflow_finish(g->ctx[r], "meso_group_voxelize_sdf");
    if (rc != MESO_OK) { g->ctx[r]->cubes_valid = false; if (first == MESO_OK) first = rc; }
  }
  return first;
}

int meso_group_volume_upload_blocks(MesoGroup* g, const MesoGPUChunk* chunks, int64_t n_chunks, const MesoGPUBlock* blocks, int64_t n_blocks,
                                    uint32_t flags, int64_t* n_accepted) {
  NEED_GROUP(g);
  for (int r = 0; r < g->n; r++) {
    const int rc = meso_volume_upload_blocks(g->ctx[r], chunks, n_chunks, blocks, n_blocks, flags, r == 0 ? n_accepted : nullptr);
    if (rc != MESO_OK) return rc;
  }
  return MESO_OK;
}

int meso_group_carve_sphere(MesoGroup* g, const int32_t center[3], int32_t radius, int64_t* n_dirty) {
  NEED_GROUP(g);
  for (int r = 0; r < g->n; r++) {
    const int rc = meso_carve_enqueue(g->ctx[r], center, radius);
    if (rc != MESO_OK) return rc;
  }
  int first = MESO_OK;
  for (int r = 0; r < g->n; r++) {
    int64_t nd = 0;
    const int rc = meso_carve_finish(g->ctx[r], &nd);
    if (rc != MESO_OK && first == MESO_OK) first = rc;
    if (r == 0 && n_dirty) *n_dirty = nd;
  }
  return first;
}

// ---- frames -------------------------------------------------------------------------------------------------------
static int ensure_slabs(MesoGroup* g, int width, int height) {
  const int tiles_y = (height + MESO_TILE_H - 1) / MESO_TILE_H;
  const int rows = ((tiles_y + g->n - 1) / g->n) * MESO_TILE_H;
  const size_t bytes = (size_t)rows * width * sizeof(MesoHitRecord);
  if (g->ring_w == width && g->ring_h == height && g->rows_per_slab == rows && g->slab_bytes >= bytes) return MESO_OK;
  const int rc = meso_group_sync(g);
  if (rc != MESO_OK) return rc;
  for (int s = 0; s < MESO_FRAME_RING; s++) g->busy[s] = false;
  for (int r = 0; r < g->n; r++) {
    CK(cudaSetDevice(g->ctx[r]->device));
    for (int s = 0; s < MESO_FRAME_RING; s++) {
      cudaFree(g->slab[s][r]); g->slab[s][r] = nullptr;
      CK(cudaMalloc(&g->slab[s][r], bytes));
    }
  }
  g->slab_bytes = bytes; g->ring_w = width; g->ring_h = height; g->rows_per_slab = rows;
  return MESO_OK;
}

int meso_group_frame_wait(MesoGroup* g, int slot) {
  NEED_GROUP(g);
  if (slot < 0 || slot >= MESO_FRAME_RING) return fail(MESO_ERR_ARGUMENT, "meso_group_frame_wait: slot out of range");
  if (!g->busy[slot]) return MESO_OK;
  for (int r = 0; r < g->n; r++) {
    CK(cudaSetDevice(g->ctx[r]->device));
    CK(cudaEventSynchronize(g->copied[slot][r]));
    g->ctx[r]->ring_busy[slot] = false;
  }
  g->busy[slot] = false;
  return MESO_OK;
}

int meso_group_raymarch_async(MesoGroup* g, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3],
                              MesoHitRecord* host_records, int slot) {
  NEED_GROUP(g);
  if (!cam || !host_records || width <= 0 || height <= 0 || width > 65536 || height > 65536) return fail(MESO_ERR_ARGUMENT, "meso_group_raymarch: bad argument");
  if (slot < 0 || slot >= MESO_FRAME_RING) return fail(MESO_ERR_ARGUMENT, "meso_group_raymarch: slot out of range");
  for (int r = 0; r < g->n; r++) if (!g->ctx[r]->has_scene) return fail(MESO_ERR_ARGUMENT, "no scene: call meso_group_scene_create first");
  int rc = meso_group_frame_wait(g, slot);        // the slot's previous frame has left its slabs
  if (rc != MESO_OK) return rc;
  rc = ensure_slabs(g, width, height);
  if (rc != MESO_OK) return rc;
  const size_t bpp = (flags & MESO_FLAG_RGBA8) ? 4 : sizeof(MesoHitRecord);
  FrameMap fm{};
  for (int r = 0; r < g->n; r++) fm.slab[r] = g->slab[slot][r];
  fm.rows_per_slab = g->rows_per_slab; fm.n_slabs = g->n;
  // 1. every member traces its tiles, storing each record into the slab (member) that owns its row
  for (int r = 0; r < g->n; r++) {
    MesoCtx* c = g->ctx[r];
    CK(cudaSetDevice(c->device));
    MesoRaySetup rs;
    rc = meso_ray_setup(cam, c->v.origin, width, height, light, &rs);
    if (rc != MESO_OK) return rc;
    const CubeTables* cubes = nullptr;
    rc = meso_cubes_for(c, flags, &cubes);
    if (rc != MESO_OK) return rc;
    LaunchCtx lc = c->lc();
    lc.stream = c->band_stream[slot & 1];         // frames alternate between two streams: the tail of one overlaps the next
    CK(cudaEventRecord(c->band_fork, c->stream));
    CK(cudaStreamWaitEvent(lc.stream, c->band_fork, 0));
    launch_raymarch(lc, c->v, rs, width, height, flags, r, g->n, MESO_LAYOUT_SLABS, nullptr, nullptr, nullptr, nullptr, 0, -1, cubes, &fm);
    CK_LAST("group raymarch");
    CK(cudaEventRecord(g->traced[slot][r], lc.stream));
    CK(cudaEventRecord(c->ring_traced[slot], lc.stream));   // volume edits on this member are ordered behind this frame
    c->ring_busy[slot] = true;
  }
  // 2. every member waits (on the device) for all members' kernels -- its slab receives rows from each of them -- and
  //    sends its slab to the host with one DMA over its own PCIe link
  for (int r = 0; r < g->n; r++) {
    MesoCtx* c = g->ctx[r];
    CK(cudaSetDevice(c->device));
    for (int k = 0; k < g->n; k++) CK(cudaStreamWaitEvent(g->copy_stream[r], g->traced[slot][k], 0));
    const int row0 = r * g->rows_per_slab;
    const int rows = std::min(g->rows_per_slab, height - row0);
    if (rows > 0)
      CK(cudaMemcpyAsync(reinterpret_cast<char*>(host_records) + (size_t)row0 * width * bpp, g->slab[slot][r], (size_t)rows * width * bpp,
                         cudaMemcpyDeviceToHost, g->copy_stream[r]));
    CK(cudaEventRecord(g->copied[slot][r], g->copy_stream[r]));
  }
  g->busy[slot] = true;
  return MESO_OK;
}

int meso_group_raymarch(MesoGroup* g, const MesoGPUUniformCamera* cam, int width, int height, uint32_t flags, const float light[3],
                        MesoHitRecord* host_records) {
  const int rc = meso_group_raymarch_async(g, cam, width, height, flags, light, host_records, 0);
  return rc != MESO_OK ? rc : meso_group_frame_wait(g, 0);
}

// ---- quads ----------------------------------------------------------------------------------------------------------
// Every member meshes its chunks (c % N == r) into its own device list; the 8-byte counts come back through pinned
// memory; every list is then copied to the host at its prefix offset, all members at once (N PCIe links).
// counts (optional, N entries) receives the per-member quad counts: host_quads is the concatenation in member order.
int meso_group_mesh(MesoGroup* g, MesoQuad* host_quads, int64_t cap, int64_t* n_quads, int64_t* counts) {
  NEED_GROUP(g);
  if (!n_quads || cap < 0 || (cap > 0 && !host_quads)) return fail(MESO_ERR_ARGUMENT, "meso_group_mesh: bad argument");
  for (int r = 0; r < g->n; r++) {
    MesoCtx* c = g->ctx[r];
    NEED_SCENE(c);
    if (cap > c->cap_quads) {
      CK(cudaStreamSynchronize(c->stream));
      cudaFree(c->d_quads); c->d_quads = nullptr; c->cap_quads = 0;
      CK(cudaMalloc(&c->d_quads, (size_t)std::max<int64_t>(cap, 1) * sizeof(MesoQuad)));
      c->cap_quads = cap;
    }
    const int rc = meso_ensure_mesh_buffers(c);
    if (rc != MESO_OK) return rc;
    launch_mesh(c->lc(), c->v, r, g->n, c->mesh_scratch(), c->d_quads, cap, c->d_quad_count);
    CK_LAST("group mesh");
    CK(cudaMemcpyAsync(&g->h_counts[r], c->d_quad_count, 8, cudaMemcpyDeviceToHost, c->stream));
  }
  int64_t total = 0;
  std::vector<int64_t> off((size_t)g->n);
  for (int r = 0; r < g->n; r++) {
    CK(cudaSetDevice(g->ctx[r]->device));
    CK(cudaStreamSynchronize(g->ctx[r]->stream));
    off[(size_t)r] = total;
    total += (int64_t)g->h_counts[r];
    if (counts) counts[r] = (int64_t)g->h_counts[r];
  }
  *n_quads = total;
  if (total > cap) return fail(MESO_ERR_ARGUMENT, "meso_group_mesh: cap too small for the gathered list (n_quads holds the size needed)");
  for (int r = 0; r < g->n; r++) {
    const int64_t nr = (int64_t)g->h_counts[r];
    if (nr == 0) continue;
    CK(cudaSetDevice(g->ctx[r]->device));
    CK(cudaMemcpyAsync(host_quads + off[(size_t)r], g->ctx[r]->d_quads, (size_t)nr * sizeof(MesoQuad), cudaMemcpyDeviceToHost, g->ctx[r]->stream));
  }
  for (int r = 0; r < g->n; r++) {
    CK(cudaSetDevice(g->ctx[r]->device));
    CK(cudaStreamSynchronize(g->ctx[r]->stream));
  }
  return MESO_OK;
}

// Device-resident gather: the list lives on member 0 (d_quads: its memory, cap records); member r's mesh kernel writes
// its quads straight into segment r = [r * (cap / N), (r + 1) * (cap / N)) over NVLink, reserving slots on a counter in
// its OWN memory (no cross-GPU atomic).  segment_counts[r] = quads in segment r.  compact != 0 closes the gaps with
// device-local copies on member 0 (segments are moved down in order), giving one contiguous list of *n_quads records.
int meso_group_mesh_device(MesoGroup* g, void* d_quads_on_member0, int64_t cap, int64_t* n_quads, int64_t* segment_counts, int compact) {
  NEED_GROUP(g);
  if (!d_quads_on_member0 || cap < g->n || !n_quads) return fail(MESO_ERR_ARGUMENT, "meso_group_mesh_device: bad argument");
  const int64_t seg = cap / g->n;
  for (int r = 0; r < g->n; r++) {
    MesoCtx* c = g->ctx[r];
    NEED_SCENE(c);
    const int rc = meso_ensure_mesh_buffers(c);
    if (rc != MESO_OK) return rc;
    launch_mesh(c->lc(), c->v, r, g->n, c->mesh_scratch(), reinterpret_cast<MesoQuad*>(d_quads_on_member0) + (size_t)r * seg, seg, c->d_quad_count);
    CK_LAST("group mesh (device list)");
    CK(cudaMemcpyAsync(&g->h_counts[r], c->d_quad_count, 8, cudaMemcpyDeviceToHost, c->stream));
  }
  int64_t total = 0;
  bool overflow = false;
  for (int r = 0; r < g->n; r++) {
    CK(cudaSetDevice(g->ctx[r]->device));
    CK(cudaStreamSynchronize(g->ctx[r]->stream));
    const int64_t nr = (int64_t)g->h_counts[r];
    if (nr > seg) overflow = true;
    if (segment_counts) segment_counts[r] = nr;
    total += nr;
  }
  *n_quads = total;
  if (overflow) return fail(MESO_ERR_ARGUMENT, "meso_group_mesh_device: a segment overflowed (cap / N records per member)");
  if (compact) {
    MesoCtx* c0 = g->ctx[0];
    CK(cudaSetDevice(c0->device));
    int64_t at = (int64_t)g->h_counts[0];
    for (int r = 1; r < g->n; r++) {
      const int64_t nr = (int64_t)g->h_counts[r];
      // ranges may overlap when a segment moves down by less than its length: memmove semantics through chunked copies
      MesoQuad* dst = reinterpret_cast<MesoQuad*>(d_quads_on_member0) + at;
      const MesoQuad* src = reinterpret_cast<MesoQuad*>(d_quads_on_member0) + (size_t)r * seg;
      const int64_t gap = (int64_t)(src - dst);
      for (int64_t done = 0; done < nr && gap > 0;) {
        const int64_t step = std::min(nr - done, gap);       // a chunk no longer than the gap never overlaps its source
        CK(cudaMemcpyAsync(dst + done, src + done, (size_t)step * sizeof(MesoQuad), cudaMemcpyDeviceToDevice, c0->stream));
        done += step;
      }
      at += nr;
    }
    CK(cudaStreamSynchronize(c0->stream));
  }
  return MESO_OK;
}

// Re-mesh the bricks of the last (replicated) carve's dirty list + their six neighbours, sharded over the members by key,
// quads delivered to the host at prefix offsets like meso_group_mesh.
int meso_group_remesh_dirty(MesoGroup* g, MesoQuad* host_quads, int64_t cap, int64_t* n_quads) {
  NEED_GROUP(g);
  if (!n_quads || cap < 0 || (cap > 0 && !host_quads)) return fail(MESO_ERR_ARGUMENT, "meso_group_remesh_dirty: bad argument");
  // the single-GPU entry point already honours the member's partition (keys are sharded by hash, see meso_remesh_dirty);
  // the members run one after the other here: dirty lists are a few hundred bricks, 0.1 ms each
  int64_t total = 0;
  for (int r = 0; r < g->n; r++) {
    int64_t nr = 0;
    const int64_t room = std::max<int64_t>(cap - total, 0);
    const int rc = meso_remesh_dirty(g->ctx[r], host_quads ? host_quads + std::min(total, cap) : nullptr, room, &nr, nullptr, 0, nullptr);
    if (rc != MESO_OK) return rc;
    total += nr;
  }
  *n_quads = total;
  if (total > cap) return fail(MESO_ERR_ARGUMENT, "meso_group_remesh_dirty: cap too small (n_quads holds the size needed)");
  return MESO_OK;
}

}  // extern "C"
