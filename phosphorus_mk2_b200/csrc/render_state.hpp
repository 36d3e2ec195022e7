// render_state.hpp — device-resident scene and frame state of one context (render.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/phos_cuda.h"
#include "camera.cuh"
#include "ctx.hpp"
#include "integrate.cuh"

namespace phos {

// DevCamera (camera.cuh): camera_t (reference src/entities/camera.hpp:10-40) with the per-frame constants of
// camera::perspective_kernel_t hoisted (src/kernels/cpu/camera.hpp:113-122)

// Wavefront state for up to `capacity` concurrent paths (SoA, all in HBM).
struct Wavefront {
  uint64_t capacity = 0;
  uint64_t pixel_capacity = 0;
  phos_rays rays[2] = {};     // closest-hit ray streams, ping-pong between bounces
  phos_rays shadow = {};      // next-event shadow-ray stream
  uint32_t* slot_path[2] = {nullptr, nullptr};  // slot -> path, ping-pong
  uint32_t* count = nullptr;  // [0..1] live slots of stream 0 / 1 (device counters), [2] scratch
  float* n = nullptr;         // shading normal per slot, 3 x capacity
  float* lightw = nullptr;    // per slot, 4 x capacity: light side of the next-event estimate
  float* beta[2] = {nullptr, nullptr};  // per SLOT, 3 x capacity, ping-pong with the ray streams (wavefront.cu FrameArgs)
  float* rad[2] = {nullptr, nullptr};   // per slot: radiance gathered so far
  float* rad_final = nullptr;           // per PATH, 3 x capacity: written once when the path ends
  uint32_t* pixel = nullptr;  // film pixel (y * W + x) per tile-pixel of the current batch
  uint32_t* perm = nullptr;   // slots of every window of the stream in shading-class order (wavefront.cu, bin_window_kernel)
};
constexpr uint32_t kShadeClasses = 1024;  // 1 + 2 * materials (occluded / unoccluded each), clamped

struct RenderState {
  DevCamera camera;
  phos_tile* d_tiles = nullptr;
  unsigned long long* d_tile_offsets = nullptr;
  uint32_t tile_capacity = 0;

  DevScene scene = {};
  std::vector<void*> scene_allocs;
  uint32_t num_meshes = 0, num_materials = 0;

  float* film = nullptr;  // W * H * 4
  float* film_normals = nullptr;  // W * H * 3, the NORMALS channel (allocated by phos_cuda_enable_normals)
  float* d_jitter = nullptr;
  uint32_t jitter_capacity = 0;
  uint64_t paths_cap = 0;  // most paths in flight the device memory affords (asked once per scene upload; 0 = not yet)
  // the wavefronts of a frame: its sample batches go round-robin over up to kMaxWavefronts wavefronts on as many streams
  // (wavefront.cu); wfs[0] also serves phos_cuda_wavefront_rays
  static constexpr int kMaxWavefronts = 4;
  Wavefront wfs[kMaxWavefronts];

  int upload(phos_ctx* ctx, const phos_scene_desc* scene);
  bool set_tiles(phos_ctx* ctx, const phos_tile* tiles, const unsigned long long* offsets, uint32_t n);
  bool ensure_wavefront(phos_ctx* ctx, uint64_t paths, uint64_t pixels, int which = 0);
  void release_wavefront();
  void release();
};

}  // namespace phos
