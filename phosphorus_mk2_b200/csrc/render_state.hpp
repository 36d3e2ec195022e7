// render_state.hpp — device-resident scene and frame state of one context (render.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"

namespace phos {

// camera_t (reference src/entities/camera.hpp:10-40) with the per-frame constants of
// camera::perspective_kernel_t hoisted (src/kernels/cpu/camera.hpp:113-122)
struct DevCamera {
  float m[16];  // to_world, row-vector convention
  float zoom;   // 1.12 * tan(fov / 2)
  float stepx, stepy, ratio;
  uint32_t width, height;
};

struct RenderState {
  DevCamera camera;
  phos_tile* d_tiles = nullptr;
  unsigned long long* d_tile_offsets = nullptr;
  uint32_t tile_capacity = 0;

  int upload(phos_ctx* ctx, const phos_scene_desc* scene);
  bool set_tiles(phos_ctx* ctx, const phos_tile* tiles, const unsigned long long* offsets, uint32_t n);
  void release();
};

}  // namespace phos
