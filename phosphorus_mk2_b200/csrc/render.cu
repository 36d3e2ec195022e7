// render.cu — device side of the frame pipeline around the ray-query kernel: scene upload,
// primary-ray generation, (wavefront shading / NEE / integration / film: see integrate.cuh).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"
#include "render_state.hpp"

namespace phos {

void phos_render_release(phos_ctx* ctx) {
  if (!ctx->render) return;
  ctx->render->release();
  delete ctx->render;
  ctx->render = nullptr;
}

// camera::perspective_kernel_t (reference src/kernels/cpu/camera.hpp:78-159), pinhole branch, one
// thread per pixel of one tile; the film jitter (jx, jy) is shared by every pixel of the sample
// (src/sampling.cpp:98-111).  Arithmetic in the reference's order with explicit rounding:
//   ndcx = (x - 0.5) * (1/W) - 0.5 ; ndcy = 0.5 - (y - 0.5) * (1/H)                  (:124-127)
//   d = ((ndcx + jx/W) * (W/H) * zoom, (ndcy + jy/H) * zoom, -1), normalised          (:132-135)
//   p = (0,0,0) * M + row 3 ; d = d * M, row-vector convention, mul + two fmadd        (matrix.hpp:58-104)
// One deliberate difference: normalize() multiplies by _mm256_rcp_ps(sqrt(l)) in the reference
// (~12-bit, micro-architecture specific, src/math/simd/vector.hpp:126-133); here it is the correctly
// rounded 1/sqrt(l).
__global__ void camera_rays_kernel(const DevCamera cam, const phos_tile* __restrict__ tiles,
                                   const unsigned long long* __restrict__ offsets, float jx, float jy, phos_rays out) {
  const phos_tile t = tiles[blockIdx.y];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.w * t.h) return;
  const uint32_t x = k % t.w, y = k / t.w;
  const float sx = (float)(t.x + x), sy = (float)(t.y + y);
  const float ndcy = __fsub_rn(0.5f, __fmul_rn(__fadd_rn(-0.5f, sy), cam.stepy));
  const float ndcx = __fsub_rn(__fmul_rn(__fadd_rn(-0.5f, sx), cam.stepx), 0.5f);
  float dx = __fmul_rn(__fmul_rn(__fadd_rn(ndcx, __fmul_rn(jx, cam.stepx)), cam.ratio), cam.zoom);
  float dy = __fmul_rn(__fadd_rn(ndcy, __fmul_rn(jy, cam.stepy)), cam.zoom);
  float dz = -1.0f;
  const float l = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
  const float ool = __fdiv_rn(1.0f, __fsqrt_rn(l));
  dx = __fmul_rn(dx, ool);
  dy = __fmul_rn(dy, ool);
  dz = __fmul_rn(dz, ool);
  const float* m = cam.m;
  const unsigned long long i = offsets[blockIdx.y] + k;
  out.px[i] = __fadd_rn(__fmaf_rn(0.0f, m[8], __fmaf_rn(0.0f, m[4], __fmul_rn(0.0f, m[0]))), m[12]);
  out.py[i] = __fadd_rn(__fmaf_rn(0.0f, m[9], __fmaf_rn(0.0f, m[5], __fmul_rn(0.0f, m[1]))), m[13]);
  out.pz[i] = __fadd_rn(__fmaf_rn(0.0f, m[10], __fmaf_rn(0.0f, m[6], __fmul_rn(0.0f, m[2]))), m[14]);
  out.wx[i] = __fmaf_rn(dz, m[8], __fmaf_rn(dy, m[4], __fmul_rn(dx, m[0])));
  out.wy[i] = __fmaf_rn(dz, m[9], __fmaf_rn(dy, m[5], __fmul_rn(dx, m[1])));
  out.wz[i] = __fmaf_rn(dz, m[10], __fmaf_rn(dy, m[6], __fmul_rn(dx, m[2])));
  out.d[i] = 3.402823466e+38f;
  out.flags[i] = 0u;
}

DevCamera make_camera(const phos_camera& c) {
  DevCamera d;
  memcpy(d.m, c.to_world, sizeof(d.m));
  d.zoom = 1.12f * std::tan(c.fov * 0.5f);
  d.stepx = 1.0f / (float)c.film_width;
  d.stepy = 1.0f / (float)c.film_height;
  d.ratio = (float)c.film_width / (float)c.film_height;
  d.width = c.film_width;
  d.height = c.film_height;
  return d;
}

}  // namespace phos

using namespace phos;

extern "C" {

int phos_cuda_upload_scene(phos_ctx* ctx, const phos_scene_desc* scene) {
  if (!ctx || !scene) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (scene->camera.aperture_radius != 0.0f) return fail(ctx, PHOS_ERR_INVALID, "thin-lens cameras are not supported (pinhole only)");
  if (scene->camera.film_width == 0 || scene->camera.film_height == 0) return fail(ctx, PHOS_ERR_INVALID, "empty film");
  cudaStreamSynchronize(ctx->stream);
  phos_render_release(ctx);
  ctx->render = new RenderState();
  ctx->render->camera = make_camera(scene->camera);
  return ctx->render->upload(ctx, scene);
}

int phos_cuda_camera_rays(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, float jx, float jy,
                          const phos_rays* device_rays) {
  if (!ctx || !tiles || !device_rays) return PHOS_ERR_INVALID;
  if (!ctx->render) return fail(ctx, PHOS_ERR_INVALID, "camera_rays before upload_scene");
  if (n_tiles == 0) return PHOS_OK;
  cudaSetDevice(ctx->device);
  RenderState& R = *ctx->render;
  std::vector<unsigned long long> offsets(n_tiles);
  unsigned long long total = 0;
  uint32_t max_px = 0;
  for (uint32_t i = 0; i < n_tiles; ++i) {
    if (tiles[i].x + tiles[i].w > R.camera.width || tiles[i].y + tiles[i].h > R.camera.height)
      return fail(ctx, PHOS_ERR_INVALID, "tile outside the film");
    offsets[i] = total;
    total += (unsigned long long)tiles[i].w * tiles[i].h;
    max_px = std::max(max_px, tiles[i].w * tiles[i].h);
  }
  if (max_px == 0) return PHOS_OK;
  if (!R.set_tiles(ctx, tiles, offsets.data(), n_tiles)) return PHOS_ERR_CUDA;
  const uint32_t block = 256;
  for (uint32_t first = 0; first < n_tiles; first += 65535) {  // grid.y limit
    const uint32_t cnt = std::min<uint32_t>(65535, n_tiles - first);
    dim3 grid((max_px + block - 1) / block, cnt);
    camera_rays_kernel<<<grid, block, 0, ctx->stream>>>(R.camera, R.d_tiles + first, R.d_tile_offsets + first, jx, jy,
                                                        *device_rays);
    ctx->launches++;
  }
  return cuda_ok(ctx, cudaGetLastError(), "camera_rays_kernel launch") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_flush_l2(phos_ctx* ctx) {
  if (!ctx) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t bytes = 256ull << 20;  // twice the 126 MB L2
  if (!ctx->d_flush && !cuda_ok(ctx, cudaMalloc(&ctx->d_flush, bytes), "cudaMalloc(flush buffer)")) return PHOS_ERR_CUDA;
  ctx->flush_value ^= 0xff;
  return cuda_ok(ctx, cudaMemsetAsync(ctx->d_flush, ctx->flush_value, bytes, ctx->stream), "flush_l2") ? PHOS_OK : PHOS_ERR_CUDA;
}

}  // extern "C"

namespace phos {

int RenderState::upload(phos_ctx*, const phos_scene_desc*) { return PHOS_OK; }

bool RenderState::set_tiles(phos_ctx* ctx, const phos_tile* tiles, const unsigned long long* offsets, uint32_t n) {
  if (n > tile_capacity) {
    cudaStreamSynchronize(ctx->stream);
    if (d_tiles) cudaFree(d_tiles);
    if (d_tile_offsets) cudaFree(d_tile_offsets);
    d_tiles = nullptr;
    d_tile_offsets = nullptr;
    tile_capacity = 0;
    if (!cuda_ok(ctx, cudaMalloc(&d_tiles, n * sizeof(phos_tile)), "cudaMalloc(tiles)") ||
        !cuda_ok(ctx, cudaMalloc(&d_tile_offsets, n * sizeof(unsigned long long)), "cudaMalloc(tile offsets)"))
      return false;
    tile_capacity = n;
  }
  // pageable sources: the copies are staged before the call returns, so the caller's arrays may go away
  return cuda_ok(ctx, cudaMemcpyAsync(d_tiles, tiles, n * sizeof(phos_tile), cudaMemcpyHostToDevice, ctx->stream), "upload tiles") &&
         cuda_ok(ctx, cudaMemcpyAsync(d_tile_offsets, offsets, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream),
                 "upload tile offsets");
}

void RenderState::release() {
  if (d_tiles) cudaFree(d_tiles);
  if (d_tile_offsets) cudaFree(d_tile_offsets);
  d_tiles = nullptr;
  d_tile_offsets = nullptr;
  tile_capacity = 0;
}

}  // namespace phos
