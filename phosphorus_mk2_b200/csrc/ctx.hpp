// ctx.hpp — the per-GPU context behind the opaque phos_ctx handle of include/phos_cuda.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/phos_cuda.h"

namespace phos {

constexpr int kPipe = 4;                     // host-pointer trace: chunks in flight
constexpr uint64_t kPipeChunk = 1ull << 18;  // rays per chunk (12 MiB in, 6 MiB out): small enough that the
                                             // up- and down-link of PCIe are both busy most of the call

struct PipeLane {
  cudaStream_t stream = nullptr;
  phos_rays rays = {};  // device staging
  uint64_t capacity = 0;
};

struct RenderState;  // render.cu

}  // namespace phos

struct phos_ctx {
  int device = 0;
  int sm_count = 0;
  int trace_blocks_per_sm = 1;
  phos_options opt = {16, 16, 9};
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  void* d_nodes = nullptr;
  void* d_tris = nullptr;
  bool has_accel = false;
  phos_accel_stats stats = {};
  // [0..1] traversal counters, [8] device-trace cursor, [16..16+kPipe) pipeline cursors, [24..] render
  unsigned long long* d_counters = nullptr;
  uint64_t launches = 0;
  phos::PipeLane pipe[phos::kPipe];
  phos::RenderState* render = nullptr;
  void* d_flush = nullptr;  // L2 flush scratch (bench hygiene)
  int flush_value = 0;
};

namespace phos {
bool cuda_ok(phos_ctx* ctx, cudaError_t e, const char* what);
int fail(phos_ctx* ctx, int code, const char* msg);
bool alloc_rays(phos_ctx* ctx, uint64_t n, phos_rays& out);
void free_rays(phos_rays& r);
int launch_trace(phos_ctx* ctx, const phos_rays& dev, uint64_t n, cudaStream_t stream, unsigned long long* cursor,
                 bool count, const uint32_t* n_ptr = nullptr);
void phos_render_release(phos_ctx* ctx);  // render.cu
}  // namespace phos
