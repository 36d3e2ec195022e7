// ctx.hpp — the per-GPU context behind the opaque phos_ctx handle of include/phos_cuda.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/phos_cuda.h"
#include "../../include/phos_scene.h"

namespace phos {

#ifndef PHOS_PIPE
#define PHOS_PIPE 4
#endif
constexpr int kPipe = PHOS_PIPE;             // host-pointer trace: chunks in flight (<= 8: their cursors are d_counters[16..24))
constexpr uint64_t kPipeChunk = 1ull << 17;  // rays per chunk (512 KiB per array row, 4 MiB in): measured best on three B200
                                             // boxes with 4 chunks in flight and two up-copy streams — 1200-1230 Mrays/s per
                                             // 2 M-ray frame against 1130-1160 at 2^18 and 1110-1140 at 112-120 Ki (rows that
                                             // do not start on a 64 KiB boundary of the page-locked slab); 3 / 6 / 8 chunks in
                                             // flight: -1..5 % (profiles/r02_e2e_chunks.log)

// One staging slot of the host-pointer trace pipeline.  Copies in, traversal and copies out run on three
// DEDICATED streams (phos_ctx::s_in / s_cmp / s_out) chained by these events: with one stream per slot the
// driver put up- and down-copies of different slots on one in-order hardware queue and the whole call
// serialised (measured: 2.86 ms = 1.85 up + 0.92 down; duplex-capable link).
struct PipeLane {
  phos_rays rays = {};  // device staging
  uint64_t capacity = 0;
  cudaEvent_t ev_in = nullptr, ev_cmp = nullptr, ev_out = nullptr;
  bool used = false;
};

struct RenderState;  // render.cu

}  // namespace phos

struct phos_ctx {
  int device = 0;
  int sm_count = 0;
  int trace_blocks_per_sm = 1;
  size_t max_pitch = 0;  // cudaDeviceProp::memPitch: the longest row a pitched copy accepts
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
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr, s_in2 = nullptr;
  phos::RenderState* render = nullptr;
  float* d_rcp_table = nullptr;  // the host's RCPSS sampled over the leading mantissa bits (phos_cuda_reference_normalize)
  int rcp_bits = 11;
  bool reference_rcp = false;
  void* nccl_comm = nullptr;  // the communicator of phos_cuda_film_reduce (comm.cu)
  bool nccl_owned = false;
  int nccl_ranks = 0;
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
void comm_release(phos_ctx* ctx);         // comm.cu
bool sample_host_rcp(std::vector<float>& table, int& bits);  // rcp_table.cpp
// every vertex index of every face below its mesh's vertex count (a malformed scene must fail with PHOS_ERR_INVALID on
// the host, not fault on the device: a device fault is sticky and kills the context)
bool scene_indices_ok(const phos_scene_desc* d);
}  // namespace phos
