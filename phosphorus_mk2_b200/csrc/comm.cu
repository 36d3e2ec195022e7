// comm.cu — the one collective of a multi-GPU frame: ncclReduce(sum) of the W*H*4 float film onto the root rank, enqueued
// on the context's stream right behind the frame's kernels (no host synchronisation in between).
//
// The reference shards a frame over the workers of ONE process through one shared tile cursor (src/jobs/tiles.hpp:40-47,
// src/xpu/cpu.cpp:223-238) and every worker hands its tiles to the same film_t (src/film.hpp:10-16); with one process
// per GPU the ranks' films meet here instead: disjoint tiles make the sum a gather, weighted sample ranges
// make it the average.  N cuda_t devices inside one process (xpu_t::discover, src/xpu.cpp:7-9) need no reduce at all —
// each feeds film_t::add_tile from its own tiles.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy the host process already loaded — torch's under torchrun —
// or the system one), so libphos_cuda.so keeps no link-time dependency and single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"
#include "render_state.hpp"

namespace phos {

// the handful of NCCL 2.x declarations used (ABI-stable across 2.x; nccl.h is not needed to build the library)
struct NcclUniqueId {
  char internal[PHOS_NCCL_ID_BYTES];
};
using NcclComm = void*;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))dlsym(api.lib, "ncclReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Reduce && api.GetErrorString;
  });
  return api;
}

static int nccl_fail(phos_ctx* ctx, const char* what, int rc) {
  std::string msg = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error");
  return fail(ctx, PHOS_ERR_CUDA, msg.c_str());
}

// After a sample-partitioned reduce every rank has contributed alpha = 1 to every pixel: back to 1 where rendered.
__global__ void film_alpha_kernel(float* __restrict__ film, uint32_t pixels) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < pixels && film[4 * (size_t)i + 3] > 1.0f) film[4 * (size_t)i + 3] = 1.0f;
}

void comm_release(phos_ctx* ctx) {
  if (ctx->nccl_comm && ctx->nccl_owned && nccl().ok) nccl().CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nccl_owned = false;
}

}  // namespace phos

using namespace phos;

extern "C" {

int phos_cuda_comm_unique_id(uint8_t id[PHOS_NCCL_ID_BYTES]) {
  if (!id) return PHOS_ERR_INVALID;
  if (!nccl().ok) return fail(nullptr, PHOS_ERR_INVALID, "libnccl.so.2 not found: multi-GPU film reduce unavailable");
  NcclUniqueId u;
  const int rc = nccl().GetUniqueId(&u);
  if (rc) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
  memcpy(id, u.internal, PHOS_NCCL_ID_BYTES);
  return PHOS_OK;
}

int phos_cuda_comm_init(phos_ctx* ctx, int n_ranks, int rank, const uint8_t id[PHOS_NCCL_ID_BYTES]) {
  if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return PHOS_ERR_INVALID;
  if (!nccl().ok) return fail(ctx, PHOS_ERR_INVALID, "libnccl.so.2 not found: multi-GPU film reduce unavailable");
  cudaSetDevice(ctx->device);
  comm_release(ctx);
  NcclUniqueId u;
  memcpy(u.internal, id, PHOS_NCCL_ID_BYTES);
  NcclComm comm = nullptr;
  const int rc = nccl().CommInitRank(&comm, n_ranks, u, rank);
  if (rc) return nccl_fail(ctx, "ncclCommInitRank", rc);
  ctx->nccl_comm = comm;
  ctx->nccl_owned = true;
  ctx->nccl_ranks = n_ranks;
  return PHOS_OK;
}

int phos_cuda_comm_adopt(phos_ctx* ctx, void* nccl_comm, int n_ranks) {
  if (!ctx || !nccl_comm) return PHOS_ERR_INVALID;
  if (!nccl().ok) return fail(ctx, PHOS_ERR_INVALID, "libnccl.so.2 not found: multi-GPU film reduce unavailable");
  comm_release(ctx);
  ctx->nccl_comm = nccl_comm;
  ctx->nccl_owned = false;
  ctx->nccl_ranks = n_ranks;
  return PHOS_OK;
}

void phos_cuda_comm_destroy(phos_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
}

int phos_cuda_film_reduce(phos_ctx* ctx, int root) {
  if (!ctx || !ctx->render || !ctx->render->film) return PHOS_ERR_INVALID;
  if (!ctx->nccl_comm) return fail(ctx, PHOS_ERR_INVALID, "film_reduce before comm_init / comm_adopt");
  if (root < 0 || root >= ctx->nccl_ranks) return fail(ctx, PHOS_ERR_INVALID, "film_reduce: root out of range");
  cudaSetDevice(ctx->device);
  const DevCamera& c = ctx->render->camera;
  const size_t pixels = (size_t)c.width * c.height;
  float* film = ctx->render->film;
  const int rc = nccl().Reduce(film, film, pixels * 4, kNcclFloat32, kNcclSum, root, ctx->nccl_comm, ctx->stream);
  if (rc) return nccl_fail(ctx, "ncclReduce(film)", rc);
  film_alpha_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, ctx->stream>>>(film, (uint32_t)pixels);
  ctx->launches++;
  return cuda_ok(ctx, cudaGetLastError(), "film_alpha_kernel launch") ? PHOS_OK : PHOS_ERR_CUDA;
}

}  // extern "C"
