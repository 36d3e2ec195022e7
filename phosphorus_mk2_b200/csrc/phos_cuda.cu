// phos_cuda.cu — the C ABI of libphos_cuda.so (include/phos_cuda.h): context, acceleration-structure
// upload, ray-stream buffers and the trace entry points.  No CPU fallback exists anywhere in this
// library: without a CUDA device every entry point fails with PHOS_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>

#include <sched.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"
#include "phos_internal.hpp"
#include "trace.cuh"

namespace phos {

std::string g_create_error;

bool cuda_ok(phos_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  char buf[512];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  if (ctx) ctx->err = buf;
  else g_create_error = buf;
  return false;
}

int fail(phos_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  else g_create_error = msg;
  return code;
}

bool scene_indices_ok(const phos_scene_desc* d) {
  for (uint32_t m = 0; m < d->num_meshes; ++m) {
    if (d->vert_offset[m + 1] < d->vert_offset[m] || d->face_offset[m + 1] < d->face_offset[m]) return false;
    const uint32_t nv = d->vert_offset[m + 1] - d->vert_offset[m];
    const uint32_t* f = d->faces + 3 * (size_t)d->face_offset[m];
    const size_t n = 3 * (size_t)(d->face_offset[m + 1] - d->face_offset[m]);
    uint32_t worst = 0;
    for (size_t i = 0; i < n; ++i) worst = std::max(worst, f[i]);
    if (n && worst >= nv) return false;
  }
  return true;
}

void free_rays(phos_rays& r) {
  if (r.px) cudaFree(r.px);  // one slab, see alloc_rays
  memset(&r, 0, sizeof(r));
}

// 12 arrays of n 32-bit values in one allocation, each starting on a 256-byte boundary
bool alloc_rays(phos_ctx* ctx, uint64_t n, phos_rays& out) {
  const uint64_t stride = (n * 4 + 255) / 256 * 256;
  char* base = nullptr;
  if (!cuda_ok(ctx, cudaMalloc(&base, std::max<uint64_t>(stride, 256) * 12), "cudaMalloc(ray stream)")) return false;
  // slab order: the eight arrays a query reads first, then the surface record
  float** f[] = {&out.px, &out.py, &out.pz, &out.wx, &out.wy, &out.wz, &out.d};
  for (int i = 0; i < 7; ++i) *f[i] = (float*)(base + stride * i);
  out.flags = (uint32_t*)(base + stride * 7);
  out.mesh = (uint32_t*)(base + stride * 8);
  out.face = (uint32_t*)(base + stride * 9);
  out.u = (float*)(base + stride * 10);
  out.v = (float*)(base + stride * 11);
  return true;
}

// Results of one pipeline chunk straight into the caller's page-locked arrays (zero-copy stores over
// PCIe): d and flags for every ray, the surface record only where the traversal wrote one — the staging
// `mesh` row is preset to 0xffffffff, a value no hit can produce (mesh ids are 16 + 16 bits below 0xffff,
// reference src/triangle.hpp:28-31).  That is what lets the up-link carry 32 B per ray instead of 48: the
// surface record of rays that miss, and of shadow rays (it holds the sampled light's ids, spt.hpp:120-124),
// never has to visit the device to come back untouched.  Four rays per thread, 16-byte stores, grid-stride.
__global__ void __launch_bounds__(256) writeback_kernel(const phos_rays dev, const phos_rays host, uint32_t n) {
 for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4u; i < n; i += gridDim.x * blockDim.x * 4u) {
  if (i + 4u <= n) {
    const uint4 m = *reinterpret_cast<const uint4*>(dev.mesh + i);
    *reinterpret_cast<float4*>(host.d + i) = *reinterpret_cast<const float4*>(dev.d + i);
    *reinterpret_cast<uint4*>(host.flags + i) = *reinterpret_cast<const uint4*>(dev.flags + i);
    const bool all4 = m.x != 0xffffffffu && m.y != 0xffffffffu && m.z != 0xffffffffu && m.w != 0xffffffffu;
    if (all4) {
      *reinterpret_cast<uint4*>(host.mesh + i) = m;
      *reinterpret_cast<uint4*>(host.face + i) = *reinterpret_cast<const uint4*>(dev.face + i);
      *reinterpret_cast<float4*>(host.u + i) = *reinterpret_cast<const float4*>(dev.u + i);
      *reinterpret_cast<float4*>(host.v + i) = *reinterpret_cast<const float4*>(dev.v + i);
      continue;
    }
    const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k)
      if (mm[k] != 0xffffffffu) {
        host.mesh[i + k] = mm[k];
        host.face[i + k] = dev.face[i + k];
        host.u[i + k] = dev.u[i + k];
        host.v[i + k] = dev.v[i + k];
      }
    continue;
  }
  for (uint32_t k = i; k < n; ++k) {
    host.d[k] = dev.d[k];
    host.flags[k] = dev.flags[k];
    const uint32_t m = dev.mesh[k];
    if (m != 0xffffffffu) {
      host.mesh[k] = m;
      host.face[k] = dev.face[k];
      host.u[k] = dev.u[k];
      host.v[k] = dev.v[k];
    }
  }
 }
}

static const void* in_ptr(const phos_rays& r, int k) {
  const void* p[] = {r.px, r.py, r.pz, r.wx, r.wy, r.wz, r.d, r.flags, r.mesh, r.face, r.u, r.v};
  return p[k];
}
static void* out_ptr(const phos_rays& r, int k) {
  void* p[] = {r.d, r.flags, r.mesh, r.face, r.u, r.v};
  return p[k];
}

int launch_trace(phos_ctx* ctx, const phos_rays& dev, uint64_t n, cudaStream_t stream, unsigned long long* cursor,
                 bool count, const uint32_t* n_ptr) {
  if (n == 0) return PHOS_OK;
  if (n >= (1ull << 32)) return fail(ctx, PHOS_ERR_INVALID, "more than 2^32 - 1 rays in one device stream");
  TraceArgs a;
  a.rays = dev;
  a.n = n;
  a.accel.nodes = (const uint4*)ctx->d_nodes;
  a.accel.tris = (const uint4*)ctx->d_tris;
  a.cursor = cursor;
  a.counters = ctx->d_counters;
  a.n_ptr = n_ptr;  // n is then only the capacity that sizes the grid
  if (!cuda_ok(ctx, cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), stream), "memset(cursor)")) return PHOS_ERR_CUDA;
  a.tma_ok = 1;  // TMA bulk copies need 16-byte aligned sources
  for (int k = 0; k < 8; ++k) a.tma_ok &= ((uintptr_t)in_ptr(dev, k) & 15u) == 0;
  const uint64_t want = (n + (uint64_t)kChunk * kTraceWarps - 1) / ((uint64_t)kChunk * kTraceWarps);
  const int grid = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * ctx->trace_blocks_per_sm);
  // trees that fit the shared-memory stack (every re-grouped tree so far) run the instantiation without a spill tier
  // (PHOS_TRACE_DEEP=1 forces the deep instantiation: tests/test_gpu_trace.py)
  // (a ray holds at most one pending sibling group per level below the root: max_depth entries)
  const bool deep = ctx->stats.max_depth > (uint32_t)kSmemStack || std::getenv("PHOS_TRACE_DEEP") != nullptr;
#ifdef PHOS_TAIL_PROBE
  // tuning probe: per-warp start / stream-dry / end times of this launch, appended to $PHOS_TAIL_PROBE_FILE
  static unsigned long long* d_probe = nullptr;
  const size_t probe_words = 5ull * (size_t)grid * kTraceWarps;
  const char* probe_file = std::getenv("PHOS_TAIL_PROBE_FILE");
  if (!d_probe) cudaMalloc(&d_probe, 5ull * 8 * 148 * 16 * kTraceWarps);
  a.probe = probe_file ? d_probe : nullptr;
  if (probe_file) cudaMemsetAsync(d_probe, 0, probe_words * 8, stream);
#endif
  if (count) trace_kernel<true, true><<<grid, kTraceBlock, 0, stream>>>(a);
  else if (deep) trace_kernel<false, true><<<grid, kTraceBlock, 0, stream>>>(a);
  else trace_kernel<false, false><<<grid, kTraceBlock, 0, stream>>>(a);
  ctx->launches++;
  if (!cuda_ok(ctx, cudaGetLastError(), "trace_kernel launch")) return PHOS_ERR_CUDA;
#ifdef PHOS_TAIL_PROBE
  if (probe_file) {
    std::vector<unsigned long long> h(probe_words + 2);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data() + 2, d_probe, probe_words * 8, cudaMemcpyDeviceToHost);
    h[0] = probe_words / 5;
    h[1] = n;
    if (FILE* f = fopen(probe_file, "ab")) {
      fwrite(h.data(), 8, h.size(), f);
      fclose(f);
    }
  }
#endif
  return PHOS_OK;
}

}  // namespace phos

using namespace phos;

extern "C" {

int phos_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* phos_cuda_last_error(phos_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

phos_ctx* phos_cuda_create(int device, const phos_options* options) {
  const int n = phos_cuda_device_count();
  if (n <= 0) {
    g_create_error = "no CUDA device visible: libphos_cuda has no CPU fallback";
    return nullptr;
  }
  if (device < 0 || device >= n) {
    g_create_error = "device index out of range";
    return nullptr;
  }
  if (!cuda_ok(nullptr, cudaSetDevice(device), "cudaSetDevice")) return nullptr;
  cudaDeviceProp prop;
  if (!cuda_ok(nullptr, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return nullptr;
  if (prop.major < 10) {
    g_create_error = "libphos_cuda is built for sm_100a (B200) only";
    return nullptr;
  }
  phos_ctx* ctx = new phos_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_pitch = prop.memPitch;
  if (options) ctx->opt = *options;
  else ctx->opt = phos_options{16, 16, 9};
  bool ok = cuda_ok(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  for (cudaStream_t* st : {&ctx->s_in, &ctx->s_cmp, &ctx->s_out, &ctx->s_in2})
    ok = ok && cuda_ok(nullptr, cudaStreamCreateWithFlags(st, cudaStreamNonBlocking), "cudaStreamCreate");
  for (int i = 0; ok && i < kPipe; ++i)
    for (cudaEvent_t* ev : {&ctx->pipe[i].ev_in, &ctx->pipe[i].ev_cmp, &ctx->pipe[i].ev_out})
      ok = ok && cuda_ok(nullptr, cudaEventCreateWithFlags(ev, cudaEventDisableTiming), "cudaEventCreate");
  ok = ok && cuda_ok(nullptr, cudaEventCreate(&ctx->ev_begin), "cudaEventCreate") &&
       cuda_ok(nullptr, cudaEventCreate(&ctx->ev_end), "cudaEventCreate") &&
       cuda_ok(nullptr, cudaMalloc(&ctx->d_counters, 64 * sizeof(unsigned long long)), "cudaMalloc(counters)") &&
       cuda_ok(nullptr, cudaMemset(ctx->d_counters, 0, 64 * sizeof(unsigned long long)), "cudaMemset(counters)");
  // tuning probe: shared-memory carve-out of the traversal kernel in % of the maximum (the rest of the 256 KB is L1); the
  // default leaves the choice to the driver (the smallest carve-out that fits the resident CTAs)
  if (const char* e = std::getenv("PHOS_TRACE_CARVEOUT")) {
    const int pct = std::atoi(e);
    cudaFuncSetAttribute(trace_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(trace_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaGetLastError();
  }
  int blocks = 0;
  ok = ok && cuda_ok(nullptr, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, trace_kernel<false, false>, kTraceBlock, 0),
                     "occupancy(trace_kernel)");
  if (!ok) {
    phos_cuda_destroy(ctx);
    return nullptr;
  }
  ctx->trace_blocks_per_sm = std::max(1, blocks);
  return ctx;
}

void phos_cuda_destroy(phos_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kPipe; ++i) {
    free_rays(ctx->pipe[i].rays);
    for (cudaEvent_t ev : {ctx->pipe[i].ev_in, ctx->pipe[i].ev_cmp, ctx->pipe[i].ev_out})
      if (ev) cudaEventDestroy(ev);
  }
  for (cudaStream_t st : {ctx->s_in, ctx->s_cmp, ctx->s_out, ctx->s_in2})
    if (st) cudaStreamDestroy(st);
  comm_release(ctx);
  phos_render_release(ctx);
  if (ctx->d_nodes) cudaFree(ctx->d_nodes);
  if (ctx->d_tris) cudaFree(ctx->d_tris);
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->d_flush) cudaFree(ctx->d_flush);
  if (ctx->d_rcp_table) cudaFree(ctx->d_rcp_table);
  if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
  if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int phos_cuda_upload_accel(phos_ctx* ctx, const void* nodes288, uint32_t n_nodes, const void* packets384,
                           uint32_t n_packets) {
  if (!ctx) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  PackedAccel packed;
  std::string err;
  const auto t0 = std::chrono::steady_clock::now();
  if (!repack_accel((const RefNode*)nodes288, n_nodes, (const RefPacket*)packets384, n_packets, packed, err))
    return fail(ctx, PHOS_ERR_ACCEL, err.c_str());
  if (packed.max_depth + 2 > (uint32_t)(kSmemStack + kSpillStack))
    return fail(ctx, PHOS_ERR_ACCEL, "tree deeper than the traversal stack");
  const auto t1 = std::chrono::steady_clock::now();
  cudaDeviceSynchronize();
  if (ctx->d_nodes) cudaFree(ctx->d_nodes);
  if (ctx->d_tris) cudaFree(ctx->d_tris);
  ctx->d_nodes = ctx->d_tris = nullptr;
  ctx->has_accel = false;
  const size_t bn = packed.nodes.size() * sizeof(GNode), bt = packed.tris.size() * sizeof(GTri);
  if (!cuda_ok(ctx, cudaMalloc(&ctx->d_nodes, bn), "cudaMalloc(nodes)") ||
      !cuda_ok(ctx, cudaMalloc(&ctx->d_tris, bt), "cudaMalloc(triangles)") ||
      !cuda_ok(ctx, cudaMemcpy(ctx->d_nodes, packed.nodes.data(), bn, cudaMemcpyHostToDevice), "upload nodes") ||
      !cuda_ok(ctx, cudaMemcpy(ctx->d_tris, packed.tris.data(), bt, cudaMemcpyHostToDevice), "upload triangles"))
    return PHOS_ERR_CUDA;
  const auto t2 = std::chrono::steady_clock::now();
  phos_accel_stats& s = ctx->stats;
  s.ref_nodes = n_nodes;
  s.ref_packets = n_packets;
  s.nodes = (uint32_t)packed.nodes.size();
  s.triangles = (uint32_t)packed.tris.size();
  s.max_depth = packed.max_depth;
  s.max_leaf_triangles = packed.max_leaf_tris;
  s.bytes_nodes = bn;
  s.bytes_triangles = bt;
  s.repack_seconds = std::chrono::duration<double>(t1 - t0).count();
  s.upload_seconds = std::chrono::duration<double>(t2 - t1).count();
  ctx->has_accel = true;
  return PHOS_OK;
}

int phos_cuda_accel_stats(phos_ctx* ctx, phos_accel_stats* out) {
  if (!ctx || !out) return PHOS_ERR_INVALID;
  if (!ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "no acceleration structure uploaded");
  *out = ctx->stats;
  return PHOS_OK;
}

int phos_cuda_trace_device(phos_ctx* ctx, const phos_rays* rays, uint64_t n) {
  if (!ctx || !rays) return PHOS_ERR_INVALID;
  if (!ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "trace before upload_accel");
  cudaSetDevice(ctx->device);
  return launch_trace(ctx, *rays, n, ctx->stream, ctx->d_counters + 8, false);
}

int phos_cuda_trace_count(phos_ctx* ctx, const phos_rays* rays, uint64_t n, uint64_t* out_nodes, uint64_t* out_tris) {
  if (!ctx || !rays) return PHOS_ERR_INVALID;
  if (!ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "trace before upload_accel");
  cudaSetDevice(ctx->device);
  if (!cuda_ok(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream), "memset")) return PHOS_ERR_CUDA;
  const int rc = launch_trace(ctx, *rays, n, ctx->stream, ctx->d_counters + 8, true);
  if (rc) return rc;
  unsigned long long c[2];
  if (!cuda_ok(ctx, cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream), "read counters") ||
      !cuda_ok(ctx, cudaStreamSynchronize(ctx->stream), "sync"))
    return PHOS_ERR_CUDA;
  if (out_nodes) *out_nodes = c[0];
  if (out_tris) *out_tris = c[1];
  return PHOS_OK;
}

int phos_cuda_trace_profile(phos_ctx* ctx, const phos_rays* rays, uint64_t n, uint64_t out[8]) {
  if (!ctx || !rays || !out) return PHOS_ERR_INVALID;
  if (!ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "trace before upload_accel");
  cudaSetDevice(ctx->device);
  if (!cuda_ok(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream), "memset")) return PHOS_ERR_CUDA;
  const int rc = launch_trace(ctx, *rays, n, ctx->stream, ctx->d_counters + 8, true);
  if (rc) return rc;
  unsigned long long c[8];
  if (!cuda_ok(ctx, cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream), "read counters") ||
      !cuda_ok(ctx, cudaStreamSynchronize(ctx->stream), "sync"))
    return PHOS_ERR_CUDA;
  for (int k = 0; k < 8; ++k) out[k] = c[k];
  return PHOS_OK;
}

// Host-pointer trace: the stream is cut into chunks that flow through kPipe staging slots; all copies in
// run on one stream, all traversals on a second, all copies out on a third, chained per slot by events,
// so the up-link, the SMs and the down-link work on different chunks at the same time.
int phos_cuda_trace(phos_ctx* ctx, const phos_rays* rays, uint64_t n) {
  if (!ctx || !rays) return PHOS_ERR_INVALID;
  if (!ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "trace before upload_accel");
  if (n == 0) return PHOS_OK;
  cudaSetDevice(ctx->device);
  uint64_t chunk = std::min<uint64_t>(kPipeChunk, std::max<uint64_t>(32768, (n + 2 * kPipe - 1) / (2 * kPipe)));
  if (const char* e = std::getenv("PHOS_PIPE_CHUNK")) chunk = std::max<uint64_t>(1024, std::strtoull(e, nullptr, 10));  // tuning
  // chunk starts stay 16-byte aligned (TMA staging, 16-byte write-back stores); larger chunks start their rows on 64 KiB
  // boundaries of the caller's arrays (the copy engine is measurably faster on those, ctx.hpp)
  chunk = chunk >= 16384 ? (chunk + 16383) / 16384 * 16384 : (chunk + 1023) / 1024 * 1024;
  for (int i = 0; i < kPipe; ++i) {
    ctx->pipe[i].used = false;
    if (ctx->pipe[i].capacity < chunk) {
      cudaStreamSynchronize(ctx->s_out);
      free_rays(ctx->pipe[i].rays);
      ctx->pipe[i].capacity = 0;
      if (!alloc_rays(ctx, chunk, ctx->pipe[i].rays)) return PHOS_ERR_CUDA;
      ctx->pipe[i].capacity = chunk;
    }
  }
  // A stream whose twelve arrays sit in one slab at a constant stride (px, py, pz, wx, wy, wz, d, flags, mesh,
  // face, u, v — what phos_cuda_host_alloc-based callers and phos_cuda_rays_alloc produce) moves with ONE
  // pitched copy per chunk and direction instead of 12 + 6: every copy call costs the host ~20 us here,
  // as much as the wire time of a 1 MB array, so the call count, not PCIe, was the limit.
  const char* slab[12] = {(const char*)rays->px, (const char*)rays->py, (const char*)rays->pz,    (const char*)rays->wx,
                          (const char*)rays->wy, (const char*)rays->wz, (const char*)rays->d,     (const char*)rays->flags,
                          (const char*)rays->mesh, (const char*)rays->face, (const char*)rays->u, (const char*)rays->v};
  const ptrdiff_t hstride = slab[1] - slab[0];
  bool pitched = hstride >= (ptrdiff_t)(n * 4) && (size_t)hstride <= ctx->max_pitch;  // longer rows take the 12-copy path
  for (int k = 2; pitched && k < 12; ++k) pitched = slab[k] - slab[k - 1] == hstride;
  // Page-locked, device-mapped host arrays (cudaHostAlloc / cudaHostRegister; under unified addressing the
  // device pointer is the host pointer): only the eight input arrays go up (32 B per ray) and the results are
  // stored straight into the caller's arrays by writeback_kernel.  Anything else (pageable memory) takes the
  // copy-everything path: 48 B per ray up so that untouched surface records come back as they were, 24 B down.
  bool sparse = pitched && (((uintptr_t)slab[0] | (uintptr_t)hstride) & 15u) == 0 && chunk % 4 == 0;
  if (const char* e = std::getenv("PHOS_E2E_SPARSE")) sparse = sparse && std::atoi(e) != 0;
  if (sparse) {
    cudaPointerAttributes at;
    memset(&at, 0, sizeof(at));
    const char* last = slab[11] + (n - 1) * 4;
    sparse = cudaPointerGetAttributes(&at, slab[0]) == cudaSuccess && at.type == cudaMemoryTypeHost &&
             at.devicePointer == (void*)slab[0];
    sparse = sparse && cudaPointerGetAttributes(&at, last) == cudaSuccess && at.type == cudaMemoryTypeHost &&
             at.devicePointer == (void*)last;
    cudaGetLastError();  // an unregistered pointer reports an error on older drivers: not ours
    if (sparse) {  // ... and both ends belong to ONE mapped allocation (two registered regions with a hole between them do not)
      // (the driver entry point comes through the runtime: the library does not link libcuda, so it still loads on a box
      // without a driver and fails in phos_cuda_create instead)
      typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qr;
      CUdeviceptr b0 = 0, b1 = 0;
      size_t s0 = 0, s1 = 0;
      sparse = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
               ((range_fn)fn)(&b0, &s0, (CUdeviceptr)(uintptr_t)slab[0]) == CUDA_SUCCESS &&
               ((range_fn)fn)(&b1, &s1, (CUdeviceptr)(uintptr_t)last) == CUDA_SUCCESS && b0 == b1;
      cudaGetLastError();
    }
  }
  int slot = 0, launch_rc = 0;
  bool ok = true;
  // The write-back runs on a handful of CTAs: its posted writes share the link's outbound queue with the read
  // requests of the up-copies, and a full grid starves them (measured, profiles/r01_e2e_pipeline.log: 4 CTAs
  // 1.84 ms per 2 M-ray frame, a full grid 2.06-2.24 ms; moving d / flags by copy engine instead changes nothing).
  // consecutive chunks go up on two streams: the next copy is already queued at the copy engine when one ends
  // (profiles/r01_e2e_pipeline.log: 1.98-2.01 -> 1.84-1.88 ms per 2 M-ray frame at 256 Ki-ray chunks)
  int in_streams = 2;
  if (const char* e = std::getenv("PHOS_E2E_IN_STREAMS")) in_streams = std::atoi(e);
  int wb_ctas = 4;
  if (const char* e = std::getenv("PHOS_E2E_WB_CTAS")) wb_ctas = std::max(1, std::atoi(e));
  const char* dbg = std::getenv("PHOS_E2E_DEBUG");  // timing probes only (tools/e2e_probe.py): "noin" / "noout" skip a stage
  const bool dbg_noin = dbg && strstr(dbg, "noin"), dbg_noout = dbg && strstr(dbg, "noout");
  // (Chunk sizes tapering towards both ends of the stream — chunk / 4, chunk / 2, chunk ... chunk / 2, chunk / 4 — measured
  // twice, r01 and profiles/r02_e2e_chunks.log: -1 % at 128 Ki, +1 % at 256 Ki.  Not kept.)
  for (uint64_t base = 0; ok && base < n; base += chunk, slot = (slot + 1) % kPipe) {
    const uint64_t cnt = std::min(chunk, n - base);
    PipeLane& L = ctx->pipe[slot];
    const size_t dstride = (size_t)((const char*)L.rays.py - (const char*)L.rays.px);
    cudaStream_t sin = (in_streams > 1 && (slot & 1)) ? ctx->s_in2 : ctx->s_in;  // up-copies of consecutive chunks on two streams
    if (L.used) ok = cuda_ok(ctx, cudaStreamWaitEvent(sin, L.ev_out, 0), "pipeline wait");  // slot free again
    if (ok && pitched && !dbg_noin)
      ok = cuda_ok(ctx, cudaMemcpy2DAsync(L.rays.px, dstride, slab[0] + base * 4, (size_t)hstride, cnt * 4, sparse ? 8 : 12,
                                          cudaMemcpyHostToDevice, sin),
                   "H2D rays");
    if (ok && sparse) ok = cuda_ok(ctx, cudaMemsetAsync(L.rays.mesh, 0xff, cnt * 4, sin), "memset(mesh)");
    for (int k = 0; ok && !pitched && k < 12; ++k) {
      const char* src = (const char*)in_ptr(*rays, k) + base * 4;
      ok = cuda_ok(ctx, cudaMemcpyAsync((void*)in_ptr(L.rays, k), src, cnt * 4, cudaMemcpyHostToDevice, sin), "H2D rays");
    }
    ok = ok && cuda_ok(ctx, cudaEventRecord(L.ev_in, sin), "pipeline record") &&
         cuda_ok(ctx, cudaStreamWaitEvent(ctx->s_cmp, L.ev_in, 0), "pipeline wait");
    if (!ok) break;
    const int rc = launch_trace(ctx, L.rays, cnt, ctx->s_cmp, ctx->d_counters + 16 + slot, false);
    if (rc) {  // copies of earlier chunks may still be writing the caller's arrays: drain like any other failure
      launch_rc = rc;
      ok = false;
      break;
    }
    ok = cuda_ok(ctx, cudaEventRecord(L.ev_cmp, ctx->s_cmp), "pipeline record") &&
         cuda_ok(ctx, cudaStreamWaitEvent(ctx->s_out, L.ev_cmp, 0), "pipeline wait");
    if (ok && sparse && !dbg_noout) {
      phos_rays h = *rays;
      float** hf[] = {&h.px, &h.py, &h.pz, &h.wx, &h.wy, &h.wz, &h.d, &h.u, &h.v};
      for (float** q : hf) *q += base;
      h.mesh += base;
      h.face += base;
      h.flags += base;
      const unsigned full = (unsigned)((cnt + 1023) / 1024);
      writeback_kernel<<<std::min<unsigned>(full, (unsigned)wb_ctas), 256, 0, ctx->s_out>>>(L.rays, h, (uint32_t)cnt);
      ctx->launches++;
      ok = cuda_ok(ctx, cudaGetLastError(), "writeback_kernel launch");
    } else if (ok && pitched && !dbg_noout) {  // rows 6..11 of the slab: d, flags, mesh, face, u, v
      ok = cuda_ok(ctx, cudaMemcpy2DAsync((void*)(slab[6] + base * 4), (size_t)hstride, L.rays.d, dstride, cnt * 4, 6, cudaMemcpyDeviceToHost, ctx->s_out),
                   "D2H rays");
    }
    for (int k = 0; ok && !pitched && k < 6; ++k) {
      char* dst = (char*)out_ptr(*rays, k) + base * 4;
      ok = cuda_ok(ctx, cudaMemcpyAsync(dst, out_ptr(L.rays, k), cnt * 4, cudaMemcpyDeviceToHost, ctx->s_out), "D2H rays");
    }
    ok = ok && cuda_ok(ctx, cudaEventRecord(L.ev_out, ctx->s_out), "pipeline record");
    L.used = true;
  }
  if (!ok) {
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->s_in2);
    cudaStreamSynchronize(ctx->s_cmp);
    cudaStreamSynchronize(ctx->s_out);
    return launch_rc ? launch_rc : PHOS_ERR_CUDA;
  }
  return cuda_ok(ctx, cudaStreamSynchronize(ctx->s_out), "trace pipeline") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_rays_alloc(phos_ctx* ctx, uint64_t n, phos_rays* out) {
  if (!ctx || !out) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  memset(out, 0, sizeof(*out));
  return alloc_rays(ctx, n, *out) ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_rays_free(phos_ctx* ctx, phos_rays* r) {
  if (!ctx || !r) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_rays(*r);
  return PHOS_OK;
}

int phos_cuda_rays_upload(phos_ctx* ctx, const phos_rays* host, const phos_rays* device, uint64_t n) {
  if (!ctx || !host || !device) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  for (int k = 0; k < 12; ++k)
    if (!cuda_ok(ctx, cudaMemcpyAsync((void*)in_ptr(*device, k), in_ptr(*host, k), n * 4, cudaMemcpyHostToDevice, ctx->stream),
                 "rays_upload"))
      return PHOS_ERR_CUDA;
  return PHOS_OK;
}

int phos_cuda_rays_download(phos_ctx* ctx, const phos_rays* device, const phos_rays* host, uint64_t n) {
  if (!ctx || !host || !device) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  for (int k = 0; k < 12; ++k)
    if (!cuda_ok(ctx, cudaMemcpyAsync((void*)in_ptr(*host, k), in_ptr(*device, k), n * 4, cudaMemcpyDeviceToHost, ctx->stream),
                 "rays_download"))
      return PHOS_ERR_CUDA;
  return cuda_ok(ctx, cudaStreamSynchronize(ctx->stream), "rays_download sync") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_synchronize(phos_ctx* ctx) {
  if (!ctx) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return cuda_ok(ctx, cudaStreamSynchronize(ctx->stream), "synchronize") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_timer_begin(phos_ctx* ctx) {
  if (!ctx) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return cuda_ok(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream), "timer_begin") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_timer_end(phos_ctx* ctx, float* out_ms) {
  if (!ctx || !out_ms) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!cuda_ok(ctx, cudaEventRecord(ctx->ev_end, ctx->stream), "timer_end") ||
      !cuda_ok(ctx, cudaEventSynchronize(ctx->ev_end), "timer_end sync") ||
      !cuda_ok(ctx, cudaEventElapsedTime(out_ms, ctx->ev_begin, ctx->ev_end), "elapsed"))
    return PHOS_ERR_CUDA;
  return PHOS_OK;
}

uint64_t phos_cuda_launch_count(phos_ctx* ctx) { return ctx ? ctx->launches : 0; }

// Pin the calling thread (and the threads it spawns later) to the CPUs next to the GPU: page-locked ray arrays
// are then first-touched on that NUMA node and the copy engines / zero-copy stores of phos_cuda_trace do not cross
// the socket interconnect.  With 8 ranks on one box, unbound ranks share it and the host-pointer path drops to a
// third of its single-GPU rate.  Returns the number of CPUs bound to, 0 when the topology is not exposed.
int phos_cuda_bind_host_to_device(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return 0;
  char line[4096] = {0};
  const bool got = fgets(line, sizeof(line), f) != nullptr;
  fclose(f);
  if (!got) return 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {  // "0-31,64-95"
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k < 1) continue;
    if (k == 1) b = a;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
      CPU_SET(c, &set);
      ++n;
    }
  }
  if (n == 0 || sched_setaffinity(0, sizeof(set), &set) != 0) return 0;
  return n;
}

void* phos_cuda_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void phos_cuda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
