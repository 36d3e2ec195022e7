// TEST INFRASTRUCTURE ONLY.  Host emulation of the device traversal function.
//
// Compiles phosphorus_mk2_b200/csrc/trace_ray.cuh — the exact source the CUDA kernel runs — for the
// host by mapping the handful of CUDA intrinsics it uses onto libm / compiler builtins, and links it
// with the product's own re-pack (repack.cpp).  This lets the CPU test suite (`-m "not gpu"`) check
// the packed layout and the traversal logic against the oracle without a GPU.  It is never part of
// the product: libphos_cuda.so does not contain it and no product module loads it.
#include <cuda_runtime.h>  // vector types only

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (hi << s) | (lo >> (32u - s)) : hi; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint4 __ldg(const uint4* p) { return *p; }
static inline uint32_t __ldg(const uint32_t* p) { return *p; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t t = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((t >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
}

#include "../../phosphorus_mk2_b200/csrc/phos_internal.hpp"
#include "../../phosphorus_mk2_b200/csrc/trace_ray.cuh"

extern "C" {

// returns 0 on success; err (>= 256 bytes) receives the re-pack error text otherwise
int emul_trace(const void* nodes288, uint32_t n_nodes, const void* packets384, uint32_t n_packets, const phos_rays* rays,
               uint64_t n, uint64_t* out_nodes, uint64_t* out_tris, uint32_t* out_stats /*4*/, char* err) {
  using namespace phos;
  PackedAccel packed;
  std::string e;
  if (!repack_accel((const RefNode*)nodes288, n_nodes, (const RefPacket*)packets384, n_packets, packed, e)) {
    if (err) strncpy(err, e.c_str(), 255);
    return 1;
  }
  if (out_stats) {
    out_stats[0] = (uint32_t)packed.nodes.size();
    out_stats[1] = (uint32_t)packed.tris.size();
    out_stats[2] = packed.max_depth;
    out_stats[3] = packed.max_leaf_tris;
  }
  DevAccel A;
  A.nodes = (const uint4*)packed.nodes.data();
  A.tris = (const uint4*)packed.tris.data();
  static uint2 column[kSmemStack * kTraceBlock];
  Stack st;
  st.smem = column;
  uint32_t nn = 0, nt = 0;
  uint64_t tn = 0, tt = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Ray r;
    r.flags = rays->flags[i];
    if (r.flags & PHOS_MASKED) continue;
    r.ox = rays->px[i]; r.oy = rays->py[i]; r.oz = rays->pz[i];
    r.wx = rays->wx[i]; r.wy = rays->wy[i]; r.wz = rays->wz[i];
    r.d = rays->d[i];
    r.tri = kNoTri;
    r.u = r.v = 0.0f;
    nn = nt = 0;
    if (trace_ray<true>(A, r, st, &nn, &nt)) {
      rays->d[i] = r.d;
      rays->flags[i] = r.flags;
      if (!(r.flags & PHOS_SHADOW)) {
        const uint4 ids = A.tris[3ull * r.tri + 2];
        rays->mesh[i] = ids.y; rays->face[i] = ids.z; rays->u[i] = r.u; rays->v[i] = r.v;
      }
    }
    tn += nn;
    tt += nt;
  }
  if (out_nodes) *out_nodes = tn;
  if (out_tris) *out_tris = tt;
  return 0;
}

// FNV-1a digest of the packed arrays (re-pack determinism across thread counts / knobs)
int emul_repack_digest(const void* nodes288, uint32_t n_nodes, const void* packets384, uint32_t n_packets, uint64_t* digest,
                       uint32_t* out_stats /*4*/) {
  using namespace phos;
  PackedAccel packed;
  std::string e;
  if (!repack_accel((const RefNode*)nodes288, n_nodes, (const RefPacket*)packets384, n_packets, packed, e)) return 1;
  uint64_t h = 1469598103934665603ull;
  auto eat = [&](const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  };
  eat(packed.nodes.data(), packed.nodes.size() * sizeof(GNode));
  eat(packed.tris.data(), packed.tris.size() * sizeof(GTri));
  *digest = h;
  if (out_stats) {
    out_stats[0] = (uint32_t)packed.nodes.size();
    out_stats[1] = (uint32_t)packed.tris.size();
    out_stats[2] = packed.max_depth;
    out_stats[3] = packed.max_leaf_tris;
  }
  return 0;
}
}
