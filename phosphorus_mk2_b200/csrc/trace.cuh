// trace.cuh — the ray-query kernel: persistent CTAs draining a ray stream through trace_ray()
// (trace_ray.cuh holds the traversal and the triangle test).
#pragma once
#include "trace_ray.cuh"

namespace phos {

struct TraceArgs {
  phos_rays rays;  // device pointers
  unsigned long long n;
  DevAccel accel;
  unsigned long long* cursor;    // work counter (zeroed before launch)
  unsigned long long* counters;  // [2]: nodes visited, triangles tested (kCount only)
};

// Persistent CTAs: every warp pulls 32 consecutive rays at a time from a global cursor until the
// stream is drained, so long rays do not hold back a whole CTA's worth of the stream.
template <bool kCount>
__global__ void __launch_bounds__(kTraceBlock) trace_kernel(const TraceArgs P) {
  __shared__ uint2 s_stack[kSmemStack * kTraceBlock];
  Stack st;
  st.smem = s_stack + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t n_nodes = 0, n_tris = 0;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.cursor, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= P.n) break;
    const unsigned long long i = base + lane;
    if (i >= P.n) continue;
    Ray r;
    r.flags = P.rays.flags[i];
    if (r.flags & PHOS_MASKED) continue;
    r.ox = P.rays.px[i];
    r.oy = P.rays.py[i];
    r.oz = P.rays.pz[i];
    r.wx = P.rays.wx[i];
    r.wy = P.rays.wy[i];
    r.wz = P.rays.wz[i];
    r.d = P.rays.d[i];
    r.order = 0xffffffffu;
    r.mesh = r.face = 0u;
    r.u = r.v = 0.0f;
    if (trace_ray<kCount>(P.accel, r, st, &n_nodes, &n_tris)) {
      P.rays.d[i] = r.d;
      P.rays.flags[i] = r.flags;
      if (!(r.flags & PHOS_SHADOW)) {
        P.rays.mesh[i] = r.mesh;
        P.rays.face[i] = r.face;
        P.rays.u[i] = r.u;
        P.rays.v[i] = r.v;
      }
    }
  }
  if (kCount) {
    atomicAdd(P.counters, (unsigned long long)n_nodes);
    atomicAdd(P.counters + 1, (unsigned long long)n_tris);
  }
}

}  // namespace phos
