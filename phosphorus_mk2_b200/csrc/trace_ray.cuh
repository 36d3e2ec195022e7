// trace_ray.cuh — closest-hit / any-hit traversal of the packed 8-wide BVH + Möller–Trumbore (sm_100a).
//
// Replaces stream_mbvh_kernel_t::trace (reference src/kernels/cpu/stream_bvh_kernel.cpp:17-161) and
// moeller_trumbore_t<8>::iterate_rays / iterate_triangles (src/accel/triangle.hpp:125-287) for ray
// streams of any length.  Semantics per ray:
//   MASKED  -> never traced (src/kernels/cpu/detail/stream.hpp:28);
//   SHADOW  -> any-hit: first accepted triangle sets HIT and shrinks d, surface record untouched
//              (triangle.hpp:188-195), traversal of that ray stops (stream_bvh_kernel.cpp:61-64);
//   else    -> closest hit: HIT, d, mesh, face, u, v; equal-t candidates resolved to the triangle the
//              reference's brute-force kernel (linear_bvh_kernel.cpp:14-19) finds first, i.e. the
//              lowest (packet, lane) position, carried per triangle as GTri::order.
// The triangle test is the reference's arithmetic, operation for operation, with explicit
// round-to-nearest intrinsics so nvcc cannot contract or reorder it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/phos_cuda.h"

namespace phos {

struct DevAccel {
  const uint4* nodes;  // GNode, 5 x uint4 each
  const uint4* tris;   // GTri, 3 x uint4 each
};

constexpr int kTraceBlock = 128;  // threads per CTA
constexpr int kSmemStack = 24;    // traversal-stack entries per ray held in shared memory
constexpr int kSpillStack = 72;   // further entries (local memory; only touched by trees deeper than kSmemStack)

// relative slack of the slab test: covers the fp32 error of the fma plane formulation for rays far
// from the node (<= 10 * 2^-24 relative, DESIGN.md "conservative traversal")
#define PHOS_SLAB_SLACK 1.000003814697265625f /* 1 + 2^-18 */
// relative slack of the distance cull: nodes are entered while t_near <= d * (1 + 2^-15), so a
// triangle whose Möller–Trumbore distance undercuts the current best by rounding error is still seen
#define PHOS_CULL_SLACK 1.000030517578125f /* 1 + 2^-15 */

struct Ray {
  float ox, oy, oz;
  float wx, wy, wz;
  float d;
  uint32_t flags;
  // hit record
  uint32_t mesh, face, order;
  float u, v;
};

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  // simd::vector3_t::dot (reference src/math/simd/vector.hpp:98-100): madd(x, x', madd(y, y', z * z'))
  return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, __fmul_rn(az, bz)));
}

// One ray against one packed triangle.  Returns true when every mask of triangle.hpp:154-159 holds
// against tmax (strict ds < tmax is applied by the caller together with the tie rule).
__device__ __forceinline__ bool mt_triangle(const uint4 a, const uint4 b, const uint4 c, const Ray& r, float& ds,
                                            float& us, float& vs) {
  const float v0x = __uint_as_float(a.x), v0y = __uint_as_float(a.y), v0z = __uint_as_float(a.z);
  const float e0x = __uint_as_float(a.w), e0y = __uint_as_float(b.x), e0z = __uint_as_float(b.y);
  const float e1x = __uint_as_float(b.z), e1y = __uint_as_float(b.w), e1z = __uint_as_float(c.x);
  const float tx = __fsub_rn(r.ox, v0x), ty = __fsub_rn(r.oy, v0y), tz = __fsub_rn(r.oz, v0z);
  // p = wi x e1 ; cross = msub(a, b, c * d) (vector.hpp:102-109)
  const float px = __fmaf_rn(r.wy, e1z, -__fmul_rn(r.wz, e1y));
  const float py = __fmaf_rn(r.wz, e1x, -__fmul_rn(r.wx, e1z));
  const float pz = __fmaf_rn(r.wx, e1y, -__fmul_rn(r.wy, e1x));
  const float det = dot3_rn(e0x, e0y, e0z, px, py, pz);
  const float ood = __fdiv_rn(1.0f, det);
  // q = t x e0
  const float qx = __fmaf_rn(ty, e0z, -__fmul_rn(tz, e0y));
  const float qy = __fmaf_rn(tz, e0x, -__fmul_rn(tx, e0z));
  const float qz = __fmaf_rn(tx, e0y, -__fmul_rn(ty, e0x));
  us = __fmul_rn(dot3_rn(tx, ty, tz, px, py, pz), ood);
  vs = __fmul_rn(dot3_rn(r.wx, r.wy, r.wz, qx, qy, qz), ood);
  ds = __fmul_rn(dot3_rn(e1x, e1y, e1z, qx, qy, qz), ood);
  const bool xmask = (det > 0.00000001f) || (det < -0.00000001f);
  const bool umask = us >= 0.0f;
  const bool vmask = (vs >= 0.0f) && (__fadd_rn(us, vs) <= 1.0f);
  const bool dmask = ds >= 0.0f;
  return xmask && umask && vmask && dmask;
}

// sum of the 4-bit counts of slots below `slot`
__device__ __forceinline__ uint32_t nibble_prefix(uint32_t counts, uint32_t slot) {
  const uint32_t below = slot ? (counts & (0xffffffffu >> (32u - 4u * slot))) : 0u;
  const uint32_t pairs = (below & 0x0f0f0f0fu) + ((below >> 4) & 0x0f0f0f0fu);
  return (pairs * 0x01010101u) >> 24;
}

__device__ __forceinline__ float qbyte(uint32_t w0, uint32_t w1, int i) {
  const uint32_t w = i < 4 ? w0 : w1;
  return (float)((w >> (8 * (i & 3))) & 0xffu);
}

// Stack entry = a group of sibling inner nodes still to visit:
//   x = child_base of the parent, y = imask (slot space, bits 0-7) | pending hit keys (bits 8-15),
// key = slot ^ octant so that ascending keys approximate front-to-back order for this ray.
struct Stack {
  uint2* smem;  // this thread's column: entry e at smem[e * kTraceBlock]
  uint2 spill[kSpillStack];
  int sp;
  __device__ __forceinline__ void push(uint2 v) {
    if (sp < kSmemStack) smem[sp * kTraceBlock] = v;
    else if (sp < kSmemStack + kSpillStack) spill[sp - kSmemStack] = v;
    ++sp;  // deeper than both cannot happen: upload rejects trees deeper than the two stacks
  }
  __device__ __forceinline__ uint2 pop() {
    --sp;
    return sp < kSmemStack ? smem[sp * kTraceBlock] : spill[sp - kSmemStack];
  }
};

// Trace one ray; updates r in place.  Returns true when r changed (a hit was recorded).
template <bool kCount>
__device__ __forceinline__ bool trace_ray(const DevAccel& A, Ray& r, Stack& st, uint32_t* n_nodes, uint32_t* n_tris) {
  const bool shadow = (r.flags & PHOS_SHADOW) != 0;
  // 1/dir, IEEE; a zero (or denormal-small) component is treated as +-2^-60 so no 0 * inf arises
  const float tiny = 8.67361737988403547e-19f;  // 2^-60
  const float dx = fabsf(r.wx) < tiny ? copysignf(tiny, r.wx) : r.wx;
  const float dy = fabsf(r.wy) < tiny ? copysignf(tiny, r.wy) : r.wy;
  const float dz = fabsf(r.wz) < tiny ? copysignf(tiny, r.wz) : r.wz;
  const float idx = __fdiv_rn(1.0f, dx), idy = __fdiv_rn(1.0f, dy), idz = __fdiv_rn(1.0f, dz);
  const bool negx = idx < 0.0f, negy = idy < 0.0f, negz = idz < 0.0f;
  const uint32_t oct = (negx ? 1u : 0u) | (negy ? 2u : 0u) | (negz ? 4u : 0u);

  bool changed = false;
  st.sp = 0;
  uint2 cur = make_uint2(0u, 1u | ((1u << oct) << 8));  // the root as a one-node group in slot 0

  for (;;) {
    if ((cur.y >> 8) == 0u) {
      if (st.sp == 0) break;
      cur = st.pop();
      continue;
    }
    const uint32_t key = __ffs(cur.y >> 8) - 1;
    cur.y &= ~(0x100u << key);
    const uint32_t slot = key ^ oct;
    const uint32_t node = cur.x + __popc(cur.y & ((1u << slot) - 1u) & 0xffu);
    if (cur.y >> 8) st.push(cur);

    const uint4* np = A.nodes + 5ull * node;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    if (kCount) ++*n_nodes;
    const uint32_t imask = n0.w >> 24;
    const float sx = __uint_as_float((n0.w & 0xffu) << 23) * idx;
    const float sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * idy;
    const float sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * idz;
    const float bx = (__uint_as_float(n0.x) - r.ox) * idx;
    const float by = (__uint_as_float(n0.y) - r.oy) * idy;
    const float bz = (__uint_as_float(n0.z) - r.oz) * idz;
    // near / far plane bytes by direction sign: lo = {n2.xy, n2.zw, n3.xy}, hi = {n3.zw, n4.xy, n4.zw}
    const uint32_t nx0 = negx ? n3.z : n2.x, nx1 = negx ? n3.w : n2.y;
    const uint32_t fx0 = negx ? n2.x : n3.z, fx1 = negx ? n2.y : n3.w;
    const uint32_t ny0 = negy ? n4.x : n2.z, ny1 = negy ? n4.y : n2.w;
    const uint32_t fy0 = negy ? n2.z : n4.x, fy1 = negy ? n2.w : n4.y;
    const uint32_t nz0 = negz ? n4.z : n3.x, nz1 = negz ? n4.w : n3.y;
    const uint32_t fz0 = negz ? n3.x : n4.z, fz1 = negz ? n3.y : n4.w;
    const float dmax = r.d * PHOS_CULL_SLACK;
    uint32_t hit_inner = 0u, hit_leaf = 0u;  // key space
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float tnx = __fmaf_rn(qbyte(nx0, nx1, i), sx, bx), tfx = __fmaf_rn(qbyte(fx0, fx1, i), sx, bx);
      const float tny = __fmaf_rn(qbyte(ny0, ny1, i), sy, by), tfy = __fmaf_rn(qbyte(fy0, fy1, i), sy, by);
      const float tnz = __fmaf_rn(qbyte(nz0, nz1, i), sz, bz), tfz = __fmaf_rn(qbyte(fz0, fz1, i), sz, bz);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
      const float tf = fminf(fminf(tfx, tfy), fminf(tfz, dmax));
      const uint32_t h = tn <= tf * PHOS_SLAB_SLACK ? 1u : 0u;
      const uint32_t inner = (imask >> i) & 1u;
      hit_inner |= (h & inner) << (i ^ oct);
      hit_leaf |= (h & (inner ^ 1u)) << (i ^ oct);
    }

    // leaves first (they shrink d before any child node is opened), nearest octant first
    const uint32_t counts = n1.z;
    while (hit_leaf) {
      const uint32_t lkey = __ffs(hit_leaf) - 1;
      hit_leaf &= hit_leaf - 1;
      const uint32_t lslot = lkey ^ oct;
      const uint32_t cnt = (counts >> (4 * lslot)) & 15u;  // 0 for an empty slot
      const uint4* tp = A.tris + 3ull * (n1.y + nibble_prefix(counts, lslot));
      for (uint32_t k = 0; k < cnt; ++k, tp += 3) {
        const uint4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (kCount) ++*n_tris;
        float ds, us, vs;
        if (!mt_triangle(a, b, c, r, ds, us, vs)) continue;
        if (shadow) {
          if (ds < r.d) {
            r.d = ds;
            r.flags |= PHOS_HIT;
            return true;
          }
        } else if (ds < r.d || (ds == r.d && (r.flags & PHOS_HIT) && c.w < r.order)) {
          r.d = ds;
          r.u = us;
          r.v = vs;
          r.mesh = c.y;
          r.face = c.z;
          r.order = c.w;
          r.flags |= PHOS_HIT;
          changed = true;
        }
      }
    }
    cur = make_uint2(n1.x, imask | (hit_inner << 8));
  }
  return changed;
}

}  // namespace phos
