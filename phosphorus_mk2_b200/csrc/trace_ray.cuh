// trace_ray.cuh — the per-ray pieces of the ray-query path: Möller–Trumbore, the 8-wide node test
// on the packed layout, the traversal stack and a plain one-ray traversal loop (sm_100a).
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
  // the bits of 1.0f, passed as a kernel parameter so that the plane decode is PRMT w, <selector immediate>,
  // c[0x0][..]: with a constant ptxas can see, it sometimes flips to PRMT w, <selector register>, 0x3F800000 and
  // spends 40 instructions per node materialising the eight selectors
  uint32_t one = 0x3F800000u;
};

constexpr int kTraceBlock = 128;  // threads per CTA
#ifndef PHOS_SMEM_STACK
#define PHOS_SMEM_STACK 12
#endif
constexpr int kSmemStack = PHOS_SMEM_STACK;  // traversal-stack entries per ray held in shared memory
constexpr int kSpillStack = 80;   // further entries (local memory; only touched when the stack runs deeper)

// relative slack of the slab test: covers the fp32 error of the fma plane formulation for rays far
// from the node (<= ~10 * 2^-24 relative, DESIGN.md "conservative traversal")
#define PHOS_SLAB_SLACK 1.000003814697265625f /* 1 + 2^-18 */
// relative slack of the distance cull: nodes are entered while t_near <= d * (1 + 2^-15), so a
// triangle whose Möller–Trumbore distance undercuts the current best by rounding error is still seen
#define PHOS_CULL_SLACK 1.000030517578125f /* 1 + 2^-15 */
// Measured and dropped (profiles/): packed fp32 (FFMA2) plane evaluation, 1.5-2.5 % slower than scalar FFMA
// (r01_sweep_ffma2.log); decoding some of the plane bytes off the ALU pipe — I2F.U8 on the XU pipe, or one PRMT into a
// half2 of 1024 + q widened by two HADD2.F32 on the FMA pipe — no gain in any mix (r01_sweep_decode_forms.log): the
// kernel is bound by the number of instructions issued, not by one pipe.

struct Ray {
  float ox, oy, oz;
  float wx, wy, wz;
  float d;
  uint32_t flags;
  // hit record: the triangle (index into the packed array; its ids and tie-break order are read back from
  // there when needed — three registers fewer per ray than carrying mesh, face and order) and u, v
  uint32_t tri;
  float u, v;
};

constexpr uint32_t kNoTri = 0xffffffffu;

// per-ray constants of the slab test
struct RayDir {
  float idx, idy, idz;  // ~1 / dir; a zero / tiny component is treated as +-2^-60
  uint32_t oct;         // bit a set <=> dir[a] < 0
};

// 1 / x to 1 ulp (MUFU.RCP).  The reciprocal directions only feed the conservative box test: a relative error
// e <= 2^-23 scales every plane distance of one axis by (1 + e), i.e. shifts near / far comparisons across axes by
// <= 2 e — inside PHOS_SLAB_SLACK next to the ~10 * 2^-24 of the fma form, and far inside PHOS_CULL_SLACK.  Hit
// records come from Möller–Trumbore alone, which never sees these values.  (An IEEE division is ~10 instructions
// and a slow-path branch; three of them per ray were 2.6 % of all instructions this kernel issues.)
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  float r = 1.0f / x;
#ifdef PHOS_EMUL_RCP_PERTURB  // host emulation only (tests/emul): the worst a 1-ulp reciprocal may return, either way
  static unsigned toggle = 0;
  r = nextafterf(r, (toggle++ & 1u) ? INFINITY : -INFINITY);
#endif
  return r;
#endif
}

__device__ __forceinline__ RayDir make_raydir(float wx, float wy, float wz) {
  const float tiny = 8.67361737988403547e-19f;  // 2^-60: no 0 * inf can arise
  const float dx = fabsf(wx) < tiny ? copysignf(tiny, wx) : wx;
  const float dy = fabsf(wy) < tiny ? copysignf(tiny, wy) : wy;
  const float dz = fabsf(wz) < tiny ? copysignf(tiny, wz) : wz;
  RayDir r;
  r.idx = fast_rcp(dx);
  r.idy = fast_rcp(dy);
  r.idz = fast_rcp(dz);
  r.oct = (r.idx < 0.0f ? 1u : 0u) | (r.idy < 0.0f ? 2u : 0u) | (r.idz < 0.0f ? 4u : 0u);
  return r;
}

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  // simd::vector3_t::dot (reference src/math/simd/vector.hpp:98-100): madd(x, x', madd(y, y', z * z'))
  return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, __fmul_rn(az, bz)));
}

// One ray against one packed triangle.  Returns true when the masks of triangle.hpp:154-159 hold,
// except ds < d, which the caller applies together with the tie rule.
__device__ __forceinline__ bool mt_triangle(const uint4 a, const uint4 b, const uint4 c, float ox, float oy, float oz,
                                            float wx, float wy, float wz, float& ds, float& us, float& vs) {
  const float v0x = __uint_as_float(a.x), v0y = __uint_as_float(a.y), v0z = __uint_as_float(a.z);
  const float e0x = __uint_as_float(a.w), e0y = __uint_as_float(b.x), e0z = __uint_as_float(b.y);
  const float e1x = __uint_as_float(b.z), e1y = __uint_as_float(b.w), e1z = __uint_as_float(c.x);
  const float tx = __fsub_rn(ox, v0x), ty = __fsub_rn(oy, v0y), tz = __fsub_rn(oz, v0z);
  // p = wi x e1 ; cross = msub(a, b, c * d) (vector.hpp:102-109)
  const float px = __fmaf_rn(wy, e1z, -__fmul_rn(wz, e1y));
  const float py = __fmaf_rn(wz, e1x, -__fmul_rn(wx, e1z));
  const float pz = __fmaf_rn(wx, e1y, -__fmul_rn(wy, e1x));
  const float det = dot3_rn(e0x, e0y, e0z, px, py, pz);
  const float ood = __fdiv_rn(1.0f, det);
  // q = t x e0
  const float qx = __fmaf_rn(ty, e0z, -__fmul_rn(tz, e0y));
  const float qy = __fmaf_rn(tz, e0x, -__fmul_rn(tx, e0z));
  const float qz = __fmaf_rn(tx, e0y, -__fmul_rn(ty, e0x));
  us = __fmul_rn(dot3_rn(tx, ty, tz, px, py, pz), ood);
  vs = __fmul_rn(dot3_rn(wx, wy, wz, qx, qy, qz), ood);
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

// byte i of the 8 quantised planes (w0 = slots 0-3, w1 = slots 4-7) as the float 1 + q * 2^-15:
// one PRMT drops the byte into bits 8-15 under the exponent of 1.0 — no integer-to-float conversion
// (I2F runs on the quarter-rate XU pipe; it was the top pipe of the first version of this kernel).
// `one` = DevAccel::one.
__device__ __forceinline__ float qplane(uint32_t w0, uint32_t w1, int i, uint32_t one) {
  return __uint_as_float(__byte_perm(i < 4 ? w0 : w1, one, 0x7604u + ((uint32_t)(i & 3) << 4)));
}

// (a & mask) | (b & ~mask) as ONE LOP3 (nvcc narrows the masks of the plain expression and emits two)
__device__ __forceinline__ uint32_t bitselect(uint32_t a, uint32_t b, uint32_t mask) {
#ifdef __CUDA_ARCH__
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
  return d;
#else
  return (a & mask) | (b & ~mask);
#endif
}

struct NodeHits {
  uint32_t inner;       // hit inner children, key space (key = slot ^ oct)
  uint32_t leaf;        // hit leaf children, key space
  uint32_t imask;       // inner-child mask, slot space
  uint32_t child_base;  // first inner child
  uint32_t tri_base;    // first triangle of the node's leaves
  uint32_t counts;      // 4-bit triangle count per slot
};

// Test a ray against the 8 quantised child boxes of one node.  dmax = cull distance.
__device__ __forceinline__ NodeHits node_test(const DevAccel& A, uint32_t node, float ox, float oy, float oz, const RayDir& rd,
                                              float dmax) {
  const uint4* np = A.nodes + 5ull * node;
  const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
  NodeHits h;
  h.imask = n0.w >> 24;
  h.child_base = n1.x;
  h.tri_base = n1.y;
  h.counts = n1.z;
  // plane(q) = og + (2^15 + q) * 2^e  ->  t = fma(1 + q 2^-15, 2^(e+15) / dir, (og - org) / dir)
  const float sx = __uint_as_float(((n0.w & 0xffu) + 15u) << 23) * rd.idx;
  const float sy = __uint_as_float((((n0.w >> 8) & 0xffu) + 15u) << 23) * rd.idy;
  const float sz = __uint_as_float((((n0.w >> 16) & 0xffu) + 15u) << 23) * rd.idz;
  const float bx = (__uint_as_float(n0.x) - ox) * rd.idx;
  const float by = (__uint_as_float(n0.y) - oy) * rd.idy;
  const float bz = (__uint_as_float(n0.z) - oz) * rd.idz;
  // near / far plane bytes by direction sign: lo = {n2.xy, n2.zw, n3.xy}, hi = {n3.zw, n4.xy, n4.zw}
  const bool negx = rd.oct & 1u, negy = rd.oct & 2u, negz = rd.oct & 4u;
  const uint32_t nx0 = negx ? n3.z : n2.x, nx1 = negx ? n3.w : n2.y;
  const uint32_t fx0 = negx ? n2.x : n3.z, fx1 = negx ? n2.y : n3.w;
  const uint32_t ny0 = negy ? n4.x : n2.z, ny1 = negy ? n4.y : n2.w;
  const uint32_t fy0 = negy ? n2.z : n4.x, fy1 = negy ? n2.w : n4.y;
  const uint32_t nz0 = negz ? n4.z : n3.x, nz1 = negz ? n4.w : n3.y;
  const uint32_t fz0 = negz ? n3.x : n4.z, fz1 = negz ? n3.y : n4.w;
  // A child is hit when max(tn, 0) <= min(tf, dmax) * slack, i.e. tn <= tf * slack and tn <= dmax * slack and tf >= 0
  // (and dmax >= 0, which only matters for rays that can accept nothing).  Each condition is the SIGN of a value
  // computed on the FMA pipe; one LOP3 ors the three signs and one funnel shift appends the bit: 4 ALU-pipe
  // instructions per child instead of the 6.3 of the clamp / compare / select form (+4 % Mrays/s,
  // profiles/r01_sweep_sign_hits.log).  A NaN (inf - inf) reads as "hit": conservative.
  const float dms = dmax * PHOS_SLAB_SLACK;
  const uint32_t one = A.one;
  uint32_t miss = 0u;
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    const float tnx = __fmaf_rn(qplane(nx0, nx1, i, one), sx, bx), tfx = __fmaf_rn(qplane(fx0, fx1, i, one), sx, bx);
    const float tny = __fmaf_rn(qplane(ny0, ny1, i, one), sy, by), tfy = __fmaf_rn(qplane(fy0, fy1, i, one), sy, by);
    const float tnz = __fmaf_rn(qplane(nz0, nz1, i, one), sz, bz), tfz = __fmaf_rn(qplane(fz0, fz1, i, one), sz, bz);
    const float tn = fmaxf(fmaxf(tnx, tny), tnz);
    const float tf = fminf(fminf(tfx, tfy), tfz);
    const float a = __fmaf_rn(tf, PHOS_SLAB_SLACK, -tn);
    const float b = __fsub_rn(dms, tn);
    miss = __funnelshift_l(__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(tf), miss, 1);
  }
  const uint32_t hits = ~miss & 0xffu;  // slot space
  // slot space -> key space: bit i moves to bit i ^ oct (three conditional swaps)
  uint32_t both = (hits & h.imask) | ((hits & ~h.imask & 0xffu) << 8);
  if (rd.oct & 1u) both = bitselect(both << 1, both >> 1, 0xaaaaaaaau);
  if (rd.oct & 2u) both = bitselect(both << 2, both >> 2, 0xccccccccu);
  if (rd.oct & 4u) both = bitselect(both << 4, both >> 4, 0xf0f0f0f0u);
  h.inner = both & 0xffu;
  h.leaf = both >> 8;
  return h;
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

// take the nearest pending child out of a group: returns its node index
__device__ __forceinline__ uint32_t take_child(uint2& group, uint32_t oct) {
  const uint32_t key = __ffs(group.y >> 8) - 1;
  group.y &= ~(0x100u << key);
  const uint32_t slot = key ^ oct;
  return group.x + __popc(group.y & ((1u << slot) - 1u) & 0xffu);
}

__device__ __forceinline__ uint32_t tri_order(const DevAccel& A, uint32_t tri) {
  return __ldg(reinterpret_cast<const uint32_t*>(A.tris + 3ull * tri + 2) + 3);
}

// closest-hit / any-hit acceptance of one Möller–Trumbore result against triangle `tri`; returns true
// when r changed.  An exact tie in t goes to the lower GTri::order; both orders are fetched only then
// (ties are rare, and nothing stays live in a register for them).
__device__ __forceinline__ bool accept_hit(const DevAccel& A, Ray& r, float ds, float us, float vs, uint32_t tri) {
  if (r.flags & PHOS_SHADOW) {
    if (!(ds < r.d)) return false;
    r.d = ds;
    r.tri = tri;  // "something was accepted": the surface record of a shadow ray is never written
    r.flags |= PHOS_HIT;
    return true;
  }
  if (ds < r.d || (ds == r.d && r.tri != kNoTri && tri_order(A, tri) < tri_order(A, r.tri))) {
    r.d = ds;
    r.u = us;
    r.v = vs;
    r.tri = tri;
    r.flags |= PHOS_HIT;
    return true;
  }
  return false;
}

// Plain one-ray traversal (node, then its leaves, nearest octant first).  The production kernel
// (trace.cuh) schedules the same node_test / mt_triangle / accept_hit steps warp-wide; this loop is
// the straight-line statement of the algorithm, kept for the small-batch path and the host emulation
// used by the CPU test-suite.  Returns true when r changed.
template <bool kCount>
__device__ __forceinline__ bool trace_ray(const DevAccel& A, Ray& r, Stack& st, uint32_t* n_nodes, uint32_t* n_tris) {
  const RayDir rd = make_raydir(r.wx, r.wy, r.wz);
  bool changed = false;
  st.sp = 0;
  uint2 cur = make_uint2(0u, 1u | ((1u << rd.oct) << 8));  // the root as a one-node group in slot 0
  for (;;) {
    if ((cur.y >> 8) == 0u) {
      if (st.sp == 0) break;
      cur = st.pop();
      continue;
    }
    const uint32_t node = take_child(cur, rd.oct);
    if (cur.y >> 8) st.push(cur);
    NodeHits h = node_test(A, node, r.ox, r.oy, r.oz, rd, r.d * PHOS_CULL_SLACK);
    if (kCount) ++*n_nodes;
    // leaves first (they shrink d before any child node is opened), nearest octant first
    while (h.leaf) {
      const uint32_t lslot = (__ffs(h.leaf) - 1) ^ rd.oct;
      h.leaf &= h.leaf - 1;
      const uint32_t cnt = (h.counts >> (4 * lslot)) & 15u;  // 0 for an empty slot
      const uint32_t first = h.tri_base + nibble_prefix(h.counts, lslot);
      const uint4* tp = A.tris + 3ull * first;
      for (uint32_t k = 0; k < cnt; ++k, tp += 3) {
        const uint4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (kCount) ++*n_tris;
        float ds, us, vs;
        if (!mt_triangle(a, b, c, r.ox, r.oy, r.oz, r.wx, r.wy, r.wz, ds, us, vs)) continue;
        if (accept_hit(A, r, ds, us, vs, first + k)) {
          changed = true;
          if (r.flags & PHOS_SHADOW) return true;
        }
      }
    }
    cur = make_uint2(h.child_base, h.imask | (h.inner << 8));
  }
  return changed;
}

}  // namespace phos
