// TEST / TUNING INFRASTRUCTURE ONLY (never part of the product).  A warp-level simulator of trace_kernel's scheduling on
// the host: 32 lanes run the device's own node_test / mt_triangle / accept_hit (trace_ray.cuh compiled for the host, as
// tests/emul does) under the kernel's two-ballot loop — refill at >= refill_min idle lanes, one node step or one triangle
// step per iteration by the weighted vote, eager commit of the next node — and count warp-level steps.  It exists to cost
// scheduling POLICIES in issue slots (291 per warp node step, 216 per warp triangle step, ~9 per ray: the fit of
// profiles/ncu_capture.json) before any of them is built on the GPU:
//   defer = 0  the shipped policy: a lane with pending leaf triangles waits for a triangle step
//   defer = Q  such a lane keeps taking node steps (its hit leaves go into a queue of up to Q leaf sets) while the warp
//              runs node steps anyway; closest-hit rays then traverse with a stale distance bound until their leaves are tested
//   defer_shadow_only: only any-hit rays may defer (they have no bound to go stale)
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
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
  const uint64_t t = ((uint64_t)y << 32) | x; uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((t >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
}
#include "../../phosphorus_mk2_b200/csrc/phos_internal.hpp"
#include "../../phosphorus_mk2_b200/csrc/trace_ray.cuh"
using namespace phos;

struct LeafSet { uint32_t lt, counts, base, tptr; };  // lt: hit leaf slots (key space, bits 0-7) | triangles left in the open leaf (bits 8-11)
struct Lane {
  Ray r; RayDir rd; bool has_ray = false; uint64_t ridx = 0;
  uint32_t node = 0xffffffffu;
  std::vector<uint2> stack;
  std::vector<LeafSet> leaves;  // FIFO; [0] is the one being tested
};
struct Params { int refill_min, tri_bias, defer, defer_shadow_only, tri_min_lanes, chunks_per_warp_hint; };

extern "C" int warp_sim(const void* nodes288, uint32_t n_nodes, const void* packets384, uint32_t n_packets, const phos_rays* rays,
                        uint64_t n, int n_warps, const Params* P, uint64_t* out /*[10]*/, float* out_d, uint32_t* out_flags) {
  static PackedAccel packed; static bool have = false;
  std::string e;
  if (!have) { if (!repack_accel((const RefNode*)nodes288, n_nodes, (const RefPacket*)packets384, n_packets, packed, e)) return 1; have = true; }
  DevAccel A; A.nodes = (const uint4*)packed.nodes.data(); A.tris = (const uint4*)packed.tris.data();
  uint64_t w_node = 0, w_tri = 0, l_node = 0, l_tri = 0, nodes = 0, tris = 0, iters = 0, refills = 0, traced = 0, spec_nodes = 0;
  const uint64_t n_chunks = (n + 31) / 32;
  for (int w = 0; w < n_warps; ++w) {
    Lane L[32];
    uint64_t chunk = w;       // static round-robin chunks: warp w takes chunks w, w + n_warps, ...
    uint64_t cur_base = 0; uint32_t cur_cnt = 0, taken = 0;
    auto advance = [&]() -> bool { if (chunk >= n_chunks) return false; cur_base = chunk * 32; cur_cnt = (uint32_t)std::min<uint64_t>(32, n - cur_base); taken = 0; chunk += n_warps; return true; };
    for (;;) {
      ++iters;
      bool tri_work[32], node_work[32]; int n_tri = 0, n_node = 0, n_busy = 0;
      for (int l = 0; l < 32; ++l) {
        Lane& a = L[l];
        tri_work[l] = a.has_ray && !a.leaves.empty();
        const bool may_defer = P->defer > 0 && (int)a.leaves.size() < P->defer && (!P->defer_shadow_only || (a.r.flags & PHOS_SHADOW));
        node_work[l] = a.has_ray && a.node != 0xffffffffu && (!tri_work[l] || may_defer);
        n_tri += tri_work[l]; n_node += node_work[l]; n_busy += (tri_work[l] || node_work[l]);
      }
      if (32 - n_busy >= P->refill_min) {
        for (int l = 0; l < 32; ++l) {  // retire
          Lane& a = L[l];
          if (a.has_ray && !tri_work[l] && a.node == 0xffffffffu) { out_d[a.ridx] = a.r.d; out_flags[a.ridx] = a.r.flags; a.has_ray = false; }
        }
        bool any = false;
        for (int l = 0; l < 32; ++l) {
          Lane& a = L[l];
          if (a.has_ray) continue;
          if (taken == cur_cnt && !advance()) break;
          const uint64_t k = cur_base + taken++;
          const uint32_t fl = rays->flags[k];
          out_d[k] = rays->d[k]; out_flags[k] = fl;
          if (fl & PHOS_MASKED) continue;
          a.r.ox = rays->px[k]; a.r.oy = rays->py[k]; a.r.oz = rays->pz[k]; a.r.wx = rays->wx[k]; a.r.wy = rays->wy[k]; a.r.wz = rays->wz[k];
          a.r.d = rays->d[k]; a.r.flags = fl; a.r.tri = kNoTri; a.r.u = a.r.v = 0;
          a.rd = make_raydir(a.r.wx, a.r.wy, a.r.wz); a.ridx = k; a.node = 0; a.stack.clear(); a.leaves.clear(); a.has_ray = true; any = true; ++traced;
        }
        if (any) { ++refills; continue; }
        bool left = false;
        for (int l = 0; l < 32; ++l) left = left || L[l].has_ray;
        if (!left) break;
        if (n_busy == 0) continue;
      }
      bool tri_step = P->tri_bias * n_tri >= n_node;
      if (P->defer > 0 && P->tri_min_lanes > 0) tri_step = n_tri >= P->tri_min_lanes || n_node == 0 || P->tri_bias * n_tri >= n_node;
      if (n_tri == 0) tri_step = false;
      if (tri_step) {
        ++w_tri; l_tri += n_tri;
        for (int l = 0; l < 32; ++l) {
          if (!tri_work[l]) continue;
          Lane& a = L[l]; LeafSet& s = a.leaves.front();
          if ((s.lt >> 8) == 0u) {
            const uint32_t slot = (__ffs(s.lt) - 1) ^ a.rd.oct; s.lt &= s.lt - 1u;
            s.lt |= ((s.counts >> (4u * slot)) & 15u) << 8; s.tptr = s.base + nibble_prefix(s.counts, slot);
          }
          bool done = false;
          if (s.lt >> 8) {
            const int cnt = s.lt >= 0x200u ? 2 : 1;
            for (int k = 0; k < cnt && !done; ++k) {
              const uint4* tp = A.tris + 3ull * (s.tptr + k); float ds, us, vs; ++tris;
              if (mt_triangle(tp[0], tp[1], tp[2], a.r.ox, a.r.oy, a.r.oz, a.r.wx, a.r.wy, a.r.wz, ds, us, vs) && accept_hit(A, a.r, ds, us, vs, s.tptr + k))
                done = (a.r.flags & PHOS_SHADOW) != 0u;
            }
            s.tptr += cnt; s.lt -= 0x100u * cnt;
          }
          if (done) { a.leaves.clear(); a.node = 0xffffffffu; a.stack.clear(); }
          else if (s.lt == 0u) a.leaves.erase(a.leaves.begin());
        }
      } else {
        ++w_node; l_node += n_node;
        for (int l = 0; l < 32; ++l) {
          if (!node_work[l]) continue;
          Lane& a = L[l];
          if (tri_work[l]) ++spec_nodes;
          const NodeHits h = node_test(A, a.node, a.r.ox, a.r.oy, a.r.oz, a.rd, a.r.d * PHOS_CULL_SLACK); ++nodes;
          if (h.leaf) a.leaves.push_back(LeafSet{h.leaf, h.counts, h.tri_base, 0u});
          uint2 g = make_uint2(h.child_base, h.imask | (h.inner << 8));
          if ((g.y >> 8) == 0u && !a.stack.empty()) { g = a.stack.back(); a.stack.pop_back(); }
          if (g.y >> 8) { a.node = take_child(g, a.rd.oct); if (g.y >> 8) a.stack.push_back(g); } else a.node = 0xffffffffu;
        }
      }
    }
  }
  out[0] = w_node; out[1] = w_tri; out[2] = l_node; out[3] = l_tri; out[4] = nodes; out[5] = tris; out[6] = iters; out[7] = refills; out[8] = traced; out[9] = spec_nodes;
  return 0;
}
