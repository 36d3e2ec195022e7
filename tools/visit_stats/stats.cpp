// experimental: how many node visits would candidate group bounds skip?  (host emulation of the device traversal)
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <cstdio>
#include <algorithm>
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

// per-child entry distances of a node (slot space), same arithmetic as node_test
static void child_tn(const DevAccel& A, uint32_t node, float ox, float oy, float oz, const RayDir& rd, float tn[8]) {
  const uint4* np = A.nodes + 5ull * node;
  const uint4 n0 = np[0], n2 = np[2], n3 = np[3], n4 = np[4];
  const float sx = __uint_as_float(((n0.w & 0xffu) + 15u) << 23) * rd.idx;
  const float sy = __uint_as_float((((n0.w >> 8) & 0xffu) + 15u) << 23) * rd.idy;
  const float sz = __uint_as_float((((n0.w >> 16) & 0xffu) + 15u) << 23) * rd.idz;
  const float bx = (__uint_as_float(n0.x) - ox) * rd.idx, by = (__uint_as_float(n0.y) - oy) * rd.idy, bz = (__uint_as_float(n0.z) - oz) * rd.idz;
  const bool negx = rd.oct & 1u, negy = rd.oct & 2u, negz = rd.oct & 4u;
  const uint32_t nx0 = negx ? n3.z : n2.x, nx1 = negx ? n3.w : n2.y;
  const uint32_t ny0 = negy ? n4.x : n2.z, ny1 = negy ? n4.y : n2.w;
  const uint32_t nz0 = negz ? n4.z : n3.x, nz1 = negz ? n4.w : n3.y;
  for (int i = 0; i < 8; ++i) {
    const float a = fmaf(qplane(nx0, nx1, i, 0x3F800000u), sx, bx), b = fmaf(qplane(ny0, ny1, i, 0x3F800000u), sy, by), c = fmaf(qplane(nz0, nz1, i, 0x3F800000u), sz, bz);
    tn[i] = fmaxf(fmaxf(a, b), fmaxf(c, 0.0f));
  }
}

struct Entry { uint2 g; float bound; float back_bound; float ctn[8]; };

// mode 0: baseline; 1: exact group bound (min tn over remaining hit inner children); 2: back-half bound over hit inner+leaf;
// 3: back-half bound inner only; 4: per-child distances (ideal)
extern "C" int visit_stats(const void* nodes288, uint32_t n_nodes, const void* packets384, uint32_t n_packets, const phos_rays* rays,
                           uint64_t n, int mode, uint64_t* out /* visits, nohit visits, skipped groups/children, hits */,
                           uint32_t* per_ray /* node visits of every ray, or null */) {
  static PackedAccel packed; static bool have = false;
  std::string e;
  if (!have) { if (!repack_accel((const RefNode*)nodes288, n_nodes, (const RefPacket*)packets384, n_packets, packed, e)) return 1; have = true; }
  DevAccel A; A.nodes = (const uint4*)packed.nodes.data(); A.tris = (const uint4*)packed.tris.data();
  uint64_t visits = 0, nohit = 0, skipped = 0, hits = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Ray r; r.flags = rays->flags[i]; if (r.flags & PHOS_MASKED) continue;
    r.ox = rays->px[i]; r.oy = rays->py[i]; r.oz = rays->pz[i]; r.wx = rays->wx[i]; r.wy = rays->wy[i]; r.wz = rays->wz[i];
    r.d = rays->d[i]; r.tri = kNoTri; r.u = r.v = 0;
    const RayDir rd = make_raydir(r.wx, r.wy, r.wz);
    Entry st[64]; int sp = 0;
    Entry cur; cur.g = make_uint2(0u, 1u | ((1u << rd.oct) << 8)); cur.bound = 0; cur.back_bound = 0; for (int k = 0; k < 8; ++k) cur.ctn[k] = 0;
    bool done = false;
    const uint64_t visits_before = visits;
    while (!done) {
      if ((cur.g.y >> 8) == 0u) { if (sp == 0) break; cur = st[--sp];
        if (mode == 1 && cur.bound > r.d) { ++skipped; cur.g.y &= 0xffu; continue; }
        continue; }
      // next key
      const uint32_t key = __ffs(cur.g.y >> 8) - 1;
      if ((mode == 2 || mode == 3) && key >= 4 && cur.back_bound > r.d) { ++skipped; cur.g.y &= 0xffu; continue; }
      if (mode == 4 && cur.ctn[key ^ rd.oct] > r.d) { ++skipped; cur.g.y &= ~(0x100u << key); continue; }
      const uint32_t node = take_child(cur.g, rd.oct);
      if (cur.g.y >> 8) st[sp++] = cur;
      NodeHits h = node_test(A, node, r.ox, r.oy, r.oz, rd, r.d * PHOS_CULL_SLACK);
      ++visits; if (!h.inner && !h.leaf) ++nohit;
      float tn[8]; child_tn(A, node, r.ox, r.oy, r.oz, rd, tn);
      while (h.leaf) {
        const uint32_t lslot = (__ffs(h.leaf) - 1) ^ rd.oct; h.leaf &= h.leaf - 1;
        const uint32_t cnt = (h.counts >> (4 * lslot)) & 15u; const uint32_t first = h.tri_base + nibble_prefix(h.counts, lslot);
        const uint4* tp = A.tris + 3ull * first;
        for (uint32_t k = 0; k < cnt; ++k, tp += 3) { float ds, us, vs;
          if (!mt_triangle(tp[0], tp[1], tp[2], r.ox, r.oy, r.oz, r.wx, r.wy, r.wz, ds, us, vs)) continue;
          if (accept_hit(A, r, ds, us, vs, first + k) && (r.flags & PHOS_SHADOW)) { done = true; break; } }
        if (done) break;
      }
      // note: leaf hit mask is consumed above, so recompute the masks for the bounds from a fresh test is not needed: use tn + hits
      Entry nx; nx.g = make_uint2(h.child_base, h.imask | (h.inner << 8));
      for (int k = 0; k < 8; ++k) nx.ctn[k] = tn[k];
      // exact bound: min tn over hit inner children except the first in key order
      float b = INFINITY, bb = INFINITY; bool first = true;
      for (uint32_t k = 0; k < 8; ++k) if (h.inner >> k & 1u) { const float t = tn[k ^ rd.oct]; if (!first) b = fminf(b, t); first = false; if (k >= 4) bb = fminf(bb, t); }
      nx.bound = b; nx.back_bound = bb;
      cur = nx;
    }
    if (r.flags & PHOS_HIT) ++hits;
    if (per_ray) per_ray[i] = (uint32_t)(visits - visits_before);
  }
  out[0] = visits; out[1] = nohit; out[2] = skipped; out[3] = hits;
  return 0;
}
