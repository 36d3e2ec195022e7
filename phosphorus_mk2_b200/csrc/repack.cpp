// repack.cpp — reference 8-wide BVH (288 B nodes / 384 B packets) -> GPU layout (80 B / 48 B).
//
// Input is mbvh_t::root / mbvh_t::triangles exactly as the reference builder publishes them
// (reference src/accel/bvh.cpp:42-47; node semantics src/accel/bvh/node.hpp:25-67; leaf = packets
// offset .. offset + ceil(num/8) - 1, src/kernels/cpu/stream_bvh_kernel.cpp:126-142).
//
// What is kept: every triangle, bit for bit (e0, e1, v0 as the reference packet holds them, ids, and
// its packet * 8 + lane position, the brute-force visiting order that breaks exact-t ties), and the
// reference's leaves — the set of triangles the SAH build decided to keep together.
// What changes, because the result of a ray query does not depend on it:
//   * the inner topology is re-grouped.  The reference widens a node by re-splitting its SMALLEST
//     child (binned_sah_builder.hpp:198-213), which yields long chains of "7 small children + 1
//     huge child" (depth 20 at 10 M triangles, 5.4 of 8 slots used).  Here the reference leaves are
//     re-grouped top-down with a binned SAH over leaf boxes, always splitting the LARGEST range
//     until a node has 8 children: a balanced 8-wide tree over the same leaves (depth ~8);
//   * child boxes are quantised outwards to 8 bits per plane on a per-node power-of-two grid — any
//     ray that meets the real box meets the quantised one;
//   * the half-empty 8-wide SoA packets (4.2 of 8 lanes used) become 48 B triangles, stored so that
//     a node's leaf triangles are contiguous; nodes are numbered breadth-first so a node's inner
//     children are contiguous (popcount addressing);
//   * children are placed in slots so that slot ^ ray-octant approximates front-to-back order;
//   * small sibling leaves may be fused (<= PHOS_REPACK_MERGE triangles) and a leaf of more than 15
//     triangles (the reference's uint8 count allows 255 and wraps beyond, node.hpp:19) is cut into
//     <= 15-triangle leaves.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <numeric>

#include "phos_internal.hpp"

namespace phos {
namespace {

struct DBox {
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
  double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  void grow(const DBox& b) {
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], b.lo[a]);
      hi[a] = std::max(hi[a], b.hi[a]);
    }
  }
  void grow(double x, double y, double z) {
    const double p[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], p[a]);
      hi[a] = std::max(hi[a], p[a]);
    }
  }
  double half_area() const {
    const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx < 0 ? 0.0 : dx * dy + dx * dz + dy * dz;
  }
};

// Box of a triangle as Möller–Trumbore sees it (v0, v0 + e0, v0 + e1), padded by one fp32 ulp of the
// largest coordinate because e0 = fl(b - a) does not reproduce b exactly.
DBox tri_box(const GTri& t) {
  DBox b;
  b.grow(t.v0x, t.v0y, t.v0z);
  b.grow((double)t.v0x + t.e0x, (double)t.v0y + t.e0y, (double)t.v0z + t.e0z);
  b.grow((double)t.v0x + t.e1x, (double)t.v0y + t.e1y, (double)t.v0z + t.e1z);
  for (int a = 0; a < 3; ++a) {
    const double m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])) * 1.2e-7 + 1e-37;
    b.lo[a] -= m;
    b.hi[a] += m;
  }
  return b;
}

constexpr uint32_t kMaxLeaf = 15;
constexpr int kBins = 16;

// one reference leaf (or a <= 15-triangle piece of an oversized one): the unit that is re-grouped
struct Prim {
  DBox box;
  double c[3];         // box centre
  uint32_t tri_begin;  // into the extracted triangle array
  uint32_t tri_count;
};

struct Range {
  uint32_t begin, end;  // into the prim index array
  DBox box;
  uint32_t tris;
  uint32_t count() const { return end - begin; }
};

struct Builder {
  const std::vector<Prim>& prims;
  std::vector<uint32_t>& idx;

  Range make_range(uint32_t b, uint32_t e) const {
    Range r;
    r.begin = b;
    r.end = e;
    r.tris = 0;
    for (uint32_t i = b; i < e; ++i) {
      r.box.grow(prims[idx[i]].box);
      r.tris += prims[idx[i]].tri_count;
    }
    return r;
  }

  // binned SAH split of a range of leaves (cost weighted by triangle count); median split when the
  // centroids do not separate.  Always produces two non-empty halves for count >= 2.
  void split(const Range& r, Range& l, Range& rr) {
    DBox cb;
    for (uint32_t i = r.begin; i < r.end; ++i) cb.grow(prims[idx[i]].c[0], prims[idx[i]].c[1], prims[idx[i]].c[2]);
    int best_axis = -1, best_bin = 0;
    double best_cost = DBL_MAX;
    for (int a = 0; a < 3; ++a) {
      const double ext = cb.hi[a] - cb.lo[a];
      if (!(ext > 0.0)) continue;
      DBox bb[kBins];
      uint32_t bn[kBins] = {0};
      const double k = kBins / ext;
      for (uint32_t i = r.begin; i < r.end; ++i) {
        const Prim& p = prims[idx[i]];
        const int b = std::min(kBins - 1, (int)((p.c[a] - cb.lo[a]) * k));
        bb[b].grow(p.box);
        bn[b] += p.tri_count;
      }
      DBox suf[kBins];
      uint32_t sufn[kBins];
      DBox acc;
      uint32_t n = 0;
      for (int j = kBins - 1; j >= 0; --j) {
        acc.grow(bb[j]);
        n += bn[j];
        suf[j] = acc;
        sufn[j] = n;
      }
      DBox left;
      uint32_t nl = 0;
      for (int j = 0; j < kBins - 1; ++j) {
        left.grow(bb[j]);
        nl += bn[j];
        if (nl == 0 || sufn[j + 1] == 0) continue;
        const double cost = left.half_area() * nl + suf[j + 1].half_area() * sufn[j + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = a;
          best_bin = j;
        }
      }
    }
    uint32_t mid = r.begin;
    if (best_axis >= 0) {
      const double k = kBins / (cb.hi[best_axis] - cb.lo[best_axis]);
      const double lo = cb.lo[best_axis];
      const int a = best_axis, bin = best_bin;
      uint32_t* m = std::partition(idx.data() + r.begin, idx.data() + r.end, [&](uint32_t i) {
        return std::min(kBins - 1, (int)((prims[i].c[a] - lo) * k)) <= bin;
      });
      mid = (uint32_t)(m - idx.data());
    }
    if (mid == r.begin || mid == r.end) {  // degenerate: split by count along the widest centroid axis
      int a = 0;
      for (int k = 1; k < 3; ++k)
        if (cb.hi[k] - cb.lo[k] > cb.hi[a] - cb.lo[a]) a = k;
      mid = r.begin + r.count() / 2;
      std::nth_element(idx.begin() + r.begin, idx.begin() + mid, idx.begin() + r.end,
                       [&](uint32_t x, uint32_t y) { return prims[x].c[a] < prims[y].c[a]; });
    }
    l = make_range(r.begin, mid);
    rr = make_range(mid, r.end);
  }
};

}  // namespace

bool repack_accel(const RefNode* nodes, uint32_t n_nodes, const RefPacket* packets, uint32_t n_packets, PackedAccel& out,
                  std::string& err) {
  out = PackedAccel();
  uint32_t merge_limit = 0;  // fuse sibling leaves up to this many triangles (0: keep the reference leaves)
  if (const char* e = std::getenv("PHOS_REPACK_MERGE")) merge_limit = std::min<uint32_t>(kMaxLeaf, (uint32_t)std::atoi(e));
  if (!nodes || !packets || n_nodes == 0 || n_packets == 0) {
    err = "empty acceleration structure (the reference builder emits no node for < 8 triangles)";
    return false;
  }

  // Real packet count of every leaf: ceil(num/8), unless the gap to the next leaf's first packet
  // shows the uint8 count wrapped (true count = num + 256 k).
  auto slot_live = [&](const RefNode& rn, int i) {
    return !(rn.bounds[i] > rn.bounds[i + 24] || rn.bounds[i + 8] > rn.bounds[i + 32] || rn.bounds[i + 16] > rn.bounds[i + 40]);
  };
  std::vector<uint32_t> leaf_offsets;
  for (uint32_t n = 0; n < n_nodes; ++n)
    for (int i = 0; i < 8; ++i)
      if (nodes[n].flags[i] == 1 && slot_live(nodes[n], i)) leaf_offsets.push_back(nodes[n].offset[i]);
  std::sort(leaf_offsets.begin(), leaf_offsets.end());
  auto leaf_packets = [&](uint32_t offset, uint32_t num) -> uint32_t {
    uint32_t npk = (num + 7) / 8;
    auto it = std::upper_bound(leaf_offsets.begin(), leaf_offsets.end(), offset);
    const uint32_t next = it == leaf_offsets.end() ? n_packets : *it;
    const uint32_t gap = next > offset ? next - offset : 0;
    if (gap > npk) {
      for (uint32_t k = 1; k <= 4096; ++k) {
        const uint32_t cand = (num + 256 * k + 7) / 8;
        if (cand == gap) return gap;
        if (cand > gap) break;
      }
    }
    return npk;
  };

  // ---- A. walk the reference tree: validate it, pull out every leaf's triangles --------------------
  std::vector<GTri> tris;
  std::vector<Prim> prims;
  tris.reserve((size_t)n_packets * 5);
  {
    std::vector<uint8_t> visited(n_nodes, 0);
    std::vector<uint32_t> stack{0};
    while (!stack.empty()) {
      const uint32_t n = stack.back();
      stack.pop_back();
      if (n >= n_nodes || visited[n]) {
        err = "node graph is not a tree (bad or repeated child index)";
        return false;
      }
      visited[n] = 1;
      const RefNode& rn = nodes[n];
      bool any = false;
      for (int i = 7; i >= 0; --i) {
        if (!slot_live(rn, i)) continue;
        any = true;
        for (int a = 0; a < 6; ++a)
          if (!std::isfinite(rn.bounds[i + 8 * a])) {
            err = "non-finite child bounds";
            return false;
          }
        if (rn.flags[i] != 1) {
          stack.push_back(rn.offset[i]);
          continue;
        }
        const uint32_t npk = leaf_packets(rn.offset[i], rn.num[i]);
        if ((uint64_t)rn.offset[i] + npk > n_packets) {
          err = "leaf packet range out of bounds";
          return false;
        }
        const uint32_t first = (uint32_t)tris.size();
        for (uint32_t p = rn.offset[i]; p < rn.offset[i] + npk; ++p) {
          const RefPacket& pk = packets[p];
          if (pk.num > 8) {
            err = "packet with more than 8 triangles";
            return false;
          }
          for (uint32_t j = 0; j < pk.num; ++j) {
            GTri t;
            t.v0x = pk.v0x[j]; t.v0y = pk.v0y[j]; t.v0z = pk.v0z[j];
            t.e0x = pk.e0x[j]; t.e0y = pk.e0y[j]; t.e0z = pk.e0z[j];
            t.e1x = pk.e1x[j]; t.e1y = pk.e1y[j]; t.e1z = pk.e1z[j];
            t.meshid = pk.meshid[j];
            t.faceid = pk.faceid[j];
            t.order = p * 8 + j;
            tris.push_back(t);
          }
        }
        // one prim per <= 15 triangles (an oversized leaf is cut along its longest axis)
        uint32_t cnt = (uint32_t)tris.size() - first;
        if (cnt > kMaxLeaf) {
          DBox cb;
          for (uint32_t k = first; k < first + cnt; ++k) cb.grow(tri_box(tris[k]));
          int ax = 0;
          for (int a = 1; a < 3; ++a)
            if (cb.hi[a] - cb.lo[a] > cb.hi[ax] - cb.lo[ax]) ax = a;
          std::sort(tris.begin() + first, tris.end(), [&](const GTri& x, const GTri& y) {
            const DBox bx = tri_box(x), by = tri_box(y);
            return bx.lo[ax] + bx.hi[ax] < by.lo[ax] + by.hi[ax];
          });
        }
        for (uint32_t b = 0; b < cnt;) {
          const uint32_t pieces = (cnt - b + kMaxLeaf - 1) / kMaxLeaf;
          const uint32_t take = (cnt - b + pieces - 1) / pieces;
          Prim p;
          p.tri_begin = first + b;
          p.tri_count = take;
          for (uint32_t k = 0; k < take; ++k) p.box.grow(tri_box(tris[first + b + k]));
          for (int a = 0; a < 3; ++a) p.c[a] = 0.5 * (p.box.lo[a] + p.box.hi[a]);
          prims.push_back(p);
          b += take;
        }
      }
      if (!any) {
        err = "inner node without children";
        return false;
      }
    }
  }
  if (tris.empty() || prims.empty()) {
    err = "acceleration structure holds no triangles";
    return false;
  }

  // ---- B. re-group the leaves into a balanced 8-wide tree, breadth first ------------------------------
  std::vector<uint32_t> idx(prims.size());
  std::iota(idx.begin(), idx.end(), 0u);
  Builder B{prims, idx};
  struct Work {
    Range range;
    uint32_t depth;
  };
  std::deque<Work> queue;
  queue.push_back({B.make_range(0, (uint32_t)prims.size()), 0});
  out.nodes.reserve(prims.size() / 4 + 16);
  out.tris.reserve(tris.size());

  auto is_leaf = [&](const Range& r) { return r.count() == 1 || r.tris <= merge_limit; };

  while (!queue.empty()) {
    const Work w = queue.front();
    queue.pop_front();
    const uint32_t self = (uint32_t)out.nodes.size();
    out.nodes.emplace_back();
    out.max_depth = std::max(out.max_depth, w.depth);

    // children: split the largest splittable range until there are 8
    Range child[8];
    int nc = 1;
    child[0] = w.range;
    if (w.range.count() == 1) {
      // a single leaf under a node of its own only happens for a one-leaf tree
    } else {
      while (nc < 8) {
        int pick = -1;
        double best = -1.0;
        for (int i = 0; i < nc; ++i) {
          if (is_leaf(child[i]) && !(nc == 1)) continue;
          if (child[i].count() < 2) continue;
          const double a = child[i].box.half_area();
          if (a > best) {
            best = a;
            pick = i;
          }
        }
        if (pick < 0) break;
        Range l, r;
        B.split(child[pick], l, r);
        child[pick] = l;
        child[nc++] = r;
      }
    }

    // ---- slot assignment: greedy max of dot(child centre - node centre, slot sign vector) -------
    DBox nb;
    for (int i = 0; i < nc; ++i) nb.grow(child[i].box);
    int slot_of[8];
    {
      double score[8][8];
      for (int i = 0; i < nc; ++i)
        for (int s = 0; s < 8; ++s) {
          double v = 0.0;
          for (int a = 0; a < 3; ++a) {
            const double rel = 0.5 * (child[i].box.lo[a] + child[i].box.hi[a]) - 0.5 * (nb.lo[a] + nb.hi[a]);
            v += ((s >> a) & 1) ? rel : -rel;
          }
          score[i][s] = v;
        }
      bool child_done[8] = {false}, slot_used[8] = {false};
      for (int round = 0; round < nc; ++round) {
        double best = -DBL_MAX;
        int bi = -1, bs = -1;
        for (int i = 0; i < nc; ++i) {
          if (child_done[i]) continue;
          for (int s = 0; s < 8; ++s)
            if (!slot_used[s] && score[i][s] > best) {
              best = score[i][s];
              bi = i;
              bs = s;
            }
        }
        child_done[bi] = true;
        slot_used[bs] = true;
        slot_of[bi] = bs;
      }
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
    for (int i = 0; i < nc; ++i) child_in_slot[slot_of[i]] = i;

    // ---- quantisation grid -----------------------------------------------------------------------
    GNode g;
    memset(&g, 0, sizeof(g));
    // Plane q of an axis sits at og + (2^15 + q) * 2^e: the kernel decodes a byte with one PRMT into
    // the float m = 1 + q * 2^-15 (bits 0x3F800000 | q << 8) and evaluates t = fma(m, 2^(e+15) / dir,
    // (og - org) / dir).  `og` is the stored fp32 origin, 2^15 cells below the grid.
    // margin, in grid cells, swallows the fp32 error of that formulation for rays that start inside or
    // near the node: the rounding of (og - org) / dir is worth <= ~2^-8 cells (DESIGN.md "conservative
    // traversal"); rays far from the node are covered by the relative slack in the kernel.  The grid
    // starts 2 margins below the true minimum so the lowest plane keeps its margin (q >= 0).
    const double margin = 1.0 / 64.0;
    double o[3], scale[3];
    uint8_t ebyte[3];
    for (int a = 0; a < 3; ++a) {
      const double ext0 = nb.hi[a] - nb.lo[a];
      int e = ext0 > 0.0 ? (int)std::ceil(std::log2(ext0 / 252.0)) : -100;
      e = std::max(-100, std::min(100, e));
      float of = 0.0f;
      for (;; ++e) {
        const double sc = std::ldexp(1.0, e);
        const double want = nb.lo[a] - 2.0 * margin * sc - 32768.0 * sc;
        of = (float)want;
        if ((double)of > want) of = std::nextafterf(of, -FLT_MAX);  // fp32 origin at or below the wanted one
        if ((nb.hi[a] - ((double)of + 32768.0 * sc)) / sc + 2.0 * margin <= 254.0 || e >= 100) break;
      }
      scale[a] = std::ldexp(1.0, e);
      o[a] = (double)of + 32768.0 * scale[a];
      ebyte[a] = (uint8_t)(e + 127);
      (&g.ox)[a] = of;
    }
    g.ex = ebyte[0];
    g.ey = ebyte[1];
    g.ez = ebyte[2];

    uint8_t* qlo[3] = {g.qlox, g.qloy, g.qloz};
    uint8_t* qhi[3] = {g.qhix, g.qhiy, g.qhiz};
    for (int s = 0; s < 8; ++s) {
      for (int a = 0; a < 3; ++a) {
        qlo[a][s] = 255;  // empty: lo > hi
        qhi[a][s] = 0;
      }
    }

    // ---- emit children in slot order ----------------------------------------------------------------
    g.child_base = (uint32_t)(out.nodes.size() + queue.size());
    g.tri_base = (uint32_t)out.tris.size();
    for (int s = 0; s < 8; ++s) {
      const int ci = child_in_slot[s];
      if (ci < 0) continue;
      const Range& c = child[ci];
      for (int a = 0; a < 3; ++a) {
        const double ql = std::floor((c.box.lo[a] - o[a]) / scale[a] - margin);
        const double qh = std::ceil((c.box.hi[a] - o[a]) / scale[a] + margin);
        if (ql < 0.0 || qh > 255.0 || o[a] + (ql + margin) * scale[a] > c.box.lo[a] ||
            o[a] + (qh - margin) * scale[a] < c.box.hi[a]) {
          err = "internal: quantised box does not contain the child box";
          return false;
        }
        qlo[a][s] = (uint8_t)ql;
        qhi[a][s] = (uint8_t)qh;
      }
      if ((c.count() == 1 || (nc > 1 && is_leaf(c))) && c.tris <= kMaxLeaf) {
        g.counts |= c.tris << (4 * s);
        out.max_leaf_tris = std::max(out.max_leaf_tris, c.tris);
        for (uint32_t i = c.begin; i < c.end; ++i) {
          const Prim& p = prims[idx[i]];
          out.tris.insert(out.tris.end(), tris.begin() + p.tri_begin, tris.begin() + p.tri_begin + p.tri_count);
        }
      } else {
        g.imask |= (uint8_t)(1u << s);
        queue.push_back({c, w.depth + 1});
      }
    }
    out.nodes[self] = g;
  }
  if (out.tris.size() != tris.size()) {
    err = "internal: triangle count changed during re-pack";
    return false;
  }
  return true;
}

}  // namespace phos
