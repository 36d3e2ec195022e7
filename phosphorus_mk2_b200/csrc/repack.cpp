// repack.cpp — reference 8-wide BVH (288 B nodes / 384 B packets) -> GPU layout (80 B / 48 B).
//
// Input is mbvh_t::root / mbvh_t::triangles exactly as the reference builder publishes them
// (reference src/accel/bvh.cpp:42-47; node semantics src/accel/bvh/node.hpp:25-67; leaf = packets
// offset .. offset + ceil(num/8) - 1, src/kernels/cpu/stream_bvh_kernel.cpp:126-142).  The topology
// and the triangle set are kept; only the storage changes:
//   * nodes are renumbered breadth-first so a node's inner children are contiguous (popcount
//     addressing) and its leaf children's triangles are contiguous;
//   * child boxes are quantised outwards to 8 bits per plane — any ray that meets the real box
//     meets the quantised one, so the set of triangles a ray can reach never shrinks;
//   * the half-empty 8-wide SoA packets (4.2 of 8 lanes used on average) become 48 B triangles;
//   * children are placed in slots so that slot ^ octant approximates front-to-back order;
//   * a leaf of more than 15 triangles (the reference's uint8 count allows up to 255 and silently
//     wraps beyond, node.hpp:19) is turned into a small sub-tree of <= 15-triangle leaves.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <deque>

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
  bool empty() const { return lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]; }
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

struct Child {
  DBox box;
  int32_t ref_node = -1;    // inner child coming from the reference tree
  std::vector<GTri> tris;   // leaf child, or synthetic inner child when > 15 triangles
  bool synthetic_inner = false;
};

struct Work {
  int32_t ref_node = -1;
  std::vector<GTri> tris;  // synthetic node: split these
  uint32_t depth = 0;
};

constexpr uint32_t kMaxLeaf = 15;

// split an oversized triangle list into <= 8 spatially sorted groups
void split_synthetic(std::vector<GTri>& tris, std::vector<Child>& out) {
  DBox cb;
  std::vector<std::pair<double, uint32_t>> key(tris.size());
  for (const GTri& t : tris) cb.grow(tri_box(t));
  int axis = 0;
  for (int a = 1; a < 3; ++a)
    if (cb.hi[a] - cb.lo[a] > cb.hi[axis] - cb.lo[axis]) axis = a;
  for (uint32_t i = 0; i < tris.size(); ++i) {
    const DBox b = tri_box(tris[i]);
    key[i] = {b.lo[axis] + b.hi[axis], i};
  }
  std::sort(key.begin(), key.end());
  const size_t groups = std::min<size_t>(8, (tris.size() + kMaxLeaf - 1) / kMaxLeaf);
  for (size_t g = 0; g < groups; ++g) {
    const size_t b = tris.size() * g / groups, e = tris.size() * (g + 1) / groups;
    Child c;
    for (size_t i = b; i < e; ++i) {
      c.tris.push_back(tris[key[i].second]);
      c.box.grow(tri_box(c.tris.back()));
    }
    c.synthetic_inner = c.tris.size() > kMaxLeaf;
    out.push_back(std::move(c));
  }
}

}  // namespace

bool repack_accel(const RefNode* nodes, uint32_t n_nodes, const RefPacket* packets, uint32_t n_packets, PackedAccel& out,
                  std::string& err) {
  out = PackedAccel();
  if (!nodes || !packets || n_nodes == 0 || n_packets == 0) {
    err = "empty acceleration structure (the reference builder emits no node for < 8 triangles)";
    return false;
  }

  // Real packet count of every leaf: ceil(num/8), unless the gap to the next leaf's first packet
  // shows the uint8 count wrapped (true count = num + 256 k).
  std::vector<uint32_t> leaf_offsets;
  for (uint32_t n = 0; n < n_nodes; ++n)
    for (int i = 0; i < 8; ++i)
      if (nodes[n].flags[i] == 1 && !(nodes[n].bounds[i] > nodes[n].bounds[i + 24])) leaf_offsets.push_back(nodes[n].offset[i]);
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

  std::vector<uint8_t> visited(n_nodes, 0);
  std::deque<Work> queue;
  Work root;
  root.ref_node = 0;
  queue.push_back(std::move(root));
  out.nodes.reserve(n_nodes + 16);
  out.tris.reserve((size_t)n_packets * 5);

  while (!queue.empty()) {
    Work w = std::move(queue.front());
    queue.pop_front();
    const uint32_t self = (uint32_t)out.nodes.size();
    out.nodes.emplace_back();
    out.max_depth = std::max(out.max_depth, w.depth);

    // ---- gather children ---------------------------------------------------------------------
    std::vector<Child> children;
    if (w.ref_node >= 0) {
      if ((uint32_t)w.ref_node >= n_nodes || visited[w.ref_node]) {
        err = "node graph is not a tree (bad or repeated child index)";
        return false;
      }
      visited[w.ref_node] = 1;
      const RefNode& rn = nodes[w.ref_node];
      for (int i = 0; i < 8; ++i) {
        if (rn.bounds[i] > rn.bounds[i + 24] || rn.bounds[i + 8] > rn.bounds[i + 32] || rn.bounds[i + 16] > rn.bounds[i + 40])
          continue;  // empty slot: min = +FLT_MAX, max = -FLT_MAX
        Child c;
        c.box.lo[0] = rn.bounds[i];      c.box.lo[1] = rn.bounds[i + 8];  c.box.lo[2] = rn.bounds[i + 16];
        c.box.hi[0] = rn.bounds[i + 24]; c.box.hi[1] = rn.bounds[i + 32]; c.box.hi[2] = rn.bounds[i + 40];
        for (int a = 0; a < 3; ++a)
          if (!std::isfinite(c.box.lo[a]) || !std::isfinite(c.box.hi[a])) {
            err = "non-finite child bounds";
            return false;
          }
        if (rn.flags[i] == 1) {
          const uint32_t npk = leaf_packets(rn.offset[i], rn.num[i]);
          if (npk == 0) continue;
          if ((uint64_t)rn.offset[i] + npk > n_packets) {
            err = "leaf packet range out of bounds";
            return false;
          }
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
              c.tris.push_back(t);
            }
          }
          if (c.tris.empty()) continue;
          c.synthetic_inner = c.tris.size() > kMaxLeaf;
        } else {
          c.ref_node = (int32_t)rn.offset[i];
        }
        children.push_back(std::move(c));
      }
    } else {
      split_synthetic(w.tris, children);
    }
    if (children.empty()) {
      err = "inner node without children";
      return false;
    }

    // ---- slot assignment: greedy max of dot(child centre - node centre, slot sign vector) -------
    DBox nb;
    for (const Child& c : children) nb.grow(c.box);
    int slot_of[8];
    {
      const size_t nc = children.size();
      double score[8][8];
      for (size_t i = 0; i < nc; ++i)
        for (int s = 0; s < 8; ++s) {
          double v = 0.0;
          for (int a = 0; a < 3; ++a) {
            const double rel = 0.5 * (children[i].box.lo[a] + children[i].box.hi[a]) - 0.5 * (nb.lo[a] + nb.hi[a]);
            v += ((s >> a) & 1) ? rel : -rel;
          }
          score[i][s] = v;
        }
      bool child_done[8] = {false}, slot_used[8] = {false};
      for (size_t round = 0; round < nc; ++round) {
        double best = -DBL_MAX;
        int bi = -1, bs = -1;
        for (size_t i = 0; i < nc; ++i) {
          if (child_done[i]) continue;
          for (int s = 0; s < 8; ++s)
            if (!slot_used[s] && score[i][s] > best) {
              best = score[i][s];
              bi = (int)i;
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
    for (size_t i = 0; i < children.size(); ++i) child_in_slot[slot_of[i]] = (int)i;

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
      Child& c = children[ci];
      for (int a = 0; a < 3; ++a) {
        double ql = std::floor((c.box.lo[a] - o[a]) / scale[a] - margin);
        double qh = std::ceil((c.box.hi[a] - o[a]) / scale[a] + margin);
        if (ql < 0.0 || qh > 255.0 || o[a] + (ql + margin) * scale[a] > c.box.lo[a] ||
            o[a] + (qh - margin) * scale[a] < c.box.hi[a]) {
          err = "internal: quantised box does not contain the child box";
          return false;
        }
        qlo[a][s] = (uint8_t)ql;
        qhi[a][s] = (uint8_t)qh;
      }
      if (c.ref_node >= 0 || c.synthetic_inner) {
        g.imask |= (uint8_t)(1u << s);
        Work cw;
        cw.ref_node = c.ref_node;
        cw.depth = w.depth + 1;
        if (c.synthetic_inner) cw.tris = std::move(c.tris);
        queue.push_back(std::move(cw));
      } else {
        g.counts |= (uint32_t)c.tris.size() << (4 * s);
        out.max_leaf_tris = std::max<uint32_t>(out.max_leaf_tris, (uint32_t)c.tris.size());
        for (const GTri& t : c.tris) out.tris.push_back(t);
      }
    }
    out.nodes[self] = g;
  }
  if (out.tris.empty()) {
    err = "acceleration structure holds no triangles";
    return false;
  }
  return true;
}

}  // namespace phos
