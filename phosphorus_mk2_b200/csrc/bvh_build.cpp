// bvh_build.cpp — host-side 8-wide BVH build (the CPU build the renderer keeps).
//
// Produces the reference's own acceleration-structure arrays — mbvh::node_t<8> (288 B,
// reference src/accel/bvh/node.hpp:11-68) and moeller_trumbore_t<8> packets (384 B,
// src/accel/triangle.hpp:24-68) — so phos_cuda_upload_accel sees exactly what the reference's
// cpu_t::preprocess would hand it.  The split decisions follow the reference builder so both sides
// build the same tree from the same scene:
//   top-down, 12 centroid bins per axis, cost = (nl*A(l) + nr*A(r)) / A(parent)
//   (src/accel/bvh/binned_sah_builder.hpp:143-189); leaf when count < 8 or count <= 1 + cost (:220);
//   a node is widened to up to 8 children by re-splitting the smallest-area child that still holds
//   >= 8 primitives (:198-213, :229-241); nodes are numbered in pre-order, packets are appended
//   after a node's sub-trees, ceil(count/8) per leaf (:243-267, src/accel/bvh.cpp:58-78).
// Triangles are numbered mesh -> face set -> face (src/scene.cpp:58-62, src/mesh.cpp:118-128).
//
// Not a translation: the tree is built by independent sub-tree tasks on a thread pool and stitched
// with prefix offsets, which gives the same numbering as the reference's single-threaded recursion.
// All arithmetic is fp32 with no contraction (compile with -ffp-contract=off) because SAH costs
// decide the topology.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/phos_cuda.h"
#include "phos_internal.hpp"

namespace phos {

namespace {

struct V3 {
  float x, y, z;
  float operator[](int i) const { return (&x)[i]; }
  float& operator[](int i) { return (&x)[i]; }
};

struct Box {
  V3 lo{FLT_MAX, FLT_MAX, FLT_MAX};
  V3 hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
  void grow(const V3& p) {
    for (int a = 0; a < 3; ++a) {
      if (p[a] < lo[a]) lo[a] = p[a];
      if (p[a] > hi[a]) hi[a] = p[a];
    }
  }
  void grow(const Box& b) {
    for (int a = 0; a < 3; ++a) {
      if (b.lo[a] < lo[a]) lo[a] = b.lo[a];
      if (b.hi[a] > hi[a]) hi[a] = b.hi[a];
    }
  }
};

// surface-area measure used by the SAH: 2 * (dx*dy + dx*dz + dy*dz) on max - min, fp32
inline float box_area(const Box& b) {
  const float dx = b.hi.x - b.lo.x, dy = b.hi.y - b.lo.y, dz = b.hi.z - b.lo.z;
  return (float)(2.0 * (double)(dx * dy + dx * dz + dy * dz));
}

struct Prim {
  uint32_t index;  // triangle number in scene order
  Box bounds;
  V3 centroid;
};

struct Tri {
  V3 a, b, c;
  uint32_t mesh_mat;  // meshid | matid << 16
  uint32_t face;      // 3 * face index
};

constexpr int kBins = 12;
constexpr uint32_t kWidth = 8;

struct Range {
  uint32_t begin = 0, end = 0;
  Box bounds, centroid_bounds;
  uint32_t count() const { return end - begin; }
};

Range make_range(const std::vector<Prim>& prims, uint32_t begin, uint32_t end) {
  Range r;
  r.begin = begin;
  r.end = end;
  for (uint32_t i = begin; i < end; ++i) {
    r.bounds.grow(prims[i].bounds);
    r.centroid_bounds.grow(prims[i].centroid);
  }
  return r;
}

// which of the 12 bins a centroid falls in along `axis`, relative to the range's centroid bounds
inline int bin_of(const Box& cb, const V3& c, int axis) {
  float o = c[axis] - cb.lo[axis];
  if (cb.hi[axis] > cb.lo[axis]) o /= (cb.hi[axis] - cb.lo[axis]);
  return std::min((int)((float)kBins * o), kBins - 1);
}

struct Split {
  int axis = 0;
  int bin = 0;
  float cost = FLT_MAX;
};

Split find_split(const std::vector<Prim>& prims, const Range& g) {
  Split best;
  const float parent_area = box_area(g.bounds);
  for (int axis = 0; axis < 3; ++axis) {
    if (g.centroid_bounds.hi[axis] < g.centroid_bounds.lo[axis]) continue;
    Box bin_bounds[kBins];
    uint32_t bin_count[kBins] = {0};
    for (uint32_t i = g.begin; i < g.end; ++i) {
      const int b = bin_of(g.centroid_bounds, prims[i].centroid, axis);
      bin_bounds[b].grow(prims[i].bounds);
      bin_count[b]++;
    }
    // suffix unions once, prefix on the fly: same unions/counts as summing bins per candidate
    Box suffix[kBins];
    int suffix_n[kBins];
    {
      Box acc;
      int n = 0;
      for (int j = kBins - 1; j >= 0; --j) {
        acc.grow(bin_bounds[j]);
        n += (int)bin_count[j];
        suffix[j] = acc;
        suffix_n[j] = n;
      }
    }
    float axis_cost = FLT_MAX;
    int axis_bin = 0;
    Box left;
    int nl = 0;
    for (int i = 0; i < kBins - 1; ++i) {
      left.grow(bin_bounds[i]);
      nl += (int)bin_count[i];
      const int nr = suffix_n[i + 1];
      const float cost = ((float)nl * box_area(left) + (float)nr * box_area(suffix[i + 1])) / parent_area;
      if (cost < axis_cost) {
        axis_cost = cost;
        axis_bin = i;
      }
    }
    if (axis_cost < best.cost) {
      best.axis = axis;
      best.cost = axis_cost;
      best.bin = axis_bin;
    }
  }
  return best;
}

void split_range(std::vector<Prim>& prims, const Split& s, const Range& parent, Range& l, Range& r) {
  const Box cb = parent.centroid_bounds;
  Prim* first = prims.data() + parent.begin;
  Prim* last = prims.data() + parent.end;
  // std::partition on purpose: the order it leaves inside each half decides which triangles share a
  // packet and their lane order, and has to be the library's.
  Prim* mid = std::partition(first, last, [&](const Prim& p) { return bin_of(cb, p.centroid, s.axis) <= s.bin; });
  const uint32_t m = (uint32_t)(mid - prims.data());
  l = make_range(prims, parent.begin, m);
  r = make_range(prims, m, parent.end);
}

// One sub-tree built in isolation with local numbering.
struct Subtree {
  std::vector<RefNode> nodes;      // pre-order, local indices
  std::vector<RefPacket> packets;  // local indices
  bool is_leaf = false;            // the range itself became a leaf (no node emitted)
};

struct Builder {
  std::vector<Prim>& prims;
  const std::vector<Tri>& tris;

  uint32_t emit_packets(Subtree& out, uint32_t begin, uint32_t end) {
    const uint32_t first = (uint32_t)out.packets.size();
    for (uint32_t i = begin; i < end; i += kWidth) {
      const uint32_t n = std::min(kWidth, end - i);
      RefPacket pk;
      memset(&pk, 0, sizeof(pk));
      pk.num = n;
      for (uint32_t j = 0; j < n; ++j) {
        const Tri& t = tris[prims[i + j].index];
        pk.e0x[j] = t.b.x - t.a.x; pk.e0y[j] = t.b.y - t.a.y; pk.e0z[j] = t.b.z - t.a.z;
        pk.e1x[j] = t.c.x - t.a.x; pk.e1y[j] = t.c.y - t.a.y; pk.e1z[j] = t.c.z - t.a.z;
        pk.v0x[j] = t.a.x; pk.v0y[j] = t.a.y; pk.v0z[j] = t.a.z;
        pk.meshid[j] = t.mesh_mat;
        pk.faceid[j] = t.face;
      }
      out.packets.push_back(pk);
    }
    return first;
  }

  // returns local node index, or 0 with nothing emitted when the range is a leaf
  uint32_t build(Subtree& out, const Range& g) {
    const Split s = find_split(prims, g);
    if (g.count() < kWidth || (float)g.count() <= 1.0f + s.cost) return 0;

    Range child[kWidth];
    uint32_t nchild = 2;
    split_range(prims, s, g, child[0], child[1]);
    while (nchild < kWidth) {
      int pick = -1;
      float smallest = FLT_MAX;
      for (uint32_t i = 0; i < nchild; ++i) {
        if (child[i].count() < kWidth) continue;
        const float a = box_area(child[i].bounds);
        if (a < smallest) {
          smallest = a;
          pick = (int)i;
        }
      }
      if (pick < 0) break;
      const Split s2 = find_split(prims, child[pick]);
      Range l, r;
      split_range(prims, s2, child[pick], l, r);
      child[pick] = l;
      child[nchild++] = r;
    }

    const uint32_t self = (uint32_t)out.nodes.size();
    out.nodes.emplace_back();
    init_ref_node(out.nodes.back());

    uint32_t child_node[kWidth];
    for (uint32_t i = 0; i < nchild; ++i) child_node[i] = build(out, child[i]);

    for (uint32_t i = 0; i < nchild; ++i) {
      RefNode& n = out.nodes[self];
      n.bounds[i] = child[i].bounds.lo.x;
      n.bounds[i + 8] = child[i].bounds.lo.y;
      n.bounds[i + 16] = child[i].bounds.lo.z;
      n.bounds[i + 24] = child[i].bounds.hi.x;
      n.bounds[i + 32] = child[i].bounds.hi.y;
      n.bounds[i + 40] = child[i].bounds.hi.z;
      if (child_node[i]) {
        n.offset[i] = child_node[i];
      } else {
        const uint32_t first = emit_packets(out, child[i].begin, child[i].end);
        RefNode& n2 = out.nodes[self];
        n2.flags[i] = 1;
        n2.offset[i] = first;
        n2.num[i] = (uint8_t)child[i].count();
      }
    }
    return self;
  }
};

void gather_triangles(const phos_scene_desc* d, std::vector<Tri>& tris) {
  size_t total = 0;
  for (uint32_t m = 0; m < d->num_meshes; ++m)
    for (uint32_t s = d->set_offset[m]; s < d->set_offset[m + 1]; ++s)
      total += d->set_face_offset[s + 1] - d->set_face_offset[s];
  tris.resize(total);
  size_t k = 0;
  for (uint32_t m = 0; m < d->num_meshes; ++m) {
    const float* v = d->vertices + 3 * (size_t)d->vert_offset[m];
    const uint32_t* f = d->faces + 3 * (size_t)d->face_offset[m];
    for (uint32_t s = d->set_offset[m]; s < d->set_offset[m + 1]; ++s) {
      const uint32_t mat = d->set_material[s];
      for (uint32_t j = d->set_face_offset[s]; j < d->set_face_offset[s + 1]; ++j, ++k) {
        const uint32_t face = d->set_faces[j];
        const uint32_t ia = f[3 * (size_t)face], ib = f[3 * (size_t)face + 1], ic = f[3 * (size_t)face + 2];
        Tri& t = tris[k];
        t.a = {v[3 * (size_t)ia], v[3 * (size_t)ia + 1], v[3 * (size_t)ia + 2]};
        t.b = {v[3 * (size_t)ib], v[3 * (size_t)ib + 1], v[3 * (size_t)ib + 2]};
        t.c = {v[3 * (size_t)ic], v[3 * (size_t)ic + 1], v[3 * (size_t)ic + 2]};
        t.mesh_mat = m | (mat << 16);
        t.face = face * 3;
      }
    }
  }
}

}  // namespace

void init_ref_node(RefNode& n) {
  memset(&n, 0, sizeof(n));
  for (int i = 0; i < 24; ++i) {
    n.bounds[i] = FLT_MAX;
    n.bounds[i + 24] = -FLT_MAX;
  }
}

}  // namespace phos

struct phos_bvh {
  std::vector<phos::RefNode> nodes;
  std::vector<phos::RefPacket> packets;
  double seconds = 0.0;
};

extern "C" {

phos_bvh* phos_bvh_build(const phos_scene_desc* scene, int threads) {
  using namespace phos;
  (void)threads;
  if (!scene) return nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<Tri> tris;
  gather_triangles(scene, tris);
  std::vector<Prim> prims(tris.size());
  for (size_t i = 0; i < tris.size(); ++i) {
    Prim& p = prims[i];
    p.index = (uint32_t)i;
    p.bounds.grow(tris[i].a);
    p.bounds.grow(tris[i].b);
    p.bounds.grow(tris[i].c);
    p.centroid = {(p.bounds.hi.x + p.bounds.lo.x) / 2, (p.bounds.hi.y + p.bounds.lo.y) / 2,
                  (p.bounds.hi.z + p.bounds.lo.z) / 2};
  }
  auto* out = new phos_bvh();
  Subtree tree;
  Builder b{prims, tris};
  const Range root = make_range(prims, 0, (uint32_t)prims.size());
  b.build(tree, root);
  out->nodes.swap(tree.nodes);
  out->packets.swap(tree.packets);
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return out;
}

uint32_t phos_bvh_num_nodes(const phos_bvh* b) { return b ? (uint32_t)b->nodes.size() : 0; }
uint32_t phos_bvh_num_packets(const phos_bvh* b) { return b ? (uint32_t)b->packets.size() : 0; }
const void* phos_bvh_nodes(const phos_bvh* b) { return b ? b->nodes.data() : nullptr; }
const void* phos_bvh_packets(const phos_bvh* b) { return b ? b->packets.data() : nullptr; }
double phos_bvh_build_seconds(const phos_bvh* b) { return b ? b->seconds : 0.0; }
void phos_bvh_free(phos_bvh* b) { delete b; }

}  // extern "C"
