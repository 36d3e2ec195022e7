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
// Not a translation: the top of the tree is expanded node by node, every range below 32 Ki primitives
// is built as an independent task on all host threads, and the pieces are laid out with prefix
// offsets — the same arrays as the reference's single-threaded recursion, bit for bit.
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
#include "host_parallel.hpp"
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

// Ranges beyond this many primitives are scanned in parallel chunks.  Every scan of the build reduces
// with min / max and integer counts only, so chunking cannot change a single bit of the result.
constexpr size_t kScanGrain = 1u << 16;
int g_scan_threads = 1;  // set by phos_bvh_build for the duration of one build

Range make_range(const std::vector<Prim>& prims, uint32_t begin, uint32_t end) {
  Range r;
  r.begin = begin;
  r.end = end;
  const size_t n = end - begin;
  if (n <= 2 * kScanGrain || g_scan_threads <= 1) {
    for (uint32_t i = begin; i < end; ++i) {
      r.bounds.grow(prims[i].bounds);
      r.centroid_bounds.grow(prims[i].centroid);
    }
    return r;
  }
  std::vector<Range> part((n + kScanGrain - 1) / kScanGrain);
  parallel_chunks(n, kScanGrain, g_scan_threads, [&](size_t b, size_t e, size_t c) {
    for (size_t i = begin + b; i < begin + e; ++i) {
      part[c].bounds.grow(prims[i].bounds);
      part[c].centroid_bounds.grow(prims[i].centroid);
    }
  });
  for (const Range& q : part) {
    r.bounds.grow(q.bounds);
    r.centroid_bounds.grow(q.centroid_bounds);
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
    if (g.count() <= 2 * kScanGrain || g_scan_threads <= 1) {
      for (uint32_t i = g.begin; i < g.end; ++i) {
        const int b = bin_of(g.centroid_bounds, prims[i].centroid, axis);
        bin_bounds[b].grow(prims[i].bounds);
        bin_count[b]++;
      }
    } else {
      struct Part {
        Box bounds[kBins];
        uint32_t count[kBins] = {0};
      };
      std::vector<Part> part((g.count() + kScanGrain - 1) / kScanGrain);
      parallel_chunks(g.count(), kScanGrain, g_scan_threads, [&](size_t b0, size_t b1, size_t c) {
        for (size_t i = g.begin + b0; i < g.begin + b1; ++i) {
          const int b = bin_of(g.centroid_bounds, prims[i].centroid, axis);
          part[c].bounds[b].grow(prims[i].bounds);
          part[c].count[b]++;
        }
      });
      for (const Part& q : part)
        for (int b = 0; b < kBins; ++b) {
          bin_bounds[b].grow(q.bounds[b]);
          bin_count[b] += q.count[b];
        }
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

  // The children of the node a range becomes (binned_sah_builder.hpp:215-241): one SAH split, then
  // the smallest-area child that still holds >= 8 primitives is re-split until there are 8 children.
  // Returns 0 when the range is a leaf.
  uint32_t node_children(const Range& g, Range child[kWidth]) {
    const Split s = find_split(prims, g);
    if (g.count() < kWidth || (float)g.count() <= 1.0f + s.cost) return 0;
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
    return nchild;
  }

  static void set_child_bounds(RefNode& n, uint32_t i, const Range& c) {
    n.bounds[i] = c.bounds.lo.x;
    n.bounds[i + 8] = c.bounds.lo.y;
    n.bounds[i + 16] = c.bounds.lo.z;
    n.bounds[i + 24] = c.bounds.hi.x;
    n.bounds[i + 32] = c.bounds.hi.y;
    n.bounds[i + 40] = c.bounds.hi.z;
  }

  // One whole sub-tree, depth first on the calling thread: nodes in pre-order, a node's leaf packets
  // after the packets of its sub-trees (binned_sah_builder.hpp:243-267).  Returns the local node index,
  // or 0 with nothing emitted when the range is a leaf.
  uint32_t build(Subtree& out, const Range& g) {
    Range child[kWidth];
    const uint32_t nchild = node_children(g, child);
    if (nchild == 0) return 0;

    const uint32_t self = (uint32_t)out.nodes.size();
    out.nodes.emplace_back();
    init_ref_node(out.nodes.back());

    uint32_t child_node[kWidth];
    for (uint32_t i = 0; i < nchild; ++i) child_node[i] = build(out, child[i]);

    for (uint32_t i = 0; i < nchild; ++i) {
      set_child_bounds(out.nodes[self], i, child[i]);
      if (child_node[i]) {
        out.nodes[self].offset[i] = child_node[i];
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

// ---- the parallel build ---------------------------------------------------------------------------------
// The top of the tree (ranges of >= `grain` primitives) is expanded node by node into a skeleton of
// Pieces; every smaller range is one task that builds its whole sub-tree depth first into a private
// Subtree.  Afterwards the pieces are laid out in the reference's numbering (node: pre-order; packets:
// sub-trees first, then the node's own leaves) by a size pass and an offset pass, and every task's arrays
// are copied to their final place once, in parallel.  The result is the single-threaded recursion's, bit
// for bit: sub-trees own disjoint slices of the primitive array and all scans are order-independent.
struct Piece {
  bool whole = false;  // a finished sub-tree (task) rather than one skeleton node
  Subtree sub;         // whole: local numbering, root at 0; sub.is_leaf: the range is a leaf
  Range range;
  uint32_t nchild = 0;
  Range child[kWidth];
  std::unique_ptr<Piece> kid[kWidth];  // null: cannot be an inner node (fewer than 8 primitives)
  uint64_t n_nodes = 0, n_packets = 0; // totals of this piece's sub-tree (size pass)
  uint64_t node_base = 0, packet_base = 0;
  bool leaf() const { return whole ? sub.is_leaf : nchild == 0; }
};

struct ParallelBuild {
  Builder& b;
  TaskBag& bag;
  uint32_t grain;

  void expand(Piece* p) {
    if (p->range.count() < grain) {  // one task: the whole sub-tree
      p->whole = true;
      const uint32_t root = b.build(p->sub, p->range);
      p->sub.is_leaf = (root == 0 && p->sub.nodes.empty());
      return;
    }
    p->nchild = b.node_children(p->range, p->child);
    for (uint32_t i = 0; i < p->nchild; ++i) {
      if (p->child[i].count() < kWidth) continue;  // a leaf for certain
      p->kid[i].reset(new Piece());
      Piece* k = p->kid[i].get();
      k->range = p->child[i];
      bag.add([this, k] { expand(k); });
    }
  }

  static uint64_t leaf_packets(const Range& r) { return (r.count() + kWidth - 1) / kWidth; }

  void measure(Piece* p) {
    if (p->whole) {
      p->n_nodes = p->sub.nodes.size();
      p->n_packets = p->sub.packets.size();
      return;
    }
    if (p->nchild == 0) return;
    p->n_nodes = 1;
    for (uint32_t i = 0; i < p->nchild; ++i) {
      Piece* k = p->kid[i].get();
      if (k) measure(k);
      if (k && !k->leaf()) {
        p->n_nodes += k->n_nodes;
        p->n_packets += k->n_packets;
      } else {
        p->n_packets += leaf_packets(p->child[i]);
      }
    }
  }

  // offsets top-down; skeleton nodes and their leaf packets are written here, tasks are collected for
  // the parallel copy
  void place(Piece* p, uint64_t node_base, uint64_t packet_base, std::vector<RefNode>& nodes, std::vector<RefPacket>& packets,
             std::vector<Piece*>& tasks) {
    p->node_base = node_base;
    p->packet_base = packet_base;
    if (p->whole) {
      tasks.push_back(p);
      return;
    }
    RefNode& n = nodes[node_base];
    init_ref_node(n);
    uint64_t cn = node_base + 1, cp = packet_base;
    for (uint32_t i = 0; i < p->nchild; ++i) {
      Builder::set_child_bounds(n, i, p->child[i]);
      Piece* k = p->kid[i].get();
      if (k && !k->leaf()) {
        n.offset[i] = (uint32_t)cn;
        place(k, cn, cp, nodes, packets, tasks);
        cn += k->n_nodes;
        cp += k->n_packets;
      }
    }
    for (uint32_t i = 0; i < p->nchild; ++i) {
      Piece* k = p->kid[i].get();
      if (k && !k->leaf()) continue;
      Subtree tmp;
      b.emit_packets(tmp, p->child[i].begin, p->child[i].end);
      std::copy(tmp.packets.begin(), tmp.packets.end(), packets.begin() + cp);
      n.flags[i] = 1;
      n.offset[i] = (uint32_t)cp;
      n.num[i] = (uint8_t)p->child[i].count();
      cp += tmp.packets.size();
    }
  }

  static void copy_task(const Piece* p, std::vector<RefNode>& nodes, std::vector<RefPacket>& packets) {
    const uint32_t nb = (uint32_t)p->node_base, pb = (uint32_t)p->packet_base;
    for (size_t j = 0; j < p->sub.nodes.size(); ++j) {
      RefNode n = p->sub.nodes[j];
      for (uint32_t i = 0; i < kWidth; ++i) {
        if (n.bounds[i] > n.bounds[i + 24]) continue;  // unused slot (init_ref_node)
        n.offset[i] += n.flags[i] == 1 ? pb : nb;
      }
      nodes[nb + j] = n;
    }
    std::copy(p->sub.packets.begin(), p->sub.packets.end(), packets.begin() + pb);
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

namespace phos {
bool scene_indices_ok(const phos_scene_desc* d);  // phos_cuda.cu
}

extern "C" {

phos_bvh* phos_bvh_build(const phos_scene_desc* scene, int threads) {
  using namespace phos;
  if (!scene || !scene->num_meshes || !scene->vert_offset || !scene->vertices || !scene->face_offset || !scene->faces ||
      !scene_indices_ok(scene))  // a face with a vertex index outside its mesh would be read out of bounds right below
    return nullptr;
  const int nthreads = threads > 0 ? std::min(threads, 64) : worker_count();
  uint32_t grain = 1u << 15;  // ranges below this are one task (PHOS_BUILD_GRAIN: tests use a tiny one)
  if (const char* e = std::getenv("PHOS_BUILD_GRAIN")) grain = (uint32_t)std::max(8, std::atoi(e));
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<Tri> tris;
  gather_triangles(scene, tris);
  std::vector<Prim> prims(tris.size());
  parallel_chunks(tris.size(), 1u << 16, nthreads, [&](size_t b0, size_t b1, size_t) {
    for (size_t i = b0; i < b1; ++i) {
      Prim& p = prims[i];
      p.index = (uint32_t)i;
      p.bounds.grow(tris[i].a);
      p.bounds.grow(tris[i].b);
      p.bounds.grow(tris[i].c);
      p.centroid = {(p.bounds.hi.x + p.bounds.lo.x) / 2, (p.bounds.hi.y + p.bounds.lo.y) / 2,
                    (p.bounds.hi.z + p.bounds.lo.z) / 2};
    }
  });
  auto* out = new phos_bvh();
  Builder b{prims, tris};
  g_scan_threads = nthreads;
  Piece top;
  top.range = make_range(prims, 0, (uint32_t)prims.size());
  {
    TaskBag bag;
    ParallelBuild pb{b, bag, grain};
    bag.add([&] { pb.expand(&top); });
    bag.run(nthreads);
    pb.measure(&top);
    if (!top.leaf()) {
      out->nodes.resize(top.n_nodes);
      out->packets.resize(top.n_packets);
      std::vector<Piece*> tasks;
      pb.place(&top, 0, 0, out->nodes, out->packets, tasks);
      parallel_chunks(tasks.size(), 1, nthreads, [&](size_t t0, size_t t1, size_t) {
        for (size_t t = t0; t < t1; ++t) ParallelBuild::copy_task(tasks[t], out->nodes, out->packets);
      });
    }
  }
  g_scan_threads = 1;
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
