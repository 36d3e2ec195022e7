// repack.cpp — reference 8-wide BVH (288 B nodes / 384 B packets) -> GPU layout (80 B / 48 B).
//
// Input is mbvh_t::root / mbvh_t::triangles exactly as the reference builder publishes them
// (reference src/accel/bvh.cpp:42-47; node semantics src/accel/bvh/node.hpp:25-67; leaf = packets
// offset .. offset + ceil(num/8) - 1, src/kernels/cpu/stream_bvh_kernel.cpp:126-142).
//
// What is kept: every triangle, bit for bit (e0, e1, v0 as the reference packet holds them, ids, and
// its packet * 8 + lane position, the brute-force visiting order that breaks exact-t ties).
// What changes, because the result of a ray query does not depend on it:
//   * the reference's leaves (4.3 triangles on average) are cut into pieces of at most 2 triangles: on
//     the GPU a Möller–Trumbore test costs as much as two box tests and runs at far lower lane
//     occupancy than the converged node step, so letting the node test cull single triangles wins
//     (+17..27 % measured, profiles/r01_sweep_leaf_size.log);
//   * the inner topology is re-grouped.  The reference widens a node by re-splitting its SMALLEST
//     child (binned_sah_builder.hpp:198-213), which yields long chains of "7 small children + 1
//     huge child" (depth 20 at 10 M triangles, 5.4 of 8 slots used).  Here the pieces are
//     re-grouped top-down with a binned SAH over their boxes, always splitting the LARGEST range
//     until a node has 8 children: a balanced 8-wide tree (depth ~9), built level by level on all
//     host threads (the result does not depend on the thread count);
//   * child boxes are quantised outwards to 8 bits per plane on a per-node power-of-two grid — any
//     ray that meets the real box meets the quantised one;
//   * the half-empty 8-wide SoA packets (4.2 of 8 lanes used) become 48 B triangles, stored so that
//     a node's leaf triangles are contiguous; nodes are numbered breadth-first so a node's inner
//     children are contiguous (popcount addressing);
//   * children are placed in slots so that slot ^ ray-octant approximates front-to-back order;
//   * tuning knobs (environment): PHOS_REPACK_LEAF (piece size, 15 = keep the reference leaves; the
//     reference's uint8 count allows 255 and wraps beyond, node.hpp:19, so oversized leaves are always
//     cut), PHOS_REPACK_MERGE (fuse sibling pieces up to that many triangles), PHOS_THREADS.
#include <algorithm>
#include <functional>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include "host_parallel.hpp"
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

constexpr uint32_t kMaxLeaf = 15;      // 4-bit triangle count per slot
constexpr uint32_t kDefaultLeaf = 2;   // reference leaves are cut into pieces of at most this many triangles
constexpr int kBins = 16;

// one piece (<= leaf_limit triangles) of a reference leaf: the unit that is re-grouped
struct Prim {
  DBox box;
  double c[3];         // box centre
  uint32_t tri_begin;  // into the extracted triangle array
  uint32_t tri_count;
};

struct Range {
  uint32_t begin = 0, end = 0;  // into the prim index array
  DBox box;
  uint32_t tris = 0;
  uint32_t count() const { return end - begin; }
};

struct Bins {
  DBox box[3][kBins];
  uint32_t tris[3][kBins];
  DBox cbox;  // centroid bounds (first pass only)
  void clear() {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < kBins; ++b) {
        box[a][b] = DBox();
        tris[a][b] = 0;
      }
    cbox = DBox();
  }
  void merge(const Bins& o) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < kBins; ++b) {
        box[a][b].grow(o.box[a][b]);
        tris[a][b] += o.tris[a][b];
      }
    cbox.grow(o.cbox);
  }
};

struct Builder {
  const std::vector<Prim>& prims;
  std::vector<uint32_t>& idx;
  int threads;  // workers available INSIDE one split (1 while nodes are processed in parallel)

  static constexpr size_t kGrain = 1u << 16;

  Range make_range(uint32_t b, uint32_t e) const {
    Range r;
    r.begin = b;
    r.end = e;
    const size_t chunks = ((size_t)(e - b) + kGrain - 1) / kGrain;
    if (chunks <= 1) {
      for (uint32_t i = b; i < e; ++i) {
        r.box.grow(prims[idx[i]].box);
        r.tris += prims[idx[i]].tri_count;
      }
      return r;
    }
    std::vector<Range> part(chunks);
    parallel_chunks(e - b, kGrain, threads, [&](size_t cb, size_t ce, size_t c) {
      Range& q = part[c];
      for (size_t i = b + cb; i < b + ce; ++i) {
        q.box.grow(prims[idx[i]].box);
        q.tris += prims[idx[i]].tri_count;
      }
    });
    for (const Range& q : part) {
      r.box.grow(q.box);
      r.tris += q.tris;
    }
    return r;
  }

  // binned SAH split of a range of prims (cost weighted by triangle count); median split when the
  // centroids do not separate.  Always produces two non-empty halves for count >= 2.  Min / max and
  // integer sums only, so the result does not depend on how the range is chunked over threads.
  void split(const Range& r, Range& l, Range& rr) const {
    const size_t n = r.count();
    const size_t chunks = (n + kGrain - 1) / kGrain;
    Bins one;  // the common case (one chunk) needs no heap
    std::vector<Bins> many(chunks > 1 ? chunks : 0);
    Bins* part = chunks > 1 ? many.data() : &one;
    // pass 1: centroid bounds
    parallel_chunks(n, kGrain, threads, [&](size_t cb, size_t ce, size_t c) {
      DBox b;
      for (size_t i = r.begin + cb; i < r.begin + ce; ++i) {
        const Prim& p = prims[idx[i]];
        b.grow(p.c[0], p.c[1], p.c[2]);
      }
      part[c].cbox = b;
    });
    DBox cb;
    for (size_t c = 0; c < chunks; ++c) cb.grow(part[c].cbox);
    double k[3];
    bool live[3];
    for (int a = 0; a < 3; ++a) {
      const double ext = cb.hi[a] - cb.lo[a];
      live[a] = ext > 0.0;
      k[a] = live[a] ? kBins / ext : 0.0;
    }
    // pass 2: all three axes at once
    parallel_chunks(n, kGrain, threads, [&](size_t b0, size_t b1, size_t c) {
      Bins& B = part[c];
      B.clear();
      for (size_t i = r.begin + b0; i < r.begin + b1; ++i) {
        const Prim& p = prims[idx[i]];
        for (int a = 0; a < 3; ++a) {
          if (!live[a]) continue;
          const int b = std::min(kBins - 1, (int)((p.c[a] - cb.lo[a]) * k[a]));
          B.box[a][b].grow(p.box);
          B.tris[a][b] += p.tri_count;
        }
      }
    });
    for (size_t c = 1; c < chunks; ++c) part[0].merge(part[c]);
    const Bins& B = part[0];

    int best_axis = -1, best_bin = 0;
    double best_cost = DBL_MAX;
    DBox best_l, best_r;
    uint32_t best_nl = 0, best_nr = 0;
    for (int a = 0; a < 3; ++a) {
      if (!live[a]) continue;
      DBox suf[kBins];
      uint32_t sufn[kBins];
      DBox acc;
      uint32_t cnt = 0;
      for (int j = kBins - 1; j >= 0; --j) {
        acc.grow(B.box[a][j]);
        cnt += B.tris[a][j];
        suf[j] = acc;
        sufn[j] = cnt;
      }
      DBox left;
      uint32_t nl = 0;
      for (int j = 0; j < kBins - 1; ++j) {
        left.grow(B.box[a][j]);
        nl += B.tris[a][j];
        if (nl == 0 || sufn[j + 1] == 0) continue;
        const double cost = left.half_area() * nl + suf[j + 1].half_area() * sufn[j + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = a;
          best_bin = j;
          best_l = left;
          best_r = suf[j + 1];
          best_nl = nl;
          best_nr = sufn[j + 1];
        }
      }
    }
    uint32_t mid = r.begin;
    if (best_axis >= 0) {
      const double kk = k[best_axis], lo = cb.lo[best_axis];
      const int a = best_axis, bin = best_bin;
      uint32_t* m = std::partition(idx.data() + r.begin, idx.data() + r.end, [&](uint32_t i) {
        return std::min(kBins - 1, (int)((prims[i].c[a] - lo) * kk)) <= bin;
      });
      mid = (uint32_t)(m - idx.data());
    }
    if (mid == r.begin || mid == r.end) {  // degenerate: split by count along the widest centroid axis
      int a = 0;
      for (int q = 1; q < 3; ++q)
        if (cb.hi[q] - cb.lo[q] > cb.hi[a] - cb.lo[a]) a = q;
      mid = r.begin + r.count() / 2;
      std::nth_element(idx.begin() + r.begin, idx.begin() + mid, idx.begin() + r.end, [&](uint32_t x, uint32_t y) {
        return prims[x].c[a] < prims[y].c[a] || (prims[x].c[a] == prims[y].c[a] && x < y);
      });
      l = make_range(r.begin, mid);
      rr = make_range(mid, r.end);
      return;
    }
    // the halves' boxes and triangle counts are the unions of their bins: no further pass
    l.begin = r.begin;
    l.end = mid;
    l.box = best_l;
    l.tris = best_nl;
    rr.begin = mid;
    rr.end = r.end;
    rr.box = best_r;
    rr.tris = best_nr;
  }
};

// everything one node of the packed tree needs before its children / triangles are numbered
struct NodeOut {
  GNode g;
  Range slot[8];  // child range per slot (count() == 0: empty slot)
  bool ok = true;
};

}  // namespace

bool repack_accel(const RefNode* nodes, uint32_t n_nodes, const RefPacket* packets, uint32_t n_packets, PackedAccel& out,
                  std::string& err) {
  out = PackedAccel();
  uint32_t merge_limit = 0;  // fuse sibling pieces up to this many triangles (0: never)
  if (const char* e = std::getenv("PHOS_REPACK_MERGE")) merge_limit = std::min<uint32_t>(kMaxLeaf, (uint32_t)std::atoi(e));
  // Leaf size.  A Möller–Trumbore test costs about as many instructions as testing two child boxes and
  // runs at far lower lane occupancy than the converged node step, so small leaves win: measured on the
  // B200 (profiles/r01_sweep_leaf_size.log) 2 triangles per leaf is +17..27 % over keeping the reference's
  // leaves (4.3 triangles on average).  PHOS_REPACK_LEAF overrides (15 = keep the reference leaves).
  uint32_t leaf_limit = kDefaultLeaf;
  if (const char* e = std::getenv("PHOS_REPACK_LEAF")) leaf_limit = std::max<uint32_t>(1u, std::min<uint32_t>(kMaxLeaf, (uint32_t)std::atoi(e)));
  const int threads = worker_count();
  const bool verbose = std::getenv("PHOS_REPACK_VERBOSE") != nullptr;
  // how a reference leaf is cut into 2-triangle pieces: 1 (default) = the pairing with the smallest total box area,
  // 0 = chop the run sorted along the leaf's longest axis (the first version)
  const bool pair_by_area = std::getenv("PHOS_REPACK_PAIR") == nullptr || std::atoi(std::getenv("PHOS_REPACK_PAIR")) != 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const auto t_start = now();
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

  // ---- A. walk the reference tree: validate it, list its leaves (depth-first, slot 7 first) -------------
  struct RefLeaf {
    uint32_t packet, npk;    // packets[packet .. packet + npk)
    uint32_t tri_first, cnt; // into the extracted triangle array
    uint32_t prim_first;
  };
  std::vector<RefLeaf> leaves;
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
        leaves.push_back({rn.offset[i], npk, 0, 0, 0});
      }
      if (!any) {
        err = "inner node without children";
        return false;
      }
    }
  }
  // triangle / piece numbering of every leaf (prefix sums over the packet counts)
  uint64_t n_tris = 0, n_prims = 0;
  for (RefLeaf& lf : leaves) {
    uint32_t cnt = 0;
    for (uint32_t p = lf.packet; p < lf.packet + lf.npk; ++p) {
      if (packets[p].num > 8) {
        err = "packet with more than 8 triangles";
        return false;
      }
      cnt += packets[p].num;
    }
    lf.tri_first = (uint32_t)n_tris;
    lf.cnt = cnt;
    lf.prim_first = (uint32_t)n_prims;
    n_tris += cnt;
    n_prims += (cnt + leaf_limit - 1) / leaf_limit;
  }
  if (n_tris == 0 || n_prims == 0) {
    err = "acceleration structure holds no triangles";
    return false;
  }
  if (n_tris > 0xfffffff0ull) {
    err = "more than 2^32 triangles";
    return false;
  }
  // pull out the triangles and cut every leaf into pieces, leaves in parallel
  std::vector<GTri> tris(n_tris);
  std::vector<Prim> prims(n_prims);
  parallel_chunks(leaves.size(), 4096, threads, [&](size_t b, size_t e, size_t) {
    std::vector<std::pair<double, uint32_t>> key;
    std::vector<GTri> tmp;
    std::vector<DBox> tb;
    for (size_t li = b; li < e; ++li) {
      const RefLeaf& lf = leaves[li];
      uint32_t k = lf.tri_first;
      for (uint32_t p = lf.packet; p < lf.packet + lf.npk; ++p) {
        const RefPacket& pk = packets[p];
        for (uint32_t j = 0; j < pk.num; ++j) {
          GTri& t = tris[k++];
          t.v0x = pk.v0x[j]; t.v0y = pk.v0y[j]; t.v0z = pk.v0z[j];
          t.e0x = pk.e0x[j]; t.e0y = pk.e0y[j]; t.e0z = pk.e0z[j];
          t.e1x = pk.e1x[j]; t.e1y = pk.e1y[j]; t.e1z = pk.e1z[j];
          t.meshid = pk.meshid[j];
          t.faceid = pk.faceid[j];
          t.order = p * 8 + j;
        }
      }
      const uint32_t cnt = lf.cnt, first = lf.tri_first;
      tb.resize(cnt);
      for (uint32_t i = 0; i < cnt; ++i) tb[i] = tri_box(tris[first + i]);
      if (cnt > leaf_limit) {  // cut along the longest axis of the leaf: order the triangles by box centre
        DBox cb;
        for (uint32_t i = 0; i < cnt; ++i) cb.grow(tb[i]);
        int ax = 0;
        for (int a = 1; a < 3; ++a)
          if (cb.hi[a] - cb.lo[a] > cb.hi[ax] - cb.lo[ax]) ax = a;
        key.resize(cnt);
        for (uint32_t i = 0; i < cnt; ++i) key[i] = {tb[i].lo[ax] + tb[i].hi[ax], i};
        std::sort(key.begin(), key.end());
        if (pair_by_area && leaf_limit == 2 && cnt <= 10) {
          // instead of chopping the sorted run into pairs, take the perfect matching (one triangle left alone when cnt
          // is odd) with the smallest total surface area of the pair boxes: 5-7 % fewer triangle tests per ray, a
          // little fewer node tests (profiles/r01_tree_pairing.log), no measurable re-pack time (<= 105 matchings of
          // <= 8 triangles per leaf; PHOS_REPACK_PAIR=0 restores the chop)
          auto half_area = [](const DBox& b) {
            const double x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
            return x * y + y * z + z * x;
          };
          double pa[10][10];
          for (uint32_t i = 0; i < cnt; ++i)
            for (uint32_t j = i + 1; j < cnt; ++j) {
              DBox u = tb[key[i].second];
              u.grow(tb[key[j].second]);
              pa[i][j] = half_area(u);
            }
          double best = 1e300;
          int best_pairs[10], cur_pairs[10];
          std::function<void(uint32_t, double, int, bool)> rec = [&](uint32_t used, double cost, int np, bool single_used) {
            if (cost >= best) return;
            uint32_t i = 0;
            while (i < cnt && (used >> i & 1u)) ++i;
            if (i == cnt) {
              best = cost;
              for (int q = 0; q < np; ++q) best_pairs[q] = cur_pairs[q];
              return;
            }
            if ((cnt & 1u) && !single_used) {  // i stays alone
              cur_pairs[np] = (int)(i | (0xffu << 8));
              rec(used | (1u << i), cost + half_area(tb[key[i].second]), np + 1, true);
            }
            for (uint32_t j = i + 1; j < cnt; ++j)
              if (!(used >> j & 1u)) {
                cur_pairs[np] = (int)(i | (j << 8));
                rec(used | (1u << i) | (1u << j), cost + pa[i][j], np + 1, single_used);
              }
          };
          rec(0u, 0.0, 0, false);
          const int np = (int)((cnt + 1) / 2);
          std::vector<std::pair<double, uint32_t>> k2;
          int single = -1;
          for (int q = 0; q < np; ++q) {
            const int a = best_pairs[q] & 0xff, b = best_pairs[q] >> 8;
            if (b == 0xff) { single = a; continue; }
            k2.push_back(key[a]);
            k2.push_back(key[b]);
          }
          if (single >= 0) k2.push_back(key[single]);  // the chop below takes pairs first, the odd one last
          key.swap(k2);
        }
        tmp.assign(tris.begin() + first, tris.begin() + first + cnt);
        std::vector<DBox> tb2(cnt);
        for (uint32_t i = 0; i < cnt; ++i) {
          tris[first + i] = tmp[key[i].second];
          tb2[i] = tb[key[i].second];
        }
        tb.swap(tb2);
      }
      uint32_t pi = lf.prim_first;
      for (uint32_t o = 0; o < cnt;) {
        const uint32_t pieces = (cnt - o + leaf_limit - 1) / leaf_limit;
        const uint32_t take = (cnt - o + pieces - 1) / pieces;
        Prim& p = prims[pi++];
        p.tri_begin = first + o;
        p.tri_count = take;
        p.box = DBox();
        for (uint32_t q = 0; q < take; ++q) p.box.grow(tb[o + q]);
        for (int a = 0; a < 3; ++a) p.c[a] = 0.5 * (p.box.lo[a] + p.box.hi[a]);
        o += take;
      }
    }
  });
  const auto t_extract = now();

  // ---- B. re-group the pieces into a balanced 8-wide tree, one level at a time ----------------------------
  // Nodes are numbered breadth first, a node's inner children and a node's leaf triangles are
  // contiguous.  Every node of a level only touches its own slice of the index array, so a level's nodes
  // are processed in parallel; while a level has fewer nodes than workers (the top of the tree, where the
  // ranges are huge) the nodes are taken one by one and the binning inside each split is parallel instead.
  std::vector<uint32_t> idx(prims.size());
  std::iota(idx.begin(), idx.end(), 0u);
  auto is_leaf = [&](const Range& r) { return r.count() == 1 || r.tris <= merge_limit; };

  auto process = [&](const Range& range, const Builder& B, NodeOut& o) {
    // children: split the largest splittable range until there are 8
    Range child[8];
    int nc = 1;
    child[0] = range;
    if (range.count() > 1 && range.count() <= 8 && merge_limit == 0) {
      // splitting would end with every piece as a child of its own: skip the SAH passes
      nc = (int)range.count();
      for (int i = 0; i < nc; ++i) {
        const Prim& p = prims[idx[range.begin + i]];
        child[i].begin = range.begin + i;
        child[i].end = range.begin + i + 1;
        child[i].box = p.box;
        child[i].tris = p.tri_count;
      }
    } else if (range.count() > 1) {
      while (nc < 8) {
        int pick = -1;
        double best = -1.0;
        for (int i = 0; i < nc; ++i) {
          if (is_leaf(child[i]) && !(nc == 1)) continue;
          if (child[i].count() < 2) continue;
          // largest area first.  Tried and dropped (tools/tree_quality.py): weighting by triangle count
          // (+1.6..10 % node tests), sparing ranges that already fit one node (no change), and clustering
          // small ranges bottom-up into full nodes (half the nodes, but only -3.7 % node tests at +5 %
          // triangle tests)
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
    for (int i = 0; i < nc; ++i) o.slot[slot_of[i]] = child[i];

    // ---- quantisation grid -----------------------------------------------------------------------
    GNode& g = o.g;
    memset(&g, 0, sizeof(g));
    // Plane q of an axis sits at og + (2^15 + q) * 2^e: the kernel decodes a byte with one PRMT into
    // the float m = 1 + q * 2^-15 (bits 0x3F800000 | q << 8) and evaluates t = fma(m, 2^(e+15) / dir,
    // (og - org) / dir).  `og` is the stored fp32 origin, 2^15 cells below the grid.
    // margin, in grid cells, swallows the fp32 error of that formulation for rays that start inside or
    // near the node: the rounding of (og - org) / dir is worth <= ~2^-8 cells (DESIGN.md "conservative
    // traversal"); rays far from the node are covered by the relative slack in the kernel.  The grid
    // starts 2 margins below the true minimum so the lowest plane keeps its margin (q >= 0).
    const double margin = 1.0 / 64.0;
    double og[3], scale[3];
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
      og[a] = (double)of + 32768.0 * scale[a];
      (&g.ox)[a] = of;
      (&g.ex)[a] = (uint8_t)(e + 127);
    }

    uint8_t* qlo[3] = {g.qlox, g.qloy, g.qloz};
    uint8_t* qhi[3] = {g.qhix, g.qhiy, g.qhiz};
    for (int s = 0; s < 8; ++s) {
      for (int a = 0; a < 3; ++a) {
        qlo[a][s] = 255;  // empty: lo > hi
        qhi[a][s] = 0;
      }
      const Range& c = o.slot[s];
      if (c.count() == 0) continue;
      for (int a = 0; a < 3; ++a) {
        const double ql = std::floor((c.box.lo[a] - og[a]) / scale[a] - margin);
        const double qh = std::ceil((c.box.hi[a] - og[a]) / scale[a] + margin);
        if (ql < 0.0 || qh > 255.0 || og[a] + (ql + margin) * scale[a] > c.box.lo[a] ||
            og[a] + (qh - margin) * scale[a] < c.box.hi[a]) {
          o.ok = false;
          return;
        }
        qlo[a][s] = (uint8_t)ql;
        qhi[a][s] = (uint8_t)qh;
      }
      if ((c.count() == 1 || (nc > 1 && is_leaf(c))) && c.tris <= kMaxLeaf)
        g.counts |= c.tris << (4 * s);
      else
        g.imask |= (uint8_t)(1u << s);
    }
  };

  std::vector<Range> level, next_level;
  {
    Builder top{prims, idx, threads};
    level.push_back(top.make_range(0, (uint32_t)prims.size()));
  }
  out.nodes.reserve(prims.size() / 3 + 16);
  out.tris.resize(tris.size());
  uint32_t tri_cursor = 0, depth = 0;
  std::vector<NodeOut> outs;
  std::vector<uint32_t> tri_base;
  constexpr size_t kBatch = 1u << 15;  // nodes in flight (bounds the scratch memory of a wide level)
  while (!level.empty()) {
    const size_t ln = level.size();
    const auto t_level = now();
    const uint32_t next_first = (uint32_t)(out.nodes.size() + ln);  // first node of the next level
    const bool wide = ln >= (size_t)threads * 4;
    next_level.clear();
    for (size_t b0 = 0; b0 < ln; b0 += kBatch) {
      const size_t bn = std::min(kBatch, ln - b0);
      outs.assign(bn, NodeOut());
      tri_base.resize(bn);
      if (wide) {
        const Builder B{prims, idx, 1};
        parallel_chunks(bn, 16, threads, [&](size_t b, size_t e, size_t) {
          for (size_t j = b; j < e; ++j) process(level[b0 + j], B, outs[j]);
        });
      } else {
        const Builder B{prims, idx, threads};
        for (size_t j = 0; j < bn; ++j) process(level[b0 + j], B, outs[j]);
      }
      // number the children and the leaf triangles in node order
      for (size_t j = 0; j < bn; ++j) {
        NodeOut& o = outs[j];
        if (!o.ok) {
          err = "internal: quantised box does not contain the child box";
          return false;
        }
        o.g.child_base = next_first + (uint32_t)next_level.size();
        o.g.tri_base = tri_base[j] = tri_cursor;
        for (int s = 0; s < 8; ++s) {
          if (o.slot[s].count() == 0) continue;
          if (o.g.imask & (1u << s)) {
            next_level.push_back(o.slot[s]);
          } else {
            tri_cursor += o.slot[s].tris;
            out.max_leaf_tris = std::max(out.max_leaf_tris, o.slot[s].tris);
          }
        }
        out.nodes.push_back(o.g);
      }
      parallel_chunks(bn, 64, threads, [&](size_t b, size_t e, size_t) {
        for (size_t j = b; j < e; ++j) {
          const NodeOut& o = outs[j];
          uint32_t w = tri_base[j];
          for (int s = 0; s < 8; ++s) {
            if (o.slot[s].count() == 0 || (o.g.imask & (1u << s))) continue;
            for (uint32_t i = o.slot[s].begin; i < o.slot[s].end; ++i) {
              const Prim& p = prims[idx[i]];
              std::copy(tris.begin() + p.tri_begin, tris.begin() + p.tri_begin + p.tri_count, out.tris.begin() + w);
              w += p.tri_count;
            }
          }
        }
      });
    }
    if (verbose) fprintf(stderr, "[phos repack]   level %u: %zu nodes, %.2f s\n", depth, ln, secs(t_level, now()));
    out.max_depth = depth++;
    level.swap(next_level);
  }
  if (tri_cursor != tris.size()) {
    err = "internal: triangle count changed during re-pack";
    return false;
  }
  if (verbose)
    fprintf(stderr, "[phos repack] %u triangles -> %zu nodes, depth %u, leaf <= %u; extract %.2f s, re-group %.2f s, %d threads\n",
            (uint32_t)tris.size(), out.nodes.size(), out.max_depth, out.max_leaf_tris, secs(t_start, t_extract),
            secs(t_extract, now()), threads);
  return true;
}

}  // namespace phos
