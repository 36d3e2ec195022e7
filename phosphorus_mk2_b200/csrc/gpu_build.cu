// gpu_build.cu — the packed acceleration structure built ON the device.
//
// phos_cuda_upload_accel keeps the reference's CPU build (bvh::from, src/accel/bvh/binned_sah_builder.hpp:215-281)
// and re-packs its arrays; that costs seconds per 10 M triangles on the host.  phos_cuda_build_accel goes from the
// scene's triangles straight to the packed layout of phos_internal.hpp (GNode 80 B / GTri 48 B) in a few tens of
// milliseconds, for interactive scene edits and 30 M-triangle scenes where the host build dominates:
//   1. gather:   every triangle in the reference's numbering (mesh -> face set -> face, src/scene.cpp:58-62) becomes a
//                GTri with exactly the fp32 v0, e0 = b - a, e1 = c - a the reference packet would hold
//                (src/accel/triangle.hpp:48-50) + an outward-rounded box of what Moeller-Trumbore sees;
//   2. order:    63-bit Morton code of the box centre, radix sort (cub::DeviceRadixSort — a library primitive, used
//                for plumbing only);
//   3. topology: binary radix tree over the sorted codes (Karras 2012), boxes and triangle counts bottom-up;
//   4. collapse: level by level, one thread per 8-wide node: starting from a binary node, the child with the largest
//                box that still holds more than `leaf` triangles is replaced by its two children until there are 8;
//                children are placed in octant slots, quantised outwards on the node's power-of-two grid (the same
//                arithmetic and safety margins as repack.cpp) and numbered so that a node's inner children and a
//                node's leaf triangles are contiguous.
// The result answers every query exactly like the re-packed reference tree (same triangles, same Moeller-Trumbore
// arithmetic, conservative boxes); only exact ties in t between two triangles may resolve differently, because the
// tie-break key is the triangle's scene order here and its position in the reference's packet array there.
// A Morton-ordered tree is a worse SAH tree than the reference's: expect ~20-35 % more node tests per ray
// (profiles/): this entry point trades traversal speed for build time.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"
#include "phos_internal.hpp"
#include "trace_ray.cuh"

namespace phos {
namespace {

struct BuildScene {
  const float* verts;
  const uint32_t* faces;
  const uint32_t* vert_offset;
  const uint32_t* face_offset;
  const uint32_t* tri_meshmat;  // meshid | matid << 16 per triangle, scene order
  const uint32_t* tri_face;     // mesh-local face index per triangle
};

struct FBox {
  float lo[3], hi[3];
};

__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// 1. gather ---------------------------------------------------------------------------------------------------
__global__ void gather_kernel(const BuildScene S, uint32_t n, GTri* __restrict__ tris, FBox* __restrict__ boxes,
                              int* __restrict__ scene_bounds /* 6 ordered ints: lo xyz, hi xyz */) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t mm = S.tri_meshmat[t], mesh = mm & 0xffffu, face = S.tri_face[t];
  const uint32_t* f = S.faces + 3 * ((size_t)S.face_offset[mesh] + face);
  const size_t vo = S.vert_offset[mesh];
  const float* pa = S.verts + 3 * (vo + f[0]);
  const float* pb = S.verts + 3 * (vo + f[1]);
  const float* pc = S.verts + 3 * (vo + f[2]);
  GTri g;
  g.v0x = pa[0]; g.v0y = pa[1]; g.v0z = pa[2];
  g.e0x = __fsub_rn(pb[0], pa[0]); g.e0y = __fsub_rn(pb[1], pa[1]); g.e0z = __fsub_rn(pb[2], pa[2]);
  g.e1x = __fsub_rn(pc[0], pa[0]); g.e1y = __fsub_rn(pc[1], pa[1]); g.e1z = __fsub_rn(pc[2], pa[2]);
  g.meshid = mm;
  g.faceid = 3u * face;
  g.order = t;
  tris[t] = g;
  // box of v0, v0 + e0, v0 + e1 (what Moeller-Trumbore sees; e0 = fl(b - a) does not reproduce b exactly), padded
  // by one fp32 ulp of the largest coordinate and rounded outwards to fp32 — repack.cpp tri_box
  const double v0[3] = {g.v0x, g.v0y, g.v0z}, e0[3] = {g.e0x, g.e0y, g.e0z}, e1[3] = {g.e1x, g.e1y, g.e1z};
  FBox b;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double lo = fmin(v0[a], fmin(v0[a] + e0[a], v0[a] + e1[a]));
    double hi = fmax(v0[a], fmax(v0[a] + e0[a], v0[a] + e1[a]));
    const double m = fmax(fabs(lo), fabs(hi)) * 1.2e-7 + 1e-37;
    lo -= m;
    hi += m;
    b.lo[a] = __double2float_rd(lo);
    b.hi[a] = __double2float_ru(hi);
    atomicMin(scene_bounds + a, float_to_ordered(b.lo[a]));
    atomicMax(scene_bounds + 3 + a, float_to_ordered(b.hi[a]));
  }
  boxes[t] = b;
}

// 2. Morton codes ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__global__ void morton_kernel(const FBox* __restrict__ boxes, uint32_t n, const int* __restrict__ scene_bounds,
                              unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned long long code = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double lo = ordered_to_float(scene_bounds[a]), hi = ordered_to_float(scene_bounds[3 + a]);
    const double c = 0.5 * ((double)boxes[t].lo[a] + (double)boxes[t].hi[a]);
    const double ext = hi - lo;
    double u = ext > 0.0 ? (c - lo) / ext : 0.0;
    u = fmin(fmax(u, 0.0), 1.0);
    const unsigned long long q = (unsigned long long)fmin(u * 2097152.0, 2097151.0);
    code |= spread21(q) << a;
  }
  keys[t] = code;
  vals[t] = t;
}

// 3. binary radix tree (Karras 2012) ----------------------------------------------------------------------------
constexpr uint32_t kLeafBit = 0x80000000u;

__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);  // equal codes: fall back to the index
  return __clzll((long long)(a ^ b));
}

__global__ void radix_tree_kernel(const unsigned long long* __restrict__ keys, int n, uint32_t* __restrict__ left,
                                  uint32_t* __restrict__ right, uint32_t* __restrict__ parent_inner,
                                  uint32_t* __restrict__ parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const uint32_t lc = lo == gamma ? (kLeafBit | (uint32_t)gamma) : (uint32_t)gamma;
  const uint32_t rc = hi == gamma + 1 ? (kLeafBit | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
  left[i] = lc;
  right[i] = rc;
  if (lc & kLeafBit) parent_leaf[gamma] = (uint32_t)i;
  else parent_inner[gamma] = (uint32_t)i;
  if (rc & kLeafBit) parent_leaf[gamma + 1] = (uint32_t)i;
  else parent_inner[gamma + 1] = (uint32_t)i;
  if (i == 0) parent_inner[0] = 0xffffffffu;
}

__global__ void refit_kernel(int n, const uint32_t* __restrict__ sorted, const FBox* __restrict__ tri_boxes,
                             const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                             const uint32_t* __restrict__ parent_inner, const uint32_t* __restrict__ parent_leaf,
                             FBox* __restrict__ node_boxes, uint32_t* __restrict__ node_count, uint32_t* __restrict__ visits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t cur = parent_leaf[i];
  while (cur != 0xffffffffu) {
    __threadfence();
    if (atomicAdd(&visits[cur], 1u) == 0u) return;  // the sibling sub-tree is not done yet: it will continue
    const uint32_t lc = left[cur], rc = right[cur];
    // boxes / counts written by other threads moments ago: read them from L2 (a stale L1 line could hold a neighbour)
    auto load_box = [&](uint32_t ref) -> FBox {
      if (ref & kLeafBit) return tri_boxes[sorted[ref & ~kLeafBit]];
      FBox r;
      const float* p = reinterpret_cast<const float*>(node_boxes + ref);
      for (int k = 0; k < 3; ++k) {
        r.lo[k] = __ldcg(p + k);
        r.hi[k] = __ldcg(p + 3 + k);
      }
      return r;
    };
    const FBox a = load_box(lc), b = load_box(rc);
    FBox u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      u.lo[k] = fminf(a.lo[k], b.lo[k]);
      u.hi[k] = fmaxf(a.hi[k], b.hi[k]);
    }
    node_boxes[cur] = u;
    node_count[cur] = ((lc & kLeafBit) ? 1u : __ldcg(node_count + lc)) + ((rc & kLeafBit) ? 1u : __ldcg(node_count + rc));
    cur = parent_inner[cur];
  }
}

// 4. collapse to 8-wide, quantise, emit ------------------------------------------------------------------------
struct CollapseArgs {
  const uint32_t* items;  // binary node of every wide node of this level
  uint32_t count;
  uint32_t level_first, next_first;
  uint32_t* next_items;
  uint32_t* next_count;
  uint32_t* tri_cursor;
  uint32_t* error;
  uint32_t leaf_limit;
  const uint32_t* sorted;
  const GTri* tris_in;
  const FBox* tri_boxes;
  const uint32_t* left;
  const uint32_t* right;
  const FBox* node_boxes;
  const uint32_t* node_count;
  GNode* nodes;
  GTri* tris_out;
};

__device__ __forceinline__ double half_area(const FBox& b) {
  const double dx = (double)b.hi[0] - b.lo[0], dy = (double)b.hi[1] - b.lo[1], dz = (double)b.hi[2] - b.lo[2];
  return dx * dy + dx * dz + dy * dz;
}

__global__ void collapse_kernel(const CollapseArgs A) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= A.count) return;
  auto box_of = [&](uint32_t ref) -> FBox { return (ref & kLeafBit) ? A.tri_boxes[A.sorted[ref & ~kLeafBit]] : A.node_boxes[ref]; };
  auto count_of = [&](uint32_t ref) -> uint32_t { return (ref & kLeafBit) ? 1u : A.node_count[ref]; };

  const uint32_t b = A.items[w];
  uint32_t child[8];
  int nc = 2;
  child[0] = A.left[b];
  child[1] = A.right[b];
  while (nc < 8) {  // split the largest child that is still more than one leaf piece
    int pick = -1;
    double best = -1.0;
    for (int i = 0; i < nc; ++i) {
      if ((child[i] & kLeafBit) || count_of(child[i]) <= A.leaf_limit) continue;
      const double a = half_area(A.node_boxes[child[i]]);
      if (a > best) {
        best = a;
        pick = i;
      }
    }
    if (pick < 0) break;
    const uint32_t c = child[pick];
    child[pick] = A.left[c];
    child[nc++] = A.right[c];
  }
  FBox cb[8];
  FBox nb;
  for (int k = 0; k < 3; ++k) {
    nb.lo[k] = 3.402823466e+38f;
    nb.hi[k] = -3.402823466e+38f;
  }
  for (int i = 0; i < nc; ++i) {
    cb[i] = box_of(child[i]);
    for (int k = 0; k < 3; ++k) {
      nb.lo[k] = fminf(nb.lo[k], cb[i].lo[k]);
      nb.hi[k] = fmaxf(nb.hi[k], cb[i].hi[k]);
    }
  }
  // slot assignment: greedy max of dot(child centre - node centre, slot sign vector) (repack.cpp)
  int slot_of[8];
  {
    bool child_done[8] = {false, false, false, false, false, false, false, false};
    bool slot_used[8] = {false, false, false, false, false, false, false, false};
    for (int round = 0; round < nc; ++round) {
      double best = -1.7976931348623157e308;
      int bi = -1, bs = -1;
      for (int i = 0; i < nc; ++i) {
        if (child_done[i]) continue;
        double rel[3];
        for (int a = 0; a < 3; ++a)
          rel[a] = 0.5 * ((double)cb[i].lo[a] + cb[i].hi[a]) - 0.5 * ((double)nb.lo[a] + nb.hi[a]);
        for (int s = 0; s < 8; ++s) {
          if (slot_used[s]) continue;
          double v = 0.0;
          for (int a = 0; a < 3; ++a) v += ((s >> a) & 1) ? rel[a] : -rel[a];
          if (v > best) {
            best = v;
            bi = i;
            bs = s;
          }
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

  // quantisation grid: plane q of an axis sits at og + (2^15 + q) * 2^e (see repack.cpp for the margins)
  GNode g;
  memset(&g, 0, sizeof(g));
  const double margin = 1.0 / 64.0;
  double og[3], scale[3];
  for (int a = 0; a < 3; ++a) {
    const double ext0 = (double)nb.hi[a] - nb.lo[a];
    int e = ext0 > 0.0 ? (int)ceil(log2(ext0 / 252.0)) : -100;
    e = max(-100, min(100, e));
    float of = 0.0f;
    for (;; ++e) {
      const double sc = ldexp(1.0, e);
      const double want = (double)nb.lo[a] - 2.0 * margin * sc - 32768.0 * sc;
      of = __double2float_rd(want);  // fp32 origin at or below the wanted one
      if (((double)nb.hi[a] - ((double)of + 32768.0 * sc)) / sc + 2.0 * margin <= 254.0 || e >= 100) break;
    }
    scale[a] = ldexp(1.0, e);
    og[a] = (double)of + 32768.0 * scale[a];
    (&g.ox)[a] = of;
    (&g.ex)[a] = (uint8_t)(e + 127);
  }
  uint8_t* qlo[3] = {g.qlox, g.qloy, g.qloz};
  uint8_t* qhi[3] = {g.qhix, g.qhiy, g.qhiz};
  uint32_t n_inner = 0, n_leaf_tris = 0;
  for (int s = 0; s < 8; ++s) {
    for (int a = 0; a < 3; ++a) {
      qlo[a][s] = 255;  // empty: lo > hi
      qhi[a][s] = 0;
    }
    const int ci = child_in_slot[s];
    if (ci < 0) continue;
    for (int a = 0; a < 3; ++a) {
      const double ql = floor(((double)cb[ci].lo[a] - og[a]) / scale[a] - margin);
      const double qh = ceil(((double)cb[ci].hi[a] - og[a]) / scale[a] + margin);
      if (ql < 0.0 || qh > 255.0 || og[a] + (ql + margin) * scale[a] > (double)cb[ci].lo[a] ||
          og[a] + (qh - margin) * scale[a] < (double)cb[ci].hi[a]) {
        atomicExch(A.error, 1u);
        return;
      }
      qlo[a][s] = (uint8_t)ql;
      qhi[a][s] = (uint8_t)qh;
    }
    const uint32_t cnt = count_of(child[ci]);
    if ((child[ci] & kLeafBit) || cnt <= A.leaf_limit) {
      g.counts |= cnt << (4 * s);
      n_leaf_tris += cnt;
    } else {
      g.imask |= (uint8_t)(1u << s);
      ++n_inner;
    }
  }
  // number the inner children (contiguous, slot order) and the leaf triangles (contiguous, slot order)
  const uint32_t ibase = n_inner ? atomicAdd(A.next_count, n_inner) : 0u;
  const uint32_t tbase = n_leaf_tris ? atomicAdd(A.tri_cursor, n_leaf_tris) : 0u;
  g.child_base = A.next_first + ibase;
  g.tri_base = tbase;
  uint32_t iw = ibase, tw = tbase;
  for (int s = 0; s < 8; ++s) {
    const int ci = child_in_slot[s];
    if (ci < 0) continue;
    if (g.imask & (1u << s)) {
      A.next_items[iw++] = child[ci];
      continue;
    }
    // the piece's triangles: the leaves of this (tiny) binary sub-tree, left to right
    uint32_t stack[16];
    int sp = 0;
    stack[sp++] = child[ci];
    while (sp) {
      const uint32_t r = stack[--sp];
      if (r & kLeafBit) {
        A.tris_out[tw++] = A.tris_in[A.sorted[r & ~kLeafBit]];
      } else if (sp + 2 <= 16) {
        stack[sp++] = A.right[r];
        stack[sp++] = A.left[r];
      } else {
        atomicExch(A.error, 2u);
        return;
      }
    }
  }
  A.nodes[A.level_first + w] = g;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  bool alloc(phos_ctx* ctx, size_t n, const char* what) { return cuda_ok(ctx, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)), what); }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

}  // namespace
}  // namespace phos

using namespace phos;

extern "C" int phos_cuda_build_accel(phos_ctx* ctx, const phos_scene_desc* d) {
  if (!ctx || !d) return PHOS_ERR_INVALID;
  const uint32_t nm = d->num_meshes;
  if (nm == 0 || !d->vert_offset || !d->vertices || !d->face_offset || !d->faces || !d->set_offset || !d->set_material ||
      !d->set_face_offset || !d->set_faces)
    return fail(ctx, PHOS_ERR_INVALID, "incomplete scene description");
  if (!scene_indices_ok(d)) return fail(ctx, PHOS_ERR_INVALID, "face with a vertex index outside its mesh");
  cudaSetDevice(ctx->device);
  const auto t0 = std::chrono::steady_clock::now();
  // triangles in the reference's numbering: mesh -> face set -> face (src/scene.cpp:58-62, src/mesh.cpp:118-128)
  std::vector<uint32_t> meshmat, face;
  for (uint32_t m = 0; m < nm; ++m)
    for (uint32_t s = d->set_offset[m]; s < d->set_offset[m + 1]; ++s) {
      const uint32_t mat = d->set_material[s];
      for (uint32_t j = d->set_face_offset[s]; j < d->set_face_offset[s + 1]; ++j) {
        if (d->set_faces[j] >= d->face_offset[m + 1] - d->face_offset[m]) return fail(ctx, PHOS_ERR_INVALID, "face set index out of range");
        meshmat.push_back(m | (mat << 16));
        face.push_back(d->set_faces[j]);
      }
    }
  const size_t n64 = meshmat.size();
  if (n64 < 2) return fail(ctx, PHOS_ERR_ACCEL, "empty acceleration structure (fewer than 2 triangles)");
  if (n64 >= 0x7fffffffull) return fail(ctx, PHOS_ERR_INVALID, "more than 2^31 triangles");
  const uint32_t n = (uint32_t)n64;
  uint32_t leaf_limit = 2;
  if (const char* e = std::getenv("PHOS_REPACK_LEAF")) leaf_limit = std::max(1, std::min(15, std::atoi(e)));
  const size_t nv = d->vert_offset[nm], nf = d->face_offset[nm];

  DevBuf<float> verts;
  DevBuf<uint32_t> faces, voff, foff, dmm, dface, vals_in, vals, left, right, par_i, par_l, ncount, visits, items_a, items_b, counters;
  DevBuf<int> bounds;
  DevBuf<GTri> tris_in;
  DevBuf<FBox> tboxes, nboxes;
  DevBuf<unsigned long long> keys_in, keys;
  DevBuf<unsigned char> sort_tmp;
  bool ok = verts.alloc(ctx, 3 * nv, "cudaMalloc(build)") && faces.alloc(ctx, 3 * nf, "cudaMalloc(build)") &&
            voff.alloc(ctx, nm + 1, "cudaMalloc(build)") && foff.alloc(ctx, nm + 1, "cudaMalloc(build)") &&
            dmm.alloc(ctx, n, "cudaMalloc(build)") && dface.alloc(ctx, n, "cudaMalloc(build)") &&
            tris_in.alloc(ctx, n, "cudaMalloc(build)") && tboxes.alloc(ctx, n, "cudaMalloc(build)") &&
            bounds.alloc(ctx, 6, "cudaMalloc(build)") && keys_in.alloc(ctx, n, "cudaMalloc(build)") &&
            keys.alloc(ctx, n, "cudaMalloc(build)") && vals_in.alloc(ctx, n, "cudaMalloc(build)") &&
            vals.alloc(ctx, n, "cudaMalloc(build)") && left.alloc(ctx, n, "cudaMalloc(build)") &&
            right.alloc(ctx, n, "cudaMalloc(build)") && par_i.alloc(ctx, n, "cudaMalloc(build)") &&
            par_l.alloc(ctx, n, "cudaMalloc(build)") && nboxes.alloc(ctx, n, "cudaMalloc(build)") &&
            ncount.alloc(ctx, n, "cudaMalloc(build)") && visits.alloc(ctx, n, "cudaMalloc(build)") &&
            items_a.alloc(ctx, n, "cudaMalloc(build)") && items_b.alloc(ctx, n, "cudaMalloc(build)") &&
            counters.alloc(ctx, 4, "cudaMalloc(build)");
  if (!ok) return PHOS_ERR_CUDA;
  cudaStream_t st = ctx->stream;
  auto up = [&](void* dst, const void* src, size_t bytes) { return cuda_ok(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st), "upload (build)"); };
  ok = up(verts.p, d->vertices, 12 * nv) && up(faces.p, d->faces, 12 * nf) && up(voff.p, d->vert_offset, 4 * (nm + 1)) &&
       up(foff.p, d->face_offset, 4 * (nm + 1)) && up(dmm.p, meshmat.data(), 4ull * n) && up(dface.p, face.data(), 4ull * n);
  const int init_bounds[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000};  // ordered +FLT_MAX / -FLT_MAX
  ok = ok && up(bounds.p, init_bounds, sizeof(init_bounds));
  if (!ok) return PHOS_ERR_CUDA;
  cudaStreamSynchronize(st);
  const auto t1 = std::chrono::steady_clock::now();

  const uint32_t T = 256, G = (n + T - 1) / T;
  BuildScene S{verts.p, faces.p, voff.p, foff.p, dmm.p, dface.p};
  gather_kernel<<<G, T, 0, st>>>(S, n, tris_in.p, tboxes.p, bounds.p);
  morton_kernel<<<G, T, 0, st>>>(tboxes.p, n, bounds.p, keys_in.p, vals_in.p);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.p, keys.p, vals_in.p, vals.p, (int)n, 0, 63, st);
  if (!sort_tmp.alloc(ctx, tmp_bytes, "cudaMalloc(sort)")) return PHOS_ERR_CUDA;
  cub::DeviceRadixSort::SortPairs(sort_tmp.p, tmp_bytes, keys_in.p, keys.p, vals_in.p, vals.p, (int)n, 0, 63, st);
  radix_tree_kernel<<<G, T, 0, st>>>(keys.p, (int)n, left.p, right.p, par_i.p, par_l.p);
  if (!cuda_ok(ctx, cudaMemsetAsync(visits.p, 0, 4ull * n, st), "memset (build)")) return PHOS_ERR_CUDA;
  refit_kernel<<<G, T, 0, st>>>((int)n, vals.p, tboxes.p, left.p, right.p, par_i.p, par_l.p, nboxes.p, ncount.p, visits.p);
  ctx->launches += 4;
  if (!cuda_ok(ctx, cudaGetLastError(), "build kernels")) return PHOS_ERR_CUDA;

  // output arrays: at most n - 1 wide nodes (every wide node consumes at least one binary node)
  GNode* out_nodes = nullptr;
  GTri* out_tris = nullptr;
  if (!cuda_ok(ctx, cudaMalloc(&out_nodes, (size_t)n * sizeof(GNode)), "cudaMalloc(nodes)") ||
      !cuda_ok(ctx, cudaMalloc(&out_tris, (size_t)n * sizeof(GTri)), "cudaMalloc(triangles)")) {
    if (out_nodes) cudaFree(out_nodes);
    return PHOS_ERR_CUDA;
  }
  const uint32_t root_item = 0;
  if (!cuda_ok(ctx, cudaMemsetAsync(counters.p, 0, 16, st), "memset (build)") ||  // [0] next level count, [1] triangle cursor, [2] error
      !up(items_a.p, &root_item, 4)) {
    cudaFree(out_nodes);
    cudaFree(out_tris);
    return PHOS_ERR_CUDA;
  }
  uint32_t level_first = 0, level_count = 1, depth = 0;
  uint32_t* cur = items_a.p;
  uint32_t* nxt = items_b.p;
  bool fail_cuda = false;
  while (level_count) {
    CollapseArgs A;
    A.items = cur;
    A.count = level_count;
    A.level_first = level_first;
    A.next_first = level_first + level_count;
    A.next_items = nxt;
    A.next_count = counters.p;
    A.tri_cursor = counters.p + 1;
    A.error = counters.p + 2;
    A.leaf_limit = leaf_limit;
    A.sorted = vals.p;
    A.tris_in = tris_in.p;
    A.tri_boxes = tboxes.p;
    A.left = left.p;
    A.right = right.p;
    A.node_boxes = nboxes.p;
    A.node_count = ncount.p;
    A.nodes = out_nodes;
    A.tris_out = out_tris;
    collapse_kernel<<<(level_count + 127) / 128, 128, 0, st>>>(A);
    ctx->launches++;
    uint32_t h[3] = {0, 0, 0};
    if (!cuda_ok(ctx, cudaMemcpyAsync(h, counters.p, 12, cudaMemcpyDeviceToHost, st), "build readback") ||
        !cuda_ok(ctx, cudaStreamSynchronize(st), "collapse_kernel")) {
      fail_cuda = true;
      break;
    }
    if (h[2]) {
      cudaFree(out_nodes);
      cudaFree(out_tris);
      return fail(ctx, PHOS_ERR_ACCEL, h[2] == 1 ? "internal: quantised box does not contain the child box" : "internal: leaf piece deeper than expected");
    }
    level_first += level_count;
    level_count = h[0];
    cudaMemsetAsync(counters.p, 0, 4, st);
    std::swap(cur, nxt);
    if (level_count) ++depth;
    if ((uint64_t)level_first + level_count > n) {
      cudaFree(out_nodes);
      cudaFree(out_tris);
      return fail(ctx, PHOS_ERR_ACCEL, "internal: more wide nodes than binary nodes");
    }
  }
  if (fail_cuda) {
    cudaFree(out_nodes);
    cudaFree(out_tris);
    return PHOS_ERR_CUDA;
  }
  if (depth + 2 > (uint32_t)(kSmemStack + kSpillStack)) {
    cudaFree(out_nodes);
    cudaFree(out_tris);
    return fail(ctx, PHOS_ERR_ACCEL, "tree deeper than the traversal stack");
  }
  const uint32_t n_nodes = level_first;
  // shrink the node array to its real size
  GNode* nodes = nullptr;
  if (!cuda_ok(ctx, cudaMalloc(&nodes, (size_t)n_nodes * sizeof(GNode)), "cudaMalloc(nodes)") ||
      !cuda_ok(ctx, cudaMemcpy(nodes, out_nodes, (size_t)n_nodes * sizeof(GNode), cudaMemcpyDeviceToDevice), "compact nodes")) {
    cudaFree(out_nodes);
    cudaFree(out_tris);
    if (nodes) cudaFree(nodes);
    return PHOS_ERR_CUDA;
  }
  cudaFree(out_nodes);
  cudaDeviceSynchronize();
  if (ctx->d_nodes) cudaFree(ctx->d_nodes);
  if (ctx->d_tris) cudaFree(ctx->d_tris);
  ctx->d_nodes = nodes;
  ctx->d_tris = out_tris;
  const auto t2 = std::chrono::steady_clock::now();
  phos_accel_stats& s = ctx->stats;
  s = phos_accel_stats{};
  s.nodes = n_nodes;
  s.triangles = n;
  s.max_depth = depth;
  s.max_leaf_triangles = leaf_limit;
  s.bytes_nodes = (uint64_t)n_nodes * sizeof(GNode);
  s.bytes_triangles = (uint64_t)n * sizeof(GTri);
  s.upload_seconds = std::chrono::duration<double>(t1 - t0).count();  // flatten + copy the scene in
  s.repack_seconds = std::chrono::duration<double>(t2 - t1).count();  // the device build itself
  ctx->has_accel = true;
  return PHOS_OK;
}
