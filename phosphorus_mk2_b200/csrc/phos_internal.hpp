// phos_internal.hpp — record layouts shared by the host build / re-pack and the CUDA kernels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace phos {

// ---- the reference's records, as uploaded (host side only) --------------------------------------
// mbvh::node_t<8> (reference src/accel/bvh/node.hpp:11-23), 288 B
struct RefNode {
  float bounds[48];  // minx[8] miny[8] minz[8] maxx[8] maxy[8] maxz[8]
  uint32_t offset[8];
  uint8_t num[8];
  uint32_t flags[8];  // 1 = leaf
  uint8_t pad[24];
};
static_assert(sizeof(RefNode) == 288, "node_t<8> is 288 bytes");

// accel::triangle::moeller_trumbore_t<8> (reference src/accel/triangle.hpp:24-38), 384 B
struct RefPacket {
  float e0x[8], e0y[8], e0z[8];
  float e1x[8], e1y[8], e1z[8];
  float v0x[8], v0y[8], v0z[8];
  uint32_t num;
  uint32_t meshid[8];
  uint32_t faceid[8];
  uint8_t pad[28];
};
static_assert(sizeof(RefPacket) == 384, "moeller_trumbore_t<8> is 384 bytes");

void init_ref_node(RefNode& n);

// ---- GPU-packed records --------------------------------------------------------------------------
// Compressed 8-wide node, 80 B = 5 x 128-bit loads.  Child boxes are quantised to 8 bits per plane
// on a per-node grid origin + q * 2^e (per-axis power-of-two scale), rounded outwards (conservative)
// with a safety margin that absorbs the fp32 error of the traversal's fma formulation
// (repack.cpp).  Slot i of a node is either empty, an inner child or a leaf child:
//   imask bit i  -> inner child; its node index is child_base + popc(imask & ((1 << i) - 1))
//   counts nibble i (4 bits) -> leaf child with that many triangles (1..15); its first triangle is
//                  tri_base + sum of the nibbles of the lower slots
//   empty        -> qlo > qhi on every axis (never hit)
// Children are assigned to slots so that (slot ^ ray octant) approximates front-to-back order.
struct alignas(16) GNode {
  float ox, oy, oz;      // grid origin
  uint8_t ex, ey, ez;    // biased IEEE exponents of the per-axis scale: scale = as_float(e << 23)
  uint8_t imask;
  uint32_t child_base;
  uint32_t tri_base;
  uint32_t counts;
  uint32_t spare;
  uint8_t qlox[8], qloy[8], qloz[8];
  uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(GNode) == 80, "GNode is 80 bytes");

// One triangle, 48 B = 3 x 128-bit loads: exactly the fp32 values the reference packet holds
// (e0 = b - a, e1 = c - a, v0 = a computed once on the host, src/accel/triangle.hpp:48-50), the
// hit record ids, and the triangle's position in the reference packet array (packet * 8 + lane),
// which is the brute-force kernel's visiting order and therefore the tie-break key.
struct alignas(16) GTri {
  float v0x, v0y, v0z, e0x;
  float e0y, e0z, e1x, e1y;
  float e1z;
  uint32_t meshid;  // meshid | matid << 16
  uint32_t faceid;  // 3 * face index
  uint32_t order;   // packet * 8 + lane in the uploaded packet array
};
static_assert(sizeof(GTri) == 48, "GTri is 48 bytes");

struct PackedAccel {
  std::vector<GNode> nodes;
  std::vector<GTri> tris;
  uint32_t max_depth = 0;  // node levels below the root
  uint32_t max_leaf_tris = 0;
};

// Re-pack the reference arrays; returns false and sets err on malformed input.
bool repack_accel(const RefNode* nodes, uint32_t n_nodes, const RefPacket* packets, uint32_t n_packets, PackedAccel& out,
                  std::string& err);

}  // namespace phos
