/* phos_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar CPU restatement of the ray-query hot path of jkrueger/phosphorus_mk2, used as
 * the checker for the CUDA device library.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; the product never does.
 *
 * Parity pin: the reference ships NO tests, golden vectors or fixtures for this path
 * (SURVEY.md §4, §8c).  This restatement is therefore pinned against the reference ITSELF:
 * oracle/_ref/libphos_ref.so is the reference's own accel/bvh.cpp + kernels/cpu/{stream,linear}
 * _bvh_kernel.cpp compiled from /root/reference (oracle/Makefile), and tests/test_oracle_pin.py
 * checks bit-equality of orc_brute_force with the reference's linear_mbvh_kernel_t and classifies
 * every disagreement with its stream_mbvh_kernel_t; tests/golden/ holds vectors generated that way
 * (tests/golden/make_golden.py) so the pin also holds on the GPU box where /root/reference is absent.
 *
 * Every function cites the reference file:line it follows.  All arithmetic is fp32 with the FMA
 * shapes the reference writes explicitly (src/math/simd/vector.hpp:98-109); compile with
 * -ffp-contract=off so nothing else is fused.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

/* flag bits, src/state.hpp:33-36 */
#define ORC_HIT 1u
#define ORC_MASKED 2u
#define ORC_SHADOW 4u
#define ORC_SPECULAR 8u

/* mbvh::node_t<8>, src/accel/bvh/node.hpp:11-23 — 288 bytes */
typedef struct {
  float bounds[48]; /* minx[8] miny[8] minz[8] maxx[8] maxy[8] maxz[8]  (node.hpp:40-47) */
  uint32_t offset[8];
  uint8_t num[8];
  uint32_t flags[8]; /* 1 = leaf (node.hpp:63-67) */
  uint8_t pad[24];
} orc_node;

/* accel::triangle::moeller_trumbore_t<8>, src/accel/triangle.hpp:24-38 — 384 bytes (release build) */
typedef struct {
  float e0x[8], e0y[8], e0z[8];
  float e1x[8], e1y[8], e1z[8];
  float v0x[8], v0y[8], v0z[8];
  uint32_t num;
  uint32_t meshid[8];
  uint32_t faceid[8];
  uint8_t pad[28];
} orc_packet;

/* flat SoA ray stream: the fields of ray_t<N>, src/state.hpp:39-57, for arbitrary n */
typedef struct {
  float *px, *py, *pz, *wx, *wy, *wz, *d;
  uint32_t *mesh, *face;
  float *u, *v;
  uint32_t* flags;
} orc_rays;

uint32_t orc_sizeof_node(void) { return (uint32_t)sizeof(orc_node); }
uint32_t orc_sizeof_packet(void) { return (uint32_t)sizeof(orc_packet); }

/* simd::vector3_t<8>::dot, src/math/simd/vector.hpp:98-100: madd(x, r.x, madd(y, r.y, mul(z, r.z))) */
static inline float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fmaf(ax, bx, fmaf(ay, by, az * bz));
}

/* Möller–Trumbore for one ray against one packet lane.
 * Follows moeller_trumbore_t<8>::iterate_rays, src/accel/triangle.hpp:143-164 (same arithmetic as
 * iterate_triangles :222-247): cross = msub(a, b, mul(c, d)) (vector.hpp:102-109), IEEE divide.
 * Returns 1 and writes ds/us/vs if every mask of :154-159 holds against the ray's current d. */
static inline int mt_lane(const orc_packet* k, int j, float ox, float oy, float oz, float wx, float wy, float wz,
                          float d, float* ds_out, float* us_out, float* vs_out) {
  const float e0x = k->e0x[j], e0y = k->e0y[j], e0z = k->e0z[j];
  const float e1x = k->e1x[j], e1y = k->e1y[j], e1z = k->e1z[j];
  const float tx = ox - k->v0x[j], ty = oy - k->v0y[j], tz = oz - k->v0z[j];
  /* p = wi.cross(e1) */
  const float px = fmaf(wy, e1z, -(wz * e1y));
  const float py = fmaf(wz, e1x, -(wx * e1z));
  const float pz = fmaf(wx, e1y, -(wy * e1x));
  const float det = dot3(e0x, e0y, e0z, px, py, pz);
  const float ood = 1.0f / det;
  /* q = t.cross(e0) */
  const float qx = fmaf(ty, e0z, -(tz * e0y));
  const float qy = fmaf(tz, e0x, -(tx * e0z));
  const float qz = fmaf(tx, e0y, -(ty * e0x));
  const float us = dot3(tx, ty, tz, px, py, pz) * ood;
  const float vs = dot3(wx, wy, wz, qx, qy, qz) * ood;
  const float ds = dot3(e1x, e1y, e1z, qx, qy, qz) * ood;
  const int xmask = (det > 0.00000001f) | (det < -0.00000001f);
  const int umask = us >= 0.0f;
  const int vmask = (vs >= 0.0f) & ((us + vs) <= 1.0f);
  const int dmask = (ds >= 0.0f) & (ds < d);
  *ds_out = ds;
  *us_out = us;
  *vs_out = vs;
  return xmask & umask & vmask & dmask;
}

/* One ray against one packet: iterate_rays, src/accel/triangle.hpp:166-199 — among the lanes that
 * pass, the smallest ds strictly below the ray's d wins, lowest lane on ties (__bscf walks lanes
 * upwards, strict '<'); lanes >= num are ignored (:174).  Shadow rays keep their surface record
 * (:188-195) but still shrink d and set HIT (state.hpp:118-123).  Returns 1 if the ray was updated. */
static int packet_vs_ray(const orc_packet* k, orc_rays* r, size_t i) {
  const float ox = r->px[i], oy = r->py[i], oz = r->pz[i];
  const float wx = r->wx[i], wy = r->wy[i], wz = r->wz[i];
  float closest = r->d[i];
  float bu = 0.0f, bv = 0.0f;
  int idx = -1;
  for (uint32_t j = 0; j < k->num && j < 8; ++j) {
    float ds, us, vs;
    if (mt_lane(k, (int)j, ox, oy, oz, wx, wy, wz, r->d[i], &ds, &us, &vs) && ds < closest) {
      closest = ds;
      idx = (int)j;
      bu = us;
      bv = vs;
    }
  }
  if (idx < 0) return 0;
  if (!(r->flags[i] & ORC_SHADOW)) {
    r->mesh[i] = k->meshid[idx];
    r->face[i] = k->faceid[idx];
    r->u[i] = bu;
    r->v[i] = bv;
  }
  r->flags[i] |= ORC_HIT;
  r->d[i] = closest;
  return 1;
}

/* Brute force: linear_mbvh_kernel_t::trace, src/kernels/cpu/linear_bvh_kernel.cpp:14-19 — every
 * packet in array order against every ray; MASKED rays are not traced (detail/stream.hpp:28).
 * This is the reference's own exact ground truth. */
void orc_brute_force(const orc_packet* packets, uint32_t n_packets, orc_rays* rays, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    if (rays->flags[i] & ORC_MASKED) continue;
    for (uint32_t p = 0; p < n_packets; ++p) packet_vs_ray(&packets[p], rays, (size_t)i);
  }
}

/* per-batch traversal counters (N_node, N_pkt per ray; N_tri added for the packed
 * GPU leaf layout, which stores triangles, not 8-wide packets) */
typedef struct {
  uint64_t rays;       /* rays traced (unmasked) */
  uint64_t nodes;      /* 8-wide nodes box-tested */
  uint64_t packets;    /* triangle packets MT-tested */
  uint64_t triangles;  /* triangles MT-tested (sum of leaf prim counts) */
  uint64_t max_stack;  /* deepest per-ray stack */
} orc_counters;

typedef struct {
  uint32_t offset;
  uint32_t prims; /* 0xffffffff = inner node */
  float dist;
} orc_ref;

/* Single-ray, front-to-back restatement of the stream traversal
 * (intersect<>, src/kernels/cpu/stream_bvh_kernel.cpp:17-148) on the reference tree:
 *   - MASKED rays never enter (detail/stream.hpp:24-32);
 *   - slab test per child as simd::intersect<8>, src/math/simd/aabb.hpp:33-61:
 *     near = max(tnear_x, tnear_y, tnear_z, 0), far = min(tfar_x, tfar_y, tfar_z, d), hit <=> near <= far,
 *     planes picked by the sign of 1/dir (:35-44); NaN compares false (ordered compare);
 *   - a SHADOW ray that is already HIT is dropped at the next inner node (:61-64);
 *   - a leaf runs every packet of the child, ceil(prims/8) of them from `offset` (:126-142).
 * Deliberate differences, both stated in SURVEY.md A.4/F4: 1/dir is the exact IEEE quotient
 * evaluated in double (the reference uses the ~12-bit _mm256_rcp_ps, vector.hpp:94-96, which makes
 * its own traversal drop 1e-4..1e-3 of true hits) and the interval is widened by 1e-6 relative so
 * the accepted set is a superset of the brute-force hit set; children are visited nearest-first
 * instead of the reference's per-stream summed-distance order (:99-115).  Neither changes the
 * result: the closest accepted triangle, ties broken like the brute-force kernel (lowest packet,
 * then lowest lane).  Validated == orc_brute_force in tests/test_oracle_pin.py.
 * mode_ties: bit 0: 1 = break exact-t ties by lowest (packet, lane) like brute force (default),
 *                   0 = first-found strict '<' (the stream kernel's own rule);
 *            bit 1: 1 = REFERENCE-EMULATION slab test (x86 only): 1/dir from the RCPSS instruction and
 *                   fp32 planes exactly as aabb.hpp:33-61 / stream_bvh_kernel.cpp:66-72, no widening —
 *                   reproduces the reference stream kernel's false misses; used only to show that
 *                   the oracle models the reference renderer (tests/test_oracle_render.py). */
void orc_traverse(const orc_node* nodes, const orc_packet* packets, orc_rays* rays, uint64_t n, orc_counters* c,
                  int mode_ties) {
  orc_ref stack[512];
  orc_counters cnt;
  memset(&cnt, 0, sizeof(cnt));
  for (uint64_t i = 0; i < n; ++i) {
    if (rays->flags[i] & ORC_MASKED) continue;
    cnt.rays++;
    const double ox = rays->px[i], oy = rays->py[i], oz = rays->pz[i];
    const double idx = 1.0 / (double)rays->wx[i], idy = 1.0 / (double)rays->wy[i], idz = 1.0 / (double)rays->wz[i];
    const int shadow = (rays->flags[i] & ORC_SHADOW) != 0;
    const int emulate = (mode_ties & 2) != 0;
    float ridx = 0, ridy = 0, ridz = 0;
#if defined(__x86_64__)
    if (emulate) {
      ridx = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(rays->wx[i])));
      ridy = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(rays->wy[i])));
      ridz = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(rays->wz[i])));
    }
#endif
    /* best (packet, lane) for tie-breaking */
    uint32_t best_packet = 0xffffffffu;
    int top = 0;
    stack[top].offset = 0;
    stack[top].prims = 0xffffffffu;
    stack[top].dist = 0.0f;
    top++;
    while (top > 0) {
      if ((uint64_t)top > cnt.max_stack) cnt.max_stack = (uint64_t)top;
      const orc_ref cur = stack[--top];
      const double d = rays->d[i];
      if (!emulate && (double)cur.dist > d * (1.0 + 1e-6) + 1e-30) continue;
      if (cur.prims == 0xffffffffu) {
        if (shadow && (rays->flags[i] & ORC_HIT)) continue;
        const orc_node* nd = &nodes[cur.offset];
        cnt.nodes++;
        orc_ref hit[8];
        int nh = 0;
        for (int k = 0; k < 8; ++k) {
          const double bx0 = nd->bounds[k], by0 = nd->bounds[k + 8], bz0 = nd->bounds[k + 16];
          const double bx1 = nd->bounds[k + 24], by1 = nd->bounds[k + 32], bz1 = nd->bounds[k + 40];
          if (bx0 > bx1) continue; /* empty child: min=+FLT_MAX, max=-FLT_MAX (node.hpp:25-29) */
          if (emulate) {
            const float o_x = rays->px[i], o_y = rays->py[i], o_z = rays->pz[i], dcur = rays->d[i];
            const float mnx = ((ridx >= 0.0f ? nd->bounds[k] : nd->bounds[k + 24]) - o_x) * ridx;
            const float mny = ((ridy >= 0.0f ? nd->bounds[k + 8] : nd->bounds[k + 32]) - o_y) * ridy;
            const float mnz = ((ridz >= 0.0f ? nd->bounds[k + 16] : nd->bounds[k + 40]) - o_z) * ridz;
            const float mxx = ((ridx >= 0.0f ? nd->bounds[k + 24] : nd->bounds[k]) - o_x) * ridx;
            const float mxy = ((ridy >= 0.0f ? nd->bounds[k + 32] : nd->bounds[k + 8]) - o_y) * ridy;
            const float mxz = ((ridz >= 0.0f ? nd->bounds[k + 40] : nd->bounds[k + 16]) - o_z) * ridz;
            const float nn = fmaxf(fmaxf(mnx, mny), fmaxf(mnz, 0.0f));
            const float ff = fminf(fminf(mxx, mxy), fminf(mxz, dcur));
            if (nn <= ff) {
              hit[nh].offset = nd->offset[k];
              hit[nh].prims = nd->flags[k] == 1 ? nd->num[k] : 0xffffffffu;
              hit[nh].dist = nn;
              nh++;
            }
            continue;
          }
          const double nx = ((idx >= 0.0 ? bx0 : bx1) - ox) * idx, fx = ((idx >= 0.0 ? bx1 : bx0) - ox) * idx;
          const double ny = ((idy >= 0.0 ? by0 : by1) - oy) * idy, fy = ((idy >= 0.0 ? by1 : by0) - oy) * idy;
          const double nz = ((idz >= 0.0 ? bz0 : bz1) - oz) * idz, fz = ((idz >= 0.0 ? bz1 : bz0) - oz) * idz;
          /* a NaN term (0 * inf: origin on a slab plane of a flat, axis-parallel ray) carries no
           * constraint — treat it as such (conservative), unlike the ordered compare of the reference */
          double near = 0.0, far = d;
          if (nx == nx && nx > near) near = nx;
          if (ny == ny && ny > near) near = ny;
          if (nz == nz && nz > near) near = nz;
          if (fx == fx && fx < far) far = fx;
          if (fy == fy && fy < far) far = fy;
          if (fz == fz && fz < far) far = fz;
          if (near * (1.0 - 1e-6) <= far * (1.0 + 1e-6) + 1e-30) {
            hit[nh].offset = nd->offset[k];
            hit[nh].prims = nd->flags[k] == 1 ? nd->num[k] : 0xffffffffu;
            hit[nh].dist = (float)(near * (1.0 - 2e-6));
            nh++;
          }
        }
        /* push far-to-near so the nearest child pops first */
        for (int a = 1; a < nh; ++a) {
          orc_ref t = hit[a];
          int b = a - 1;
          while (b >= 0 && hit[b].dist < t.dist) {
            hit[b + 1] = hit[b];
            --b;
          }
          hit[b + 1] = t;
        }
        for (int a = 0; a < nh; ++a) stack[top++] = hit[a];
      } else {
        uint32_t index = cur.offset;
        uint32_t prims = 0;
        do {
          const orc_packet* k = &packets[index];
          cnt.packets++;
          cnt.triangles += k->num;
          if (!(mode_ties & 1) || shadow) {
            if (packet_vs_ray(k, rays, (size_t)i)) best_packet = index;
          } else {
            /* closest-hit with brute-force tie rule: accept ds == d from a lower packet index */
            const float ox_ = rays->px[i], oy_ = rays->py[i], oz_ = rays->pz[i];
            const float wx_ = rays->wx[i], wy_ = rays->wy[i], wz_ = rays->wz[i];
            for (uint32_t j = 0; j < k->num && j < 8; ++j) {
              float ds, us, vs;
              const float dcur = rays->d[i];
              /* same masks as mt_lane but 'ds <= d' so an equal-t candidate can be inspected */
              int ok = mt_lane(k, (int)j, ox_, oy_, oz_, wx_, wy_, wz_, INFINITY, &ds, &us, &vs);
              if (!ok) continue;
              int take = 0;
              if (ds < dcur) take = 1;
              else if (ds == dcur && (rays->flags[i] & ORC_HIT) && best_packet != 0xffffffffu && index < best_packet)
                take = 1;
              if (take) {
                rays->mesh[i] = k->meshid[j];
                rays->face[i] = k->faceid[j];
                rays->u[i] = us;
                rays->v[i] = vs;
                rays->flags[i] |= ORC_HIT;
                rays->d[i] = ds;
                best_packet = index;
              }
            }
          }
          prims += 8;
          ++index;
        } while (prims < cur.prims);
      }
    }
  }
  if (c) *c = cnt;
}

/* ---------------------------------------------------------------------------------------------
 * Primary-ray generation: camera::perspective_kernel_t::operator(), src/kernels/cpu/camera.hpp:78-159,
 * pinhole branch, for one film sample jitter (jx, jy) shared by all pixels of the sample
 * (src/sampling.cpp:98-111).  Pixel (px, py) in film coordinates:
 *   ndcx = (px - 0.5) / W - 0.5         (:127  (nhalf + sx) * stepx - half)
 *   ndcy = 0.5 - (py - 0.5) / H         (:124  half - (nhalf + sy) * stepy)
 *   d    = ((ndcx + jx/W) * (W/H) * zoom, (ndcy + jy/H) * zoom, -1), zoom = 1.12 tan(fov/2) (:113,:132-133)
 *   normalise; p = (0,0,0) * M (point), d = d * M (vector), Imath row-vector convention
 *   (src/math/simd/matrix.hpp:58-104: mul then two fmadd, translation added last).
 * Deliberate difference: vector3_t::normalize (vector.hpp:126-133) multiplies by the ~12-bit
 * _mm256_rcp_ps(sqrt(l)); that approximation is micro-architecture specific, so the
 * restatement — and the GPU — use the correctly rounded 1/sqrt(l).  Directions differ from the
 * reference's by a scale factor within 1 +- 3.7e-4 (hit distances scale inversely; hit points and
 * ids are unaffected).  Rays are written in row-major pixel order over the rectangle [x0,x0+w) x [y0,y0+h).
 */
void orc_camera_rays(const float* to_world /*16*/, float fov, uint32_t W, uint32_t H, uint32_t x0, uint32_t y0,
                     uint32_t w, uint32_t h, float jx, float jy, orc_rays* out) {
  const float zoom = 1.12f * tanf(fov * 0.5f);
  const float stepx = 1.0f / (float)W, stepy = 1.0f / (float)H;
  const float ratio = (float)W / (float)H;
  const float* m = to_world;
  size_t k = 0;
  for (uint32_t y = 0; y < h; ++y) {
    const float sy = (float)(y0 + y);
    const float ndcy = 0.5f - (-0.5f + sy) * stepy;
    for (uint32_t x = 0; x < w; ++x, ++k) {
      const float sx = (float)(x0 + x);
      const float ndcx = (-0.5f + sx) * stepx - 0.5f;
      float dx = (ndcx + jx * stepx) * ratio * zoom;
      float dy = (ndcy + jy * stepy) * zoom;
      float dz = -1.0f;
      const float l = dot3(dx, dy, dz, dx, dy, dz);
      const float ool = 1.0f / sqrtf(l);
      dx *= ool;
      dy *= ool;
      dz *= ool;
      /* transform_point of (0,0,0): mul, fmadd, fmadd, then add row 3 */
      out->px[k] = fmaf(0.0f, m[8], fmaf(0.0f, m[4], 0.0f * m[0])) + m[12];
      out->py[k] = fmaf(0.0f, m[9], fmaf(0.0f, m[5], 0.0f * m[1])) + m[13];
      out->pz[k] = fmaf(0.0f, m[10], fmaf(0.0f, m[6], 0.0f * m[2])) + m[14];
      out->wx[k] = fmaf(dz, m[8], fmaf(dy, m[4], dx * m[0]));
      out->wy[k] = fmaf(dz, m[9], fmaf(dy, m[5], dx * m[1]));
      out->wz[k] = fmaf(dz, m[10], fmaf(dy, m[6], dx * m[2]));
      out->d[k] = FLT_MAX;
      out->flags[k] = 0;
    }
  }
}
