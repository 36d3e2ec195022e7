// wavefront.cu — the wavefront path tracer that replaces the per-thread tile loop of the reference
// (tile_renderer_t::render_tile / trace_rays, src/xpu/cpu.cpp:148-205) on the device:
//
//   paths_init      camera rays + path state for pixels x samples-in-flight            (camera.hpp:78-159)
//   per bounce:     trace (closest hit)                                                 (trace.cuh)
//                   shade_nee   interaction, closure, light sample -> shadow-ray queue  (deferred_shading_kernel.hpp,
//                                                                                        spt.hpp:95-149)
//                   [normals_channel: the NORMALS channel, first bounce only              (cpu.cpp:194-196)]
//                   trace (any hit, the shadow queue)
//                   integrate   radiance (+ environment on a miss), Russian roulette, bsdf_t::sample
//                               over the material's closure list, and compaction
//                               of the survivors into the next ray stream with warp-
//                               aggregated queue appends (ballot + popc + one atomic)    (spt.hpp:161-328)
//   film_accumulate radiance / (spp * pps) into the device film                          (cpu.cpp:175-198)
//
// Queue lengths live in HBM and every kernel (the traversal kernel included) reads them there, so a
// whole frame is enqueued without a single host synchronisation.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"
#include "render_state.hpp"

namespace phos {

struct FrameArgs {
  DevCamera cam;
  DevScene scene;
  uint32_t P;          // pixels in the batch (all tiles)
  uint32_t Q;          // paths in flight = P * samples in this batch
  uint32_t spp_begin;  // first sample index of the batch
  uint32_t seed;
  uint32_t max_depth;
  float scale;  // 1 / (spp_total * pps)
  const float* jitter;  // [2 * spp_total]: jx then jy
  const uint32_t* pixel;
  // Path state travels WITH the slot (throughput beta and radiance so far, [3][Q] each, ping-pong like the ray streams):
  // compaction appends survivors in whatever order the warps arrive, so state indexed by path id turns every access of
  // the later bounces into a 32-byte sector per 4-byte value (r02: 7 GB of DRAM reads in one integrate launch).  A path
  // that ends writes its radiance once to rad_final[path]; the depth of every live path is the bounce index.
  const float* beta_in;
  const float* rad_in;
  float* beta_out;
  float* rad_out;
  float* rad_final;  // per path, [3][Q]: what film_accumulate sums, in sample order
  uint32_t bounce;
  uint32_t export_light_ids;  // shade_nee also writes mesh / face / u / v of the shadow rays (16 B per slot nobody reads in a frame)
  float* n;
  float* lightw;  // per slot, [4][Q]: (light.e * 4) and 1 / pdf of the next-event sample (shade_nee -> integrate)
  uint32_t* count;  // [2]
};

// film pixel id of every tile-pixel of the batch: slot order = tile after tile, row-major inside
__global__ void pixel_table_kernel(const phos_tile* __restrict__ tiles, const unsigned long long* __restrict__ offsets, uint32_t W,
                                   uint32_t* __restrict__ pixel) {
  const phos_tile t = tiles[blockIdx.y];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.w * t.h) return;
  pixel[offsets[blockIdx.y] + k] = (t.y + k / t.w) * W + (t.x + k % t.w);
}

// prepare_sample (cpu.cpp:116-131): path q = (sample q / P, pixel q % P) starts in slot q
__global__ void paths_init_kernel(const FrameArgs A, phos_rays rays, uint32_t* __restrict__ slot_path) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) {
    A.count[0] = A.Q;
    A.count[1] = 0u;
  }
  if (q >= A.Q) return;
  const uint32_t s = A.spp_begin + q / A.P, pix = A.pixel[q % A.P];
  float o[3], w[3];
  // lens sample: the reference draws two fresh uniforms per slot and sample (sampling.cpp:104-109)
  const bool thin = A.cam.aperture_radius != 0.0f;
  const float lu = thin ? rng(A.seed, pix, s, 0, DIM_LENS) : 0.5f, lv = thin ? rng(A.seed, pix, s, 1, DIM_LENS) : 0.5f;
  camera_ray(A.cam, pix % A.cam.width, pix / A.cam.width, A.jitter[2 * s], A.jitter[2 * s + 1], lu, lv, o, w);
  rays.px[q] = o[0];
  rays.py[q] = o[1];
  rays.pz[q] = o[2];
  rays.wx[q] = w[0];
  rays.wy[q] = w[1];
  rays.wz[q] = w[2];
  rays.d[q] = 3.402823466e+38f;
  rays.flags[q] = 0u;
  slot_path[q] = q;
  A.beta_out[q] = 1.0f;  // (paths_init is handed the first bounce's input buffers as its outputs)
  A.beta_out[q + (size_t)A.Q] = 1.0f;
  A.beta_out[q + 2 * (size_t)A.Q] = 1.0f;
  A.rad_out[q] = 0.0f;
  A.rad_out[q + (size_t)A.Q] = 0.0f;
  A.rad_out[q + 2 * (size_t)A.Q] = 0.0f;
}

#ifndef PHOS_GS_BLOCKS
#define PHOS_GS_BLOCKS 16
#endif
constexpr uint32_t kGridStrideBlocksPerSm = PHOS_GS_BLOCKS;  // grid cap of the grid-stride kernels (blocks of 256 per SM)

// Interaction + next-event shadow ray for every live slot: build_interactions
// (deferred_shading_kernel.hpp:39-72), fresh_light_samples (sampling.cpp:160-180), area_light_t::sample
// (light.cpp:47-71), light_sampler_t (spt.hpp:116-148).
__device__ __forceinline__ void shade_nee_slot(const FrameArgs& A, const phos_rays& rays, const uint32_t* __restrict__ slot_path,
                                               const phos_rays& sh, uint32_t i) {
  const uint32_t flags = rays.flags[i];
  if (!(flags & PHOS_HIT) || A.scene.nlights == 0u) {  // a missed slot gets a masked query (spt.hpp:139-143)
    sh.flags[i] = PHOS_SHADOW | PHOS_MASKED;
    if (!(flags & PHOS_HIT)) return;
  }
  const v3 o = V(rays.px[i], rays.py[i], rays.pz[i]), w = V(rays.wx[i], rays.wy[i], rays.wz[i]);
  const uint32_t mm = rays.mesh[i], face = rays.face[i];
  const v3 P = add(o, scl(w, rays.d[i]));  // hits->p = p + wi * d
  const v3 n = shading_normal(A.scene, mm & 0xffffu, face, rays.u[i], rays.v[i]);
  A.n[i] = n.x;
  A.n[i + (size_t)A.Q] = n.y;
  A.n[i + 2 * (size_t)A.Q] = n.z;
  if (A.scene.nlights == 0u) return;

  const uint32_t q = slot_path[i];
  const uint32_t s = A.spp_begin + q / A.P, pix = A.pixel[q % A.P], depth = A.bounce;
  const uint32_t nl = A.scene.nlights;
  const float xl = rng(A.seed, pix, s, depth, DIM_LIGHT);
  const uint32_t l = (uint32_t)fminf(floorf(xl * nl), (float)(nl - 1));
  const float ux = rng(A.seed, pix, s, depth, DIM_LIGHT_U), uy = rng(A.seed, pix, s, depth, DIM_LIGHT_V);
  const uint32_t first = A.scene.light_first[l], num = A.scene.light_first[l + 1] - first;
  uint32_t k = (uint32_t)floorf(ux * num);  // uniform by index, pdf 1 / area (light.cpp:55,67)
  if (k > num - 1) k = num - 1;
  const float remapped = fminf(ux * num - k, 1.0f - FLT_EPSILON);
  const float sq = sqrtf(remapped);  // triangle_t::sample (mesh.cpp:318-324)
  const float bu = 1 - sq, bv = uy * sq;
  const uint32_t lm = A.scene.light_tri_mesh[first + k], lf = A.scene.light_tri_face[first + k];
  const v3 a = scene_vert(A.scene, lm & 0xffffu, lf, 0), b = scene_vert(A.scene, lm & 0xffffu, lf, 1),
           c = scene_vert(A.scene, lm & 0xffffu, lf, 2);
  const v3 L = add(add(scl(a, bu), scl(b, bv)), scl(c, 1 - bu - bv));  // barycentric_to_point (mesh.cpp:314-316)
  const float light_pdf = (1.0f / A.scene.light_area[l]) / nl;
  const v3 so = add(P, scl(n, 0.0001f));  // simd::offset
  v3 wi = sub(L, so);
  const float l2 = __fmaf_rn(wi.x, wi.x, __fmaf_rn(wi.y, wi.y, __fmul_rn(wi.z, wi.z)));
  const float dist = sqrtf(l2) - 0.0001f;
  const float ool = inv_length(A.cam, l2);  // the reference multiplies by the ~12-bit rcpps here (simd/vector.hpp:126-133)
  wi = V(wi.x * ool, wi.y * ool, wi.z * ool);
  const bool ish = __fmaf_rn(n.x, wi.x, __fmaf_rn(n.y, wi.y, __fmul_rn(n.z, wi.z))) >= 0.0f;  // simd::in_same_hemisphere
  sh.px[i] = so.x;
  sh.py[i] = so.y;
  sh.pz[i] = so.z;
  sh.wx[i] = wi.x;
  sh.wy[i] = wi.y;
  sh.wz[i] = wi.z;
  sh.d[i] = dist;
  if (A.export_light_ids) {  // the sampled light's ids ride on the shadow ray (spt.hpp:120-124): only the exported stream needs
    sh.mesh[i] = lm;         // them (phos_cuda_wavefront_rays) — inside a frame `integrate` gets the light's side of li() below
    sh.face[i] = lf;
    sh.u[i] = bu;
    sh.v[i] = bv;
  }
  sh.flags[i] = ish ? PHOS_SHADOW : (PHOS_SHADOW | PHOS_MASKED);
  // The light's side of li() (spt.hpp:237-249) is known here: (light.e * 4) and 1 / pdf with pdf = light pdf * d^2 /
  // |n_light . -wi|.  `integrate` multiplies by the BSDF value if the shadow ray arrives — it reads these 4 floats instead
  // of the light's ids off the shadow ray, its vertices, normal and material (a shadow ray that is not hit keeps d = dist).
  if (ish) {
    const v3 light_n = shading_normal(A.scene, lm & 0xffffu, lf, bu, bv);
    const DevMaterial* lmt = A.scene.mats + (lm >> 16);
    const float pdf = light_pdf * dist * dist / fabsf(dot(light_n, neg(wi)));
    const v3 le4 = scl(V(__ldg(&lmt->e[0]), __ldg(&lmt->e[1]), __ldg(&lmt->e[2])), 4);
    const size_t Q = A.Q;
    A.lightw[i] = le4.x;
    A.lightw[i + Q] = le4.y;
    A.lightw[i + 2 * Q] = le4.z;
    A.lightw[i + 3 * Q] = 1.0f / pdf;
  }
}

// Grid-stride over the live slots (the queue length lives in HBM) on a grid capped at kGridStrideBlocksPerSm blocks
// per SM: with the 16-64 M-slot wavefronts of a frame, one block per 256 slots would be up to 260 k blocks per launch,
// most of which start, read the count and leave in the late bounces.  Measured equal within the box-to-box noise at
// 64 Mi paths (config 4: 198-216 ms per frame against 206 ms) and 3 % slower at 4 Mi paths
// (profiles/r01_render_wavefront_size.log, r01_render_variants.log).
__global__ void shade_nee_kernel(const FrameArgs A, const phos_rays rays, const uint32_t* __restrict__ slot_path, int cur,
                                 phos_rays sh) {
  const uint32_t count = A.count[cur];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    shade_nee_slot(A, rays, slot_path, sh, i);
}

// ---- shading classes ------------------------------------------------------------------------------------------------
// The reference buckets the hits of a stream by material before it shades them (deferred_shading_kernel.hpp:27-33,63: one
// OSL shader group runs over one bucket).  Here the integrator's cost is the closure code of the slot's material — a warp
// whose 32 slots hold 5 materials runs all 5 (r01: 11 of 32 lanes per instruction) — so `integrate` visits the slots of
// every kBinWindow-slot window of the stream class by class:
//   class 0                          the ray missed (environment, path ends)
//   class 1 + 4 * mat + 2 * l + a    hit on material `mat`; l = 1 when the next-event shadow ray arrived (li() evaluates the
//                                    BSDF), a = 1 when the path survives Russian roulette (bsdf_t::sample runs)
// bin_window_kernel counting-sorts each window on its own (one block per window: histogram in shared memory, scan,
// scatter — no global atomics) into `perm`; the block of `integrate` that owns the window reads slot perm[k] instead of
// slot k.  The sort is LOCAL on purpose: the streams are SoA, and a stream-wide permutation turns every 4-byte field read
// into its own 32-byte sector (measured: 3.4 ms instead of 0.9 ms per launch on the bounce-2 stream of the Cornell box,
// profiles/r02_shading_classes.md); inside a window the gathers hit lines the same block is fetching anyway.
#ifndef PHOS_BIN_ROUNDS
#define PHOS_BIN_ROUNDS 4
#endif
constexpr int kBinRounds = PHOS_BIN_ROUNDS;          // slots per thread of a 256-thread block
constexpr uint32_t kBinWindow = 256u * kBinRounds;  // slots per window

#ifndef PHOS_CLASS_ALIVE
#define PHOS_CLASS_ALIVE 1
#endif
__device__ __forceinline__ uint32_t shade_class(const FrameArgs& A, const phos_rays& rays, const phos_rays& sh,
                                                const uint32_t* __restrict__ slot_path, uint32_t i) {
  if (!(rays.flags[i] & PHOS_HIT)) return 0u;
  const uint32_t mat = min(rays.mesh[i] >> 16, (kShadeClasses - 4u) / 4u - 1u);
  const uint32_t lit = (sh.flags[i] & (PHOS_HIT | PHOS_MASKED)) ? 0u : 1u;
  uint32_t alive = 1u;
#if PHOS_CLASS_ALIVE
  // terminate_path as integrate_kernel decides it (only the grouping depends on this copy, not the image)
  const uint32_t depth = A.bounce + 1u;
  if (depth >= A.max_depth) {
    alive = 0u;
  } else if (depth >= 3u) {
    const size_t Q = A.Q;
    const uint32_t q = slot_path[i];
    const float yb = 0.212671f * A.beta_in[i] + 0.715160f * A.beta_in[i + Q] + 0.072169f * A.beta_in[i + 2 * Q];
    alive = rng(A.seed, A.pixel[q % A.P], A.spp_begin + q / A.P, depth - 1u, DIM_RR) >= fmaxf(0.05f, 1.0f - yb) ? 1u : 0u;
  }
#endif
  return 1u + 4u * mat + 2u * lit + alive;
}

__global__ void __launch_bounds__(256) bin_window_kernel(const FrameArgs A, int cur, const phos_rays rays, const phos_rays sh,
                                                         const uint32_t* __restrict__ slot_path, uint32_t* __restrict__ perm) {
  __shared__ uint32_t hist[kShadeClasses];  // slots per class, then first position of the class in the window
  __shared__ uint32_t warp_sum[8];
  const uint32_t n = A.count[cur];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t base = blockIdx.x * kBinWindow; base < n; base += gridDim.x * kBinWindow) {  // block-uniform trip count
    for (uint32_t c = threadIdx.x; c < kShadeClasses; c += 256u) hist[c] = 0u;
    __syncthreads();
    uint32_t cls[kBinRounds], off[kBinRounds];
#pragma unroll
    for (int j = 0; j < kBinRounds; ++j) {
      const uint32_t i = base + j * 256u + threadIdx.x;
      cls[j] = 0xffffffffu;
      off[j] = 0u;
      if (i < n) {
        const uint32_t c = shade_class(A, rays, sh, slot_path, i);
        const unsigned peers = __match_any_sync(__activemask(), c);
        const int leader = __ffs(peers) - 1;
        uint32_t o = 0u;
        if ((int)lane == leader) o = atomicAdd(&hist[c], (uint32_t)__popc(peers));
        off[j] = __shfl_sync(peers, o, leader) + __popc(peers & ((1u << lane) - 1u));
        cls[j] = c;
      }
    }
    __syncthreads();
    // exclusive scan of hist[0 .. ncls): every thread owns 4 consecutive classes
    {
      const uint32_t c0 = threadIdx.x * 4u;
      const uint32_t a = hist[c0], b = hist[c0 + 1], c = hist[c0 + 2], d = hist[c0 + 3];
      uint32_t incl = a + b + c + d;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, s);
        if ((int)lane >= s) incl += v;
      }
      if (lane == 31u) warp_sum[warp] = incl;
      __syncthreads();
      uint32_t before = incl - (a + b + c + d);
      for (unsigned w = 0; w < warp; ++w) before += warp_sum[w];
      hist[c0] = before;
      hist[c0 + 1] = before + a;
      hist[c0 + 2] = before + a + b;
      hist[c0 + 3] = before + a + b + c;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kBinRounds; ++j)
      if (cls[j] != 0xffffffffu) perm[base + hist[cls[j]] + off[j]] = base + j * 256u + threadIdx.x;
    __syncthreads();
  }
}

// integrator_t::operator() (spt.hpp:161-210) with li (:212-255), sample_bsdf (:257-305) and
// terminate_path (:307-328); survivors are appended to the next ray stream.
#ifndef PHOS_INT_HOIST_LIT
#define PHOS_INT_HOIST_LIT 1
#endif
#ifndef PHOS_INTEGRATE_MIN_BLOCKS
// 3 blocks per SM = 80 registers.  With the loads of a slot batched into two round trips (above) the kernel holds ~25 loaded
// values at once: at 64 registers (4 blocks, the r01 choice for the unbatched form) it spills 172 bytes and the batching
// loses 1 %; at 80 it spills 68 bytes and gains 1.4 % on the Cornell box, 0.4 % on config 4; 48 registers: -10 %
// (profiles/r02_render_integrate_batched_loads.log); at 2 blocks (102 registers, no spill at all) it loses 4.3 % / 1.4 %: the
// spill-free kernel is not the fast one.
#define PHOS_INTEGRATE_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(256, PHOS_INTEGRATE_MIN_BLOCKS) integrate_kernel(const FrameArgs A, const phos_rays rays, const phos_rays sh,
                                 const uint32_t* __restrict__ slot_path, int cur, phos_rays next, uint32_t* __restrict__ next_path,
                                 const uint32_t* __restrict__ perm) {
 // grid-stride over the live slots, a whole warp at a time (the compaction below is a warp collective); with `perm` the
 // slots come class by class (bin_* kernels above)
 const uint32_t count = A.count[cur];
 // a block walks whole windows of kBinWindow consecutive positions (4 rounds of 256): with `perm` these are the slots of one
 // window of the stream class by class, fetched by this block alone
 for (uint32_t k0 = blockIdx.x * kBinWindow + threadIdx.x; (k0 & ~(kBinWindow - 1u)) < count; k0 += gridDim.x * kBinWindow)
 for (uint32_t k = k0; k < k0 + kBinWindow && (k & ~31u) < count; k += 256u) {
  const bool valid = k < count;
  const uint32_t i = (perm && valid) ? perm[k] : k;
  bool alive = false;
  uint32_t q = 0, nflags = 0;
  v3 no = V(0, 0, 0), nw = V(0, 0, 0), nbeta = V(0, 0, 0), rad = V(0, 0, 0);
  // This kernel waits on memory (ncu: 58 % of the stall samples are long-scoreboard, 20 % of the issue slots used): what
  // matters is the number of DEPENDENT round trips per slot.  Everything that hangs off the slot index alone is fetched in
  // one of them — the hit flags, the path id, the material word and the verdict of the next-event shadow ray — and the
  // fields of a hit slot in the next (the shadow direction and the light terms only where the shadow ray arrived).
  uint32_t fl = 0, meshw = 0, sflags = PHOS_MASKED;
  if (valid) {
    fl = rays.flags[i];
    q = slot_path[i];
    meshw = rays.mesh[i];
    sflags = sh.flags[i];
  }
  if (valid && !(fl & PHOS_HIT)) {  // a miss adds beta * e_env (spt.hpp:199-202) and ends the path
    const uint32_t p = q;
    const size_t Q = A.Q;
    v3 rad = V(A.rad_in[i], A.rad_in[i + Q], A.rad_in[i + 2 * Q]);
    if (A.scene.environment >= 0) {
      const DevMaterial* env = A.scene.mats + A.scene.environment;
      rad.x += A.beta_in[i] * __ldg(&env->e[0]);
      rad.y += A.beta_in[i + Q] * __ldg(&env->e[1]);
      rad.z += A.beta_in[i + 2 * Q] * __ldg(&env->e[2]);
    }
    A.rad_final[p] = rad.x;
    A.rad_final[p + Q] = rad.y;
    A.rad_final[p + 2 * Q] = rad.z;
  } else if (valid) {
    const size_t Q = A.Q;
    const bool lit = !(sflags & (PHOS_HIT | PHOS_MASKED));  // the next-event shadow ray arrived
    const DevMaterial* mt = A.scene.mats + (meshw >> 16);
    // second round trip: all of these are independent loads
    const uint32_t kind = __ldg(&mt->kind), nlobes = __ldg(&mt->nlobes);
    const uint32_t s = A.spp_begin + q / A.P, pix = A.pixel[q % A.P];
    v3 beta = V(A.beta_in[i], A.beta_in[i + Q], A.beta_in[i + 2 * Q]);
    rad = V(A.rad_in[i], A.rad_in[i + Q], A.rad_in[i + 2 * Q]);
    const v3 o = V(rays.px[i], rays.py[i], rays.pz[i]), w = V(rays.wx[i], rays.wy[i], rays.wz[i]);
    const float dist = rays.d[i];
    const v3 n = V(A.n[i], A.n[i + Q], A.n[i + 2 * Q]);
    v3 swi = V(0, 0, 0), lw = V(0, 0, 0);
    float lw_rcp_pdf = 0.0f;
    if (PHOS_INT_HOIST_LIT && lit) {
      swi = V(sh.wx[i], sh.wy[i], sh.wz[i]);
      lw = V(A.lightw[i], A.lightw[i + Q], A.lightw[i + 2 * Q]);
      lw_rcp_pdf = A.lightw[i + 3 * Q];
    }
    uint32_t depth = A.bounce;
    const v3 P = add(o, scl(w, dist));
    const v3 wo = neg(w);
    const bool emitter = kind == PHOS_MAT_EMITTER;
    if (emitter && (depth == 0 || (fl & PHOS_SPECULAR)))
      rad = add(rad, mul(beta, V(__ldg(&mt->e[0]), __ldg(&mt->e[1]), __ldg(&mt->e[2]))));
    if (lit && nlobes != 0) {  // li(), spt.hpp:212-255; a 0-lobe BSDF evaluates to 0
      if (!PHOS_INT_HOIST_LIT) {
        swi = V(sh.wx[i], sh.wy[i], sh.wz[i]);
        lw = V(A.lightw[i], A.lightw[i + Q], A.lightw[i + 2 * Q]);
        lw_rcp_pdf = A.lightw[i + 3 * Q];
      }
      const v3 f = bsdf_f(mt, n, swi, wo);
      const v3 li = scl(mul(lw, f), lw_rcp_pdf);  // (light.e * 4) * f * (1 / pdf)
      rad = add(rad, mul(beta, li));
    }
    ++depth;
    // terminate_path
    float wgt = 1.0f;
    alive = depth < A.max_depth;
    if (alive && depth >= 3) {
      const float yb = 0.212671f * beta.x + 0.715160f * beta.y + 0.072169f * beta.z;
      const float qq = fmaxf(0.05f, 1.0f - yb);
      alive = rng(A.seed, pix, s, depth - 1, DIM_RR) >= qq;
      if (alive) wgt = (1.0f / (1.0f - qq));
    }
    beta = scl(beta, wgt);
    if (alive && nlobes == 0) alive = false;  // 0-lobe BSDF: the path ends on an emitter
    if (alive) {
      // bsdf_t::sample (bsdf.cpp:133-248), then sample_bsdf (spt.hpp:289-303)
      const float su = rng(A.seed, pix, s, depth - 1, DIM_BSDF_U);
      const float sv = rng(A.seed, pix, s, depth - 1, DIM_BSDF_V);
      v3 sampled = V(0, 0, 0), f = V(0, 0, 0);
      float pdf = 0.0f;
      alive = bsdf_sample(mt, n, su, sv, wo, sampled, f, pdf, nflags);
      if (alive && ((f.x == 0.0f && f.y == 0.0f && f.z == 0.0f) || pdf == 0.0f)) alive = false;
      if (alive) {
        const float weight = dot(n, sampled);
        beta = mul(beta, scl(f, fabsf(weight) / pdf));
        no = add(P, scl(n, weight < 0.0f ? -0.0001f : 0.0001f));  // offset() (math/vector.hpp:14-21)
        nw = sampled;
      }
    }
    nbeta = beta;
    if (!alive) {  // the path ends here: its radiance goes to the film
      A.rad_final[q] = rad.x;
      A.rad_final[q + Q] = rad.y;
      A.rad_final[q + 2 * Q] = rad.z;
    }
  }
  // compaction: one atomic per warp reserves slots for all its survivors
  const unsigned live = __ballot_sync(0xffffffffu, alive);
  if (live == 0u) continue;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == (unsigned)(__ffs(live) - 1)) base = atomicAdd(&A.count[cur ^ 1], (uint32_t)__popc(live));
  base = __shfl_sync(0xffffffffu, base, __ffs(live) - 1);
  if (alive) {
    const uint32_t slot = base + __popc(live & ((1u << lane) - 1u));
    next.px[slot] = no.x;
    next.py[slot] = no.y;
    next.pz[slot] = no.z;
    next.wx[slot] = nw.x;
    next.wy[slot] = nw.y;
    next.wz[slot] = nw.z;
    next.d[slot] = 3.402823466e+38f;
    next.flags[slot] = (nflags & BSDF_SPECULAR_F) ? PHOS_SPECULAR : 0u;  // rays->specular_bounce (spt.hpp:302)
    next_path[slot] = q;
    const size_t Q = A.Q;
    A.beta_out[slot] = nbeta.x;
    A.beta_out[slot + Q] = nbeta.y;
    A.beta_out[slot + 2 * Q] = nbeta.z;
    A.rad_out[slot] = rad.x;
    A.rad_out[slot + Q] = rad.y;
    A.rad_out[slot + 2 * Q] = rad.z;
  }
 }
}

// channels.primary->add(x, y, r * (1 / (spp * pps))) per sample (cpu.cpp:175-198), samples in order
__global__ void film_accumulate_kernel(const FrameArgs A, float* __restrict__ film) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.P) return;
  float* px = film + 4 * (size_t)A.pixel[t];
  float r = px[0], g = px[1], b = px[2];
  for (uint32_t q = t; q < A.Q; q += A.P) {
    r += A.rad_final[q] * A.scale;
    g += A.rad_final[q + (size_t)A.Q] * A.scale;
    b += A.rad_final[q + 2 * (size_t)A.Q] * A.scale;
  }
  px[0] = r;
  px[1] = g;
  px[2] = b;
  px[3] = 1.0f;
}

// render_buffer_t::NORMALS (cpu.cpp:97,194-196): channels.normals->set(x, y, primary->n) after every sample
// whose primary ray hit, samples in order — so a pixel ends up with the shading normal of its LAST sample
// that hit.  Runs after the first shade_nee of a sample batch (slot = path = tile pixel + sample * P there).
__global__ void normals_channel_kernel(const FrameArgs A, const phos_rays rays, uint32_t samples, float* __restrict__ film_n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.P) return;
  for (uint32_t s = samples; s-- > 0;) {
    const size_t q = t + (size_t)s * A.P;
    if (!(rays.flags[q] & PHOS_HIT)) continue;
    float* o = film_n + 3 * (size_t)A.pixel[t];
    o[0] = A.n[q];
    o[1] = A.n[q + (size_t)A.Q];
    o[2] = A.n[q + 2 * (size_t)A.Q];
    return;
  }
}

__global__ void zero_u32_kernel(uint32_t* p) { *p = 0u; }

// sample::stratified_2d (math/sampling.hpp:67-82) as sampler_t::preprocess uses it (sampling.cpp:96-99):
// one film jitter per sample index, spd = lround(sqrt(spp)) strata per axis.  Entries >= spd^2 are
// uninitialised in the reference (and spd^2 > spp overruns its array); 0.5 here.
void film_jitter(uint32_t seed, uint32_t spp, std::vector<float>& out) {
  out.assign(2 * (size_t)spp, 0.5f);
  const uint32_t num = (uint32_t)lroundf(sqrtf((float)spp));
  const float step = 1.0f / (float)num;
  float dy = 0.0f;
  uint32_t call = 0;
  for (uint32_t i = 0; i < num; ++i, dy += step) {
    float dx = 0.0f;
    for (uint32_t j = 0; j < num; ++j, dx += step) {
      const float a = rng(seed, 0xffffffffu, call++, 0, DIM_FILM);
      const float b = rng(seed, 0xffffffffu, call++, 0, DIM_FILM);
      if (j * num + i < spp) {
        out[2 * (size_t)(j * num + i)] = dx + a * step;
        out[2 * (size_t)(j * num + i) + 1] = dy + b * step;
      }
    }
  }
}

}  // namespace phos

using namespace phos;

extern "C" {

int phos_cuda_render(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, uint32_t spp_begin, uint32_t spp_end,
                     uint32_t spp_total, uint64_t seed64) {
  if (!ctx || !tiles) return PHOS_ERR_INVALID;
  if (!ctx->render || !ctx->render->film) return fail(ctx, PHOS_ERR_INVALID, "render before upload_scene");
  if (!ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "render before upload_accel");
  if (spp_begin >= spp_end || spp_end > spp_total) return fail(ctx, PHOS_ERR_INVALID, "bad sample range");
  if (n_tiles == 0) return PHOS_OK;
  cudaSetDevice(ctx->device);
  RenderState& R = *ctx->render;
  const uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));

  std::vector<unsigned long long> offsets(n_tiles);
  unsigned long long total = 0;
  uint32_t max_px = 0;
  for (uint32_t i = 0; i < n_tiles; ++i) {
    if (tiles[i].x + tiles[i].w > R.camera.width || tiles[i].y + tiles[i].h > R.camera.height)
      return fail(ctx, PHOS_ERR_INVALID, "tile outside the film");
    offsets[i] = total;
    total += (unsigned long long)tiles[i].w * tiles[i].h;
    max_px = std::max(max_px, tiles[i].w * tiles[i].h);
  }
  if (total == 0) return PHOS_OK;
  if (total > 0x7fffffffull) return fail(ctx, PHOS_ERR_INVALID, "too many pixels in one render call");
  const uint32_t P = (uint32_t)total;
  // Paths in flight per batch.  Every trace launch ends in a tail in which the last long rays finish on a draining
  // machine (a ray lives ~50 us; the slowest of a launch several times that), and a batch has 2 x depth of them: at
  // 4 Mi paths the tails were a quarter of a config-4 frame (profiles/r01_render_wavefront_size.log: 4 / 8 / 16 / 32 /
  // 64 Mi paths = 513 / 582 / 628 / 656 / 670 M samples/s).  64 Mi paths (over both wavefronts) are 13 GB of wavefront
  // state (kBytesPerPath per path) on a 180 GB device; never more than a quarter of what is free.  PHOS_WAVEFRONT_PATHS overrides.
  constexpr uint64_t kBytesPerPath = 248ull;  // ensure_wavefront: 3 ray streams, slot maps, normals, light terms, 2 x (beta, rad), rad_final, perm
  uint64_t target = 64ull << 20;
  if (R.paths_cap == 0) {  // asked once per scene upload: cudaMemGetInfo is a driver round trip, not something for every frame
    size_t free_b = 0, total_b = 0;
    uint64_t held = 0;
    for (const Wavefront& w : R.wfs) held += w.capacity;
    R.paths_cap = ~0ull;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
      R.paths_cap = std::max<uint64_t>(((uint64_t)free_b + held * kBytesPerPath) / 4ull / kBytesPerPath, 1ull << 20);
    else
      cudaGetLastError();
  }
  target = std::min<uint64_t>(target, R.paths_cap);
  if (const char* e = std::getenv("PHOS_WAVEFRONT_PATHS")) target = std::max<uint64_t>(1ull << 16, std::strtoull(e, nullptr, 10));
  // Several wavefronts: the sample batches of a frame go round-robin over nw wavefronts on nw streams.  r01 had this off
  // (the tail of one batch's launches under the other's full launches bought 1.7 % on config 4 and cost 20 % on the Cornell
  // box); with the round-2 shading kernels, which are bound by memory latency while the traversal is bound by issue slots,
  // the two overlap: config 4 747 -> 764 M samples/s, Cornell box 1296 -> 1383, 30 M-triangle field 576 -> 585
  // (profiles/r02_render_wavefronts.log) — two wavefronts by default (PHOS_WAVEFRONTS=1..4).  Film accumulation stays in
  // sample order (events chain the streams there), so the image does not depend on it.
  // Always one wavefront when the NORMALS channel is on (its "last sample that hit" is order dependent).
  int nw = 2;
  if (const char* e = std::getenv("PHOS_WAVEFRONTS")) nw = std::max(1, std::min(RenderState::kMaxWavefronts, std::atoi(e)));
  if (R.film_normals) nw = 1;
  nw = (int)std::min<uint32_t>((uint32_t)nw, spp_end - spp_begin);
  uint32_t batch = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(spp_end - spp_begin, target / ((uint64_t)P * nw)));
  if (nw > 1) batch = std::min<uint32_t>(batch, (spp_end - spp_begin + nw - 1) / nw);  // at least nw batches to overlap
  for (int w = 0; w < nw; ++w)
    if (!R.ensure_wavefront(ctx, (uint64_t)P * batch, P, w)) return PHOS_ERR_CUDA;
  if (!R.set_tiles(ctx, tiles, offsets.data(), n_tiles)) return PHOS_ERR_CUDA;
  Wavefront* WF[RenderState::kMaxWavefronts];
  for (int w = 0; w < RenderState::kMaxWavefronts; ++w) WF[w] = &R.wfs[w];
  cudaStream_t st = ctx->stream;
  // (the extra streams are those of the host-pointer query pipeline: one host thread drives a context, never both at once)
  cudaStream_t streams[RenderState::kMaxWavefronts] = {ctx->stream, ctx->s_cmp, ctx->s_in, ctx->s_out};

  for (int w = 0; w < nw; ++w)
  for (uint32_t first = 0; first < n_tiles; first += 65535) {
    const uint32_t cnt = std::min<uint32_t>(65535, n_tiles - first);
    pixel_table_kernel<<<dim3((max_px + 255) / 256, cnt), 256, 0, st>>>(R.d_tiles + first, R.d_tile_offsets + first, R.camera.width,
                                                                         WF[w]->pixel);
    ctx->launches++;
  }
  // film jitter table for this (seed, spp_total)
  {
    std::vector<float> jit;
    film_jitter(seed, spp_total, jit);
    if (R.jitter_capacity < spp_total) {
      cudaStreamSynchronize(st);
      if (R.d_jitter) cudaFree(R.d_jitter);
      R.d_jitter = nullptr;
      if (!cuda_ok(ctx, cudaMalloc(&R.d_jitter, 2 * sizeof(float) * spp_total), "cudaMalloc(jitter)")) return PHOS_ERR_CUDA;
      R.jitter_capacity = spp_total;
    }
    if (!cuda_ok(ctx, cudaMemcpyAsync(R.d_jitter, jit.data(), 2 * sizeof(float) * spp_total, cudaMemcpyHostToDevice, st), "upload jitter"))
      return PHOS_ERR_CUDA;
    cudaStreamSynchronize(st);  // jit goes out of scope
  }

  FrameArgs AW[RenderState::kMaxWavefronts];
  for (int w = 0; w < nw; ++w) {
    FrameArgs& A = AW[w];
    const Wavefront& W = *WF[w];
    A.cam = R.camera;
    A.scene = R.scene;
    A.P = P;
    A.seed = seed;
    A.max_depth = ctx->opt.path_depth;
    A.scale = 1.0f / (float)(spp_total * ctx->opt.paths_per_sample);
    A.jitter = R.d_jitter;
    A.pixel = W.pixel;
    A.beta_in = A.beta_out = W.beta[0];
    A.rad_in = A.rad_out = W.rad[0];
    A.rad_final = W.rad_final;
    A.bounce = 0;
    A.export_light_ids = std::getenv("PHOS_SHADE_EXPORT_IDS") ? 1u : 0u;  // (A/B switch of profiles/r02_render_shadow_ids.log)
    A.n = W.n;
    A.lightw = W.lightw;
    A.count = W.count;
  }
  // the second stream starts after the set-up on the first (pixel tables, jitter table); film_accumulate launches are
  // chained in batch order; the first stream ends by waiting for the second
  cudaEvent_t ev_setup = nullptr, ev_film[RenderState::kMaxWavefronts] = {nullptr, nullptr, nullptr, nullptr};
  bool ok = true;
  if (nw > 1) {
    ok = cuda_ok(ctx, cudaEventCreateWithFlags(&ev_setup, cudaEventDisableTiming), "event") && cuda_ok(ctx, cudaEventRecord(ev_setup, st), "event record");
    for (int w = 0; ok && w < nw; ++w) ok = cuda_ok(ctx, cudaEventCreateWithFlags(&ev_film[w], cudaEventDisableTiming), "event");
    for (int w = 1; ok && w < nw; ++w) ok = cuda_ok(ctx, cudaStreamWaitEvent(streams[w], ev_setup, 0), "event wait");
  }
  // shading classes (above): on unless the scene has a single material or PHOS_SHADE_BIN=0
  bool bin = R.num_materials > 1;
  if (const char* e = std::getenv("PHOS_SHADE_BIN")) bin = std::atoi(e) != 0;
  int rc = PHOS_OK;
  uint32_t k = 0;
  for (uint32_t s0 = spp_begin; ok && rc == PHOS_OK && s0 < spp_end; s0 += batch, ++k) {
    const int w = (int)(k % (uint32_t)nw);
    cudaStream_t sk = streams[w];
    FrameArgs& A = AW[w];
    Wavefront& W = *WF[w];
    unsigned long long* cursors = ctx->d_counters + 24 + 2 * w;
    const uint32_t ns = std::min(batch, spp_end - s0);
    A.Q = P * ns;
    A.spp_begin = s0;
    const uint32_t blocks = (A.Q + 255) / 256;
    const uint32_t gs_blocks = std::min<uint32_t>(blocks, (uint32_t)ctx->sm_count * kGridStrideBlocksPerSm);  // grid-stride kernels
    A.beta_out = W.beta[0];
    A.rad_out = W.rad[0];
    paths_init_kernel<<<blocks, 256, 0, sk>>>(A, W.rays[0], W.slot_path[0]);
    ctx->launches++;
    // (path_depth 0 still traces the primary rays and adds what they see: the reference's loop tests the depth after the
    // first bounce, spt.hpp:307-328)
    for (uint32_t b = 0; b < std::max<uint32_t>(1u, A.max_depth) && rc == PHOS_OK; ++b) {
      const int cur = (int)(b & 1u);
      A.bounce = b;
      A.beta_in = W.beta[cur];
      A.rad_in = W.rad[cur];
      A.beta_out = W.beta[cur ^ 1];
      A.rad_out = W.rad[cur ^ 1];
      zero_u32_kernel<<<1, 1, 0, sk>>>(W.count + (cur ^ 1));
      ctx->launches++;
      rc = launch_trace(ctx, W.rays[cur], A.Q, sk, cursors, false, W.count + cur);
      if (rc) break;
      shade_nee_kernel<<<gs_blocks, 256, 0, sk>>>(A, W.rays[cur], W.slot_path[cur], cur, W.shadow);
      ctx->launches++;
      if (b == 0 && R.film_normals) {
        normals_channel_kernel<<<(P + 255) / 256, 256, 0, sk>>>(A, W.rays[0], ns, R.film_normals);
        ctx->launches++;
      }
      rc = launch_trace(ctx, W.shadow, A.Q, sk, cursors + 1, false, W.count + cur);
      if (rc) break;
      if (bin) {
        bin_window_kernel<<<gs_blocks, 256, 0, sk>>>(A, cur, W.rays[cur], W.shadow, W.slot_path[cur], W.perm);
        ctx->launches++;
      }
      integrate_kernel<<<gs_blocks, 256, 0, sk>>>(A, W.rays[cur], W.shadow, W.slot_path[cur], cur, W.rays[cur ^ 1], W.slot_path[cur ^ 1],
                                                  bin ? W.perm : nullptr);
      ctx->launches++;
    }
    if (rc) break;
    if (nw > 1 && k > 0) ok = cuda_ok(ctx, cudaStreamWaitEvent(sk, ev_film[(k - 1) % (uint32_t)nw], 0), "event wait");  // sample order
    film_accumulate_kernel<<<(P + 255) / 256, 256, 0, sk>>>(A, R.film);
    ctx->launches++;
    if (nw > 1) ok = ok && cuda_ok(ctx, cudaEventRecord(ev_film[k % (uint32_t)nw], sk), "event record");
  }
  if (nw > 1) {
    // everything the other streams did is ordered before whatever the caller enqueues next on the first: the film
    // accumulations are chained, so the last batch's event covers all of them
    if (ok && rc == PHOS_OK && k > 0)
      ok = cuda_ok(ctx, cudaStreamWaitEvent(st, ev_film[(k - 1) % (uint32_t)nw], 0), "event wait");
    else
      for (int w = 1; w < nw; ++w) cudaStreamSynchronize(streams[w]);
    if (ev_setup) cudaEventDestroy(ev_setup);
    for (cudaEvent_t e : ev_film)
      if (e) cudaEventDestroy(e);
  }
  if (rc) return rc;
  return ok && cuda_ok(ctx, cudaGetLastError(), "render launch") ? PHOS_OK : PHOS_ERR_CUDA;
}

// The ray streams of BASELINE config 3, produced by the pipeline itself: run ONE bounce of sample
// `sample` over the tiles and hand back either the next-event shadow-ray stream (which = 1: one query per
// primary slot, SHADOW or SHADOW|MASKED, tmax = distance to the sampled light point) or the BSDF-sampled
// bounce-ray stream (which = 0: cosine-sampled diffuse / GGX rays from the primary hits, compacted).
int phos_cuda_wavefront_rays(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, uint32_t sample, uint32_t spp_total,
                             uint64_t seed64, int which, const phos_rays* out, uint64_t capacity, uint64_t* out_count) {
  if (!ctx || !tiles || !out || !out_count) return PHOS_ERR_INVALID;
  if (!ctx->render || !ctx->render->film || !ctx->has_accel) return fail(ctx, PHOS_ERR_INVALID, "wavefront_rays before upload_scene / upload_accel");
  if (sample >= spp_total) return fail(ctx, PHOS_ERR_INVALID, "bad sample index");
  cudaSetDevice(ctx->device);
  RenderState& R = *ctx->render;
  const uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
  std::vector<unsigned long long> offsets(n_tiles);
  unsigned long long total = 0;
  uint32_t max_px = 0;
  for (uint32_t i = 0; i < n_tiles; ++i) {
    if (tiles[i].x + tiles[i].w > R.camera.width || tiles[i].y + tiles[i].h > R.camera.height)
      return fail(ctx, PHOS_ERR_INVALID, "tile outside the film");
    offsets[i] = total;
    total += (unsigned long long)tiles[i].w * tiles[i].h;
    max_px = std::max(max_px, tiles[i].w * tiles[i].h);
  }
  if (total == 0 || total > 0x7fffffffull || n_tiles > 65535) return fail(ctx, PHOS_ERR_INVALID, "bad tile list");
  const uint32_t P = (uint32_t)total;
  if (capacity < P) return fail(ctx, PHOS_ERR_INVALID, "output stream too small");
  if (!R.ensure_wavefront(ctx, P, P) || !R.set_tiles(ctx, tiles, offsets.data(), n_tiles)) return PHOS_ERR_CUDA;
  Wavefront& W = R.wfs[0];
  cudaStream_t st = ctx->stream;
  pixel_table_kernel<<<dim3((max_px + 255) / 256, n_tiles), 256, 0, st>>>(R.d_tiles, R.d_tile_offsets, R.camera.width, W.pixel);
  std::vector<float> jit;
  film_jitter(seed, spp_total, jit);
  if (R.jitter_capacity < spp_total) {
    cudaStreamSynchronize(st);
    if (R.d_jitter) cudaFree(R.d_jitter);
    R.d_jitter = nullptr;
    if (!cuda_ok(ctx, cudaMalloc(&R.d_jitter, 2 * sizeof(float) * spp_total), "cudaMalloc(jitter)")) return PHOS_ERR_CUDA;
    R.jitter_capacity = spp_total;
  }
  if (!cuda_ok(ctx, cudaMemcpyAsync(R.d_jitter, jit.data(), 2 * sizeof(float) * spp_total, cudaMemcpyHostToDevice, st), "upload jitter"))
    return PHOS_ERR_CUDA;
  cudaStreamSynchronize(st);
  FrameArgs A;
  A.cam = R.camera;
  A.scene = R.scene;
  A.P = P;
  A.Q = P;
  A.spp_begin = sample;
  A.seed = seed;
  A.max_depth = std::max<uint32_t>(2, ctx->opt.path_depth);  // the first bounce must be allowed to continue
  A.scale = 0.0f;
  A.jitter = R.d_jitter;
  A.pixel = W.pixel;
  A.beta_in = A.beta_out = W.beta[0];
  A.rad_in = A.rad_out = W.rad[0];
  A.rad_final = W.rad_final;
  A.bounce = 0;
  A.export_light_ids = 1;
  A.n = W.n;
  A.lightw = W.lightw;
  A.count = W.count;
  const uint32_t blocks = (P + 255) / 256;
  const uint32_t gs_blocks = std::min<uint32_t>(blocks, (uint32_t)ctx->sm_count * kGridStrideBlocksPerSm);
  paths_init_kernel<<<blocks, 256, 0, st>>>(A, W.rays[0], W.slot_path[0]);
  int rc = launch_trace(ctx, W.rays[0], P, st, ctx->d_counters + 24, false, W.count);
  if (rc) return rc;
  shade_nee_kernel<<<gs_blocks, 256, 0, st>>>(A, W.rays[0], W.slot_path[0], 0, W.shadow);
  const phos_rays* src = &W.shadow;
  uint32_t n = P;
  if (which == 0) {
    rc = launch_trace(ctx, W.shadow, P, st, ctx->d_counters + 25, false, W.count);
    if (rc) return rc;
    A.beta_out = W.beta[1];
    A.rad_out = W.rad[1];
    integrate_kernel<<<gs_blocks, 256, 0, st>>>(A, W.rays[0], W.shadow, W.slot_path[0], 0, W.rays[1], W.slot_path[1], nullptr);
    if (!cuda_ok(ctx, cudaMemcpyAsync(&n, W.count + 1, 4, cudaMemcpyDeviceToHost, st), "read queue length") ||
        !cuda_ok(ctx, cudaStreamSynchronize(st), "wavefront_rays"))
      return PHOS_ERR_CUDA;
    src = &W.rays[1];
  }
  ctx->launches += 5;
  const void* sp[12] = {src->px, src->py, src->pz, src->wx, src->wy, src->wz, src->d, src->mesh, src->face, src->u, src->v, src->flags};
  void* dp[12] = {out->px, out->py, out->pz, out->wx, out->wy, out->wz, out->d, out->mesh, out->face, out->u, out->v, out->flags};
  for (int k = 0; k < 12; ++k)
    if (n && !cuda_ok(ctx, cudaMemcpyAsync(dp[k], sp[k], (size_t)n * 4, cudaMemcpyDeviceToDevice, st), "copy stream")) return PHOS_ERR_CUDA;
  if (!cuda_ok(ctx, cudaStreamSynchronize(st), "wavefront_rays")) return PHOS_ERR_CUDA;
  *out_count = n;
  return PHOS_OK;
}

int phos_cuda_film_clear(phos_ctx* ctx) {
  if (!ctx || !ctx->render || !ctx->render->film) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t bytes = (size_t)ctx->render->camera.width * ctx->render->camera.height * 4 * sizeof(float);
  return cuda_ok(ctx, cudaMemsetAsync(ctx->render->film, 0, bytes, ctx->stream), "film_clear") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_film_device_ptr(phos_ctx* ctx, void** out_ptr, uint64_t* out_floats) {
  if (!ctx || !ctx->render || !ctx->render->film || !out_ptr) return PHOS_ERR_INVALID;
  *out_ptr = ctx->render->film;
  if (out_floats) *out_floats = (uint64_t)ctx->render->camera.width * ctx->render->camera.height * 4;
  return PHOS_OK;
}

int phos_cuda_enable_normals(phos_ctx* ctx, int on) {
  if (!ctx || !ctx->render || !ctx->render->film) return fail(ctx, PHOS_ERR_INVALID, "enable_normals before upload_scene");
  cudaSetDevice(ctx->device);
  RenderState& R = *ctx->render;
  cudaStreamSynchronize(ctx->stream);
  if (!on) {
    if (R.film_normals) cudaFree(R.film_normals);
    R.film_normals = nullptr;
    return PHOS_OK;
  }
  const size_t bytes = (size_t)R.camera.width * R.camera.height * 3 * sizeof(float);
  if (!R.film_normals && !cuda_ok(ctx, cudaMalloc(&R.film_normals, bytes), "cudaMalloc(normals channel)")) return PHOS_ERR_CUDA;
  return cuda_ok(ctx, cudaMemset(R.film_normals, 0, bytes), "clear normals channel") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_film_read_normals(phos_ctx* ctx, float* xyz, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
  if (!ctx || !ctx->render || !xyz) return PHOS_ERR_INVALID;
  if (!ctx->render->film_normals) return fail(ctx, PHOS_ERR_INVALID, "normals channel not enabled");
  const DevCamera& c = ctx->render->camera;
  if (x + w > c.width || y + h > c.height) return fail(ctx, PHOS_ERR_INVALID, "rectangle outside the film");
  if (w == 0 || h == 0) return PHOS_OK;
  cudaSetDevice(ctx->device);
  const float* src = ctx->render->film_normals + 3 * ((size_t)y * c.width + x);
  if (!cuda_ok(ctx,
               cudaMemcpy2DAsync(xyz, (size_t)w * 12, src, (size_t)c.width * 12, (size_t)w * 12, h, cudaMemcpyDeviceToHost, ctx->stream),
               "film_read_normals") ||
      !cuda_ok(ctx, cudaStreamSynchronize(ctx->stream), "film_read_normals sync"))
    return PHOS_ERR_CUDA;
  return PHOS_OK;
}

int phos_cuda_film_read(phos_ctx* ctx, float* rgba, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
  if (!ctx || !ctx->render || !ctx->render->film || !rgba) return PHOS_ERR_INVALID;
  const DevCamera& c = ctx->render->camera;
  if (x + w > c.width || y + h > c.height) return fail(ctx, PHOS_ERR_INVALID, "rectangle outside the film");
  if (w == 0 || h == 0) return PHOS_OK;
  cudaSetDevice(ctx->device);
  const float* src = ctx->render->film + 4 * ((size_t)y * c.width + x);
  if (!cuda_ok(ctx,
               cudaMemcpy2DAsync(rgba, (size_t)w * 16, src, (size_t)c.width * 16, (size_t)w * 16, h, cudaMemcpyDeviceToHost, ctx->stream),
               "film_read") ||
      !cuda_ok(ctx, cudaStreamSynchronize(ctx->stream), "film_read sync"))
    return PHOS_ERR_CUDA;
  return PHOS_OK;
}

}  // extern "C"
