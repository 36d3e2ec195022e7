// integrate.cuh — device math of the wavefront path tracer: counter-based random numbers, tangent
// frames, the diffuse and GGX-reflection lobes, area-light sampling and shading normals (sm_100a).
//
// Re-states, for the renderer's built-in closure subset, what the reference evaluates per slot in
//   deferred_shading_kernel_t::build_interactions (src/kernels/cpu/deferred_shading_kernel.hpp:39-72),
//   mesh_t::shading_parameters (src/mesh.cpp:169-258),
//   spt::light_sampler_t / integrator_t (src/kernels/cpu/spt.hpp:95-328),
//   bsdf_t::f / sample (src/bsdf.cpp:113-248), lambert.hpp, microfacet.hpp:174-435, params.hpp:86-99,
//   orthogonal_base.hpp, math/sampling.hpp:23-36, math/fresnel.hpp.
// Expression shapes (float vs double sub-expressions, operation order) follow the reference; this
// translation unit is compiled with -fmad=false so nothing is contracted.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

// The lobe functions are called from three places each (bsdf_f, the sampled lobe and the matched-lobe sum of
// bsdf_sample); inlined everywhere the integrator grows to 11.6 k instructions (186 KB) and stalls on instruction
// fetch (ncu: no_instruction 3.9 per issue).  Out of line it is a third of that.
#ifndef PHOS_LOBE_INLINE
#define PHOS_LOBE_FN static __device__ __noinline__
#else
#define PHOS_LOBE_FN __device__ __forceinline__
#endif

#include <cfloat>
#include <cstdint>

namespace phos {

#define PHOS_PI 3.14159265358979323846     /* M_PI  */
#define PHOS_1_PI 0.31830988618379067154   /* M_1_PI */

struct v3 {
  float x, y, z;
};
__device__ __forceinline__ v3 V(float x, float y, float z) { return v3{x, y, z}; }
__device__ __forceinline__ v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ v3 scl(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ v3 cross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float len(v3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ v3 normalized(v3 a) {  // Imath normalize(): divide by the length unless it is 0
  const float l = len(a);
  return l != 0.0f ? V(a.x / l, a.y / l, a.z / l) : a;
}

// ---- counter-based random numbers (DESIGN.md "random numbers"): a pure function of
// (seed, film pixel, sample, bounce, dimension), so an image does not depend on how pixels or samples
// are partitioned over tiles, streams or GPUs.  The reference's sampler is one shared sequential
// std::mt19937 (src/sampling.cpp:43-76) and cannot be reproduced draw for draw.
__host__ __device__ __forceinline__ uint32_t rng_mix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ float rng(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t dim) {
  uint32_t h = rng_mix(seed + 0x9E3779B9u * (pixel + 1u));
  h = rng_mix(h ^ (0x85EBCA6Bu * (sample + 1u)));
  h = rng_mix(h ^ (0xC2B2AE35u * (bounce * 8u + dim + 1u)));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}
enum { DIM_LIGHT = 0, DIM_LIGHT_U = 1, DIM_LIGHT_V = 2, DIM_BSDF_U = 3, DIM_BSDF_V = 4, DIM_RR = 5, DIM_LENS = 6, DIM_FILM = 7 };

// ---- scene tables in HBM -------------------------------------------------------------------------------
// One lobe of bsdf_t after add_lobe + precompute (bsdf.hpp:52-83, bsdf/params.hpp): type = bsdf_t::type_t
// (= PHOS_LOBE_*), flags = bsdf::REFLECT / TRANSMIT / SPECULAR ..., p0 / p1 = GGX alpha_x / alpha_y, Oren-Nayar
// a / b, eta (reflection, refraction), sheen r.
enum { BSDF_DIFFUSE_F = 1, BSDF_GLOSSY_F = 2, BSDF_SPECULAR_F = 4, BSDF_REFLECT_F = 8, BSDF_TRANSMIT_F = 16 };
struct DevLobe {
  uint32_t type, flags;
  float w[3];
  float p0, p1, p2;  // p2: eta of the GGX transmission lobe
};
// What material_t::evaluate leaves in shading_result_t for a hit on this material: the closure list
// eval_closure builds (material.cpp:218-305) and the emission.
struct DevMaterial {
  uint32_t kind;    // PHOS_MAT_*
  uint32_t nlobes;  // 0 for emitters / the background
  float e[3];       // (power / pi) * Cs (emitter), Cs * power (background)
  DevLobe lobes[8];
};

struct DevScene {
  const float* verts;
  const float* normals;  // may be null
  const uint32_t* faces;
  const uint32_t* vert_offset;
  const uint32_t* face_offset;
  const uint8_t* mesh_smooth;  // 0 flat, 1 smooth, 2 mixed: per face in face_smooth
  const uint8_t* face_smooth;  // per face (scene-wide face index); null when no mesh mixes smooth and flat faces
  const DevMaterial* mats;
  int32_t environment;  // material id of the environment or -1 (scene_t::environment())
  uint32_t nlights;
  const uint32_t* light_first;     // [nlights + 1]
  const float* light_area;         // [nlights]
  const uint32_t* light_tri_mesh;  // meshid | matid << 16
  const uint32_t* light_tri_face;  // 3 * face index
};

__device__ __forceinline__ v3 scene_vert(const DevScene& S, uint32_t mesh, uint32_t face3, int k) {
  const size_t vi = (size_t)__ldg(S.faces + 3 * (size_t)__ldg(S.face_offset + mesh) + face3 + k) + __ldg(S.vert_offset + mesh);
  return V(__ldg(S.verts + 3 * vi), __ldg(S.verts + 3 * vi + 1), __ldg(S.verts + 3 * vi + 2));
}

// mesh_t::shading_parameters, mesh.cpp:169-206: interpolated vertex normal (w on a, u on b, v on c)
// for smooth meshes, geometric normal (v1 - v0) x (v2 - v0) otherwise; never face-forwarded (:209-215)
__device__ __forceinline__ v3 shading_normal(const DevScene& S, uint32_t mesh, uint32_t face3, float u, float v) {
  uint32_t smooth = __ldg(S.mesh_smooth + mesh);
  if (smooth == 2u) smooth = __ldg(S.face_smooth + (size_t)__ldg(S.face_offset + mesh) + face3 / 3u);  // the face's own flag
  if (smooth && S.normals) {
    const float w = 1 - u - v;
    const size_t f = 3 * (size_t)__ldg(S.face_offset + mesh) + face3;
    const size_t vo = __ldg(S.vert_offset + mesh);
    const size_t ia = __ldg(S.faces + f) + vo, ib = __ldg(S.faces + f + 1) + vo, ic = __ldg(S.faces + f + 2) + vo;
    const v3 n0 = V(S.normals[3 * ia], S.normals[3 * ia + 1], S.normals[3 * ia + 2]);
    const v3 n1 = V(S.normals[3 * ib], S.normals[3 * ib + 1], S.normals[3 * ib + 2]);
    const v3 n2 = V(S.normals[3 * ic], S.normals[3 * ic + 1], S.normals[3 * ic + 2]);
    return normalized(add(add(scl(n0, w), scl(n1, u)), scl(n2, v)));
  }
  const v3 v0 = scene_vert(S, mesh, face3, 0), v1 = scene_vert(S, mesh, face3, 1), v2 = scene_vert(S, mesh, face3, 2);
  return normalized(cross(sub(v1, v0), sub(v2, v0)));
}

// ---- frames (orthogonal_base.hpp:11-69) -----------------------------------------------------------------
struct Base {
  v3 a, b, c;
};
__device__ __forceinline__ Base make_base(v3 n) {
  Base r;
  const v3 t = (n.x != n.y || n.x != n.z) ? V(n.z - n.y, n.x - n.z, n.y - n.x) : V(n.z - n.y, n.x + n.z, -n.y - n.x);
  r.a = normalized(t);
  r.b = n;
  r.c = normalized(cross(r.a, n));
  return r;
}
__device__ __forceinline__ v3 to_world(const Base& b, v3 v) { return add(add(scl(b.a, v.x), scl(b.b, v.y)), scl(b.c, v.z)); }
__device__ __forceinline__ v3 to_local(const Base& b, v3 v) {
  const v3 ia = V(b.a.x, b.b.x, b.c.x), ib = V(b.a.y, b.b.y, b.c.y), ic = V(b.a.z, b.b.z, b.c.z);
  return add(add(scl(ia, v.x), scl(ib, v.y)), scl(ic, v.z));
}

// ---- tangent-space trigonometry (math/vector.hpp:24-71), y is up ------------------------------------------
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }
__device__ __forceinline__ float cos2_theta(v3 v) { return v.y * v.y; }
__device__ __forceinline__ float sin2_theta(v3 v) { return fmaxf(0.0f, 1.0f - cos2_theta(v)); }
__device__ __forceinline__ float sin_theta(v3 v) { return sqrtf(sin2_theta(v)); }
__device__ __forceinline__ float tan_theta(v3 v) { return sin_theta(v) / v.y; }
__device__ __forceinline__ float tan2_theta(v3 v) { return sin2_theta(v) / cos2_theta(v); }
__device__ __forceinline__ float cos_phi(v3 v) {
  const float s = sin_theta(v);
  return s == 0 ? 1 : clampf(v.x / s, -1.f, 1.f);
}
__device__ __forceinline__ float sin_phi(v3 v) {
  const float s = sin_theta(v);
  return s == 0 ? 0 : clampf(v.z / s, -1.f, 1.f);
}

// fresnel::dielectric, math/fresnel.hpp:6-28
__device__ __forceinline__ float fresnel_dielectric(float cosi, float eta) {
  if (eta == 0) return 1;
  if (cosi < 0.0f) eta = 1.0f / eta;
  const float c = fabsf(cosi);
  float g = eta * eta - 1.0f + c * c;
  if (g > 0.0f) {
    g = sqrtf(g);
    const float A = (g - c) / (g + c);
    const float B = (c * (g + c) - 1.0f) / (c * (g - c) + 1.0f);
    return 0.5f * A * A * (1 + B * B);
  }
  return 1.0f;
}

// ---- GGX (microfacet.hpp:306-435) ------------------------------------------------------------------------------
__device__ __forceinline__ float ggx_D(float ax, float ay, v3 v) {
  const float t2 = tan2_theta(v);
  if (isinf(t2)) return 0.0f;
  const float c2 = cos2_theta(v);
  const float c4 = c2 * c2;
  const float cp = cos_phi(v), sp = sin_phi(v);
  const float e = (cp * cp / (ax * ax) + sp * sp / (ay * ay)) * t2;
  return (float)(1.0f / (PHOS_PI * ax * ay * c4 * (1 + e) * (1 + e)));  // M_PI makes the denominator double
}
__device__ __forceinline__ float ggx_Lambda(float ax, float ay, v3 v) {
  const float att = fabsf(tan_theta(v));
  if (isinf(att)) return 0.0f;
  const float cp = cos_phi(v), sp = sin_phi(v);
  const float alpha = sqrtf(cp * cp * ax * ay + sp * sp * ax * ay);
  const float a2t2 = (alpha * att) * (alpha * att);
  return (-1.0f + sqrtf(1.0f + a2t2)) * 0.5f;
}
__device__ __forceinline__ float ggx_G1(float ax, float ay, v3 v) { return 1.0f / (1.0f + ggx_Lambda(ax, ay, v)); }
__device__ __forceinline__ void ggx_sample_slope(float cos_theta, float& slope_x, float& slope_y, float u, float v) {
  if (cos_theta > .9999) {
    const float r = sqrtf(u / (1 - u));
    const float phi = (float)(6.28318530718 * v);
    slope_x = r * cosf(phi);
    slope_y = r * sinf(phi);
    return;
  }
  const float sin_t = sqrtf(fmaxf(0.0f, 1.0f - (cos_theta * cos_theta)));
  const float tan_t = sin_t / cos_theta;
  const float a = 1.0f / tan_t;
  const float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
  const float A = 2.0f * u / g1 - 1.0f;
  float tmp = 1.0f / (A * A - 1.0f);
  if (tmp > 1e10) tmp = 1e10;
  const float B = tan_t;
  const float D = sqrtf(fmaxf((float)(B * B * tmp * tmp - (A * A - B * B) * tmp), 0.0f));
  const float slope_x1 = B * tmp - D;
  const float slope_x2 = B * tmp + D;
  slope_x = (A < 0.0f || slope_x2 > 1.0f / tan_t) ? slope_x1 : slope_x2;
  float S;
  if (v > 0.5f) {
    S = 1.0f;
    v = 2.0f * (v - 0.5f);
  } else {
    S = -1.0f;
    v = 2.0f * (0.5f - v);
  }
  const float z =
      (v * (v * (v * 0.27385f - 0.73369f) + 0.46341f)) / (v * (v * (v * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
  slope_y = S * z * sqrtf(1.0f + slope_x * slope_x);
}
__device__ __forceinline__ v3 ggx_sample(float ax, float ay, v3 wi, float& pdf, float u, float v) {
  const v3 stretched = normalized(V(ax * wi.x, wi.y, ay * wi.z));
  float slope_x, slope_y;
  ggx_sample_slope(stretched.y, slope_x, slope_y, u, v);
  const float tmp = cos_phi(stretched) * slope_x - sin_phi(stretched) * slope_y;
  slope_y = sin_phi(stretched) * slope_x + cos_phi(stretched) * slope_y;
  slope_x = tmp;
  slope_x = slope_x * ax;
  slope_y = slope_y * ay;
  const v3 wh = normalized(V(-slope_x, 1.0f, -slope_y));
  pdf = (ggx_D(ax, ay, wh) * ggx_G1(ax, ay, wi) * fabsf(dot(wi, wh)) / fabsf(wi.y));
  return wh;
}
// cook_torrance::f, microfacet.hpp:174-215 (Fresnel eta hard-wired to 0.5, :209)
PHOS_LOBE_FN float ct_f(v3 n, float ax, float ay, v3 wi, v3 wo) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi), lo = to_local(base, wo);
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  v3 wh = add(li, lo);
  const float cos_ti = fabsf(li.y), cos_to = fabsf(lo.y);
  if (cos_ti == 0 || cos_to == 0) return 0.0f;
  if (wh.x == 0 || wh.y == 0 || wh.z == 0) return 0.0f;
  wh = normalized(wh);
  const float d = ggx_D(ax, ay, wh);
  const float g = 1.0f / (1.0f + ggx_Lambda(ax, ay, li) + ggx_Lambda(ax, ay, lo));
  const float whu = (float)(wh.x * 0.0f + wh.y * 1.0 + wh.z * 0.0f);
  const float f = fresnel_dielectric(dot(lo, whu < 0.0f ? neg(wh) : wh), 0.5f);
  return d * g * f * (1.0f / (4.0f * cos_ti * cos_to));
}
// cook_torrance::sample, microfacet.hpp:238-277; returns 0 (black) on the early-outs
PHOS_LOBE_FN float ct_sample(v3 n, float ax, float ay, v3 wi, v3& wo, float u, float v, float& opdf) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi);
  if (li.y == 0.0f) return 0.0f;
  float dpdf;
  const v3 wh = ggx_sample(ax, ay, li, dpdf, u, v);
  if (dot(li, wh) < 0.0f) return 0.0f;
  const v3 lo = add(neg(li), scl(wh, 2.0f * dot(li, wh)));
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  opdf = dpdf / (4.0f * dot(li, wh));
  wo = to_world(base, lo);
  return ct_f(n, ax, ay, wi, wo);
}


// ---- the other lobes and bsdf_t itself -----------------------------------------------------------------------------
// oren_nayar::f, bsdf/oren_nayar.hpp:9-47
PHOS_LOBE_FN float oren_nayar_f(v3 n, float a, float b, v3 wi, v3 wo) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi), lo = to_local(base, wo);
  const float cos_theta_i = fabsf(li.y), cos_theta_o = fabsf(lo.y);
  const float sin_theta_i = sin_theta(li), sin_theta_o = sin_theta(lo);
  float max_cos = 0.0f;
  if (sin_theta_i > 0.0001f && sin_theta_o > 0.0001f) {
    const float sin_phi_i = sin_phi(li), cos_phi_i = cos_phi(li);
    const float sin_phi_o = sin_phi(lo), cos_phi_o = cos_phi(lo);
    const float dcos = cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o;
    max_cos = fmaxf(0.0f, dcos);
  }
  float sin_alpha, tan_beta;
  if (cos_theta_i > cos_theta_o) {
    sin_alpha = sin_theta_o;
    tan_beta = sin_theta_i / cos_theta_i;
  } else {
    sin_alpha = sin_theta_i;
    tan_beta = sin_theta_o / cos_theta_o;
  }
  const float result = (a + b * max_cos * sin_alpha * tan_beta);
  return (float)(result * PHOS_1_PI);
}
// microfacet::sheen, bsdf/sheen.hpp:15-66.  The reference keeps L(0.5, r) in a function-local static initialised by
// the first sheen lobe ever evaluated (sheen.hpp:56); here it is that of the lobe at hand — identical as long as
// a scene uses one sheen roughness.
__device__ __forceinline__ float sheen_L(float x, float r) {
  const float t = (1.0f - r) * (1.0f - r);
  const float a = t * 25.3245f + (1.0f - t) * 21.5473f;
  const float b = t * 3.32435f + (1.0f - t) * 3.82987f;
  const float c = t * 0.16801f + (1.0f - t) * 0.19823f;
  const float d = t * -1.27393f + (1.0f - t) * -1.97760f;
  const float e = t * -4.85967f + (1.0f - t) * -4.32054f;
  const float xc = powf(x, c);
  return a / (1 + b * xc) + d * x + e;
}
__device__ __forceinline__ float sheen_D(float r, v3 v) {
  const float st = sin_theta(v);
  const float oor = 1.0f / r;
  return (float)((2.0f + oor) * powf(st, oor) / (2.0f * PHOS_PI));
}
__device__ __forceinline__ float sheen_Lambda(float r, v3 v) {
  const float L5 = sheen_L(0.5f, r);
  const float ct = v.y;
  const float l = (ct < 0.5f) ? sheen_L(ct, r) : 2.0f * L5 - sheen_L(1.0f - ct, r);
  return expf(l);
}
// cook_torrance::f with the sheen distribution (bsdf.cpp:88-96, microfacet.hpp:174-215)
PHOS_LOBE_FN float sheen_f(v3 n, float r, v3 wi, v3 wo) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi), lo = to_local(base, wo);
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  v3 wh = add(li, lo);
  const float cos_ti = fabsf(li.y), cos_to = fabsf(lo.y);
  if (cos_ti == 0 || cos_to == 0) return 0.0f;
  if (wh.x == 0 || wh.y == 0 || wh.z == 0) return 0.0f;
  wh = normalized(wh);
  const float d = sheen_D(r, wh);
  const float g = 1.0f / (1.0f + sheen_Lambda(r, li) + sheen_Lambda(r, lo));
  const float whu = (float)(wh.x * 0.0f + wh.y * 1.0 + wh.z * 0.0f);
  const float f = fresnel_dielectric(dot(lo, whu < 0.0f ? neg(wh) : wh), 0.5f);
  return d * g * f * (1.0f / (4.0f * cos_ti * cos_to));
}
// cook_torrance::pdf, microfacet.hpp:217-236 — G1 is handed the WORLD-space wi there, as here
PHOS_LOBE_FN float ct_pdf(v3 n, float ax, float ay, v3 wi, v3 wo) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi), lo = to_local(base, wo);
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  const v3 wh = normalized(add(li, lo));
  return (ggx_D(ax, ay, wh) * ggx_G1(ax, ay, wi) * fabsf(dot(li, wh)) / fabsf(li.y)) / (4.0f * dot(li, wh));
}
// cook_torrance::refract::{f, pdf, sample}, microfacet.hpp:36-172 (GGX transmission), to the letter: pdf divides by
// sqrt_denom and multiplies by it again (:113), and tests the hemisphere on the WORLD-space vectors (:108)
PHOS_LOBE_FN float ctr_f(v3 n, float ax, float ay, float peta, v3 wi, v3 wo) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi), lo = to_local(base, wo);
  if ((li.y * lo.y) > 0.0f) return 0.0f;
  const float eta = li.y > 0.0f ? peta : 1.0f / peta;
  const float cos_ti = li.y, cos_to = lo.y;
  if (cos_ti == 0.0f || cos_to == 0.0f) return 0.0f;
  v3 wh = normalized(add(li, scl(lo, eta)));
  if (wh.y < 0) wh = neg(wh);
  if (dot(lo, wh) * dot(li, wh) > 0) return 0.0f;
  const float f = fresnel_dielectric(dot(lo, wh), eta);
  const float sqrt_denom = dot(li, wh) + eta * dot(lo, wh);
  const float factor = 1.0f / eta;
  const float d = ggx_D(ax, ay, wh);
  const float g = 1.0f / (1.0f + ggx_Lambda(ax, ay, li) + ggx_Lambda(ax, ay, lo));
  return (1.0f - f) * fabsf(d * g * eta * eta * fabsf(dot(lo, wh)) * fabsf(dot(li, wh)) * factor * factor /
                            (cos_ti * cos_to * sqrt_denom * sqrt_denom));
}
PHOS_LOBE_FN float ctr_pdf(v3 n, float ax, float ay, float peta, v3 wi, v3 wo) {
  const Base base = make_base(n);
  const v3 li = to_local(base, wi), lo = to_local(base, wo);
  const float eta = li.y > 0.0f ? peta : 1.0f / peta;
  if ((double)dot(wo, wi) > 0.0) return 0;
  const v3 wh = normalized(add(li, scl(lo, eta)));
  const float sqrt_denom = dot(li, wh) + eta * dot(lo, wh);
  const float dwh_dwi = fabsf(eta * eta * dot(lo, wh)) / sqrt_denom * sqrt_denom;
  return (ggx_D(ax, ay, wh) * wh.y) * dwh_dwi;
}
PHOS_LOBE_FN float ctr_sample(v3 n, float ax, float ay, float peta, v3 wi, v3& wo, float u, float v, float& opdf) {
  if (peta == 1.0f) {
    wo = neg(wi);
    opdf = 1.0f;
    return 1.0f;
  }
  const Base base = make_base(n);
  const v3 li = to_local(base, wi);
  if (li.y == 0.0f) return 0.0f;
  float dpdf;
  const v3 wh = ggx_sample(ax, ay, li, dpdf, u, v);
  if (dot(wh, li) < 0.0f) return 0.0f;
  const float eta = li.y > 0.0f ? 1.0f / peta : peta;
  const float cos_ti = dot(wh, li);
  const float sin2_ti = fmaxf(0.0f, 1.0f - cos_ti * cos_ti);
  const float sin2_tt = eta * eta * sin2_ti;
  if (sin2_tt >= 1.0f) return 0.0f;
  const float cos_tt = sqrtf(1.0f - sin2_tt);
  const v3 lo = add(scl(neg(li), eta), scl(wh, eta * cos_ti - cos_tt));
  const float sqrt_denom = dot(li, wh) + eta * dot(lo, wh);
  const float dwh_dwi = fabsf((eta * eta * dot(lo, wh)) / (sqrt_denom * sqrt_denom));
  opdf = dpdf * dwh_dwi;
  wo = to_world(base, lo);
  return ctr_f(n, ax, ay, peta, wi, wo);
}
// sample::hemisphere::cosine_weighted + orthogonal_base_t::to_world (math/sampling.hpp:23-36, lambert.hpp:24-36)
PHOS_LOBE_FN v3 cosine_sample(v3 n, float sx, float sy, float& pdf) {
  const Base base = make_base(n);
  const float rr = sqrtf(sx);
  const float theta = (float)(2 * PHOS_PI * sy);
  const float x = rr * cosf(theta), y = rr * sinf(theta);
  const v3 lo = V(x, sqrtf(fmaxf(0.0f, 1.0f - sx)), y);
  pdf = lo.y * (float)(1.0f / PHOS_PI);
  return to_world(base, lo);
}
// eval(), bsdf.cpp:25-107: grey value of one lobe and its pdf
__device__ __forceinline__ float lobe_eval(const DevLobe& l, v3 n, v3 wi, v3 wo, float& pdf) {
  switch (l.type) {
    case 1: pdf = (float)(dot(n, wi) * PHOS_1_PI); return (float)PHOS_1_PI;                         // Diffuse
    case 2: pdf = (float)(dot(n, wi) * PHOS_1_PI); return oren_nayar_f(n, l.p0, l.p1, wi, wo);     // OrenNayar
    case 16: pdf = ct_pdf(n, l.p0, l.p1, wi, wo); return ct_f(n, l.p0, l.p1, wi, wo);              // Microfacet (GGX)
    case 32: pdf = (float)(dot(n, wi) * PHOS_1_PI); return sheen_f(n, l.p0, wi, wo);               // Sheen
    case 272: pdf = ctr_pdf(n, l.p0, l.p1, l.p2, wi, wo); return ctr_f(n, l.p0, l.p1, l.p2, wi, wo);  // Microfacet, refract = 1
    default: pdf = 0.0f; return 0.0f;                                                               // Reflection, Refraction, Transparent
  }
}
// bsdf_t::f, bsdf.cpp:113-131
__device__ __forceinline__ v3 bsdf_f(const DevMaterial* __restrict__ m, v3 n, v3 wi, v3 wo) {
  v3 out = V(0, 0, 0);
  const uint32_t nl = m->nlobes;
  for (uint32_t i = 0; i < nl; ++i) {
    const DevLobe l = m->lobes[i];
    float ignored;
    const float e = lobe_eval(l, n, wi, wo, ignored);
    const float atl = dot(n, wi);
    const bool reflect = atl * dot(n, wo) > 0.0f;
    if ((reflect && (l.flags & BSDF_REFLECT_F)) || (!reflect && (l.flags & BSDF_TRANSMIT_F)))
      out = add(out, scl(mul(V(e, e, e), V(l.w[0], l.w[1], l.w[2])), atl));
  }
  return out;
}
// bsdf_t::sample, bsdf.cpp:133-248.  false when the path ends (a lobe that bails out before setting its pdf
// leaves the reference with an uninitialised one — frozen to "path ends"; the same for total internal reflection,
// where refraction::sample returns an uninitialised colour, refraction.hpp:45).
__device__ __forceinline__ bool bsdf_sample(const DevMaterial* __restrict__ m, v3 n, float sx, float sy, v3 wi, v3& wo, v3& f,
                                            float& opdf, uint32_t& oflags) {
  const uint32_t nl = m->nlobes;
  if (nl == 0) return false;
  const uint32_t index = min((uint32_t)floorf(sx * nl), nl - 1);
  const float u = fminf(sx * nl - index, 1.0f - FLT_EPSILON);
  const DevLobe l = m->lobes[index];
  float pdf = 0.0f, r = 0.0f;
  switch (l.type) {
    case 1: wo = cosine_sample(n, u, sy, pdf); r = (float)PHOS_1_PI; break;
    case 2: wo = cosine_sample(n, u, sy, pdf); r = oren_nayar_f(n, l.p0, l.p1, wi, wo); break;
    case 16:
      r = ct_sample(n, l.p0, l.p1, wi, wo, u, sy, pdf);
      if (r == 0.0f) return false;
      break;
    case 32: wo = cosine_sample(n, u, sy, pdf); r = sheen_f(n, l.p0, wi, wo); break;
    case 272:
      r = ctr_sample(n, l.p0, l.p1, l.p2, wi, wo, u, sy, pdf);
      if (r == 0.0f) return false;
      break;
    case 4: {  // reflection.hpp:8-21
      const float ct = dot(n, wi);
      pdf = 1.0f;
      wo = add(neg(wi), scl(n, 2.0f * ct));
      r = 1.0f;
      break;
    }
    case 8: {  // refraction.hpp:10-46
      pdf = 1.0f;
      float ct = dot(n, wi);
      const float st = fmaxf(0.0f, 1.0f - ct * ct);
      v3 nn;
      float eta = l.p0;
      if (ct > 0) {
        nn = n;
        eta = 1.0f / eta;
      } else {
        nn = neg(n);
        ct = -ct;
      }
      const float arg = 1.0f - (eta * eta * st);
      if (!(arg >= 0.0f)) return false;
      const float dnp = sqrtf(arg);
      const float nk = eta * ct - dnp;
      wo = add(scl(neg(wi), eta), scl(nn, nk));
      r = 1.0f;
      break;
    }
    case 128: wo = neg(wi); pdf = 1.0f; r = 1.0f; break;  // Transparent
    default: return false;
  }
  if (pdf == 0.0f) return false;
  v3 result = mul(V(r, r, r), V(l.w[0], l.w[1], l.w[2]));
  int matched = 1;
  for (uint32_t i = 0; i < nl; ++i) {
    if (i == index) continue;
    const DevLobe o = m->lobes[i];
    if ((l.flags & o.flags) != o.flags) continue;
    const bool reflect = dot(n, wi) * dot(n, wo) > 0.0f;
    if ((reflect && (o.flags & BSDF_REFLECT_F)) || (!reflect && (o.flags & BSDF_TRANSMIT_F))) {
      float lobe_pdf = 0.0f;
      const float e = lobe_eval(o, n, wi, wo, lobe_pdf);
      result = add(result, mul(V(e, e, e), V(o.w[0], o.w[1], o.w[2])));
      pdf += lobe_pdf;
      ++matched;
    }
  }
  pdf /= matched;
  f = result;
  opdf = pdf;
  oflags = l.flags;
  return true;
}

}  // namespace phos
