/* phos_oracle_render.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar CPU restatement of the reference's path-tracing pipeline for the renderer's
 * built-in closure subset: deferred shading, next-event estimation, the integrator and the film
 * hand-off.  It is the checker for the CUDA wavefront integrator; only tests/, smoke() and the
 * bench's CPU-baseline leg may load it.
 *
 * Follows (reference file:line):
 *   tile pipeline ........ tile_renderer_t::render_tile / trace_rays, src/xpu/cpu.cpp:148-205
 *   interactions ......... deferred_shading_kernel_t::build_interactions, deferred_shading_kernel.hpp:39-72
 *   shading normal ....... mesh_t::shading_parameters, src/mesh.cpp:169-258
 *   closures ............. diffuse_bsdf_node.osl:20-25, glossy_bsdf_node.osl:26-34, diffuse_emitter_node.osl:18
 *                          through material.cpp:218-305 (here: oracle/ref_shim/material_stub.cpp's table)
 *   NEE .................. spt::light_sampler_t, spt.hpp:95-149; sampler_t::fresh_light_samples,
 *                          sampling.cpp:160-180; area_light_t, light.cpp:30-71; triangle_t::sample, mesh.cpp:314-324
 *   integrator ........... spt::integrator_t, spt.hpp:161-328
 *   BSDF ................. bsdf_t::f / sample, bsdf.cpp:113-248; lambert.hpp; cook_torrance + ggx_t,
 *                          microfacet.hpp:174-435; params.hpp:86-99; orthogonal_base.hpp; math/sampling.hpp:23-36
 *
 * Stated differences from the linked reference renderer (DESIGN.md "image parity"):
 *   1. random numbers: the reference draws from one shared, sequential std::mt19937
 *      (sampling.cpp:43-76, order- and thread-dependent).  Here every draw is a pure function of
 *      (seed, pixel, sample, bounce, dimension) — orc_rnd below — which is what the GPU uses too;
 *   2. vector3_t<8>::normalize (simd/vector.hpp:126-133) multiplies by the ~12-bit _mm256_rcp_ps;
 *      with rcp_mode = 0 the shadow-ray and camera-ray directions are normalised exactly.  With
 *      rcp_mode = 1 (x86 only) the same RCPSS instruction is used, which reproduces the reference's
 *      systematic shadow-ray overshoot (a shadow ray whose direction is > 1e-4 / length too long
 *      hits the light's own triangle and is counted as occluded), and rays are traced with the
 *      reference's approximate slab test; rcp_mode = 2 keeps the RCPSS normalisation but traces
 *      exactly — what the device does under phos_cuda_reference_normalize;
 *   3. a 0-lobe BSDF (pure emitter) is sampled by the reference through uninitialised lobe data
 *      (bsdf.cpp:140-153); here the path simply ends there.
 * All arithmetic keeps the reference's float / double expression shapes; compile with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../include/phos_scene.h"

/* ---- pieces of phos_oracle.c reused here ------------------------------------------------------------ */
typedef struct {
  float *px, *py, *pz, *wx, *wy, *wz, *d;
  uint32_t *mesh, *face;
  float *u, *v;
  uint32_t* flags;
} orc_rays;
void orc_traverse(const void* nodes, const void* packets, orc_rays* rays, uint64_t n, void* counters, int mode_ties);

#define ORC_HIT 1u
#define ORC_MASKED 2u
#define ORC_SHADOW 4u
#define ORC_SPECULAR 8u

typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scl(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* Imath Vec3::dot */
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float len(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 normalized(v3 a) { /* Imath normalize(): divide by the length unless it is 0 */
  const float l = len(a);
  return l != 0.0f ? V(a.x / l, a.y / l, a.z / l) : a;
}

/* ---- counter-based random numbers shared with the GPU (DESIGN.md "random numbers") ------------------ */
static inline uint32_t orc_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
float orc_rnd(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t dim) {
  uint32_t h = orc_mix(seed + 0x9E3779B9u * (pixel + 1u));
  h = orc_mix(h ^ (0x85EBCA6Bu * (sample + 1u)));
  h = orc_mix(h ^ (0xC2B2AE35u * (bounce * 8u + dim + 1u)));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}
enum { DIM_LIGHT = 0, DIM_LIGHT_U = 1, DIM_LIGHT_V = 2, DIM_BSDF_U = 3, DIM_BSDF_V = 4, DIM_RR = 5, DIM_LENS = 6, DIM_FILM = 7 };

/* film jitter per sample index: sample::stratified_2d (math/sampling.hpp:67-82) over
 * spd = lround(sqrt(spp)) strata per axis, as sampler_t::preprocess calls it (sampling.cpp:96-99).
 * Entries >= spd^2 (spp not a perfect square) are uninitialised in the reference; 0.5 here. */
void orc_film_jitter(uint32_t seed, uint32_t spp, float* jx, float* jy) {
  const uint32_t num = (uint32_t)lroundf(sqrtf((float)spp));
  for (uint32_t i = 0; i < spp; ++i) jx[i] = jy[i] = 0.5f;
  const float step = 1.0f / (float)num;
  float dy = 0.0f;
  uint32_t call = 0;
  for (uint32_t i = 0; i < num; ++i, dy += step) {
    float dx = 0.0f;
    for (uint32_t j = 0; j < num; ++j, dx += step) {
      const float a = orc_rnd(seed, 0xffffffffu, call++, 0, DIM_FILM);
      const float b = orc_rnd(seed, 0xffffffffu, call++, 0, DIM_FILM);
      if (j * num + i < spp) {
        jx[j * num + i] = dx + a * step;
        jy[j * num + i] = dy + b * step;
      }
    }
  }
}

/* ---- derived scene tables ------------------------------------------------------------------------------ */
/* one lobe of bsdf_t after add_lobe + precompute (bsdf.hpp:52-83, bsdf/params.hpp): p0 / p1 = GGX alpha_x / alpha_y,
 * Oren-Nayar a / b, eta (reflection, refraction), sheen r */
enum { BSDF_DIFFUSE_F = 1, BSDF_GLOSSY_F = 2, BSDF_SPECULAR_F = 4, BSDF_REFLECT_F = 8, BSDF_TRANSMIT_F = 16 };
typedef struct { uint32_t type, flags; v3 w; float p0, p1, p2; } lobe1; /* p2: eta of the GGX transmission lobe */
typedef struct bsdf1 { uint32_t n; lobe1 l[PHOS_MAX_LOBES]; } bsdf1;
typedef struct {
  const phos_scene_desc* d;
  uint32_t nlights;
  uint32_t* light_first; /* [nlights+1] into light_tri_* */
  float* light_area;
  uint32_t* light_tri_mesh; /* meshid | matid << 16 */
  uint32_t* light_tri_face; /* 3 * face index */
  v3* emission;             /* per material: (power / pi) * Cs (emitter), Cs * power (background) */
  struct bsdf1* bsdf;       /* per material: the closure list eval_closure would build (material.cpp:218-305) */
} orc_scene;

static inline v3 vert(const phos_scene_desc* d, uint32_t mesh, uint32_t face3, int k) {
  const uint32_t vi = d->faces[3 * (size_t)d->face_offset[mesh] + face3 + k] + d->vert_offset[mesh];
  return V(d->vertices[3 * (size_t)vi], d->vertices[3 * (size_t)vi + 1], d->vertices[3 * (size_t)vi + 2]);
}

/* microfacet_t::roughness_to_alpha + precompute clamp, params.hpp:86-99 */
static float roughness_to_alpha(float roughness) {
  roughness = fmaxf(roughness, (float)1e-5);
  const float x = logf(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

/* bsdf_t::add_lobe + T::precompute for one closure */
static void add_lobe(bsdf1* b, uint32_t type, const float w[3], float param, float param2) {
  if (b->n >= PHOS_MAX_LOBES) return;
  lobe1* l = &b->l[b->n++];
  l->type = type;
  l->w = V(w[0], w[1], w[2]);
  l->p0 = l->p1 = l->p2 = 0.0f;
  switch (type) {
    case PHOS_LOBE_MICROFACET_REFRACT: /* add_lobe<microfacet_t> with refract = 1: flags = TRANSMIT (bsdf.hpp:69-83) */
      l->flags = BSDF_TRANSMIT_F;
      l->p0 = l->p1 = fminf(1.0f, fmaxf(0.0001f, roughness_to_alpha(param)));
      l->p2 = param2;
      break;
    case PHOS_LOBE_DIFFUSE: l->flags = BSDF_REFLECT_F | BSDF_DIFFUSE_F; break;
    case PHOS_LOBE_OREN_NAYAR: { /* oren_nayar_t::precompute, params.hpp:37-42 */
      l->flags = BSDF_REFLECT_F | BSDF_DIFFUSE_F;
      const float sg = (float)(param * (M_PI / 180.0f)); /* trig::radians */
      const float s2 = sg * sg;
      l->p0 = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
      l->p1 = 0.45f * s2 / (s2 + 0.09f);
      break;
    }
    case PHOS_LOBE_REFLECTION: l->flags = BSDF_REFLECT_F | BSDF_SPECULAR_F; l->p0 = param; break;
    case PHOS_LOBE_REFRACTION: l->flags = BSDF_TRANSMIT_F | BSDF_SPECULAR_F; l->p0 = param; break;
    case PHOS_LOBE_MICROFACET: /* add_lobe<microfacet_t>: flags = REFLECT (refract = 0), bsdf.hpp:69-83 */
      l->flags = BSDF_REFLECT_F;
      l->p0 = l->p1 = fminf(1.0f, fmaxf(0.0001f, roughness_to_alpha(param)));
      break;
    case PHOS_LOBE_SHEEN: l->flags = BSDF_REFLECT_F | BSDF_GLOSSY_F; l->p0 = param; break;
    case PHOS_LOBE_TRANSPARENT: l->flags = BSDF_TRANSMIT_F; break; /* empty_params_t, material.cpp:98-103 */
    default: b->n--; break;
  }
}

void* orc_scene_create(const phos_scene_desc* d) {
  orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
  s->d = d;
  s->emission = (v3*)calloc(d->num_materials + 1, sizeof(v3));
  s->bsdf = (bsdf1*)calloc(d->num_materials + 1, sizeof(bsdf1));
  for (uint32_t m = 0; m < d->num_materials; ++m) {
    const phos_material* mt = &d->materials[m];
    const float one[3] = {1.0f * mt->cs[0], 1.0f * mt->cs[1], 1.0f * mt->cs[2]}; /* w * component->w, w = (1,1,1) */
    switch (mt->kind) {
      case PHOS_MAT_DIFFUSE: /* diffuse_bsdf_node.osl:20-25 */
        if (mt->roughness == 0) add_lobe(&s->bsdf[m], PHOS_LOBE_DIFFUSE, one, 0.0f, 0.0f);
        else add_lobe(&s->bsdf[m], PHOS_LOBE_OREN_NAYAR, one, mt->roughness, 0.0f);
        break;
      case PHOS_MAT_GLOSSY: /* glossy_bsdf_node.osl:26-34 */
        if (mt->roughness == 0.0f) add_lobe(&s->bsdf[m], PHOS_LOBE_REFLECTION, one, 0.0f, 0.0f);
        else add_lobe(&s->bsdf[m], PHOS_LOBE_MICROFACET, one, mt->roughness * mt->roughness, 0.0f);
        break;
      case PHOS_MAT_EMITTER: { /* diffuse_emitter_node.osl:18 */
        const float k = (float)(mt->power / M_PI);
        s->emission[m] = V(1.0f * (k * mt->cs[0]), 1.0f * (k * mt->cs[1]), 1.0f * (k * mt->cs[2]));
        break;
      }
      case PHOS_MAT_BACKGROUND: /* background_node.osl: Cs * power * background() */
        s->emission[m] = V(1.0f * (mt->cs[0] * mt->power), 1.0f * (mt->cs[1] * mt->power), 1.0f * (mt->cs[2] * mt->power));
        break;
      case PHOS_MAT_LAYERED:
        for (uint32_t k = 0; k < mt->num_lobes && k < PHOS_MAX_LOBES; ++k)
          add_lobe(&s->bsdf[m], mt->lobes[k].type, mt->lobes[k].weight, mt->lobes[k].param, mt->lobes[k].param2);
        break;
      default: break;
    }
  }
  /* lights: one per emissive face set, mesh -> set order (mesh.cpp:108-116, scene.cpp:49-55) */
  uint32_t nl = 0, nt = 0;
  for (uint32_t m = 0; m < d->num_meshes; ++m)
    for (uint32_t st = d->set_offset[m]; st < d->set_offset[m + 1]; ++st)
      if (d->materials[d->set_material[st]].kind == PHOS_MAT_EMITTER) {
        nl++;
        nt += d->set_face_offset[st + 1] - d->set_face_offset[st];
      }
  s->nlights = nl;
  s->light_first = (uint32_t*)calloc(nl + 1, sizeof(uint32_t));
  s->light_area = (float*)calloc(nl + 1, sizeof(float));
  s->light_tri_mesh = (uint32_t*)calloc(nt + 1, sizeof(uint32_t));
  s->light_tri_face = (uint32_t*)calloc(nt + 1, sizeof(uint32_t));
  uint32_t l = 0, t = 0;
  for (uint32_t m = 0; m < d->num_meshes; ++m)
    for (uint32_t st = d->set_offset[m]; st < d->set_offset[m + 1]; ++st) {
      const uint32_t mat = d->set_material[st];
      if (d->materials[mat].kind != PHOS_MAT_EMITTER) continue;
      s->light_first[l] = t;
      float area = 0.0f; /* area_light_t::preprocess, light.cpp:30-45 */
      for (uint32_t j = d->set_face_offset[st]; j < d->set_face_offset[st + 1]; ++j, ++t) {
        const uint32_t face3 = d->set_faces[j] * 3;
        s->light_tri_mesh[t] = m | (mat << 16);
        s->light_tri_face[t] = face3;
        const v3 a = vert(d, m, face3, 0), b = vert(d, m, face3, 1), c = vert(d, m, face3, 2);
        area += 0.5f * len(cross(sub(b, a), sub(c, a))); /* triangle_t::area, mesh.cpp:291-298 */
      }
      s->light_area[l] = area;
      l++;
    }
  s->light_first[nl] = t;
  return s;
}

void orc_scene_destroy(void* h) {
  orc_scene* s = (orc_scene*)h;
  if (!s) return;
  free(s->light_first); free(s->light_area); free(s->light_tri_mesh); free(s->light_tri_face);
  free(s->bsdf); free(s->emission); free(s);
}
uint32_t orc_scene_num_lights(void* h) { return ((orc_scene*)h)->nlights; }

/* ---- frames (orthogonal_base.hpp:11-69) --------------------------------------------------------------- */
typedef struct { v3 a, b, c; } base_t;
static base_t make_base(v3 n) {
  base_t r;
  const v3 t = (n.x != n.y || n.x != n.z) ? V(n.z - n.y, n.x - n.z, n.y - n.x) : V(n.z - n.y, n.x + n.z, -n.y - n.x);
  r.a = normalized(t);
  r.b = n;
  r.c = normalized(cross(r.a, n));
  return r;
}
static inline v3 to_world(const base_t* b, v3 v) { return add(add(scl(b->a, v.x), scl(b->b, v.y)), scl(b->c, v.z)); }
static inline v3 to_local(const base_t* b, v3 v) { /* v.x * ia + v.y * ib + v.z * ic, ia = (a.x, b.x, c.x) ... */
  const v3 ia = V(b->a.x, b->b.x, b->c.x), ib = V(b->a.y, b->b.y, b->c.y), ic = V(b->a.z, b->b.z, b->c.z);
  return add(add(scl(ia, v.x), scl(ib, v.y)), scl(ic, v.z));
}

/* ---- tangent-space trigonometry (math/vector.hpp:24-71) ------------------------------------------------- */
static inline float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }
static inline float cos2_theta(v3 v) { return v.y * v.y; }
static inline float sin2_theta(v3 v) { return fmaxf(0.0f, 1.0f - cos2_theta(v)); }
static inline float sin_theta(v3 v) { return sqrtf(sin2_theta(v)); }
static inline float tan_theta(v3 v) { return sin_theta(v) / v.y; }
static inline float tan2_theta(v3 v) { return sin2_theta(v) / cos2_theta(v); }
static inline float cos_phi(v3 v) { const float s = sin_theta(v); return s == 0 ? 1 : clampf(v.x / s, -1.f, 1.f); }
static inline float sin_phi(v3 v) { const float s = sin_theta(v); return s == 0 ? 0 : clampf(v.z / s, -1.f, 1.f); }

/* fresnel::dielectric, math/fresnel.hpp:6-28 */
static float fresnel_dielectric(float cosi, float eta) {
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

/* ---- GGX (microfacet.hpp:306-435), alpha_x = alpha_y = alpha ---------------------------------------------- */
static float ggx_D(float ax, float ay, v3 v) {
  const float t2 = tan2_theta(v);
  if (isinf(t2)) return 0.0f;
  const float c2 = cos2_theta(v);
  const float c4 = c2 * c2;
  const float cp = cos_phi(v), sp = sin_phi(v);
  const float e = (cp * cp / (ax * ax) + sp * sp / (ay * ay)) * t2;
  return (float)(1.0f / (M_PI * ax * ay * c4 * (1 + e) * (1 + e)));
}
static float ggx_Lambda(float ax, float ay, v3 v) {
  const float att = fabsf(tan_theta(v));
  if (isinf(att)) return 0.0f;
  const float cp = cos_phi(v), sp = sin_phi(v);
  const float alpha = sqrtf(cp * cp * ax * ay + sp * sp * ax * ay);
  const float a2t2 = (alpha * att) * (alpha * att);
  return (-1.0f + sqrtf(1.0f + a2t2)) * 0.5f;
}
static inline float ggx_G1(float ax, float ay, v3 v) { return 1.0f / (1.0f + ggx_Lambda(ax, ay, v)); }
static void ggx_sample_slope(float cos_theta, float* slope_x, float* slope_y, float u, float v) {
  if (cos_theta > .9999) {
    const float r = sqrtf(u / (1 - u));
    const float phi = (float)(6.28318530718 * v);
    *slope_x = r * cosf(phi);
    *slope_y = r * sinf(phi);
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
  *slope_x = (A < 0.0f || slope_x2 > 1.0f / tan_t) ? slope_x1 : slope_x2;
  float S;
  if (v > 0.5f) { S = 1.0f; v = 2.0f * (v - 0.5f); }
  else { S = -1.0f; v = 2.0f * (0.5f - v); }
  const float z = (v * (v * (v * 0.27385f - 0.73369f) + 0.46341f)) / (v * (v * (v * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
  *slope_y = S * z * sqrtf(1.0f + *slope_x * *slope_x);
}
/* ggx_t::sample: wi is the tangent-space direction; G1 is evaluated on it as well (local, :432) */
static v3 ggx_sample(float ax, float ay, v3 wi, float* pdf, float u, float v) {
  const v3 stretched = normalized(V(ax * wi.x, wi.y, ay * wi.z));
  float slope_x, slope_y;
  ggx_sample_slope(stretched.y, &slope_x, &slope_y, u, v);
  const float tmp = cos_phi(stretched) * slope_x - sin_phi(stretched) * slope_y;
  slope_y = sin_phi(stretched) * slope_x + cos_phi(stretched) * slope_y;
  slope_x = tmp;
  slope_x = slope_x * ax;
  slope_y = slope_y * ay;
  const v3 wh = normalized(V(-slope_x, 1.0f, -slope_y));
  *pdf = (ggx_D(ax, ay, wh) * ggx_G1(ax, ay, wi) * fabsf(dot(wi, wh)) / fabsf(wi.y));
  return wh;
}
/* cook_torrance::f, microfacet.hpp:174-215 (eta hard-wired 0.5, :209) */
static float ct_f(v3 n, float ax, float ay, v3 wi, v3 wo) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi), lo = to_local(&base, wo);
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  v3 wh = add(li, lo);
  const float cos_ti = fabsf(li.y), cos_to = fabsf(lo.y);
  if (cos_ti == 0 || cos_to == 0) return 0.0f;
  if (wh.x == 0 || wh.y == 0 || wh.z == 0) return 0.0f;
  wh = normalized(wh);
  const float d = ggx_D(ax, ay, wh);
  const float g = 1.0f / (1.0f + ggx_Lambda(ax, ay, li) + ggx_Lambda(ax, ay, lo));
  const float whu = (float)(wh.x * 0.0f + wh.y * 1.0 + wh.z * 0.0f); /* wh.dot({0, 1.0, 0}) */
  const float f = fresnel_dielectric(dot(lo, whu < 0.0f ? neg(wh) : wh), 0.5f);
  return d * g * f * (1.0f / (4.0f * cos_ti * cos_to));
}
/* cook_torrance::sample, microfacet.hpp:238-277; returns 0 (black) on the early-outs */
static float ct_sample(v3 n, float ax, float ay, v3 wi, v3* wo, float u, float v, float* opdf) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi);
  if (li.y == 0.0f) return 0.0f;
  float dpdf;
  const v3 wh = ggx_sample(ax, ay, li, &dpdf, u, v);
  if (dot(li, wh) < 0.0f) return 0.0f;
  const v3 lo = add(neg(li), scl(wh, 2.0f * dot(li, wh)));
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  *opdf = dpdf / (4.0f * dot(li, wh));
  *wo = to_world(&base, lo);
  return ct_f(n, ax, ay, wi, *wo);
}

/* ---- the other lobes and bsdf_t itself ------------------------------------------------------------------------ */
/* oren_nayar::f, bsdf/oren_nayar.hpp:9-47 */
static float oren_nayar_f(v3 n, float a, float b, v3 wi, v3 wo) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi), lo = to_local(&base, wo);
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
  return (float)(result * M_1_PI);
}
/* microfacet::sheen, bsdf/sheen.hpp:15-66.  The reference keeps L(0.5, r) in a function-local static that is
 * initialised by the FIRST sheen lobe ever evaluated (sheen.hpp:56); here it is that of the lobe at hand —
 * identical as long as a scene uses one sheen roughness. */
static float sheen_L(float x, float r) {
  static const float p0[] = {25.3245f, 3.32435f, 0.16801f, -1.27393f, -4.85967f};
  static const float p1[] = {21.5473f, 3.82987f, 0.19823f, -1.97760f, -4.32054f};
  const float t = (1.0f - r) * (1.0f - r);
#define SHEEN_INTERP(i) (t * p0[i] + (1.0f - t) * p1[i])
  const float a = SHEEN_INTERP(0), b = SHEEN_INTERP(1), c = SHEEN_INTERP(2), d = SHEEN_INTERP(3), e = SHEEN_INTERP(4);
#undef SHEEN_INTERP
  const float xc = powf(x, c);
  return a / (1 + b * xc) + d * x + e;
}
static float sheen_D(float r, v3 v) {
  const float st = sin_theta(v);
  const float oor = 1.0f / r;
  return (float)((2.0f + oor) * powf(st, oor) / (2.0f * M_PI));
}
static float sheen_Lambda(float r, v3 v) {
  const float L5 = sheen_L(0.5f, r);
  const float ct = v.y;
  const float l = (ct < 0.5f) ? sheen_L(ct, r) : 2.0f * L5 - sheen_L(1.0f - ct, r);
  return expf(l);
}
/* cook_torrance::f with the sheen distribution (bsdf.cpp:88-96, microfacet.hpp:174-215) */
static float sheen_f(v3 n, float r, v3 wi, v3 wo) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi), lo = to_local(&base, wo);
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
/* cook_torrance::pdf, microfacet.hpp:217-236 — G1 is handed the WORLD-space wi there, as here */
static float ct_pdf(v3 n, float ax, float ay, v3 wi, v3 wo) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi), lo = to_local(&base, wo);
  if (!((li.y * lo.y) > 0.0f)) return 0.0f;
  const v3 wh = normalized(add(li, lo));
  return (ggx_D(ax, ay, wh) * ggx_G1(ax, ay, wi) * fabsf(dot(li, wh)) / fabsf(li.y)) / (4.0f * dot(li, wh));
}
/* cook_torrance::refract::{f, pdf, sample}, microfacet.hpp:36-172 (GGX transmission), to the letter: pdf divides by
 * sqrt_denom and multiplies by it again (:113), and tests the hemisphere on the WORLD-space vectors (:108) */
static float ctr_f(v3 n, float ax, float ay, float peta, v3 wi, v3 wo) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi), lo = to_local(&base, wo);
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
static float ctr_pdf(v3 n, float ax, float ay, float peta, v3 wi, v3 wo) {
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi), lo = to_local(&base, wo);
  const float eta = li.y > 0.0f ? peta : 1.0f / peta;
  if (dot(wo, wi) > 0.0) return 0;
  const v3 wh = normalized(add(li, scl(lo, eta)));
  const float sqrt_denom = dot(li, wh) + eta * dot(lo, wh);
  const float dwh_dwi = fabsf(eta * eta * dot(lo, wh)) / sqrt_denom * sqrt_denom;
  return (ggx_D(ax, ay, wh) * wh.y) * dwh_dwi;
}
static float ctr_sample(v3 n, float ax, float ay, float peta, v3 wi, v3* wo, float u, float v, float* opdf) {
  if (peta == 1.0f) {
    *wo = neg(wi);
    *opdf = 1.0f;
    return 1.0f;
  }
  const base_t base = make_base(n);
  const v3 li = to_local(&base, wi);
  if (li.y == 0.0f) return 0.0f;
  float dpdf;
  const v3 wh = ggx_sample(ax, ay, li, &dpdf, u, v);
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
  *opdf = dpdf * dwh_dwi;
  *wo = to_world(&base, lo);
  return ctr_f(n, ax, ay, peta, wi, *wo);
}
/* sample::hemisphere::cosine_weighted + orthogonal_base_t::to_world (math/sampling.hpp:23-36) */
static v3 cosine_sample(v3 n, float sx, float sy, float* pdf) {
  const base_t base = make_base(n);
  const float rr = sqrtf(sx);
  const float theta = (float)(2 * M_PI * sy);
  const float x = rr * cosf(theta), y = rr * sinf(theta);
  const v3 lo = V(x, sqrtf(fmaxf(0.0f, 1.0f - sx)), y);
  *pdf = lo.y * (float)(1.0f / M_PI);
  return to_world(&base, lo);
}
/* eval(), bsdf.cpp:25-107: grey value of one lobe and its pdf */
static float lobe_eval(const lobe1* l, v3 n, v3 wi, v3 wo, float* pdf) {
  switch (l->type) {
    case PHOS_LOBE_DIFFUSE: *pdf = (float)(dot(n, wi) * M_1_PI); return (float)M_1_PI;
    case PHOS_LOBE_OREN_NAYAR: *pdf = (float)(dot(n, wi) * M_1_PI); return oren_nayar_f(n, l->p0, l->p1, wi, wo);
    case PHOS_LOBE_MICROFACET: *pdf = ct_pdf(n, l->p0, l->p1, wi, wo); return ct_f(n, l->p0, l->p1, wi, wo);
    case PHOS_LOBE_SHEEN: *pdf = (float)(dot(n, wi) * M_1_PI); return sheen_f(n, l->p0, wi, wo);
    case PHOS_LOBE_MICROFACET_REFRACT: *pdf = ctr_pdf(n, l->p0, l->p1, l->p2, wi, wo); return ctr_f(n, l->p0, l->p1, l->p2, wi, wo);
    default: *pdf = 0.0f; return 0.0f; /* Reflection, Refraction, Transparent */
  }
}
/* bsdf_t::f, bsdf.cpp:113-131 */
static v3 bsdf_f(const bsdf1* b, v3 n, v3 wi, v3 wo) {
  v3 out = V(0, 0, 0);
  for (uint32_t i = 0; i < b->n; ++i) {
    float ignored;
    const float e = lobe_eval(&b->l[i], n, wi, wo, &ignored);
    const float atl = dot(n, wi);
    const int reflect = atl * dot(n, wo) > 0.0f;
    if ((reflect && (b->l[i].flags & BSDF_REFLECT_F)) || (!reflect && (b->l[i].flags & BSDF_TRANSMIT_F)))
      out = add(out, scl(mul(V(e, e, e), b->l[i].w), atl));
  }
  return out;
}
/* bsdf_t::sample, bsdf.cpp:133-248.  Returns 0 when the path ends (black sample or pdf 0; a lobe that bails out
 * before setting its pdf leaves the reference with an uninitialised one — frozen to "path ends" here). */
static int bsdf_sample(const bsdf1* b, v3 n, float sx, float sy, v3 wi, v3* wo, v3* f, float* opdf, uint32_t* oflags) {
  if (b->n == 0) return 0;
  uint32_t index = (uint32_t)floorf(sx * b->n);
  if (index > b->n - 1) index = b->n - 1;
  const float u = fminf(sx * b->n - index, 1.0f - FLT_EPSILON);
  const lobe1* l = &b->l[index];
  float pdf = 0.0f, r = 0.0f;
  switch (l->type) {
    case PHOS_LOBE_DIFFUSE: *wo = cosine_sample(n, u, sy, &pdf); r = (float)M_1_PI; break;
    case PHOS_LOBE_OREN_NAYAR: *wo = cosine_sample(n, u, sy, &pdf); r = oren_nayar_f(n, l->p0, l->p1, wi, *wo); break;
    case PHOS_LOBE_MICROFACET:
      r = ct_sample(n, l->p0, l->p1, wi, wo, u, sy, &pdf);
      if (r == 0.0f) return 0;
      break;
    case PHOS_LOBE_SHEEN: *wo = cosine_sample(n, u, sy, &pdf); r = sheen_f(n, l->p0, wi, *wo); break;
    case PHOS_LOBE_MICROFACET_REFRACT:
      r = ctr_sample(n, l->p0, l->p1, l->p2, wi, wo, u, sy, &pdf);
      if (r == 0.0f) return 0;
      break;
    case PHOS_LOBE_REFLECTION: { /* reflection.hpp:8-21 */
      const float ct = dot(n, wi);
      pdf = 1.0f;
      *wo = add(neg(wi), scl(n, 2.0f * ct));
      r = 1.0f;
      break;
    }
    case PHOS_LOBE_REFRACTION: { /* refraction.hpp:10-46 */
      pdf = 1.0f;
      float ct = dot(n, wi);
      const float st = fmaxf(0.0f, 1.0f - ct * ct);
      v3 nn;
      float eta = l->p0;
      if (ct > 0) { nn = n; eta = 1.0f / eta; }
      else { nn = neg(n); ct = -ct; }
      const float arg = 1.0f - (eta * eta * st);
      if (!(arg >= 0.0f)) return 0; /* total internal reflection: black */
      const float dnp = sqrtf(arg);
      const float nk = eta * ct - dnp;
      *wo = add(scl(neg(wi), eta), scl(nn, nk));
      r = 1.0f;
      break;
    }
    case PHOS_LOBE_TRANSPARENT: *wo = neg(wi); pdf = 1.0f; r = 1.0f; break;
    default: return 0;
  }
  if (pdf == 0.0f) return 0;
  v3 result = mul(V(r, r, r), l->w);
  int matched = 1;
  for (uint32_t i = 0; i < b->n; ++i) {
    if (i == index || (l->flags & b->l[i].flags) != b->l[i].flags) continue;
    const int reflect = dot(n, wi) * dot(n, *wo) > 0.0f;
    if ((reflect && (b->l[i].flags & BSDF_REFLECT_F)) || (!reflect && (b->l[i].flags & BSDF_TRANSMIT_F))) {
      float lobe_pdf = 0.0f;
      const float e = lobe_eval(&b->l[i], n, wi, *wo, &lobe_pdf);
      result = add(result, mul(V(e, e, e), b->l[i].w));
      pdf += lobe_pdf;
      ++matched;
    }
  }
  pdf /= matched;
  *f = result;
  *opdf = pdf;
  *oflags = l->flags;
  return 1;
}

/* parity hooks: bsdf_t::f / bsdf_t::sample of one material for given directions (tests pin them to the compiled
 * reference's own bsdf_t) */
void orc_bsdf_f(const phos_material* mt, const float* n, const float* wi, const float* wo, float* out3) {
  phos_scene_desc d;
  memset(&d, 0, sizeof(d));
  d.num_materials = 1;
  d.materials = mt;
  orc_scene* S = (orc_scene*)orc_scene_create(&d);
  const v3 f = bsdf_f(&S->bsdf[0], V(n[0], n[1], n[2]), V(wi[0], wi[1], wi[2]), V(wo[0], wo[1], wo[2]));
  out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
  orc_scene_destroy(S);
}
int orc_bsdf_sample(const phos_material* mt, const float* n, const float* wi, float sx, float sy, float* wo3, float* f3,
                    float* pdf, uint32_t* flags) {
  phos_scene_desc d;
  memset(&d, 0, sizeof(d));
  d.num_materials = 1;
  d.materials = mt;
  orc_scene* S = (orc_scene*)orc_scene_create(&d);
  v3 wo = V(0, 0, 0), f = V(0, 0, 0);
  *pdf = 0.0f;
  *flags = 0;
  const int ok = bsdf_sample(&S->bsdf[0], V(n[0], n[1], n[2]), sx, sy, V(wi[0], wi[1], wi[2]), &wo, &f, pdf, flags);
  wo3[0] = wo.x; wo3[1] = wo.y; wo3[2] = wo.z;
  f3[0] = f.x; f3[1] = f.y; f3[2] = f.z;
  orc_scene_destroy(S);
  return ok && !(f.x == 0.0f && f.y == 0.0f && f.z == 0.0f) && *pdf != 0.0f; /* what sample_bsdf continues on (spt.hpp:291) */
}

/* ---- shading normal: mesh_t::shading_parameters, mesh.cpp:169-206 ------------------------------------------ */
static v3 shading_normal(const phos_scene_desc* d, uint32_t mesh, uint32_t face3, float u, float v) {
  unsigned smooth = d->mesh_smooth[mesh];
  if (smooth == 2) smooth = d->face_smooth[(size_t)d->face_offset[mesh] + face3 / 3]; /* mesh_t::details_t::smooth[face], mesh.cpp:10-18 */
  if (smooth && d->normals) {
    const float w = 1 - u - v;
    const size_t f = 3 * (size_t)d->face_offset[mesh] + face3;
    const size_t ia = d->faces[f] + d->vert_offset[mesh], ib = d->faces[f + 1] + d->vert_offset[mesh],
                 ic = d->faces[f + 2] + d->vert_offset[mesh];
    const v3 n0 = V(d->normals[3 * ia], d->normals[3 * ia + 1], d->normals[3 * ia + 2]);
    const v3 n1 = V(d->normals[3 * ib], d->normals[3 * ib + 1], d->normals[3 * ib + 2]);
    const v3 n2 = V(d->normals[3 * ic], d->normals[3 * ic + 1], d->normals[3 * ic + 2]);
    return normalized(add(add(scl(n0, w), scl(n1, u)), scl(n2, v)));
  }
  const v3 v0 = vert(d, mesh, face3, 0), v1 = vert(d, mesh, face3, 1), v2 = vert(d, mesh, face3, 2);
  return normalized(cross(sub(v1, v0), sub(v2, v0)));
}

/* vector3_t<8>::normalize, simd/vector.hpp:126-133: x * rcp(sqrt(dot)) with dot = madd(x,x,madd(y,y,z*z)) */
static v3 simd_normalize(v3 a, int rcp_mode) {
  const float l = fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z));
  float ool;
#if defined(__x86_64__)
  if (rcp_mode) ool = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(sqrtf(l))));
  else
#endif
    ool = 1.0f / sqrtf(l);
  (void)rcp_mode;
  return V(a.x * ool, a.y * ool, a.z * ool);
}

/* camera::perspective_kernel_t for one pixel (kernels/cpu/camera.hpp:113-152): film jitter (jx, jy), lens
 * sample (lu, lv) used only when aperture_radius != 0 (entities/camera.hpp:37-39).
 * The thin lens follows the reference to the letter, including what looks like slips in
 * simd::concentric_sample_disc (math/simd/sampling.hpp:7-32): the [-1,1) `offset` is computed and never used
 * (the raw [0,1) samples are), the constants named pi_o_4 / pi_o_2 hold 4/pi and 2/pi, and simd::select(m, l, r)
 * is blendv(l, r, m), i.e. m ? r : l (math/simd/float8.hpp:103-105) — so r and theta are taken from the
 * "other" branch.  Pinned against the compiled reference kernel in tests/test_oracle_pin.py. */
typedef struct { const float* m; float zoom, stepx, stepy, ratio, focal_distance, aperture_radius; } cam1;
static void camera_ray1(const cam1* c, uint32_t px, uint32_t py, float jx, float jy, float lu, float lv, int rcp_mode, v3* o, v3* w) {
  const float* m = c->m;
  const float sy = (float)py, sx = (float)px;
  const float ndcy = 0.5f - (-0.5f + sy) * c->stepy;
  const float ndcx = (-0.5f + sx) * c->stepx - 0.5f;
  v3 dd = V((ndcx + jx * c->stepx) * c->ratio * c->zoom, (ndcy + jy * c->stepy) * c->zoom, -1.0f);
  dd = simd_normalize(dd, rcp_mode);
  v3 p = V(0.0f, 0.0f, 0.0f);
  if (c->aperture_radius != 0.0f) {
    const float pi_o_2 = (float)(2.0f / M_PI), pi_o_4 = (float)(4.0f / M_PI);
    const int x_gt_y = fabsf(lu) > fabsf(lv);
    const float r = x_gt_y ? lv : lu;
    const float theta1 = pi_o_4 * (lv / lu);
    const float theta2 = pi_o_2 - pi_o_4 * (lu / lv);
    const float theta = x_gt_y ? theta2 : theta1;
    const float lx = (r * cosf(theta)) * c->aperture_radius, ly = (r * sinf(theta)) * c->aperture_radius;
    const float ft = fabsf(c->focal_distance / dd.z);
    p = V(lx, ly, 0.0f);
    dd = V(dd.x * ft - p.x, dd.y * ft - p.y, dd.z * ft - p.z);
    dd = simd_normalize(dd, rcp_mode);
  }
  /* transform_point / transform_vector (math/simd/matrix.hpp:58-104): mul, fmadd, fmadd, (+ row 3) */
  *o = V(fmaf(p.z, m[8], fmaf(p.y, m[4], p.x * m[0])) + m[12], fmaf(p.z, m[9], fmaf(p.y, m[5], p.x * m[1])) + m[13],
         fmaf(p.z, m[10], fmaf(p.y, m[6], p.x * m[2])) + m[14]);
  *w = V(fmaf(dd.z, m[8], fmaf(dd.y, m[4], dd.x * m[0])), fmaf(dd.z, m[9], fmaf(dd.y, m[5], dd.x * m[1])),
         fmaf(dd.z, m[10], fmaf(dd.y, m[6], dd.x * m[2])));
}

/* camera rays of a rectangle with caller-chosen samples (parity hook for the thin lens; slot = y * w + x) */
void orc_camera_rays_lens(const phos_camera* cam, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, float jx, float jy,
                          const float* lens_u, const float* lens_v, int rcp_mode, float* px, float* py, float* pz,
                          float* wx, float* wy, float* wz) {
  const cam1 c = {cam->to_world, 1.12f * tanf(cam->fov * 0.5f), 1.0f / (float)cam->film_width, 1.0f / (float)cam->film_height,
                  (float)cam->film_width / (float)cam->film_height, cam->focal_distance, cam->aperture_radius};
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      const size_t k = (size_t)y * w + x;
      v3 o, d;
      camera_ray1(&c, x0 + x, y0 + y, jx, jy, lens_u ? lens_u[k] : 0.5f, lens_v ? lens_v[k] : 0.5f, rcp_mode, &o, &d);
      px[k] = o.x; py[k] = o.y; pz[k] = o.z;
      wx[k] = d.x; wy[k] = d.y; wz[k] = d.z;
    }
}

/* one ray through the oracle traversal */
typedef struct { v3 o, w; float d; uint32_t mesh, face; float u, v; uint32_t flags; } ray1;
static void trace1(const void* nodes, const void* packets, ray1* r, int rcp_mode) {
  orc_rays s = {&r->o.x, &r->o.y, &r->o.z, &r->w.x, &r->w.y, &r->w.z, &r->d, &r->mesh, &r->face, &r->u, &r->v, &r->flags};
  orc_traverse(nodes, packets, &s, 1, NULL, rcp_mode == 1 ? 3 : 1); /* rcp_mode 1: the reference's own approximate slab test */
}

/* Path-trace samples [spp_begin, spp_end) of the pixel rectangle [x0,x0+w) x [y0,y0+h) and add
 * radiance / (spp_total * pps) into film (W*H*4 floats, interleaved RGBA; alpha is set to 1). */
void orc_render(void* scene_h, const void* nodes, const void* packets, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                uint32_t spp_begin, uint32_t spp_end, uint32_t spp_total, uint32_t pps, uint32_t max_depth, uint64_t seed64,
                int rcp_mode, float* film, float* normals /* W*H*3 or NULL: the NORMALS channel, cpu.cpp:194-196 */) {
  const orc_scene* S = (const orc_scene*)scene_h;
  const phos_scene_desc* d = S->d;
  const uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
  const uint32_t W = d->camera.film_width, H = d->camera.film_height;
  float* jx = (float*)malloc(sizeof(float) * spp_total);
  float* jy = (float*)malloc(sizeof(float) * spp_total);
  orc_film_jitter(seed, spp_total, jx, jy);
  const cam1 cam = {d->camera.to_world, 1.12f * tanf(d->camera.fov * 0.5f), 1.0f / (float)W, 1.0f / (float)H, (float)W / (float)H,
                    d->camera.focal_distance, d->camera.aperture_radius};
  const int thin = d->camera.aperture_radius != 0.0f;
  const float scale = 1.0f / (spp_total * pps);
  const uint32_t nl = S->nlights;

  for (uint32_t py = y0; py < y0 + h; ++py)
    for (uint32_t px = x0; px < x0 + w; ++px) {
      const uint32_t pixel = py * W + px;
      float* out = film + 4 * (size_t)pixel;
      out[3] = 1.0f;
      for (uint32_t s = spp_begin; s < spp_end; ++s) {
        /* camera ray: camera.hpp:113-152 (see orc_camera_rays in phos_oracle.c) */
        ray1 r;
        {
          /* lens sample: the reference draws two fresh uniforms per slot and sample (sampling.cpp:104-109) */
          const float lu = thin ? orc_rnd(seed, pixel, s, 0, DIM_LENS) : 0.5f, lv = thin ? orc_rnd(seed, pixel, s, 1, DIM_LENS) : 0.5f;
          camera_ray1(&cam, px, py, jx[s], jy[s], lu, lv, rcp_mode, &r.o, &r.w);
          r.d = FLT_MAX; r.flags = 0; r.mesh = r.face = 0; r.u = r.v = 0;
        }
        v3 beta = V(1, 1, 1), rad = V(0, 0, 0);
        uint32_t depth = 0;
        for (;;) {
          trace1(nodes, packets, &r, rcp_mode);
          if (!(r.flags & ORC_HIT)) { /* miss: out += beta * e_env (spt.hpp:199-202) */
            if (d->environment >= 0) rad = add(rad, mul(beta, S->emission[d->environment]));
            break;
          }
          /* interaction: deferred_shading_kernel.hpp:47-62 */
          const v3 P = add(r.o, scl(r.w, r.d));
          const v3 wo = neg(r.w);
          const uint32_t mesh = r.mesh & 0xffffu, mat = r.mesh >> 16;
          const v3 n = shading_normal(d, mesh, r.face, r.u, r.v);
          if (depth == 0 && normals) { /* channels.normals->set(x, y, primary->n): the last sample that hits wins */
            normals[3 * (size_t)pixel] = n.x; normals[3 * (size_t)pixel + 1] = n.y; normals[3 * (size_t)pixel + 2] = n.z;
          }
          const phos_material* mt = &d->materials[mat];
          const bsdf1* bs = &S->bsdf[mat];
          const v3 e = mt->kind == PHOS_MAT_EMITTER ? S->emission[mat] : V(0, 0, 0);
          /* NEE: fresh_light_samples (sampling.cpp:160-180) + light_sampler_t (spt.hpp:116-148) */
          ray1 sh;
          float light_pdf = 0.0f;
          int have_light = nl > 0;
          if (have_light) {
            const float xl = orc_rnd(seed, pixel, s, depth, DIM_LIGHT);
            const uint32_t l = (uint32_t)fminf(floorf(xl * nl), (float)(nl - 1));
            const float ux = orc_rnd(seed, pixel, s, depth, DIM_LIGHT_U), uy = orc_rnd(seed, pixel, s, depth, DIM_LIGHT_V);
            const uint32_t num = S->light_first[l + 1] - S->light_first[l];
            uint32_t i = (uint32_t)floorf(ux * num); /* light.cpp:55 */
            if (i > num - 1) i = num - 1;
            const float one_minus_epsilon = 1.0f - FLT_EPSILON;
            const float remapped = fminf(ux * num - i, one_minus_epsilon);
            const float sx_ = sqrtf(remapped); /* triangle_t::sample, mesh.cpp:318-324 */
            const float bu = 1 - sx_, bv = uy * sx_;
            const uint32_t t = S->light_first[l] + i;
            const uint32_t lmesh = S->light_tri_mesh[t] & 0xffffu, lface = S->light_tri_face[t];
            const v3 a = vert(d, lmesh, lface, 0), b = vert(d, lmesh, lface, 1), c = vert(d, lmesh, lface, 2);
            const v3 L = add(add(scl(a, bu), scl(b, bv)), scl(c, 1 - bu - bv)); /* barycentric_to_point, mesh.cpp:314-316 */
            light_pdf = (1.0f / S->light_area[l]) / nl;
            sh.mesh = S->light_tri_mesh[t]; sh.face = lface; sh.u = bu; sh.v = bv;
            sh.o = add(P, scl(n, 0.0001f)); /* simd::offset */
            v3 wi = sub(L, sh.o);
            sh.d = sqrtf(fmaf(wi.x, wi.x, fmaf(wi.y, wi.y, wi.z * wi.z))) - 0.0001f;
            wi = simd_normalize(wi, rcp_mode);
            sh.w = wi;
            const int ish = fmaf(n.x, wi.x, fmaf(n.y, wi.y, n.z * wi.z)) >= 0.0f; /* simd::in_same_hemisphere, >= */
            sh.flags = ish ? ORC_SHADOW : (ORC_SHADOW | ORC_MASKED);
            if (!(sh.flags & ORC_MASKED)) trace1(nodes, packets, &sh, rcp_mode);
          }
          /* integrate: spt.hpp:161-210 */
          if (depth == 0 || (r.flags & ORC_SPECULAR)) rad = add(rad, mul(beta, e));
          if (have_light && !(sh.flags & (ORC_HIT | ORC_MASKED)) && bs->n != 0) {
            /* li, spt.hpp:212-255 */
            const v3 f = bsdf_f(bs, n, sh.w, wo);
            const uint32_t lmesh = sh.mesh & 0xffffu, lmat = sh.mesh >> 16;
            const v3 light_n = shading_normal(d, lmesh, sh.face, sh.u, sh.v);
            const v3 le = S->emission[lmat];
            const float pdf = light_pdf * sh.d * sh.d / fabsf(dot(light_n, neg(sh.w)));
            const v3 li = scl(mul(scl(le, 4), f), 1.0f / pdf); /* (light.e * 4) * f * (1 / pdf) */
            rad = add(rad, mul(beta, li));
          }
          ++depth;
          /* sample_bsdf, spt.hpp:257-305 with terminate_path, :307-328 */
          {
            float wgt = 1.0f;
            int alive = depth < max_depth;
            if (alive && depth >= 3) {
              const float yb = 0.212671f * beta.x + 0.715160f * beta.y + 0.072169f * beta.z;
              const float q = fmaxf(0.05f, 1.0f - yb);
              alive = orc_rnd(seed, pixel, s, depth - 1, DIM_RR) >= q;
              if (alive) wgt = (1.0f / (1.0f - q));
            }
            beta = scl(beta, wgt);
            if (!alive) break;
          }
          if (bs->n == 0) break; /* 0-lobe BSDF (an emitter): the path ends (difference 3) */
          {
            /* bsdf_t::sample, bsdf.cpp:133-248, then sample_bsdf, spt.hpp:289-303 */
            const float sx_ = orc_rnd(seed, pixel, s, depth - 1, DIM_BSDF_U);
            const float sy_ = orc_rnd(seed, pixel, s, depth - 1, DIM_BSDF_V);
            v3 sampled = V(0, 0, 0), f = V(0, 0, 0);
            float pdf = 0.0f;
            uint32_t lflags = 0;
            if (!bsdf_sample(bs, n, sx_, sy_, wo, &sampled, &f, &pdf, &lflags)) break;
            if ((f.x == 0.0f && f.y == 0.0f && f.z == 0.0f) || pdf == 0.0f) break;
            const float weight = dot(n, sampled);
            beta = mul(beta, scl(f, fabsf(weight) / pdf));
            r.o = add(P, scl(n, weight < 0.0f ? -0.0001f : 0.0001f)); /* offset(), math/vector.hpp:14-21 */
            r.w = sampled;
            r.d = FLT_MAX;
            r.flags = (lflags & BSDF_SPECULAR_F) ? ORC_SPECULAR : 0; /* rays->specular_bounce, spt.hpp:302 */
          }
        }
        out[0] += rad.x * scale; /* channel_t::add, cpu.cpp:191 */
        out[1] += rad.y * scale;
        out[2] += rad.z * scale;
      }
    }
  free(jx);
  free(jy);
}
