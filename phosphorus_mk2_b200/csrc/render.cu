// render.cu — device side of the frame pipeline around the ray-query kernel: scene upload,
// primary-ray generation, (wavefront shading / NEE / integration / film: see integrate.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/phos_cuda.h"
#include "ctx.hpp"
#include "render_state.hpp"

namespace phos {

void phos_render_release(phos_ctx* ctx) {
  if (!ctx->render) return;
  ctx->render->release();
  delete ctx->render;
  ctx->render = nullptr;
}

// camera::perspective_kernel_t (reference src/kernels/cpu/camera.hpp:78-159), pinhole branch, for one
// pixel; the film jitter (jx, jy) is shared by every pixel of the sample (src/sampling.cpp:98-111).
// Arithmetic in the reference's order with explicit rounding:
//   ndcx = (x - 0.5) * (1/W) - 0.5 ; ndcy = 0.5 - (y - 0.5) * (1/H)                  (:124-127)
//   d = ((ndcx + jx/W) * (W/H) * zoom, (ndcy + jy/H) * zoom, -1), normalised          (:132-135)
//   p = (0,0,0) * M + row 3 ; d = d * M, row-vector convention, mul + two fmadd        (matrix.hpp:58-104)
// One deliberate difference: normalize() multiplies by _mm256_rcp_ps(sqrt(l)) in the reference
// (~12-bit, micro-architecture specific, src/math/simd/vector.hpp:126-133); here it is the correctly
// rounded 1/sqrt(l).
__global__ void camera_rays_kernel(const DevCamera cam, const phos_tile* __restrict__ tiles,
                                   const unsigned long long* __restrict__ offsets, float jx, float jy, uint32_t seed,
                                   uint32_t sample, phos_rays out) {
  const phos_tile t = tiles[blockIdx.y];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.w * t.h) return;
  float o[3], w[3];
  const uint32_t px = t.x + k % t.w, py = t.y + k / t.w, pix = py * cam.width + px;
  // lens sample of (pixel, sample): the same two draws the wavefront takes (wavefront.cu paths_init_kernel)
  const bool thin = cam.aperture_radius != 0.0f;
  const float lu = thin ? rng(seed, pix, sample, 0, DIM_LENS) : 0.5f, lv = thin ? rng(seed, pix, sample, 1, DIM_LENS) : 0.5f;
  camera_ray(cam, px, py, jx, jy, lu, lv, o, w);
  const unsigned long long i = offsets[blockIdx.y] + k;
  out.px[i] = o[0];
  out.py[i] = o[1];
  out.pz[i] = o[2];
  out.wx[i] = w[0];
  out.wy[i] = w[1];
  out.wz[i] = w[2];
  out.d[i] = 3.402823466e+38f;
  out.flags[i] = 0u;
}

DevCamera make_camera(const phos_camera& c) {
  DevCamera d;
  d.rcp_tab = nullptr;
  d.rcp_shift = 0;
  memcpy(d.m, c.to_world, sizeof(d.m));
  d.zoom = 1.12f * std::tan(c.fov * 0.5f);
  d.stepx = 1.0f / (float)c.film_width;
  d.stepy = 1.0f / (float)c.film_height;
  d.ratio = (float)c.film_width / (float)c.film_height;
  d.focal_distance = c.focal_distance;
  d.aperture_radius = c.aperture_radius;
  d.width = c.film_width;
  d.height = c.film_height;
  return d;
}

}  // namespace phos

using namespace phos;

extern "C" {

int phos_cuda_upload_scene(phos_ctx* ctx, const phos_scene_desc* scene) {
  if (!ctx || !scene) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (scene->camera.film_width == 0 || scene->camera.film_height == 0) return fail(ctx, PHOS_ERR_INVALID, "empty film");
  cudaStreamSynchronize(ctx->stream);
  phos_render_release(ctx);
  ctx->render = new RenderState();
  ctx->render->camera = make_camera(scene->camera);
  ctx->render->camera.rcp_tab = ctx->reference_rcp ? ctx->d_rcp_table : nullptr;
  ctx->render->camera.rcp_shift = 23u - (uint32_t)ctx->rcp_bits;
  return ctx->render->upload(ctx, scene);
}

int phos_cuda_reference_normalize(phos_ctx* ctx, int on) {
  if (!ctx) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (on && !ctx->d_rcp_table) {
    std::vector<float> table;
    int bits = 0;
    if (!sample_host_rcp(table, bits))
      return fail(ctx, PHOS_ERR_INVALID, "reference_normalize: this host has no RCPSS that fits the leading-bits model");
    if (!cuda_ok(ctx, cudaMalloc(&ctx->d_rcp_table, table.size() * sizeof(float)), "cudaMalloc(rcp table)") ||
        !cuda_ok(ctx, cudaMemcpy(ctx->d_rcp_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice), "upload rcp table"))
      return PHOS_ERR_CUDA;
    ctx->rcp_bits = bits;
  }
  ctx->reference_rcp = on != 0;
  if (ctx->render) {  // frames already uploaded switch too (the stream is drained first: kernels hold the camera by value)
    cudaStreamSynchronize(ctx->stream);
    ctx->render->camera.rcp_tab = ctx->reference_rcp ? ctx->d_rcp_table : nullptr;
    ctx->render->camera.rcp_shift = 23u - (uint32_t)ctx->rcp_bits;
  }
  return PHOS_OK;
}

int phos_cuda_camera_rays(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, float jx, float jy,
                          const phos_rays* device_rays) {
  return phos_cuda_camera_rays_lens(ctx, tiles, n_tiles, jx, jy, 0, 0, device_rays);
}

int phos_cuda_camera_rays_lens(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, float jx, float jy, uint64_t seed64,
                               uint32_t sample, const phos_rays* device_rays) {
  if (!ctx || !tiles || !device_rays) return PHOS_ERR_INVALID;
  if (!ctx->render) return fail(ctx, PHOS_ERR_INVALID, "camera_rays before upload_scene");
  if (n_tiles == 0) return PHOS_OK;
  cudaSetDevice(ctx->device);
  RenderState& R = *ctx->render;
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
  if (max_px == 0) return PHOS_OK;
  if (!R.set_tiles(ctx, tiles, offsets.data(), n_tiles)) return PHOS_ERR_CUDA;
  const uint32_t block = 256;
  for (uint32_t first = 0; first < n_tiles; first += 65535) {  // grid.y limit
    const uint32_t cnt = std::min<uint32_t>(65535, n_tiles - first);
    dim3 grid((max_px + block - 1) / block, cnt);
    camera_rays_kernel<<<grid, block, 0, ctx->stream>>>(R.camera, R.d_tiles + first, R.d_tile_offsets + first, jx, jy,
                                                        (uint32_t)(seed64 ^ (seed64 >> 32)), sample, *device_rays);
    ctx->launches++;
  }
  return cuda_ok(ctx, cudaGetLastError(), "camera_rays_kernel launch") ? PHOS_OK : PHOS_ERR_CUDA;
}

int phos_cuda_flush_l2(phos_ctx* ctx) {
  if (!ctx) return PHOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t bytes = 256ull << 20;  // twice the 126 MB L2
  if (!ctx->d_flush && !cuda_ok(ctx, cudaMalloc(&ctx->d_flush, bytes), "cudaMalloc(flush buffer)")) return PHOS_ERR_CUDA;
  ctx->flush_value ^= 0xff;
  return cuda_ok(ctx, cudaMemsetAsync(ctx->d_flush, ctx->flush_value, bytes, ctx->stream), "flush_l2") ? PHOS_OK : PHOS_ERR_CUDA;
}

}  // extern "C"

namespace phos {

namespace {
template <typename T>
bool to_device(phos_ctx* ctx, RenderState& R, const T* host, size_t count, const T** out) {
  *out = nullptr;
  if (count == 0) count = 1;  // keep pointers non-null
  void* d = nullptr;
  if (!cuda_ok(ctx, cudaMalloc(&d, count * sizeof(T)), "cudaMalloc(scene)")) return false;
  R.scene_allocs.push_back(d);
  if (host && !cuda_ok(ctx, cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice), "upload scene")) return false;
  *out = (const T*)d;
  return true;
}

// microfacet_t::roughness_to_alpha + the precompute clamp (reference src/bsdf/params.hpp:86-99)
float roughness_to_alpha(float roughness) {
  roughness = std::max(roughness, (float)1e-5);
  const float x = std::log(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

float tri_area(const float* a, const float* b, const float* c) {  // triangle_t::area (src/mesh.cpp:291-298)
  const float abx = b[0] - a[0], aby = b[1] - a[1], abz = b[2] - a[2];
  const float acx = c[0] - a[0], acy = c[1] - a[1], acz = c[2] - a[2];
  const float zx = aby * acz - abz * acy, zy = abz * acx - abx * acz, zz = abx * acy - aby * acx;
  return 0.5f * std::sqrt(zx * zx + zy * zy + zz * zz);
}
}  // namespace

// What tile_renderer_t reads from scene_t while rendering: geometry for shading normals and light
// sampling, the closure table, the area lights (one per emissive face set, in mesh -> set order:
// src/mesh.cpp:108-116, src/scene.cpp:49-55) with their total areas (src/light.cpp:30-45).
int RenderState::upload(phos_ctx* ctx, const phos_scene_desc* d) {
  paths_cap = 0;  // free memory changes with the scene: ask again at the next render
  const uint32_t nm = d->num_meshes;
  if (nm == 0 || !d->vert_offset || !d->vertices || !d->face_offset || !d->faces || !d->mesh_smooth || !d->set_offset ||
      !d->set_material || !d->set_face_offset || !d->set_faces || (d->num_materials && !d->materials))
    return fail(ctx, PHOS_ERR_INVALID, "incomplete scene description");
  if (!scene_indices_ok(d)) return fail(ctx, PHOS_ERR_INVALID, "face with a vertex index outside its mesh");
  num_meshes = nm;
  num_materials = d->num_materials;
  const size_t nv = d->vert_offset[nm], nf = d->face_offset[nm];
  for (uint32_t m = 0; m < nm; ++m)
    if (d->mesh_smooth[m] && !d->normals) return fail(ctx, PHOS_ERR_INVALID, "smooth mesh without vertex normals");
  bool mixed = false;
  for (uint32_t m = 0; m < nm; ++m) {
    if (d->mesh_smooth[m] > 2) return fail(ctx, PHOS_ERR_INVALID, "mesh_smooth must be 0, 1 or 2");
    mixed = mixed || d->mesh_smooth[m] == 2;
  }
  if (mixed && !d->face_smooth) return fail(ctx, PHOS_ERR_INVALID, "mesh_smooth = 2 without face_smooth");
  const uint32_t nsets = d->set_offset[nm];
  for (uint32_t s = 0; s < nsets; ++s)
    if (d->set_material[s] >= d->num_materials) return fail(ctx, PHOS_ERR_INVALID, "face set with an unknown material");

  if (d->environment >= 0 && ((uint32_t)d->environment >= d->num_materials || d->materials[d->environment].kind != PHOS_MAT_BACKGROUND))
    return fail(ctx, PHOS_ERR_INVALID, "environment is not a background material of the scene");
  scene.environment = d->environment;
  std::vector<DevMaterial> mats(std::max<uint32_t>(1, d->num_materials));
  memset(mats.data(), 0, mats.size() * sizeof(DevMaterial));
  // bsdf_t::add_lobe + T::precompute for one closure (bsdf.hpp:52-83, bsdf/params.hpp)
  auto add_lobe = [&](DevMaterial& o, uint32_t type, const float* w, float param, float param2 = 0.0f) -> bool {
    if (o.nlobes >= PHOS_MAX_LOBES) return false;
    DevLobe& l = o.lobes[o.nlobes++];
    l.type = type;
    for (int c = 0; c < 3; ++c) l.w[c] = w[c];
    l.p0 = l.p1 = l.p2 = 0.0f;
    switch (type) {
      case PHOS_LOBE_MICROFACET_REFRACT:  // add_lobe<microfacet_t> with refract = 1: flags = TRANSMIT
        l.flags = BSDF_TRANSMIT_F;
        l.p0 = l.p1 = std::min(1.0f, std::max(0.0001f, roughness_to_alpha(param)));
        l.p2 = param2;
        return param2 != 0.0f;  // eta 0 divides by zero in the reference as well
      case PHOS_LOBE_DIFFUSE: l.flags = BSDF_REFLECT_F | BSDF_DIFFUSE_F; return true;
      case PHOS_LOBE_OREN_NAYAR: {  // oren_nayar_t::precompute, params.hpp:37-42
        l.flags = BSDF_REFLECT_F | BSDF_DIFFUSE_F;
        const float sg = (float)(param * (M_PI / 180.0f));
        const float s2 = sg * sg;
        l.p0 = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
        l.p1 = 0.45f * s2 / (s2 + 0.09f);
        return true;
      }
      case PHOS_LOBE_REFLECTION: l.flags = BSDF_REFLECT_F | BSDF_SPECULAR_F; l.p0 = param; return true;
      case PHOS_LOBE_REFRACTION: l.flags = BSDF_TRANSMIT_F | BSDF_SPECULAR_F; l.p0 = param; return true;
      case PHOS_LOBE_MICROFACET:  // add_lobe<microfacet_t> with refract = 0: flags = REFLECT (bsdf.hpp:69-83)
        l.flags = BSDF_REFLECT_F;
        l.p0 = l.p1 = std::min(1.0f, std::max(0.0001f, roughness_to_alpha(param)));
        return true;
      case PHOS_LOBE_SHEEN: l.flags = BSDF_REFLECT_F | BSDF_GLOSSY_F; l.p0 = param; return true;
      case PHOS_LOBE_TRANSPARENT: l.flags = BSDF_TRANSMIT_F; return true;  // empty_params_t, material.cpp:98-103
      default: return false;
    }
  };
  for (uint32_t m = 0; m < d->num_materials; ++m) {
    const phos_material& in = d->materials[m];
    DevMaterial& o = mats[m];
    o.kind = in.kind;
    const float cw[3] = {1.0f * in.cs[0], 1.0f * in.cs[1], 1.0f * in.cs[2]};  // w * component->w with w = (1, 1, 1)
    bool ok = true;
    switch (in.kind) {
      case PHOS_MAT_DIFFUSE:  // diffuse_bsdf_node.osl:20-25
        ok = in.roughness == 0 ? add_lobe(o, PHOS_LOBE_DIFFUSE, cw, 0.0f) : add_lobe(o, PHOS_LOBE_OREN_NAYAR, cw, in.roughness);
        break;
      case PHOS_MAT_GLOSSY:  // glossy_bsdf_node.osl:26-34
        ok = in.roughness == 0.0f ? add_lobe(o, PHOS_LOBE_REFLECTION, cw, 0.0f)
                                  : add_lobe(o, PHOS_LOBE_MICROFACET, cw, in.roughness * in.roughness);
        break;
      case PHOS_MAT_EMITTER: {  // diffuse_emitter_node.osl:18
        const float k = (float)(in.power / M_PI);
        for (int c = 0; c < 3; ++c) o.e[c] = 1.0f * (k * in.cs[c]);
        break;
      }
      case PHOS_MAT_BACKGROUND:  // background_node.osl: Cs * power * background()
        for (int c = 0; c < 3; ++c) o.e[c] = 1.0f * (in.cs[c] * in.power);
        break;
      case PHOS_MAT_LAYERED:
        if (in.num_lobes > PHOS_MAX_LOBES) return fail(ctx, PHOS_ERR_INVALID, "more than 8 closures in one material (bsdf_t::MaxLobes)");
        for (uint32_t k = 0; ok && k < in.num_lobes; ++k) ok = add_lobe(o, in.lobes[k].type, in.lobes[k].weight, in.lobes[k].param, in.lobes[k].param2);
        break;
      default: ok = false;
    }
    if (!ok) return fail(ctx, PHOS_ERR_INVALID, "material outside the built-in closure set");
  }
  std::vector<uint32_t> light_first, light_tri_mesh, light_tri_face;
  std::vector<float> light_area;
  for (uint32_t m = 0; m < nm; ++m)
    for (uint32_t s = d->set_offset[m]; s < d->set_offset[m + 1]; ++s) {
      const uint32_t mat = d->set_material[s];
      if (d->materials[mat].kind != PHOS_MAT_EMITTER) continue;
      light_first.push_back((uint32_t)light_tri_mesh.size());
      float area = 0.0f;
      for (uint32_t j = d->set_face_offset[s]; j < d->set_face_offset[s + 1]; ++j) {
        const uint32_t face = d->set_faces[j];
        if (face >= d->face_offset[m + 1] - d->face_offset[m]) return fail(ctx, PHOS_ERR_INVALID, "face set index out of range");
        light_tri_mesh.push_back(m | (mat << 16));
        light_tri_face.push_back(face * 3);
        const uint32_t* f = d->faces + 3 * ((size_t)d->face_offset[m] + face);
        const float* v = d->vertices + 3 * (size_t)d->vert_offset[m];
        area += tri_area(v + 3 * (size_t)f[0], v + 3 * (size_t)f[1], v + 3 * (size_t)f[2]);
      }
      light_area.push_back(area);
    }
  scene.nlights = (uint32_t)light_area.size();
  light_first.push_back((uint32_t)light_tri_mesh.size());

  const DevMaterial* dm = nullptr;
  bool ok = to_device(ctx, *this, d->vertices, 3 * nv, &scene.verts) && to_device(ctx, *this, d->faces, 3 * nf, &scene.faces) &&
            to_device(ctx, *this, d->vert_offset, (size_t)nm + 1, &scene.vert_offset) &&
            to_device(ctx, *this, d->face_offset, (size_t)nm + 1, &scene.face_offset) &&
            to_device(ctx, *this, d->mesh_smooth, (size_t)nm, &scene.mesh_smooth) &&
            to_device(ctx, *this, mats.data(), mats.size(), &dm) &&
            to_device(ctx, *this, light_first.data(), light_first.size(), &scene.light_first) &&
            to_device(ctx, *this, light_area.data(), light_area.size(), &scene.light_area) &&
            to_device(ctx, *this, light_tri_mesh.data(), light_tri_mesh.size(), &scene.light_tri_mesh) &&
            to_device(ctx, *this, light_tri_face.data(), light_tri_face.size(), &scene.light_tri_face);
  scene.mats = dm;
  scene.face_smooth = nullptr;
  if (ok && mixed) ok = to_device(ctx, *this, d->face_smooth, nf, &scene.face_smooth);
  scene.normals = nullptr;
  if (ok && d->normals) ok = to_device(ctx, *this, d->normals, 3 * nv, &scene.normals);
  if (!ok) return PHOS_ERR_CUDA;
  const size_t film_bytes = (size_t)camera.width * camera.height * 4 * sizeof(float);
  if (!cuda_ok(ctx, cudaMalloc(&film, film_bytes), "cudaMalloc(film)") || !cuda_ok(ctx, cudaMemset(film, 0, film_bytes), "clear film"))
    return PHOS_ERR_CUDA;
  return PHOS_OK;
}

static void release_one(Wavefront& wf) {
  free_rays(wf.rays[0]);
  free_rays(wf.rays[1]);
  free_rays(wf.shadow);
  void* ptrs[] = {wf.slot_path[0], wf.slot_path[1], wf.count, wf.n, wf.lightw, wf.beta[0], wf.beta[1], wf.rad[0], wf.rad[1], wf.rad_final, wf.pixel, wf.perm};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  wf = Wavefront();
}

bool RenderState::ensure_wavefront(phos_ctx* ctx, uint64_t paths, uint64_t pixels, int which) {
  Wavefront& wf = wfs[which];
  if (wf.capacity < paths || wf.pixel_capacity < pixels) {
    for (cudaStream_t s : {ctx->stream, ctx->s_cmp, ctx->s_in, ctx->s_out}) cudaStreamSynchronize(s);  // the wavefront streams
    release_one(wf);
    const uint64_t cap = std::max<uint64_t>(paths, 1024);
    bool ok = alloc_rays(ctx, cap, wf.rays[0]) && alloc_rays(ctx, cap, wf.rays[1]) && alloc_rays(ctx, cap, wf.shadow);
    auto get = [&](void** p, size_t bytes) { return ok && (ok = cuda_ok(ctx, cudaMalloc(p, bytes), "cudaMalloc(wavefront)")); };
    get((void**)&wf.slot_path[0], cap * 4);
    get((void**)&wf.slot_path[1], cap * 4);
    get((void**)&wf.count, 16);
    get((void**)&wf.n, cap * 12);
    get((void**)&wf.lightw, cap * 16);
    for (int k = 0; k < 2; ++k) {
      get((void**)&wf.beta[k], cap * 12);
      get((void**)&wf.rad[k], cap * 12);
    }
    get((void**)&wf.rad_final, cap * 12);
    get((void**)&wf.pixel, std::max<uint64_t>(pixels, 1024) * 4);
    get((void**)&wf.perm, cap * 4);
    if (!ok) {
      release_one(wf);
      return false;
    }
    wf.capacity = cap;
    wf.pixel_capacity = std::max<uint64_t>(pixels, 1024);
  }
  return true;
}

void RenderState::release_wavefront() {
  for (Wavefront& w : wfs) release_one(w);
}

bool RenderState::set_tiles(phos_ctx* ctx, const phos_tile* tiles, const unsigned long long* offsets, uint32_t n) {
  if (n > tile_capacity) {
    cudaStreamSynchronize(ctx->stream);
    if (d_tiles) cudaFree(d_tiles);
    if (d_tile_offsets) cudaFree(d_tile_offsets);
    d_tiles = nullptr;
    d_tile_offsets = nullptr;
    tile_capacity = 0;
    if (!cuda_ok(ctx, cudaMalloc(&d_tiles, n * sizeof(phos_tile)), "cudaMalloc(tiles)") ||
        !cuda_ok(ctx, cudaMalloc(&d_tile_offsets, n * sizeof(unsigned long long)), "cudaMalloc(tile offsets)"))
      return false;
    tile_capacity = n;
  }
  // pageable sources: the copies are staged before the call returns, so the caller's arrays may go away
  return cuda_ok(ctx, cudaMemcpyAsync(d_tiles, tiles, n * sizeof(phos_tile), cudaMemcpyHostToDevice, ctx->stream), "upload tiles") &&
         cuda_ok(ctx, cudaMemcpyAsync(d_tile_offsets, offsets, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream),
                 "upload tile offsets");
}

void RenderState::release() {
  release_wavefront();
  for (void* p : scene_allocs) cudaFree(p);
  scene_allocs.clear();
  if (film) cudaFree(film);
  if (film_normals) cudaFree(film_normals);
  film_normals = nullptr;
  if (d_jitter) cudaFree(d_jitter);
  film = nullptr;
  d_jitter = nullptr;
  jitter_capacity = 0;
  if (d_tiles) cudaFree(d_tiles);
  if (d_tile_offsets) cudaFree(d_tile_offsets);
  d_tiles = nullptr;
  d_tile_offsets = nullptr;
  tile_capacity = 0;
}

}  // namespace phos
