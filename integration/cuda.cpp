// integration/cuda.cpp — drop-in replacement for the reference's src/xpu/cuda.cpp (empty bodies,
// reference src/xpu/cuda.cpp:5-14).  Everything on the device goes through the C ABI of
// include/phos_cuda.h; this file only adapts the reference's types:
//   preprocess  = cpu_t::details_t::reset (src/xpu/cpu.cpp:35-44: the reference's own CPU BVH build)
//                 + one upload of mbvh_t::root / triangles and of the flattened scene;
//   start/join  = cpu_t::start / join (src/xpu/cpu.cpp:223-244) with ONE worker that takes tiles off the shared
//                 queue (frame.tiles, src/jobs/tiles.hpp:40-47 — the same cursor every other device of the
//                 session drains, plugins/blender/session.cpp:85-99) in guided chunks, renders a chunk as one
//                 wavefront, reads its film rows back ONCE into page-locked memory and hands every tile to
//                 frame.film->add_tile (src/xpu/cpu.cpp:201-204) while the GPU already renders the next chunk.
//                 Errors inside the worker surface in join() (a throw out of a std::thread would terminate).
#include "cuda.hpp"

#include "accel/bvh.hpp"
#include "accel/bvh/binned_sah_builder.hpp"
#include "jobs/tiles.hpp"
#include "material.hpp"
#include "mesh.hpp"
#include "options.hpp"
#include "scene.hpp"
#include "state.hpp"
#include "utils/allocator.hpp"

#include <phos_cuda.h>

#include "light.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

// The material system has to say which built-in closure a material is (diffuse_bsdf_node /
// glossy_bsdf_node / diffuse_emitter_node and their Cs, roughness, power): with OSL that is a
// ShadingSystem::getattribute query on the shader group; the table-driven stand-in answers directly.
bool material_builtin_closure(const material_t* m, uint32_t* kind, float cs[3], float* roughness, float* power);
// ... and, for a material whose shader group flattens to a list of closures (mix / add trees): that list
int material_builtin_lobes(const material_t* m, uint32_t* type, float* weight3, float* param, float* param2);

namespace {

void check(phos_ctx* c, int rc) {
  if (rc != PHOS_OK) throw std::runtime_error(std::string("libphos_cuda: ") + phos_cuda_last_error(c));
}

// mbvh_t never publishes its node count (src/accel/bvh.cpp:42-47): largest reachable index + 1
uint32_t count_nodes(const accel::mbvh_t& a) {
  if (!a.root) return 0;
  uint32_t max_index = 0;
  std::vector<uint32_t> stack{0};
  while (!stack.empty()) {
    const uint32_t n = stack.back();
    stack.pop_back();
    if (n > max_index) max_index = n;
    const auto& node = a.root[n];
    for (int i = 0; i < 8; ++i)
      if (!(node.bounds[i] > node.bounds[i + 24]) && node.flags[i] != 1) stack.push_back(node.offset[i]);
  }
  return max_index + 1;
}

// scene_t / mesh_t public arrays (src/mesh.hpp:68-77) -> include/phos_scene.h
struct flat_scene_t {
  std::vector<uint32_t> vert_offset, face_offset, set_offset, set_material, set_face_offset, set_faces, faces;
  std::vector<float> vertices, normals;
  std::vector<uint8_t> mesh_smooth, face_smooth;
  std::vector<phos_material> materials;
  phos_scene_desc desc;
};

void flatten(const scene_t& scene, flat_scene_t& f) {
  const uint32_t nm = scene.num_meshes();
  // face sets per mesh: scene_t::triangles walks mesh -> set -> face (src/scene.cpp:58-62)
  std::vector<triangle_t> tris;
  scene.triangles(tris);
  std::vector<uint32_t> num_sets(nm, 0);
  for (const auto& t : tris) num_sets[t.meshid()] = std::max(num_sets[t.meshid()], t.set + 1);
  bool all_normals = true, any_mixed = false;
  f.vert_offset.assign(1, 0);
  f.face_offset.assign(1, 0);
  f.set_offset.assign(1, 0);
  f.set_face_offset.assign(1, 0);
  for (uint32_t m = 0; m < nm; ++m) {
    const mesh_t* mesh = scene.mesh(m);
    uint32_t nv = 0;
    for (uint32_t i = 0; i < 3 * mesh->num_faces; ++i) nv = std::max(nv, mesh->faces[i] + 1);
    const bool has_n = mesh->has_per_vertex_normals() && mesh->normals;
    all_normals = all_normals && has_n;
    for (uint32_t v = 0; v < nv; ++v) {
      f.vertices.insert(f.vertices.end(), {mesh->vertices[v].x, mesh->vertices[v].y, mesh->vertices[v].z});
      if (has_n) f.normals.insert(f.normals.end(), {mesh->normals[v].x, mesh->normals[v].y, mesh->normals[v].z});
      else f.normals.insert(f.normals.end(), {0.0f, 0.0f, 0.0f});
    }
    f.faces.insert(f.faces.end(), mesh->faces, mesh->faces + 3 * mesh->num_faces);
    // the per-face flag of builder_t::add_face(a, b, c, smooth) (src/mesh.hpp:46-66, src/mesh.cpp:10-18,202-206)
    bool smooth = mesh->num_faces > 0, flat = mesh->num_faces > 0;
    for (uint32_t face = 0; face < mesh->num_faces; ++face) {
      const bool sm = mesh->is_smooth(face);
      (sm ? flat : smooth) = false;
      f.face_smooth.push_back(sm ? 1 : 0);
    }
    f.mesh_smooth.push_back(smooth ? 1 : (flat || !mesh->num_faces) ? 0 : 2);  // 2: mixed, read face_smooth
    any_mixed = any_mixed || f.mesh_smooth.back() == 2;
    for (uint32_t s = 0; s < num_sets[m]; ++s) {
      f.set_material.push_back(mesh->sets[s].material);
      f.set_faces.insert(f.set_faces.end(), mesh->sets[s].faces, mesh->sets[s].faces + mesh->sets[s].num_faces);
      f.set_face_offset.push_back((uint32_t)f.set_faces.size());
    }
    f.vert_offset.push_back(f.vert_offset.back() + nv);
    f.face_offset.push_back(f.face_offset.back() + mesh->num_faces);
    f.set_offset.push_back(f.set_offset.back() + num_sets[m]);
  }
  for (uint32_t i = 0; i < scene.num_materials(); ++i) {
    phos_material pm = {};
    if (!material_builtin_closure(scene.material(i), &pm.kind, pm.cs, &pm.roughness, &pm.power))
      throw std::runtime_error("cuda_t: material outside the built-in closure set");
    if (pm.kind == PHOS_MAT_LAYERED) {
      uint32_t type[PHOS_MAX_LOBES];
      float weight[3 * PHOS_MAX_LOBES], param[PHOS_MAX_LOBES], param2[PHOS_MAX_LOBES];
      pm.num_lobes = (uint32_t)material_builtin_lobes(scene.material(i), type, weight, param, param2);
      for (uint32_t k = 0; k < pm.num_lobes; ++k) {
        pm.lobes[k].type = type[k];
        for (int c = 0; c < 3; ++c) pm.lobes[k].weight[c] = weight[3 * k + c];
        pm.lobes[k].param = param[k];
        pm.lobes[k].param2 = param2[k];
      }
    }
    f.materials.push_back(pm);
  }
  phos_scene_desc& d = f.desc;
  d.num_meshes = nm;
  d.vert_offset = f.vert_offset.data();
  d.vertices = f.vertices.data();
  d.normals = all_normals ? f.normals.data() : nullptr;
  d.face_offset = f.face_offset.data();
  d.faces = f.faces.data();
  d.mesh_smooth = f.mesh_smooth.data();
  d.face_smooth = any_mixed ? f.face_smooth.data() : nullptr;
  d.set_offset = f.set_offset.data();
  d.set_material = f.set_material.data();
  d.set_face_offset = f.set_face_offset.data();
  d.set_faces = f.set_faces.data();
  d.num_materials = (uint32_t)f.materials.size();
  d.materials = f.materials.data();
  const camera_t& c = scene.camera;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) d.camera.to_world[4 * i + j] = c.to_world.x[i][j];
  d.camera.fov = c.fov;
  d.camera.focal_distance = c.focal_distance;
  d.camera.aperture_radius = c.aperture_radius;
  d.camera.film_width = c.film.width;
  d.camera.film_height = c.film.height;
  // scene_t::environment() (src/scene.cpp:126-128): what a ray that leaves the scene picks up (spt.hpp:199-202)
  d.environment = scene.environment() ? (int32_t)scene.environment()->matid() : -1;
}

}  // namespace

std::atomic<int> cuda_t::instances(0);

int cuda_t::device_count() { return phos_cuda_device_count(); }

cuda_t* cuda_t::make(const parsed_options_t& options, int device) {
  phos_options po{options.samples_per_pixel, options.paths_per_sample, options.path_depth};
  phos_ctx* ctx = phos_cuda_create(device, &po);
  if (!ctx) throw std::runtime_error(std::string("libphos_cuda: ") + phos_cuda_last_error(nullptr));
  auto* d = new cuda_t();
  d->ctx = ctx;
  d->spp = options.samples_per_pixel;
  ++instances;
  // A device that shares frames with cpu_t normalises the way cpu_t does (the host's _mm256_rcp_ps, src/math/simd/
  // vector.hpp:126-133) so that its tiles match the CPU's; PHOS_EXACT_NORMALIZE=1 renders with exact arithmetic instead.
  const char* exact = std::getenv("PHOS_EXACT_NORMALIZE");
  if (!(exact && exact[0] == '1')) check(ctx, phos_cuda_reference_normalize(ctx, 1));
  return d;
}

cuda_t::~cuda_t() {
  if (worker.joinable()) worker.join();
  if (pinned) phos_cuda_host_free(pinned);
  if (ctx) --instances;
  phos_cuda_destroy(ctx);
}

void cuda_t::preprocess(const scene_t& scene) {
  accel::mbvh_t accel;
  {
    accel::mbvh_t::builder_t::scoped_t builder(accel.builder());
    std::vector<triangle_t> triangles;
    scene.triangles(triangles);
    bvh::from(builder, triangles);
  }
  check(ctx, phos_cuda_upload_accel(ctx, accel.root, count_nodes(accel), accel.triangles, accel.num_triangles));
  flat_scene_t flat;
  flatten(scene, flat);
  check(ctx, phos_cuda_upload_scene(ctx, &flat.desc));
  film_w = scene.camera.film.width;
  film_h = scene.camera.film.height;
}

namespace {

// A claim on the shared tile cursor: tiles [first, first + count) of frame.tiles->tiles — exactly what `count` calls of
// tiles_t::next would hand out (src/jobs/tiles.hpp:40-47), taken with one fetch_add.  Guided: half of this device's share of
// what is left, so that N devices (GPUs, or a GPU next to the CPU device) finish together; at least kMinClaim tiles, because
// a wavefront over few pixels leaves the GPU idle in its launch tails, at most kMaxClaim (64 Mi paths at 64 spp).
constexpr uint32_t kMinClaim = 64, kMaxClaim = 1024;

struct claim_t {
  uint32_t first = 0, count = 0;
  uint32_t y0 = 0, y1 = 0;  // film rows covered
  std::vector<phos_tile> tiles;
};

bool take_tiles(job::tiles_t* q, int peers, claim_t& c) {
  const uint32_t seen = q->tile.load();
  if (seen >= q->size) return false;
  const uint32_t share = (q->size - seen) / (2u * (uint32_t)std::max(1, peers));
  const uint32_t want = std::min(kMaxClaim, std::max(kMinClaim, share));
  c.first = q->tile.fetch_add(want);
  if (c.first >= q->size) return false;
  c.count = std::min(want, q->size - c.first);
  c.tiles.clear();
  c.y0 = ~0u;
  c.y1 = 0;
  for (uint32_t i = 0; i < c.count; ++i) {
    const auto& t = q->tiles[c.first + i];
    c.tiles.push_back(phos_tile{t.x, t.y, t.w, t.h});
    c.y0 = std::min(c.y0, t.y);
    c.y1 = std::max(c.y1, t.y + t.h);
  }
  return true;
}

}  // namespace

void cuda_t::start(const scene_t&, frame_state_t& frame) {
  if (worker.joinable()) worker.join();  // a device is started once per view (session.cpp:224-229)
  error = nullptr;
  check(ctx, phos_cuda_film_clear(ctx));
  // the NORMALS channel only when the frame's tile format asks for it (render_buffer_t::NORMALS, cpu.cpp:97)
  bool want_normals = false;
  for (const auto& ch : frame.tiles->format.channels) want_normals = want_normals || ch.name == render_buffer_t::NORMALS;
  check(ctx, phos_cuda_enable_normals(ctx, want_normals ? 1 : 0));
  // page-locked landing zone for the film rows of a chunk (RGBA, then the NORMALS channel): one copy per chunk
  const size_t need = (size_t)film_w * film_h * (want_normals ? 7 : 4) * sizeof(float);
  if (pinned_bytes < need) {
    if (pinned) phos_cuda_host_free(pinned);
    pinned = static_cast<float*>(phos_cuda_host_alloc(need));
    pinned_bytes = pinned ? need : 0;
    if (!pinned) throw std::runtime_error("cuda_t: phos_cuda_host_alloc failed");
  }
  frame_state_t* fs = &frame;
  worker = std::thread([this, fs, want_normals] {
    try {
      allocator_t allocator(1024 * 1024 * 4);
      float* rgba = pinned;
      float* nrm = want_normals ? pinned + (size_t)film_w * film_h * 4 : nullptr;
      claim_t cur, nxt;
      bool have = take_tiles(fs->tiles, instances.load(), cur);
      if (have) check(ctx, phos_cuda_render(ctx, cur.tiles.data(), cur.count, 0, spp, spp, seed));  // asynchronous
      while (have) {
        // the rows of this chunk, once, into page-locked memory (blocks until the chunk is rendered) ...
        check(ctx, phos_cuda_film_read(ctx, rgba + (size_t)cur.y0 * film_w * 4, 0, cur.y0, film_w, cur.y1 - cur.y0));
        if (nrm) check(ctx, phos_cuda_film_read_normals(ctx, nrm + (size_t)cur.y0 * film_w * 3, 0, cur.y0, film_w, cur.y1 - cur.y0));
        // ... the next chunk goes to the GPU before the host starts slicing this one into tiles
        const bool more = take_tiles(fs->tiles, instances.load(), nxt);
        if (more) check(ctx, phos_cuda_render(ctx, nxt.tiles.data(), nxt.count, 0, spp, spp, seed));
        for (const phos_tile& c : cur.tiles) {
          allocator_scope_t scope(allocator);
          render_buffer_t buffer(fs->tiles->format);
          buffer.allocate(allocator, c.w, c.h);
          if (auto* primary = buffer.channel(render_buffer_t::PRIMARY))
            for (uint32_t y = 0; y < c.h; ++y) {
              const float* p = rgba + 4 * ((size_t)(c.y + y) * film_w + c.x);
              for (uint32_t x = 0; x < c.w; ++x, p += 4) primary->set(x, y, Imath::V3f(p[0], p[1], p[2]));
            }
          if (auto* normals = nrm ? buffer.channel(render_buffer_t::NORMALS) : nullptr)
            for (uint32_t y = 0; y < c.h; ++y) {
              const float* p = nrm + 3 * ((size_t)(c.y + y) * film_w + c.x);
              for (uint32_t x = 0; x < c.w; ++x, p += 3) normals->set(x, y, Imath::V3f(p[0], p[1], p[2]));
            }
          fs->film->add_tile(Imath::V2i(c.x, c.y), Imath::V2i(c.w, c.h), buffer);
        }
        tiles_done += cur.count;
        ++claims;
        std::swap(cur, nxt);
        have = more;
      }
    } catch (...) {
      error = std::current_exception();  // rethrown by join()
    }
  });
}

void cuda_t::join() {
  if (worker.joinable()) worker.join();
  if (error) {
    std::exception_ptr e = error;
    error = nullptr;
    std::rethrow_exception(e);
  }
}
