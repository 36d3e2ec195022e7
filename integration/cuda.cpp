// integration/cuda.cpp — drop-in replacement for the reference's src/xpu/cuda.cpp (empty bodies,
// reference src/xpu/cuda.cpp:5-14).  Everything on the device goes through the C ABI of
// include/phos_cuda.h; this file only adapts the reference's types:
//   preprocess  = cpu_t::details_t::reset (src/xpu/cpu.cpp:35-44: the reference's own CPU BVH build)
//                 + one upload of mbvh_t::root / triangles and of the flattened scene;
//   start/join  = cpu_t::start / join (src/xpu/cpu.cpp:223-244) with ONE worker that drains the shared
//                 tile queue in chunks and hands every tile to frame.film->add_tile (src/xpu/cpu.cpp:201-204).
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
  std::vector<uint8_t> mesh_smooth;
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
  bool all_normals = true;
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
    bool smooth = mesh->num_faces > 0, flat = mesh->num_faces > 0;
    for (uint32_t face = 0; face < mesh->num_faces; ++face) (mesh->is_smooth(face) ? flat : smooth) = false;
    if (!smooth && !flat && mesh->num_faces) throw std::runtime_error("cuda_t: meshes mixing smooth and flat faces are not supported");
    f.mesh_smooth.push_back(smooth ? 1 : 0);
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

int cuda_t::device_count() { return phos_cuda_device_count(); }

cuda_t* cuda_t::make(const parsed_options_t& options, int device) {
  phos_options po{options.samples_per_pixel, options.paths_per_sample, options.path_depth};
  phos_ctx* ctx = phos_cuda_create(device, &po);
  if (!ctx) throw std::runtime_error(std::string("libphos_cuda: ") + phos_cuda_last_error(nullptr));
  auto* d = new cuda_t();
  d->ctx = ctx;
  d->spp = options.samples_per_pixel;
  return d;
}

cuda_t::~cuda_t() {
  if (worker.joinable()) worker.join();
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
}

void cuda_t::start(const scene_t&, frame_state_t& frame) {
  if (worker.joinable()) worker.join();  // a device is started once per view (session.cpp:224-229)
  check(ctx, phos_cuda_film_clear(ctx));
  // the NORMALS channel only when the frame's tile format asks for it (render_buffer_t::NORMALS, cpu.cpp:97)
  bool want_normals = false;
  for (const auto& ch : frame.tiles->format.channels) want_normals = want_normals || ch.name == render_buffer_t::NORMALS;
  check(ctx, phos_cuda_enable_normals(ctx, want_normals ? 1 : 0));
  frame_state_t* fs = &frame;
  worker = std::thread([this, fs] {
    allocator_t allocator(1024 * 1024 * 4);
    std::vector<phos_tile> chunk;
    std::vector<float> rgba, nrm;
    job::tiles_t::tile_t t;
    for (;;) {
      chunk.clear();
      while (chunk.size() < 4096 && fs->tiles->next(t)) chunk.push_back(phos_tile{t.x, t.y, t.w, t.h});
      if (chunk.empty()) break;
      check(ctx, phos_cuda_render(ctx, chunk.data(), (uint32_t)chunk.size(), 0, spp, spp, seed));
      for (const phos_tile& c : chunk) {
        allocator_scope_t scope(allocator);
        render_buffer_t buffer(fs->tiles->format);
        buffer.allocate(allocator, c.w, c.h);
        rgba.resize(4u * c.w * c.h);
        check(ctx, phos_cuda_film_read(ctx, rgba.data(), c.x, c.y, c.w, c.h));
        if (auto* primary = buffer.channel(render_buffer_t::PRIMARY))
          for (uint32_t y = 0; y < c.h; ++y)
            for (uint32_t x = 0; x < c.w; ++x) {
              const float* p = &rgba[4u * (y * c.w + x)];
              primary->set(x, y, Imath::V3f(p[0], p[1], p[2]));
            }
        if (auto* normals = buffer.channel(render_buffer_t::NORMALS)) {
          nrm.resize(3u * c.w * c.h);
          check(ctx, phos_cuda_film_read_normals(ctx, nrm.data(), c.x, c.y, c.w, c.h));
          for (uint32_t y = 0; y < c.h; ++y)
            for (uint32_t x = 0; x < c.w; ++x) {
              const float* p = &nrm[3u * (y * c.w + x)];
              normals->set(x, y, Imath::V3f(p[0], p[1], p[2]));
            }
        }
        fs->film->add_tile(Imath::V2i(c.x, c.y), Imath::V2i(c.w, c.h), buffer);
      }
    }
  });
}

void cuda_t::join() {
  if (worker.joinable()) worker.join();
}
