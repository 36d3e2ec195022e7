// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Never linked into the product.
//
// A thin extern "C" face over the UNMODIFIED reference sources compiled where they lie under
// /root/reference/src (see oracle/Makefile).  Everything that computes here is the reference's
// own code: mesh_t::builder_t (src/mesh.hpp:46-66), scene_t (src/scene.hpp), bvh::from
// (src/accel/bvh/binned_sah_builder.hpp:272-281), stream_mbvh_kernel_t / linear_mbvh_kernel_t
// (src/kernels/cpu/*_bvh_kernel.cpp), cpu_t (src/xpu/cpu.cpp).  This file only moves flat arrays
// in and out so Python tests and bench.py's reference arm can drive it through ctypes.
#include "../../include/phos_scene.h"

#include "accel/bvh.hpp"
#include "bsdf.hpp"
#include "accel/bvh/binned_sah_builder.hpp"
#include "accel/triangle.hpp"
#include "kernels/cpu/linear_bvh_kernel.hpp"
#include "kernels/cpu/stream_bvh_kernel.hpp"
#include "light.hpp"
#include "material.hpp"
#include "mesh.hpp"
#include "options.hpp"
#include "scene.hpp"
#include "state.hpp"
#include "xpu/cpu.hpp"
#include "xpu/cuda.hpp"
#include "xpu.hpp"
// after scene.hpp: the kernel header uses camera_t without including entities/camera.hpp itself
#include "kernels/cpu/camera.hpp"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct ref_scene {
  scene_t scene;
  accel::mbvh_t accel;
  bool built = false;
  double build_seconds = 0.0;
};

void fill_scene(scene_t& scene, const phos_scene_desc* d) {
  for (uint32_t i = 0; i < d->num_materials; ++i) {
    const phos_material& m = d->materials[i];
    auto* mat = new material_t();
    material_t::builder_t::scoped_t b(mat->builder());
    const char* node = m.kind == PHOS_MAT_EMITTER ? "diffuse_emitter_node"
                     : m.kind == PHOS_MAT_GLOSSY ? "glossy_bsdf_node"
                     : m.kind == PHOS_MAT_BACKGROUND ? "background_node"
                     : m.kind == PHOS_MAT_LAYERED ? "layered_node" : "diffuse_bsdf_node";
    b->shader(node, "layer0", "surface");
    b->parameter("Cs", Imath::Color3f(m.cs[0], m.cs[1], m.cs[2]));
    b->parameter("roughness", m.roughness);
    b->parameter("power", m.power);
    for (uint32_t k = 0; k < m.num_lobes && k < PHOS_MAX_LOBES; ++k) {
      const std::string p = "lobe" + std::to_string(k) + ".";
      b->parameter(p + "type", (int)m.lobes[k].type);
      b->parameter(p + "weight", Imath::Color3f(m.lobes[k].weight[0], m.lobes[k].weight[1], m.lobes[k].weight[2]));
      b->parameter(p + "param", m.lobes[k].param);
      b->parameter(p + "param2", m.lobes[k].param2);
    }
    scene.add("material" + std::to_string(i), mat);
  }
  // scene.add(light_t::make_infinite(material)) as codecs/scene.cpp:34-38 does for world.environment
  if (d->environment >= 0 && (uint32_t)d->environment < d->num_materials)
    scene.add(light_t::make_infinite(scene.material("material" + std::to_string(d->environment))));
  for (uint32_t m = 0; m < d->num_meshes; ++m) {
    auto* mesh = new mesh_t();
    {
      mesh_t::builder_t::scoped_t b(mesh->builder());
      for (uint32_t v = d->vert_offset[m]; v < d->vert_offset[m + 1]; ++v) {
        b->add_vertex(Imath::V3f(d->vertices[3 * v], d->vertices[3 * v + 1], d->vertices[3 * v + 2]));
        if (d->normals) {
          b->add_normal(Imath::V3f(d->normals[3 * v], d->normals[3 * v + 1], d->normals[3 * v + 2]));
        }
      }
      for (uint32_t f = d->face_offset[m]; f < d->face_offset[m + 1]; ++f) {
        const bool smooth = d->mesh_smooth[m] == 2 ? d->face_smooth[f] != 0 : d->mesh_smooth[m] != 0;
        b->add_face(d->faces[3 * f], d->faces[3 * f + 1], d->faces[3 * f + 2], smooth);
      }
      for (uint32_t s = d->set_offset[m]; s < d->set_offset[m + 1]; ++s) {
        std::vector<uint32_t> faces(d->set_faces + d->set_face_offset[s], d->set_faces + d->set_face_offset[s + 1]);
        b->add_face_set(d->set_material[s], faces);
      }
      b->set_normals_per_vertex();
    }
    scene.add(mesh);
  }
  camera_t& c = scene.camera;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) c.to_world.x[i][j] = d->camera.to_world[4 * i + j];
  c.fov = d->camera.fov;
  c.focal_distance = d->camera.focal_distance;
  c.aperture_radius = d->camera.aperture_radius;
  c.film.width = d->camera.film_width;
  c.film.height = d->camera.film_height;
  scene.preprocess();
}

// view of one 1024-slot window of flat SoA arrays as the reference's ray_t<1024>
void load_stream(ray_t<>* r, const phos_scene_desc*, const float* const* f, const uint32_t* const* u, size_t base,
                 uint32_t n) {
  memcpy(r->p.x, f[0] + base, n * 4);
  memcpy(r->p.y, f[1] + base, n * 4);
  memcpy(r->p.z, f[2] + base, n * 4);
  memcpy(r->wi.x, f[3] + base, n * 4);
  memcpy(r->wi.y, f[4] + base, n * 4);
  memcpy(r->wi.z, f[5] + base, n * 4);
  memcpy(r->d, f[6] + base, n * 4);
  memcpy(r->u, f[7] + base, n * 4);
  memcpy(r->v, f[8] + base, n * 4);
  memcpy(r->mesh, u[0] + base, n * 4);
  memcpy(r->face, u[1] + base, n * 4);
  memcpy(r->flags, u[2] + base, n * 4);
}

void store_stream(const ray_t<>* r, float* const* f, uint32_t* const* u, size_t base, uint32_t n) {
  memcpy(f[6] + base, r->d, n * 4);
  memcpy(f[7] + base, r->u, n * 4);
  memcpy(f[8] + base, r->v, n * 4);
  memcpy(u[0] + base, r->mesh, n * 4);
  memcpy(u[1] + base, r->face, n * 4);
  memcpy(u[2] + base, r->flags, n * 4);
}

struct memory_film_t : public film_t<> {
  float* rgba;              // W*H*4, row-major
  float* normals = nullptr; // W*H*3 when the NORMALS channel was requested
  uint32_t width, height;
  void add_tile(const Imath::V2i& pos, const Imath::V2i& size, const render_buffer_t& buffer) override {
    const auto* ch = buffer.channel(render_buffer_t::PRIMARY);
    const auto* nc = normals ? buffer.channel(render_buffer_t::NORMALS) : nullptr;
    for (int y = 0; y < size.y; ++y)
      for (int x = 0; x < size.x; ++x) {
        float px[4] = {0, 0, 0, 0};
        ch->get(x, y, px);
        float* out = rgba + 4 * ((size_t)(pos.y + y) * width + (pos.x + x));
        out[0] = px[0]; out[1] = px[1]; out[2] = px[2]; out[3] = 1.0f;
        if (nc) {
          float n[4] = {0, 0, 0, 0};
          nc->get(x, y, n);
          float* o = normals + 3 * ((size_t)(pos.y + y) * width + (pos.x + x));
          o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
        }
      }
  }
};

}  // namespace

extern "C" {

void* ref_scene_create(const phos_scene_desc* desc) {
  auto* s = new ref_scene();
  fill_scene(s->scene, desc);
  return s;
}

void ref_scene_destroy(void* h) { delete static_cast<ref_scene*>(h); }

uint32_t ref_scene_num_lights(void* h) { return static_cast<ref_scene*>(h)->scene.num_lights(); }

// cpu_t::details_t::reset (src/xpu/cpu.cpp:35-44)
double ref_accel_build(void* h) {
  auto* s = static_cast<ref_scene*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  s->accel.reset();
  {
    accel::mbvh_t::builder_t::scoped_t builder(s->accel.builder());
    std::vector<triangle_t> triangles;
    s->scene.triangles(triangles);
    bvh::from(builder, triangles);
  }
  s->built = true;
  s->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return s->build_seconds;
}

uint32_t ref_accel_num_packets(void* h) { return static_cast<ref_scene*>(h)->accel.num_triangles; }

// num_nodes is never published by the reference (src/accel/bvh.cpp:42-47); recover it by walking.
uint32_t ref_accel_num_nodes(void* h) {
  auto* s = static_cast<ref_scene*>(h);
  if (!s->accel.root) return 0;
  uint32_t max_index = 0;
  std::vector<uint32_t> stack{0};
  while (!stack.empty()) {
    const uint32_t n = stack.back();
    stack.pop_back();
    if (n > max_index) max_index = n;
    const auto& node = s->accel.root[n];
    for (int i = 0; i < 8; ++i) {
      const bool empty = node.bounds[i] > node.bounds[i + 24];
      if (!empty && node.flags[i] != 1) stack.push_back(node.offset[i]);
    }
  }
  return max_index + 1;
}

uint32_t ref_sizeof_node(void) { return sizeof(mbvh::node_t<8>); }
uint32_t ref_sizeof_packet(void) { return sizeof(accel::mbvh_t::triangle_t); }

void ref_accel_copy(void* h, void* nodes, uint32_t n_nodes, void* packets, uint32_t n_packets) {
  auto* s = static_cast<ref_scene*>(h);
  memcpy(nodes, s->accel.root, (size_t)n_nodes * sizeof(mbvh::node_t<8>));
  memcpy(packets, s->accel.triangles, (size_t)n_packets * sizeof(accel::mbvh_t::triangle_t));
}

// Trace n rays given as flat SoA arrays, cut into 1024-ray streams (config::STREAM_SIZE), on
// `threads` host threads each owning one kernel instance and pulling streams from an atomic cursor
// — the shape of cpu_t::start (src/xpu/cpu.cpp:223-238).  kind 0 = stream_mbvh_kernel_t,
// 1 = linear_mbvh_kernel_t (brute force).  f[0..8] = px,py,pz,wx,wy,wz,d,u,v ; u[0..2] = mesh,face,flags.
// Returns seconds spent tracing (excludes thread start-up of kernel state allocation).
double ref_trace(void* h, int kind, float* const* f, uint32_t* const* u, uint64_t n, int threads) {
  auto* s = static_cast<ref_scene*>(h);
  const uint64_t streams = (n + 1023) / 1024;
  std::atomic<uint64_t> cursor(0);
  if (threads < 1) threads = 1;
  std::atomic<int> ready(0);
  std::atomic<bool> go(false);
  std::vector<std::thread> pool;
  std::chrono::steady_clock::time_point t0;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, kind]() {
      stream_mbvh_kernel_t stream_kernel(&s->accel);
      linear_mbvh_kernel_t linear_kernel(&s->accel);
      ray_t<>* rays;
      posix_memalign((void**)&rays, 32, sizeof(ray_t<>));
      active_t<> active;
      ++ready;
      while (!go.load()) std::this_thread::yield();
      for (;;) {
        const uint64_t i = cursor++;
        if (i >= streams) break;
        const size_t base = (size_t)i * 1024;
        const uint32_t cnt = (uint32_t)std::min<uint64_t>(1024, n - base);
        load_stream(rays, nullptr, f, u, base, cnt);
        active.reset(0);
        active.num = cnt;
        if (kind == 0) {
          stream_kernel.trace(rays, active);
        } else {
          // the linear kernel has no MASKED filter of its own (src/kernels/cpu/linear_bvh_kernel.cpp:14-19):
          // hand it only the unmasked slots, as lanes_t::init does for the stream kernel.
          active_t<> unmasked;
          for (uint32_t k = 0; k < cnt; ++k)
            if (!rays->is_masked(k)) unmasked.add(k);
          linear_kernel.trace(rays, unmasked);
        }
        store_stream(rays, f, u, base, cnt);
      }
      free(rays);
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  t0 = std::chrono::steady_clock::now();
  go.store(true);
  for (auto& th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Render a frame with the reference's own CPU device (cpu_t::preprocess/start/join) into an RGBA
// float image.  Wall clock around start -> join like src/core.cpp:158-177.  Returns seconds.
double ref_render_on(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int single_threaded, int use_cuda, float* rgba);
double ref_render_aov(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int single_threaded, int use_cuda, float* rgba, float* normals);
double ref_render(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int single_threaded, float* rgba) {
  return ref_render_on(h, spp, pps, depth, single_threaded, 0, rgba);
}

// The same frame through ANY xpu_t: use_cuda = 1 drives the drop-in GPU device (integration/cuda.cpp -> libphos_cuda.so)
// with exactly the calls session_t::details_t::render makes (plugins/blender/session.cpp:73-94).
// Returns seconds, or -1 with a message on stderr if the device raised.
double ref_render_on(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int single_threaded, int use_cuda, float* rgba) {
  return ref_render_aov(h, spp, pps, depth, single_threaded, use_cuda, rgba, nullptr);
}

// ... and with the NORMALS channel (render_buffer_t::NORMALS, src/xpu/cpu.cpp:97,194-196) when `normals` is given
double ref_render_aov(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int single_threaded, int use_cuda, float* rgba,
                      float* normals) {
  auto* s = static_cast<ref_scene*>(h);
  parsed_options_t options;
  options.samples_per_pixel = spp;
  options.paths_per_sample = pps;
  options.path_depth = depth;
  options.single_threaded = single_threaded != 0;
  options.host_only = true;

  const uint32_t W = s->scene.camera.film.width, H = s->scene.camera.film.height;
  render_buffer_t::descriptor_t format;
  format.request(render_buffer_t::PRIMARY, 4);
  if (normals) format.request(render_buffer_t::NORMALS, 3);

  xpu_t* device = nullptr;
  try {
#ifdef REF_WITH_CUDA
    device = use_cuda ? static_cast<xpu_t*>(cuda_t::make(options, 0)) : static_cast<xpu_t*>(cpu_t::make(options));
#else
    if (use_cuda) throw std::runtime_error("this build holds the reference alone (load libphos_ref_cuda.so for cuda_t)");
    device = cpu_t::make(options);
#endif
    device->preprocess(s->scene);
  } catch (const std::exception& e) {
    std::cerr << "ref_render_on: " << e.what() << std::endl;
    delete device;
    return -1.0;
  }

  job::tiles_t* tiles = job::tiles_t::make(W, H, 32, format);
  memory_film_t film;
  film.rgba = rgba;
  film.normals = normals;
  film.width = W;
  film.height = H;
  sampler_t* sampler = new sampler_t(options);
  frame_state_t state(sampler, tiles, &film);
  sampler->preprocess(s->scene);

  const auto t0 = std::chrono::steady_clock::now();
  device->start(s->scene, state);
  device->join();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  delete device;
  delete tiles;
  // sampler_t::~sampler_t deletes posix_memalign'd memory (src/sampling.cpp:83-87); leak it instead.
  return dt;
}

#ifdef REF_WITH_CUDA
// Several frames on ONE cuda_t that is made and preprocessed once and then started / joined per frame — what
// session_t does per view (plugins/blender/session.cpp:224-229: prepare_devices once, render per view).  A fresh tile
// queue, sampler and (zeroed) film per frame; secs[i] = seconds around start..join of frame i, rgba holds the last
// frame.  Frame 0 is the cold one (the device allocates its wavefront state and page-locked read-back slabs in it).
// Returns 0, or -1 with a message on stderr if the device raised.
int ref_render_frames_cuda(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int frames, float* rgba, double* secs) {
  auto* s = static_cast<ref_scene*>(h);
  parsed_options_t options;
  options.samples_per_pixel = spp;
  options.paths_per_sample = pps;
  options.path_depth = depth;
  options.single_threaded = true;
  options.host_only = true;
  const uint32_t W = s->scene.camera.film.width, H = s->scene.camera.film.height;
  render_buffer_t::descriptor_t format;
  format.request(render_buffer_t::PRIMARY, 4);
  xpu_t* device = nullptr;
  int rc = 0;
  try {
    device = cuda_t::make(options, 0);
    device->preprocess(s->scene);
    for (int f = 0; f < frames; ++f) {
      std::memset(rgba, 0, sizeof(float) * 4 * (size_t)W * H);
      job::tiles_t* tiles = job::tiles_t::make(W, H, 32, format);
      memory_film_t film;
      film.rgba = rgba;
      film.normals = nullptr;
      film.width = W;
      film.height = H;
      sampler_t* sampler = new sampler_t(options);  // (leaked like everywhere in this file: its destructor is broken)
      frame_state_t state(sampler, tiles, &film);
      sampler->preprocess(s->scene);
      const auto t0 = std::chrono::steady_clock::now();
      device->start(s->scene, state);
      device->join();
      secs[f] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      delete tiles;
    }
  } catch (const std::exception& e) {
    std::cerr << "ref_render_frames_cuda: " << e.what() << std::endl;
    rc = -1;
  }
  delete device;
  return rc;
}

int ref_cuda_device_count(void) { return cuda_t::device_count(); }

// How many devices xpu_t::discover (src/xpu.cpp:7-9 + integration/xpu_discover.patch) returns, and how many are GPUs.
int ref_discover(int host_only, int* out_cuda) {
  parsed_options_t options;
  options.host_only = host_only != 0;
  options.samples_per_pixel = 1;
  options.paths_per_sample = 1;
  const int before = cuda_t::instances.load();
  std::vector<xpu_t*> devices = xpu_t::discover(options);
  const int n_cuda = cuda_t::instances.load() - before;  // (built without RTTI: count the live cuda_t devices)
  for (xpu_t* d : devices) delete d;
  if (out_cuda) *out_cuda = n_cuda;
  return (int)devices.size();
}

// One frame on SEVERAL devices that share frame.tiles, exactly the loops of session_t::details_t::render / prepare_devices
// (plugins/blender/session.cpp:85-99,121-132): n_cuda cuda_t devices (GPU i % device_count) and, with_cpu, the reference's
// own cpu_t next to them.  tiles_per_device[i] receives the tiles cuda device i rendered (the rest went to cpu_t).
// Returns seconds around start..join, or -1 if a device raised (message on stderr).
double ref_render_devices(void* h, uint32_t spp, uint32_t pps, uint32_t depth, int n_cuda, int with_cpu, int cpu_single_threaded,
                          float* rgba, int* tiles_per_device) {
  auto* s = static_cast<ref_scene*>(h);
  parsed_options_t options;
  options.samples_per_pixel = spp;
  options.paths_per_sample = pps;
  options.path_depth = depth;
  options.single_threaded = cpu_single_threaded != 0;
  options.host_only = false;
  const uint32_t W = s->scene.camera.film.width, H = s->scene.camera.film.height;
  render_buffer_t::descriptor_t format;
  format.request(render_buffer_t::PRIMARY, 4);
  std::vector<xpu_t*> devices;
  std::vector<cuda_t*> gpus;
  double dt = -1.0;
  try {
    const int n_gpu = cuda_t::device_count();
    for (int i = 0; i < n_cuda; ++i) {
      gpus.push_back(cuda_t::make(options, n_gpu > 0 ? i % n_gpu : 0));
      devices.push_back(gpus.back());
    }
    if (with_cpu) devices.push_back(cpu_t::make(options));
    for (xpu_t* d : devices) d->preprocess(s->scene);
    job::tiles_t* tiles = job::tiles_t::make(W, H, 32, format);
    memory_film_t film;
    film.rgba = rgba;
    film.width = W;
    film.height = H;
    sampler_t* sampler = new sampler_t(options);
    frame_state_t state(sampler, tiles, &film);
    sampler->preprocess(s->scene);
    const auto t0 = std::chrono::steady_clock::now();
    for (xpu_t* d : devices) d->start(s->scene, state);
    for (xpu_t* d : devices) d->join();
    dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t i = 0; tiles_per_device && i < gpus.size(); ++i) tiles_per_device[i] = (int)gpus[i]->tiles_done;
    delete tiles;
  } catch (const std::exception& e) {
    std::cerr << "ref_render_devices: " << e.what() << std::endl;
    dt = -1.0;
  }
  for (xpu_t* d : devices) delete d;
  return dt;
}

// cuda_t::join must surface what the worker threw (a throw out of a std::thread would terminate the process): start a
// frame whose tile queue holds a tile outside the film.  Returns 1 if join() raised, 0 if it did not.
int ref_cuda_join_raises(void* h) {
  auto* s = static_cast<ref_scene*>(h);
  parsed_options_t options;
  options.samples_per_pixel = 1;
  options.paths_per_sample = 1;
  options.path_depth = 2;
  options.host_only = false;
  const uint32_t W = s->scene.camera.film.width, H = s->scene.camera.film.height;
  render_buffer_t::descriptor_t format;
  format.request(render_buffer_t::PRIMARY, 4);
  cuda_t* device = cuda_t::make(options, 0);
  int raised = 0;
  std::vector<float> rgba((size_t)W * H * 4);
  try {
    device->preprocess(s->scene);
    job::tiles_t* tiles = job::tiles_t::make(W, H, 32, format);
    tiles->tiles[0].x = W;  // outside the film: phos_cuda_render rejects it inside the worker
    memory_film_t film;
    film.rgba = rgba.data();
    film.width = W;
    film.height = H;
    sampler_t* sampler = new sampler_t(options);
    frame_state_t state(sampler, tiles, &film);
    sampler->preprocess(s->scene);
    device->start(s->scene, state);
    try {
      device->join();
    } catch (const std::exception&) {
      raised = 1;
    }
    delete tiles;
  } catch (const std::exception& e) {
    std::cerr << "ref_cuda_join_raises: " << e.what() << std::endl;
    raised = -1;
  }
  delete device;
  return raised;
}
#else
int ref_cuda_device_count(void) { return 0; }
#endif

uint32_t ref_hardware_concurrency(void) { return std::thread::hardware_concurrency(); }


// material_t::evaluate(allocator, hits, active) for one slot with shading normal n -> the slot's bsdf_t
static bsdf_t* build_bsdf(ref_scene* s, uint32_t mat, const float* n, allocator_t& allocator) {
  interaction_t<>* hits = new (allocator) interaction_t<>();
  hits->n.from(0, Imath::V3f(n[0], n[1], n[2]));
  active_t<> active;
  active.reset(0);
  active.num = 1;
  active.index[0] = 0;
  s->scene.material(mat)->evaluate(allocator, hits, active);
  return hits->bsdf[0];
}

// bsdf_t::f / bsdf_t::sample (src/bsdf.cpp:113-248) of material `mat` of the scene at shading normal n: the material
// builds its bsdf_t exactly as in a render (material_t::evaluate -> add_lobe + precompute), the rest is the
// reference's own code.
void ref_bsdf_f(ref_scene* s, uint32_t mat, const float* n, const float* wi, const float* wo, float* out3) {
  allocator_t allocator(1 << 22);
  bsdf_t* bsdf = build_bsdf(s, mat, n, allocator);
  const auto f = bsdf->f(Imath::V3f(wi[0], wi[1], wi[2]), Imath::V3f(wo[0], wo[1], wo[2]));
  out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
}
int ref_bsdf_sample(ref_scene* s, uint32_t mat, const float* n, const float* wi, float sx, float sy, float* wo3, float* f3,
                    float* pdf, uint32_t* flags) {
  allocator_t allocator(1 << 22);
  bsdf_t* bsdf = build_bsdf(s, mat, n, allocator);
  Imath::V3f wo(0.0f);
  float p = 0.0f;  // a lobe that bails out early leaves the caller's pdf untouched: 0 here
  uint32_t fl = 0;
  const auto f = bsdf->sample(Imath::V2f(sx, sy), Imath::V3f(wi[0], wi[1], wi[2]), wo, p, fl);
  wo3[0] = wo.x; wo3[1] = wo.y; wo3[2] = wo.z;
  f3[0] = f.x; f3[1] = f.y; f3[2] = f.z;
  *pdf = p;
  *flags = fl;
  return !(f.x == 0.0f && f.y == 0.0f && f.z == 0.0f) && p != 0.0f;
}

// camera::perspective_kernel_t (src/kernels/cpu/camera.hpp:78-159) on one tile with caller-chosen samples:
// one film jitter for the whole tile, one lens sample per slot (slot = y * tile.w + x; tile.w % 8 == 0,
// tile.w * tile.h <= 1024).  Writes p / wi of the tile's slots.
void ref_camera_rays(ref_scene* s, uint32_t tx, uint32_t ty, uint32_t tw, uint32_t th, float jx, float jy,
                     const float* lens_x, const float* lens_y, float* px, float* py, float* pz, float* wx, float* wy,
                     float* wz) {
  struct tile_t { uint32_t x, y, w, h; } tile = {tx, ty, tw, th};
  sampler_t::pixel_samples_t* samples;
  posix_memalign((void**)&samples, 32, sizeof(sampler_t::pixel_samples_t));
  const uint32_t n = tw * th;
  for (uint32_t j = 0; j < 128; ++j)
    for (uint32_t k = 0; k < 8; ++k) {
      samples->film[j].x[k] = jx;
      samples->film[j].y[k] = jy;
      const uint32_t slot = j * 8 + k;
      samples->lens[j].x[k] = slot < n ? lens_x[slot] : 0.5f;
      samples->lens[j].y[k] = slot < n ? lens_y[slot] : 0.5f;
    }
  ray_t<>* rays;
  posix_memalign((void**)&rays, 32, sizeof(ray_t<>));
  camera::perspective_kernel_t kernel;
  kernel(s->scene.camera, tile, *samples, rays);
  memcpy(px, rays->p.x, n * 4);
  memcpy(py, rays->p.y, n * 4);
  memcpy(pz, rays->p.z, n * 4);
  memcpy(wx, rays->wi.x, n * 4);
  memcpy(wy, rays->wi.y, n * 4);
  memcpy(wz, rays->wi.z, n * 4);
  free(rays);
  free(samples);
}
}  // extern "C"
