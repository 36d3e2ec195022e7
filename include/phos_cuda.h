/* phos_cuda.h — C ABI of libphos_cuda.so, the B200 (sm_100a) device backend for phosphorus.
 *
 * This is what the reference's GPU device slot binds.  The reference declares `cuda_t : xpu_t`
 * (src/xpu/cuda.hpp:8-13) with empty bodies (src/xpu/cuda.cpp:5-14); the drop-in fills
 * `cuda_t::{make,preprocess,start,join}` by calling the entry points below (INTEGRATION.md shows
 * the exact glue).  Plain C types only, int status everywhere (0 = PHOS_OK), no exceptions cross
 * the boundary, caller-owned host buffers, one context per GPU, one host thread per context.
 *
 * Reference interface each group replaces / serves:
 *   phos_cuda_device_count/create/destroy .. cuda_t::make + xpu_t::discover (src/xpu.cpp:7-9,
 *                                            src/xpu/cuda.cpp:10-14), options from parsed_options_t
 *                                            (src/options.hpp:6-43)
 *   phos_bvh_*  ............................ cpu_t::details_t::reset (src/xpu/cpu.cpp:35-44): the
 *                                            host binned-SAH 8-wide build, bvh::from
 *                                            (src/accel/bvh/binned_sah_builder.hpp:215-281) +
 *                                            accel::builder_t (src/accel/bvh.cpp:22-79); emits the
 *                                            reference's own node_t<8> (288 B) / packet (384 B) arrays
 *   phos_cuda_upload_accel ................. xpu_t::preprocess (src/xpu.hpp:20): takes mbvh_t::root /
 *                                            mbvh_t::triangles (src/accel/bvh.hpp:29-36) as they are
 *   phos_cuda_trace[_device] ............... stream_mbvh_kernel_t::trace(ray_t<>*, active_t<>&)
 *                                            (src/kernels/cpu/stream_bvh_kernel.hpp:19-25): same SoA
 *                                            fields, same flag semantics, in place
 *   phos_cuda_upload_scene ................. what tile_renderer_t reads from scene_t while rendering
 *                                            (src/xpu/cpu.cpp:85-99, deferred_shading_kernel.hpp, spt.hpp)
 *   phos_cuda_camera_rays .................. camera::perspective_kernel_t (src/kernels/cpu/camera.hpp:78-159)
 *   phos_cuda_render ....................... tile_renderer_t::render_tile over a list of tiles
 *                                            (src/xpu/cpu.cpp:156-205), i.e. xpu_t::start's work
 *   phos_cuda_film_* ....................... film_t<>::add_tile hand-off (src/film.hpp:10-16)
 */
#ifndef PHOS_CUDA_H
#define PHOS_CUDA_H

#include <stdint.h>

#include "phos_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PHOS_OK 0
#define PHOS_ERR_INVALID 1  /* bad argument / call order                        */
#define PHOS_ERR_CUDA 2     /* a CUDA runtime call failed (see last_error)      */
#define PHOS_ERR_NO_DEVICE 3 /* no CUDA device: there is NO CPU fallback         */
#define PHOS_ERR_ACCEL 4    /* malformed acceleration structure                 */

/* ray flag bits, src/state.hpp:33-36 */
#define PHOS_HIT 1u
#define PHOS_MASKED 2u
#define PHOS_SHADOW 4u
#define PHOS_SPECULAR 8u

typedef struct phos_ctx phos_ctx;
typedef struct phos_bvh phos_bvh;

/* the parsed_options_t fields a device reads (src/options.hpp:27-32) */
typedef struct phos_options {
  uint32_t samples_per_pixel; /* default 16 */
  uint32_t paths_per_sample;  /* default 16; film is divided by spp*pps (src/xpu/cpu.cpp:191) */
  uint32_t path_depth;        /* default 9  */
} phos_options;

/* A ray stream: the fields of ray_t<N> (src/state.hpp:39-57) as 12 SoA pointers over n rays.
 * d = tmax in / hit distance out (FLT_MAX default); mesh = meshid | matid << 16; face = 3 * face index. */
typedef struct phos_rays {
  float *px, *py, *pz;
  float *wx, *wy, *wz;
  float* d;
  uint32_t *mesh, *face;
  float *u, *v;
  uint32_t* flags;
} phos_rays;

/* job::tiles_t::tile_t (src/jobs/tiles.hpp:13-20) */
typedef struct phos_tile {
  uint32_t x, y, w, h;
} phos_tile;

typedef struct phos_accel_stats {
  uint32_t ref_nodes, ref_packets;  /* as uploaded (288 B / 384 B records)                  */
  uint32_t nodes, triangles;        /* as packed for the GPU (80 B / 48 B records)          */
  uint32_t max_depth;               /* node levels below the root, sizes the traversal stack */
  uint32_t max_leaf_triangles;
  uint64_t bytes_nodes, bytes_triangles;
  double   repack_seconds, upload_seconds;
} phos_accel_stats;

/* ---- device discovery / lifetime ------------------------------------------------------------ */
int         phos_cuda_device_count(void);
phos_ctx*   phos_cuda_create(int device, const phos_options* options); /* NULL on failure */
void        phos_cuda_destroy(phos_ctx* ctx);
const char* phos_cuda_last_error(phos_ctx* ctx); /* ctx may be NULL: error of the last failed create */

/* ---- host BVH build (the CPU build the reference keeps) --------------------------------------- */
phos_bvh*   phos_bvh_build(const phos_scene_desc* scene, int threads /* <=0: hardware concurrency */);
uint32_t    phos_bvh_num_nodes(const phos_bvh* bvh);
uint32_t    phos_bvh_num_packets(const phos_bvh* bvh);
const void* phos_bvh_nodes(const phos_bvh* bvh);   /* num_nodes   x 288 B, mbvh::node_t<8> layout        */
const void* phos_bvh_packets(const phos_bvh* bvh); /* num_packets x 384 B, moeller_trumbore_t<8> layout  */
double      phos_bvh_build_seconds(const phos_bvh* bvh);
void        phos_bvh_free(phos_bvh* bvh);

/* ---- acceleration structure upload (re-pack on host, one upload) ----------------------------- */
int phos_cuda_upload_accel(phos_ctx* ctx, const void* nodes288, uint32_t n_nodes, const void* packets384,
                           uint32_t n_packets);
int phos_cuda_accel_stats(phos_ctx* ctx, phos_accel_stats* out);
/* The packed structure built ON the device straight from the scene's triangles (no reference tree involved): Morton
 * order, binary radix tree, collapse to the same 8-wide quantised layout.  Milliseconds instead of the seconds of the
 * host build (bvh::from, src/accel/bvh/binned_sah_builder.hpp:215-281) + re-pack, at the price of a lower-quality
 * tree (more node tests per ray).  Queries return the same hits; only exact ties in t between two triangles may
 * resolve differently (tie-break = scene triangle order instead of the reference's packet order).
 * In the stats, upload_seconds = flatten + copy the scene in, repack_seconds = the device build. */
int phos_cuda_build_accel(phos_ctx* ctx, const phos_scene_desc* scene);

/* ---- ray queries ----------------------------------------------------------------------------- */
/* Trace n rays in place.  Per ray, exactly as the reference's trace(rays, active):
 *   MASKED            -> not traced, untouched;
 *   SHADOW            -> any-hit: on an accepted hit set HIT and shrink d, leave mesh/face/u/v alone;
 *   otherwise         -> closest hit: set HIT, d, mesh, face, u, v.
 * `rays` holds HOST pointers; the call is chunked so copies overlap traversal.  Blocking.
 *   - page-locked arrays from phos_cuda_host_alloc laid out as one slab at a constant stride in the order
 *     px, py, pz, wx, wy, wz, d, flags, mesh, face, u, v (or any cudaHostAlloc / cudaHostRegister memory
 *     laid out like that): the eight input arrays go up (32 B/ray), d and flags come back for every ray and
 *     the surface record only where the traversal wrote one, stored straight into the caller's arrays
 *     (8 B/ray + 16 B per closest hit);
 *   - any other host memory: all twelve arrays go up (48 B/ray: the surface record of a ray that is not
 *     hit must come back as it was) and d, flags, mesh, face, u, v come back (24 B/ray). */
int phos_cuda_trace(phos_ctx* ctx, const phos_rays* rays, uint64_t n);
/* Same, `rays` holds DEVICE pointers of this ctx's device; asynchronous on the ctx stream. */
int phos_cuda_trace_device(phos_ctx* ctx, const phos_rays* rays, uint64_t n);

/* ---- device buffers for callers that keep ray streams resident ------------------------------- */
int phos_cuda_rays_alloc(phos_ctx* ctx, uint64_t n, phos_rays* out_device_rays);
int phos_cuda_rays_free(phos_ctx* ctx, phos_rays* device_rays);
int phos_cuda_rays_upload(phos_ctx* ctx, const phos_rays* host, const phos_rays* device, uint64_t n);   /* inputs  */
int phos_cuda_rays_download(phos_ctx* ctx, const phos_rays* device, const phos_rays* host, uint64_t n); /* outputs */
int phos_cuda_synchronize(phos_ctx* ctx);
/* CUDA-event timing on the ctx stream (bench.py's device clock): begin, work, end -> milliseconds */
int phos_cuda_timer_begin(phos_ctx* ctx);
int phos_cuda_timer_end(phos_ctx* ctx, float* out_ms);
/* number of kernels this library launched on ctx since creation */
uint64_t phos_cuda_launch_count(phos_ctx* ctx);
/* Trace device-resident rays with traversal counters on: total 8-wide nodes box-tested and
 * triangles Moeller-Trumbore-tested over the batch (the N_node / N_tri of the roofline model).  Blocking. */
int phos_cuda_trace_count(phos_ctx* ctx, const phos_rays* rays, uint64_t n, uint64_t* out_nodes, uint64_t* out_tris);
/* The same launch with the warp-level step statistics of the traversal kernel (bench.py's issue-slot figures):
 * out[0] nodes box-tested, [1] triangles tested (summed over lanes), [2] warp-level node steps, [3] warp-level triangle
 * steps, [4] lanes that took part in the node steps, [5] in the triangle steps, [6] loop iterations of all warps,
 * [7] rays traced (MASKED rays are not).  Blocking. */
int phos_cuda_trace_profile(phos_ctx* ctx, const phos_rays* rays, uint64_t n, uint64_t out[8]);
/* Bind the calling host thread (and threads it spawns later) to the CPUs local to `device` (sysfs local_cpulist of
 * its PCI function), so page-locked ray arrays allocated afterwards live on the GPU's NUMA node.  One rank per GPU
 * calls this first.  Returns the number of CPUs bound to, 0 if the topology is not exposed (nothing changes). */
int phos_cuda_bind_host_to_device(int device);
/* page-locked host memory for ray streams handed to phos_cuda_trace (pageable memory works, slower) */
void* phos_cuda_host_alloc(uint64_t bytes);
void  phos_cuda_host_free(void* p);

/* Write a 256 MiB scratch buffer on the ctx stream so the 126 MB L2 holds none of the caller's data
 * (timing hygiene between benchmark iterations). */
int phos_cuda_flush_l2(phos_ctx* ctx);

/* ---- scene and frame pipeline ------------------------------------------------------------------ */
/* Upload what tile_renderer_t reads from scene_t while rendering (src/xpu/cpu.cpp:85-99): camera,
 * and for the path tracer meshes, materials, lights (pinhole and thin-lens cameras). */
int phos_cuda_upload_scene(phos_ctx* ctx, const phos_scene_desc* scene);

/* Reference-compatible normalisation.  The reference normalises camera and shadow-ray directions as
 * x * _mm256_rcp_ps(sqrt(dot)) (src/math/simd/vector.hpp:94-96,126-133): a ~12-bit reciprocal that leaves shadow rays up
 * to 3.7e-4 too long, so that a share of them hit the light's own triangle and count as occluded (spt.hpp:116-148) — its
 * images are 13-40 % darker than exact arithmetic.  on = 1: the library samples the RCPSS of the HOST it runs on into a
 * table (a function of the leading mantissa bits, verified at sampling time) and the device kernels normalise through it,
 * so cuda_t tiles match the cpu_t tiles next to them in a shared frame (plugins/blender/session.cpp:85-99).  on = 0
 * (default): correctly rounded 1 / sqrt.  The traversal is exact in both modes.  Applies to the following renders. */
int phos_cuda_reference_normalize(phos_ctx* ctx, int on);

/* camera::perspective_kernel_t (src/kernels/cpu/camera.hpp:78-159): primary rays of every pixel of
 * the given tiles for one film jitter (jx, jy) shared by the whole sample, into device ray arrays.
 * The rays of tile k start at slot sum_{j<k} w_j * h_j, row-major inside the tile (slot = y * w + x,
 * src/xpu/cpu.cpp:177); unlike the reference, partial tiles are not padded to 1024 slots. */
int phos_cuda_camera_rays(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, float jx, float jy,
                          const phos_rays* device_rays);
/* The same with the thin lens of camera_t (aperture_radius != 0, src/entities/camera.hpp:37-39;
 * camera.hpp:140-147): the lens sample of a ray is the renderer's own draw for (seed, film pixel, sample),
 * so these are the primary rays phos_cuda_render traces for that sample.  phos_cuda_camera_rays is this
 * call with seed 0, sample 0. */
int phos_cuda_camera_rays_lens(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, float jx, float jy, uint64_t seed,
                               uint32_t sample, const phos_rays* device_rays);

/* tile_renderer_t::render_tile over a list of tiles (src/xpu/cpu.cpp:156-205), as a wavefront: path-trace
 * samples [spp_begin, spp_end) of every pixel of the given tiles and accumulate
 * radiance / (spp_total * paths_per_sample) into the ctx's device film (film size = camera film size).
 * Depth and paths_per_sample come from the phos_options the ctx was created with.  Random numbers
 * are a pure function of (seed, film pixel, sample, bounce, dimension): the image does not depend on
 * how tiles or sample ranges are split over calls or GPUs.  Asynchronous on the ctx stream. */
int phos_cuda_render(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, uint32_t spp_begin, uint32_t spp_end,
                     uint32_t spp_total, uint64_t seed);

/* The ray streams of a path-traced frame, produced by the pipeline itself (BASELINE config 3's ray
 * sets): run one bounce of sample `sample` over the tiles and copy into `device_out` (capacity >= number
 * of tile pixels) either the BSDF-sampled bounce rays from the primary hits, compacted (which = 0), or the
 * next-event shadow rays, one per primary slot, SHADOW or SHADOW|MASKED (which = 1).  Blocking. */
int phos_cuda_wavefront_rays(phos_ctx* ctx, const phos_tile* tiles, uint32_t n_tiles, uint32_t sample, uint32_t spp_total,
                             uint64_t seed, int which, const phos_rays* device_out, uint64_t capacity, uint64_t* out_count);

/* ---- film: the film_t<>::add_tile hand-off (src/film.hpp:10-16) -------------------------------------- */
int phos_cuda_film_clear(phos_ctx* ctx);
/* device pointer of the W*H*4 float film (for the per-frame NCCL reduce done by the caller) */
int phos_cuda_film_device_ptr(phos_ctx* ctx, void** out_ptr, uint64_t* out_floats);
/* copy a rectangle of the film to host as interleaved RGBA (alpha = 1 where rendered): the tile
 * buffer film_t<>::add_tile receives.  Blocking. */
int phos_cuda_film_read(phos_ctx* ctx, float* rgba, uint32_t x, uint32_t y, uint32_t w, uint32_t h);
/* render_buffer_t::NORMALS (src/buffer.hpp:10, src/xpu/cpu.cpp:97,194-196): per pixel the shading normal of the
 * last sample whose primary ray hit (0 where none did).  Off by default; enabling allocates and clears a W*H*3
 * float channel that phos_cuda_render then fills; read it back as interleaved xyz.  (Not part of the multi-GPU
 * film reduce: under tile partitioning read each rank's own tiles; under sample partitioning the rank holding the
 * last sample range has the reference's values.) */
int phos_cuda_enable_normals(phos_ctx* ctx, int on);
int phos_cuda_film_read_normals(phos_ctx* ctx, float* xyz, uint32_t x, uint32_t y, uint32_t w, uint32_t h);

/* ---- multi-GPU frames: ONE NCCL reduce of the film per frame (one process per GPU) -------------
 * The reference shards a frame over the workers of one process through one tile cursor (src/jobs/tiles.hpp:40-47,
 * src/xpu/cpu.cpp:223-238) into one film_t (src/film.hpp:10-16).  With one process per GPU each rank renders its tiles
 * (or its sample range) into its own device film and the films meet in
 *     ncclReduce(film, film, W*H*4, ncclFloat, ncclSum, root)
 * enqueued on the context's stream behind the frame's kernels (asynchronous; phos_cuda_film_read on the root then sees
 * the whole frame).  Disjoint tiles make the sum a gather, weighted sample ranges make it the average; alpha is clamped
 * back to 1.  NCCL is bound at run time (dlopen "libnccl.so.2"), the library has no link-time dependency on it.
 * N cuda_t devices inside ONE process (xpu_t::discover) need none of this: each hands its own tiles to film_t::add_tile.
 *   comm_unique_id : ncclGetUniqueId on one rank; the host ships the 128 bytes to the others (any channel);
 *   comm_init      : ncclCommInitRank for this context's GPU (collective: every rank calls it);
 *   comm_adopt     : use a communicator the host already owns (an ncclComm_t whose device is this context's) instead;
 *   film_reduce    : the per-frame reduce (collective). */
#define PHOS_NCCL_ID_BYTES 128
int  phos_cuda_comm_unique_id(uint8_t id[PHOS_NCCL_ID_BYTES]);
int  phos_cuda_comm_init(phos_ctx* ctx, int n_ranks, int rank, const uint8_t id[PHOS_NCCL_ID_BYTES]);
int  phos_cuda_comm_adopt(phos_ctx* ctx, void* nccl_comm, int n_ranks);
void phos_cuda_comm_destroy(phos_ctx* ctx);
int  phos_cuda_film_reduce(phos_ctx* ctx, int root);

#ifdef __cplusplus
}
#endif
#endif /* PHOS_CUDA_H */
