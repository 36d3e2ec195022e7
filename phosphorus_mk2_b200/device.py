"""The B200 device behind the renderer's device interface.

Mirrors the reference's ``xpu_t`` (src/xpu.hpp:12-40) and its GPU slot ``cuda_t``
(src/xpu/cuda.hpp:8-13): ``discover`` / ``make`` / ``preprocess`` / ``start`` / ``join`` keep their
names, argument meaning and error behaviour (failures raise, as the reference throws
``std::runtime_error``).  ``trace`` is the kernel functor the reference's tile renderer calls
(``stream_mbvh_kernel_t::trace(rays, active)``, src/kernels/cpu/stream_bvh_kernel.hpp:19-25).
Everything computes inside libphos_cuda.so; this file only moves pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass

import numpy as np

from . import lib as _lib
from .lib import PhosAccelStats, PhosError, PhosOptions, PhosTile
from .rays import PhosRays, RayBatch
from .scene import Scene

NODE_BYTES, PACKET_BYTES = 288, 384


@dataclass
class Options:
    """parsed_options_t (src/options.hpp:6-43): the fields a device reads."""

    samples_per_pixel: int = 16
    paths_per_sample: int = 16
    path_depth: int = 9
    single_threaded: bool = False
    host_only: bool = False


def make_tiles(width: int, height: int, tile_size: int = 32) -> list[tuple[int, int, int, int]]:
    """job::tiles_t::make (src/jobs/tiles.hpp:49-89): row-major list of (x, y, w, h), partial tiles
    on the right / bottom edges."""
    out = []
    for y in range(0, height, tile_size):
        for x in range(0, width, tile_size):
            out.append((x, y, min(tile_size, width - x), min(tile_size, height - y)))
    return out


def tile_array(tiles) -> "C.Array":
    arr = (PhosTile * len(tiles))()
    for i, (x, y, w, h) in enumerate(tiles):
        arr[i].x, arr[i].y, arr[i].w, arr[i].h = x, y, w, h
    return arr


class Accel:
    """accel::mbvh_t (src/accel/bvh.hpp:17-49): the host-built 8-wide BVH in the reference's own
    node (288 B) / packet (384 B) layout, produced by the library's host builder."""

    def __init__(self, scene: Scene, threads: int = 0):
        self._L = _lib.load()
        d = scene.desc()
        self._h = self._L.phos_bvh_build(C.byref(d), threads)
        if not self._h:
            raise PhosError("phos_bvh_build failed")
        self.num_nodes = self._L.phos_bvh_num_nodes(self._h)
        self.num_packets = self._L.phos_bvh_num_packets(self._h)
        self.build_seconds = self._L.phos_bvh_build_seconds(self._h)

    @property
    def root(self) -> int:  # mbvh_t::root
        return self._L.phos_bvh_nodes(self._h)

    @property
    def triangles(self) -> int:  # mbvh_t::triangles
        return self._L.phos_bvh_packets(self._h)

    def nodes_array(self) -> np.ndarray:
        return np.ctypeslib.as_array(C.cast(self.root, C.POINTER(C.c_uint8)), (self.num_nodes * NODE_BYTES,)).copy()

    def packets_array(self) -> np.ndarray:
        return np.ctypeslib.as_array(C.cast(self.triangles, C.POINTER(C.c_uint8)), (self.num_packets * PACKET_BYTES,)).copy()

    def __del__(self):
        try:
            if self._h:
                self._L.phos_bvh_free(self._h)
                self._h = None
        except Exception:
            pass


class DeviceRays:
    """A ray stream resident in HBM (12 SoA arrays)."""

    def __init__(self, dev: "CudaDevice", n: int):
        self.dev, self.n = dev, int(n)
        self.s = PhosRays()
        dev._check(dev._L.phos_cuda_rays_alloc(dev._ctx, self.n, C.byref(self.s)))

    def upload(self, host: RayBatch):
        assert host.n == self.n
        hs = host.as_struct()
        self.dev._check(self.dev._L.phos_cuda_rays_upload(self.dev._ctx, C.byref(hs), C.byref(self.s), self.n))

    def download(self, host: RayBatch | None = None) -> RayBatch:
        host = host if host is not None else RayBatch(self.n)
        hs = host.as_struct()
        self.dev._check(self.dev._L.phos_cuda_rays_download(self.dev._ctx, C.byref(self.s), C.byref(hs), self.n))
        return host

    def free(self):
        if self.s.px:
            self.dev._L.phos_cuda_rays_free(self.dev._ctx, C.byref(self.s))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CudaDevice:
    """cuda_t : xpu_t — one B200."""

    def __init__(self, options: Options, device: int = 0, lib_path: str | None = None):
        self._L = _lib.load(lib_path)
        self.options = options
        self.device = device
        o = PhosOptions(options.samples_per_pixel, options.paths_per_sample, options.path_depth)
        self._ctx = self._L.phos_cuda_create(device, C.byref(o))
        if not self._ctx:
            raise PhosError(self._L.phos_cuda_last_error(None).decode())
        self.accel = None
        self.has_comm = False

    # ---- xpu_t -----------------------------------------------------------------------------------
    @staticmethod
    def discover(options: Options) -> list["CudaDevice"]:
        """xpu_t::discover (src/xpu.cpp:7-9): one device per visible GPU; empty under host_only."""
        if options.host_only:
            return []
        return [CudaDevice.make(options, i) for i in range(_lib.load().phos_cuda_device_count())]

    @staticmethod
    def make(options: Options, device: int = 0) -> "CudaDevice":
        return CudaDevice(options, device)

    def preprocess(self, scene: Scene, accel: Accel | None = None) -> None:
        """cpu_t::preprocess (src/xpu/cpu.cpp:219-221 -> details_t::reset :35-44): build the 8-wide
        BVH on the host, then re-pack and upload it once."""
        if os.environ.get("PHOS_ACCEL_BUILDER") == "device":  # the structure built on the GPU instead (build_accel)
            self.accel = accel
            self.build_accel(scene)
            return
        self.accel = accel if accel is not None else Accel(scene)
        self.upload_accel(self.accel.root, self.accel.num_nodes, self.accel.triangles, self.accel.num_packets)

    def upload_accel(self, nodes, n_nodes: int, packets, n_packets: int) -> None:
        """Upload mbvh_t::root / mbvh_t::triangles given as addresses or uint8 numpy arrays."""
        if isinstance(nodes, np.ndarray):
            self._keep = (nodes, packets)
            nodes, packets = nodes.ctypes.data, packets.ctypes.data
        self._check(self._L.phos_cuda_upload_accel(self._ctx, nodes, n_nodes, packets, n_packets))

    def build_accel(self, scene: Scene) -> None:
        """The packed structure built on the device straight from the scene's triangles (Morton order + radix tree +
        8-wide collapse): milliseconds instead of the host build + re-pack, a lower-quality tree, the same hits."""
        d = scene.desc()
        self._check(self._L.phos_cuda_build_accel(self._ctx, C.byref(d)))

    def accel_stats(self) -> PhosAccelStats:
        s = PhosAccelStats()
        self._check(self._L.phos_cuda_accel_stats(self._ctx, C.byref(s)))
        return s

    # ---- the kernel functor ------------------------------------------------------------------------
    def trace(self, rays: RayBatch) -> RayBatch:
        """trace(rays, active) in place on a host ray stream (copies in, traces, copies out)."""
        s = rays.as_struct()
        self._check(self._L.phos_cuda_trace(self._ctx, C.byref(s), rays.n))
        return rays

    def trace_device(self, rays: DeviceRays) -> None:
        self._check(self._L.phos_cuda_trace_device(self._ctx, C.byref(rays.s), rays.n))

    def trace_device_n(self, rays: DeviceRays, n: int) -> None:
        """Trace the first n rays of a device stream."""
        assert n <= rays.n
        self._check(self._L.phos_cuda_trace_device(self._ctx, C.byref(rays.s), n))

    def trace_count(self, rays: DeviceRays) -> tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.phos_cuda_trace_count(self._ctx, C.byref(rays.s), rays.n, C.byref(a), C.byref(b)))
        return a.value, b.value

    def trace_profile(self, rays: DeviceRays, n: int | None = None) -> dict:
        """One launch of the counting instantiation: per-lane fetch counts and warp-level step statistics."""
        out = (C.c_uint64 * 8)()
        self._check(self._L.phos_cuda_trace_profile(self._ctx, C.byref(rays.s), rays.n if n is None else n, out))
        keys = ("nodes", "tris", "warp_node_steps", "warp_tri_steps", "lanes_node_steps", "lanes_tri_steps", "warp_iterations", "rays")
        return dict(zip(keys, (int(v) for v in out)))

    def upload_scene(self, scene: Scene) -> None:
        d = scene.desc()
        self._check(self._L.phos_cuda_upload_scene(self._ctx, C.byref(d)))
        self._film_wh = (scene.camera.film_width, scene.camera.film_height)

    def camera_rays(self, tiles, rays: DeviceRays, jx: float = 0.5, jy: float = 0.5, seed: int = 0, sample: int = 0) -> None:
        """camera::perspective_kernel_t over a list of (x, y, w, h) tiles into a device stream; with a thin-lens
        camera the lens samples are the renderer's draws for (seed, pixel, sample)."""
        arr = tile_array(tiles)
        assert sum(t[2] * t[3] for t in tiles) <= rays.n
        self._check(self._L.phos_cuda_camera_rays_lens(self._ctx, arr, len(tiles), jx, jy, seed, sample, C.byref(rays.s)))

    def render(self, tiles, spp_begin: int, spp_end: int, spp_total: int, seed: int = 0) -> None:
        """tile_renderer_t::render_tile over a tile list (src/xpu/cpu.cpp:156-205): accumulate samples
        [spp_begin, spp_end) of the tiles' pixels into the device film.  Asynchronous."""
        arr = tile_array(tiles)
        self._check(self._L.phos_cuda_render(self._ctx, arr, len(tiles), spp_begin, spp_end, spp_total, seed))

    def wavefront_rays(self, tiles, rays: DeviceRays, which: str = "bounce", sample: int = 0, spp_total: int = 1,
                       seed: int = 0) -> int:
        """One bounce of the pipeline; `rays` receives the bounce-ray stream (compacted) or the
        next-event shadow-ray stream.  Returns the number of rays written."""
        arr = tile_array(tiles)
        n = C.c_uint64(0)
        self._check(self._L.phos_cuda_wavefront_rays(self._ctx, arr, len(tiles), sample, spp_total, seed,
                                                     0 if which == "bounce" else 1, C.byref(rays.s), rays.n, C.byref(n)))
        return int(n.value)

    def film_clear(self) -> None:
        self._check(self._L.phos_cuda_film_clear(self._ctx))

    def film_read(self, x: int = 0, y: int = 0, w: int | None = None, h: int | None = None) -> np.ndarray:
        """The RGBA float tile buffer film_t<>::add_tile would receive for the rectangle."""
        w = self._film_wh[0] - x if w is None else w
        h = self._film_wh[1] - y if h is None else h
        out = np.zeros((h, w, 4), np.float32)
        self._check(self._L.phos_cuda_film_read(self._ctx, out.ctypes.data, x, y, w, h))
        return out

    def reference_normalize(self, on: bool = True) -> None:
        """Normalise camera / shadow-ray directions through the host's RCPSS, as the reference's vector3_t<8>::normalize
        does (src/math/simd/vector.hpp:126-133): images then match the reference's cpu_t instead of exact arithmetic."""
        self._check(self._L.phos_cuda_reference_normalize(self._ctx, 1 if on else 0))

    def enable_normals(self, on: bool = True) -> None:
        """Request the NORMALS channel (render_buffer_t::NORMALS, cpu.cpp:97,194-196) for the following renders."""
        self._check(self._L.phos_cuda_enable_normals(self._ctx, 1 if on else 0))

    def film_read_normals(self, x: int = 0, y: int = 0, w: int | None = None, h: int | None = None) -> np.ndarray:
        w = self._film_wh[0] - x if w is None else w
        h = self._film_wh[1] - y if h is None else h
        out = np.zeros((h, w, 3), np.float32)
        self._check(self._L.phos_cuda_film_read_normals(self._ctx, out.ctypes.data, x, y, w, h))
        return out

    def film_device_ptr(self) -> tuple[int, int]:
        p, n = C.c_void_p(0), C.c_uint64(0)
        self._check(self._L.phos_cuda_film_device_ptr(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- multi-GPU: the per-frame film reduce (one process per GPU) ---------------------------------------
    def comm_init(self, dist) -> None:
        """Create this context's NCCL communicator over the ranks of an initialised ``torch.distributed`` group: rank 0
        draws the unique id (ncclGetUniqueId), the group ships its 128 bytes, every rank joins (ncclCommInitRank)."""
        ident = (C.c_uint8 * 128)()
        if dist.get_rank() == 0:
            self._check(self._L.phos_cuda_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self._check(self._L.phos_cuda_comm_init(self._ctx, dist.get_world_size(), dist.get_rank(), ident))
        self.has_comm = True

    def film_reduce(self, root: int = 0) -> None:
        """ncclReduce(sum) of the device film onto `root`, enqueued behind the frame's kernels (asynchronous)."""
        self._check(self._L.phos_cuda_film_reduce(self._ctx, root))

    def flush_l2(self) -> None:
        self._check(self._L.phos_cuda_flush_l2(self._ctx))

    def device_rays(self, n: int) -> DeviceRays:
        return DeviceRays(self, n)

    def synchronize(self) -> None:
        self._check(self._L.phos_cuda_synchronize(self._ctx))

    def timer_begin(self) -> None:
        self._check(self._L.phos_cuda_timer_begin(self._ctx))

    def timer_end(self) -> float:
        ms = C.c_float(0)
        self._check(self._L.phos_cuda_timer_end(self._ctx, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self._L.phos_cuda_launch_count(self._ctx))

    # ---- plumbing ------------------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise PhosError(f"libphos_cuda error {rc}: {self._L.phos_cuda_last_error(self._ctx).decode()}")

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._L.phos_cuda_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_ray_batch(n: int) -> RayBatch:
    """A RayBatch whose arrays live in one page-locked slab (fast, truly asynchronous copies).  The slab is released by
    ``batch.free()`` or when the batch is collected; ``copy()`` / ``slice()`` return ordinary pageable batches."""
    L = _lib.load()
    r = RayBatch(0)
    r.n = int(n)
    stride = (n * 4 + 255) // 256 * 256
    base = L.phos_cuda_host_alloc(max(stride, 256) * 12)
    if not base:
        raise PhosError("phos_cuda_host_alloc failed")
    r._pinned = base
    r._release = weakref.finalize(r, L.phos_cuda_host_free, base)  # the arrays are views: do not keep them past the batch
    r.free = r._release
    # one slab, constant stride, in the library's own array order: phos_cuda_trace then moves a chunk with a
    # single pitched copy per direction (see csrc/phos_cuda.cu)
    order = ("px", "py", "pz", "wx", "wy", "wz", "d", "flags", "mesh", "face", "u", "v")
    for i, k in enumerate(order):
        ct = C.c_uint32 if k in ("mesh", "face", "flags") else C.c_float
        arr = np.ctypeslib.as_array(C.cast(base + stride * i, C.POINTER(ct)), (n,))
        setattr(r, k, arr)
    return r
