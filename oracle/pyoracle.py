"""ctypes loaders for the test oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this
module; the product package never does.

  Oracle  — oracle/libphos_oracle.so, the plain-C restatement (phos_oracle.c).
  RefLib  — oracle/_ref/libphos_ref.so, the reference's own sources compiled from /root/reference
            (present when `make -C oracle ref` ran in a container that has the reference; the
            built .so travels to the GPU box, the sources do not).  RefLib(cuda=True) loads
            oracle/_ref/libphos_ref_cuda.so instead: the same objects plus the drop-in GPU device
            (integration/cuda.cpp, integration/xpu_discover.patch) linked against libphos_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from phosphorus_mk2_b200.rays import PhosRays, RayBatch
from phosphorus_mk2_b200.scene import PhosSceneDesc, Scene

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libphos_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libphos_ref.so")            # the reference alone: maps no product code
REF_CUDA_SO = os.path.join(HERE, "_ref", "libphos_ref_cuda.so")  # + the drop-in cuda_t (integration/) -> libphos_cuda.so
NODE_BYTES, PACKET_BYTES = 288, 384


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes", C.c_uint64), ("packets", C.c_uint64),
                ("triangles", C.c_uint64), ("max_stack", C.c_uint64)]


def build(target: str = "oracle") -> None:
    subprocess.run(["make", "-C", HERE, target], check=True, capture_output=True)


class Oracle:
    def __init__(self):
        build("oracle")  # no-op when up to date
        self.lib = L = C.CDLL(ORACLE_SO)
        L.orc_brute_force.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(PhosRays), C.c_uint64]
        L.orc_traverse.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(PhosRays), C.c_uint64, C.POINTER(Counters), C.c_int]
        L.orc_camera_rays.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.POINTER(PhosRays)]
        L.orc_sizeof_node.restype = C.c_uint32
        L.orc_sizeof_packet.restype = C.c_uint32
        assert L.orc_sizeof_node() == NODE_BYTES and L.orc_sizeof_packet() == PACKET_BYTES
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(PhosSceneDesc)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_num_lights.argtypes = [C.c_void_p]
        L.orc_scene_num_lights.restype = C.c_uint32
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_uint32] * 9 + [C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_rnd.restype = C.c_float
        L.orc_rnd.argtypes = [C.c_uint32] * 5
        L.orc_film_jitter.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_bsdf_f.argtypes = [C.c_void_p] * 5
        L.orc_bsdf_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float] + [C.c_void_p] * 4
        L.orc_bsdf_sample.restype = C.c_int
        L.orc_camera_rays_lens.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6

    def brute_force(self, packets: np.ndarray, rays: RayBatch) -> RayBatch:
        out = rays.copy()
        s = out.as_struct()
        self.lib.orc_brute_force(packets.ctypes.data, len(packets) // PACKET_BYTES, C.byref(s), out.n)
        return out

    def traverse(self, nodes: np.ndarray, packets: np.ndarray, rays: RayBatch, ties: bool = True):
        out = rays.copy()
        s = out.as_struct()
        c = Counters()
        self.lib.orc_traverse(nodes.ctypes.data, packets.ctypes.data, C.byref(s), out.n, C.byref(c), 1 if ties else 0)
        return out, c

    def camera_rays(self, cam, x0=0, y0=0, w=None, h=None, jx=0.5, jy=0.5) -> RayBatch:
        w = cam.film_width if w is None else w
        h = cam.film_height if h is None else h
        out = RayBatch(w * h)
        s = out.as_struct()
        m = (C.c_float * 16)(*np.asarray(cam.to_world, np.float32).ravel())
        self.lib.orc_camera_rays(m, cam.fov, cam.film_width, cam.film_height, x0, y0, w, h, jx, jy, C.byref(s))
        return out


    def render(self, scene: Scene, nodes: np.ndarray, packets: np.ndarray, spp: int, pps: int = 1, depth: int = 9,
               seed: int = 0, region=None, spp_range=None, rcp_mode: bool | int = False, film: np.ndarray | None = None,
               normals: np.ndarray | None = None):
        """Scalar path tracer over a pixel rectangle (default: whole film); returns the RGBA film.  `normals`
        (H, W, 3 float32) receives the NORMALS channel (cpu.cpp:194-196) when given."""
        cam = scene.camera
        d = scene.desc()
        h = self.lib.orc_scene_create(C.byref(d))
        if film is None:
            film = np.zeros((cam.film_height, cam.film_width, 4), np.float32)
        x0, y0, w, hh = region if region is not None else (0, 0, cam.film_width, cam.film_height)
        s0, s1 = spp_range if spp_range is not None else (0, spp)
        self.lib.orc_render(h, nodes.ctypes.data, packets.ctypes.data, x0, y0, w, hh, s0, s1, spp, pps, depth, seed,
                            int(rcp_mode), film.ctypes.data, normals.ctypes.data if normals is not None else None)
        self.lib.orc_scene_destroy(h)
        return film

    def _mat_ptr(self, scene: Scene, mat: int):
        d = scene.desc()
        return d, C.addressof(d.materials[mat])

    def bsdf_f(self, scene: Scene, mat: int, n, wi, wo) -> np.ndarray:
        """bsdf_t::f of material `mat` at shading normal n (restatement)."""
        d, mp = self._mat_ptr(scene, mat)
        a = [np.ascontiguousarray(x, np.float32) for x in (n, wi, wo)]
        out = np.zeros(3, np.float32)
        self.lib.orc_bsdf_f(mp, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, out.ctypes.data)
        return out

    def bsdf_sample(self, scene: Scene, mat: int, n, wi, sx: float, sy: float):
        """bsdf_t::sample (restatement): (alive, wo, f, pdf, flags)."""
        d, mp = self._mat_ptr(scene, mat)
        a = [np.ascontiguousarray(x, np.float32) for x in (n, wi)]
        wo, f = np.zeros(3, np.float32), np.zeros(3, np.float32)
        pdf, fl = C.c_float(0), C.c_uint32(0)
        ok = self.lib.orc_bsdf_sample(mp, a[0].ctypes.data, a[1].ctypes.data, sx, sy, wo.ctypes.data, f.ctypes.data,
                                      C.addressof(pdf), C.addressof(fl))
        return bool(ok), wo, f, pdf.value, fl.value

    def camera_rays_lens(self, scene: Scene, x0, y0, w, h, jx, jy, lens_u, lens_v, rcp_mode: bool = False):
        """camera::perspective_kernel_t for the rectangle with one film jitter and one lens sample per slot
        (thin lens when scene.camera.aperture_radius != 0).  Returns (p, wi) as two (n, 3) float32 arrays."""
        d = scene.desc()
        n = w * h
        out = [np.zeros(n, np.float32) for _ in range(6)]
        lu = np.ascontiguousarray(lens_u, np.float32)
        lv = np.ascontiguousarray(lens_v, np.float32)
        cam_ptr = C.addressof(d) + type(d).camera.offset
        self.lib.orc_camera_rays_lens(cam_ptr, x0, y0, w, h, jx, jy, lu.ctypes.data, lv.ctypes.data, 1 if rcp_mode else 0,
                                      *[a.ctypes.data for a in out])
        return np.stack(out[:3], 1), np.stack(out[3:], 1)

    def film_jitter(self, seed: int, spp: int):
        jx, jy = np.zeros(spp, np.float32), np.zeros(spp, np.float32)
        self.lib.orc_film_jitter(seed, spp, jx.ctypes.data, jy.ctypes.data)
        return jx, jy


class RefScene:
    """A scene inside the compiled reference: reference mesh_t/scene_t + its own mbvh_t."""

    def __init__(self, lib, scene: Scene):
        self.lib = lib
        self._scene = scene
        d = scene.desc()
        self.h = lib.ref_scene_create(C.byref(d))
        self.build_seconds = None

    def build(self):
        self.build_seconds = self.lib.ref_accel_build(self.h)
        return self.build_seconds

    def accel(self):
        """(nodes288 bytes, packets384 bytes) exactly as the reference builder laid them out."""
        nn, np_ = self.lib.ref_accel_num_nodes(self.h), self.lib.ref_accel_num_packets(self.h)
        nodes = np.zeros(nn * NODE_BYTES, np.uint8)
        packets = np.zeros(np_ * PACKET_BYTES, np.uint8)
        self.lib.ref_accel_copy(self.h, nodes.ctypes.data, nn, packets.ctypes.data, np_)
        return nodes, packets

    def trace(self, rays: RayBatch, kind: str = "stream", threads: int = 1):
        out = rays.copy()
        f, u = out.float_ptrs()
        secs = self.lib.ref_trace(self.h, 0 if kind == "stream" else 1, f, u, out.n, threads)
        return out, secs

    def render(self, spp: int, pps: int = 1, depth: int = 9, single_threaded: bool = True, normals: np.ndarray | None = None,
               cuda: bool = False):
        """cpu_t (or, cuda=True, the drop-in cuda_t) start / join; `normals` (H, W, 3) requests the NORMALS channel."""
        cam = self._scene.camera
        img = np.zeros((cam.film_height, cam.film_width, 4), np.float32)
        secs = self.lib.ref_render_aov(self.h, spp, pps, depth, 1 if single_threaded else 0, 1 if cuda else 0, img.ctypes.data,
                                       normals.ctypes.data if normals is not None else None)
        if secs < 0:
            raise RuntimeError("device raised (see stderr)")
        return img, secs

    def render_devices(self, spp: int, pps: int = 1, depth: int = 9, n_cuda: int = 1, with_cpu: bool = False,
                       cpu_single_threaded: bool = False):
        """One frame on several devices sharing frame.tiles (session.cpp:85-99): n_cuda cuda_t (+ the reference's cpu_t).
        Returns (image, seconds, tiles rendered per cuda device)."""
        cam = self._scene.camera
        img = np.zeros((cam.film_height, cam.film_width, 4), np.float32)
        per = (C.c_int * (n_cuda + 1))()
        secs = self.lib.ref_render_devices(self.h, spp, pps, depth, n_cuda, 1 if with_cpu else 0, 1 if cpu_single_threaded else 0,
                                           img.ctypes.data, per)
        if secs < 0:
            raise RuntimeError("a device raised (see stderr)")
        return img, secs, [int(per[i]) for i in range(n_cuda)]

    def cuda_join_raises(self) -> int:
        return int(self.lib.ref_cuda_join_raises(self.h))

    def render_cuda(self, spp: int, pps: int = 1, depth: int = 9):
        """The same frame through the drop-in GPU device: the reference's host code (scene_t, tiles_t,
        sampler, film_t) driving cuda_t::preprocess/start/join (integration/cuda.cpp -> libphos_cuda.so)."""
        cam = self._scene.camera
        img = np.zeros((cam.film_height, cam.film_width, 4), np.float32)
        secs = self.lib.ref_render_on(self.h, spp, pps, depth, 1, 1, img.ctypes.data)
        if secs < 0:
            raise RuntimeError("cuda_t raised (see stderr)")
        return img, secs

    def render_cuda_frames(self, spp: int, pps: int = 1, depth: int = 9, frames: int = 3):
        """`frames` frames on ONE cuda_t made and preprocessed once, started / joined per frame (session.cpp:224-229: one
        render per view on prepared devices).  Returns (last image, [seconds start..join per frame]); frame 0 is cold."""
        cam = self._scene.camera
        img = np.zeros((cam.film_height, cam.film_width, 4), np.float32)
        secs = (C.c_double * frames)()
        if self.lib.ref_render_frames_cuda(self.h, spp, pps, depth, frames, img.ctypes.data, secs) != 0:
            raise RuntimeError("cuda_t raised (see stderr)")
        return img, list(secs)

    def camera_rays(self, x0, y0, w, h, jx, jy, lens_u, lens_v):
        """The reference's own camera::perspective_kernel_t on one tile (w % 8 == 0, w * h <= 1024)."""
        n = w * h
        assert w % 8 == 0 and n <= 1024
        out = [np.zeros(n, np.float32) for _ in range(6)]
        lu = np.ascontiguousarray(lens_u, np.float32)
        lv = np.ascontiguousarray(lens_v, np.float32)
        self.lib.ref_camera_rays(self.h, x0, y0, w, h, jx, jy, lu.ctypes.data, lv.ctypes.data, *[a.ctypes.data for a in out])
        return np.stack(out[:3], 1), np.stack(out[3:], 1)

    def bsdf_f(self, mat: int, n, wi, wo) -> np.ndarray:
        """The reference's own bsdf_t::f for material `mat` (built through material_t::evaluate)."""
        a = [np.ascontiguousarray(x, np.float32) for x in (n, wi, wo)]
        out = np.zeros(3, np.float32)
        self.lib.ref_bsdf_f(self.h, mat, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, out.ctypes.data)
        return out

    def bsdf_sample(self, mat: int, n, wi, sx: float, sy: float):
        a = [np.ascontiguousarray(x, np.float32) for x in (n, wi)]
        wo, f = np.zeros(3, np.float32), np.zeros(3, np.float32)
        pdf, fl = C.c_float(0), C.c_uint32(0)
        ok = self.lib.ref_bsdf_sample(self.h, mat, a[0].ctypes.data, a[1].ctypes.data, sx, sy, wo.ctypes.data, f.ctypes.data,
                                      C.addressof(pdf), C.addressof(fl))
        return bool(ok), wo, f, pdf.value, fl.value

    def num_lights(self):
        return self.lib.ref_scene_num_lights(self.h)

    def __del__(self):
        try:
            self.lib.ref_scene_destroy(self.h)
        except Exception:
            pass


class RefLib:
    @staticmethod
    def available(cuda: bool = False) -> bool:
        return os.path.exists(REF_CUDA_SO if cuda else REF_SO)

    def __init__(self, cuda: bool = False):
        self.cuda = cuda
        self.lib = L = C.CDLL(REF_CUDA_SO if cuda else REF_SO)
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.POINTER(PhosSceneDesc)]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_num_lights.argtypes = [C.c_void_p]
        L.ref_scene_num_lights.restype = C.c_uint32
        L.ref_accel_build.argtypes = [C.c_void_p]
        L.ref_accel_build.restype = C.c_double
        for fn in (L.ref_accel_num_nodes, L.ref_accel_num_packets):
            fn.argtypes = [C.c_void_p]
            fn.restype = C.c_uint32
        L.ref_accel_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.ref_trace.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_uint32)),
                                C.c_uint64, C.c_int]
        L.ref_trace.restype = C.c_double
        L.ref_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.ref_render.restype = C.c_double
        L.ref_render_on.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.ref_render_on.restype = C.c_double
        L.ref_render_aov.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_render_aov.restype = C.c_double
        if hasattr(L, "ref_render_frames_cuda"):
            L.ref_render_frames_cuda.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
            L.ref_render_frames_cuda.restype = C.c_int
        L.ref_bsdf_f.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 4
        L.ref_bsdf_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float] + [C.c_void_p] * 4
        L.ref_bsdf_sample.restype = C.c_int
        L.ref_camera_rays.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [C.c_float, C.c_float] + [C.c_void_p] * 8
        L.ref_cuda_device_count.restype = C.c_int
        if cuda:
            L.ref_discover.argtypes = [C.c_int, C.POINTER(C.c_int)]
            L.ref_discover.restype = C.c_int
            L.ref_render_devices.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.POINTER(C.c_int)]
            L.ref_render_devices.restype = C.c_double
            L.ref_cuda_join_raises.argtypes = [C.c_void_p]
            L.ref_cuda_join_raises.restype = C.c_int
        L.ref_hardware_concurrency.restype = C.c_uint32
        L.ref_sizeof_node.restype = C.c_uint32
        L.ref_sizeof_packet.restype = C.c_uint32
        assert L.ref_sizeof_node() == NODE_BYTES and L.ref_sizeof_packet() == PACKET_BYTES

    def scene(self, scene: Scene) -> RefScene:
        return RefScene(self.lib, scene)

    def discover(self, host_only: bool = False):
        """(devices, of which GPUs) xpu_t::discover returns (src/xpu.cpp:7-9 + integration/xpu_discover.patch)."""
        n = C.c_int(0)
        total = self.lib.ref_discover(1 if host_only else 0, C.byref(n))
        return int(total), int(n.value)

    def hardware_concurrency(self) -> int:
        return int(self.lib.ref_hardware_concurrency())
