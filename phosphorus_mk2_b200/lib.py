"""ctypes binding of libphos_cuda.so (include/phos_cuda.h).

The library is the product: there is no Python, numpy or CPU implementation of the ray-query path
behind it.  If the shared object is missing, or no CUDA device is visible when a context is
created, this module raises — it never falls back.
"""
from __future__ import annotations

import ctypes as C
import os

from .rays import PhosRays
from .scene import PhosSceneDesc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libphos_cuda.so")

PHOS_OK, PHOS_ERR_INVALID, PHOS_ERR_CUDA, PHOS_ERR_NO_DEVICE, PHOS_ERR_ACCEL = 0, 1, 2, 3, 4


class PhosError(RuntimeError):
    pass


class PhosOptions(C.Structure):
    _fields_ = [("samples_per_pixel", C.c_uint32), ("paths_per_sample", C.c_uint32), ("path_depth", C.c_uint32)]


class PhosTile(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("w", C.c_uint32), ("h", C.c_uint32)]


class PhosAccelStats(C.Structure):
    _fields_ = [
        ("ref_nodes", C.c_uint32), ("ref_packets", C.c_uint32),
        ("nodes", C.c_uint32), ("triangles", C.c_uint32),
        ("max_depth", C.c_uint32), ("max_leaf_triangles", C.c_uint32),
        ("bytes_nodes", C.c_uint64), ("bytes_triangles", C.c_uint64),
        ("repack_seconds", C.c_double), ("upload_seconds", C.c_double),
    ]


# every symbol include/phos_cuda.h declares: name -> (restype, argtypes)
_VP, _U32, _U64, _I = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
_RP = C.POINTER(PhosRays)
SYMBOLS = {
    "phos_cuda_device_count": (_I, []),
    "phos_cuda_create": (_VP, [_I, C.POINTER(PhosOptions)]),
    "phos_cuda_destroy": (None, [_VP]),
    "phos_cuda_last_error": (C.c_char_p, [_VP]),
    "phos_bvh_build": (_VP, [C.POINTER(PhosSceneDesc), _I]),
    "phos_bvh_num_nodes": (_U32, [_VP]),
    "phos_bvh_num_packets": (_U32, [_VP]),
    "phos_bvh_nodes": (_VP, [_VP]),
    "phos_bvh_packets": (_VP, [_VP]),
    "phos_bvh_build_seconds": (C.c_double, [_VP]),
    "phos_bvh_free": (None, [_VP]),
    "phos_cuda_upload_accel": (_I, [_VP, _VP, _U32, _VP, _U32]),
    "phos_cuda_accel_stats": (_I, [_VP, C.POINTER(PhosAccelStats)]),
    "phos_cuda_trace": (_I, [_VP, _RP, _U64]),
    "phos_cuda_trace_device": (_I, [_VP, _RP, _U64]),
    "phos_cuda_trace_count": (_I, [_VP, _RP, _U64, C.POINTER(_U64), C.POINTER(_U64)]),
    "phos_cuda_trace_profile": (_I, [_VP, _RP, _U64, C.POINTER(_U64)]),
    "phos_cuda_rays_alloc": (_I, [_VP, _U64, _RP]),
    "phos_cuda_rays_free": (_I, [_VP, _RP]),
    "phos_cuda_rays_upload": (_I, [_VP, _RP, _RP, _U64]),
    "phos_cuda_rays_download": (_I, [_VP, _RP, _RP, _U64]),
    "phos_cuda_synchronize": (_I, [_VP]),
    "phos_cuda_timer_begin": (_I, [_VP]),
    "phos_cuda_timer_end": (_I, [_VP, C.POINTER(C.c_float)]),
    "phos_cuda_launch_count": (_U64, [_VP]),
    "phos_cuda_host_alloc": (_VP, [_U64]),
    "phos_cuda_host_free": (None, [_VP]),
    "phos_cuda_flush_l2": (_I, [_VP]),
    "phos_cuda_upload_scene": (_I, [_VP, C.POINTER(PhosSceneDesc)]),
    "phos_cuda_bind_host_to_device": (_I, [C.c_int]),
    "phos_cuda_build_accel": (_I, [_VP, _VP]),
    "phos_cuda_camera_rays": (_I, [_VP, C.POINTER(PhosTile), _U32, C.c_float, C.c_float, _RP]),
    "phos_cuda_camera_rays_lens": (_I, [_VP, C.POINTER(PhosTile), _U32, C.c_float, C.c_float, C.c_uint64, _U32, _RP]),
    "phos_cuda_render": (_I, [_VP, C.POINTER(PhosTile), _U32, _U32, _U32, _U32, _U64]),
    "phos_cuda_wavefront_rays": (_I, [_VP, C.POINTER(PhosTile), _U32, _U32, _U32, _U64, _I, _RP, _U64, C.POINTER(_U64)]),
    "phos_cuda_film_clear": (_I, [_VP]),
    "phos_cuda_film_device_ptr": (_I, [_VP, C.POINTER(_VP), C.POINTER(_U64)]),
    "phos_cuda_film_read": (_I, [_VP, _VP, _U32, _U32, _U32, _U32]),
    "phos_cuda_enable_normals": (_I, [_VP, C.c_int]),
    "phos_cuda_film_read_normals": (_I, [_VP, _VP, _U32, _U32, _U32, _U32]),
    "phos_cuda_reference_normalize": (_I, [_VP, _I]),
    "phos_cuda_comm_unique_id": (_I, [_VP]),
    "phos_cuda_comm_init": (_I, [_VP, _I, _I, _VP]),
    "phos_cuda_comm_adopt": (_I, [_VP, _VP, _I]),
    "phos_cuda_comm_destroy": (None, [_VP]),
    "phos_cuda_film_reduce": (_I, [_VP, _I]),
}

_libs: dict = {}


def load(path: str | None = None) -> C.CDLL:
    """Load libphos_cuda.so (or a tuning variant given by `path` / $PHOS_CUDA_LIB) and bind every
    declared symbol; raises PhosError if it is not built."""
    path = path or os.environ.get("PHOS_CUDA_LIB") or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise PhosError(f"{path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C phosphorus_mk2_b200/csrc`); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib
