"""Ray stream container: the fields of the reference's ``ray_t<N>`` (src/state.hpp:39-57) as flat
numpy SoA arrays of arbitrary length, plus the flag bits of src/state.hpp:33-36."""
from __future__ import annotations

import ctypes as C

import numpy as np

HIT, MASKED, SHADOW, SPECULAR = 1, 2, 4, 8
FLT_MAX = np.float32(3.4028234663852886e38)

_F = ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v")
_U = ("mesh", "face", "flags")


class PhosRays(C.Structure):
    """phos_rays (include/phos_cuda.h): 12 SoA pointers, host or device."""

    _fields_ = [
        ("px", C.c_void_p), ("py", C.c_void_p), ("pz", C.c_void_p),
        ("wx", C.c_void_p), ("wy", C.c_void_p), ("wz", C.c_void_p),
        ("d", C.c_void_p), ("mesh", C.c_void_p), ("face", C.c_void_p),
        ("u", C.c_void_p), ("v", C.c_void_p), ("flags", C.c_void_p),
    ]


class RayBatch:
    """n rays, SoA.  ``d`` is tmax on input and hit distance on output (default FLT_MAX)."""

    def __init__(self, n: int):
        self.n = int(n)
        for k in _F:
            setattr(self, k, np.zeros(self.n, np.float32))
        for k in _U:
            setattr(self, k, np.zeros(self.n, np.uint32))
        self.d[:] = FLT_MAX

    @classmethod
    def from_arrays(cls, p, w, d=None, flags=None) -> "RayBatch":
        p = np.asarray(p, np.float32).reshape(-1, 3)
        w = np.asarray(w, np.float32).reshape(-1, 3)
        r = cls(len(p))
        r.px[:], r.py[:], r.pz[:] = p[:, 0], p[:, 1], p[:, 2]
        r.wx[:], r.wy[:], r.wz[:] = w[:, 0], w[:, 1], w[:, 2]
        if d is not None:
            r.d[:] = d
        if flags is not None:
            r.flags[:] = flags
        return r

    def copy(self) -> "RayBatch":
        r = RayBatch(self.n)
        for k in _F + _U:
            getattr(r, k)[:] = getattr(self, k)
        return r

    def slice(self, lo: int, hi: int) -> "RayBatch":
        r = RayBatch(hi - lo)
        for k in _F + _U:
            getattr(r, k)[:] = getattr(self, k)[lo:hi]
        return r

    def as_struct(self) -> PhosRays:
        s = PhosRays()
        for k in _F + _U:
            a = getattr(self, k)
            assert a.flags["C_CONTIGUOUS"] and len(a) == self.n
            setattr(s, k, a.ctypes.data)
        return s

    @property
    def hit(self) -> np.ndarray:
        return (self.flags & HIT) != 0

    def float_ptrs(self):
        """(float*[9], uint32*[3]) pointer tables in the order px,py,pz,wx,wy,wz,d,u,v / mesh,face,flags."""
        f = (C.POINTER(C.c_float) * 9)(*[getattr(self, k).ctypes.data_as(C.POINTER(C.c_float)) for k in _F])
        u = (C.POINTER(C.c_uint32) * 3)(*[getattr(self, k).ctypes.data_as(C.POINTER(C.c_uint32)) for k in _U])
        return f, u
