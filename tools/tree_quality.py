#!/usr/bin/env python3
"""CPU: node / triangle tests per ray of the PACKED tree (re-pack knobs from the environment), counted by the
host emulation of the device traversal (tests/emul) on coherent camera rays and incoherent aimed rays.
A quick guide for re-pack heuristics before spending GPU time; usage: tools/tree_quality.py [spheres|terrain] [n]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.pyoracle import Oracle
from phosphorus_mk2_b200 import raysets, scenes
from phosphorus_mk2_b200.device import Accel
from phosphorus_mk2_b200.rays import PhosRays

L = C.CDLL(os.path.join(ROOT, "tests", "emul", "libemul_trace.so"))
L.emul_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(PhosRays), C.c_uint64,
                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_char_p]
which = sys.argv[1] if len(sys.argv) > 1 else "spheres"
sc = scenes.sphere_field() if which == "spheres" else scenes.terrain(n=int(sys.argv[2]) if len(sys.argv) > 2 else 1000)
acc = Accel(sc)
cam = sc.camera
prim = Oracle().camera_rays(cam)
sel = np.arange(0, prim.n, 13)
coh = prim.slice(0, prim.n).take(sel) if hasattr(prim, "take") else None
if coh is None:
    from phosphorus_mk2_b200.rays import RayBatch
    coh = RayBatch(len(sel))
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
        getattr(coh, f)[:] = getattr(prim, f)[sel]
inc = raysets.aimed_rays(sc, 100000, seed=5)
for name, rays in (("camera", coh), ("aimed", inc)):
    r = rays.copy()
    s = r.as_struct()
    nn, nt = C.c_uint64(), C.c_uint64()
    stats = (C.c_uint32 * 4)()
    err = C.create_string_buffer(256)
    t0 = time.time()
    rc = L.emul_trace(acc.root, acc.num_nodes, acc.triangles, acc.num_packets, C.byref(s), r.n, C.byref(nn), C.byref(nt), stats, err)
    assert rc == 0, err.value
    print(f"{which} {name:7s} nodes/ray {nn.value / r.n:6.2f} tris/ray {nt.value / r.n:6.2f}  packed nodes {stats[0]} depth {stats[2]} "
          f"hit {np.mean((r.flags & 1) != 0):.3f}  ({time.time() - t0:.1f} s)", flush=True)
