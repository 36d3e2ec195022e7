"""CPU experiment (test infrastructure, not product): node visits per ray of the device traversal (host emulation of
csrc/trace_ray.cuh) with candidate per-group distance bounds.  python tools/visit_stats/run.py spheres|terrain"""
import ctypes as C, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from phosphorus_mk2_b200 import raysets, scenes
from phosphorus_mk2_b200.device import Accel
from phosphorus_mk2_b200.rays import RayBatch, PhosRays
from oracle.pyoracle import Oracle
import subprocess
so = os.path.join(HERE, "libstats.so")
subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include",
                os.path.join(HERE, "stats.cpp"), os.path.join(ROOT, "phosphorus_mk2_b200", "csrc", "repack.cpp"), "-o", so], check=True)
L = C.CDLL(so)
L.visit_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(PhosRays), C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
which = sys.argv[1]
sc = scenes.sphere_field() if which == "spheres" else scenes.terrain(n=700)
a = Accel(sc); nodes, packets = a.nodes_array(), a.packets_array()
def camera_rays(sc, w=480, h=270):
    cam = sc.camera; m = cam.to_world.astype(np.float64)
    eye = m[3, :3]; r, u, f = m[0, :3], m[1, :3], -m[2, :3]
    zoom = np.tan(cam.fov / 2); asp = h / w
    xs = (np.arange(w) + 0.5) / w * 2 - 1; ys = (np.arange(h) + 0.5) / h * 2 - 1
    X, Y = np.meshgrid(xs, ys)
    d = f[None, None, :] + X[..., None] * zoom * r[None, None, :] - Y[..., None] * zoom * asp * u[None, None, :]
    d = d.reshape(-1, 3); d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(eye, d.shape)
    return RayBatch.from_arrays(o.astype(np.float32), d.astype(np.float32))
sets = {"camera": camera_rays(sc), "aimed": raysets.aimed_rays(sc, 100000, seed=5)}
# bounce-like: from the camera hits, cosine-weighted about +y
orc = Oracle(); hit, _ = orc.traverse(nodes, packets, sets["camera"])
m = (hit.flags & 1) != 0
cam = sets["camera"]
P = np.stack([cam.px, cam.py, cam.pz], 1)[m].astype(np.float64) + np.stack([cam.wx, cam.wy, cam.wz], 1)[m] * hit.d[m][:, None]
rng = np.random.default_rng(3); u1, u2 = rng.random(len(P)), rng.random(len(P))
rr = np.sqrt(u1); th = 2 * np.pi * u2
D = np.stack([rr * np.cos(th), np.sqrt(1 - u1), rr * np.sin(th)], 1)
sets["bounce"] = RayBatch.from_arrays((P + np.array([0, 1e-3, 0])).astype(np.float32), D.astype(np.float32))
names = {0: "baseline", 1: "exact group bound", 2: "back-half bound", 4: "per-child distances"}
for sname, rays in sets.items():
    base = None
    for mode in (0, 1, 2, 4):
        rb = rays.copy(); s = rb.as_struct(); out = (C.c_uint64 * 4)()
        per = np.zeros(rb.n, np.uint32) if mode == 0 else None
        assert L.visit_stats(nodes.ctypes.data, len(nodes) // 288, packets.ctypes.data, len(packets) // 384, C.byref(s), rb.n, mode, out,
                             per.ctypes.data if per is not None else None) == 0
        v, nh, sk, h = list(out)
        if base is None: base = v
        if per is not None:
            q = np.percentile(per, [50, 90, 99, 99.9, 100])
            print(f"{which:8s} {sname:7s} node visits per ray: mean {per.mean():.1f}, median {q[0]:.0f}, p90 {q[1]:.0f}, p99 {q[2]:.0f}, p99.9 {q[3]:.0f}, max {q[4]:.0f}", flush=True)
        print(f"{which:8s} {sname:7s} {names[mode]:20s} visits/ray {v/rb.n:6.2f} ({100*v/base:5.1f} %)  no-hit visits {100*nh/v:4.1f} %  skipped {sk/rb.n:5.2f}/ray  hits {h}", flush=True)
