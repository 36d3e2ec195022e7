"""CPU experiment (test / tuning infrastructure, not product): warp-level step counts of trace_kernel's scheduling under
candidate policies, on the host emulation of the device traversal (tools/warp_sim/sim.cpp).
python tools/warp_sim/run.py spheres|terrain"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from phosphorus_mk2_b200 import raysets, scenes
from phosphorus_mk2_b200.device import Accel
from phosphorus_mk2_b200.rays import RayBatch, PhosRays
from oracle.pyoracle import Oracle

so = os.path.join(HERE, "libsim.so")
subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include",
                os.path.join(HERE, "sim.cpp"), os.path.join(ROOT, "phosphorus_mk2_b200", "csrc", "repack.cpp"), "-o", so], check=True)
L = C.CDLL(so)


class Params(C.Structure):
    _fields_ = [("refill_min", C.c_int), ("tri_bias", C.c_int), ("defer", C.c_int), ("defer_shadow_only", C.c_int),
                ("tri_min_lanes", C.c_int), ("hint", C.c_int)]


L.warp_sim.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(PhosRays), C.c_uint64, C.c_int, C.POINTER(Params),
                       C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
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
    # tile order like the renderer's streams: 32 x 32 tiles, row-major inside
    idx = np.arange(w * h).reshape(h, w)
    order = np.concatenate([idx[y:y + 32, x:x + 32].ravel() for y in range(0, h, 32) for x in range(0, w, 32)])
    d = d[order]
    o = np.broadcast_to(eye, d.shape)
    return RayBatch.from_arrays(o.astype(np.float32), d.astype(np.float32))


sets = {"camera": camera_rays(sc)}
orc = Oracle(); hit, _ = orc.traverse(nodes, packets, sets["camera"])
m = (hit.flags & 1) != 0
cam = sets["camera"]
Pp = np.stack([cam.px, cam.py, cam.pz], 1)[m].astype(np.float64) + np.stack([cam.wx, cam.wy, cam.wz], 1)[m] * hit.d[m][:, None]
rng = np.random.default_rng(3); u1, u2 = rng.random(len(Pp)), rng.random(len(Pp))
rr = np.sqrt(u1); th = 2 * np.pi * u2
D = np.stack([rr * np.cos(th), np.sqrt(1 - u1), rr * np.sin(th)], 1)
sets["bounce"] = RayBatch.from_arrays((Pp + np.array([0, 1e-3, 0])).astype(np.float32), D.astype(np.float32))
sh = sets["bounce"].copy(); sh.flags[:] = 4
sets["shadow"] = sh

NODE, TRI, RAY = 291.0, 216.0, 9.2
policies = [("shipped: bias 3, refill 6", dict(refill_min=6, tri_bias=3, defer=0)),
            ("bias 2", dict(refill_min=6, tri_bias=2, defer=0)),
            ("defer 1", dict(refill_min=6, tri_bias=3, defer=1)),
            ("defer 2", dict(refill_min=6, tri_bias=3, defer=2)),
            ("defer 2, bias 2", dict(refill_min=6, tri_bias=2, defer=2)),
            ("defer 2, bias 1", dict(refill_min=6, tri_bias=1, defer=2)),
            ("defer 4, bias 1", dict(refill_min=6, tri_bias=1, defer=4)),
            ("defer 4, tri step from 16 lanes", dict(refill_min=6, tri_bias=1, defer=4, tri_min_lanes=16)),
            ("defer 2, shadow rays only", dict(refill_min=6, tri_bias=3, defer=2, defer_shadow_only=1)),
            ("defer 4, bias 1, shadow rays only", dict(refill_min=6, tri_bias=1, defer=4, defer_shadow_only=1))]
if os.environ.get("GRID"):
    policies = [("shipped: bias 3, refill 6", dict(refill_min=6, tri_bias=3, defer=0))]
    for q in (2, 3):
        for b in (2, 3, 4):
            for rf in (4, 6, 8):
                policies.append((f"defer {q}, bias {b}, refill {rf}", dict(refill_min=rf, tri_bias=b, defer=q)))
for sname, rays in sets.items():
    base = None; ref = None
    for pname, kw in policies:
        p = Params(**kw)
        rb = rays.copy(); s = rb.as_struct(); out = (C.c_uint64 * 10)()
        od = np.zeros(rb.n, np.float32); of = np.zeros(rb.n, np.uint32)
        assert L.warp_sim(nodes.ctypes.data, len(nodes) // 288, packets.ctypes.data, len(packets) // 384, C.byref(s), rb.n, 64, C.byref(p), out,
                          od.ctypes.data, of.ctypes.data) == 0
        wn, wt, ln, lt, nn, nt, it, rf, tr, sp = list(out)
        cost = (NODE * wn + TRI * wt) * 1.0 / tr + RAY  # warp instructions per ray (x 32 lanes of issue width)
        if base is None:
            base, ref = cost, (od.copy(), of.copy())
        closest = (rays.flags & 4) == 0
        same = np.array_equal(of, ref[1]) and np.array_equal(od[closest], ref[0][closest])
        print(f"{which:8s} {sname:7s} {pname:36s} node steps/ray {wn/tr:6.3f} @ {ln/max(wn,1):4.1f} lanes  tri steps/ray {wt/tr:6.3f} @ {lt/max(wt,1):4.1f} lanes  "
              f"nodes/ray {nn/tr:6.2f} tris/ray {nt/tr:5.2f}  speculative {100*sp/max(nn,1):4.1f} %  instr/ray {cost:7.1f} ({100*cost/base:5.1f} %)  same={same}", flush=True)
