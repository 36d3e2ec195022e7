#!/usr/bin/env python3
"""Generate the golden ray-query vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference): `make -C oracle ref` compiles the reference's
own accel/bvh.cpp, kernels/cpu/{stream,linear}_bvh_kernel.cpp ... into oracle/_ref/libphos_ref.so;
this script drives it through ctypes and records, per case:
  nodes, packets      the reference builder's own arrays (288 B / 384 B records)
  rays_*              the input ray stream (p, wi, d, flags, and the pre-set mesh/face/u/v)
  linear_*            outputs of the reference's brute-force kernel (linear_mbvh_kernel_t) — exact MT
  stream_*            outputs of the reference's stream kernel (stream_mbvh_kernel_t) — uses rcpps
The reference ships no tests or vectors of its own, so these are the pin.
Files are small .npz archives, committed; the GPU box has no /root/reference and only reads them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import RefLib  # noqa: E402
from phosphorus_mk2_b200 import raysets, scenes  # noqa: E402
from phosphorus_mk2_b200.rays import MASKED, SHADOW, RayBatch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags")
OUTS = ("d", "u", "v", "mesh", "face", "flags")


def special_rays(scene) -> RayBatch:
    """Edge cases: axis-parallel directions (zero components), rays starting exactly on geometry,
    rays along triangle edges / through vertices, zero-length tmax, negative zero components."""
    lo, hi = raysets.scene_bounds(scene)
    c = 0.5 * (lo + hi)
    v = np.concatenate([m.vertices for m in scene.meshes]).astype(np.float64)
    rng = np.random.default_rng(99)
    P, W, D = [], [], []
    for axis in range(3):
        for sgn in (1.0, -1.0):
            for k in range(24):
                tgt = v[rng.integers(0, len(v))] + (rng.random(3) - 0.5) * 0.05 * (hi - lo) * (k % 2)
                o = tgt.copy()
                o[axis] = (lo[axis] - 1.0) if sgn > 0 else (hi[axis] + 1.0)
                w = np.zeros(3)
                w[axis] = sgn
                if k % 3 == 0:
                    w[(axis + 1) % 3] = -0.0
                P.append(o), W.append(w), D.append(3.4028234663852886e38)
    for k in range(96):  # origin exactly on a vertex, random direction
        P.append(v[rng.integers(0, len(v))])
        w = rng.normal(size=3)
        W.append(w / np.linalg.norm(w)), D.append(3.4028234663852886e38)
    for k in range(96):  # aimed exactly at a vertex from the centre region
        o = c + (rng.random(3) - 0.5) * 0.3 * (hi - lo) + np.array([0, 0.6 * (hi - lo)[1], 0])
        w = v[rng.integers(0, len(v))] - o
        P.append(o), W.append(w / np.linalg.norm(w)), D.append(3.4028234663852886e38)
    for k in range(32):  # tmax = 0 and tiny tmax
        o = c + np.array([0, 0.6 * (hi - lo)[1], 0])
        w = v[rng.integers(0, len(v))] - o
        P.append(o), W.append(w / np.linalg.norm(w)), D.append(0.0 if k % 2 else 1e-3)
    return RayBatch.from_arrays(np.array(P, np.float32), np.array(W, np.float32), d=np.array(D, np.float32))


def record(ref, name, scene, batches):
    rs = ref.scene(scene)
    rs.build()
    nodes, packets = rs.accel()
    out = {"nodes": nodes, "packets": packets}
    for bname, rays in batches.items():
        lin, _ = rs.trace(rays, "linear")
        st, _ = rs.trace(rays, "stream")
        for f in FIELDS:
            out[f"{bname}_rays_{f}"] = getattr(rays, f)
        for f in OUTS:
            out[f"{bname}_linear_{f}"] = getattr(lin, f)
            out[f"{bname}_stream_{f}"] = getattr(st, f)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "tris", scene.num_triangles(), "nodes", len(nodes) // 288, "packets", len(packets) // 384,
          "->", os.path.getsize(path) // 1024, "KiB")


def main():
    assert RefLib.available(), "build oracle/_ref first: make -C oracle ref"
    ref = RefLib()
    n = 2048
    for name, scene in (("cornell", scenes.cornell_box(64, 64)), ("heightfield24", scenes.heightfield(24)),
                        ("spheres2", scenes.sphere_field(2, 12, 6, 64, 64))):
        aimed = raysets.aimed_rays(scene, n, seed=11)
        rnd = raysets.random_rays(scene, n, seed=12)
        shadow = raysets.as_shadow(raysets.aimed_rays(scene, n, seed=13), seed=14)
        mixed = raysets.aimed_rays(scene, n, seed=15)
        sh = raysets.as_shadow(mixed, seed=16)
        pick = np.arange(n) % 3 == 0  # every third ray is a shadow query, some of them masked
        for f in ("d", "mesh", "face", "u", "v", "flags"):
            getattr(mixed, f)[pick] = getattr(sh, f)[pick]
        record(ref, name, scene, {"aimed": aimed, "random": rnd, "shadow": shadow, "mixed": mixed,
                                  "special": special_rays(scene)})


if __name__ == "__main__":
    main()
