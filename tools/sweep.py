#!/usr/bin/env python3
"""Tuning sweep (GPU box): time the traversal kernel of several library variants on the bench
workloads.  Usage: python tools/sweep.py [--workloads spheres,terrain,terrain_inc] lib1.so lib2.so ...
Variants are built with `make -C phosphorus_mk2_b200/csrc VARIANT=name DEFS="-D..."`."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phosphorus_mk2_b200 import raysets, scenes  # noqa: E402
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="spheres,terrain_inc")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--terrain-n", type=int, default=2237)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    for wl in a.workloads.split(","):
        t0 = time.time()
        if wl == "spheres":
            sc = scenes.sphere_field()
        else:
            sc = scenes.terrain(n=a.terrain_n)
        acc = Accel(sc)
        host = None
        if wl.endswith("_inc"):
            host = raysets.aimed_rays(sc, 1920 * 1080, seed=5)
        elif wl.endswith("_shadow"):
            host = raysets.as_shadow(raysets.aimed_rays(sc, 1920 * 1080, seed=5), seed=6, masked_fraction=0.0)
        print(f"# {wl}: {sc.num_triangles()} tris, build {acc.build_seconds:.1f}s, setup {time.time()-t0:.1f}s", flush=True)
        ref = None
        for lib in a.libs:
            dev = CudaDevice(Options(), 0, lib_path=os.path.abspath(lib))
            dev.preprocess(sc, acc)
            dev.upload_scene(sc)
            cam = sc.camera
            n = cam.film_width * cam.film_height
            dr = dev.device_rays(n)
            tiles = make_tiles(cam.film_width, cam.film_height)

            def fresh():
                if host is None:
                    dev.camera_rays(tiles, dr)
                else:
                    dr.upload(host)
            ms = []
            for i in range(3 + a.steps):
                dev.flush_l2()
                fresh()
                dev.timer_begin()
                dev.trace_device(dr)
                t = dev.timer_end()
                if i >= 3:
                    ms.append(t)
            fresh()
            nodes, tris = dev.trace_count(dr)
            out = dr.download()
            sig = (int(out.flags.sum()), float(out.d[out.hit].astype(np.float64).sum()), int(out.face[out.hit].astype(np.uint64).sum()))
            if ref is None:
                ref = sig
            print(f"{wl:12s} {os.path.basename(lib):34s} {n/np.mean(ms)/1e3:9.1f} Mrays/s  min {n/np.min(ms)/1e3:9.1f}  "
                  f"nodes/ray {nodes/n:.2f} tris/ray {tris/n:.2f} hit {out.hit.mean():.3f} same={sig == ref}", flush=True)
            dr.free()
            dev.close()


if __name__ == "__main__":
    main()
