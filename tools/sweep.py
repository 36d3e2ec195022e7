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
    ap.add_argument("--film", default="1920x1080", help="film size WxH (rays per launch of the camera workloads)")
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    fw, fh = (int(v) for v in a.film.split("x"))
    for wl in a.workloads.split(","):
        t0 = time.time()
        if wl == "spheres":
            sc = scenes.sphere_field(width=fw, height=fh)
        else:
            sc = scenes.terrain(n=a.terrain_n, width=fw, height=fh)
        acc = Accel(sc)
        host = None
        wave = None
        if wl.endswith("_bounce"):
            wave = "bounce"  # BASELINE config 3, ray set A: BSDF-sampled bounce rays from the primary hits
        elif wl.endswith("_nee"):
            wave = "shadow"  # ray set B: next-event shadow rays
        elif wl.endswith("_inc"):
            host = raysets.aimed_rays(sc, 1920 * 1080, seed=5)
        elif wl.endswith("_shadow"):
            host = raysets.as_shadow(raysets.aimed_rays(sc, 1920 * 1080, seed=5), seed=6, masked_fraction=0.0)
        print(f"# {wl}: {sc.num_triangles()} tris, build {acc.build_seconds:.1f}s, setup {time.time()-t0:.1f}s", flush=True)
        ref = None
        for spec in a.libs:  # "lib.so" or "lib.so:ENV=val,ENV2=val" (re-pack knobs are read from the environment)
            lib, _, envs = spec.partition(":")
            saved = {}
            for kv in filter(None, envs.split(",")):
                k, _, v = kv.partition("=")
                saved[k] = os.environ.get(k)
                os.environ[k] = v
            dev = CudaDevice(Options(), 0, lib_path=os.path.abspath(lib))
            dev.preprocess(sc, acc)
            dev.upload_scene(sc)
            cam = sc.camera
            n = cam.film_width * cam.film_height
            dr = dev.device_rays(n)
            tiles = make_tiles(cam.film_width, cam.film_height)

            nrays = n
            if wave:
                pristine = dev.device_rays(n)
                nrays = dev.wavefront_rays(tiles, pristine, wave, 0, 1, 42)
                host_w = pristine.download()
                pristine.free()

            def fresh():
                if wave:
                    dr.upload(host_w)
                elif host is None:
                    dev.camera_rays(tiles, dr)
                else:
                    dr.upload(host)
            ms = []
            for i in range(3 + a.steps):
                dev.flush_l2()
                fresh()
                dev.timer_begin()
                if wave:
                    dev.trace_device_n(dr, nrays)
                else:
                    dev.trace_device(dr)
                t = dev.timer_end()
                if i >= 3:
                    ms.append(t)
            fresh()
            if wave:
                sub = dev.device_rays(max(nrays, 1))
                h2 = host_w.slice(0, max(nrays, 1))
                sub.upload(h2)
                nodes, tris = dev.trace_count(sub)
                out = sub.download()
                sub.free()
                traced = int(((h2.flags & 2) == 0).sum())
            else:
                nodes, tris = dev.trace_count(dr)
                out = dr.download()
                traced = n
            # order-independent signature (the pipeline's compaction order varies from run to run): integer sums of bit patterns
            sig = (int(out.flags.sum()), int(out.d[out.hit].view(np.uint32).astype(np.uint64).sum()), int(out.face[out.hit].astype(np.uint64).sum()))
            if ref is None:
                ref = sig
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            print(f"{wl:14s} {os.path.basename(spec):30s} {traced/np.mean(ms)/1e3:9.1f} Mrays/s  min {traced/np.min(ms)/1e3:9.1f}  "
                  f"depth {dev.accel_stats().max_depth} rays {traced} nodes/ray {nodes/max(traced,1):.2f} tris/ray {tris/max(traced,1):.2f} hit {out.hit.mean():.3f} same={sig == ref}", flush=True)
            dr.free()
            dev.close()


if __name__ == "__main__":
    main()
