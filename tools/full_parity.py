#!/usr/bin/env python3
"""GPU box: EVERY ray of the BASELINE config-3 streams (10 M-triangle terrain, 1920 x 1080) against the
oracle traversal; prints the mismatch count and classifies each mismatch (north star: ids exact except
provably tied or edge-grazing rays, reported as a count, <= 1e-5 of the rays)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Oracle
from parity import mismatches
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles
from phosphorus_mk2_b200.rays import HIT

which = sys.argv[1] if len(sys.argv) > 1 else "terrain"
lib_path = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else None  # a tuning variant of the device library
sc = scenes.terrain() if which == "terrain" else scenes.sphere_field()
acc = Accel(sc); nodes, packets = acc.nodes_array(), acc.packets_array()
dev = CudaDevice(Options(), 0, lib_path=lib_path); dev.preprocess(sc, acc); dev.upload_scene(sc)
orc = Oracle(); cam = sc.camera
tiles = make_tiles(cam.film_width, cam.film_height); n = cam.film_width * cam.film_height
for stream in ("primary", "bounce", "shadow"):
    dr = dev.device_rays(n)
    if stream == "primary":
        dev.camera_rays(tiles, dr); k = n
    else:
        k = dev.wavefront_rays(tiles, dr, stream, 0, 1, 42)
    rays = dr.download().slice(0, k)
    dev.trace_device_n(dr, k); got = dr.download().slice(0, k); dr.free()
    t0 = time.time(); want, cnt = orc.traverse(nodes, packets, rays); dt = time.time() - t0
    bad = mismatches(rays, got, want)
    print(f"{which} {os.path.basename(lib_path or 'libphos_cuda.so')} {stream}: {k} rays, oracle {dt:.1f} s, mismatches {len(bad)} ({len(bad)/k:.2e})", flush=True)
    for i in bad[:10]:
        gh, wh = bool(got.flags[i] & HIT), bool(want.flags[i] & HIT)
        rel = abs(float(got.d[i]) - float(want.d[i])) / max(abs(float(want.d[i])), 1e-30) if gh and wh else float("nan")
        print(f"   ray {i}: gpu hit={gh} d={got.d[i]:.9g} face={got.face[i]} | oracle hit={wh} d={want.d[i]:.9g} face={want.face[i]} | rel dt {rel:.2e} "
              f"u,v oracle=({want.u[i]:.3e},{want.v[i]:.3e})")
