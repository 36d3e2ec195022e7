#!/usr/bin/env python3
"""GPU box: wall clock of phos_cuda_trace on page-locked host arrays over pipeline chunk sizes / write-back CTAs / up-copy streams."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles, pinned_ray_batch
F = ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags")
KEYS = ("PHOS_PIPE_CHUNK", "PHOS_E2E_WB_CTAS", "PHOS_E2E_IN_STREAMS")
sc = scenes.sphere_field(); acc = Accel(sc)
dev = CudaDevice.make(Options(), 0); dev.preprocess(sc, acc); dev.upload_scene(sc)
cam = sc.camera; n = cam.film_width * cam.film_height
dr = dev.device_rays(n); dev.camera_rays(make_tiles(cam.film_width, cam.film_height), dr); pristine = dr.download()
h = pinned_ray_batch(n)


def run(label, **env):
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(12):
        for f in F: getattr(h, f)[:] = getattr(pristine, f)
        t0 = time.perf_counter(); dev.trace(h); ts.append(time.perf_counter() - t0)
    t = np.mean(ts[2:])
    print(f"{label:60s} {t*1e3:6.2f} ms (min {min(ts)*1e3:5.2f}) {n/t/1e6:7.0f} Mrays/s", flush=True)


chunks = sys.argv[1].split(",") if len(sys.argv) > 1 else ("65536", "98304", "131072", "163840", "196608", "262144")
for rep in range(2):
    for ch in chunks:
        run(f"chunk {ch}", PHOS_PIPE_CHUNK=ch)
    for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ()):
        run(f"chunk {chunks[0]}, {c} write-back CTAs", PHOS_PIPE_CHUNK=chunks[0], PHOS_E2E_WB_CTAS=c)
