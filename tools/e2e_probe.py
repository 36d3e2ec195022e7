#!/usr/bin/env python3
"""GPU box: where the time of phos_cuda_trace on page-locked host arrays goes (stage-skipping probes)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles, pinned_ray_batch
sc = scenes.sphere_field(); acc = Accel(sc)
dev = CudaDevice.make(Options(), 0); dev.preprocess(sc, acc); dev.upload_scene(sc)
cam = sc.camera; n = cam.film_width * cam.film_height
tiles = make_tiles(cam.film_width, cam.film_height)
dr = dev.device_rays(n); dev.camera_rays(tiles, dr); pristine = dr.download()
h = pinned_ray_batch(n)
F = ("px","py","pz","wx","wy","wz","d","u","v","mesh","face","flags")
KEYS = ("PHOS_E2E_DEBUG", "PHOS_E2E_SPARSE", "PHOS_PIPE_CHUNK", "PHOS_E2E_WB", "PHOS_E2E_WB_CTAS")
def run(label, **env):
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(8):
        for f in F: getattr(h, f)[:] = getattr(pristine, f)
        t0 = time.perf_counter(); dev.trace(h); ts.append(time.perf_counter() - t0)
    t = np.mean(ts[2:])
    print(f"{label:60s} {t*1e3:6.2f} ms  {n/t/1e6:7.0f} Mrays/s", flush=True)
KEYS = ("PHOS_E2E_DEBUG", "PHOS_E2E_SPARSE", "PHOS_PIPE_CHUNK", "PHOS_E2E_WB", "PHOS_E2E_WB_CTAS")
for ch in ("131072", "262144", "524288"):
    run(f"chunk {ch}: copy-everything (48 up / 24 down)", PHOS_E2E_SPARSE="0", PHOS_PIPE_CHUNK=ch)
    run(f"chunk {ch}: copy-everything, no upload", PHOS_E2E_SPARSE="0", PHOS_PIPE_CHUNK=ch, PHOS_E2E_DEBUG="noin")
    run(f"chunk {ch}: copy-everything, no download", PHOS_E2E_SPARSE="0", PHOS_PIPE_CHUNK=ch, PHOS_E2E_DEBUG="noout")
    run(f"chunk {ch}: sparse (32 up / zero-copy write-back, 4 CTAs)", PHOS_PIPE_CHUNK=ch)
    run(f"chunk {ch}: sparse, no upload", PHOS_PIPE_CHUNK=ch, PHOS_E2E_DEBUG="noin")
    run(f"chunk {ch}: sparse, no write-back", PHOS_PIPE_CHUNK=ch, PHOS_E2E_DEBUG="noout")
    for c in ("1", "2", "8", "16", "1024"):
        run(f"chunk {ch}: sparse, {c} write-back CTAs", PHOS_PIPE_CHUNK=ch, PHOS_E2E_WB_CTAS=c)
