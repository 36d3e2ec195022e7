#!/usr/bin/env python3
"""GPU box probe: raw H2D / D2H rates of ray-stream arrays and the host-pointer trace pipeline."""
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
for f in F: getattr(h, f)[:] = getattr(pristine, f)
def t(fn, reps=5):
    fn(); dev.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    dev.synchronize(); return (time.perf_counter() - t0) / reps
up = t(lambda: dr.upload(h)); print(f"H2D 12 arrays ({48*n/1e6:.0f} MB): {up*1e3:.2f} ms = {48*n/up/1e9:.1f} GB/s")
dn = t(lambda: dr.download(h)); print(f"D2H 12 arrays ({48*n/1e6:.0f} MB): {dn*1e3:.2f} ms = {48*n/dn/1e9:.1f} GB/s")
for f in F: getattr(h, f)[:] = getattr(pristine, f)
def tr():
    h.d[:] = pristine.d; h.flags[:] = pristine.flags
    t0 = time.perf_counter(); dev.trace(h); return time.perf_counter() - t0
tr(); ts = [tr() for _ in range(5)]
print(f"phos_cuda_trace (host pointers, 48 B up + 24 B down per ray): {np.mean(ts)*1e3:.2f} ms = {n/np.mean(ts)/1e6:.0f} Mrays/s")
