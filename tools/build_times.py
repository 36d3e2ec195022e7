#!/usr/bin/env python3
"""GPU box: preprocess time of the two builders (host: reference-identical SAH build + re-pack + upload; device:
phos_cuda_build_accel) and the tree each produces, on the BASELINE scenes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options

for name, make in (("spheres 1M", scenes.sphere_field), ("terrain 10M", scenes.terrain), ("instanced 30M", scenes.instanced_field)):
    sc = make()
    dev = CudaDevice.make(Options(), 0)
    t0 = time.perf_counter(); acc = Accel(sc); t1 = time.perf_counter()
    dev.preprocess(sc, acc); t2 = time.perf_counter()
    st = dev.accel_stats()
    print(f"{name:14s} host  : build {t1 - t0:6.2f} s + re-pack {st.repack_seconds:5.2f} s + upload {st.upload_seconds:5.2f} s = {t2 - t0:6.2f} s; "
          f"{st.nodes} nodes, depth {st.max_depth}", flush=True)
    del acc
    t0 = time.perf_counter(); dev.build_accel(sc); t1 = time.perf_counter()
    st = dev.accel_stats()
    print(f"{name:14s} device: flatten + copy-in {st.upload_seconds:5.2f} s + build {st.repack_seconds * 1e3:7.1f} ms = {t1 - t0:6.2f} s; "
          f"{st.nodes} nodes, depth {st.max_depth}", flush=True)
    dev.close()
