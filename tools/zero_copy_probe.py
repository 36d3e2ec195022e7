#!/usr/bin/env python3
"""GPU box probe: trace_kernel reading its input arrays straight from page-locked HOST memory (TMA bulk copies over PCIe,
no separate up-copy, no staging in HBM) — how fast is the up-link when the SMs' copy units pull it, and does it interfere
with result stores less than the copy engine does?  Mixed pointer sets go through phos_cuda_trace_device unchanged (UVA)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phosphorus_mk2_b200 import scenes  # noqa: E402
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, PhosRays, make_tiles, pinned_ray_batch  # noqa: E402

IN = ("px", "py", "pz", "wx", "wy", "wz", "d", "flags")
OUT = ("mesh", "face", "u", "v")


def main():
    sc = scenes.sphere_field()
    cam = sc.camera
    n = cam.film_width * cam.film_height
    dev = CudaDevice.make(Options(), 0)
    acc = Accel(sc)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    dr = dev.device_rays(n)
    dev.camera_rays(make_tiles(cam.film_width, cam.film_height), dr)
    dev.synchronize()
    host0 = dr.download()          # pristine inputs (pageable)
    hb = pinned_ray_batch(n)       # page-locked slab the kernel reads
    ref = None

    def restore(which_host):
        for k in IN + OUT:
            getattr(hb, k)[:] = getattr(host0, k)
        dr.upload(host0)
        dev.synchronize()

    def struct(host_fields):
        s = PhosRays()
        hs = hb.as_struct()
        for k in IN + OUT:
            setattr(s, k, getattr(hs, k) if k in host_fields else getattr(dr.s, k))
        return s

    cases = [("all 12 arrays in HBM", ()),
             ("p, wi read from host (24 B/ray over PCIe)", ("px", "py", "pz", "wx", "wy", "wz")),
             ("p, wi, d, flags read from host, d / flags written back to host in place (32 B up)", IN),
             ("everything in host memory (inputs read, hit records scattered back)", IN + OUT)]
    for name, hf in cases:
        ts = []
        for rep in range(6):
            restore(hf)
            dev.flush_l2()
            s = struct(hf)
            dev.timer_begin()
            dev._check(dev._L.phos_cuda_trace_device(dev._ctx, C.byref(s), n))
            ts.append(dev.timer_end())
        # result check against the all-device run
        dev.synchronize()
        d = np.array(hb.d) if "d" in hf else np.array(dr.download().d)
        if ref is None:
            ref = d
        same = bool(np.array_equal(d, ref))
        ts = ts[1:]
        print(f"{name:90s} {np.mean(ts):7.3f} ms  min {min(ts):7.3f}  {n / np.mean(ts) / 1e3:7.1f} Mrays/s  same={same}", flush=True)
    dev.close()


if __name__ == "__main__":
    main()
