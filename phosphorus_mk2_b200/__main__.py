"""`python -m phosphorus_mk2_b200 <options> scene.yaml` — the reference's command line (src/core.cpp:17-185) on the
B200 device: parse options, import the scene, discover devices, preprocess, start / join, finalize the film.

  -o <path>     output image (.pfm or .npy; the reference writes .exr through OpenImageIO)
  -s <samples>  samples per pixel          (default 16, src/options.hpp:7)
  -p <paths>    paths per sample           (default 16)
  -d <depth>    maximum path depth         (default 9)
  -n <path>     also write the NORMALS channel (parsed_options_t::render_normals) as .npy

There is no `-c` (CPU only) here: this package has no host renderer, and without a GPU the run fails."""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import codec, frame
from .device import CudaDevice, Options


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="phosphorus", description="usage: phosphorus <options> scene")
    ap.add_argument("-o", "--output", default="out.pfm")
    ap.add_argument("-s", "--spp", type=int, default=16)
    ap.add_argument("-p", "--paths", type=int, default=16)
    ap.add_argument("-d", "--depth", type=int, default=9)
    ap.add_argument("-n", "--normals", default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("scene")
    a = ap.parse_args(argv)

    print(f"Importing scene: {a.scene}")
    scene = codec.load_scene(a.scene)
    print("Discovering devices")
    options = Options(a.spp, a.paths, a.depth)
    devices = CudaDevice.discover(options)
    if not devices:
        print("no CUDA device: this renderer has no CPU fallback", file=sys.stderr)
        return 1
    dev = devices[0]  # one process drives one GPU; several GPUs = several ranks (bench.py --render, frame.py)
    for d in devices[1:]:
        d.close()
    cam = scene.camera
    sink = codec.FileFilm(cam.film_width, cam.film_height, a.output)
    state = frame.FrameState(frame.Tiles.make(cam.film_width, cam.film_height, 32), sink, a.spp, a.seed)
    print("Preprocessing")
    dev.preprocess(scene)
    dev.upload_scene(scene)
    if a.normals:
        dev.enable_normals()
    print("Rendering...")
    t0 = time.perf_counter()
    frame.join(frame.start(dev, scene, state))
    print(f"Rendering time: {time.perf_counter() - t0:.3f}")
    sink.finalize()
    if a.normals:
        np.save(a.normals, dev.film_read_normals())
    dev.close()
    print("Done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
