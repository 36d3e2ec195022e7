"""Comparison helpers shared by the parity tests."""
from __future__ import annotations

import os

import numpy as np

from phosphorus_mk2_b200.rays import HIT, MASKED, SHADOW, RayBatch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags")
OUTS = ("d", "u", "v", "mesh", "face", "flags")
BATCHES = ("aimed", "random", "shadow", "mixed", "special")
CASES = ("cornell", "heightfield24", "spheres2")


def load_golden(case: str):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    return z


def golden_rays(z, batch: str) -> RayBatch:
    n = len(z[f"{batch}_rays_px"])
    r = RayBatch(n)
    for f in FIELDS:
        getattr(r, f)[:] = z[f"{batch}_rays_{f}"]
    return r


def golden_out(z, batch: str, kind: str, rays: RayBatch) -> RayBatch:
    r = rays.copy()
    for f in OUTS:
        getattr(r, f)[:] = z[f"{batch}_{kind}_{f}"]
    return r


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def mismatches(inp: RayBatch, got: RayBatch, want: RayBatch) -> np.ndarray:
    """Indices of rays whose result differs from `want` under the path's parity rule:
      closest-hit rays : flags, d, and for hit rays mesh, face, u, v — all bit-exact;
      SHADOW rays      : flags bit-exact (the HIT verdict); a ray that is not hit keeps d; a hit ray
                         holds SOME accepted distance < tmax (which one is traversal-order dependent
                         in the reference itself, stream_bvh_kernel.cpp:61-64); surface untouched;
      MASKED rays      : every field untouched."""
    shadow = (inp.flags & SHADOW) != 0
    masked = (inp.flags & MASKED) != 0
    bad = got.flags != want.flags
    hit = (want.flags & HIT) != 0
    closest = ~shadow & ~masked
    bad |= closest & (bits(got.d) != bits(want.d))
    for f in ("mesh", "face", "u", "v"):
        bad |= closest & hit & (bits(getattr(got, f)) != bits(getattr(want, f)))
        bad |= (shadow | masked | (closest & ~hit)) & (bits(getattr(got, f)) != bits(getattr(inp, f)))
    bad |= shadow & ~masked & ~hit & (bits(got.d) != bits(inp.d))
    bad |= shadow & ~masked & hit & ~((got.d < inp.d) & (got.d >= 0))
    bad |= masked & (bits(got.d) != bits(inp.d))
    return np.nonzero(bad)[0]


def classify_vs_stream(inp: RayBatch, stream: RayBatch, exact: RayBatch):
    """Split disagreements between the reference stream kernel and the exact result into
    (stream_missed, stream_farther, tied, other) index arrays."""
    shadow = (inp.flags & SHADOW) != 0
    masked = (inp.flags & MASKED) != 0
    closest = ~shadow & ~masked
    sh, eh = (stream.flags & HIT) != 0, (exact.flags & HIT) != 0
    differ = closest & ((sh != eh) | (eh & ((stream.mesh != exact.mesh) | (stream.face != exact.face))))
    missed = differ & eh & ~sh
    farther = differ & eh & sh & (stream.d > exact.d)
    tied = differ & eh & sh & (stream.d == exact.d)
    other = differ & ~missed & ~farther & ~tied
    return tuple(np.nonzero(x)[0] for x in (missed, farther, tied, other))
