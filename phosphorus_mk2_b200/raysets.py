"""Deterministic synthetic ray batches (numpy) for parity tests and benches."""
from __future__ import annotations

import numpy as np

from .rays import RayBatch, SHADOW, MASKED


def scene_bounds(scene):
    lo = np.min([m.vertices.min(axis=0) for m in scene.meshes], axis=0)
    hi = np.max([m.vertices.max(axis=0) for m in scene.meshes], axis=0)
    return lo.astype(np.float64), hi.astype(np.float64)


def aimed_rays(scene, n: int, seed: int = 1) -> RayBatch:
    """Incoherent rays: origins uniform in the scene's bounding box inflated by 50 %, each aimed at
    a uniformly chosen vertex of the scene jittered by 1 % of the extent (nearly all hit)."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(scene)
    ext = hi - lo
    org = lo - 0.25 * ext + rng.random((n, 3)) * 1.5 * ext
    sizes = np.array([len(m.vertices) for m in scene.meshes], np.float64)
    mesh_idx = rng.choice(len(scene.meshes), size=n, p=sizes / sizes.sum())
    tgt = np.empty((n, 3))
    for mi in np.unique(mesh_idx):
        sel = np.nonzero(mesh_idx == mi)[0]
        v = scene.meshes[mi].vertices
        tgt[sel] = v[rng.integers(0, len(v), len(sel))]
    tgt += (rng.random((n, 3)) - 0.5) * 0.02 * ext
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return RayBatch.from_arrays(org.astype(np.float32), d.astype(np.float32))


def random_rays(scene, n: int, seed: int = 2) -> RayBatch:
    """Incoherent rays with uniformly random directions from origins inside the inflated bounds
    (a mix of hits and misses)."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bounds(scene)
    ext = hi - lo
    org = lo - 0.1 * ext + rng.random((n, 3)) * 1.2 * ext
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return RayBatch.from_arrays(org.astype(np.float32), d.astype(np.float32))


def as_shadow(rays: RayBatch, seed: int = 3, masked_fraction: float = 0.1) -> RayBatch:
    """Turn a batch into occlusion queries: finite tmax around the typical hit distance, SHADOW set,
    a fraction additionally MASKED (never traced), as spt::light_sampler_t emits them
    (src/kernels/cpu/spt.hpp:138-143)."""
    rng = np.random.default_rng(seed)
    out = rays.copy()
    lo = np.array([out.px.min(), out.py.min(), out.pz.min()])
    hi = np.array([out.px.max(), out.py.max(), out.pz.max()])
    diag = float(np.linalg.norm(hi - lo))
    out.d[:] = (rng.random(out.n) * diag).astype(np.float32)
    out.flags[:] = SHADOW
    out.flags[rng.random(out.n) < masked_fraction] |= MASKED
    out.mesh[:] = 0xABCD0001  # stands for the sampled light's ids: must come back untouched
    out.face[:] = 777
    out.u[:] = 0.25
    out.v[:] = 0.5
    return out
