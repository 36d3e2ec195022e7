"""Synthetic scenes for the BASELINE.json configs, built through the scene API.

All generators are deterministic (fixed seeds) and numpy-vectorised so the 10-30 M triangle
configs build in seconds.  Windings follow the reference's convention that normals are never
face-forwarded (src/mesh.cpp:202-215): walls face into the room, boxes face
outwards, lights face down.
"""
from __future__ import annotations

import math

import numpy as np

from .scene import MAT_DIFFUSE, MAT_EMITTER, MAT_GLOSSY, Camera, Material, Mesh, Scene


def _quad(verts, faces, a, b, c, d, want_normal):
    """Append quad a-b-c-d as two triangles whose geometric normal points along want_normal."""
    a, b, c, d = (np.asarray(v, np.float64) for v in (a, b, c, d))
    n = np.cross(b - a, c - a)
    if np.dot(n, want_normal) < 0:
        a, b, c, d = a, d, c, b
    i = len(verts)
    verts.extend([a, b, c, d])
    faces.append((i, i + 1, i + 2))
    faces.append((i, i + 2, i + 3))


def _box(verts, faces, lo, hi, rot_y=0.0, outward=True):
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    c = 0.5 * (lo + hi)
    h = 0.5 * (hi - lo)
    cs, sn = math.cos(rot_y), math.sin(rot_y)

    def P(sx, sy, sz):
        x, y, z = sx * h[0], sy * h[1], sz * h[2]
        return np.array([c[0] + cs * x + sn * z, c[1] + y, c[2] - sn * x + cs * z])

    def N(x, y, z):
        v = np.array([cs * x + sn * z, y, -sn * x + cs * z])
        return v if outward else -v

    _quad(verts, faces, P(-1, -1, 1), P(1, -1, 1), P(1, 1, 1), P(-1, 1, 1), N(0, 0, 1))
    _quad(verts, faces, P(-1, -1, -1), P(1, -1, -1), P(1, 1, -1), P(-1, 1, -1), N(0, 0, -1))
    _quad(verts, faces, P(-1, -1, -1), P(-1, -1, 1), P(-1, 1, 1), P(-1, 1, -1), N(-1, 0, 0))
    _quad(verts, faces, P(1, -1, -1), P(1, -1, 1), P(1, 1, 1), P(1, 1, -1), N(1, 0, 0))
    _quad(verts, faces, P(-1, 1, -1), P(1, 1, -1), P(1, 1, 1), P(-1, 1, 1), N(0, 1, 0))
    _quad(verts, faces, P(-1, -1, -1), P(1, -1, -1), P(1, -1, 1), P(-1, -1, 1), N(0, -1, 0))


def cornell_box(width: int = 512, height: int = 512, light_y: float = 1.98) -> Scene:
    """Config 1: 38-triangle Cornell box, two area lights (SURVEY.md §8d "Config 1").

    5 wall quads (10 tris), short box + tall box (12 each), 2 emissive quads under the ceiling (4).
    Room spans x in [-1,1], y in [0,2], z in [-1,1], open towards +z; camera at (0,1,3.8) looking -z.
    """
    s = Scene()
    white = s.add_material(Material(MAT_DIFFUSE, (0.73, 0.73, 0.73)))
    red = s.add_material(Material(MAT_DIFFUSE, (0.65, 0.05, 0.05)))
    green = s.add_material(Material(MAT_DIFFUSE, (0.12, 0.45, 0.15)))
    glossy = s.add_material(Material(MAT_GLOSSY, (0.73, 0.73, 0.73), roughness=0.3))
    light = s.add_material(Material(MAT_EMITTER, (1.0, 1.0, 1.0), power=15.0))

    v, f = [], []
    _quad(v, f, (-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1), (0, 1, 0))  # floor
    _quad(v, f, (-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), (0, -1, 0))  # ceiling
    _quad(v, f, (-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), (0, 0, 1))  # back
    _quad(v, f, (-1, 0, -1), (-1, 0, 1), (-1, 2, 1), (-1, 2, -1), (1, 0, 0))  # left (red)
    _quad(v, f, (1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), (-1, 0, 0))  # right (green)
    s.add(Mesh(np.array(v), np.array(f), [(white, np.arange(0, 6)), (red, np.arange(6, 8)), (green, np.arange(8, 10))]))

    v, f = [], []
    _box(v, f, (0.15, 0.0, 0.05), (0.75, 0.6, 0.65), rot_y=math.radians(-17.0))
    s.add(Mesh(np.array(v), np.array(f), [(white, np.arange(12))]))

    v, f = [], []
    _box(v, f, (-0.75, 0.0, -0.65), (-0.15, 1.2, -0.05), rot_y=math.radians(20.0))
    s.add(Mesh(np.array(v), np.array(f), [(glossy, np.arange(12))]))

    v, f = [], []
    # (light_y well below the ceiling takes the 1 / d^2 spikes of next-event estimation onto the ceiling 2 cm above the
    # lights out of the image: plain means then converge fast enough for percent-level comparisons)
    _quad(v, f, (-0.6, light_y, -0.25), (-0.1, light_y, -0.25), (-0.1, light_y, 0.25), (-0.6, light_y, 0.25), (0, -1, 0))
    _quad(v, f, (0.1, light_y, -0.25), (0.6, light_y, -0.25), (0.6, light_y, 0.25), (0.1, light_y, 0.25), (0, -1, 0))
    # two face sets -> two area lights (one light per emissive face set, src/mesh.cpp:108-116)
    s.add(Mesh(np.array(v), np.array(f), [(light, np.arange(0, 2)), (light, np.arange(2, 4))]))

    cam = Camera(fov=math.radians(65.0), film_width=width, film_height=height)  # reference NDC spans +-0.5: = 39 deg conventional
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = (0.0, 1.0, 3.8)
    cam.to_world = m
    s.camera = cam
    return s


def uv_sphere(center, radius, seg_u=64, seg_v=32):
    """UV sphere with seg_u x seg_v quads x 2 triangles (pole quads degenerate to zero-area tris)."""
    u = np.linspace(0.0, 2.0 * np.pi, seg_u + 1)
    t = np.linspace(0.0, np.pi, seg_v + 1)
    uu, tt = np.meshgrid(u, t, indexing="xy")  # (seg_v+1, seg_u+1)
    n = np.stack([np.sin(tt) * np.cos(uu), np.cos(tt), np.sin(tt) * np.sin(uu)], axis=-1).reshape(-1, 3)
    verts = (np.asarray(center, np.float64) + radius * n).astype(np.float32)
    j, i = np.meshgrid(np.arange(seg_v), np.arange(seg_u), indexing="ij")
    a = (j * (seg_u + 1) + i).ravel()
    b = a + 1
    c = a + (seg_u + 1)
    d = c + 1
    # outward winding: (a, b, d), (a, d, c) has normal along +n for this parametrisation
    faces = np.concatenate([np.stack([a, d, b], 1), np.stack([a, c, d], 1)], axis=1).reshape(-1, 3)
    return verts, faces.astype(np.uint32), n.astype(np.float32)


def sphere_field(grid: int = 16, seg_u: int = 64, seg_v: int = 32, width: int = 1920, height: int = 1080,
                 smooth: bool = False) -> Scene:
    """Config 2: grid x grid UV spheres, radius 0.4 on a unit grid, one mesh per sphere.

    Default 16 x 16 x (64 x 32 x 2) = 1 048 576 triangles, single diffuse material, 1920 x 1080
    pinhole camera looking at the field at a slant so primary rays are coherent and mostly hit.
    """
    s = Scene()
    grey = s.add_material(Material(MAT_DIFFUSE, (0.7, 0.7, 0.7)))
    nf = seg_u * seg_v * 2
    half = 0.5 * (grid - 1)
    for gy in range(grid):
        for gx in range(grid):
            v, f, n = uv_sphere((gx - half, 0.0, gy - half), 0.4, seg_u, seg_v)
            s.add(Mesh(v, f, [(grey, np.arange(nf))], smooth=smooth, normals=n if smooth else None))
    cam = Camera(fov=math.radians(50.0), film_width=width, film_height=height)
    ext = float(grid)
    cam.to_world = Camera.look_at((0.0, 0.62 * ext, 0.78 * ext), (0.0, 0.0, -0.02 * ext))
    s.camera = cam
    return s


def mixed_shading_spheres(width: int = 64, height: int = 48) -> Scene:
    """2 x 2 spheres whose faces alternate between smooth and flat shading in bands of rows — the per-face flag of
    mesh_t::builder_t::add_face(a, b, c, smooth) (src/mesh.hpp:46-66, src/mesh.cpp:202-206) — under an emissive quad."""
    s = sphere_field(2, 24, 12, width, height, smooth=True)
    for k, m in enumerate(s.meshes):
        band = (np.arange(len(m.faces)) // 48 + k) % 2  # 24 quads x 2 triangles per row of the sphere
        m.face_smooth = band.astype(np.uint8)
        m.smooth = False
    light = s.add_material(Material(MAT_EMITTER, (1.0, 0.9, 0.8), power=30.0))
    v = np.array([[-1.5, 2.0, -1.5], [1.5, 2.0, -1.5], [1.5, 2.0, 1.5], [-1.5, 2.0, 1.5]], np.float32)
    nrm = np.tile(np.array([[0, -1, 0]], np.float32), (4, 1))
    s.add(Mesh(v, np.array([[0, 1, 2], [0, 2, 3]]), [(light, np.arange(2))], smooth=False, normals=nrm))
    return s


def _value_noise(x, y, seed):
    """2-D value noise on an integer lattice with smoothstep interpolation (vectorised)."""
    xi, yi = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    fx, fy = x - xi, y - yi

    def h(ix, iy):
        k = (ix * 374761393 + iy * 668265263 + seed * 2147483647) & 0xFFFFFFFF
        k = ((k ^ (k >> 13)) * 1274126177) & 0xFFFFFFFF
        k = k ^ (k >> 16)
        return (k & 0xFFFFFF).astype(np.float64) / float(1 << 24)

    sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    v00, v10, v01, v11 = h(xi, yi), h(xi + 1, yi), h(xi, yi + 1), h(xi + 1, yi + 1)
    return (v00 * (1 - sx) + v10 * sx) * (1 - sy) + (v01 * (1 - sx) + v11 * sx) * sy


def terrain(n: int = 2237, extent: float = 100.0, seed: int = 1234, width: int = 1920, height: int = 1080,
            glossy_fraction: float = 0.0) -> Scene:
    """Config 3/4: n x n vertex displaced-terrain grid, (n-1)^2 x 2 triangles.

    n = 2237 gives 9 999 392 triangles.  Height is a 5-octave value-noise fBm (seed 1234) of
    amplitude 0.15 x extent.  One terrain mesh (diffuse, optionally a GGX face set covering
    `glossy_fraction` of the faces in whole-row bands) plus an emissive sky quad pair facing down.
    """
    s = Scene()
    ground = s.add_material(Material(MAT_DIFFUSE, (0.55, 0.5, 0.42)))
    sky = s.add_material(Material(MAT_EMITTER, (1.0, 1.0, 1.0), power=40.0))
    ggx = s.add_material(Material(MAT_GLOSSY, (0.8, 0.8, 0.85), roughness=0.2)) if glossy_fraction > 0 else None

    lin = np.linspace(-0.5 * extent, 0.5 * extent, n)
    xx, zz = np.meshgrid(lin, lin, indexing="xy")
    hgt = np.zeros_like(xx)
    amp, freq = 1.0, 6.0 / extent
    norm = 0.0
    for o in range(5):
        hgt += amp * _value_noise(xx * freq + 17.0 * o, zz * freq - 31.0 * o, seed + o)
        norm += amp
        amp *= 0.5
        freq *= 2.0
    hgt = (hgt / norm - 0.5) * 2.0 * 0.15 * extent
    verts = np.stack([xx, hgt, zz], axis=-1).reshape(-1, 3).astype(np.float32)
    j, i = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = (j * n + i).ravel()
    b, c = a + 1, a + n
    d = c + 1
    # upward (+y) normals: (a, c, b): (c-a) x (b-a) = (0,0,dz) x (dx,0,0) -> +y
    faces = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, d], 1)], axis=1).reshape(-1, 3).astype(np.uint32)
    nf = len(faces)
    if ggx is not None:
        rows = n - 1
        per_row = 2 * (n - 1)
        band = max(1, int(round(1.0 / glossy_fraction)))
        row_of_face = np.arange(nf) // per_row
        is_g = (row_of_face % band) == 0
        sets = [(ground, np.nonzero(~is_g)[0]), (ggx, np.nonzero(is_g)[0])]
        del rows
    else:
        sets = [(ground, np.arange(nf))]
    s.add(Mesh(verts, faces, sets))

    v, f = [], []
    e = 0.35 * extent
    y = 0.6 * extent
    _quad(v, f, (-e, y, -e), (e, y, -e), (e, y, e), (-e, y, e), (0, -1, 0))
    s.add(Mesh(np.array(v), np.array(f), [(sky, np.arange(2))]))

    cam = Camera(fov=math.radians(55.0), film_width=width, film_height=height)
    cam.to_world = Camera.look_at((0.0, 0.33 * extent, 0.62 * extent), (0.0, -0.05 * extent, 0.0))
    s.camera = cam
    return s


def heightfield(g: int, seed: int = 7) -> Scene:
    """Small g x g x 2 triangle heightfield used by parity tests (the survey's probe mesh family)."""
    return terrain(n=g + 1, extent=10.0, seed=seed, width=256, height=256)


def instanced_field(grid: int = 86, width: int = 3840, height: int = 2160) -> Scene:
    """Config 5: ~30 M triangles as world-space-baked copies of one 4096-triangle object (the reference
    has no instancing: every copy is its own mesh_t), lit by one large emissive quad above the field.
    grid = 86 -> 7396 copies = 30 294 016 triangles (+ 2 for the light)."""
    s = sphere_field(grid=grid, width=width, height=height)
    light = s.add_material(Material(MAT_EMITTER, (1.0, 0.95, 0.9), power=60.0))
    e = 0.5 * grid
    v, f = [], []
    _quad(v, f, (-e, 0.45 * grid, -e), (e, 0.45 * grid, -e), (e, 0.45 * grid, e), (-e, 0.45 * grid, e), (0, -1, 0))
    s.add(Mesh(np.array(v), np.array(f), [(light, np.arange(2))]))
    return s


def cornell_lobes(width: int = 64, height: int = 64, environment: bool = True) -> Scene:
    """The Cornell box re-dressed with every closure of the widened subset: Oren-Nayar walls,
    a mirror short box, a glass (sharp refraction) + sheen mix on the tall box, a diffuse / GGX mix on the floor
    side walls, a transparent panel in front, and a constant background seen through the open front."""
    from .scene import (LOBE_REFRACTION, LOBE_SHEEN, LOBE_TRANSPARENT, MAT_BACKGROUND, MAT_DIFFUSE, MAT_GLOSSY, MAT_LAYERED,
                        Material)
    sc = cornell_box(width, height)
    white, red, green, box = 0, 1, 2, 3  # cornell_box's material order: white, red, green, tall box, emitter
    sc.materials[white] = Material(MAT_DIFFUSE, (0.73, 0.73, 0.73), roughness=20.0)  # oren_nayar(N, 20)
    sc.materials[red] = Material.mix(Material(MAT_DIFFUSE, (0.65, 0.05, 0.05)), Material(MAT_GLOSSY, (0.9, 0.9, 0.9), roughness=0.25), 0.3)
    sc.materials[green] = Material.mix(Material(MAT_DIFFUSE, (0.12, 0.45, 0.15)),
                                       Material(MAT_LAYERED, lobes=((LOBE_SHEEN, (0.8, 0.8, 0.8), 0.4),)), 0.5)
    sc.materials[box] = Material.mix(Material(MAT_GLOSSY, (0.9, 0.9, 0.9), roughness=0.0),
                                     Material(MAT_LAYERED, lobes=((LOBE_REFRACTION, (0.9, 0.95, 1.0), 1.5),)), 0.6)
    if environment:
        sc.environment = sc.add_material(Material(MAT_BACKGROUND, (0.3, 0.4, 0.6), power=0.5))
    return sc
