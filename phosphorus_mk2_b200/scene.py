"""Host-side scene model: the reference's scene/mesh/material/camera API, flattened for the C ABI.

Mirrors the names and meaning of the reference's public scene API so tests read like the
reference's own callers (reference: src/scene.hpp:14-50 ``scene_t``, src/mesh.hpp:15-138
``mesh_t`` + ``builder_t``, src/material.hpp ``material_t``, src/entities/camera.hpp:10-40
``camera_t``).  Geometry is held in numpy arrays (10-30 M triangle scenes) and handed to the native
library as one ``phos_scene_desc`` (include/phos_scene.h).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

MAT_DIFFUSE, MAT_GLOSSY, MAT_EMITTER, MAT_BACKGROUND, MAT_LAYERED = 0, 1, 2, 3, 4
# closure ids = bsdf_t::type_t (src/bsdf.hpp:14-24)
LOBE_DIFFUSE, LOBE_OREN_NAYAR, LOBE_REFLECTION, LOBE_REFRACTION, LOBE_MICROFACET, LOBE_SHEEN, LOBE_TRANSPARENT = 1, 2, 4, 8, 16, 32, 128
LOBE_MICROFACET_REFRACT = 16 | 256  # GGX transmission (rough glass): (type, weight, roughness, eta)
MAX_LOBES = 8


class PhosLobe(C.Structure):
    _fields_ = [("type", C.c_uint32), ("weight", C.c_float * 3), ("param", C.c_float), ("param2", C.c_float)]


class PhosMaterial(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("cs", C.c_float * 3), ("roughness", C.c_float), ("power", C.c_float),
                ("num_lobes", C.c_uint32), ("lobes", PhosLobe * MAX_LOBES)]


class PhosCamera(C.Structure):
    _fields_ = [
        ("to_world", C.c_float * 16),
        ("fov", C.c_float),
        ("focal_distance", C.c_float),
        ("aperture_radius", C.c_float),
        ("film_width", C.c_uint32),
        ("film_height", C.c_uint32),
    ]


class PhosSceneDesc(C.Structure):
    _fields_ = [
        ("num_meshes", C.c_uint32),
        ("vert_offset", C.POINTER(C.c_uint32)),
        ("vertices", C.POINTER(C.c_float)),
        ("normals", C.POINTER(C.c_float)),
        ("face_offset", C.POINTER(C.c_uint32)),
        ("faces", C.POINTER(C.c_uint32)),
        ("mesh_smooth", C.POINTER(C.c_uint8)),
        ("set_offset", C.POINTER(C.c_uint32)),
        ("set_material", C.POINTER(C.c_uint32)),
        ("set_face_offset", C.POINTER(C.c_uint32)),
        ("set_faces", C.POINTER(C.c_uint32)),
        ("num_materials", C.c_uint32),
        ("materials", C.POINTER(PhosMaterial)),
        ("camera", PhosCamera),
        ("environment", C.c_int32),
        ("face_smooth", C.POINTER(C.c_uint8)),
    ]


@dataclass
class Material:
    """One of the renderer's built-in closures (include/phos_scene.h PHOS_MAT_*)."""

    kind: int
    cs: tuple = (1.0, 1.0, 1.0)
    roughness: float = 0.0
    power: float = 1.0
    lobes: tuple = ()  # MAT_LAYERED: ((LOBE_*, (r, g, b), param), ...), at most 8

    def is_emitter(self) -> bool:  # material_t::is_emitter, src/material.cpp:487-489
        return self.kind == MAT_EMITTER

    @staticmethod
    def mix(a: "Material", b: "Material", fac: float) -> "Material":
        """mix_closure_node.osl:20: A * (1 - fac) + B * fac, flattened the way eval_closure walks the MUL / ADD
        tree (src/material.cpp:218-305): A's closures first with weights * (1 - fac), then B's * fac."""
        out = []
        for m, k in ((a, np.float32(1.0) - np.float32(fac)), (b, np.float32(fac))):
            for lobe in m.closures():
                t, w = lobe[0], lobe[1]
                out.append((t, tuple(float(np.float32(k) * np.float32(c)) for c in w)) + tuple(lobe[2:]))
        assert len(out) <= MAX_LOBES
        return Material(MAT_LAYERED, lobes=tuple(out))

    def closures(self):
        """The closure list the node hands to eval_closure: (type, weight, param) per lobe."""
        if self.kind == MAT_LAYERED:
            return list(self.lobes)
        if self.kind == MAT_DIFFUSE:  # diffuse_bsdf_node.osl:20-25
            return [(LOBE_DIFFUSE, self.cs, 0.0)] if self.roughness == 0 else [(LOBE_OREN_NAYAR, self.cs, self.roughness)]
        if self.kind == MAT_GLOSSY:  # glossy_bsdf_node.osl:26-34
            r2 = float(np.float32(self.roughness) * np.float32(self.roughness))
            return [(LOBE_REFLECTION, self.cs, 0.0)] if self.roughness == 0 else [(LOBE_MICROFACET, self.cs, r2)]
        return []


@dataclass
class Mesh:
    """mesh_t: vertices, faces and face sets (material id -> faces), src/mesh.hpp:68-77."""

    vertices: np.ndarray  # (nv, 3) float32
    faces: np.ndarray  # (nf, 3) uint32, mesh-local
    sets: list  # [(material_id, face_index_array)] in set order
    smooth: bool = False
    normals: np.ndarray | None = None  # (nv, 3) float32, per vertex
    face_smooth: np.ndarray | None = None  # (nf,) bool: builder_t::add_face(a, b, c, smooth) per face (overrides `smooth`)

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32).reshape(-1, 3)
        self.faces = np.ascontiguousarray(self.faces, dtype=np.uint32).reshape(-1, 3)
        self.sets = [(int(m), np.ascontiguousarray(f, dtype=np.uint32).ravel()) for m, f in self.sets]
        if self.normals is not None:
            self.normals = np.ascontiguousarray(self.normals, dtype=np.float32).reshape(-1, 3)
            assert len(self.normals) == len(self.vertices)
        if self.face_smooth is not None:
            self.face_smooth = np.ascontiguousarray(self.face_smooth, dtype=np.uint8).ravel()
            assert len(self.face_smooth) == len(self.faces)
            self.smooth = bool(self.face_smooth.all())
        if self.smooth or (self.face_smooth is not None and self.face_smooth.any()):
            assert self.normals is not None, "smooth faces interpolate per-vertex normals"

    @property
    def num_faces(self) -> int:
        return len(self.faces)


@dataclass
class Camera:
    """camera_t. ``to_world`` uses Imath's row-vector convention (translation in row 3)."""

    to_world: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    fov: float = math.radians(39.3)
    film_width: int = 512
    film_height: int = 512
    focal_distance: float = 1.0
    aperture_radius: float = 0.0

    @staticmethod
    def look_at(eye, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
        """Row-vector camera-to-world matrix for a camera looking down its local -z."""
        eye = np.asarray(eye, dtype=np.float64)
        f = np.asarray(target, dtype=np.float64) - eye
        f /= np.linalg.norm(f)
        r = np.cross(f, np.asarray(up, dtype=np.float64))
        r /= np.linalg.norm(r)
        u = np.cross(r, f)
        m = np.eye(4)
        m[0, :3], m[1, :3], m[2, :3], m[3, :3] = r, u, -f, eye
        return m.astype(np.float32)


class Scene:
    """scene_t: flat list of meshes, materials and one camera (src/scene.hpp:14-50)."""

    def __init__(self):
        self.meshes: list[Mesh] = []
        self.materials: list[Material] = []
        self.camera = Camera()
        self.environment: int | None = None  # material id of the environment (MAT_BACKGROUND), scene_t::environment()
        self._keep = None

    # scene_t::add(name, material) / add(mesh): ids are assigned in insertion order (scene.cpp:85-98)
    def add_material(self, material: Material) -> int:
        self.materials.append(material)
        return len(self.materials) - 1

    def add(self, mesh: Mesh) -> int:
        assert len(self.meshes) < 0xFFFF, "meshid is packed into 16 bits (src/accel/triangle.hpp:62-66)"
        self.meshes.append(mesh)
        return len(self.meshes) - 1

    def num_meshes(self) -> int:
        return len(self.meshes)

    def num_materials(self) -> int:
        return len(self.materials)

    def num_triangles(self) -> int:
        return sum(sum(len(f) for _, f in m.sets) for m in self.meshes)

    def desc(self) -> PhosSceneDesc:
        """Flatten into a phos_scene_desc; the backing arrays stay alive on ``self``."""
        nm = len(self.meshes)
        vert_offset = np.zeros(nm + 1, np.uint32)
        face_offset = np.zeros(nm + 1, np.uint32)
        set_offset = np.zeros(nm + 1, np.uint32)
        for i, m in enumerate(self.meshes):
            vert_offset[i + 1] = vert_offset[i] + len(m.vertices)
            face_offset[i + 1] = face_offset[i] + len(m.faces)
            set_offset[i + 1] = set_offset[i] + len(m.sets)
        vertices = np.concatenate([m.vertices for m in self.meshes]).astype(np.float32, copy=False)
        has_normals = all(m.normals is not None for m in self.meshes)
        normals = np.concatenate([m.normals for m in self.meshes]) if has_normals else None
        def smooth_kind(m):  # 0 flat, 1 smooth, 2 mixed (per face)
            if m.face_smooth is None or m.face_smooth.all() or not m.face_smooth.any():
                return 1 if (m.smooth or (m.face_smooth is not None and len(m.face_smooth) and m.face_smooth.all())) else 0
            return 2

        smooth = np.array([smooth_kind(m) for m in self.meshes], np.uint8)
        if not has_normals:
            assert not smooth.any(), "smooth faces need normals on every mesh"
        faces = np.concatenate([m.faces for m in self.meshes]).astype(np.uint32, copy=False)
        face_smooth = None
        if (smooth == 2).any():
            face_smooth = np.concatenate([m.face_smooth if m.face_smooth is not None else np.full(len(m.faces), 1 if m.smooth else 0, np.uint8)
                                          for m in self.meshes]).astype(np.uint8)
        set_material, set_faces_l = [], []
        for m in self.meshes:
            for mat, f in m.sets:
                assert 0 <= mat < len(self.materials)
                set_material.append(mat)
                set_faces_l.append(f)
        set_material = np.array(set_material, np.uint32)
        set_face_offset = np.zeros(len(set_faces_l) + 1, np.uint32)
        set_face_offset[1:] = np.cumsum([len(f) for f in set_faces_l])
        set_faces = np.concatenate(set_faces_l).astype(np.uint32, copy=False) if set_faces_l else np.zeros(0, np.uint32)
        mats = (PhosMaterial * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            mats[i].kind = m.kind
            mats[i].cs = (C.c_float * 3)(*m.cs)
            mats[i].roughness = m.roughness
            mats[i].power = m.power
            mats[i].num_lobes = len(m.lobes)
            for k, lobe in enumerate(m.lobes):
                t, w, prm = lobe[:3]
                mats[i].lobes[k].type = t
                mats[i].lobes[k].weight = (C.c_float * 3)(*w)
                mats[i].lobes[k].param = prm
                mats[i].lobes[k].param2 = lobe[3] if len(lobe) > 3 else 0.0

        def p(a, t):
            return a.ctypes.data_as(C.POINTER(t)) if a is not None else C.POINTER(t)()

        vertices = np.ascontiguousarray(vertices)
        faces = np.ascontiguousarray(faces)
        d = PhosSceneDesc()
        d.num_meshes = nm
        d.vert_offset = p(vert_offset, C.c_uint32)
        d.vertices = p(vertices, C.c_float)
        normals = np.ascontiguousarray(normals, dtype=np.float32) if normals is not None else None
        d.normals = p(normals, C.c_float)
        d.face_offset = p(face_offset, C.c_uint32)
        d.faces = p(faces, C.c_uint32)
        d.mesh_smooth = p(smooth, C.c_uint8)
        d.set_offset = p(set_offset, C.c_uint32)
        d.set_material = p(set_material, C.c_uint32)
        d.set_face_offset = p(set_face_offset, C.c_uint32)
        d.set_faces = p(set_faces, C.c_uint32)
        d.num_materials = len(self.materials)
        d.materials = C.cast(mats, C.POINTER(PhosMaterial))
        cam = self.camera
        d.camera.to_world = (C.c_float * 16)(*np.asarray(cam.to_world, np.float32).ravel())
        d.camera.fov = cam.fov
        d.camera.focal_distance = cam.focal_distance
        d.camera.aperture_radius = cam.aperture_radius
        d.camera.film_width = cam.film_width
        d.camera.film_height = cam.film_height
        d.environment = -1 if self.environment is None else int(self.environment)
        d.face_smooth = p(face_smooth, C.c_uint8)
        self._keep = (face_smooth, vert_offset, vertices, normals, face_offset, faces, smooth, set_offset, set_material,
                      set_face_offset, set_faces, mats)
        return d
