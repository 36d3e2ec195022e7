"""Scene ingest and film output without the reference's third-party stack.

``codec::scene::import`` (reference src/codecs/scene.cpp:41-76) reads a YAML file with four sections —
``materials`` (shader networks, src/codecs/scene/material.hpp:47-95), ``data`` (Alembic archives),
``camera`` and ``world`` (``environment: <material>``, scene.cpp:33-38) — and needs yaml-cpp, Alembic and OSL.
This module reads the same file layout with PyYAML and stands in for the two pieces that cannot exist here:

* materials: the shader network of a material is evaluated symbolically for the closure nodes of the built-in
  set (src/shaders/*_node.osl): every layer yields the closure list its OSL body would produce, ``connect``
  edges feed closure outputs into ``mix_closure_node`` / ``add_node`` inputs, and the last layer's list is what
  ``material_t::details_t::eval_closure`` (src/material.cpp:218-305) would flatten the closure tree to;
* geometry: ``data`` entries name Wavefront ``.obj`` files (triangles / convex polygons, ``usemtl`` selects
  the face set's material by name) or a procedural ``generator`` of :mod:`scenes`, instead of ``.abc``.

``FileFilm`` is ``film::file_t`` (src/film/file.cpp): ``add_tile`` + ``finalize``, writing a little-endian
RGB ``.pfm`` (or ``.npy``) instead of going through OpenImageIO.
"""
from __future__ import annotations

import math
import os
import threading

import numpy as np

from . import scenes
from .scene import (LOBE_DIFFUSE, LOBE_MICROFACET, LOBE_MICROFACET_REFRACT, LOBE_OREN_NAYAR, LOBE_REFLECTION, LOBE_REFRACTION, LOBE_SHEEN,
                    LOBE_TRANSPARENT, MAT_BACKGROUND, MAT_DIFFUSE, MAT_EMITTER, MAT_GLOSSY, MAT_LAYERED, MAX_LOBES, Camera,
                    Material, Mesh, Scene)


class SceneError(ValueError):
    pass


# ---- materials: shader networks -> closure lists ---------------------------------------------------------
def _param(p):
    """One entry of a layer's ``parameters`` list: {name, type: float|rgb|string, value}."""
    t = p.get("type", "float")
    v = p["value"]
    if t == "float":
        return float(v)
    if t == "rgb":
        if len(v) != 3:
            raise SceneError(f"rgb parameter {p.get('name')} needs three components")
        return tuple(float(c) for c in v)
    if t == "string":
        return str(v)
    raise SceneError(f"Unknown parameter type: {t}")  # material.hpp:71-73


def _scale(lobes, k):
    k = np.float32(k)
    return [(l[0], tuple(float(k * np.float32(c)) for c in l[1])) + tuple(l[2:]) for l in lobes]


def _layer_closures(node: str, prm: dict, inputs: dict):
    """What the OSL body of `node` assigns to its closure output.  Returns (kind, payload):
    ("bsdf", closure list) | ("emission", (Cs, power)) | ("background", (Cs, power))."""
    cs = prm.get("Cs", (1.0, 1.0, 1.0))
    rough = float(prm.get("roughness", 0.0))
    dist = prm.get("distribution", "ggx")
    if node == "diffuse_bsdf_node":  # diffuse_bsdf_node.osl:20-25
        return "bsdf", [(LOBE_DIFFUSE, cs, 0.0)] if rough == 0 else [(LOBE_OREN_NAYAR, cs, rough)]
    if node == "glossy_bsdf_node":  # glossy_bsdf_node.osl:26-34
        if dist == "sharp" or rough == 0.0:
            return "bsdf", [(LOBE_REFLECTION, cs, 0.0)]
        return "bsdf", [(LOBE_MICROFACET, cs, float(np.float32(rough) * np.float32(rough)))]
    if node == "refraction_bsdf_node":  # refraction_bsdf_node.osl:30-38
        if dist == "sharp" or rough == 0.0:
            return "bsdf", [(LOBE_REFRACTION, cs, float(prm.get("IoR", 0.5)))]
        return "bsdf", [(LOBE_MICROFACET_REFRACT, cs, rough, float(prm.get("IoR", 0.5)))]  # microfacet(dist, N, 0, r, r, eta, 1)
    if node == "sheen_bsdf_node":  # sheen_bsdf_node.osl
        return "bsdf", [(LOBE_SHEEN, cs, rough)]
    if node == "transparent_bsdf_node":  # transparent_bsdf.node.osl
        return "bsdf", [(LOBE_TRANSPARENT, cs, 0.0)]
    if node == "diffuse_emitter_node":  # diffuse_emitter_node.osl:18
        return "emission", (cs, float(prm.get("power", 1.0)))
    if node == "background_node":  # background_node.osl
        return "background", (prm.get("Cs", (0.0, 0.0, 0.0)), float(prm.get("power", 1.0)))
    if node == "mix_closure_node":  # mix_closure_node.osl:20: A * (1 - fac) + B * fac
        fac = np.float32(prm.get("fac", 0.5))
        return "bsdf", _scale(inputs.get("A", []), np.float32(1.0) - fac) + _scale(inputs.get("B", []), fac)
    if node == "add_node":  # closure sum
        return "bsdf", list(inputs.get("A", [])) + list(inputs.get("B", []))
    if node == "material_node":  # the network's output layer: passes its surface closure through
        for slot in ("surface", "Cin", "A"):
            if slot in inputs:
                return "bsdf", list(inputs[slot])
        return "bsdf", []
    raise SceneError(f"shader node outside the built-in closure set: {node}")


def material_from_yaml(node: dict) -> Material:
    """convert<material_t*>::decode (src/codecs/scene/material.hpp:47-95) + a symbolic run of the network."""
    if not isinstance(node, dict) or "shaders" not in node:
        raise SceneError("a material is a map with a `shaders` list")
    layers, order = {}, []
    for sh in node["shaders"]:
        prm = {p["name"]: _param(p) for p in sh.get("parameters", []) or []}
        layers[sh["layer"]] = (sh["name"], prm)
        order.append(sh["layer"])
    feeds = {}  # to-layer -> {to-slot: from-layer}
    for e in node.get("connect", []) or []:
        feeds.setdefault(e["to"]["layer"], {})[e["to"]["slot"]] = e["from"]["layer"]
    done = {}

    def run(layer, seen=()):
        if layer in done:
            return done[layer]
        if layer in seen or layer not in layers:
            raise SceneError(f"bad shader network around layer {layer}")
        name, prm = layers[layer]
        inputs = {}
        for slot, src in feeds.get(layer, {}).items():
            kind, payload = run(src, seen + (layer,))
            if kind != "bsdf":
                raise SceneError("only BSDF closures can be mixed")
            inputs[slot] = payload
        done[layer] = _layer_closures(name, prm, inputs)
        return done[layer]

    kind, payload = run(order[-1])  # OSL: the last layer of a group is its output
    if kind == "emission":
        return Material(MAT_EMITTER, payload[0], power=payload[1])
    if kind == "background":
        return Material(MAT_BACKGROUND, payload[0], power=payload[1])
    if len(payload) > MAX_LOBES:
        raise SceneError("more than 8 closures in one material (bsdf_t::MaxLobes)")
    name, prm = layers[order[-1]]
    if len(order) == 1 and name in ("diffuse_bsdf_node", "glossy_bsdf_node") and not (name == "glossy_bsdf_node" and prm.get("distribution") == "sharp"):
        return Material(MAT_DIFFUSE if name == "diffuse_bsdf_node" else MAT_GLOSSY, prm.get("Cs", (1.0, 1.0, 1.0)),
                        roughness=float(prm.get("roughness", 0.0)))
    return Material(MAT_LAYERED, lobes=tuple(payload))


# ---- geometry ------------------------------------------------------------------------------------------------
def load_obj(path: str, material_ids: dict, default_material: int, smooth: bool = False) -> Mesh:
    """A Wavefront .obj as one mesh_t: `v`, optional `vn` (used when every corner carries one and `smooth`), `f`
    (fan-triangulated), `usemtl name` opens a face set with that scene material."""
    verts, normals, faces, face_n = [], [], [], []
    sets, cur = [], [default_material, []]
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "v":
                verts.append([float(x) for x in t[1:4]])
            elif t[0] == "vn":
                normals.append([float(x) for x in t[1:4]])
            elif t[0] == "usemtl":
                if t[1] not in material_ids:
                    raise SceneError(f"{path}: unknown material {t[1]}")
                if cur[1]:
                    sets.append(tuple(cur))
                cur = [material_ids[t[1]], []]
            elif t[0] == "f":
                idx, nidx = [], []
                for c in t[1:]:
                    parts = c.split("/")
                    i = int(parts[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                    if len(parts) > 2 and parts[2]:
                        j = int(parts[2])
                        nidx.append(j - 1 if j > 0 else len(normals) + j)
                for k in range(1, len(idx) - 1):
                    cur[1].append(len(faces))
                    faces.append([idx[0], idx[k], idx[k + 1]])
                    face_n.append([nidx[0], nidx[k], nidx[k + 1]] if len(nidx) == len(idx) else None)
    if cur[1]:
        sets.append(tuple(cur))
    if not faces:
        raise SceneError(f"{path}: no faces")
    v = np.asarray(verts, np.float32)
    f = np.asarray(faces, np.uint32)
    vn = None
    if normals and all(n is not None for n in face_n):  # per-vertex normals: the last corner normal seen wins
        vn = np.zeros_like(v)
        nn = np.asarray(normals, np.float32)
        for tri, ni in zip(faces, face_n):
            for a, b in zip(tri, ni):
                vn[a] = nn[b]
    return Mesh(v, f, [(m, np.asarray(fs, np.uint32)) for m, fs in sets], smooth=bool(smooth and vn is not None), normals=vn)


def _camera_from_yaml(node: dict, base: Camera) -> Camera:
    """convert<camera_t> (src/codecs/scene/entities.hpp): position / at / up; plus fov (degrees), film size and the
    thin-lens pair, which the reference takes from the Alembic camera."""
    cam = Camera(to_world=base.to_world, fov=base.fov, film_width=base.film_width, film_height=base.film_height,
                 focal_distance=base.focal_distance, aperture_radius=base.aperture_radius)
    if "position" in node:
        cam.to_world = Camera.look_at(node["position"], node.get("at", (0.0, 0.0, 0.0)), node.get("up", (0.0, 1.0, 0.0)))
    if "fov" in node:
        cam.fov = math.radians(float(node["fov"]))
    film = node.get("film", {})
    cam.film_width = int(film.get("width", cam.film_width))
    cam.film_height = int(film.get("height", cam.film_height))
    cam.focal_distance = float(node.get("focal-distance", cam.focal_distance))
    cam.aperture_radius = float(node.get("aperture-radius", cam.aperture_radius))
    return cam


_GENERATORS = {"cornell_box": scenes.cornell_box, "cornell_lobes": scenes.cornell_lobes, "sphere_field": scenes.sphere_field,
               "terrain": scenes.terrain, "heightfield": scenes.heightfield}


def load_scene(path: str) -> Scene:
    """codec::scene::import (src/codecs/scene.cpp:41-76)."""
    import yaml
    with open(path) as fh:
        cfg = yaml.safe_load(fh) or {}
    base = os.path.dirname(os.path.abspath(path))
    sc = Scene()
    names = {}
    for name, node in (cfg.get("materials") or {}).items():  # ids in file order, scene_t::add(name, material)
        names[name] = sc.add_material(material_from_yaml(node))
    for entry in cfg.get("data") or []:
        if "generator" in entry:  # procedural stand-in for an Alembic archive; brings its own materials
            gen = _GENERATORS.get(entry["generator"])
            if gen is None:
                raise SceneError(f"No importer for: {entry['generator']}")
            sub = gen(**(entry.get("args") or {}))
            off = len(sc.materials)
            sc.materials.extend(sub.materials)
            for m in sub.meshes:
                sc.add(Mesh(m.vertices, m.faces, [(mat + off, f) for mat, f in m.sets], smooth=m.smooth, normals=m.normals))
            sc.camera = sub.camera
            if sub.environment is not None:
                sc.environment = sub.environment + off
            continue
        p = os.path.join(base, entry["path"])
        if os.path.splitext(p)[1].lower() != ".obj":
            raise SceneError("No importer for: " + p)  # scene.cpp:28-30
        default = names.get(entry.get("material", ""), 0)
        sc.add(load_obj(p, names, default, smooth=bool(entry.get("smooth", False))))
    if not sc.meshes:
        raise SceneError("scene without geometry")
    if cfg.get("camera"):
        sc.camera = _camera_from_yaml(cfg["camera"], sc.camera)
    world = cfg.get("world") or {}
    if "environment" in world:  # import_world_data, scene.cpp:33-38
        env = world["environment"]
        if env not in names or sc.materials[names[env]].kind != MAT_BACKGROUND:
            raise SceneError(f"world.environment names no background material: {env}")
        sc.environment = names[env]
    for m in sc.meshes:
        for mat, _ in m.sets:
            if sc.materials[mat].kind == MAT_BACKGROUND:
                raise SceneError("a background material cannot be assigned to geometry")
    return sc


# ---- film -----------------------------------------------------------------------------------------------------
class FileFilm:
    """film::file_t (src/film/file.cpp): tiles land in an image that `finalize` writes to `path`
    (.pfm: RGB float32, bottom row first, little endian; .npy: the RGBA array)."""

    def __init__(self, width: int, height: int, path: str):
        self.path = path
        self.rgba = np.zeros((height, width, 4), np.float32)
        self.tiles_added = 0
        self._lock = threading.Lock()

    def add_tile(self, pos, size, buffer: np.ndarray) -> None:
        (x, y), (w, h) = pos, size
        with self._lock:
            self.rgba[y:y + h, x:x + w, :] = buffer
            self.tiles_added += 1

    def finalize(self) -> None:
        if self.path.lower().endswith(".npy"):
            np.save(self.path, self.rgba)
            return
        h, w = self.rgba.shape[:2]
        with open(self.path, "wb") as fh:
            fh.write(f"PF\n{w} {h}\n-1.0\n".encode())
            fh.write(np.ascontiguousarray(self.rgba[::-1, :, :3], dtype="<f4").tobytes())


def read_pfm(path: str) -> np.ndarray:
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"PF":
            raise SceneError("not an RGB .pfm")
        w, h = (int(x) for x in fh.readline().split())
        scale = float(fh.readline())
        data = np.frombuffer(fh.read(), "<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    return data[::-1].astype(np.float32)
