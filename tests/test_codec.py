"""Scene ingest (phosphorus_mk2_b200/codec.py): the reference's YAML scene layout (src/codecs/scene.cpp) with
.obj geometry, shader networks evaluated to closure lists, the file film.  CPU only, except the CLI test."""
import os

import numpy as np
import pytest

from phosphorus_mk2_b200 import codec
from phosphorus_mk2_b200.scene import (LOBE_DIFFUSE, LOBE_MICROFACET, LOBE_REFLECTION, MAT_BACKGROUND, MAT_DIFFUSE, MAT_EMITTER,
                                       MAT_LAYERED)

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def test_yaml_scene_is_imported_like_the_reference_codec():
    sc = codec.load_scene(os.path.join(DATA, "scene.yaml"))
    names = ["floor", "body", "lid", "lamp", "sky"]  # scene_t::add(name, material): ids in file order
    kinds = [m.kind for m in sc.materials]
    assert kinds == [MAT_DIFFUSE, MAT_LAYERED, MAT_LAYERED, MAT_EMITTER, MAT_BACKGROUND]
    assert sc.materials[0].roughness == 20.0  # -> oren_nayar(N, 20)
    body = sc.materials[1].lobes                # mix_closure_node: A * (1 - fac) + B * fac
    assert [l[0] for l in body] == [LOBE_DIFFUSE, LOBE_MICROFACET]
    assert np.allclose(body[0][1], np.float32(0.75) * np.float32([0.7, 0.2, 0.1]))
    assert np.allclose(body[1][1], np.float32(0.25) * np.float32([0.9, 0.9, 0.9])) and np.isclose(body[1][2], 0.09)
    assert sc.materials[2].lobes == ((LOBE_REFLECTION, (0.95, 0.95, 0.95), 0.0),)  # distribution "sharp"
    assert sc.environment == names.index("sky")
    assert sc.num_meshes() == 2 and sc.num_triangles() == 4 + 12
    floor, box = sc.meshes
    assert [m for m, _ in floor.sets] == [0, 3] and [m for m, _ in box.sets] == [1, 2]
    assert [len(f) for _, f in box.sets] == [8, 4]
    cam = sc.camera
    assert (cam.film_width, cam.film_height) == (96, 64) and np.isclose(cam.fov, np.radians(60))
    assert np.allclose(cam.to_world[3, :3], [2.5, 2.0, 4.0])
    d = sc.desc()  # flattens for the C ABI
    assert d.environment == 4 and d.materials[1].num_lobes == 2


def test_refraction_node_sharp_and_rough():
    from phosphorus_mk2_b200.scene import LOBE_MICROFACET_REFRACT, LOBE_REFRACTION
    def mat(params):
        return codec.material_from_yaml({"shaders": [{"name": "refraction_bsdf_node", "layer": "l0", "parameters": params}]})
    ior = {"name": "IoR", "type": "float", "value": 1.5}
    assert mat([ior]).lobes == ((LOBE_REFRACTION, (1.0, 1.0, 1.0), 1.5),)
    rough = mat([ior, {"name": "roughness", "type": "float", "value": 0.2}])
    assert rough.lobes == ((LOBE_MICROFACET_REFRACT, (1.0, 1.0, 1.0), 0.2, 1.5),)  # roughness, not its square (osl:31)
    sharp = mat([ior, {"name": "roughness", "type": "float", "value": 0.2}, {"name": "distribution", "type": "string", "value": "sharp"}])
    assert sharp.lobes[0][0] == LOBE_REFRACTION


def test_unknown_nodes_and_importers_are_rejected(tmp_path):
    bad = tmp_path / "bad.yaml"
    bad.write_text("materials:\n  m:\n    shaders:\n      - {name: principled_bsdf_node, layer: l0}\ndata:\n  - {generator: cornell_box}\n")
    with pytest.raises(codec.SceneError):
        codec.load_scene(str(bad))
    abc = tmp_path / "abc.yaml"
    abc.write_text("data:\n  - {path: scene.abc}\n")
    with pytest.raises(codec.SceneError, match="No importer"):
        codec.load_scene(str(abc))


def test_generator_entries_and_file_film(tmp_path):
    y = tmp_path / "gen.yaml"
    y.write_text("data:\n  - {generator: cornell_box, args: {width: 32, height: 32}}\ncamera:\n  film: {width: 40, height: 24}\n")
    sc = codec.load_scene(str(y))
    assert sc.num_triangles() == 38 and sc.camera.film_width == 40
    film = codec.FileFilm(40, 24, str(tmp_path / "out.pfm"))
    img = np.random.default_rng(1).random((24, 40, 4), dtype=np.float32)
    for (x, yy, w, h) in ((0, 0, 32, 24), (32, 0, 8, 24)):
        film.add_tile((x, yy), (w, h), img[yy:yy + h, x:x + w])
    film.finalize()
    assert np.array_equal(codec.read_pfm(str(tmp_path / "out.pfm")), img[..., :3])


@pytest.mark.gpu
def test_command_line_renders_the_yaml_scene(tmp_path, oracle):
    """python -m phosphorus_mk2_b200 -s 8 -p 1 -d 4 -o out.pfm scene.yaml == the oracle on the imported scene."""
    from phosphorus_mk2_b200.__main__ import main
    from phosphorus_mk2_b200.device import Accel
    out, nrm = str(tmp_path / "out.pfm"), str(tmp_path / "n.npy")
    assert main(["-s", "8", "-p", "1", "-d", "4", "-o", out, "-n", nrm, os.path.join(DATA, "scene.yaml")]) == 0
    got = codec.read_pfm(out)
    sc = codec.load_scene(os.path.join(DATA, "scene.yaml"))
    acc = Accel(sc)
    wn = np.zeros((64, 96, 3), np.float32)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 8, 1, 4, seed=0, normals=wn)
    assert np.abs(got - want[..., :3]).mean() / np.abs(want[..., :3]).mean() < 1e-3
    assert np.abs(np.load(nrm) - wn).max() < 1e-5
