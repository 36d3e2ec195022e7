"""The library's host BVH build (phos_bvh_build) against the reference builder's own arrays.
CPU only: calls the host-side entry points of libphos_cuda.so, no GPU needed."""
import numpy as np
import pytest

from parity import CASES, load_golden
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel

SCENES = {"cornell": lambda: scenes.cornell_box(64, 64), "heightfield24": lambda: scenes.heightfield(24),
          "spheres2": lambda: scenes.sphere_field(2, 12, 6, 64, 64)}


def assert_same_tree(ref_nodes, ref_packets, nodes, packets):
    assert len(ref_nodes) == len(nodes) and len(ref_packets) == len(packets)
    # node_t<8>: bounds (192 B) + offset (32 B) + num (8 B) + flags (32 B); the trailing pad is unspecified
    assert np.array_equal(ref_nodes.reshape(-1, 288)[:, :264], nodes.reshape(-1, 288)[:, :264])
    rp, hp = ref_packets.reshape(-1, 384), packets.reshape(-1, 384)
    num = rp[:, 288:292].copy().view(np.uint32).ravel()
    assert np.array_equal(num, hp[:, 288:292].copy().view(np.uint32).ravel())
    # lanes >= num are uninitialised in the reference (triangle.hpp:43): compare live lanes only
    live9 = np.broadcast_to(np.arange(8)[None, None, :] < num[:, None, None], (len(rp), 9, 8))
    a = rp[:, :288].copy().view(np.uint32).reshape(-1, 9, 8)
    b = hp[:, :288].copy().view(np.uint32).reshape(-1, 9, 8)
    assert np.array_equal(a[live9], b[live9])
    live2 = live9[:, :2, :]
    a = rp[:, 292:356].copy().view(np.uint32).reshape(-1, 2, 8)
    b = hp[:, 292:356].copy().view(np.uint32).reshape(-1, 2, 8)
    assert np.array_equal(a[live2], b[live2])


@pytest.mark.parametrize("case", CASES)
def test_host_build_is_bit_identical_to_golden_reference_tree(case):
    z = load_golden(case)
    a = Accel(SCENES[case]())
    assert_same_tree(z["nodes"], z["packets"], a.nodes_array(), a.packets_array())


@pytest.mark.parametrize("make", [lambda: scenes.heightfield(150, seed=3), lambda: scenes.sphere_field(5, 20, 10, 64, 64),
                                  lambda: scenes.terrain(n=120, glossy_fraction=0.1)])
def test_host_build_is_bit_identical_to_live_reference(reflib, make):
    sc = make()
    rs = reflib.scene(sc)
    rs.build()
    rn, rp = rs.accel()
    a = Accel(sc)
    assert_same_tree(rn, rp, a.nodes_array(), a.packets_array())


def test_parallel_build_equals_single_threaded_build(monkeypatch):
    """phos_bvh_build expands the top of the tree into a skeleton, builds every range below
    PHOS_BUILD_GRAIN primitives as an independent task and lays the pieces out with prefix offsets: the
    arrays must be those of the plain depth-first recursion for any grain / thread count."""
    sc = scenes.heightfield(160, seed=5)  # 51 k triangles
    monkeypatch.setenv("PHOS_BUILD_GRAIN", "2000000000")
    monkeypatch.setenv("PHOS_THREADS", "1")
    a = Accel(sc)
    n0, p0 = a.nodes_array(), a.packets_array()
    for grain, threads in (("8", "4"), ("100", "3"), ("5000", "8")):
        monkeypatch.setenv("PHOS_BUILD_GRAIN", grain)
        monkeypatch.setenv("PHOS_THREADS", threads)
        b = Accel(sc)
        assert_same_tree(n0, p0, b.nodes_array(), b.packets_array())
        assert np.array_equal(p0, b.packets_array())


def test_fewer_than_eight_triangles_builds_no_node():
    """The reference builder returns 0 at the root for < 8 primitives (binned_sah_builder.hpp:220):
    no node at all.  The host build mirrors that; upload then rejects the empty structure."""
    from phosphorus_mk2_b200.scene import MAT_DIFFUSE, Material, Mesh, Scene
    s = Scene()
    m = s.add_material(Material(MAT_DIFFUSE))
    s.add(Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32), np.array([[0, 1, 2], [1, 3, 2]]),
               [(m, np.arange(2))]))
    a = Accel(s)
    assert a.num_nodes == 0 and a.num_packets == 0


def test_host_build_refuses_a_face_outside_its_mesh():
    """phos_bvh_build validates vertex indices before it reads vertices through them."""
    import pytest
    from phosphorus_mk2_b200.lib import PhosError
    sc = scenes.cornell_box(32, 32)
    sc.meshes[0].faces[0, 2] = 4_000_000
    with pytest.raises(PhosError):
        Accel(sc)
