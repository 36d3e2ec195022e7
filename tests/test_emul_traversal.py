"""The device traversal source (csrc/trace_ray.cuh) + the product's re-pack (csrc/repack.cpp),
compiled for the host by tests/emul, against the golden vectors and the oracle.  CPU only; the GPU
parity tests (test_gpu_trace.py) run the same checks through the C ABI on the B200."""
import ctypes as C

import numpy as np
import pytest

from parity import BATCHES, CASES, golden_out, golden_rays, load_golden, mismatches
from phosphorus_mk2_b200 import raysets, scenes
from phosphorus_mk2_b200.device import Accel


def run_emul(emul, nodes, packets, rays):
    out = rays.copy()
    s = out.as_struct()
    nn, nt = C.c_uint64(), C.c_uint64()
    stats = (C.c_uint32 * 4)()
    err = C.create_string_buffer(256)
    rc = emul.emul_trace(nodes.ctypes.data, len(nodes) // 288, packets.ctypes.data, len(packets) // 384, C.byref(s), out.n,
                         C.byref(nn), C.byref(nt), stats, err)
    assert rc == 0, err.value
    return out, nn.value, nt.value, list(stats)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("batch", BATCHES)
def test_emulated_device_traversal_matches_reference_brute_force(emul, case, batch):
    z = load_golden(case)
    rays = golden_rays(z, batch)
    want = golden_out(z, batch, "linear", rays)
    got, _, _, stats = run_emul(emul, z["nodes"], z["packets"], rays)
    assert stats[1] == int(z["packets"].reshape(-1, 384)[:, 288:292].copy().view(np.uint32).sum())  # every triangle kept
    assert len(mismatches(rays, got, want)) == 0


def test_emulated_traversal_larger_scene_against_oracle(emul, oracle):
    sc = scenes.heightfield(200, seed=9)  # 80 k triangles
    a = Accel(sc)
    nodes, packets = a.nodes_array(), a.packets_array()
    for rays in (raysets.aimed_rays(sc, 20000, seed=31), raysets.random_rays(sc, 20000, seed=32),
                 raysets.as_shadow(raysets.aimed_rays(sc, 20000, seed=33), seed=34)):
        want, _ = oracle.traverse(nodes, packets, rays)
        got, nn, nt, stats = run_emul(emul, nodes, packets, rays)
        assert len(mismatches(rays, got, want)) == 0
        assert stats[2] + 2 <= 24  # fits the shared-memory stack


def test_repack_splits_oversized_leaves(emul, oracle):
    """A pile of coincident-centroid triangles makes the SAH builder emit one big leaf (> 15
    triangles); the re-pack must split it into <= 15-triangle leaves without losing a triangle."""
    from phosphorus_mk2_b200.scene import MAT_DIFFUSE, Material, Mesh, Scene
    rng = np.random.default_rng(4)
    s = Scene()
    m = s.add_material(Material(MAT_DIFFUSE))
    n = 60
    ang = rng.random(n) * 2 * np.pi
    # all triangles share the same bounding box centre -> no useful split
    verts, faces = [], []
    for k in range(n):
        c, sn = np.cos(ang[k]), np.sin(ang[k])
        z = 0.01 * k
        verts += [[-c, -sn, z], [c, sn, z], [-sn * 0.0, 0.0, z]]
        verts[-1] = [sn, -c, z]
        verts += [[-sn, c, z]]
        i = 4 * k
        faces += [[i, i + 1, i + 2], [i, i + 3, i + 1]]
    s.add(Mesh(np.array(verts, np.float32), np.array(faces), [(m, np.arange(2 * n))]))
    # plus a far cluster so the root splits at all
    v2 = (rng.random((30, 3)) + np.array([10, 10, 0])).astype(np.float32)
    s.add(Mesh(v2, np.arange(30).reshape(10, 3), [(m, np.arange(10))]))
    a = Accel(s)
    if a.num_nodes == 0:
        pytest.skip("builder produced no node")
    nodes, packets = a.nodes_array(), a.packets_array()
    rays = raysets.aimed_rays(s, 4000, seed=5)
    want = oracle.brute_force(packets, rays)
    got, _, _, stats = run_emul(emul, nodes, packets, rays)
    assert stats[3] <= 15
    assert len(mismatches(rays, got, want)) == 0
