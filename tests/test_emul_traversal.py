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


def test_one_ulp_reciprocals_do_not_change_any_result(emul, emul_rcp, oracle):
    """The kernel takes its reciprocal directions from MUFU.RCP (1 ulp); they only feed the conservative box test.
    With every reciprocal pushed one ulp off (alternately up / down) the emulated traversal must return the same bits
    as with exact reciprocals, on the golden vectors and on a larger scene, and they must be the oracle's."""
    for case in CASES:
        z = load_golden(case)
        for batch in BATCHES:
            rays = golden_rays(z, batch)
            want = golden_out(z, batch, "linear", rays)
            got, _, _, _ = run_emul(emul_rcp, z["nodes"], z["packets"], rays)
            assert len(mismatches(rays, got, want)) == 0, (case, batch)
    sc = scenes.heightfield(200, seed=9)
    a = Accel(sc)
    nodes, packets = a.nodes_array(), a.packets_array()
    for rays in (raysets.aimed_rays(sc, 40000, seed=61), raysets.random_rays(sc, 40000, seed=62),
                 raysets.as_shadow(raysets.aimed_rays(sc, 40000, seed=63), seed=64)):
        want, _ = oracle.traverse(nodes, packets, rays)
        exact, _, _, _ = run_emul(emul, nodes, packets, rays)
        pert, _, _, _ = run_emul(emul_rcp, nodes, packets, rays)
        assert len(mismatches(rays, pert, want)) == 0
        closest = (rays.flags & 4) == 0  # an occluded shadow ray's distance depends on the visiting order; its verdict does not
        for f in ("d", "u", "v", "mesh", "face"):
            assert np.array_equal(getattr(exact, f)[closest].view(np.uint32), getattr(pert, f)[closest].view(np.uint32)), f
        assert np.array_equal(exact.flags, pert.flags)


def _digest(emul, nodes, packets):
    d = C.c_uint64()
    stats = (C.c_uint32 * 4)()
    assert emul.emul_repack_digest(nodes.ctypes.data, len(nodes) // 288, packets.ctypes.data, len(packets) // 384, C.byref(d), stats) == 0
    return d.value, list(stats)


def test_repack_is_independent_of_the_thread_count(emul, monkeypatch):
    """The re-pack runs a level's nodes on all host threads and bins big ranges in parallel chunks; the
    packed arrays must not depend on that (min / max / integer sums only)."""
    sc = scenes.heightfield(300, seed=3)  # 180 k triangles: several chunks per split at the top
    a = Accel(sc)
    nodes, packets = a.nodes_array(), a.packets_array()
    got = {}
    for t in ("1", "3", "8"):
        monkeypatch.setenv("PHOS_THREADS", t)
        got[t] = _digest(emul, nodes, packets)
    assert got["1"] == got["3"] == got["8"]
    assert got["1"][1][1] == sc.num_triangles() and got["1"][1][3] <= 2


@pytest.mark.parametrize("leaf", ["1", "2", "3", "15"])
def test_leaf_size_knob_keeps_results(emul, oracle, monkeypatch, leaf):
    """PHOS_REPACK_LEAF only changes how the reference leaves are cut; hits are the oracle's for any value."""
    monkeypatch.setenv("PHOS_REPACK_LEAF", leaf)
    sc = scenes.heightfield(120, seed=11)
    a = Accel(sc)
    nodes, packets = a.nodes_array(), a.packets_array()
    for rays in (raysets.aimed_rays(sc, 8000, seed=41), raysets.as_shadow(raysets.aimed_rays(sc, 8000, seed=43), seed=44)):
        want, _ = oracle.traverse(nodes, packets, rays)
        got, _, _, stats = run_emul(emul, nodes, packets, rays)
        assert len(mismatches(rays, got, want)) == 0
        assert stats[3] <= int(leaf) and stats[1] == sc.num_triangles()


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


def test_warp_level_schedule_of_the_kernel_gives_the_one_ray_results(oracle):
    """tools/warp_sim runs trace_kernel's scheduling on the host — 32 lanes, refill at >= 6 idle lanes, one node step or one
    triangle step per iteration by the weighted vote, the next node committed at the end of a node step — over the device's
    own node_test / mt_triangle / accept_hit.  Closest hits and shadow verdicts must be those of the oracle for every ray,
    with the shipped policy and with leaves deferred behind node steps (a different ORDER of the same tests)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    here = os.path.join(root, "tools", "warp_sim")
    so = os.path.join(here, "libsim_test.so")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include",
                    os.path.join(here, "sim.cpp"), os.path.join(root, "phosphorus_mk2_b200", "csrc", "repack.cpp"), "-o", so], check=True)
    L = C.CDLL(so)

    class Params(C.Structure):
        _fields_ = [("refill_min", C.c_int), ("tri_bias", C.c_int), ("defer", C.c_int), ("defer_shadow_only", C.c_int),
                    ("tri_min_lanes", C.c_int), ("hint", C.c_int)]

    sc = scenes.heightfield(120, seed=4)
    a = Accel(sc)
    nodes, packets = a.nodes_array(), a.packets_array()
    rays = raysets.aimed_rays(sc, 6000, seed=12)
    sh = raysets.as_shadow(raysets.aimed_rays(sc, 6000, seed=13), seed=14, masked_fraction=0.1)
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "flags"):
        getattr(rays, f)[1::3] = getattr(sh, f)[1::3]
    want, _ = oracle.traverse(nodes, packets, rays)
    closest = (rays.flags & 6) == 0
    for kw in (dict(refill_min=6, tri_bias=3, defer=0), dict(refill_min=6, tri_bias=2, defer=2)):
        s = rays.as_struct()
        out = (C.c_uint64 * 10)()
        d = np.zeros(rays.n, np.float32)
        fl = np.zeros(rays.n, np.uint32)
        p = Params(**kw)
        assert L.warp_sim(C.c_void_p(nodes.ctypes.data), C.c_uint32(len(nodes) // 288), C.c_void_p(packets.ctypes.data),
                          C.c_uint32(len(packets) // 384), C.byref(s), C.c_uint64(rays.n), C.c_int(7), C.byref(p), out,
                          C.c_void_p(d.ctypes.data), C.c_void_p(fl.ctypes.data)) == 0
        assert np.array_equal(fl, want.flags), kw
        assert np.array_equal(d[closest].view(np.uint32), want.d[closest].view(np.uint32)), kw
        assert out[8] == int(((rays.flags & 2) == 0).sum())  # every unmasked ray traced exactly once
