"""Parity of the CUDA ray-query path, called through the C ABI (libphos_cuda.so), against the
golden vectors of the reference's brute-force kernel and against the oracle.  Needs a B200."""
import numpy as np
import pytest

from parity import BATCHES, CASES, bits, classify_vs_stream, golden_out, golden_rays, load_golden, mismatches
from phosphorus_mk2_b200 import raysets, scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, pinned_ray_batch
from phosphorus_mk2_b200.lib import PhosError
from phosphorus_mk2_b200.rays import HIT, MASKED, SHADOW, RayBatch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_golden_vectors_bit_exact(device, case):
    z = load_golden(case)
    nodes, packets = z["nodes"], z["packets"]
    device.upload_accel(nodes, len(nodes) // 288, packets, len(packets) // 384)
    st = device.accel_stats()
    assert st.ref_nodes == len(nodes) // 288 and st.triangles > 0
    for batch in BATCHES:
        rays = golden_rays(z, batch)
        want = golden_out(z, batch, "linear", rays)
        got = device.trace(rays.copy())
        assert len(mismatches(rays, got, want)) == 0, (case, batch)


def test_device_resident_path_equals_host_pointer_path(device):
    sc = scenes.heightfield(64)
    device.preprocess(sc)
    rays = raysets.aimed_rays(sc, 50000, seed=41)
    a = device.trace(rays.copy())
    dr = device.device_rays(rays.n)
    dr.upload(rays)
    device.trace_device(dr)
    b = dr.download()
    for f in ("d", "u", "v", "mesh", "face", "flags"):
        assert np.array_equal(bits(getattr(a, f)), bits(getattr(b, f))), f
    nodes, tris = device.trace_count(dr)
    assert nodes > 0 and tris > 0
    dr.free()


def test_pinned_host_stream_and_odd_lengths(device, oracle):
    sc = scenes.heightfield(48)
    acc = Accel(sc)
    device.preprocess(sc, acc)
    nodes, packets = acc.nodes_array(), acc.packets_array()
    for n in (1, 31, 33, 1000, 4097):
        src = raysets.aimed_rays(sc, n, seed=50 + n)
        want, _ = oracle.traverse(nodes, packets, src)
        pr = pinned_ray_batch(n)
        for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
            getattr(pr, f)[:] = getattr(src, f)
        device.trace(pr)
        assert len(mismatches(src, pr, want)) == 0, n


def test_empty_stream_and_call_order_errors(device):
    fresh = CudaDevice.make(Options(), 0)
    with pytest.raises(PhosError):
        fresh.trace(RayBatch(8))  # no acceleration structure yet
    with pytest.raises(PhosError):
        fresh.upload_accel(np.zeros(0, np.uint8), 0, np.zeros(0, np.uint8), 0)  # empty tree is rejected
    fresh.close()
    sc = scenes.heightfield(16)
    device.preprocess(sc)
    device.trace(RayBatch(0))  # zero rays is a no-op


@pytest.mark.parametrize("make,nrays", [(lambda: scenes.heightfield(280, seed=9), 200000),
                                        (lambda: scenes.sphere_field(6, 32, 16, 64, 64), 200000)])
def test_mid_size_scenes_against_oracle(device, oracle, make, nrays):
    """~150 k triangles, 200 k rays per batch: closest-hit (incoherent, all-hit and random) and
    mixed shadow / masked streams, every ray compared with the oracle traversal (== brute force)."""
    sc = make()
    acc = Accel(sc)
    device.preprocess(sc, acc)
    nodes, packets = acc.nodes_array(), acc.packets_array()
    for rays in (raysets.aimed_rays(sc, nrays, seed=61), raysets.random_rays(sc, nrays, seed=62),
                 raysets.as_shadow(raysets.aimed_rays(sc, nrays, seed=63), seed=64)):
        want, _ = oracle.traverse(nodes, packets, rays)
        got = device.trace(rays.copy())
        bad = mismatches(rays, got, want)
        assert len(bad) == 0, bad[:10]


def test_config2_full_size_properties(device, oracle):
    """BASELINE config 2 at full size: 1 048 576-triangle sphere field, 1920 x 1080 primary rays.
    The oracle checks a 20 000-ray sample ray for ray; the whole frame is checked through
    size-independent properties: idempotence (re-tracing a hit ray with tmax = its hit distance
    finds nothing closer: the result is the closest), the hit point lies on the reported triangle,
    and a shadow query along every primary ray is occluded exactly when the primary ray hit."""
    sc = scenes.sphere_field()
    acc = Accel(sc)
    device.preprocess(sc, acc)
    cam = sc.camera
    rays = oracle.camera_rays(cam)  # the oracle only generates the input here
    assert rays.n == 1920 * 1080
    got = device.trace(rays.copy())
    hit = got.hit
    assert 0.2 < hit.mean() < 1.0
    # (1) oracle on a sample
    nodes, packets = acc.nodes_array(), acc.packets_array()
    idx = np.random.default_rng(1).choice(rays.n, 20000, replace=False)
    sample = RayBatch(len(idx))
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
        getattr(sample, f)[:] = getattr(rays, f)[idx]
    want, _ = oracle.traverse(nodes, packets, sample)
    sub = RayBatch(len(idx))
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
        getattr(sub, f)[:] = getattr(got, f)[idx]
    assert len(mismatches(sample, sub, want)) == 0
    # (2) idempotence: tmax = found distance -> strict '<' finds nothing (or an exact tie, never closer)
    again = rays.copy()
    again.d[:] = got.d
    again = device.trace(again)
    assert np.all(again.d[hit] == got.d[hit])
    assert not np.any(again.hit[~hit])
    # (3) the hit point is on the reported triangle: barycentrics in range, point matches p + t w
    mesh_id = got.mesh[hit] & 0xFFFF
    face = got.face[hit] // 3
    u, v, t = got.u[hit].astype(np.float64), got.v[hit].astype(np.float64), got.d[hit].astype(np.float64)
    assert np.all(u >= 0) and np.all(v >= 0) and np.all(u + v <= 1 + 1e-6)
    V = np.stack([m.vertices for m in sc.meshes])  # all spheres share topology
    F = sc.meshes[0].faces
    tri = F[face]
    a, b, c = V[mesh_id, tri[:, 0]], V[mesh_id, tri[:, 1]], V[mesh_id, tri[:, 2]]
    on_tri = (1 - u - v)[:, None] * a + u[:, None] * b + v[:, None] * c
    o = np.stack([rays.px, rays.py, rays.pz], 1)[hit].astype(np.float64)
    w = np.stack([rays.wx, rays.wy, rays.wz], 1)[hit].astype(np.float64)
    assert np.abs(on_tri - (o + t[:, None] * w)).max() < 1e-3
    # (4) occlusion query agrees with the closest-hit verdict
    sh = rays.copy()
    sh.flags[:] = SHADOW
    sh.mesh[:] = 12345
    sh = device.trace(sh)
    assert np.array_equal(sh.hit, hit)
    assert np.all(sh.mesh == 12345)


def test_config3_ray_sets_against_oracle(device, oracle):
    """BASELINE config 3's ray sets at reduced size (318 k-triangle terrain, 256 x 256): the incoherent
    BSDF-sampled bounce rays and the next-event shadow rays produced by the pipeline itself, traced
    through the C ABI and compared ray for ray with the oracle."""
    from phosphorus_mk2_b200.device import make_tiles
    sc = scenes.terrain(n=400, width=256, height=256)
    acc = Accel(sc)
    device.preprocess(sc, acc)
    device.upload_scene(sc)
    nodes, packets = acc.nodes_array(), acc.packets_array()
    tiles = make_tiles(256, 256)
    for which in ("bounce", "shadow"):
        dr = device.device_rays(256 * 256)
        n = device.wavefront_rays(tiles, dr, which, 0, 1, 42)
        assert n > 1000
        rays = dr.download().slice(0, n)
        dr.free()
        if which == "shadow":
            assert np.all((rays.flags & SHADOW) != 0)
        want, _ = oracle.traverse(nodes, packets, rays)
        got = device.trace(rays.copy())
        assert len(mismatches(rays, got, want)) == 0
