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


def test_deep_tree_instantiation_bit_exact(device, monkeypatch):
    """The kernel instantiation with the local-memory spill tier (trees deeper than the shared-memory stack; no
    re-grouped tree needs it) returns the same bits: forced through PHOS_TRACE_DEEP on the golden vectors."""
    monkeypatch.setenv("PHOS_TRACE_DEEP", "1")
    z = load_golden(CASES[-1])
    nodes, packets = z["nodes"], z["packets"]
    device.upload_accel(nodes, len(nodes) // 288, packets, len(packets) // 384)
    for batch in BATCHES:
        rays = golden_rays(z, batch)
        want = golden_out(z, batch, "linear", rays)
        assert len(mismatches(rays, device.trace(rays.copy()), want)) == 0, batch


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


def test_pinned_stream_sparse_write_back(device, oracle, monkeypatch):
    """Page-locked, device-mapped host arrays take the 32 B/ray up-link path: the surface record is stored
    straight into the caller's arrays only where the traversal wrote one.  Mixed closest-hit / shadow /
    masked rays with recognisable stale surface records, several pipeline chunks and a ragged tail; the
    copy-everything path (PHOS_E2E_SPARSE=0) must give the same arrays."""
    sc = scenes.heightfield(96)
    acc = Accel(sc)
    device.preprocess(sc, acc)
    nodes, packets = acc.nodes_array(), acc.packets_array()
    monkeypatch.setenv("PHOS_PIPE_CHUNK", "4096")
    n = 4096 * 5 + 1237
    src = raysets.aimed_rays(sc, n, seed=77)
    rnd = raysets.random_rays(sc, n, seed=78)  # many misses
    for f in ("px", "py", "pz", "wx", "wy", "wz"):
        getattr(src, f)[1::3] = getattr(rnd, f)[1::3]
    sh = raysets.as_shadow(src, seed=79, masked_fraction=0.2)
    for f in ("d", "flags"):
        getattr(src, f)[2::3] = getattr(sh, f)[2::3]
    src.mesh[:] = 0xABCD0001  # stale records: must survive on misses, shadow and masked rays
    src.face[:] = 777
    src.u[:] = 0.25
    src.v[:] = 0.5
    want, _ = oracle.traverse(nodes, packets, src)
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PHOS_E2E_SPARSE", mode)
        pr = pinned_ray_batch(n)
        for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
            getattr(pr, f)[:] = getattr(src, f)
        device.trace(pr)
        assert len(mismatches(src, pr, want)) == 0, mode
        got[mode] = pr
    closest = (src.flags & 4) == 0
    for f in ("d", "u", "v", "mesh", "face", "flags"):
        a, b = bits(getattr(got["1"], f)), bits(getattr(got["0"], f))
        assert np.array_equal(a[closest], b[closest]), f


def test_long_pinned_stream_equals_the_device_resident_query(device, monkeypatch):
    """A long page-locked stream through the host-pointer pipeline (several chunks in flight, slots reused, a ragged last
    chunk): every ray exactly once, same bits as the device-resident query."""
    sc = scenes.heightfield(96)
    device.preprocess(sc)
    monkeypatch.setenv("PHOS_PIPE_CHUNK", "32768")
    n = 9 * 32768 + 12345
    src = raysets.aimed_rays(sc, n, seed=91)
    sh = raysets.as_shadow(src, seed=92, masked_fraction=0.1)
    for f in ("d", "flags"):
        getattr(src, f)[1::4] = getattr(sh, f)[1::4]
    dr = device.device_rays(n)
    dr.upload(src)
    device.trace_device(dr)
    want = dr.download()
    dr.free()
    closest = (src.flags & 4) == 0
    pr = pinned_ray_batch(n)
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
        getattr(pr, f)[:] = getattr(src, f)
    device.trace(pr)
    assert np.array_equal(pr.flags, want.flags)
    for f in ("d", "u", "v", "mesh", "face"):
        assert np.array_equal(bits(getattr(pr, f))[closest], bits(getattr(want, f))[closest]), f
    pr.free()


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


def test_config3_full_size_properties(device, oracle):
    """BASELINE config 3 at full size: 9 999 394-triangle displaced terrain (+ sky quad), the bounce-ray
    and shadow-ray streams of a 1920 x 1080 frame.  A 20 000-ray sample of each stream is compared
    ray for ray with the oracle traversal of the same 10 M-triangle tree; the whole streams are checked
    through size-independent properties."""
    from phosphorus_mk2_b200.device import make_tiles
    sc = scenes.terrain()
    assert sc.num_triangles() == 9_999_394
    acc = Accel(sc)
    device.preprocess(sc, acc)
    device.upload_scene(sc)
    st = device.accel_stats()
    assert st.triangles == 9_999_394 and st.max_depth + 2 <= 12  # the re-grouped tree fits the shared-memory stack
    nodes, packets = acc.nodes_array(), acc.packets_array()
    tiles = make_tiles(1920, 1080)
    rng = np.random.default_rng(3)
    F = ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags")
    for which in ("bounce", "shadow"):
        dr = device.device_rays(1920 * 1080)
        n = device.wavefront_rays(tiles, dr, which, 0, 1, 42)
        rays = dr.download().slice(0, n)
        assert n > 1_500_000
        device.trace_device_n(dr, n)
        got = dr.download().slice(0, n)
        dr.free()
        idx = np.sort(rng.choice(n, 20000, replace=False))
        sample, sub = RayBatch(len(idx)), RayBatch(len(idx))
        for f in F:
            getattr(sample, f)[:] = getattr(rays, f)[idx]
            getattr(sub, f)[:] = getattr(got, f)[idx]
        want, _ = oracle.traverse(nodes, packets, sample)
        assert len(mismatches(sample, sub, want)) == 0, which
        if which == "bounce":
            hit = got.hit
            assert 0.2 < hit.mean() < 0.9
            # idempotence: with tmax = the reported distance nothing closer exists
            again = rays.copy()
            again.d[:] = got.d
            again = device.trace(again)
            assert np.all(again.d[hit] == got.d[hit]) and not np.any(again.hit[~hit])
            assert np.all(got.u[hit] >= 0) and np.all(got.v[hit] >= 0) and np.all(got.u[hit] + got.v[hit] <= 1 + 1e-6)
        else:
            masked = (rays.flags & MASKED) != 0
            assert np.array_equal(got.mesh, rays.mesh) and np.array_equal(got.face, rays.face)  # light ids untouched
            assert not np.any(got.hit[masked])                                                   # never traced
            assert np.all(got.d[got.hit] < rays.d[got.hit])


def test_ragged_and_degenerate_streams(device, oracle):
    """Edge cases: stream lengths around the 32-ray chunk and the 128-thread CTA, a misaligned device
    view (no TMA), all-masked and all-miss streams, zero-length and NaN-free degenerate directions."""
    sc = scenes.heightfield(40)
    acc = Accel(sc)
    device.preprocess(sc, acc)
    nodes, packets = acc.nodes_array(), acc.packets_array()
    base = raysets.aimed_rays(sc, 700, seed=77)
    for n in (2, 32, 63, 64, 65, 127, 129, 511, 700):
        rays = base.slice(0, n)
        want, _ = oracle.traverse(nodes, packets, rays)
        assert len(mismatches(rays, device.trace(rays.copy()), want)) == 0, n
    # misaligned host views (odd element offsets): the host path stages them, results identical
    off = RayBatch(699)
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
        setattr(off, f, getattr(base, f)[1:])
    want, _ = oracle.traverse(nodes, packets, base.slice(1, 700))
    cp = RayBatch(699)
    for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags"):
        getattr(cp, f)[:] = getattr(off, f)
    assert len(mismatches(base.slice(1, 700), device.trace(cp), want)) == 0
    # all masked: nothing is touched
    m = base.copy()
    m.flags[:] = MASKED | SHADOW
    out = device.trace(m.copy())
    for f in ("d", "u", "v", "mesh", "face", "flags"):
        assert np.array_equal(bits(getattr(out, f)), bits(getattr(m, f)))
    # all miss: rays leaving the scene
    away = base.copy()
    away.wy[:] = np.abs(away.wy) + 1.0
    away.py[:] = 100.0
    out = device.trace(away.copy())
    assert not out.hit.any() and np.array_equal(bits(out.d), bits(away.d))
    # zero direction components and tmax = 0
    z = base.slice(0, 64)
    z.wx[:32] = 0.0
    z.wz[32:] = -0.0
    z.d[::7] = 0.0
    want, _ = oracle.traverse(nodes, packets, z)
    assert len(mismatches(z, device.trace(z.copy()), want)) == 0
