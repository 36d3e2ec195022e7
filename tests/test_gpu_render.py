"""The CUDA wavefront path tracer, through the C ABI, against the integrator oracle driven by the
same counter-based random numbers.  Needs a B200."""
import numpy as np
import pytest

from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles
from phosphorus_mk2_b200.lib import PhosError
from phosphorus_mk2_b200.scene import MAT_DIFFUSE, MAT_GLOSSY

pytestmark = pytest.mark.gpu


def mean_rel_err(a, b):
    return float(np.abs(a[..., :3] - b[..., :3]).mean() / np.abs(b[..., :3]).mean())


def render_gpu(sc, acc, spp, depth, seed, pps=1, tiles=None, ranges=None, device=0):
    dev = CudaDevice.make(Options(samples_per_pixel=spp, paths_per_sample=pps, path_depth=depth), device)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    cam = sc.camera
    tiles = tiles if tiles is not None else make_tiles(cam.film_width, cam.film_height)
    for (a, b) in (ranges or [(0, spp)]):
        dev.render(tiles, a, b, spp, seed)
    img = dev.film_read()
    n = dev.launch_count()
    dev.close()
    return img, n


@pytest.mark.parametrize("kind,depth,spp", [("mixed", 1, 4), ("mixed", 6, 16), ("diffuse", 9, 4), ("glossy", 4, 4)])
def test_cornell_image_matches_oracle(oracle, kind, depth, spp):
    """Matched samples: tolerance 1e-3 mean relative error (north star); typically ~1e-6, the rest
    is libm vs CUDA sin/cos/log in a handful of paths."""
    sc = scenes.cornell_box(64, 64)
    for m in sc.materials:
        if kind == "diffuse" and m.kind == MAT_GLOSSY:
            m.kind = MAT_DIFFUSE
        if kind == "glossy" and m.kind == MAT_DIFFUSE:
            m.kind, m.roughness = MAT_GLOSSY, 0.5
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), spp, 1, depth, seed=11)
    got, launches = render_gpu(sc, acc, spp, depth, 11)
    assert launches > 4 * depth
    assert np.isfinite(got).all()
    assert mean_rel_err(got, want) < 1e-3
    assert np.all(got[..., 3] == 1.0)


@pytest.mark.parametrize("depth,spp", [(2, 8), (6, 16)])
def test_wider_closure_set_matches_oracle(oracle, depth, spp):
    """Oren-Nayar, mirror reflection (SPECULAR bounces count emission), sharp refraction, sheen, closure mixes
    and the constant environment (scenes.cornell_lobes): the device image against the restatement (pinned to
    the compiled reference's bsdf_t value by value) at matched samples."""
    sc = scenes.cornell_lobes(64, 64)
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), spp, 1, depth, seed=13)
    got, _ = render_gpu(sc, acc, spp, depth, 13)
    assert np.isfinite(got).all() and np.isfinite(want).all()
    assert mean_rel_err(got, want) < 1e-3
    dark = scenes.cornell_lobes(64, 64, environment=False)
    got_dark, _ = render_gpu(dark, acc, spp, depth, 13)
    assert got[..., :3].mean() > got_dark[..., :3].mean() * 1.02  # the environment is seen through the open front


@pytest.mark.parametrize("lobe", ["oren_nayar", "mirror", "glass", "rough_glass", "sheen", "transparent"])
def test_single_closures_match_oracle(oracle, lobe):
    from phosphorus_mk2_b200.scene import LOBE_MICROFACET_REFRACT, LOBE_REFRACTION, LOBE_SHEEN, LOBE_TRANSPARENT, MAT_LAYERED, Material
    sc = scenes.cornell_box(48, 48)
    new = {"oren_nayar": Material(MAT_DIFFUSE, (0.7, 0.6, 0.5), roughness=30.0),
           "mirror": Material(MAT_GLOSSY, (0.9, 0.9, 0.9), roughness=0.0),
           "glass": Material(MAT_LAYERED, lobes=((LOBE_REFRACTION, (0.95, 0.95, 0.95), 1.45),)),
           "rough_glass": Material(MAT_LAYERED, lobes=((LOBE_MICROFACET_REFRACT, (0.95, 0.95, 0.95), 0.3, 1.45),)),
           "sheen": Material(MAT_LAYERED, lobes=((LOBE_SHEEN, (0.8, 0.7, 0.9), 0.5),)),
           "transparent": Material(MAT_LAYERED, lobes=((LOBE_TRANSPARENT, (0.8, 0.9, 0.8), 0.0),))}[lobe]
    sc.materials[3] = new  # the tall box
    if lobe in ("oren_nayar", "sheen"):
        sc.materials[0] = new  # and the white walls
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 8, 1, 5, seed=4)
    got, _ = render_gpu(sc, acc, 8, 5, 4)
    assert np.isfinite(got).all()
    assert mean_rel_err(got, want) < 1e-3


def test_thin_lens_camera(oracle):
    """camera_t with aperture_radius != 0 (kernels/cpu/camera.hpp:140-147): primary rays against the oracle's
    restatement (itself pinned bit for bit to the compiled reference kernel) on the renderer's own lens draws,
    then the depth-of-field image at matched samples."""
    sc = scenes.cornell_box(64, 64)
    sc.camera.aperture_radius = 0.08
    sc.camera.focal_distance = 3.3
    acc = Accel(sc)
    dev = CudaDevice.make(Options(4, 1, 4), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    W = H = 64
    dr = dev.device_rays(W * H)
    seed, sample = 11, 2
    dev.camera_rays([(0, 0, W, H)], dr, 0.25, 0.75, seed=seed, sample=sample)
    got = dr.download()
    dr.free()
    dev.close()
    s32 = (seed ^ (seed >> 32)) & 0xFFFFFFFF
    pix = np.arange(W * H)
    lu = np.array([oracle.lib.orc_rnd(s32, int(p), sample, 0, 6) for p in pix], np.float32)
    lv = np.array([oracle.lib.orc_rnd(s32, int(p), sample, 1, 6) for p in pix], np.float32)
    wp, ww = oracle.camera_rays_lens(sc, 0, 0, W, H, 0.25, 0.75, lu, lv)
    gp = np.stack([got.px, got.py, got.pz], 1)
    gw = np.stack([got.wx, got.wy, got.wz], 1)
    assert np.abs(gp - wp).max() < 2e-6 and np.abs(gw - ww).max() < 2e-6  # CUDA vs glibc sinf / cosf
    assert np.abs(gp - gp[0]).max() > 1e-2                                # origins spread over the lens
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 8, 1, 4, seed=21)
    img, _ = render_gpu(sc, acc, 8, 4, 21)
    assert mean_rel_err(img, want) < 1e-3
    sc.camera.aperture_radius = 0.0
    pin, _ = render_gpu(sc, acc, 8, 4, 21)
    assert mean_rel_err(img, pin) > 1e-2  # and it is not the pinhole image


def test_normals_channel(oracle):
    """render_buffer_t::NORMALS: the shading normal of the last sample whose primary ray hit, against the oracle
    (same film jitter, so bit-level agreement up to CUDA vs host normalisation), whole frame, several sample
    batches, flat and smooth meshes, partial tiles."""
    from phosphorus_mk2_b200.scene import MAT_EMITTER, Material, Mesh
    spheres = scenes.sphere_field(2, 24, 12, 70, 50, smooth=True)
    light = spheres.add_material(Material(MAT_EMITTER, (1.0, 0.9, 0.8), power=30.0))
    v = np.array([[-1.5, 2.0, -1.5], [1.5, 2.0, -1.5], [1.5, 2.0, 1.5], [-1.5, 2.0, 1.5]], np.float32)
    nrm = np.tile(np.array([[0, -1, 0]], np.float32), (4, 1))
    spheres.add(Mesh(v, np.array([[0, 1, 2], [0, 2, 3]]), [(light, np.arange(2))], smooth=False, normals=nrm))
    for sc in (scenes.cornell_box(64, 64), spheres):
        acc = Accel(sc)
        cam = sc.camera
        want = np.zeros((cam.film_height, cam.film_width, 3), np.float32)
        oracle.render(sc, acc.nodes_array(), acc.packets_array(), 6, 1, 2, seed=9, normals=want)
        dev = CudaDevice.make(Options(6, 1, 2), 0)
        dev.preprocess(sc, acc)
        dev.upload_scene(sc)
        dev.enable_normals()
        tiles = make_tiles(cam.film_width, cam.film_height)
        for (a, b) in ((0, 2), (2, 3), (3, 6)):
            dev.render(tiles, a, b, 6, 9)
        got = dev.film_read_normals()
        dev.close()
        assert np.abs(got - want).max() < 1e-5
        hit = np.abs(want).sum(axis=2) > 0
        assert 0.3 < hit.mean() and np.array_equal(hit, np.abs(got).sum(axis=2) > 0)


def test_partial_tiles_and_smooth_normals(oracle):
    """Film 70 x 50 (partial tiles right and bottom), smooth-shaded spheres + an emissive quad."""
    sc = scenes.sphere_field(2, 24, 12, 70, 50, smooth=True)
    from phosphorus_mk2_b200.scene import MAT_EMITTER, Material, Mesh
    light = sc.add_material(Material(MAT_EMITTER, (1.0, 0.9, 0.8), power=30.0))
    v = np.array([[-1.5, 2.0, -1.5], [1.5, 2.0, -1.5], [1.5, 2.0, 1.5], [-1.5, 2.0, 1.5]], np.float32)
    nrm = np.tile(np.array([[0, -1, 0]], np.float32), (4, 1))
    sc.add(Mesh(v, np.array([[0, 1, 2], [0, 2, 3]]), [(light, np.arange(2))], smooth=False, normals=nrm))
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 4, 2, 4, seed=5)
    got, _ = render_gpu(sc, acc, 4, 4, 5, pps=2)
    assert mean_rel_err(got, want) < 1e-3


def test_tile_and_sample_partitioning_do_not_change_the_image():
    """The film is the same whether the frame is rendered in one call, tile by tile in scrambled
    order, or as separate sample ranges (the two multi-GPU partitionings)."""
    sc = scenes.cornell_box(96, 64)
    acc = Accel(sc)
    whole, _ = render_gpu(sc, acc, 16, 5, 3)
    tiles = make_tiles(96, 64)
    by_samples, _ = render_gpu(sc, acc, 16, 5, 3, ranges=[(0, 5), (5, 6), (6, 16)])
    assert np.array_equal(whole, by_samples)
    dev = CudaDevice.make(Options(16, 1, 5), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    for t in reversed(tiles):
        dev.render([t], 0, 16, 16, 3)
    by_tiles = dev.film_read()
    dev.close()
    assert np.array_equal(whole, by_tiles)


def test_batching_and_second_wavefront_do_not_change_the_image(monkeypatch):
    """The frame is cut into batches of PHOS_WAVEFRONT_PATHS paths that go round-robin over 1..4 wavefronts on as many
    streams (PHOS_WAVEFRONTS, default 2; film accumulation chained in sample order): same film, bit for bit."""
    sc = scenes.cornell_box(96, 64)
    acc = Accel(sc)
    monkeypatch.setenv("PHOS_WAVEFRONTS", "1")
    whole, n_whole = render_gpu(sc, acc, 16, 5, 3)
    monkeypatch.setenv("PHOS_WAVEFRONT_PATHS", "65536")  # the smallest allowed: 10 samples per batch here, then a short one
    batched, n_batched = render_gpu(sc, acc, 16, 5, 3)
    assert np.array_equal(whole, batched) and n_batched > n_whole
    monkeypatch.setenv("PHOS_WAVEFRONTS", "2")
    two, _ = render_gpu(sc, acc, 16, 5, 3)
    assert np.array_equal(whole, two)
    monkeypatch.delenv("PHOS_WAVEFRONT_PATHS")
    two_big, _ = render_gpu(sc, acc, 16, 5, 3)  # one batch per wavefront
    assert np.array_equal(whole, two_big)
    for nw in ("3", "4"):
        monkeypatch.setenv("PHOS_WAVEFRONTS", nw)
        monkeypatch.setenv("PHOS_WAVEFRONT_PATHS", "65536")
        many, _ = render_gpu(sc, acc, 16, 5, 3)
        assert np.array_equal(whole, many), nw
        monkeypatch.delenv("PHOS_WAVEFRONT_PATHS")
        assert np.array_equal(whole, render_gpu(sc, acc, 16, 5, 3)[0]), nw
    monkeypatch.delenv("PHOS_WAVEFRONTS")
    assert np.array_equal(whole, render_gpu(sc, acc, 16, 5, 3)[0])  # the default
    assert np.array_equal(render_gpu(sc, acc, 1, 5, 3)[0], render_gpu(sc, acc, 1, 5, 3)[0])  # fewer samples than wavefronts


def test_render_call_order_errors():
    from phosphorus_mk2_b200.lib import PhosError
    dev = CudaDevice.make(Options(), 0)
    with pytest.raises(PhosError):
        dev.render([(0, 0, 8, 8)], 0, 1, 1)
    sc = scenes.cornell_box(32, 32)
    dev.upload_scene(sc)
    with pytest.raises(PhosError):
        dev.render([(0, 0, 8, 8)], 0, 1, 1)  # no acceleration structure yet
    dev.preprocess(sc)
    with pytest.raises(PhosError):
        dev.render([(30, 30, 8, 8)], 0, 1, 1)  # tile outside the film
    with pytest.raises(PhosError):
        dev.render([(0, 0, 8, 8)], 2, 1, 4)  # bad sample range
    dev.close()


def test_start_join_feeds_every_tile_to_the_film_sink():
    """xpu_t::start / join over the shared tile queue and the film_t<>::add_tile hand-off."""
    from phosphorus_mk2_b200.frame import FrameState, MemoryFilm, Tiles, join, start
    sc = scenes.cornell_box(96, 80)
    acc = Accel(sc)
    direct, _ = render_gpu(sc, acc, 4, 4, 9)
    dev = CudaDevice.make(Options(4, 1, 4), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    tiles = Tiles.make(96, 80)
    film = MemoryFilm(96, 80)
    h = start(dev, sc, FrameState(tiles, film, spp=4, seed=9), chunk_tiles=4)
    join(h)
    dev.close()
    assert film.tiles_added == tiles.size
    assert np.array_equal(film.rgba, direct)


def test_two_contexts_emulate_two_ranks():
    """The multi-GPU frame on one GPU: two contexts each render their share (tile-partitioned, then
    sample-partitioned); the sum of the two films — what the NCCL reduce computes — is the frame."""
    from phosphorus_mk2_b200.frame import samples_of_rank, tiles_of_rank
    sc = scenes.cornell_box(96, 96)
    acc = Accel(sc)
    whole, _ = render_gpu(sc, acc, 16, 5, 21)
    tiles = make_tiles(96, 96)
    for mode in ("tiles", "samples"):
        films = []
        for r in range(2):
            if mode == "tiles":
                img, _ = render_gpu(sc, acc, 16, 5, 21, tiles=tiles_of_rank(tiles, r, 2))
            else:
                img, _ = render_gpu(sc, acc, 16, 5, 21, ranges=[samples_of_rank(16, r, 2)])
            films.append(img)
        total = films[0] + films[1]
        if mode == "tiles":
            assert np.array_equal(total[..., :3], whole[..., :3])  # disjoint tiles: the sum is a gather
        else:
            assert np.allclose(total[..., :3], whole[..., :3], rtol=2e-6, atol=1e-7)  # (a + b) + (c + d) vs a + b + c + d


def test_per_face_smooth_flag(oracle):
    """Meshes mixing smooth and flat faces (phos_scene_desc::face_smooth, mesh_smooth = 2): image and NORMALS channel
    against the oracle."""
    sc = scenes.mixed_shading_spheres(70, 50)
    acc = Accel(sc)
    want_n = np.zeros((50, 70, 3), np.float32)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 4, 1, 4, seed=5, normals=want_n)
    dev = CudaDevice.make(Options(4, 1, 4), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    dev.enable_normals()
    dev.render(make_tiles(70, 50), 0, 4, 4, 5)
    got, got_n = dev.film_read(), dev.film_read_normals()
    dev.close()
    assert np.abs(got_n - want_n).max() < 1e-5
    assert mean_rel_err(got, want) < 1e-3


def test_film_reduce_inside_the_library_single_rank():
    """phos_cuda_comm_* / phos_cuda_film_reduce with a one-rank communicator: the reduce is the identity on the colours
    and clamps alpha (the N-rank case runs under torchrun in test_gpu_multi.py when the box has >= 2 GPUs)."""
    import ctypes as C
    sc = scenes.cornell_box(64, 64)
    acc = Accel(sc)
    dev = CudaDevice.make(Options(4, 1, 3), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    ident = (C.c_uint8 * 128)()
    dev._check(dev._L.phos_cuda_comm_unique_id(ident))
    dev._check(dev._L.phos_cuda_comm_init(dev._ctx, 1, 0, ident))
    dev.render(make_tiles(64, 64), 0, 4, 4, 7)
    before = dev.film_read()
    dev.film_reduce(0)
    after = dev.film_read()
    with pytest.raises(PhosError):
        dev.film_reduce(1)  # root outside the communicator
    dev.close()
    assert np.array_equal(before, after) and (after[..., 3] == 1.0).all()


def test_pinhole_camera_rays_are_bit_identical_to_the_oracle(oracle):
    """phos_cuda_camera_rays, pinhole branch (camera.hpp:113-152), tile by tile in the reference's 32 x 32 order with
    partial tiles, a jitter other than the centre, and a rotated / translated camera: every origin and direction bit
    for bit the oracle's restatement (exact 1 / sqrt on both sides; no transcendental function is involved)."""
    from phosphorus_mk2_b200.scene import Camera
    sc = scenes.sphere_field(2, 16, 8, 150, 70)
    sc.camera.to_world = Camera.look_at((1.3, 2.1, 3.7), (0.2, -0.1, 0.4))
    acc = Accel(sc)
    dev = CudaDevice.make(Options(), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    tiles = make_tiles(150, 70)
    dr = dev.device_rays(150 * 70)
    for jx, jy in ((0.5, 0.5), (0.125, 0.875)):
        dev.camera_rays(tiles, dr, jx, jy)
        got = dr.download()
        o = 0
        for (x, y, w, h) in tiles:
            want = oracle.camera_rays(sc.camera, x, y, w, h, jx, jy)
            for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "flags"):
                assert np.array_equal(getattr(got, f)[o:o + w * h].view(np.uint32), getattr(want, f).view(np.uint32)), (f, x, y)
            o += w * h
    dr.free()
    dev.close()


def test_reference_normalize_reproduces_the_rcpps_images(oracle):
    """phos_cuda_reference_normalize: camera and shadow-ray directions normalised through the host's RCPSS (sampled into a
    table, rcp_table.cpp) — the image of the oracle's rcp_mode 2 (RCPSS normalisation, exact traversal) at matched
    samples, and visibly darker than exact arithmetic (the reference's shadow rays overshoot into the light)."""
    sc = scenes.cornell_box(64, 64)
    acc = Accel(sc)
    nodes, packets = acc.nodes_array(), acc.packets_array()
    dev = CudaDevice.make(Options(16, 1, 4), 0)
    dev.reference_normalize(True)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    dr = dev.device_rays(64 * 64)
    dev.camera_rays([(0, 0, 64, 64)], dr, 0.5, 0.5)
    rays = dr.download()
    dr.free()
    wp, ww = oracle.camera_rays_lens(sc, 0, 0, 64, 64, 0.5, 0.5, np.full(4096, 0.5, np.float32), np.full(4096, 0.5, np.float32), rcp_mode=True)
    assert np.array_equal(np.stack([rays.wx, rays.wy, rays.wz], 1).view(np.uint32), ww.view(np.uint32))  # RCPSS bit for bit
    dev.render(make_tiles(64, 64), 0, 16, 16, 5)
    got = dev.film_read()
    dev.reference_normalize(False)
    dev.film_clear()
    dev.render(make_tiles(64, 64), 0, 16, 16, 5)
    exact = dev.film_read()
    dev.close()
    want = oracle.render(sc, nodes, packets, 16, 1, 4, seed=5, rcp_mode=2)
    assert mean_rel_err(got, want) < 1e-3
    assert mean_rel_err(exact, oracle.render(sc, nodes, packets, 16, 1, 4, seed=5)) < 1e-3
    assert np.median(exact[..., :3]) > 1.05 * np.median(got[..., :3])


def test_reduced_config4_scene_matches_oracle(oracle):
    """BASELINE config 4's scene family at a size the oracle renders in seconds: displaced terrain with 10 % GGX face
    bands (roughness 0.2) under the sky emitter, depth 8."""
    sc = scenes.terrain(n=129, glossy_fraction=0.1, width=96, height=64)
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 8, 1, 8, seed=13)
    got, _ = render_gpu(sc, acc, 8, 8, 13)
    assert np.isfinite(got).all() and want[..., :3].mean() > 1e-3
    assert mean_rel_err(got, want) < 1e-3


def test_reduced_config5_scene_matches_oracle(oracle):
    """BASELINE config 5's scene family: a field of baked copies of one object (one mesh_t each: 9 meshes + the light),
    sample-partitioned over two calls like two ranks."""
    sc = scenes.instanced_field(grid=3, width=96, height=64)
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 8, 1, 6, seed=17)
    got, _ = render_gpu(sc, acc, 8, 6, 17, ranges=[(0, 4), (4, 8)])
    assert np.isfinite(got).all() and want[..., :3].mean() > 1e-3
    assert mean_rel_err(got, want) < 1e-3


def test_malformed_scene_is_refused_on_the_host():
    """A face whose vertex index lies outside its mesh must fail with PHOS_ERR_INVALID in upload_scene / build_accel: on the
    device it would be an out-of-bounds read, and a device fault is sticky (it kills the context)."""
    sc = scenes.cornell_box(32, 32)
    sc.meshes[1].faces[3, 1] = 10_000
    dev = CudaDevice.make(Options(1, 1, 2), 0)
    with pytest.raises(PhosError, match="vertex index"):
        dev.upload_scene(sc)
    with pytest.raises(PhosError, match="vertex index"):
        dev.build_accel(sc)
    ok = scenes.cornell_box(32, 32)  # the context is still usable
    acc = Accel(ok)
    dev.preprocess(ok, acc)
    dev.upload_scene(ok)
    dev.render(make_tiles(32, 32), 0, 1, 1, 3)
    assert np.isfinite(dev.film_read()).all()
    dev.close()


def test_path_depth_zero_still_traces_the_primary_rays(oracle):
    """path_depth 0: the reference's loop tests the depth AFTER the first bounce (spt.hpp:307-328), so the frame shows what the
    primary rays see — emitters and the environment — not black."""
    sc = scenes.cornell_lobes(48, 48)
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 4, 1, 0, seed=2)
    got, _ = render_gpu(sc, acc, 4, 0, 2)
    assert want[..., :3].max() > 0.1
    assert mean_rel_err(got, want) < 1e-3


def test_more_ranks_than_samples():
    """samples_of_rank hands some ranks an empty range when spp < world: render_partition renders nothing there and the sum of
    the films is still the frame."""
    from phosphorus_mk2_b200.frame import render_partition, samples_of_rank
    sc = scenes.cornell_box(64, 64)
    acc = Accel(sc)
    whole, _ = render_gpu(sc, acc, 2, 4, 8)
    tiles = make_tiles(64, 64)
    total = np.zeros_like(whole)
    ranges = [samples_of_rank(2, r, 5) for r in range(5)]
    assert any(a == b for a, b in ranges)
    dev = CudaDevice.make(Options(2, 1, 4), 0)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    for rng in ranges:
        render_partition(dev, tiles, rng, 2, seed=8)
        total += dev.film_read()
    dev.close()
    assert np.allclose(total[..., :3], whole[..., :3], rtol=2e-6, atol=1e-7)
