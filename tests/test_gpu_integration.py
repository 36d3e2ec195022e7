"""The drop-in boundary exercised by the reference's OWN host code: oracle/_ref contains the reference
sources compiled together with integration/cuda.{hpp,cpp} (this repository's replacement for the empty
src/xpu/cuda.* stub).  ref_render_on(use_cuda = 1) builds the scene with the reference's mesh / scene
builders, then makes exactly the calls session_t::details_t::render makes on an xpu_t —
make, preprocess, start, join — with the reference's tiles_t, sampler_t and a film_t<> sink.
Needs a B200 and oracle/_ref (built where /root/reference exists; the .so travels)."""
import numpy as np
import pytest

from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles

pytestmark = pytest.mark.gpu


def test_reference_host_code_drives_the_cuda_device(reflib, oracle):
    assert reflib.lib.ref_cuda_device_count() >= 1
    sc = scenes.cornell_box(96, 64)
    rs = reflib.scene(sc)
    got, secs = rs.render_cuda(spp=16, pps=1, depth=5)
    assert secs > 0 and np.isfinite(got).all()
    # the same frame through the Python mirror of the boundary: identical film (same library, same seed)
    dev = CudaDevice.make(Options(16, 1, 5), 0)
    acc = Accel(sc)
    dev.preprocess(sc, acc)
    dev.upload_scene(sc)
    dev.render(make_tiles(96, 64), 0, 16, 16, 0)
    direct = dev.film_read()
    dev.close()
    assert np.array_equal(got[..., :3], direct[..., :3])
    # and against the integrator oracle at matched samples
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 16, 1, 5, seed=0)
    err = float(np.abs(got[..., :3] - want[..., :3]).mean() / np.abs(want[..., :3]).mean())
    assert err < 1e-3


def test_gpu_and_cpu_device_agree_statistically(reflib):
    """cuda_t next to cpu_t behind the same interface: the GPU image is the reference CPU image up to
    Monte-Carlo noise and the reference's documented shadow-ray overshoot (GPU brighter)."""
    sc = scenes.cornell_box(48, 48)
    gpu, _ = reflib.scene(sc).render_cuda(spp=256, pps=1, depth=4)
    cpu, _ = reflib.scene(sc).render(256, 1, 4, single_threaded=True)
    g, c = np.median(gpu[..., :3]), np.median(cpu[..., :3])
    assert 1.0 < g / c < 1.6


def test_normals_channel_and_thin_lens_through_the_reference_host_code(reflib, oracle):
    """The tile format asks for render_buffer_t::NORMALS and the camera has an aperture: cuda_t, driven by the
    reference's tiles_t / film_t, hands back the same NORMALS channel the reference's cpu_t produces (up to
    the different film jitter on silhouettes) and a depth-of-field image equal to the oracle's."""
    sc = scenes.cornell_box(64, 64)
    sc.camera.aperture_radius = 0.05
    sc.camera.focal_distance = 3.3
    rs = reflib.scene(sc)
    gn = np.zeros((64, 64, 3), np.float32)
    gimg, _ = rs.render(8, 1, 3, normals=gn, cuda=True)
    cn = np.zeros_like(gn)
    rs.render(8, 1, 3, normals=cn)
    assert (np.abs(gn - cn).max(axis=2) < 2e-3).mean() > 0.9
    acc = Accel(sc)
    want = oracle.render(sc, acc.nodes_array(), acc.packets_array(), 8, 1, 3, seed=0)
    err = float(np.abs(gimg[..., :3] - want[..., :3]).mean() / np.abs(want[..., :3]).mean())
    assert err < 1e-3
