import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from phosphorus_mk2_b200 import lib
        return lib.load().phos_cuda_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own compiled hot path (oracle/_ref); absent only if never built."""
    from oracle.pyoracle import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libphos_ref.so not built (needs /root/reference at build time)")
    return RefLib()


@pytest.fixture(scope="session")
def reflib_cuda():
    """The same reference objects + the drop-in GPU device (integration/cuda.cpp, integration/xpu_discover.patch)
    linked against libphos_cuda.so: oracle/_ref/libphos_ref_cuda.so."""
    from oracle.pyoracle import RefLib
    if not RefLib.available(cuda=True):
        pytest.skip("oracle/_ref/libphos_ref_cuda.so not built (needs /root/reference at build time)")
    return RefLib(cuda=True)


def _build_emul(defs):
    import ctypes as C
    from phosphorus_mk2_b200.rays import PhosRays
    here = os.path.join(ROOT, "tests", "emul")
    tag = "".join(c if c.isalnum() else "_" for c in "".join(defs))
    so = os.path.join(here, f"libemul_trace{tag}.so")
    srcs = [os.path.join(here, "emul_trace.cpp"), os.path.join(ROOT, "phosphorus_mk2_b200", "csrc", "repack.cpp")]
    deps = srcs + [os.path.join(ROOT, "phosphorus_mk2_b200", "csrc", f) for f in ("trace_ray.cuh", "phos_internal.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                        "-I/usr/local/cuda/include", *defs, *srcs, "-o", so], check=True)
    L = C.CDLL(so)
    L.emul_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(PhosRays), C.c_uint64,
                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_char_p]
    L.emul_repack_digest.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    return L


@pytest.fixture(scope="session")
def emul():
    """Host emulation of the device traversal source (tests/emul), built on demand.  PHOS_EMUL_DEFS (e.g.
    "-DPHOS_X=1") emulates a kernel build knob."""
    return _build_emul(os.environ.get("PHOS_EMUL_DEFS", "").split())


@pytest.fixture(scope="session")
def emul_rcp():
    """The same emulation with every reciprocal direction pushed one ulp off, alternately up and down: the worst the
    device's MUFU.RCP may return.  Results must not change (the box test is conservative by more than that)."""
    return _build_emul(["-DPHOS_EMUL_RCP_PERTURB"])


@pytest.fixture(scope="session")
def device():
    """One CudaDevice on cuda:0.  Raises (does not skip) when the library or the GPU is missing."""
    from phosphorus_mk2_b200.device import CudaDevice, Options
    dev = CudaDevice.make(Options(), 0)
    yield dev
    dev.close()
