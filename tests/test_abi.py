"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

from phosphorus_mk2_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in ("phos_cuda.h", "phos_scene.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(phos_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported"


def test_python_binding_covers_the_header():
    assert sorted(lib.SYMBOLS) == declared_functions()


def test_no_device_means_no_context_and_a_clear_error():
    """Without a GPU the library must refuse to create a context (there is no CPU fallback)."""
    L = lib.load()
    if L.phos_cuda_device_count() > 0:
        return  # on a GPU box this is covered by the gpu tests
    ctx = L.phos_cuda_create(0, None)
    assert not ctx
    assert b"no CUDA device" in L.phos_cuda_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phosphorus_mk2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("test oracle", "") or f == "rays.py", (dirpath, f)
