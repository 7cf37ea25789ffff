"""The C-ABI boundary: libalq.so loads on a CPU-only box and exports exactly the entry points
include/alq.h declares; the ctypes binding mirrors them.  No compute calls here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "alq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(alq_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from active_learning_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in alq.h but not exported"


def test_ctypes_binding_mirrors_header():
    from active_learning_b200 import _lib
    assert set(_lib.SIGNATURES) == _header_functions()
    lib = _lib.load()
    assert lib.alq_version() == _lib.ABI_VERSION


def test_greedy_desc_layout_matches_header_field_order():
    from active_learning_b200 import _lib
    src = open(os.path.join(ROOT, "include", "alq.h")).read()
    body = src[src.index("typedef struct alq_greedy_desc {"):src.index("} alq_greedy_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(\w+)\s*;", body)
    assert fields == [f[0] for f in _lib.GreedyDesc._fields_]


def test_create_without_gpu_fails_cleanly():
    import torch
    if torch.cuda.is_available():
        return
    from active_learning_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.alq_create(ctypes.byref(h), 0) != 0 and not h.value
